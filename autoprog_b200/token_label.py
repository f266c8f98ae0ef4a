"""Token-label target builder -- drop-in for `tlt.data.create_token_label_target` as the reference calls it
(main_prog.py:983-1004, 1919-1932): `create_token_label_target(target, num_classes, smoothing, label_size)`.

`target` is what tlt's loader yields: int64 class ids [B] (-> smoothed one-hot [B, C]) or label maps [B, 3, 5, Hm, Wm]
(top-5 scores, top-5 class ids, augmentation record) -> [B, C, 2 + label_size^2] fp32, the class-major layout that
`TokenLabelCrossEntropy` consumes.  One CUDA kernel (csrc/token_label.cu); CUDA only, no eager fallback.  tlt is not
vendored in the reference tree: the recipe is restated from the published TokenLabeling code (oracle/token_label_cpu.py,
SURVEY.md Appendix B) and its parity is unpinned upstream.
"""
import torch

from ._lib import check, lib


def create_token_label_target(target: torch.Tensor, num_classes: int, smoothing: float = 0.1, label_size: int = 1,
                              apply_softmax: bool = True) -> torch.Tensor:
    if not target.is_cuda:
        raise RuntimeError('autoprog_b200.create_token_label_target needs a CUDA tensor (there is no CPU fallback)')
    st = torch.cuda.current_stream().cuda_stream
    if target.dim() == 1:
        labels = target.to(torch.int64).contiguous()
        out = torch.empty((labels.shape[0], num_classes), device=target.device, dtype=torch.float32)
        check(lib().apb_onehot_smooth(labels.data_ptr(), out.data_ptr(), labels.shape[0], num_classes, float(smoothing), st),
              'onehot_smooth')
        return out
    if target.dim() != 5 or target.shape[1] != 3 or target.shape[2] != 5:
        raise ValueError(f'label maps must be [B, 3, 5, Hm, Wm], got {tuple(target.shape)}')
    maps = target.to(torch.float32).contiguous()
    B, _, _, Hm, Wm = maps.shape
    out = torch.empty((B, num_classes, 2 + label_size * label_size), device=target.device, dtype=torch.float32)
    check(lib().apb_token_label_target(maps.data_ptr(), out.data_ptr(), B, num_classes, Hm, Wm, int(label_size), float(smoothing),
                                       int(apply_softmax), st), 'token_label_target')
    return out

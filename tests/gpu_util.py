import os

import pytest
import torch

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
F32_TOL = 1e-5     # north-star fp32 tolerance (relative, norm-wise)
BF16_TOL = 2e-2    # north-star bf16 tolerance against the fp32/fp64 reference output


def need_gpu():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return torch.device('cuda:0')


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-300))


def tol(dtype):
    return F32_TOL if dtype == torch.float32 else BF16_TOL

"""2-GPU debug: which parameters diverge between replicas in CUDA-graph DDP mode."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
os.environ.setdefault('TORCH_NCCL_ASYNC_ERROR_HANDLING', '0')
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank); dev = torch.device('cuda', rank)
dist.init_process_group('nccl', device_id=dev)
import autoprog_b200 as A
from autoprog_b200.ddp import DistributedDataParallel
from autoprog_b200.graph import GraphedTrainStep
from autoprog_b200.optim import FusedAdamW

def report(tag, m, opt):
    worst = []
    for n, p in m.named_parameters():
        both = [torch.empty_like(p.data) for _ in range(world)]
        dist.all_gather(both, p.data.contiguous())
        d = float((both[0] - both[1]).abs().max())
        if d > 0: worst.append((d, n))
    gw = []
    for n, p in m.named_parameters():
        if p.grad is None: continue
        both = [torch.empty_like(p.grad) for _ in range(world)]
        dist.all_gather(both, p.grad.contiguous())
        d = float((both[0] - both[1]).abs().max())
        if d > 0: gw.append((d, n))
    if rank == 0:
        print(tag, 'params differing:', len(worst), sorted(worst, reverse=True)[:6], '| grads differing:', len(gw), sorted(gw, reverse=True)[:6], flush=True)

def run(mode):
    torch.manual_seed(100 + rank)
    m = A.create_model('model_variant', variant='volo_h2_l4', img_size=64, num_classes=16).to(dev)
    opt = FusedAdamW(m, lr=1e-3, weight_decay=0.05)
    net = DistributedDataParallel(m, flat=opt.flat, bucket_mb=0.05)
    crit = A.TokenLabelCrossEntropy(dense_weight=0.5)
    torch.manual_seed(7 + rank)
    x = torch.randn(4, 3, 64, 64, device=dev); tgt = torch.softmax(torch.randn(4, 16, 18, device=dev), 1)
    report(mode + ' init', m, opt)
    np.random.seed(5)
    if mode == 'reduce_now':
        for i in range(3):
            opt.zero_grad()
            with net.no_sync():
                with A.autocast():
                    crit(net(x), tgt).backward()
            net.reduce_now()
            report(f'reduce_now step {i} after reduce', m, opt)
            opt.step()
            report(f'reduce_now step {i} after opt', m, opt)
    elif mode == 'eager':
        for i in range(3):
            opt.zero_grad()
            with A.autocast():
                crit(net(x), tgt).backward()
            report(f'eager step {i} after bwd', m, opt)
            opt.step()
            report(f'eager step {i} after opt', m, opt)
    else:
        step = GraphedTrainStep(net, crit, opt, x, tgt, bf16=True, warmup=int(os.environ.get('WARM', '3')))
        report('after warmup+capture', m, opt)
        for i in range(3):
            step()
            report(f'graph step {i}', m, opt)
        step.close()

for mode in sys.argv[1:]:
    run(mode)
dist.destroy_process_group()

"""world_size-2 gloo test (CPU) of the bucketed gradient reducer: N-rank averaged gradients == single process over the
concatenated batch; parameters without gradient (elastic depth) stay consistent; no_sync accumulates locally."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.a = nn.Linear(6, 16)
        self.n = nn.LayerNorm(16)
        self.skip = nn.Linear(16, 16)      # never used in forward: "identity layer" without gradient
        self.b = nn.Linear(16, 3)
        self.pos = nn.Parameter(torch.zeros(1, 16))

    def no_weight_decay(self):
        return {'pos'}

    def forward(self, x):
        return self.b(self.n(torch.tanh(self.a(x)) + self.pos))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from autoprog_b200.ddp import DistributedDataParallel
    from autoprog_b200.flat import FlatState
    torch.manual_seed(100 + rank)                 # different init per rank: the wrapper must broadcast rank 0's
    net = Net()
    flat = FlatState(net, weight_decay=0.05, want_shadow=False)
    ddp = DistributedDataParallel(net, flat=flat, bucket_mb=0.0005)   # tiny buckets -> several buckets
    assert len(ddp.buckets) >= 3
    torch.manual_seed(7)
    x = torch.randn(8, 6)
    y = torch.randn(8, 3)
    xs, ys = x[rank * 4:(rank + 1) * 4], y[rank * 4:(rank + 1) * 4]
    flat.zero_grad()
    ((ddp(xs) - ys) ** 2).mean().backward()
    res = {n: p.grad.clone() for n, p in net.named_parameters()}
    res['__w'] = {n: p.detach().clone() for n, p in net.named_parameters()}
    # accumulate without sync, then reduce explicitly
    flat.zero_grad()
    with ddp.no_sync():
        ((ddp(xs) - ys) ** 2).mean().backward()
    res['__local'] = net.b.weight.grad.clone()
    ddp.reduce_now()
    res['__reduced'] = net.b.weight.grad.clone()
    torch.save(res, out + f'.{rank}')
    dist.destroy_process_group()


def test_ddp_two_ranks_gloo(tmp_path):
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / 'res')
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r0, r1 = torch.load(out + '.0'), torch.load(out + '.1')
    # reference: rank-0 weights, full batch
    sys.path.insert(0, ROOT)
    net = Net()
    net.load_state_dict(r0['__w'])
    torch.manual_seed(7)
    x = torch.randn(8, 6)
    y = torch.randn(8, 3)
    ((net(x) - y) ** 2).mean().backward()
    for n, p in net.named_parameters():
        assert torch.equal(r0['__w'][n], r1['__w'][n]), n           # broadcast made the replicas identical
        if p.grad is None:
            assert float(r0[n].abs().max()) == 0.0 and float(r1[n].abs().max()) == 0.0, n
            continue
        assert torch.allclose(r0[n], p.grad, atol=1e-6), n
        assert torch.equal(r0[n], r1[n]), n
    assert not torch.allclose(r0['__local'], r1['__local'])
    assert torch.allclose(r0['__reduced'], net.b.weight.grad, atol=1e-6) and torch.equal(r0['__reduced'], r1['__reduced'])

#!/usr/bin/env python
"""Benchmark of the AutoProg hot path: a training step (fwd + TokenLabelCrossEntropy + bwd + all-reduce + AdamW/EMA step).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config NAME]

Prints ONE JSON line (rank 0).  `value` = whole-job train images/sec with inputs resident in HBM; `e2e` = the same
through the public API with pinned HOST buffers (H2D of images + token labels and D2H of the loss every step);
`roofline` = the dominant kernel family (tcgen05 GEMM) measured in situ with CUDA events; `cpu_baseline` = the oracle
port of the reference's PyTorch path on this box's host cores; `gpu_eager_baseline` = the same restatement of the
reference (plain torch ops: cuBLAS / cuDNN / ATen unfold-softmax-fold) on THIS GPU under bf16 autocast.
`--impl reference` times the CPU path alone.

--config (BASELINE.json configs; the headline stays the default):
  d1_224      volo_d1 == volo_h12_l18, 224 px, per-GPU batch 128 (configs[1], final stage)           [default]
  stages      the four AutoProg stages (l9@128, l12@160, l15@192, l18@224) + a super-net epoch with a random (r, l) per step
  d2_384      volo_d2 at 384 px, per-GPU batch 64 (configs[3])
  deit_small  deit_small progressive (elastic depth 6..12, r 128..224) with token labeling (configs[2]; extension, see DESIGN)
  micro       OutlookAttention core + TokenLabelCrossEntropy sweep r in 112..448 x B in 64..512 against the reference's
              PyTorch ops on the same GPU (configs[4])
"""
from __future__ import annotations

import argparse
import copy
import hashlib
import json
import os
import random
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# stdout carries exactly ONE line, the JSON result: fd 1 is pointed at stderr for the whole run (libraries such as NCCL print
# banners on stdout) and the saved descriptor is used only by emit()
_RESULT_FD = None


def claim_stdout():
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + '\n').encode()
    sys.stdout.flush()
    if _RESULT_FD is None:
        os.write(1, data)
    else:
        os.write(_RESULT_FD, data)

WORKLOAD = 'volo_d1 (== volo_h12_l18) AutoProg final stage: 224px, depth 18, per-GPU batch 128, bf16, TokenLabelCE(dense 0.5)'
EMA_DECAYS = [0.998, 0.9986, 0.999, 0.9996]            # scripts/train_autoprog.sh
STAGES = ((9, 128, 0.0), (12, 160, 0.1 / 3), (15, 192, 0.2 / 3), (18, 224, 0.1))   # progressive_schedule of the shipped script


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='d1_224', choices=['d1_224', 'stages', 'd2_384', 'deit_small', 'micro'])
    ap.add_argument('--batch', type=int, default=None, help='per-GPU batch (scripts/train_autoprog.sh: -b 128)')
    ap.add_argument('--model', default=None)
    ap.add_argument('--res', type=int, default=None)
    ap.add_argument('--no-ema', action='store_true')
    ap.add_argument('--cpu-steps', type=int, default=3)
    ap.add_argument('--skip-cpu', action='store_true')
    ap.add_argument('--skip-e2e', action='store_true')
    ap.add_argument('--skip-eager', action='store_true', help='skip the torch-eager GPU baseline of the reference path')
    ap.add_argument('--fp32', action='store_true', help='fp32 parity mode (CUDA-core kernels)')
    ap.add_argument('--no-graph', action='store_true', help='do not replay the step from a CUDA graph')
    ap.add_argument('--no-stages', action='store_true', help='skip the per-stage AutoProg schedule table (N=1 only)')
    ap.add_argument('--ddp-mode', default='overlap', choices=['overlap', 'split'],
                    help='N>1 graph path: all-reduce captured inside the graph and overlapped with backward, or split graphs')
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's PyTorch path (BASELINE.json configs[0]: batch 4, fp32, 224px)
# ----------------------------------------------------------------------------------------------------------------
def cpu_path_images_per_sec(steps: int, warmup: int, model_name: str, res: int, batch: int = 4):
    import numpy as np
    import torch
    from oracle import volo_cpu as O
    import autoprog_b200 as A

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    np.random.seed(0)
    arch = O.VoloArch.named(model_name, img_size=224)
    m = A.create_model(model_name, img_size=224)            # only used as a container of correctly shaped weights
    sd = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point) for k, v in m.state_dict().items()}
    x = torch.randn(batch, 3, res, res)
    g = res // 16
    tgt = torch.softmax(torch.randn(batch, 1000, 2 + g * g), dim=1)
    times = []
    for it in range(warmup + steps):
        for v in sd.values():
            v.grad = None
        t0 = time.perf_counter()
        lam = np.random.beta(1.0, 1.0)
        bbox = O.rand_bbox(g, g, lam)
        out = O.volo_forward(sd, x, arch, train=True, bbox=bbox)
        loss = O.token_label_ce(out[0], out[1], out[2], tgt, dense_weight=0.5)
        loss.backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    med = statistics.median(times)
    return batch / med, cores, f'{steps} steps of batch {batch} @ {res}px fp32 fwd+loss+bwd (median), {cores} torch threads'


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 6))
    v, cores, sample = cpu_path_images_per_sec(steps, min(args.warmup, 2), args.model or 'volo_d1', args.res or 224)
    line = {
        'impl': 'reference', 'metric': 'train images/sec', 'value': round(v, 3), 'unit': 'images/s', 'n_gpus': args.gpus,
        'steps': steps, 'warmup': min(args.warmup, 2), 'ms_per_step': round(4 / v * 1e3, 2), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'note': 'reference PyTorch path restated in oracle/volo_cpu.py, CPU, bounded sample batch 4'},
        'cpu_baseline': {'value': round(v, 3), 'unit': 'images/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': round(v, 3), 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    emit(line)


# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith('active')})
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(sm)}


def load_peaks():
    peaks = {'hbm_gbs': 6650.0, 'bf16_tflops_sustained': 1400.0, 'bf16_tflops': 1590.0, 'src': 'fallback'}
    try:
        pk = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        peaks = {'hbm_gbs': pk['hbm_gbs'], 'bf16_tflops_sustained': pk['bf16_tflops_sustained'],
                 'bf16_tflops': pk.get('bf16_tflops', pk['bf16_tflops_sustained']), 'src': 'measured'}
    except Exception:
        pass
    return peaks


def kernel_sources_sha():
    """Hash of the CUDA sources: a committed ncu traffic capture is only quoted when it was taken from THIS code."""
    h = hashlib.sha1()
    d = os.path.join(ROOT, 'autoprog_b200', 'csrc')
    for f in sorted(os.listdir(d)):
        with open(os.path.join(d, f), 'rb') as fh:
            h.update(f.encode())
            h.update(fh.read())
    return h.hexdigest()[:16]


def load_traffic():
    """profiles/dram_traffic.json (tools/capture_traffic.py, ncu --set full): DRAM bytes per launch of the headline
    kernels.  Returns {} -- and the roofline `traffic` fields become null -- when the capture predates the sources."""
    try:
        t = json.load(open(os.path.join(ROOT, 'profiles', 'dram_traffic.json')))
        if t.get('csrc_sha') != kernel_sources_sha():
            return {}
        return t['kernels']
    except Exception:
        return {}


# ----------------------------------------------------------------------------------------------------------------
class Env:
    """Process-wide state: device, rank, world, helpers for barriers / timed loops (max over ranks, on the device)."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get('RANK', '0'))
        self.local = int(os.environ.get('LOCAL_RANK', '0'))
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        torch.cuda.set_device(self.local)
        self.dev = torch.device('cuda', self.local)
        if self.world > 1:
            os.environ.setdefault('TORCH_NCCL_ASYNC_ERROR_HANDLING', '0')   # the watchdog must not poll a capturing stream
            dist.init_process_group('nccl', device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps):
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item())

    def all_agree(self, ok: bool) -> bool:
        if self.world == 1:
            return ok
        t = self.torch.tensor([1 if ok else 0], device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(int(t.item()))


class TrainJob:
    """One model + fused optimizer (+ EMAs, + DDP) + criterion with a graph-replayed (or eager) step."""

    def __init__(self, env: Env, model_name, model_kw, img_size, res, B, bf16=True, ema=True, graph=True, ddp_mode='overlap',
                 dense_tuple=False, drop_path=0.1, sample_config=None, lr=None):
        import torch
        import autoprog_b200 as A
        from autoprog_b200.optim import FusedAdamW
        from autoprog_b200.ddp import DistributedDataParallel
        self.env, self.B, self.res, self.bf16 = env, B, res, bf16
        dev = env.dev
        kw = dict(model_kw or {})
        kw.setdefault('drop_path_rate', drop_path)
        self.model = A.create_model(model_name, img_size=img_size, **kw).to(dev)
        if sample_config is not None:
            self.model.set_sample_config(sample_config)
        self.decays = EMA_DECAYS if ema else []
        self.emas = [copy.deepcopy(self.model).eval() for _ in self.decays]
        for e in self.emas:
            for p in e.parameters():
                p.requires_grad_(False)
        lr = lr if lr is not None else 1.6e-3 * (B * env.world) / 1024.0
        self.opt = FusedAdamW(self.model, lr=lr, weight_decay=0.05, ema_models=self.emas, ema_decays=self.decays)
        self.net = DistributedDataParallel(self.model, flat=self.opt.flat) if env.world > 1 else self.model
        tlce = A.TokenLabelCrossEntropy(dense_weight=0.5, cls_weight=1.0)
        if dense_tuple:      # DeiT with the token-labeling aux head returns (cls, dense): no mix-token box
            self.crit = lambda out, tgt: tlce((out[0], out[1], (0, 0, 0, 0)), tgt)
        else:
            self.crit = tlce
        g = res // 16
        self.g = g
        self.x = torch.randn(B, 3, res, res, device=dev)
        self.t = torch.softmax(torch.randn(B, 1000, 2 + g * g, device=dev), dim=1)
        self.graphed, self.note = None, 'eager launches'
        if graph:
            self._capture(sample_config, ddp_mode)

    def _capture(self, sample_config, ddp_mode):
        from autoprog_b200.graph import GraphedTrainStep
        env = self.env
        modes = ['single'] if env.world == 1 else ([ddp_mode, 'split'] if ddp_mode != 'split' else ['split'])
        for mode in modes:
            gs, err = None, ''
            try:
                gs = GraphedTrainStep(self.net, self.crit, self.opt, self.x, self.t, bf16=self.bf16, warmup=3,
                                      sample_config=sample_config, ddp_mode=mode if mode != 'single' else 'overlap')
            except Exception as e:   # noqa: BLE001 - fall back, say so in the JSON line
                err = f'{type(e).__name__}: {str(e)[:80]}'
                self.model._graph_box = None
                try:
                    self.env.torch.cuda.synchronize()
                except Exception:
                    pass
            if env.all_agree(gs is not None):        # all ranks must take the same path
                self.graphed = gs
                self.note = {'single': 'whole step replayed from one CUDA graph',
                             'overlap': 'whole step = one CUDA graph per rank; bucketed NCCL all-reduces captured inside it on a '
                                        'forked stream, overlapped with the remaining backward',
                             'split': 'two CUDA graphs per step (fwd+bwd | optimizer+EMA) around one eager bucketed NCCL all-reduce'}[mode]
                return
            if gs is not None:
                gs.close()
            self.note = f'eager launches (graph capture [{mode}] failed: {err or "on another rank"})'

    def eager_step(self, x, tgt):
        import autoprog_b200 as A
        self.opt.zero_grad()
        with A.autocast(enabled=self.bf16):
            out = self.net(x)
            loss = self.crit(out, tgt)
        loss.backward()
        self.opt.step()
        return loss

    def step(self, x=None, tgt=None):
        x = self.x if x is None else x
        tgt = self.t if tgt is None else tgt
        if self.graphed is not None:
            return self.graphed(x if x is not self.x else None, tgt if tgt is not self.t else None)
        return self.eager_step(x, tgt)

    def launches_per_step(self, K, measured):
        return self.graphed.kernels_per_step if self.graphed is not None else measured

    def close(self):
        if self.graphed is not None:
            self.graphed.close()
            self.graphed = None
        self.model._graph_box = None


def measure_device(env, job, steps, warmup, K):
    for _ in range(max(3, warmup)):
        job.step()
    sampler = ClockSampler(env.local)
    if env.rank == 0:
        sampler.start()
    l0 = K.launch_count()
    ms = env.timed(job.step, steps)
    launches = K.launch_count() - l0
    if job.graphed is not None:
        launches = job.graphed.kernels_per_step * steps      # replayed from the graph: counted at capture time
    clocks = sampler.stop() if env.rank == 0 else None
    return ms, launches, clocks


def measure_e2e(env, job, steps):
    """Pinned host buffers, H2D every step on a copy stream (double-buffered like tlt's PrefetchLoader), loss read back."""
    torch = env.torch
    B, res, g = job.B, job.res, job.g
    hx = [torch.randn(B, 3, res, res).pin_memory() for _ in range(2)]
    ht = [torch.softmax(torch.randn(B, 1000, 2 + g * g), dim=1).pin_memory() for _ in range(2)]
    dx = [torch.empty_like(job.x) for _ in range(2)]
    dtg = [torch.empty_like(job.t) for _ in range(2)]
    copy_stream = torch.cuda.Stream()
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    state = {'i': 0, 'loss': 0.0}

    def prefetch(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            dx[slot].copy_(hx[slot], non_blocking=True)
            dtg[slot].copy_(ht[slot], non_blocking=True)
            ready[slot].record(copy_stream)

    def e2e_step():
        slot = state['i'] & 1
        torch.cuda.current_stream().wait_event(ready[slot])
        prefetch(slot ^ 1)                      # next batch streams in while this step computes
        loss = job.step(dx[slot], dtg[slot])
        consumed[slot].record()
        state['loss'] = float(loss.item())      # D2H of the step's result
        state['i'] += 1

    for s in range(2):
        consumed[s].record()
    prefetch(0)
    for _ in range(3):
        e2e_step()
    ms_e2e = env.timed(e2e_step, steps)
    return {'value': round(B * env.world * steps / (ms_e2e / 1e3), 1), 'unit': 'images/s',
            'h2d_bytes_per_step': int(hx[0].numel() * 4 + ht[0].numel() * 4) * env.world, 'd2h_bytes_per_step': 4 * env.world,
            'ms_per_step': round(ms_e2e / steps, 3), 'last_loss': state['loss']}


def gpu_eager_baseline(env, model_name, img_size, res, B, steps=5, ema=True):
    """The reference path as torch eager executes it on THIS GPU (BASELINE.md §5 'beat this'): oracle/volo_cpu.py is the
    reference's nn.Module arithmetic restated functionally with stock torch ops (F.conv2d / F.linear / unfold-softmax-
    fold / F.layer_norm ...), here under torch.autocast(bf16) with torch.optim.AdamW(fused) + 4 foreach-lerp EMAs."""
    import numpy as np
    torch = env.torch
    from oracle import volo_cpu as O
    import autoprog_b200 as A
    dev = env.dev
    torch.manual_seed(0)
    arch = O.VoloArch.named(model_name, img_size=img_size)
    m = A.create_model(model_name, img_size=img_size)
    sd = {k: v.detach().to(dev).clone().requires_grad_(v.dtype.is_floating_point) for k, v in m.state_dict().items()}
    params = [v for v in sd.values() if v.requires_grad]
    opt = torch.optim.AdamW(params, lr=2e-4, weight_decay=0.05, fused=True)
    emas = [[p.detach().clone() for p in params] for _ in (EMA_DECAYS if ema else [])]
    g = res // 16
    x = torch.randn(B, 3, res, res, device=dev)
    tgt = torch.softmax(torch.randn(B, 1000, 2 + g * g, device=dev), dim=1)
    rates = O.drop_path_rates(arch, 0.1)

    def step():
        opt.zero_grad(set_to_none=True)
        lam = np.random.beta(1.0, 1.0)
        bbox = O.rand_bbox(g, g, lam)
        with torch.autocast('cuda', dtype=torch.bfloat16):
            out = O.volo_forward(sd, x, arch, train=True, bbox=bbox)
            loss = O.token_label_ce(out[0].float(), out[1].float(), out[2], tgt, dense_weight=0.5)
        loss.backward()
        opt.step()
        with torch.no_grad():
            for e, d in zip(emas, EMA_DECAYS):
                torch._foreach_lerp_(e, params, 1.0 - d)
        return loss

    del rates
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del sd, params, opt, emas
    torch.cuda.empty_cache()
    return {'value': round(B / ms * 1e3, 1), 'unit': 'images/s', 'ms_per_step': round(ms, 3),
            'what': f'oracle/volo_cpu.py (reference arithmetic, stock torch ops) on this GPU, torch.autocast(bf16), eager launches, '
                    f'torch.optim.AdamW(fused) + {len(EMA_DECAYS) if ema else 0} EMA lerps, batch {B} @ {res}px, no drop-path'}


# ----------------------------------------------------------------------------------------------------------------
def run_headline(env, args, spec):
    """One training configuration -> the JSON line of the contract."""
    import numpy as np
    torch = env.torch
    from autoprog_b200 import kernels as K
    torch.manual_seed(0)
    np.random.seed(0 + env.rank)
    B, res = spec['B'], spec['res']
    bf16 = not args.fp32
    job = TrainJob(env, spec['model'], spec.get('model_kw'), spec['img_size'], res, B, bf16=bf16, ema=not args.no_ema,
                   graph=not args.no_graph, ddp_mode=args.ddp_mode, dense_tuple=spec.get('dense_tuple', False))
    ms, launches, clocks = measure_device(env, job, args.steps, args.warmup, K)
    value = B * env.world * args.steps / (ms / 1e3)
    e2e = None if args.skip_e2e else measure_e2e(env, job, args.steps)

    # ---- rooflines in situ: one extra eager step with the kernel entry points bracketed by events (rank-local)
    roof, extra = None, {}
    job.close()
    if env.rank == 0:
        if env.world > 1:
            with job.net.no_sync():
                roof, extra = measure_rooflines(job.eager_step, job.x, job.t, K, torch, B, bf16)
        else:
            roof, extra = measure_rooflines(job.eager_step, job.x, job.t, K, torch, B, bf16)
        fb = K.fallback_count()
        extra['simt_fallbacks'] = fb          # bf16 launches that fell through to a CUDA-core kernel (must be 0)
    env.barrier()
    decays, launch_note = job.decays, job.note
    del job
    torch.cuda.empty_cache()

    stages = None
    if env.rank == 0 and env.world == 1 and spec.get('stage_table') and not args.no_stages and not args.fp32:
        stages = stage_table(env, args, K, last=(18, res, value))
    eager = None
    if env.rank == 0 and env.world == 1 and not args.skip_eager and bf16 and spec.get('eager_arch'):
        try:
            eager = gpu_eager_baseline(env, spec['eager_arch'], spec['img_size'], res, B, ema=not args.no_ema)
            eager['speedup_device_resident'] = round(value / eager['value'], 2)
        except Exception as e:   # noqa: BLE001
            eager = {'error': f'{type(e).__name__}: {str(e)[:120]}'}
    cpu = None
    if env.rank == 0 and env.world == 1 and not args.skip_cpu and spec.get('cpu_arch'):
        v, cores, sample = cpu_path_images_per_sec(args.cpu_steps, 1, spec['cpu_arch'], 224)
        cpu = {'value': round(v, 3), 'unit': 'images/s', 'cores': cores, 'kind': 'port', 'sample': sample}
    if env.rank != 0:
        return
    line = {
        'metric': 'train images/sec', 'value': round(value, 1), 'unit': 'images/s', 'n_gpus': env.world, 'steps': args.steps,
        'warmup': max(3, args.warmup), 'ms_per_step': round(ms / args.steps, 3), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'bf16' if bf16 else 'f32', 'data': 'synthetic',
        'config': {'workload': spec['workload'], 'model': spec['model'], 'res': res, 'per_gpu_batch': B, 'global_batch': B * env.world,
                   'parallelism': f'dp{env.world}', 'optimizer': f'fused AdamW + {len(decays)} EMA', 'drop_path': 0.1, 'launch': launch_note,
                   'l2': 'per-step working set (activations, GBs) exceeds the 126 MB L2; no explicit flush'},
        'e2e': e2e, 'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roof, 'cpu_baseline': cpu,
        'gpu_eager_baseline': eager,
    }
    line.update(extra)
    if stages is not None:
        line['autoprog_stages'] = stages
    emit(line)


def stage_table(env, args, K, last=None):
    """images/s of each AutoProg stage sub-net (depth, resolution, drop-path of progressive_schedule)."""
    torch = env.torch
    rows = []
    for l, r, dp in STAGES:
        if last is not None and l == last[0] and r == last[1]:
            rows.append({'l': l, 'r': r, 'images_per_s': round(last[2], 1)})
            continue
        try:
            job = TrainJob(env, 'model_variant', {'variant': f'volo_h12_l{l}'}, 224, r, 128, drop_path=dp, lr=1e-3,
                           ddp_mode=args.ddp_mode)
            for _ in range(3):
                job.step()
            ms = env.timed(job.step, args.steps)
            rows.append({'l': l, 'r': r, 'images_per_s': round(128 * env.world * args.steps / (ms / 1e3), 1),
                         'ms_per_step': round(ms / args.steps, 3)})
            job.close()
            del job
            torch.cuda.empty_cache()
        except Exception as e:   # noqa: BLE001
            rows.append({'l': l, 'r': r, 'error': f'{type(e).__name__}: {str(e)[:80]}'})
    rows.sort(key=lambda d: d['l'])
    return rows


def run_stages(env, args):
    """configs[1]: the four stages of the shipped schedule, plus a super-net epoch (random (r, l) per step, seed = epoch,
    main_prog.py:1861, 1907-1910) served by the (r, l)-keyed graph cache."""
    import numpy as np
    torch = env.torch
    from autoprog_b200 import kernels as K
    from autoprog_b200.graph import GraphCache, sample_configs, probe_throughput
    torch.manual_seed(0)
    np.random.seed(0 + env.rank)
    rows = stage_table(env, args, K)
    ok = [r for r in rows if 'images_per_s' in r]
    # every stage trains the same number of images (25 epochs each): schedule throughput = harmonic mean
    value = len(ok) / sum(1.0 / r['images_per_s'] for r in ok) if ok else 0.0
    full = next((r['images_per_s'] for r in ok if r['l'] == 18), None)
    supernet = None
    try:
        B = 128
        job = TrainJob(env, 'model_variant', {'variant': 'volo_h12_l18'}, 224, 224, B, graph=False, lr=1e-3)
        cache = GraphCache(job.net, job.crit, job.opt, bf16=True, warmup=2)
        l_list, r_list = [12, 15, 18], [160, 192, 224]          # search candidates of stage 2 -> 3 (3 x 3 set)
        tg = {r: torch.softmax(torch.randn(B, 1000, 2 + (r // 16) ** 2, device=env.dev), dim=1) for r in r_list}
        x224 = job.x
        random.seed(1)                                          # epoch 1: every rank draws the same sequence
        plan = [sample_configs(l_list, r_list)[0] for _ in range(9 * 4)]
        for cfg in plan:                                        # first pass: captures (9 configurations)
            cache.step(x224, tg[cfg['input_size']], cfg)
        cursor = {'i': 0}

        def supernet_step():
            cfg = plan[cursor['i'] % len(plan)]
            cursor['i'] += 1
            cache.step(x224, tg[cfg['input_size']], cfg)
        t_ms = env.timed(supernet_step, len(plan))
        probe = None
        if env.world == 1:
            cfg0 = dict(min_layer_num=12, max_layer_num=18, layer_num=15, input_size=192)
            import autoprog_b200 as A
            sec = probe_throughput(job.model, A.SoftTargetCrossEntropy(), x224, tg[192][:, :, 1].contiguous(), cfg0, steps=20)
            probe = {'config': 'r192_l15', 'fwd_bwd_ms': round(sec * 1e3, 3)}
        supernet = {'candidates': '3 resolutions x 3 depths (r160..224, l12..18), random (r,l) per step, seed=epoch',
                    'steps': len(plan), 'graphs_captured': cache.captures, 'images_per_s': round(B * env.world * len(plan) / (t_ms / 1e3), 1),
                    'throughput_probe': probe, 'includes': 'bilinear 224->r resize + set_sample_config + graph replay per step'}
        cache.close()
    except Exception as e:   # noqa: BLE001
        supernet = {'error': f'{type(e).__name__}: {str(e)[:160]}'}
    if env.rank != 0:
        return
    line = {'metric': 'train images/sec', 'value': round(value, 1), 'unit': 'images/s', 'n_gpus': env.world, 'steps': args.steps,
            'warmup': 3, 'ms_per_step': None, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16',
            'data': 'synthetic',
            'config': {'workload': 'volo_d1 AutoProg schedule: stages (l9@128, l12@160, l15@192, l18@224), per-GPU batch 128, bf16; value = '
                                   'harmonic mean over the four equal-length stages', 'parallelism': f'dp{env.world}'},
            'autoprog_stages': rows, 'supernet_epoch': supernet,
            'schedule_speedup_vs_full_model': round(value / full, 3) if full else None}
    emit(line)


def run_deit(env, args):
    """configs[2]: deit_small progressive (elastic depth, resolution ramp) with token labeling.  The reference has no
    runnable DeiT progressive path (prog/helpers.py:753 'TODO: deit'); this follows the VOLO recipe: l in 6..12, r in 128..224."""
    import numpy as np
    torch = env.torch
    from autoprog_b200 import kernels as K
    torch.manual_seed(0)
    np.random.seed(0 + env.rank)
    rows = []
    for l, r in ((6, 128), (8, 160), (10, 192), (12, 224)):
        try:
            job = TrainJob(env, 'deit_small_patch16_224', {'return_dense': True}, r, r, 128, dense_tuple=True, lr=1e-3,
                           sample_config={'layer_num': l, 'min_layer_num': 6, 'max_layer_num': 12}, ddp_mode=args.ddp_mode)
            for _ in range(3):
                job.step()
            ms = env.timed(job.step, args.steps)
            rows.append({'l': l, 'r': r, 'images_per_s': round(128 * env.world * args.steps / (ms / 1e3), 1),
                         'ms_per_step': round(ms / args.steps, 3), 'launch': job.note})
            job.close()
            del job
            torch.cuda.empty_cache()
        except Exception as e:   # noqa: BLE001
            rows.append({'l': l, 'r': r, 'error': f'{type(e).__name__}: {str(e)[:120]}'})
    ok = [r for r in rows if 'images_per_s' in r]
    value = len(ok) / sum(1.0 / r['images_per_s'] for r in ok) if ok else 0.0
    if env.rank != 0:
        return
    emit({'metric': 'train images/sec', 'value': round(value, 1), 'unit': 'images/s', 'n_gpus': env.world, 'steps': args.steps,
                      'warmup': 3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
                      'config': {'workload': 'deit_small (D=384, 12 layers, 6 heads x 64) progressive: elastic depth l6..12 at r128..224, '
                                             'token-label aux head + TokenLabelCE(dense 0.5), per-GPU batch 128, bf16, fused AdamW + 4 EMA; '
                                             'value = harmonic mean over the four stages', 'parallelism': f'dp{env.world}'},
                      'stages': rows})


def run_micro(env, args):
    """configs[4]: OutlookAttention core and TokenLabelCrossEntropy, fwd+bwd, r x B sweep, against the reference's
    PyTorch ops (oracle.outlook_core = unfold -> softmax -> matmul -> fold; oracle.token_label_ce) on the same GPU, bf16."""
    torch = env.torch
    if env.rank != 0:
        return
    from autoprog_b200 import kernels as K
    from oracle import volo_cpu as O
    peaks = load_peaks()
    dev, bf = env.dev, torch.bfloat16
    torch.manual_seed(0)

    def t_us(fn, n=10):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3

    rows = []
    for r in (112, 128, 160, 192, 224, 288, 320, 384, 448):
        for B in (64, 128, 256, 512):
            H = r // 8
            h = (H + 1) // 2
            N = (r // 16) ** 2
            if B * H * H * 192 * 2 > 6e9:          # bound memory: skip cells whose reference intermediates exceed ~40 GB
                continue
            row = {'r': r, 'B': B}
            try:
                v = torch.randn(B, H, H, 192, device=dev).to(bf)
                lg = (torch.randn(B, h, h, 488, device=dev) * 2).to(bf)
                dy = torch.randn_like(v)
                by = (5 * v.numel() + 3 * B * h * h * 486) * 2
                us = t_us(lambda: (K.outlook_fwd(v, lg, 6, 32 ** -0.5), K.outlook_bwd(v, lg, dy, 6, 32 ** -0.5)))
                row['outlook_us'] = round(us, 1)
                row['outlook_gbs'] = round(by / us / 1e3, 1)
                row['outlook_frac_hbm'] = round(by / us / 1e3 / peaks['hbm_gbs'], 4)
                vr = v.clone().requires_grad_(True)
                lr_ = lg[..., :486].contiguous().requires_grad_(True)

                def ref_outlook():
                    vr.grad = lr_.grad = None
                    O.outlook_core(vr, lr_, 6, 32 ** -0.5).backward(dy)
                row['outlook_ref_us'] = round(t_us(ref_outlook, 3), 1)
                row['outlook_speedup'] = round(row['outlook_ref_us'] / us, 1)
                del v, lg, dy, vr, lr_
            except Exception as e:   # noqa: BLE001
                row['outlook_error'] = f'{type(e).__name__}: {str(e)[:80]}'
            torch.cuda.empty_cache()
            try:
                xa = (torch.randn(B, N, 1000, device=dev) * 2).to(bf)
                xc = (torch.randn(B, 1000, device=dev) * 2).to(bf)
                tg = torch.softmax(torch.randn(B, 1000, 2 + N, device=dev), 1)
                by = xa.numel() * (2 * 2 + 4)
                us = t_us(lambda: K.tlce_fwd_bwd(xc, xa, tg, 4, 1.0, 0.5))
                row['tlce_us'] = round(us, 1)
                row['tlce_gbs'] = round(by / us / 1e3, 1)
                row['tlce_frac_hbm'] = round(by / us / 1e3 / peaks['hbm_gbs'], 4)
                xar, xcr = xa.clone().requires_grad_(True), xc.clone().requires_grad_(True)

                def ref_tlce():
                    xar.grad = xcr.grad = None
                    O.token_label_ce(xcr.float(), xar.float(), (0, 0, 2, 2), tg, dense_weight=0.5).backward()
                row['tlce_ref_us'] = round(t_us(ref_tlce, 3), 1)
                row['tlce_speedup'] = round(row['tlce_ref_us'] / us, 1)
                del xa, xc, tg, xar, xcr
            except Exception as e:   # noqa: BLE001
                row['tlce_error'] = f'{type(e).__name__}: {str(e)[:80]}'
            torch.cuda.empty_cache()
            rows.append(row)
    head = next((x for x in rows if x['r'] == 224 and x['B'] == 128 and 'outlook_gbs' in x), None)
    emit({'metric': 'OutlookAttention fwd+bwd HBM GB/s', 'value': head['outlook_gbs'] if head else None, 'unit': 'GB/s',
                      'n_gpus': 1, 'higher_is_better': True, 'dtype': 'bf16', 'data': 'synthetic', 'vs_baseline': None,
                      'config': {'workload': 'OutlookAttention core (C=192, 6 heads) + TokenLabelCrossEntropy (C=1000) fwd+bwd sweep, '
                                             'r in 112..448, B in 64..512; value = the r224 / B128 cell; reference = oracle torch ops on this GPU'},
                      'peak_hbm_gbs': peaks['hbm_gbs'], 'peak_source': peaks['src'], 'cells': rows})


def main():
    args = parse()
    if not (args.gpus > 1 and 'RANK' not in os.environ):      # (the torchrun re-launcher passes its children's stdout through)
        claim_stdout()
    if args.impl == 'reference':
        run_reference(args)
        return
    if args.gpus > 1 and 'RANK' not in os.environ:     # convenience: re-launch under torchrun
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={args.gpus}',
               '--master-addr', '127.0.0.1', '--master-port', '29531', os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    env = Env()
    try:
        if args.config == 'micro':
            run_micro(env, args)
        elif args.config == 'stages':
            run_stages(env, args)
        elif args.config == 'deit_small':
            run_deit(env, args)
        elif args.config == 'd2_384':
            B = args.batch or 64
            run_headline(env, args, dict(model='volo_d2', img_size=384, res=384, B=B, eager_arch='volo_d2', cpu_arch=None,
                                         workload=f'volo_d2 at 384px (48x48 outlook grid, 576 stage-2 tokens), per-GPU batch {B}, bf16, '
                                                  f'TokenLabelCE(dense 0.5)'))
        else:
            B = args.batch or 128
            model = args.model or 'volo_d1'
            res = args.res or 224
            wl = WORKLOAD if (model, res, B) == ('volo_d1', 224, 128) else f'{model} at {res}px, per-GPU batch {B}, TokenLabelCE(dense 0.5)'
            run_headline(env, args, dict(model=model, img_size=224, res=res, B=B, stage_table=(model == 'volo_d1' and res == 224),
                                         eager_arch=model if model.startswith('volo') else None,
                                         cpu_arch=model if model.startswith('volo') else None, workload=wl))
    finally:
        if env.world > 1:
            env.dist.destroy_process_group()


def measure_rooflines(train_step, x, t, K, torch, B, bf16):
    """Wrap the GEMM / attention / loss entry points with CUDA events for ONE extra (untimed) step."""
    peaks = load_peaks()
    names = ['gemm', 'outlook_fwd', 'outlook_bwd', 'tlce_fwd_bwd', 'mhsa_fwd', 'mhsa_bwd']
    rec = {n: [] for n in names}
    orig = {n: getattr(K, n) for n in names}

    def wrap(name, work):
        fn = orig[name]

        def inner(*a, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **kw)
            e1.record()
            rec[name].append((e0, e1, work(*a, **kw)))
            return out
        return inner

    es = 2 if bf16 else 4

    def gemm_work(a, b, M, N, Kd, **kw):
        oes = 4 if kw.get('out_dtype') == torch.float32 or (kw.get('out') is not None and kw['out'].dtype == torch.float32) else a.element_size()
        nout = 2 if kw.get('epilogue', 0) != K.EPI_NONE else 1          # GELU stores the pre-activation; dGELU reads it
        byts = (M * Kd + N * Kd) * a.element_size() + nout * M * N * oes
        return (2.0 * M * N * Kd, (M, N, Kd, int(kw.get('trans_a', False)), int(kw.get('trans_b', False)), kw.get('epilogue', 0)), byts)

    def mhsa_work(qkv, *a, **kw):
        Bq, N, C3 = qkv.shape
        return 4.0 * Bq * N * N * (C3 // 3)          # QK^T + PV: 2 products x 2 N^2 D per head

    setattr(K, 'gemm', wrap('gemm', gemm_work))
    setattr(K, 'outlook_fwd', wrap('outlook_fwd', lambda v, lg, *a, **kw: (2 * v.numel() + lg.numel()) * es))
    setattr(K, 'outlook_bwd', wrap('outlook_bwd', lambda v, lg, *a, **kw: (3 * v.numel() + 2 * lg.numel()) * es))
    setattr(K, 'tlce_fwd_bwd', wrap('tlce_fwd_bwd', lambda xc, xa, *a, **kw: xa.numel() * (2 * es + 4)))
    setattr(K, 'mhsa_fwd', wrap('mhsa_fwd', mhsa_work))
    setattr(K, 'mhsa_bwd', wrap('mhsa_bwd', lambda qkv, *a, **kw: 2.5 * mhsa_work(qkv)))     # 5 products vs 2
    from autoprog_b200 import ops as _ops
    side_wgrad = _ops.SIDE_WGRAD
    _ops.SIDE_WGRAD = False           # event brackets need one kernel at a time: no dgrad / wgrad overlap in this extra step
    try:
        host_s = 0.05
        for it in range(3):           # first pass warms the caching allocator, second one times the host's enqueue
            for v in rec.values():
                v.clear()
            # park the GPU for as long as the host needs to enqueue the WHOLE step (measured in the previous pass, +50 %):
            # every event pair then brackets back-to-back device work only (an eager host that falls behind would
            # otherwise leave idle time inside the spans)
            torch.cuda._sleep(int(min(host_s * 1.5, 1.0) * 2.0e9))
            h0 = time.perf_counter()
            train_step(x, t)
            host_s = time.perf_counter() - h0
            torch.cuda.synchronize()
    finally:
        _ops.SIDE_WGRAD = side_wgrad
        for n in names:
            setattr(K, n, orig[n])

    if os.environ.get('APB_BENCH_GEMM_TABLE'):
        tab = {}
        for e0, e1, (w, key, by) in rec['gemm']:
            tt = tab.setdefault(key, [0, 0.0, 0.0, 0.0])
            tt[0] += 1; tt[1] += e0.elapsed_time(e1); tt[2] += w; tt[3] += by
        for key, (n, ms, w, by) in sorted(tab.items(), key=lambda kv: -kv[1][1]):
            print(f'[gemm] M,N,K,ta,tb,epi={key} x{n}: {ms:.3f} ms  {w / ms / 1e9:.0f} TFLOP/s  {by / ms / 1e6:.0f} GB/s', file=sys.stderr)
    # per launch: the roof that binds is max(flops / tensor peak, algorithmic bytes / HBM peak)
    roof_ms = sum(max(w / (peaks['bf16_tflops_sustained'] * 1e9), by / (peaks['hbm_gbs'] * 1e6)) for _, _, (w, _, by) in rec['gemm'])
    hbm_bound_ms = sum(e0.elapsed_time(e1) for e0, e1, (w, _, by) in rec['gemm']
                       if by / (peaks['hbm_gbs'] * 1e6) > w / (peaks['bf16_tflops_sustained'] * 1e9))
    gemm_bytes = sum(by for _, _, (_, _, by) in rec['gemm'])
    rec['gemm'] = [(e0, e1, w[0]) for e0, e1, w in rec['gemm']]

    def agg(name):
        ms = sum(e0.elapsed_time(e1) for e0, e1, _ in rec[name])
        work = sum(w for _, _, w in rec[name])
        return ms, work, len(rec[name])

    traffic = load_traffic()           # {} (-> null fields) unless captured from exactly these kernel sources

    def tr(name):
        return round(traffic[name]['dram_bytes_per_launch']) if name in traffic else None

    gms, gflop, gn = agg('gemm')
    tf = gflop / (gms * 1e-3) / 1e12 if gms > 0 else 0.0
    roof = {'kernel': 'tcgen05 GEMM kernels (gemm_tc / gemm_tc2) over all Linear/patchify/conv GEMMs of one step', 'bound': 'tensor',
            'achieved': round(tf, 1), 'peak': peaks['bf16_tflops_sustained'], 'unit': 'TFLOP/s',
            'frac': round(tf / peaks['bf16_tflops_sustained'], 4), 'traffic': tr('gemm_tc'),
            'traffic_unit': 'DRAM bytes per launch (ncu dram__bytes_read+write, profiles/dram_traffic.json; null when that capture '
                            'predates the kernel sources); algorithmic = ' + str(round(gemm_bytes / max(gn, 1))) + ' B per launch',
            'launches_per_step': gn,
            'ms_per_step': round(gms, 3), 'peak_source': peaks['src'] + ' (sustained cuBLAS bf16)',
            # the step's GEMMs are a mix: K=192 layers of stage 1 are HBM-bound, the rest tensor-bound.  frac_binding =
            # sum over launches of max(flop time at tensor peak, algorithmic-byte time at HBM peak) / measured time
            'frac_binding': round(roof_ms / gms, 4) if gms > 0 else None,
            'hbm_bound_share_of_ms': round(hbm_bound_ms / gms, 3) if gms > 0 else None,
            'algorithmic_gbytes': round(gemm_bytes / 1e9, 3)}
    extra = {}
    fms, fby, fn_ = agg('outlook_fwd')
    bms, bby, _ = agg('outlook_bwd')
    if fms + bms > 0:
        gbs = (fby + bby) / ((fms + bms) * 1e-3) / 1e9
        extra['roofline_outlook'] = {'kernel': 'OutlookAttention core fwd+bwd', 'bound': 'hbm', 'achieved': round(gbs, 1),
                                     'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': round(gbs / peaks['hbm_gbs'], 4),
                                     'traffic': tr('outlook'),
                                     'algorithmic_bytes_per_launch': round((fby + bby) / max(2 * fn_, 1)),
                                     'layers': fn_, 'ms_per_step': round(fms + bms, 3)}
    tms, tby, _ = agg('tlce_fwd_bwd')
    if tms > 0:
        gbs = tby / (tms * 1e-3) / 1e9
        extra['roofline_tlce'] = {'kernel': 'TokenLabelCrossEntropy fused fwd+bwd', 'bound': 'hbm', 'achieved': round(gbs, 1),
                                  'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': round(gbs / peaks['hbm_gbs'], 4),
                                  'traffic': tr('tlce'), 'algorithmic_bytes_per_launch': round(tby),
                                  'ms_per_step': round(tms, 3)}
    mf, mfw, mn = agg('mhsa_fwd')
    mb, mbw, _ = agg('mhsa_bwd')
    if mf + mb > 0:
        tfm = (mfw + mbw) / ((mf + mb) * 1e-3) / 1e12
        extra['roofline_mhsa'] = {'kernel': 'MHSA core fwd+bwd (tcgen05 / TMEM)', 'bound': 'tensor', 'achieved': round(tfm, 1),
                                  'peak': peaks['bf16_tflops_sustained'], 'unit': 'TFLOP/s',
                                  'frac': round(tfm / peaks['bf16_tflops_sustained'], 4), 'layers': mn, 'traffic': tr('mhsa'),
                                  'fwd_ms_per_step': round(mf, 3), 'bwd_ms_per_step': round(mb, 3),
                                  'flops': 'algorithmic 4 N^2 D per head fwd, 10 N^2 D bwd (unpadded)'}
    return roof, extra


if __name__ == '__main__':
    main()

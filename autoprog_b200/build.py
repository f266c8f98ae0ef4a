"""In-tree build of the C-ABI kernel library `autoprog_b200/_apb.so` for sm_100a.

    python -m autoprog_b200.build [--force]

nvcc cross-compiles without a GPU.  The library is plain C ABI (include/autoprog_b200.h): no torch,
no pybind -- it is loaded with ctypes (autoprog_b200/_lib.py).
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'build')
LIB = os.path.join(HERE, '_apb.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
FLAGS = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=default',
         '--expt-relaxed-constexpr']


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _stamp(paths):
    h = hashlib.sha1()
    for p in sorted(paths):
        h.update(p.encode())
        with open(p, 'rb') as f:
            h.update(f.read())
    h.update(' '.join(FLAGS + ARCH).encode())
    return h.hexdigest()


def _compile(src):
    obj = os.path.join(OBJ, src[:-3] + '.o')
    cmd = [NVCC] + ARCH + FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'nvcc failed for {src}:\n{r.stdout}\n{r.stderr}')
    return obj


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = _sources()
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(HERE), 'include', 'autoprog_b200.h')]
    stamp = _stamp(deps)
    stamp_file = os.path.join(OBJ, 'stamp')
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    if verbose:
        print(f'[autoprog_b200.build] compiling {len(srcs)} CUDA sources for sm_100a ...', file=sys.stderr)
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(_compile, srcs))
    cmd = [NVCC] + ARCH + ['-shared', '-o', LIB] + objs   # no -lcuda: driver entry points are fetched at run time
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    with open(stamp_file, 'w') as f:
        f.write(stamp)
    if verbose:
        print(f'[autoprog_b200.build] wrote {LIB}', file=sys.stderr)
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv)

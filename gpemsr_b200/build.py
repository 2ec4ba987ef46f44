"""Builds gpemsr_b200/lib/libgpemsr_b200.so (C ABI, sm_100a only) with nvcc, in-tree.

    python -m gpemsr_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The library has no torch dependency: it is a
plain CUDA shared object loaded through ctypes (gpemsr_b200/_lib.py).
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
OBJDIR = os.path.join(HERE, 'build')
LIB = os.path.join(LIBDIR, 'libgpemsr_b200.so')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden', '--expt-relaxed-constexpr',
              '-DGPEMSR_BUILDING=1']


def nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found; cannot build libgpemsr_b200.so')
    return exe


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def _deps_mtime():
    files = sources() + glob.glob(os.path.join(CSRC, '*.h')) + glob.glob(os.path.join(CSRC, '*.cuh')) \
        + glob.glob(os.path.join(HERE, '..', 'include', '*.h')) + [os.path.abspath(__file__)]
    return max(os.path.getmtime(f) for f in files)


def up_to_date():
    return os.path.exists(LIB) and os.path.getmtime(LIB) >= _deps_mtime()


def _compile(src, verbose):
    obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + '.o')
    hdr_m = max([os.path.getmtime(f) for f in glob.glob(os.path.join(CSRC, '*.h')) +
                 glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(HERE, '..', 'include', '*.h'))]
                + [os.path.getmtime(os.path.abspath(__file__))])
    if os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), hdr_m):
        return obj, ''
    cmd = [nvcc()] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'nvcc failed for {src}:\n{r.stdout}\n{r.stderr}')
    return obj, r.stderr


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    if force:
        for f in glob.glob(os.path.join(OBJDIR, '*.o')):
            os.remove(f)
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        res = list(ex.map(lambda s: _compile(s, verbose), sources()))
    if verbose:
        for _, log in res:
            if log:
                print(log, file=sys.stderr)
    cmd = [nvcc(), '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB] + [o for o, _ in res]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    return LIB


if __name__ == '__main__':
    path = build(force='--force' in sys.argv, verbose='--verbose' in sys.argv)
    print(path)

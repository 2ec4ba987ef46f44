"""Ablation timing of the tap-fused GEMM kernel: which of its three roles bounds a tile?

Profiling builds of the library (gemm_core.cuh, GPEMSR_ABLATE bits: 1 = epilogue warps skip their work, 2 = the A tiles are
loaded for the first tile only, 4 = the MMA warp issues no MMAs) are compiled HERE into tools/_bin/ (git-ignored, travels with
gpurun) and each is timed on the single-layer conv cases of tools/microbench.py.  Results of an ablated build are garbage by
construction; only the time matters.  The shipped library is never touched.

    python tools/ablate.py build            # CPU container: nvcc, ~1 min per variant
    python tools/ablate.py run [case ...]   # GPU box: one subprocess per variant -> gpurun_out/ablate.jsonl
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, 'tools', '_bin')
VARIANTS = tuple(int(v) for v in os.environ.get('ABLATE_VARIANTS', '8,9,10,11,15').split(','))
sys.path.insert(0, ROOT)


def lib_path(v):
    return os.path.join(BIN, f'libgpemsr_ablate{v}.so')


def build():
    from gpemsr_b200 import build as B
    B.build()
    os.makedirs(BIN, exist_ok=True)
    src = os.path.join(B.CSRC, 'conv_igemm.cu')
    others = [os.path.join(B.OBJDIR, os.path.basename(s)[:-3] + '.o') for s in B.sources() if s != src]
    for v in VARIANTS:
        obj = os.path.join(BIN, f'conv_igemm_ablate{v}.o')
        subprocess.run([B.nvcc()] + B.NVCC_FLAGS + [f'-DGPEMSR_ABLATE={v}', '-c', src, '-o', obj], check=True)
        subprocess.run([B.nvcc(), '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', lib_path(v), obj] + others, check=True)
        os.remove(obj)
        print('built', lib_path(v), flush=True)


def child(v, cases):
    import gpemsr_b200._lib as L
    L.LIB_PATH = lib_path(v)
    import microbench as M
    for r in M.conv_bench(cases or ['hr64_1280', 'rb64_640']):
        print(json.dumps(dict(ablate=v, name=r['name'], ms=r['ms'])), flush=True)


def run(cases):
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'ablate.jsonl'), 'w') as f:
        for v in VARIANTS:
            if not os.path.exists(lib_path(v)):
                continue
            r = subprocess.run([sys.executable, os.path.abspath(__file__), 'child', str(v)] + cases, capture_output=True, text=True, timeout=300)
            out = [l for l in r.stdout.splitlines() if l.startswith('{')]
            prof = [l for l in r.stdout.splitlines() if l.startswith('cta ')]
            if prof:      # the accounting of the LAST launch of each case (7 launches per case, one block of lines per launch)
                f.write('\n'.join(prof) + '\n')
                print('\n'.join(prof[-8:]), flush=True)
            if r.returncode != 0 and not out:
                out = [json.dumps(dict(ablate=v, error=r.stderr[-400:]))]
            for l in out:
                print(l, flush=True)
                f.write(l + '\n')


if __name__ == '__main__':
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    if sys.argv[1] == 'build':
        build()
    elif sys.argv[1] == 'child':
        child(int(sys.argv[2]), sys.argv[3:])
    else:
        run(sys.argv[2:])

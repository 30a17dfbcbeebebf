"""Build libnerfstyle_b200.so in-tree with nvcc for sm_100a (no torch headers: seconds, not minutes).

    python -m nerfstyle_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'libnerfstyle_b200.so')
OBJ = os.path.join(HERE, 'build')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
SOURCES = ['raymarching.cu', 'gridencoder.cu', 'mlp.cu', 'mlp_tc.cu', 'mlp_f32.cu', 'field_tc.cu', 'nnfm.cu', 'nnfm_tc.cu', 'optim.cu', 'optim_p2p.cu', 'rays.cu', 'occupancy.cu', 'loss.cu']
FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
         '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden', '--expt-relaxed-constexpr']
FLAGS += os.environ.get('NRF_EXTRA_NVCC_FLAGS', '').split()      # experiments (e.g. -DNRF_MBAR_HINT_NS=256)


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    headers.append(os.path.join(HERE, '..', 'include', 'nerfstyle_b200.h'))
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objs = [os.path.join(OBJ, os.path.basename(s) + '.o') for s in srcs]

    def compile_one(pair):
        src, obj = pair
        if not force and not _stale(obj, [src] + headers):
            return
        cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)

    with ThreadPoolExecutor(max_workers=len(srcs)) as ex:
        list(ex.map(compile_one, zip(srcs, objs)))
    if force or _stale(OUT, objs):
        cmd = [NVCC, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', OUT] + objs + ['-lcudart']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))

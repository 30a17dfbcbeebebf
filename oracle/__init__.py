"""CPU oracle for the nerfstyle hot path -- TEST INFRASTRUCTURE ONLY.

A restatement of the reference's CUDA kernels in plain C (``oracle.c``, loaded through ctypes) plus a
torch-CPU restatement of the model glue (``field.py``) and the matching loss (``matching.py``).  It is
the parity checker: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under ``nerfstyle_b200/`` imports it, and the product
path raises if its CUDA library is missing rather than falling back to this code.

Pinning status: the reference ships no tests or golden vectors, so the oracle is pinned
by the closed-form known-answer values of SURVEY.md section 4 and, on the GPU box, against the reference's own
CUDA extensions rebuilt for sm_100a (``oracle/build_ref.sh`` -> ``oracle/_ref/``).
"""
from .ops import *  # noqa: F401,F403
from .ops import build, lib  # noqa: F401


# The reference's own Python callers of the hot path, staged (byte for byte) next to the rebuilt reference extensions so
# that the GPU box -- where /root/reference does not exist -- can run them UNMODIFIED on the drop-in ops
# (tests/refenv.py, tests/test_reference_callers_gpu.py).  Git-ignored like the rest of oracle/_ref/.
REFERENCE_PY = ['renderer.py', 'common.py', 'config.py', 'nerf_lib.py', 'loss.py', 'utils/__init__.py', 'utils/matrix.py',
                'networks/style_nerf.py', 'networks/tcnn_nerf.py', 'cfgs/renderer/default.yaml', 'cfgs/network/default.yaml']


def stage_reference_sources(ref='/root/reference'):
    import os
    import shutil
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref', 'pysrc')
    if not os.path.isdir(ref):
        return out if os.path.isdir(out) else None
    for rel in REFERENCE_PY:
        dst = os.path.join(out, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(ref, rel), dst)
    return out

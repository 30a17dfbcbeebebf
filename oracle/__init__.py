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

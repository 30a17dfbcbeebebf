"""Raw C-ABI timing: nrf_field_forward (one launch) against the four nrf_mlp_forward_ex launches it replaces.
Usage (GPU box): python tools/field_bench.py [B]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerfstyle_b200 import _lib  # noqa: E402

dev = torch.device('cuda:0')
lib = _lib.lib()


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4_100_000
    K = 8
    st = torch.cuda.current_stream().cuda_stream
    ed = (torch.randn(B, 32, device=dev) * 0.5).half()
    ec = (torch.randn(B, 32, device=dev) * 0.5).half()
    w = {n: (torch.randn(k, device=dev) * 0.1).half() for n, k in (('d', 64 * 32 + 1024), ('k', 64 * 32 + 1024), ('c1', 64 * 32 + 1024), ('c2', 64 * 16 + 4096 + 1024))}
    sig = torch.empty(B, 1, device=dev)
    rgbs = torch.empty(B, 3 + K, device=dev)
    c1 = torch.empty(B, 16, device=dev, dtype=torch.float16)

    def fused():
        lib.nrf_field_forward(ed.data_ptr(), ec.data_ptr(), w['d'].data_ptr(), w['k'].data_ptr(), w['c1'].data_ptr(), w['c2'].data_ptr(), B, K,
                              sig.data_ptr(), rgbs.data_ptr(), 3 + K, c1.data_ptr(), None, st)

    def four():
        lib.nrf_mlp_forward_ex(ed.data_ptr(), 1, w['d'].data_ptr(), B, 32, 1, 1, 64, 1, 4, sig.data_ptr(), 0, 1, st)
        lib.nrf_mlp_forward_ex(ec.data_ptr(), 1, w['c1'].data_ptr(), B, 32, 16, 1, 64, 1, 0, c1.data_ptr(), 1, 16, st)
        lib.nrf_mlp_forward_ex(c1.data_ptr(), 1, w['c2'].data_ptr(), B, 16, 3, 2, 64, 1, 2, rgbs.data_ptr(), 0, 3 + K, st)
        lib.nrf_mlp_forward_ex(ec.data_ptr(), 1, w['k'].data_ptr(), B, 32, K, 1, 64, 1, 0, rgbs.data_ptr() + 12, 0, 3 + K, st)
    tf, t4 = timeit(fused), timeit(four)
    flop = 2 * (32 * 64 + 64 + 32 * 64 + 64 * K + 32 * 64 + 64 * 16 + 16 * 64 + 64 * 64 + 64 * 3) * B
    print('B=%d  one launch %.3f ms (%.0f TFLOP/s un-padded)   four launches %.3f ms (%.0f TFLOP/s)' % (B, tf, flop / tf / 1e9, t4, flop / t4 / 1e9))


if __name__ == '__main__':
    main()

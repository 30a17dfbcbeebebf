"""Raw C-ABI timing of nrf_mlp_forward / nrf_mlp_backward for the four nets of the model, both implementations
(tcgen05 and mma.sync), with a sweep of resident CTAs per SM for the tcgen05 kernels.
Usage (GPU box): python tools/mlp_bench.py [B] [quick]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerfstyle_b200 import _lib  # noqa: E402

dev = torch.device('cuda:0')
lib = _lib.lib()
NETS = {'density': (32, 1, 1, 0), 'class': (32, 8, 1, 0), 'color1': (32, 16, 1, 0), 'color2': (16, 3, 2, 2)}


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
    quick = len(sys.argv) > 2
    only = sys.argv[3] if len(sys.argv) > 3 else None
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, (ni, no, nh, oact) in NETS.items():
        if only and name != only:
            continue
        in_pad = (ni + 15) // 16 * 16
        npar = 64 * in_pad + (nh - 1) * 4096 + 16 * 64
        params = (torch.randn(npar, device=dev) * 0.1).half()
        x = torch.randn(B, ni, device=dev).half()
        dy = (torch.randn(B, no, device=dev) * 0.01).half()
        y = torch.empty(B, no, device=dev, dtype=torch.float16)
        dx = torch.empty_like(x)
        dp = torch.zeros(npar, device=dev, dtype=torch.float32)
        io_f = B * (ni + no) * 2
        io_b = B * (2 * ni + no) * 2

        def fwd():
            flush.zero_() if False else None
            lib.nrf_mlp_forward(x.data_ptr(), 1, params.data_ptr(), B, ni, no, nh, 64, 1, oact, y.data_ptr(), 1, st)

        def bwd():
            lib.nrf_mlp_backward(x.data_ptr(), 1, params.data_ptr(), dy.data_ptr(), 1, B, ni, no, nh, 64, 1, oact, 128.0,
                                 dx.data_ptr(), 1, dp.data_ptr(), st)

        lib.nrf_mlp_set_mode(1)
        tf, tb = timeit(fwd), timeit(bwd)
        print('%-8s mma.sync          : fwd %.3f ms (%.0f GB/s io)   bwd %.3f ms (%.0f GB/s io)' % (
            name, tf, io_f / tf / 1e6, tb, io_b / tb / 1e6), flush=True)
        lib.nrf_mlp_set_mode(0)
        for ctas in ([4] if quick else [1, 2, 3, 4, 5, 6, 8]):
            lib.nrf_mlp_set_tuning(ctas, ctas)
            tf, tb = timeit(fwd), timeit(bwd)
            print('%-8s tcgen05 ctas/sm=%d : fwd %.3f ms (%.0f GB/s io)   bwd %.3f ms (%.0f GB/s io)' % (
                name, ctas, tf, io_f / tf / 1e6, tb, io_b / tb / 1e6), flush=True)
        lib.nrf_mlp_set_tuning(5, 4)
        # per-phase cycles of CTA 0 / thread 0 of the tcgen05 backward (see PROF_MARK in mlp_tc.cu)
        prof = torch.zeros(16, dtype=torch.int64, device=dev)
        for ctas in (1, 4):
            lib.nrf_mlp_set_tuning(ctas, ctas)
            lib.nrf_mlp_set_profile(prof.data_ptr())
            bwd()
            torch.cuda.synchronize()
            lib.nrf_mlp_set_profile(None)
            ntile = -(-B // 128) / (148 * min(ctas, 4 if nh == 1 else 2))
            print('   prof ctas/sm=%d cycles/tile: %s' % (ctas, ' '.join('%d:%.0f' % (i, v / ntile) for i, v in enumerate(prof.tolist()[:16]))))
        lib.nrf_mlp_set_tuning(5, 4)


if __name__ == '__main__':
    main()

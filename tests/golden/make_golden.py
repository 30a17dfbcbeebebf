"""Generates tests/golden/*.npz from the CPU oracle (oracle/oracle.c).

The reference ships no golden vectors and its kernels are CUDA-only, so these fixtures
pin the ORACLE (regression) and give the GPU tests committed expected values; the oracle itself is pinned
by the closed-form KATs (tests/test_oracle_kat.py) and, on the GPU box, by the reference's own extensions
rebuilt for sm_100a (tests/test_ref_ext_gpu.py).  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def scene(n_rays=192, H=32, C=2, seed=7):
    rs = np.random.RandomState(seed)
    o = (rs.rand(n_rays, 3).astype(np.float32) - 0.5)
    d = rs.randn(n_rays, 3).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    grid = (rs.rand(C, H ** 3) < 0.35).astype(np.float32)
    return o, d, grid


def main():
    bound, H, C, max_steps = 2.0, 32, 2, 256
    o, d, grid = scene(H=H, C=C)
    bits = oracle.packbits(grid, 0.5)
    aabb = np.array([-bound] * 3 + [bound] * 3, np.float32)
    nears, fars = oracle.near_far_from_aabb(o, d, aabb, 0.2)
    counter = np.zeros(2, np.int32)
    xyzs, dirs, deltas, rays = oracle.march_rays_train(o, d, None, bound, bits, C, H, nears, fars, counter, -1, False, 128,
                                                       True, 0., max_steps, False)
    rs = np.random.RandomState(11)
    M = xyzs.shape[0]
    sigmas = (rs.rand(M).astype(np.float32) * 6.0)
    rgbs = rs.rand(M, 5).astype(np.float32)
    ws, depth, image = oracle.composite_rays_train_forward(sigmas, rgbs, deltas, rays, 1e-4, False)
    gws = rs.randn(rays.shape[0]).astype(np.float32)
    gim = rs.randn(rays.shape[0], 5).astype(np.float32)
    gs, gr = oracle.composite_rays_train_backward(gws, gim, sigmas, rgbs, deltas, rays, ws, image, 1e-4, False)
    np.savez_compressed(os.path.join(HERE, 'march_composite.npz'), rays_o=o, rays_d=d, bitfield=bits, nears=nears,
                        fars=fars, counter=counter, rays=rays, xyzs=xyzs[:512], deltas=deltas[:512], n_rows=M,
                        sigmas=sigmas, rgbs=rgbs, weights_sum=ws, depth=depth, image=image, grad_ws=gws, grad_image=gim,
                        grad_sigmas=gs, grad_rgbs=gr, bound=bound, H=H, C=C, max_steps=max_steps)

    # hash grid: small 8-level grid (T=2^12) so the table fits in a fixture by seed only
    offs, pls = oracle.grid_offsets(3, 8, 2, 2, 16, 12, desired_resolution=512, align_corners=True)
    rs = np.random.RandomState(3)
    emb = rs.uniform(-1, 1, (int(offs[-1]), 2)).astype(np.float32)
    x = rs.rand(96, 3).astype(np.float32)
    x[0] = [0.0, 0.0, 0.0]; x[1] = [1.0, 1.0, 1.0]; x[2] = [0.5, 1.0, 0.25]; x[3] = [1.5, 0.2, 0.2]  # edges + one oob
    out, _, idx = oracle.grid_encode_forward(x, emb, offs, pls, 16, False, 0, True, 0, return_indices=True)
    out_h, _ = oracle.grid_encode_forward(x, emb.astype(np.float16), offs, pls, 16, False, 0, True, 0, half=True)
    g = rs.randn(96, 16).astype(np.float32)
    ge = oracle.grid_encode_backward(g, x, offs, int(offs[-1]), 2, pls, 16, 0, True, 0)
    np.savez_compressed(os.path.join(HERE, 'grid_small.npz'), offsets=offs, per_level_scale=pls, emb_seed=3, inputs=x,
                        outputs=out, outputs_half=out_h, indices=idx, grad=g, grad_embeddings=ge)
    print('golden fixtures written to', HERE)


if __name__ == '__main__':
    main()

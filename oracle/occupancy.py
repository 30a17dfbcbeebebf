"""CPU oracle for the occupancy-grid update -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

numpy float32 restatement of Renderer.update_state / _compute_occ_sigmas (/root/reference/renderer.py:120-194) around the
density query, operation for operation (torch's elementwise kernels do not contract multiplies and adds; dividing a tensor by
a host scalar multiplies by the scalar's float32 reciprocal).  The reference ships no vectors for this function: parity
unpinned by the reference; pinned by running the reference's op sequence itself on the GPU next to the fused kernels
(tests/test_pipeline_gpu.py::test_update_state_fused_full_phase_equals_reference_ops).
"""
import numpy as np

from .ops import morton3D, packbits

f32 = np.float32


def cascade_constants(cascade, bound, grid_size):
    """(bound_c - half_grid_size_c, half_grid_size_c) per cascade as float32 (renderer.py:124-127)."""
    out = []
    for cas in range(cascade):
        b = min(2 ** cas, bound)
        hgs = b / grid_size
        out.append((f32(b - hgs), f32(hgs)))
    return out


def cell_points(coords, cas_scale, cas_hgs, noise, grid_size):
    """coords [n,3] int cell coordinates, noise [n,3] in [0,1) -> jittered sample points [n,3] (renderer.py:127-131,155)."""
    inv = f32(1.0) / f32(grid_size - 1)
    xyz = (f32(2.0) * coords.astype(np.float32)) * inv - f32(1.0)                 # 2 * coords.float() / (H - 1) - 1
    pts = xyz * cas_scale
    pts = pts + ((noise.astype(np.float32) * f32(2.0) - f32(1.0)) * cas_hgs)
    return pts.astype(np.float32)


def points_full_morton(noise, cascade, bound, grid_size):
    """All cells of every cascade in Morton order: noise [C, H^3, 3] (Morton order) -> pts [C, H^3, 3]."""
    H = grid_size
    ar = np.arange(H, dtype=np.int32)
    xx, yy, zz = np.meshgrid(ar, ar, ar, indexing='ij')
    coords = np.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], axis=-1)
    mort = np.asarray(morton3D(coords)).astype(np.int64)
    by_morton = np.empty_like(coords)
    by_morton[mort] = coords                                                     # cell coordinates of Morton index i
    out = np.empty((cascade, H ** 3, 3), np.float32)
    for cas, (scale, hgs) in enumerate(cascade_constants(cascade, bound, grid_size)):
        out[cas] = cell_points(by_morton, scale, hgs, noise[cas], grid_size)
    return out


def grid_update(density_grid, tmp_grid, decay, density_thresh):
    """renderer.py:183-189: masked decay / max, mean of the clamped grid, threshold, bitfield."""
    g = density_grid.astype(np.float32).copy()
    valid = (g >= 0) & (tmp_grid >= 0)
    g[valid] = np.maximum(g[valid] * f32(decay), tmp_grid[valid])
    mean = float(np.mean(np.clip(g, 0, None), dtype=np.float64))
    thresh = min(mean, density_thresh)
    return g, mean, packbits(g, thresh)

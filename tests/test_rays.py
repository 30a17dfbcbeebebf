"""Ray generation (SURVEY.md 8f NEXT-1): `nerfstyle_b200.nerf_lib.generate_rays` against tests/golden/rays.npz -- outputs
of the REFERENCE's own NerfLib.generate_rays run on CPU (tests/golden/make_rays_golden.py) -- and against the numpy oracle.
Tolerance: 3e-7 absolute on unit directions (2 float32 ulp at 1.0; summation / norm order is not defined by the
reference's einsum), origins and target pixels bit-exact."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'rays.npz')
TOL = 3e-7


def synth_img(w, h):
    c, y, x = np.meshgrid(np.arange(3), np.arange(h), np.arange(w), indexing='ij')
    return (((x * 7 + y * 13 + c * 29) % 251) / 251.0).astype(np.float32)


def _cases():
    g = np.load(GOLD)
    for name in g['names']:
        name = str(name)
        w, h, fx, fy, cx, cy = g[name + '/intr']
        flip, precrop, bs, px, py, pw, ph = g[name + '/args']
        yield dict(name=name, w=int(w), h=int(h), fx=fx, fy=fy, cx=cx, cy=cy, flip=int(flip), precrop=float(precrop),
                   bsize=int(bs), patch=None if px < 0 else (int(px), int(py), int(pw), int(ph)), pose=g[name + '/pose'],
                   indices=g[name + '/indices'], origins=g[name + '/origins'], dirs=g[name + '/dirs'], target=g[name + '/target'])


def test_oracle_matches_reference_outputs():
    from oracle import rays
    n = 0
    for c in _cases():
        img = synth_img(c['w'], c['h']) if c['patch'] is None else None
        o, d, t = rays.generate_rays(c['pose'], c['w'], c['h'], c['fx'], c['fy'], c['cx'], c['cy'], img=img, patch=c['patch'],
                                     precrop=c['precrop'], indices=c['indices'] if c['bsize'] > 0 else None,
                                     camera_flip=c['flip'])
        if c['bsize'] == -2:                    # full frame stored sub-sampled
            o, d = o[c['indices']], d[c['indices']]
        assert o.shape == c['origins'].shape and np.array_equal(o, c['origins']), c['name']
        assert np.abs(d - c['dirs']).max() <= TOL, c['name']
        assert np.abs(np.linalg.norm(d.astype(np.float64), axis=1) - 1).max() < 2e-7
        if c['target'].size:
            assert np.array_equal(t, c['target']), c['name']
        n += 1
    assert n == 8


def test_host_mirror_types():
    """Intrinsics.scale / Box2D follow common.py:25-114 (values from the room data set)."""
    from nerfstyle_b200.nerf_lib import Box2D, Intrinsics, NerfLib
    base = Intrinsics(378, 504, 383.829783860205, 383.829783860205, 252.0, 189.0)
    big = base.scale(1008, 756)
    assert (big.w, big.h, big.cx, big.cy) == (1008, 756, 504.0, 378.0) and abs(big.fx - 2 * base.fx) < 1e-9
    tall = base.scale(504, 756)               # narrower aspect: focal follows the width ratio
    assert tall.fx == base.fx
    assert Box2D(3, 4, 5, 6).wrange() == slice(3, 8) and Box2D(3, 4, 5, 6).hrange() == slice(4, 10)
    lib = NerfLib()
    with pytest.raises(AssertionError):
        lib.generate_rays(np.eye(4), base)      # no device assigned (nerf_lib.py:15-19)
    with pytest.raises(AssertionError):
        lib.device = torch.device('cpu')        # nerf_lib.py:37


@pytest.mark.gpu
def test_generate_rays_matches_reference_outputs(cuda_lib, dev):
    from nerfstyle_b200.nerf_lib import Box2D, Intrinsics, NerfLib
    from oracle import rays as orays
    lib = NerfLib()
    lib.device = dev
    for c in _cases():
        intr = Intrinsics(c['h'], c['w'], c['fx'], c['fy'], c['cx'], c['cy'])
        img = torch.from_numpy(synth_img(c['w'], c['h'])).to(dev) if c['patch'] is None else None
        patch = Box2D(*c['patch']) if c['patch'] is not None else None
        bsize = c['bsize'] if c['bsize'] > 0 else None
        rb, target = lib.generate_rays(torch.from_numpy(c['pose']).to(dev), intr, img=img, patch=patch, precrop=c['precrop'],
                                       bsize=bsize, camera_flip=c['flip'],
                                       indices=torch.from_numpy(c['indices']) if bsize else None)
        o, d = rb.origins.cpu().numpy(), rb.dirs.cpu().numpy()
        eo, ed, _ = orays.generate_rays(c['pose'], c['w'], c['h'], c['fx'], c['fy'], c['cx'], c['cy'], patch=c['patch'],
                                        precrop=c['precrop'], indices=c['indices'] if bsize else None, camera_flip=c['flip'])
        assert d.shape == ed.shape and np.abs(d - ed).max() <= TOL, c['name']          # against the oracle, every ray
        if c['bsize'] == -2:
            o, d = o[c['indices']], d[c['indices']]
        assert np.array_equal(o, c['origins']), c['name']                                # against the reference's outputs
        assert np.abs(d - c['dirs']).max() <= TOL, c['name']
        if c['target'].size:
            assert np.array_equal(target.cpu().numpy(), c['target']), c['name']
        assert len(rb) == d.shape[0] or c['bsize'] == -2


@pytest.mark.gpu
def test_generate_rays_device_draw_and_frame_sizes(cuda_lib, dev):
    """The device draw is without replacement and inside the crop window; a 1008x756 frame (config 3) matches the oracle."""
    from nerfstyle_b200.nerf_lib import Intrinsics, NerfLib
    from oracle import rays as orays
    lib = NerfLib()
    lib.device = dev
    base = Intrinsics(378, 504, 383.829783860205, 383.829783860205, 252.0, 189.0)
    pose = np.eye(4, dtype=np.float32)
    pose[:3, 3] = [0.1, -0.2, 0.3]
    # identity pose, no flip: the direction through pixel (ix, iy) is ((ix+.5-cx)/fx, (iy+.5-cy)/fy, 1) normalised, so
    # the pixel can be read back from the ray -> the draw's pixels are recoverable
    gen = torch.Generator(device=dev).manual_seed(3)
    rb, _ = lib.generate_rays(pose, base, precrop=0.5, bsize=8192, generator=gen)
    d = rb.dirs.double().cpu().numpy()
    ix = np.rint(d[:, 0] / d[:, 2] * base.fx + base.cx - 0.5).astype(np.int64)
    iy = np.rint(d[:, 1] / d[:, 2] * base.fy + base.cy - 0.5).astype(np.int64)
    assert len(np.unique(iy * 504 + ix)) == 8192                        # replace=False
    assert ix.min() >= 126 and ix.max() < 126 + 252 and iy.min() >= 94 and iy.max() < 94 + 189   # dx, dy of nerf_lib.py:110-111
    with pytest.raises(ValueError):
        lib.generate_rays(pose, base, precrop=0.1, bsize=8192)
    big = base.scale(1008, 756)
    rb, _ = lib.generate_rays(pose, big, camera_flip=3)
    eo, ed, _ = orays.generate_rays(pose, 1008, 756, big.fx, big.fy, big.cx, big.cy, camera_flip=3)
    assert rb.dirs.shape == (762048, 3)
    assert np.abs(rb.dirs.cpu().numpy() - ed).max() <= TOL and np.array_equal(rb.origins.cpu().numpy(), eo)
    # empty batch is a no-op
    rb, _ = lib.generate_rays(pose, base, bsize=0, indices=torch.zeros(0, dtype=torch.int64))
    assert len(rb) == 0

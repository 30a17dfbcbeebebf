"""Scratch diagnostic: where do the reference's caller files and the host mirror diverge on the same kernels?"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import numpy as np, torch
import refenv
import test_reference_callers_gpu as T
from oracle import field
from nerfstyle_b200 import raymarching, scenes
dev = torch.device('cuda:0')
of = field.OracleField(bound=2.0, n_classes=8, half=True, seed=5, table_std=0.5)
pose = T._pose(dev)
bits = raymarching.packbits(scenes.analytic_density_grid(2, 128, 2.0).to(dev), 0.5)
with refenv.ReferenceEnv() as E:
    base = E.common.Intrinsics(378, 504, 383.829783860205, 383.829783860205, 252.0, 189.0)
    intr = base.scale(168, 126)
    model, r, _ = T._reference_stack(E, dev, of, intr=intr)
    r.density_bitfield = bits.clone()
    m2, r2 = T._mirror_stack(dev, of)
    r2.density_bitfield = bits.clone()
    rays, _ = E.nerf_lib.nerf_lib.generate_rays(pose, intr, camera_flip=3)
    o, d = rays.origins.clone(), rays.dirs.clone()
    def run(tag, fn):
        with torch.no_grad(), torch.autocast('cuda', dtype=torch.float16):
            return fn()
    a1 = run('ref', lambda: r.render_test(rays))
    a2 = run('ref', lambda: r.render_test(rays))
    b1 = run('mir', lambda: r2.render_test(o, d, sync_every=1))
    b2 = run('mir', lambda: r2.render_test(o, d, sync_every=1))
    print('ref self', [torch.equal(x, y) for x, y in zip(a1, a2)])
    print('mir self', [torch.equal(x, y) for x, y in zip(b1, b2)])
    print('ref vs mir', [torch.equal(x, y) for x, y in zip(a1, b1)], [float((x - y).abs().max()) for x, y in zip(a1, b1)])
    # mirror renderer driving the reference model, and vice versa
    r2.model = model
    c1 = run('mir+refmodel', lambda: r2.render_test(o, d, sync_every=1))
    print('mirror renderer + ref model vs ref', [torch.equal(x, y) for x, y in zip(a1, c1)])
    r2.model = m2
    r.model = m2
    d1 = run('ref+mirmodel', lambda: r.render_test(rays))
    print('ref renderer + mirror model vs ref', [torch.equal(x, y) for x, y in zip(a1, d1)], 'vs mir', [torch.equal(x, y) for x, y in zip(b1, d1)])
    r.model = model
    # no autocast at all
    with torch.no_grad():
        e1 = r2.render_test(o, d, sync_every=1)
        e2 = r2.render_test(o, d, sync_every=1)
    print('mirror no-autocast self', [torch.equal(x, y) for x, y in zip(e1, e2)])

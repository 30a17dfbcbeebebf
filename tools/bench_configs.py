"""Measurements for BASELINE.json configs 3-5 (bench.py covers config 2, tests cover config 1).

    python tools/bench_configs.py render  [--w 1008 --h 756 --frames 3]        # config 3 (torchrun for N>1: tiles shard)
    python tools/bench_configs.py nnfm                                         # config 4: matching kernel vs torch composition
    python tools/bench_configs.py sweep                                        # config 5: rays/step sweep (torchrun for N>1)
    python tools/bench_configs.py style                                        # config 4: one whole stylization step (trainers/style.py:162-207)
Each prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerfstyle_b200 import model as M, nnfm, parallel, raymarching, scenes  # noqa: E402
import bench as B  # noqa: E402


def setup():
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    return world, rank, dev


def sync(world):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def render(args):
    world, rank, dev = setup()
    torch.manual_seed(0)
    m = M.StyleTCNerf([-2., -2., -2.], [2., 2., 2.], class_dim=B.N_CLASSES).to(dev)
    r = M.Renderer(m, 2.0, raymarch_channels=3 + B.N_CLASSES, density_scale=args.density_scale).to(dev)
    if args.occupancy == 'field':
        with torch.autocast('cuda', dtype=torch.float16):
            r.update_state()
    else:
        r.density_bitfield = raymarching.packbits(scenes.analytic_density_grid(2, 128, 2.0).to(dev), 0.5)
    intr = scenes.scaled_intrinsics(args.w, args.h)
    poses = scenes.synthetic_poses(8, 1)
    n_total = args.w * args.h
    lo, hi = parallel.shard_bounds(n_total, rank, world)
    idx = torch.arange(lo, hi, device=dev)
    times = []
    for f in range(args.frames + 1):
        o, d = scenes.generate_rays(poses[f % len(poses)], intr, dev, idx)
        sync(world)
        t0 = time.perf_counter()
        with torch.no_grad(), torch.autocast('cuda', dtype=torch.float16):
            if args.loop == 'graph':
                img, depth, cls = r.render_test_graph(o, d)
            else:
                img, depth, cls = r.render_test(o, d, sync_every=args.sync_every)
        full = parallel.gather_rows(torch.cat([img, depth[:, None], cls], dim=1), n_total, rank, world)
        sync(world)
        if f > 0:
            times.append(time.perf_counter() - t0)
    t = torch.tensor([sum(times) / len(times)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        sec = float(t.item())
        print(json.dumps({'config': 'render_full_frame', 'w': args.w, 'h': args.h, 'n_gpus': world, 'rays_per_frame': n_total,
                          'ms_per_frame': round(sec * 1e3, 2), 'mrays_per_s': round(n_total / sec / 1e6, 3),
                          'occupancy': args.occupancy, 'density_scale': args.density_scale, 'frames': args.frames,
                          'sync_every': args.sync_every, 'loop': args.loop,
                          'gathered_bytes': int(n_total * (4 + B.N_CLASSES) * 4)}))
    if world > 1:
        dist.destroy_process_group()


def nnfm_bench(args):
    world, rank, dev = setup()
    from oracle import matching as om      # the reference composition (torch ops) timed beside the kernel
    N1, N2, K = 11844, 15876, 768
    g = torch.Generator().manual_seed(0)
    a = torch.randn(N1, K, generator=g).to(dev)
    b = torch.randn(N2, K, generator=g).to(dev)
    a_hat = (a / a.norm(dim=1, keepdim=True)).half()
    b_hat = (b / b.norm(dim=1, keepdim=True)).half()
    preds = torch.randint(0, 8, (N1,), generator=g).to(dev)
    clusters = (torch.arange(N2) * 8 // N2).to(dev)
    match = list(range(8))

    from nerfstyle_b200 import _lib

    def ours():
        _lib.lib().nrf_nnfm_set_mode(0)
        return nnfm.nn_match(a_hat, b_hat, preds, clusters, match)

    def ours_mma_sync():
        _lib.lib().nrf_nnfm_set_mode(1)
        out = nnfm.nn_match(a_hat, b_hat, preds, clusters, match)
        _lib.lib().nrf_nnfm_set_mode(0)
        return out

    def ref():
        with torch.autocast('cuda', dtype=torch.float16):
            return om.semantic_nn_loss(a, b, preds, clusters, match, 8)
    res = {}
    for name, fn in (('ours', ours), ('ours_mma_sync', ours_mma_sync), ('torch_composition', ref)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        res[name] = {'ms': round(ms, 3), 'tflops': round(2.0 * N1 * N2 * K / ms / 1e9, 1),
                     'peak_mem_mb': round(torch.cuda.max_memory_allocated() / 2 ** 20, 1)}
    print(json.dumps({'config': 'nnfm_matching', 'N1': N1, 'N2': N2, 'K': K, **res}))


def style(args):
    """One stylization training step the way StyleTrainer.run_iter does it (trainers/style.py:162-207), on this repo's ops:
    (1) full-frame render WITHOUT gradients (render_train on all 504x378 rays), (2) VGG-16 relu3 features of the render, the
    target and the style image, content MSE + segment-wise nearest-neighbour matching loss (loss.py:187-214, nrf_nnfm_forward),
    backward to d loss / d pixel, (3) the frame again in 200x200 patches WITH gradients (nerf_lib.generate_rays(patch=...)),
    each patch's render backpropagated with its slice of the cached pixel gradients, (4) GradScaler + Adam step on the colour
    hash table only (StyleTrainer.OPTIM_KEYS).  VGG-16 is torchvision's (weights=None: no network here) running on cuDNN --
    library code, outside the hot path; everything else is this repo's kernels."""
    import itertools
    import torchvision
    from nerfstyle_b200 import nnfm as NN
    from nerfstyle_b200.nerf_lib import Box2D, Intrinsics, NerfLib
    from nerfstyle_b200.optim import FusedAdamEMA
    world, rank, dev = setup()
    torch.manual_seed(0)
    W, H, K = 504, 378, B.N_CLASSES
    m = M.StyleTCNerf([-2., -2., -2.], [2., 2., 2.], class_dim=K).to(dev)
    r = M.Renderer(m, 2.0, raymarch_channels=3 + K).to(dev)
    with torch.autocast('cuda', dtype=torch.float16):
        r.update_state()
    r.update_occ = False
    intr = Intrinsics(H, W, scenes.ROOM['fx'], scenes.ROOM['fy'], scenes.ROOM['cx'], scenes.ROOM['cy'])
    pose = torch.from_numpy(scenes.synthetic_poses(2, 0)[0]).to(dev)
    nl = NerfLib()
    nl.device = dev
    flip = scenes.ROOM['flip_camera']
    all_idx = torch.arange(W * H, device=dev)
    tgt_rgb, tgt_cls = scenes.synthetic_target(all_idx, dict(scenes.ROOM))
    target_chw = tgt_rgb.t().reshape(3, H, W).contiguous()
    vgg = torchvision.models.vgg16(weights=None).features[:16].to(dev).eval()
    for p_ in vgg.parameters():
        p_.requires_grad_(False)

    def fx(img_chw):                       # networks/fx.py: 'relu3' = the three relu3_x maps concatenated (768 channels)
        x, outs = img_chw[None], []
        for i, layer in enumerate(vgg):
            x = layer(x)
            if i in (11, 13, 15):
                outs.append(x)
        return torch.cat(outs, dim=1)[0]
    style_img = torch.rand(3, 504, 504, device=dev)
    with torch.no_grad(), torch.autocast('cuda', dtype=torch.float16):
        style_feat = fx(style_img)
        target_feat = fx(target_chw)
    hs, ws = style_feat.shape[1:]
    clusters = (torch.arange(ws, device=dev) * K // ws)[None, :].expand(hs, ws).contiguous()
    matching = list(range(K))
    if args.freeze:      # the reference leaves requires_grad on (its optimizer just ignores the other gradients); freezing skips them
        for n_, p_ in m.named_parameters():
            p_.requires_grad_(n_ == 'x_color_embedder.embeddings')
    opt = FusedAdamEMA([m.x_color_embedder.embeddings], lr=0.01, eps=1e-15, lr_decay_steps=30000, ema_decay=None, enable_amp=True)
    ps = 200
    content_lambda, style_lambda = 1.0, 1.0
    def step():
        opt.zero_grad()
        for p_ in m.parameters():
            p_.grad = None
        with torch.no_grad(), torch.autocast('cuda', dtype=torch.float16):
            rays, _ = nl.generate_rays(pose, intr, camera_flip=flip)
            image, depth, classes = r.render_train(rays.origins, rays.dirs)
        rgb_map = image.detach().clone().requires_grad_(True)
        with torch.autocast('cuda', dtype=torch.float16):
            rgb_chw = rgb_map.t().reshape(3, H, W)
            feat = fx(rgb_chw)
            preds = torch.argmax(classes.t().reshape(K, H, W), dim=0)
            nh, nw = feat.shape[1:]                # labels_downscale (loss.py:23-28)
            preds_small = preds[torch.linspace(0, H - 1, nh, device=dev).long()[:, None], torch.linspace(0, W - 1, nw, device=dev).long()]
            content = torch.nn.functional.mse_loss(feat.float(), target_feat.float()) * content_lambda
            sty = NN.semantic_nnfm_loss(feat, style_feat, preds_small, clusters, matching) * style_lambda
            total = content + sty
        opt.scale_loss(total).backward()
        grad_map = rgb_map.grad.reshape(H, W, 3)
        for x0, y0 in itertools.product(range(0, W, ps), range(0, H, ps)):
            patch = Box2D(x=x0, y=y0, w=ps, h=ps)
            with torch.autocast('cuda', dtype=torch.float16):
                prays, _ = nl.generate_rays(pose, intr, patch=patch, camera_flip=flip)
                pimg, _, _ = r.render_train(prays.origins, prays.dirs)
            pg = grad_map[patch.hrange(), patch.wrange()].reshape(-1, 3)
            pimg.backward(pg)
        opt.step()
        return total.detach()
    step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.frames):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.frames
    print(json.dumps({'config': 'stylization_step', 'w': W, 'h': H, 'rays_per_pass': W * H, 'patches': 6, 'patch_size': ps,
                      'image_feats': list(target_feat.shape), 'style_feats': list(style_feat.shape), 'ms_per_step': round(ms, 2),
                      'rays_per_s_both_passes': round(2 * W * H / ms * 1e3, 1), 'steps': args.frames, 'loss': float(loss),
                      'optimised': 'x_color_embedder.embeddings', 'others_frozen': bool(args.freeze), 'peak_mem_gb': round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 2),
                      'vgg': 'torchvision vgg16(weights=None).features[:16], cuDNN (library, outside the hot path)'}))


def sweep(args):
    world, rank, dev = setup()
    out = []
    for log2n in range(14, 21):
        n_global = 1 << log2n
        n_local = n_global // world
        # ~520 samples per room-shaped ray at random init, ~600 B of activations + gradients per sample
        need = n_local * 520 * 600
        free, _ = torch.cuda.mem_get_info(dev)
        if need > 0.6 * free:
            out.append({'rays_per_step': n_global, 'skipped': 'needs ~%d GB per GPU' % (need >> 30)})
            continue
        ts = B.build_trainer(dev, True, world)
        host, devb = B.make_batches(10, n_local, rank, world, dev)
        assert devb[0].shape[0] == n_local
        for s in range(4):
            ts.step(*B.unpack(devb[s]), n_global=n_global)
        sync(world)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(4, 10):
            ts.step(*B.unpack(devb[s]), n_global=n_global)
        e1.record()
        sync(world)
        t = torch.tensor([e0.elapsed_time(e1) / 6], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out.append({'rays_per_step': n_global, 'ms_per_step': round(float(t.item()), 3),
                    'rays_per_s': round(n_global / float(t.item()) * 1e3, 1)})
        del ts, host, devb
        torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps({'config': 'ray_batch_sweep', 'n_gpus': world, 'points': out}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('what', choices=['render', 'nnfm', 'sweep', 'style'])
    ap.add_argument('--w', type=int, default=1008)
    ap.add_argument('--h', type=int, default=756)
    ap.add_argument('--frames', type=int, default=3)
    ap.add_argument('--occupancy', default='field', choices=['field', 'analytic'])
    ap.add_argument('--density-scale', type=float, default=1.0)
    ap.add_argument('--sync-every', type=int, default=4)
    ap.add_argument('--loop', default='graph', choices=['graph', 'host'])
    ap.add_argument('--freeze', action='store_true', help='style: requires_grad False on everything but the colour table')
    a = ap.parse_args()
    {'render': render, 'nnfm': nnfm_bench, 'sweep': sweep, 'style': style}[a.what](a)

"""nerfstyle_b200 -- B200-native (sm_100a) NeRF render/train hot path behind the reference's operator surface.

Sub-modules (each mirrors one reference interface):
  raymarching   raymarching/raymarching.py   (near_far_from_aabb, march_rays_train, composite_rays_train, ...)
  gridencoder   gridencoder/grid.py          (GridEncoder, grid_encode)
  tcnn          tinycudann.Network / Encoding (fused 64-wide MLPs on tcgen05, the four field heads in one launch, fp32 parity mode)
  nnfm          loss.py                      (nearest-neighbour feature matching: SemanticStyleLoss, NNFMStyleLoss, labels_downscale)
  nerf_lib      nerf_lib.py generate_rays    (camera rays on the device: NerfLib, Intrinsics, Box2D, RayBatch)
  model         renderer.py / style_nerf.py  (host-side mirror: StyleTCNerf, Renderer incl. the fused occupancy update and
                                              the device-driven inference loop)
  trainer/optim trainers/base.py step        (TrainStep, fused loss head, FusedAdamEMA with paired tables; for N > 1 the gradient
                                              exchange is fused into the optimizer kernel over NVLink peer memory)
  dropin        sys.modules aliases so the reference's renderer.py / networks/*.py import these unchanged

All compute goes through libnerfstyle_b200.so (C ABI: include/nerfstyle_b200.h).  No CPU fallback.
"""
__version__ = '0.2.0'

from . import _lib  # noqa: F401


def lib_path():
    return _lib.LIB_PATH

"""Imports the REFERENCE's own Python callers of the hot path -- unmodified -- on top of nerfstyle_b200.dropin.

    renderer.py, networks/style_nerf.py, networks/tcnn_nerf.py, common.py, config.py, nerf_lib.py, loss.py, utils/

The files are read from /root/reference where it is mounted (the build container) and otherwise from the staged copy
`oracle/_ref/pysrc/` that `oracle.stage_reference_sources()` makes at build() time (git-ignored, travels with gpurun,
exactly like the rebuilt reference extensions in oracle/_ref/).  Test infrastructure only.

Only THIRD-PARTY packages that are absent from this image are stood in for (matplotlib, torch_ema, dacite,
simple_parsing, imageio); every module of the reference itself is the real file.  `raymarching`, `gridencoder` and
`tinycudann` -- the three native dependencies this repository replaces -- come from dropin.install().
"""
import importlib
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, 'oracle', '_ref', 'pysrc')
REF_MODULES = ['utils', 'utils.matrix', 'common', 'config', 'nerf_lib', 'networks', 'networks.tcnn_nerf',
               'networks.style_nerf', 'renderer', 'loss']


def reference_root():
    for p in ('/root/reference', STAGED):
        if os.path.isfile(os.path.join(p, 'renderer.py')):
            return p
    return None


def _third_party_stand_ins():
    """Empty stand-ins for packages the reference imports at module scope but the hot path never calls."""
    made = {}

    def mod(name, **attrs):
        try:
            importlib.import_module(name)
            return
        except Exception:
            pass
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__nrf_stand_in__ = True
        made[name] = m
        sys.modules[name] = m
        if '.' in name:
            parent, leaf = name.rsplit('.', 1)
            setattr(sys.modules[parent], leaf, m)

    mod('matplotlib')
    mod('matplotlib.colors')
    mod('matplotlib.pyplot')
    mod('imageio')

    class ExponentialMovingAverage:                       # torch_ema: only subclassed by utils.EMA at import time
        def __init__(self, parameters, decay, use_num_updates=True):
            self.decay = decay
    mod('torch_ema', ExponentialMovingAverage=ExponentialMovingAverage)

    class DaciteConfig:                                   # dacite: config.py builds one Config(...) at class scope
        def __init__(self, *a, **k):
            pass

    class UnexpectedDataError(Exception):
        pass

    def from_dict(data_class, data, config=None):
        raise NotImplementedError('dacite stand-in: construct the config dataclasses directly')
    mod('dacite', from_dict=from_dict, Config=DaciteConfig)
    mod('dacite.exceptions', UnexpectedDataError=UnexpectedDataError)
    mod('simple_parsing')
    mod('simple_parsing.docstring', get_attribute_docstring=lambda *a, **k: None)
    return made


class ReferenceEnv:
    """Context manager: inside it `import renderer` etc. resolve to the reference's files running on the drop-in."""

    def __init__(self):
        self.root = reference_root()
        if self.root is None:
            raise RuntimeError('reference sources not found: neither /root/reference nor %s (run __graft_entry__.build() '
                               'where the reference is mounted)' % STAGED)

    def __enter__(self):
        import nerfstyle_b200.dropin as dropin
        self._saved = {k: sys.modules.get(k) for k in REF_MODULES}
        for k in REF_MODULES:
            sys.modules.pop(k, None)
        self._path = list(sys.path)
        self._stand_ins = _third_party_stand_ins()
        dropin.install(force=True)
        sys.path.insert(0, self.root)
        self.utils = importlib.import_module('utils')
        self.common = importlib.import_module('common')
        self.config = importlib.import_module('config')
        self.nerf_lib = importlib.import_module('nerf_lib')
        self.tcnn_nerf = importlib.import_module('networks.tcnn_nerf')
        self.style_nerf = importlib.import_module('networks.style_nerf')
        self.renderer = importlib.import_module('renderer')
        for m in (self.utils, self.common, self.config, self.nerf_lib, self.tcnn_nerf, self.style_nerf, self.renderer):
            assert os.path.abspath(m.__file__).startswith(os.path.abspath(self.root)), m.__file__
        return self

    def loss_module(self):
        return importlib.import_module('loss')

    def __exit__(self, *exc):
        sys.path[:] = self._path
        for k in REF_MODULES:
            sys.modules.pop(k, None)
        for k, v in self._saved.items():
            if v is not None:
                sys.modules[k] = v
        for k in self._stand_ins:
            sys.modules.pop(k, None)
        return False

    # ---- the reference's shipped configuration values (cfgs/network/default.yaml, cfgs/renderer/default.yaml) ----------
    def network_config(self, **over):
        c = self.config
        v = dict(network_seed=80000, density_out_dims=16, density_hidden_dims=64, density_hidden_layers=1,
                 rgb_hidden_dims=64, rgb_hidden_layers=2,
                 pos_enc=c.NetworkConfig.HashGridConfig(n_lvls=16, n_feats_per_lvl=2, hashmap_size=19, min_res=16,
                                                        max_res_coeff=1024),
                 dir_enc_sh_deg=4)
        v.update(over)
        return c.NetworkConfig(**v)

    def renderer_config(self, **over):
        v = dict(grid_size=128, grid_bsize=None, update_iter=16, min_near=0.2, t_thresh=1e-4, use_ndc=False,
                 flip_camera=3, max_steps=1024, update_thres=256, density_scale=1, density_thresh=10,
                 density_decay=0.95)
        v.update(over)
        return self.config.RendererConfig(**v)

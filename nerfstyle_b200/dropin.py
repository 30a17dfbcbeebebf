"""Make the reference's own imports resolve to this package.

    import nerfstyle_b200.dropin as dropin; dropin.install()
    import raymarching                      # -> nerfstyle_b200.raymarching
    from gridencoder import GridEncoder     # -> nerfstyle_b200.gridencoder
    import tinycudann as tcnn               # -> nerfstyle_b200.tcnn
    from nerf_lib import nerf_lib           # -> nerfstyle_b200.nerf_lib (device-side generate_rays), opt-in

After install() the reference's renderer.py (`import raymarching`, renderer.py:11),
networks/tcnn_nerf.py (`import tinycudann as tcnn`, `from gridencoder import GridEncoder`, :5,10) and
networks/style_nerf.py (:3) run unmodified on the sm_100a kernels.
"""
import sys
import types


def install(force=False, nerf_lib=False):
    from . import gridencoder as _ge
    from . import raymarching as _rm
    from . import tcnn as _tcnn

    def alias(name, module, public):
        if name in sys.modules and not force and getattr(sys.modules[name], '__nerfstyle_b200__', False) is False:
            raise RuntimeError('dropin.install(): a different %r module is already imported (pass force=True)' % name)
        m = types.ModuleType(name)
        m.__nerfstyle_b200__ = True
        m.__doc__ = module.__doc__
        for k in public:
            setattr(m, k, getattr(module, k))
        sys.modules[name] = m
        return m

    rm = alias('raymarching', _rm, _rm.__all__)
    sub = alias('raymarching.raymarching', _rm, _rm.__all__)
    rm.raymarching = sub
    ge = alias('gridencoder', _ge, ['GridEncoder', 'grid_encode'])
    gsub = alias('gridencoder.grid', _ge, ['GridEncoder', 'grid_encode'])
    ge.grid = gsub
    alias('tinycudann', _tcnn, ['Network', 'Encoding'])
    if nerf_lib:
        # renderer.py:9 / trainers/base.py:19 / render.py:14 do `from nerf_lib import nerf_lib`; the returned RayBatch has
        # the reference's fields (origins, dirs).  Opt-in: the reference's own nerf_lib.py also runs on this machine.
        from . import nerf_lib as _nl
        alias('nerf_lib', _nl, ['nerf_lib', 'NerfLib'])
    return rm, ge, sys.modules['tinycudann']

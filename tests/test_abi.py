"""The C-ABI library loads without a GPU and exports every symbol include/*.h declares.  CPU only."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, 'include', 'nerfstyle_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    return sorted(set(re.findall(r'\b(nrf_\w+)\s*\(', hdr)))


def test_header_symbols_exported(cuda_lib):
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(cuda_lib, n), 'symbol %s declared in the header is not exported' % n


def test_python_binding_covers_header(cuda_lib):
    from nerfstyle_b200 import _lib
    assert sorted(_lib.SIGNATURES.keys()) == _declared()


def test_no_torch_types_in_abi():
    hdr = open(os.path.join(ROOT, 'include', 'nerfstyle_b200.h')).read()
    code = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    assert 'at::' not in code and 'torch' not in code and 'Tensor' not in code


def test_error_reporting_without_gpu(cuda_lib):
    assert cuda_lib.nrf_version() >= 1
    assert cuda_lib.nrf_error_string(0) == b'ok'
    assert cuda_lib.nrf_error_string(-1) == b'invalid argument'
    # argument validation happens before any CUDA call: NULL pointers are rejected, N == 0 is a no-op
    assert cuda_lib.nrf_near_far_from_aabb(None, None, None, 4, 0.2, None, None, None) == -1
    assert cuda_lib.nrf_near_far_from_aabb(None, None, None, 0, 0.2, None, None, None) == 0
    assert cuda_lib.nrf_mlp_forward(None, 0, None, 0, 32, 1, 1, 64, 1, 0, None, 1, None) == 0
    assert cuda_lib.nrf_march_scratch_bytes(8192) >= 8192 // 64 * 4
    # entry points added for SURVEY 8f NEXT-1 / the paired tables: same rules
    assert cuda_lib.nrf_generate_rays(None, 1.0, 1.0, 0.5, 0.5, 0, 0, 4, 16, None, 0, None, 0, 0, None, None, None, None) == -1
    assert cuda_lib.nrf_generate_rays(None, 1.0, 1.0, 0.5, 0.5, 0, 0, 4, 0, None, 0, None, 0, 0, None, None, None, None) == 0
    assert cuda_lib.nrf_grid_encode_forward_pair(None, None, None, None, None, 8, 16, 0.5, 16, 0, 1, 0, 1, None, None, None, None) == -1
    assert cuda_lib.nrf_grid_encode_backward_pair(None, None, None, None, None, 0, 16, 0.5, 16, 0, 1, 0, 1, None, None) == 0
    assert cuda_lib.nrf_adam_step_pair(None, None, None, None, None, None, None, None, None, None, 4, None, 0.01, 0.0, 0.9, 0.999,
                                       1e-15, 0.05, None) == -1
    assert cuda_lib.nrf_adam_step_ex(None, None, None, None, None, None, 0, None, 0.01, 0.0, 0.9, 0.999, 1e-15, 0.05, 1, 1, None) == 0


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under nerfstyle_b200/ may import or load it."""
    pkg = os.path.join(ROOT, 'nerfstyle_b200')
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dp, f)).read()
                assert 'import oracle' not in src and 'from oracle' not in src and 'liboracle' not in src, f


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from nerfstyle_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'nope.so'))
    try:
        _lib.lib()
    except RuntimeError as e:
        assert 'no CPU fallback' in str(e)
    else:
        raise AssertionError('expected RuntimeError')

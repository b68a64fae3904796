"""CPU-side checks of the C ABI: the library loads and exports every symbol include/vlfb.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'vlfb.h')).read()
    return sorted(set(re.findall(r'\b(vlfb_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from vlfb import libvlfb
    lib = libvlfb.load()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(libvlfb.SIGNATURES) == names
    assert lib.vlfb_version() == 100


def test_argument_validation_without_gpu():
    from vlfb import libvlfb
    lib = libvlfb.load()
    assert lib.vlfb_gemm(None, None) == -1
    assert b'bad argument' in lib.vlfb_last_error()
    assert lib.vlfb_set_gemm_backend(7) == -1
    assert lib.vlfb_relu_fwd(None, None, 4, None) == -1


def test_struct_layout_matches_header():
    from vlfb import libvlfb as L
    assert ctypes.sizeof(L.ConvGeom) == 21 * 4
    assert ctypes.sizeof(L.Operand) == 32
    # a, b, g, 6 ints, d, 3 int64, alpha(+pad), 4 pointers, flags(+pad)
    assert ctypes.sizeof(L.GemmParams) == 32 + 32 + 84 + 4 + 24 + 8 + 24 + 8 + 32 + 8

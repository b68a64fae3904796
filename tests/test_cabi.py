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


def test_struct_layout_matches_header(tmp_path):
    """sizeof / offsetof of the C structs, as the C compiler sees include/vlfb.h, equal the ctypes mirror."""
    import subprocess
    from vlfb import libvlfb as L
    fields = ['a', 'b', 'g', 'M', 'split_k', 'd', 'ldd', 'alpha', 'col_scale', 'residual', 'relu_mask', 'flags']
    src = tmp_path / 'layout.c'
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "vlfb.h"\nint main(void) {\n'
                   '  printf("%zu %zu %zu\\n", sizeof(vlfb_conv_geom_t), sizeof(vlfb_operand_t), sizeof(vlfb_gemm_params_t));\n'
                   + ''.join('  printf("%%zu\\n", offsetof(vlfb_gemm_params_t, %s));\n' % f for f in fields)
                   + '  return 0;\n}\n')
    exe = tmp_path / 'layout'
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    sizes, offs = [int(v) for v in out[:3]], [int(v) for v in out[3:]]
    assert sizes == [ctypes.sizeof(L.ConvGeom), ctypes.sizeof(L.Operand), ctypes.sizeof(L.GemmParams)]
    assert offs == [getattr(L.GemmParams, f).offset for f in fields]


def test_wt_job_layout(tmp_path):
    """The device job table of vlfb_weight_transpose_multi is written from numpy: 40-byte records."""
    import subprocess
    src = tmp_path / 'job.c'
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "vlfb.h"\nint main(void) {\n'
                   '  printf("%zu %zu %zu %zu %zu\\n", sizeof(vlfb_wt_job_t), offsetof(vlfb_wt_job_t, wt), '
                   'offsetof(vlfb_wt_job_t, scale), offsetof(vlfb_wt_job_t, Co), offsetof(vlfb_wt_job_t, block_begin));\n'
                   '  return 0;\n}\n')
    exe = tmp_path / 'job'
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)])
    assert subprocess.check_output([str(exe)]).decode().split() == ['40', '8', '16', '24', '36']


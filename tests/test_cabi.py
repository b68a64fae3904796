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



def test_gemm_plan_query_and_optional_tile_widths():
    """Host-only: the (tile width, split-K, tile count) the tensor-core GEMM would use.  Default widths are powers of
    two; vlfb_set_tile_widths(1) adds 96/160/192/224 so that the 49-row-tile layers of res4/res5 fill 148 SMs."""
    import ctypes as C
    from vlfb import libvlfb as L
    lib = L.load()
    before = lib.vlfb_get_tile_widths()

    def plan(M, N, K, extra, split=1, taps=1):
        lib.vlfb_set_tile_widths(extra)
        p = L.GemmParams()
        p.M, p.N, p.K, p.batch, p.taps, p.split_k = M, N, K, 1, taps, split
        bn, sp, t = C.c_int(), C.c_int(), C.c_int()
        assert lib.vlfb_gemm_plan(C.byref(p), 148, C.byref(bn), C.byref(sp), C.byref(t)) == 0
        return bn.value, sp.value, t.value

    try:
        assert plan(6272, 512, 4608, 0) == (256, 1, 98)          # res5 3x3: 98 of 148 SMs
        assert plan(6272, 512, 4608, 1) == (192, 1, 147)         # one full round
        assert plan(6272, 256, 3072, 0) == (128, 1, 98) and plan(6272, 256, 3072, 1) == (96, 1, 147)
        assert plan(200704, 256, 64, 1) == plan(200704, 256, 64, 0)      # large-M layers keep 256
        for M, N, K in [(4, 80, 2560), (3136, 784, 256), (512, 4608, 6272), (64, 224, 802816), (1, 300, 512)]:
            for extra in (0, 1):
                bn, sp, tiles = plan(M, N, K, extra, split=0)
                assert bn % 32 == 0 and 32 <= bn <= 256 and sp >= 1
                assert extra or bn & (bn - 1) == 0
                assert bn == 32 or N > bn // 2                   # never pad N by 2x or more
                assert tiles == -(-M // 128) * -(-N // bn) * sp
        assert lib.vlfb_gemm_plan(None, 148, None, None, None) == -1
    finally:
        lib.vlfb_set_tile_widths(before)


def test_fbo_bank_scan_split_table_properties():
    """Host-only: the split of a RoI's L bank rows over CTAs never leaves an empty split, covers all rows, and the
    workspace query matches it (csrc/fbo.cu vlfb_fbo_bank_scan_splits)."""
    from vlfb import libvlfb as L
    lib = L.load()
    for D, rows_per_tile in [(1024, 8), (2048, 8), (4096, 4)]:
        for R in (1, 2, 3, 4, 16, 37, 64, 255, 256, 1000):
            for Lb in (1, 2, 7, 8, 9, 60, 300, 301, 1200, 3600):
                s = lib.vlfb_fbo_bank_scan_splits(R, Lb, D)
                per = -(-Lb // s)
                assert 1 <= s <= 1024 and (s - 1) * per < Lb <= s * per, (R, Lb, D, s)
                assert s <= -(-Lb // rows_per_tile), 'a split holds at least one tile of rows'
                assert lib.vlfb_fbo_bank_scan_workspace(R, Lb, D) == (R * s * D + R * s * 2) * 4
    for bad in [(0, 300, 2048), (4, 0, 2048), (4, 300, 512), (4, 300, 2047)]:
        assert lib.vlfb_fbo_bank_scan_splits(*bad) == 0 and lib.vlfb_fbo_bank_scan_workspace(*bad) == 0
    # many RoIs: enough CTAs for two per SM without splitting rows needlessly
    assert lib.vlfb_fbo_bank_scan_splits(1000, 17, 2048) <= 3

"""The C-ABI library loads without a GPU and exports every symbol include/vqb200.h declares."""
import ctypes
import pathlib
import re

from vector_quantization_b200 import _lib

HEADER = pathlib.Path(__file__).resolve().parents[1] / 'include' / 'vqb200.h'


def declared_symbols():
    text = re.sub(r'/\*.*?\*/', '', HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r'\b(vqb_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_expected_surface():
    names = declared_symbols()
    assert len(names) >= 24
    for must in ('vqb_assign', 'vqb_pack_rows', 'vqb_gather_ste_loss', 'vqb_quantize_backward',
                 'vqb_scatter_stats', 'vqb_kmeans_ema_update', 'vqb_cvq_update', 'vqb_fsq_forward'):
        assert must in names


def test_library_exports_every_declared_symbol():
    _lib.build()
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in declared_symbols():
        assert hasattr(lib, name), f'{name} declared in vqb200.h but not exported by libvqb200.so'


def test_binding_table_matches_header():
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_load_and_pure_host_entry_points():
    lib = _lib.load()
    assert lib.vqb_abi_version() == 3
    assert [lib.vqb_operand_dp(d) for d in (5, 8, 16, 17, 32, 33, 64, 256, 700, 768)] == \
        [16, 16, 16, 32, 32, 64, 64, 256, 704, 768]
    assert lib.vqb_operand_rows_pad(1) == 256 and lib.vqb_operand_rows_pad(257) == 512
    assert lib.vqb_operand_bytes(8192, 32, 3) == 3 * 8192 * 32 * 2
    assert lib.vqb_loss_partials_count() > 0


def test_bad_arguments_return_error_codes_not_crashes():
    lib = _lib.load()
    assert lib.vqb_pack_rows(None, 0, 10, 8, 0, 1, None, None, None, None, 0, None, 0, None, None) == -1
    assert b'null pointer' in lib.vqb_last_error()
    assert lib.vqb_assign(None, 1, 1, 0, None, 1, 1, 0, 8, None, 0, 0, None, 0, None) == -1


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    """tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, TMA -> UTMALDG (B200_PROFILING.md)."""
    import shutil
    import subprocess
    if shutil.which('cuobjdump') is None:
        import pytest
        pytest.skip('cuobjdump not available')
    sass = subprocess.run(['cuobjdump', '-sass', str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    for mnemonic in ('UTCHMMA', 'LDTM', 'UTMALDG'):
        assert mnemonic in sass, f'{mnemonic} missing from the SASS of libvqb200.so'
    assert 'sm_100a' in sass


def test_integration_sketch_matches_binding_table():
    """The reference-side ctypes stub shown in INTEGRATION.md binds the same arity / argument kinds as _lib.SIGNATURES."""
    import ctypes
    text = (HEADER.parents[1] / 'INTEGRATION.md').read_text()
    kinds = {'P': ctypes.c_void_p, 'I64': ctypes.c_int64, 'I': ctypes.c_int}
    found = dict(re.findall(r'_lib\.(vqb_[a-z0-9_]+)\.argtypes = \[([A-Z0-9, ]+)\]', text))
    assert {'vqb_pack_rows', 'vqb_assign', 'vqb_unpack_keys'} <= set(found)
    for name, args in found.items():
        sketch = [kinds[a.strip()] for a in args.split(',')]
        assert sketch == list(_lib.SIGNATURES[name][1]), name


def test_header_is_plain_c():
    """include/vqb200.h is the drop-in boundary: it must compile as C99 (cgo / JNI / FFI generators read it)."""
    import shutil
    import subprocess
    if shutil.which('gcc') is None:
        import pytest
        pytest.skip('gcc not available')
    for args in (['gcc', '-std=c99', '-pedantic', '-Werror', '-fsyntax-only', '-x', 'c'],
                 ['g++', '-std=c++17', '-Werror', '-fsyntax-only', '-x', 'c++']):
        r = subprocess.run(args + [str(HEADER)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr

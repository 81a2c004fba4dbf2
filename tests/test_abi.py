"""The C-ABI library builds, loads and exports exactly what include/csm_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

from csm_mlx_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "csm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(csmb_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _header_functions() == _lib.exported_symbols()


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    h = ctypes.CDLL(_lib.LIB_PATH)
    for name in _header_functions():
        assert hasattr(h, name), name
    lib = _lib.lib()
    assert lib.csmb_abi_version() == _lib.ABI_VERSION
    assert lib.csmb_strerror(0) == b"ok" and lib.csmb_strerror(-1) == b"invalid argument"


def test_struct_sizes_match_header():
    # csmb_llama: 6 ints + float + 6 arrays of 16 pointers + 2 pointers (8-byte aligned)
    assert ctypes.sizeof(_lib.Llama) == 32 + 6 * 16 * 8 + 16
    assert ctypes.sizeof(_lib.Sampler) == 32
    assert ctypes.sizeof(_lib.Model) == 2 * ctypes.sizeof(_lib.Llama) + 5 * 8 + 6 * 4
    assert ctypes.sizeof(_lib.Batch) == 8 + 7 * 8 + 8       # + int flags, padded to the struct's 8-byte alignment
    assert ctypes.sizeof(_lib.ChainOpts) == 3 * 4 + 4 + 8 and ctypes.sizeof(_lib.FrameOpts) == 4 * 4 + 8


def test_header_is_plain_c():
    """include/csm_b200.h is the boundary a foreign binding compiles against: it must parse as C99 with nothing but
    <stddef.h> / <stdint.h> (no C++, no CUDA, no torch types)."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        import pytest

        pytest.skip("no gcc")
    src = '#include "csm_b200.h"\nint main(void) { csmb_sampler s; csmb_model m; csmb_batch b; csmb_chain_opts c; csmb_frame_opts f; (void)s; (void)m; (void)b; (void)c; (void)f; return CSMB_ABI_VERSION - 3; }\n'
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-x", "c", "-I",
                        os.path.join(ROOT, "include"), "-"], input=src.encode(), capture_output=True)
    assert r.returncode == 0, r.stderr.decode()

"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol that
include/cvd_b200.h declares (no compute calls -- there is no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from cvd_b200 import build
    return build.build()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "cvd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cvd_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("cvd_epi_row_params", "cvd_epi_attn_fwd", "cvd_groupnorm_tokens_fwd", "cvd_layernorm_fwd",
                 "cvd_temporal_attn_fwd", "cvd_tokens_to_video_fwd", "cvd_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} is declared in include/cvd_b200.h but not exported"
    lib.cvd_abi_version.restype = ctypes.c_int
    assert lib.cvd_abi_version() == 1


def test_ctypes_binding_covers_the_header(lib_path):
    from cvd_b200 import _lib
    bound = set(_lib.SIGNATURES) | {"cvd_last_error", "cvd_abi_version"}
    assert bound == set(declared_symbols())
    assert _lib.load() is not None


def test_argument_errors_do_not_need_a_gpu(lib_path):
    """Validation happens before any CUDA call: bad shapes return CVD_E_INVALID with a message."""
    from cvd_b200 import _lib
    lib = _lib.load()
    rc = lib.cvd_layernorm_fwd(None, None, None, None, None, 4, 320, 1, 1e-5, None)
    assert rc == -1 and b"null pointer" in lib.cvd_last_error()
    rc = lib.cvd_epi_attn_fwd(1, 1, 1, None, 1, 1, 2, 2, 15, 8, 40, 1, 4, 256.0, None)
    assert rc == -1 and b"feat" in lib.cvd_last_error()
    with pytest.raises(_lib.CvdError):
        _lib.call("cvd_geglu_fwd", 1, 1, 4, 13, None)


def test_kernels_are_blackwell_native(lib_path):
    """SASS evidence: tcgen05 MMA (UTCHMMA), TMEM load/store (LDTM/STTM), TMA (UTMALDG) in the library."""
    try:
        sass = subprocess.run(["cuobjdump", "-sass", lib_path], capture_output=True, text=True, timeout=300).stdout
    except FileNotFoundError:
        pytest.skip("cuobjdump not on PATH")
    for mnemonic in ("UTCHMMA", "LDTM", "STTM", "UTMALDG"):
        assert mnemonic in sass, f"{mnemonic} missing from the SASS"
    assert "HGMMA" not in sass

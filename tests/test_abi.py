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


def test_device_shape_limits_are_refused_before_any_launch_and_match_the_test_double(lib_path, monkeypatch):
    """What the device kernels are not instantiated for is refused by argument validation, before any CUDA call (so it can
    be checked on a box without a GPU), and tests/emulated_ops.py in strict mode refuses the same shapes -- the CPU dry
    run of the GPU suite (CVD_TEST_EMULATE=1) therefore fails where the device would."""
    import torch
    import emulated_ops
    from cvd_b200 import _lib, ops
    lib = _lib.load()
    P = 1 << 20                                     # any non-null, 16-byte aligned address: never dereferenced on these paths
    cases = [
        ("cvd_epi_attn_fwd", (P, P, P, P, P, P, 2, 2, 24 * 24, 8, 40, 1, 24, 256.0, None), b"not an instantiated"),
        ("cvd_epi_attn_fwd", (P, P, P, P, P, P, 2, 2, 1024, 8, 48, 1, 32, 256.0, None), b"not an instantiated"),
        ("cvd_epi_attn_fwd", (P, P, P, P, P, P, 2, 2, 16, 8, 40, 3, 4, 256.0, None), b"partners"),
        ("cvd_epi_attn_fwd", (P + 8, P, P, P, P, P, 2, 2, 1024, 8, 40, 1, 32, 256.0, None), b"16-byte aligned"),
        ("cvd_epi_attn_fwd_lse", (P, P, P, P, P, P, P, 2, 2, 1024, 8, 48, 1, 32, 256.0, None), b"not an instantiated"),
        ("cvd_epi_attn_bwd", (P, P, P, P, P, P, P, P, P, P, P, P, 2, 2, 1024, 8, 48, 32, 256.0, None), b"not an instantiated"),
        ("cvd_epi_attn_bwd", (P, P, P, P, P, P, P, P, P, P, P, P, 2, 2, 100, 8, 40, 10, 256.0, None), b"not an instantiated"),
        ("cvd_layernorm_fwd", (P, P, P, None, P, 4, 128, 1, 1e-5, None), b"not supported"),
        ("cvd_residual_layernorm_fwd", (P, P, None, 1.0, P, P, None, P, 4, 96, 1, 1e-5, None), b"not supported"),
        ("cvd_geglu_fwd", (P, P, 4, 13, None), b"multiple of 8"),
        ("cvd_temporal_attn_fwd", (P, P, P, P, 4, 33, 8, 40, None), b"F=33"),
        ("cvd_temporal_attn_fwd", (P, P, P, P, 4, 16, 8, 41, None), b"head_dim=41"),
        ("cvd_groupnorm_tokens_fwd", (P, 0, P, P, P, P, 2, 30, 16, 64, 32, 1e-5, 0, None), b"bad shape"),
    ]
    for name, args, msg in cases:
        rc = getattr(lib, name)(*args)
        assert rc == -1 and msg in lib.cvd_last_error(), (name, rc, lib.cvd_last_error())
    # the same shapes through the strict test double
    emulated_ops.install(monkeypatch, strict_shapes=True)
    bf = torch.bfloat16

    def refused(fn, *a, **k):
        with pytest.raises(_lib.CvdError):
            fn(*a, **k)
    part = torch.tensor([1, 0], dtype=torch.int32)
    refused(ops.epi_attn_fwd, torch.zeros(2, 576, 320, dtype=bf), torch.zeros(2, 576, 320, dtype=bf),
            torch.zeros(2, 8, 48, 576, dtype=bf), torch.zeros(2, 12), part, 8, 256.0)                       # (40, 24)
    refused(ops.epi_attn_fwd, torch.zeros(2, 1024, 384, dtype=bf), torch.zeros(2, 1024, 384, dtype=bf),
            torch.zeros(2, 8, 64, 1024, dtype=bf), torch.zeros(2, 12), part, 8, 256.0)                      # (48, 32)
    refused(ops.layernorm, torch.zeros(4, 128), torch.ones(128), torch.zeros(128))
    refused(ops.geglu, torch.zeros(4, 26, dtype=bf))
    refused(ops.temporal_attn, *(torch.zeros(4, 33, 320, dtype=bf),) * 3, 8)


def test_empty_and_degenerate_sizes_are_refused_not_launched(lib_path):
    """Edge cases of every entry point -- empty batches, zero heads / partners / frames, non-square latents, unknown layout
    and dtype codes, negative counts -- come back as CVD_E_INVALID with a message (a zero-sized grid would be a CUDA launch
    error; PyTorch's silent empty result is not reproduced: the product fails loudly)."""
    from cvd_b200 import _lib
    lib = _lib.load()
    P = 1 << 20
    cases = [
        ("cvd_epi_row_params", (P, P, P, 0, 32, 256.0, 2.0, 0.5, 0, 16, None)),
        ("cvd_epi_row_params", (P, P, P, 4, 0, 256.0, 2.0, 0.5, 0, 16, None)),
        ("cvd_epi_row_params", (P, P, P, 4, 128, 256.0, 2.0, 0.5, 0, 16, None)),
        ("cvd_epi_bias_dense", (P, P, 0, 8, 256.0, None)),
        ("cvd_epi_attn_fwd", (P, P, P, P, P, P, 0, 2, 1024, 8, 40, 1, 32, 256.0, None)),
        ("cvd_epi_attn_fwd", (P, P, P, P, P, P, 2, 0, 1024, 8, 40, 1, 32, 256.0, None)),
        ("cvd_epi_attn_fwd", (P, P, P, P, P, P, 2, 2, 1024, 0, 40, 1, 32, 256.0, None)),
        ("cvd_epi_attn_fwd", (P, P, P, P, P, P, 2, 2, 1024, 8, 40, 0, 32, 256.0, None)),
        ("cvd_epi_attn_fwd", (P, P, P, P, P, P, 2, 2, 1000, 8, 40, 1, 32, 256.0, None)),
        ("cvd_epi_attn_bwd", (P, P, P, P, P, P, P, P, P, P, P, P, 0, 2, 1024, 8, 40, 32, 256.0, None)),
        ("cvd_layernorm_fwd", (P, P, P, None, P, 0, 320, 1, 1e-5, None)),
        ("cvd_layernorm_fwd", (P, P, P, P, P, 4, 320, 0, 1e-5, None)),
        ("cvd_residual_add_fwd", (P, P, None, 1.0, None, 0, 320, None)),
        ("cvd_residual_add_fwd", (P, P, None, 1.0, None, 4, 321, None)),
        ("cvd_geglu_fwd", (P, P, 0, 16, None)),
        ("cvd_tokens_to_video_fwd", (P, P, P, P, 0, 0, 320, 16, 64, 0, None)),
        ("cvd_tokens_to_video_fwd", (P, P, P, P, 0, 2, 320, 16, 64, 7, None)),
        ("cvd_temporal_attn_fwd", (P, P, P, P, 0, 16, 8, 40, None)),
        ("cvd_add_pose_tokens_fwd", (P, P, 0, P, 0, 320, 16, 64, None)),
        ("cvd_scale_add_fwd", (P, P, 1.0, P, 0, None)),
        ("cvd_scale_add_fwd", (P, P, 1.0, P, -5, None)),
        ("cvd_pairwise_fundamental", (P, P, P, P, 0, None)),
        ("cvd_plucker_embedding", (P, P, None, P, 0, 8, 8, 1, None)),
        ("cvd_plucker_embedding", (P, P, None, P, 2, 0, 8, 1, None)),
        ("cvd_groupnorm_tokens_fwd", (P, 0, P, P, P, P, 0, 320, 16, 64, 32, 1e-5, 0, None)),
        ("cvd_groupnorm_tokens_fwd", (P, 5, P, P, P, P, 2, 320, 16, 64, 32, 1e-5, 0, None)),
    ]
    for name, args in cases:
        rc = getattr(lib, name)(*args)
        assert rc == -1 and name.encode() in lib.cvd_last_error(), (name, args, rc, lib.cvd_last_error())

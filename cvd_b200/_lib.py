"""ctypes binding of ``libcvd_b200.so`` (C ABI declared in ``include/cvd_b200.h``).

The library is built in-tree by ``cvd_b200/csrc/Makefile`` (``python -m cvd_b200.build`` or
``__graft_entry__.build()``).  There is no fallback: if the shared object is missing or a
call fails, the error is raised to the caller.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CVD_B200_LIB", os.path.join(_HERE, "lib", "libcvd_b200.so"))   # override: kernel experiments only

OK = 0
DTYPE_F32, DTYPE_BF16, DTYPE_F16 = 0, 1, 2
LAYOUT_BF_HW, LAYOUT_BHW_F = 0, 1
FF_NONE, FF_HORIZONTAL, FF_RANDOM_SLOPE = 0, 1, 2
ROWP_STRIDE = 12

_c = ctypes
_P, _I, _F, _LL = _c.c_void_p, _c.c_int, _c.c_float, _c.c_longlong

# name -> argtypes; must list every symbol of include/cvd_b200.h (tests/test_abi.py checks it)
SIGNATURES = {
    "cvd_epi_row_params": [_P, _P, _P, _I, _I, _F, _F, _F, _I, _I, _P],
    "cvd_epi_bias_dense": [_P, _P, _I, _I, _F, _P],
    "cvd_epi_attn_fwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _P],
    "cvd_epi_attn_fwd_sharded": [_P, _P, _P, _I, _I, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _P],
    "cvd_epi_attn_fwd_lse": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _P],
    "cvd_epi_attn_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _P],
    "cvd_epi_attn_fwd_simt": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _F, _P],
    "cvd_epi_attn_fwd_debug": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _P],
    "cvd_groupnorm_tokens_fwd": [_P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _I, _P],
    "cvd_layernorm_fwd": [_P, _P, _P, _P, _P, _I, _I, _I, _F, _P],
    "cvd_residual_layernorm_fwd": [_P, _P, _P, _F, _P, _P, _P, _P, _I, _I, _I, _F, _P],
    "cvd_geglu_fwd": [_P, _P, _I, _I, _P],
    "cvd_residual_add_fwd": [_P, _P, _P, _F, _P, _I, _I, _P],
    "cvd_tokens_to_video_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "cvd_temporal_attn_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _P],
    "cvd_add_pose_tokens_fwd": [_P, _P, _I, _P, _I, _I, _I, _I, _P],
    "cvd_scale_add_fwd": [_P, _P, _F, _P, _LL, _P],
    "cvd_peer_gather_rows": [_P, _I, _P, _P, _I, _I, _LL, _P],
    "cvd_pairwise_fundamental": [_P, _P, _P, _P, _I, _P],
    "cvd_plucker_embedding": [_P, _P, _P, _P, _I, _I, _I, _I, _P],
}

_lib = None


class CvdError(RuntimeError):
    pass


def load():
    """Load the shared library once; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CvdError(
            f"{LIB_PATH} is missing: build it with `python -m cvd_b200.build` "
            "(nvcc, sm_100a). cvd_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    lib.cvd_last_error.restype = _c.c_char_p
    lib.cvd_last_error.argtypes = []
    lib.cvd_abi_version.restype = _I
    lib.cvd_abi_version.argtypes = []
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = _I
        fn.argtypes = argtypes
    _lib = lib
    return lib


def call(name, *args):
    """Invoke ``name`` and raise :class:`CvdError` with the library's message on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != OK:
        raise CvdError(f"{name} failed ({rc}): {lib.cvd_last_error().decode(errors='replace')}")

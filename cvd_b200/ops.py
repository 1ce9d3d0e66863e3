"""Tensor-level wrappers over the C ABI (``include/cvd_b200.h``).

PyTorch is plumbing here: it owns device memory and the stream; every op below passes raw
device pointers to ``libcvd_b200.so`` on ``torch.cuda.current_stream()``.  Inputs must be
CUDA tensors -- there is no CPU path and nothing here ever imports ``oracle/``.
"""
import torch

from . import _lib

_DT = {torch.float32: _lib.DTYPE_F32, torch.bfloat16: _lib.DTYPE_BF16, torch.float16: _lib.DTYPE_F16}

LAUNCHES = 0  # kernels launched through this module (bench.py reports it as gpu_launches)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check(t, dtype=None, name="tensor"):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.CvdError(f"{name} must be a CUDA tensor (cvd_b200 has no CPU path)")
    if not t.is_contiguous():
        raise _lib.CvdError(f"{name} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise _lib.CvdError(f"{name} must be {dtype}, got {t.dtype}")
    return t.data_ptr()


def _count(n=1):
    global LAUNCHES
    LAUNCHES += n


def dv_rows(head_dim):
    """Rows of one head of the transposed-V operand: d values, a row of ones, zero padding."""
    return (head_dim // 16 + 1) * 16


def epi_row_params(F_mats, feat, F_mat_size, theta=None, first_frame_rule=_lib.FF_HORIZONTAL,
                   pixel_band=3.0, decay_alpha=3.0, frames_per_video=16):
    """F_mats fp32 [rows,3,3] -> row table fp32 [rows,12] (EpiEncoding.get_attn_map prepass)."""
    rows = F_mats.shape[0]
    rowp = torch.empty(rows, _lib.ROWP_STRIDE, device=F_mats.device, dtype=torch.float32)
    tp = None
    if first_frame_rule == _lib.FF_RANDOM_SLOPE:
        tp = _check(theta, torch.float32, "theta")
    _lib.call("cvd_epi_row_params", _check(F_mats, torch.float32, "F_mats"), tp, rowp.data_ptr(), rows, feat,
              float(F_mat_size), float(pixel_band), float(decay_alpha), first_frame_rule, frames_per_video, _stream())
    _count()
    return rowp


def epi_bias_dense(rowp, feat, F_mat_size):
    """Parity/debug only: dense bias [rows, N, N] fp32."""
    rows, N = rowp.shape[0], feat * feat
    bias = torch.empty(rows, N, N, device=rowp.device, dtype=torch.float32)
    _lib.call("cvd_epi_bias_dense", _check(rowp, torch.float32, "rowp"), bias.data_ptr(), rows, feat,
              float(F_mat_size), _stream())
    _count()
    return bias


def _attn_args(q, k, vt, rowp, partner, heads):
    R, N, C = q.shape
    d = C // heads
    R_kv = k.shape[0]
    n_partners = partner.shape[0] // R
    assert k.shape == (R_kv, N, C) and vt.shape == (R_kv, heads, dv_rows(d), N), (q.shape, k.shape, vt.shape)
    assert partner.shape[0] == n_partners * R
    feat = int(round(N ** 0.5))
    return R, R_kv, N, d, n_partners, feat


def epi_attn_fwd(q, k, vt, rowp, partner, heads, F_mat_size, dbg_s=None):
    """tcgen05 fused epipolar attention.  q,k bf16 [R,N,C]; vt bf16 [R,H,DV,N]; -> bf16 [R,N,C]."""
    R, R_kv, N, d, n_partners, feat = _attn_args(q, k, vt, rowp, partner, heads)
    out = torch.empty_like(q)
    rp = None if rowp is None else _check(rowp, torch.float32, "rowp")
    args = [_check(q, torch.bfloat16, "q"), _check(k, torch.bfloat16, "k"), _check(vt, torch.bfloat16, "vt"), rp,
            _check(partner, torch.int32, "partner"), out.data_ptr()]
    tail = [R, R_kv, N, heads, d, n_partners, feat, float(F_mat_size), _stream()]
    if dbg_s is None:
        _lib.call("cvd_epi_attn_fwd", *args, *tail)
    else:
        _lib.call("cvd_epi_attn_fwd_debug", *args, _check(dbg_s, torch.float32, "dbg_s"), *tail)
    _count()
    return out


def epi_attn_fwd_sharded(q, k_srcs, vt_srcs, rows_per_src, rowp, partner, heads, F_mat_size):
    """View-sharded attention with the exchange fused into the kernel (``cvd_epi_attn_fwd_sharded``).

    ``k_srcs`` / ``vt_srcs``: one tensor per source rank (``[rows_per_src, N, C]`` / ``[rows_per_src, H, DV, N]`` bf16:
    the symmetric-memory buffers of every rank seen through their peer mappings, own rank included) or plain device
    pointers; ``partner`` int32 holds GLOBAL rows (source = row // rows_per_src).  The caller has ordered the call
    after the peers' writes (symmetric-memory barrier)."""
    import ctypes
    R, N, C = q.shape
    d = C // heads
    n_src = len(k_srcs)
    assert 1 <= n_src <= 8 and len(vt_srcs) == n_src
    n_partners = partner.shape[0] // R
    assert partner.shape[0] == n_partners * R
    feat = int(round(N ** 0.5))

    def ptr(t, shape):
        if isinstance(t, torch.Tensor):
            assert tuple(t.shape) == shape, (tuple(t.shape), shape)
            return _check(t, torch.bfloat16, "k / vt source")
        return int(t)
    ka = (ctypes.c_void_p * n_src)(*[ptr(t, (rows_per_src, N, C)) for t in k_srcs])
    va = (ctypes.c_void_p * n_src)(*[ptr(t, (rows_per_src, heads, dv_rows(d), N)) for t in vt_srcs])
    out = torch.empty_like(q)
    _lib.call("cvd_epi_attn_fwd_sharded", _check(q, torch.bfloat16, "q"), ctypes.cast(ka, ctypes.c_void_p),
              ctypes.cast(va, ctypes.c_void_p), n_src, rows_per_src, None if rowp is None else _check(rowp, torch.float32, "rowp"),
              _check(partner, torch.int32, "partner"), out.data_ptr(), R, N, heads, d, n_partners, feat, float(F_mat_size),
              _stream())
    _count()
    return out


def epi_attn_fwd_lse(q, k, vt, rowp, partner, heads, F_mat_size):
    """:func:`epi_attn_fwd` that also returns lse fp32 [R,H,N] (base-2 log-sum-exp of every query row's logits)."""
    R, R_kv, N, d, n_partners, feat = _attn_args(q, k, vt, rowp, partner, heads)
    out = torch.empty_like(q)
    lse = torch.empty(R, heads, N, device=q.device, dtype=torch.float32)
    _lib.call("cvd_epi_attn_fwd_lse", _check(q, torch.bfloat16, "q"), _check(k, torch.bfloat16, "k"),
              _check(vt, torch.bfloat16, "vt"), None if rowp is None else _check(rowp, torch.float32, "rowp"),
              _check(partner, torch.int32, "partner"), out.data_ptr(), lse.data_ptr(), R, R_kv, N, heads, d, n_partners, feat,
              float(F_mat_size), _stream())
    _count()
    return out, lse


def epi_attn_bwd(q, k, v, rowp, partner, out, d_out, lse, heads, F_mat_size):
    """Backward of the attention core (single partner).  q [R,N,C], k / v [R_kv,N,C] (natural layout), out / d_out [R,N,C]
    bf16; lse from :func:`epi_attn_fwd_lse`.  Returns fp32 (dq [R,N,C], dk [R_kv,N,C], dv [R_kv,N,C])."""
    R, N, C = q.shape
    R_kv = k.shape[0]
    d = C // heads
    feat = int(round(N ** 0.5))
    assert partner.shape[0] == R, "the backward supports one partner per row"
    assert k.shape == v.shape == (R_kv, N, C) and out.shape == d_out.shape == q.shape and lse.shape == (R, heads, N)
    dq = torch.zeros(R, N, C, device=q.device, dtype=torch.float32)
    dk = torch.zeros(R_kv, N, C, device=q.device, dtype=torch.float32)
    dv = torch.zeros(R_kv, N, C, device=q.device, dtype=torch.float32)
    delta = torch.empty(R, heads, N, device=q.device, dtype=torch.float32)
    _lib.call("cvd_epi_attn_bwd", _check(q, torch.bfloat16, "q"), _check(k, torch.bfloat16, "k"), _check(v, torch.bfloat16, "v"),
              None if rowp is None else _check(rowp, torch.float32, "rowp"), _check(partner, torch.int32, "partner"),
              _check(out, torch.bfloat16, "out"), _check(d_out, torch.bfloat16, "d_out"), _check(lse, torch.float32, "lse"),
              delta.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), R, R_kv, N, heads, d, feat, float(F_mat_size),
              _stream())
    _count(2)
    return dq, dk, dv


def padded_vt(v, heads):
    """v bf16 [R,N,C] -> v^T per head with the ones row, bf16 [R,H,DV,N] (the forward kernels' V layout)."""
    R, N, C = v.shape
    d = C // heads
    vt = torch.zeros(R, heads, dv_rows(d), N, device=v.device, dtype=torch.bfloat16)
    vt[:, :, :d] = v.view(R, N, heads, d).permute(0, 2, 3, 1)
    vt[:, :, d] = 1.0
    return vt


class EpiAttnFunction(torch.autograd.Function):
    """Differentiable epipolar attention core: forward = the tcgen05 kernel (+lse), backward = ``cvd_epi_attn_bwd``.
    Inputs q, k, v bf16 [R,N,C] (k, v are indexed through ``partner``); the bias (rowp) has no gradient
    (epi_module.py:329,393).  First building block of SURVEY.md section 8 row f3; the drop-in layers stay inference-only."""

    @staticmethod
    def forward(ctx, q, k, v, rowp, partner, heads, F_mat_size):
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        out, lse = epi_attn_fwd_lse(q, k, padded_vt(v, heads), rowp, partner, heads, F_mat_size)
        ctx.save_for_backward(q, k, v, out, lse, partner)
        ctx.rowp, ctx.heads, ctx.S = rowp, heads, F_mat_size
        return out

    @staticmethod
    def backward(ctx, d_out):
        q, k, v, out, lse, partner = ctx.saved_tensors
        dq, dk, dv = epi_attn_bwd(q, k, v, ctx.rowp, partner, out, d_out.contiguous().to(torch.bfloat16), lse, ctx.heads, ctx.S)
        return dq.to(q.dtype), dk.to(k.dtype), dv.to(v.dtype), None, None, None, None


def epi_attn_fwd_simt(q, k, vt, rowp, partner, heads, F_mat_size, out_dtype=torch.float32):
    """CUDA-core cross-check of :func:`epi_attn_fwd` (tests only)."""
    R, R_kv, N, d, n_partners, feat = _attn_args(q, k, vt, rowp, partner, heads)
    out = torch.empty(q.shape, device=q.device, dtype=out_dtype)
    rp = None if rowp is None else _check(rowp, torch.float32, "rowp")
    _lib.call("cvd_epi_attn_fwd_simt", _check(q, torch.bfloat16, "q"), _check(k, torch.bfloat16, "k"),
              _check(vt, torch.bfloat16, "vt"), rp, _check(partner, torch.int32, "partner"), out.data_ptr(),
              _DT[out_dtype], R, R_kv, N, heads, d, n_partners, feat, float(F_mat_size), _stream())
    _count()
    return out


def groupnorm_tokens(x, gamma, beta, groups=32, eps=1e-6, layout=_lib.LAYOUT_BF_HW):
    """x [B,C,F,h,w] -> bf16 tokens ([B*F, h*w, C] or [B*h*w, F, C])."""
    B, C, Fr, h, w = x.shape
    HW = h * w
    shape = (B * Fr, HW, C) if layout == _lib.LAYOUT_BF_HW else (B * HW, Fr, C)
    y = torch.empty(shape, device=x.device, dtype=torch.bfloat16)
    stats = torch.empty(B * Fr * groups * 2, device=x.device, dtype=torch.float32)
    _lib.call("cvd_groupnorm_tokens_fwd", _check(x, None, "x"), _DT[x.dtype], _check(gamma, torch.float32, "gamma"),
              _check(beta, torch.float32, "beta"), y.data_ptr(), stats.data_ptr(), B, C, Fr, HW, groups, float(eps),
              layout, _stream())
    _count(2)
    return y


def layernorm(x, gamma, beta, eps=1e-5, pe=None, seq_len=1):
    """fp32 tokens [..., C] -> bf16, optional positional table pe[seq_len, C] added after the norm."""
    C = x.shape[-1]
    tokens = x.numel() // C
    y = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    _lib.call("cvd_layernorm_fwd", _check(x, torch.float32, "x"), _check(gamma, torch.float32, "gamma"),
              _check(beta, torch.float32, "beta"), None if pe is None else _check(pe, torch.float32, "pe"),
              y.data_ptr(), tokens, C, seq_len, float(eps), _stream())
    _count()
    return y


def residual_layernorm_(x, y, bias, scale, gamma, beta, eps=1e-5, pe=None, seq_len=1):
    """x (fp32, in place) += (y (bf16 | None) + bias (fp32 | None)) * scale; returns LayerNorm(x) (+pe) as bf16."""
    C = x.shape[-1]
    tokens = x.numel() // C
    n = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    _lib.call("cvd_residual_layernorm_fwd", _check(x, torch.float32, "x"),
              None if y is None else _check(y, torch.bfloat16, "y"),
              None if bias is None else _check(bias, torch.float32, "bias"), float(scale),
              _check(gamma, torch.float32, "gamma"), _check(beta, torch.float32, "beta"),
              None if pe is None else _check(pe, torch.float32, "pe"), n.data_ptr(), tokens, C, seq_len, float(eps), _stream())
    _count()
    return n


def geglu(hg):
    """bf16 [..., 2*inner] -> bf16 [..., inner]: h * gelu_erf(g)."""
    inner = hg.shape[-1] // 2
    tokens = hg.numel() // hg.shape[-1]
    out = torch.empty(*hg.shape[:-1], inner, device=hg.device, dtype=torch.bfloat16)
    _lib.call("cvd_geglu_fwd", _check(hg, torch.bfloat16, "hg"), out.data_ptr(), tokens, inner, _stream())
    _count()
    return out


def residual_add_(x, y, bias=None, scale=1.0, want_bf16=False):
    """x (fp32, in place) += (y (bf16) + bias) * scale; optionally also returns a bf16 copy."""
    C = x.shape[-1]
    tokens = x.numel() // C
    xb = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    _lib.call("cvd_residual_add_fwd", _check(x, torch.float32, "x"), _check(y, torch.bfloat16, "y"),
              None if bias is None else _check(bias, torch.float32, "bias"), float(scale),
              None if xb is None else xb.data_ptr(), tokens, C, _stream())
    _count()
    return xb


def tokens_to_video(tok, bias, residual, layout=_lib.LAYOUT_BF_HW):
    """bf16 tokens (+bias) -> [B,C,F,h,w] in residual's dtype, plus the residual."""
    B, C, Fr, h, w = residual.shape
    out = torch.empty_like(residual)
    _lib.call("cvd_tokens_to_video_fwd", _check(tok, torch.bfloat16, "tok"),
              None if bias is None else _check(bias, torch.float32, "bias"), _check(residual, None, "residual"),
              out.data_ptr(), _DT[residual.dtype], B, C, Fr, h * w, layout, _stream())
    _count()
    return out


def temporal_attn(q, k, v, heads):
    """bf16 [T, F, C] x3 -> bf16 [T, F, C]: per-pixel attention over the F frames."""
    T, Fr, C = q.shape
    out = torch.empty_like(q)
    _lib.call("cvd_temporal_attn_fwd", _check(q, torch.bfloat16, "q"), _check(k, torch.bfloat16, "k"),
              _check(v, torch.bfloat16, "v"), out.data_ptr(), T, Fr, heads, C // heads, _stream())
    _count()
    return out


def add_pose_tokens(x_tok, pose):
    """x_tok bf16 [(b h w), F, C] + pose [B,C,F,h,w] -> bf16 tokens."""
    B, C, Fr, h, w = pose.shape
    out = torch.empty_like(x_tok)
    _lib.call("cvd_add_pose_tokens_fwd", _check(x_tok, torch.bfloat16, "x_tok"), _check(pose, None, "pose"),
              _DT[pose.dtype], out.data_ptr(), B, C, Fr, h * w, _stream())
    _count()
    return out


def scale_add(a, b, scale):
    """bf16: a * scale + b."""
    out = torch.empty_like(a)
    _lib.call("cvd_scale_add_fwd", _check(a, torch.bfloat16, "a"), _check(b, torch.bfloat16, "b"), float(scale),
              out.data_ptr(), a.numel(), _stream())
    _count()
    return out


def peer_gather_rows(src_ptrs, partner, dst, rows_per_src):
    """dst[r] = rows of the peers' buffers (device pointers ``src_ptrs``, one per rank) selected by the
    GLOBAL row numbers ``partner`` (int32 device tensor).  ``dst``: contiguous [rows, ...] CUDA tensor."""
    import ctypes
    rows = partner.shape[0]
    row_bytes = dst.numel() // rows * dst.element_size()
    arr = (ctypes.c_void_p * len(src_ptrs))(*[int(p) for p in src_ptrs])
    _lib.call("cvd_peer_gather_rows", ctypes.cast(arr, ctypes.c_void_p), len(src_ptrs), _check(partner, torch.int32, "partner"),
              _check(dst, None, "dst"), rows, rows_per_src, row_bytes, _stream())
    _count()
    return dst


def pairwise_fundamental(c2w, K, partner):
    """Device version of the sampler's per-call F recompute: c2w fp64 [rows,4,4], K fp64 [rows,3,3],
    partner int32 [rows] (target row of every source row) -> F_mats fp32 [rows,3,3]."""
    rows = c2w.shape[0]
    F_mats = torch.empty(rows, 3, 3, device=c2w.device, dtype=torch.float32)
    _lib.call("cvd_pairwise_fundamental", _check(c2w, torch.float64, "c2w"), _check(K, torch.float64, "K"),
              _check(partner, torch.int32, "partner"), F_mats.data_ptr(), rows, _stream())
    _count()
    return F_mats


def plucker_embedding(intrinsics, c2w, H, W, flip_flag=None, channel_major=True):
    """Device version of ``ray_condition`` (dataset_validation.py:102-144): intrinsics fp32 [V,4] (fx, fy, cx, cy),
    c2w fp32 [V,4,4] -> fp32 [V,6,H,W] (``channel_major``, the dataset's layout) or [V,H,W,6] (ray_condition's)."""
    V = c2w.shape[0]
    out = torch.empty((V, 6, H, W) if channel_major else (V, H, W, 6), device=c2w.device, dtype=torch.float32)
    fp = None
    if flip_flag is not None:
        flip_flag = flip_flag.to(device=c2w.device, dtype=torch.uint8).contiguous()
        fp = _check(flip_flag, torch.uint8, "flip_flag")
    _lib.call("cvd_plucker_embedding", _check(intrinsics, torch.float32, "intrinsics"), _check(c2w, torch.float32, "c2w"),
              fp, out.data_ptr(), V, H, W, 1 if channel_major else 0, _stream())
    _count()
    return out

"""TEST DOUBLE -- a plain-torch emulation of every ``cvd_b200.ops`` entry point the inference paths call, so that the
HOST logic of the drop-in layers (argument plumbing, partner indices, branches, layouts, the order of residual updates,
RNG consumption) can run on the CPU build box, where there is no GPU to run ``libcvd_b200.so``.

It is installed by monkeypatching ``cvd_b200.ops`` inside a test (``install(monkeypatch)``); nothing under ``cvd_b200/``
imports it and the product never falls back to it.  Every function restates the CONTRACT of its kernel as documented in
``include/cvd_b200.h`` / ``cvd_b200/ops.py`` (fp32 math; the kernels' bf16 rounding points are reproduced where the output
dtype is bf16), not the kernel's implementation.  The kernels themselves are checked on the GPU (tests/test_gpu_*.py)."""
import math

import torch
import torch.nn.functional as F

BF = torch.bfloat16

# ``install(..., strict_shapes=True)``: additionally refuse what the DEVICE kernels refuse (instantiated attention
# configurations, LayerNorm widths, alignment rules of include/cvd_b200.h), so that a CPU test can prove that a model
# configuration is runnable on the B200 before the GPU suite sees it.
STRICT = {"on": False}
_ATTN_TC = {(40, 32), (40, 64), (80, 16), (80, 32), (160, 16), (40, 16), (160, 8), (80, 8), (40, 8)}
_ATTN_SMALL = {(160, 4), (80, 4), (40, 4)}


def _strict(cond, what):
    if STRICT["on"] and not cond:
        from cvd_b200._lib import CvdError            # what the real wrapper raises for a refused call
        raise CvdError(f"the device kernel would refuse this call: {what}")


def _pixel_centres(feat, S):
    step = S / feat
    idx = torch.arange(feat, dtype=torch.float32)
    x = step * idx + (step - 1) / 2
    return torch.stack([x.repeat(feat), x.repeat_interleave(feat), torch.ones(feat * feat)], dim=-1)     # [N,3], col fastest


def epi_row_params(F_mats, feat, F_mat_size, theta=None, first_frame_rule=1, pixel_band=3.0, decay_alpha=3.0, frames_per_video=16):
    M = F_mats.float().clone()
    if first_frame_rule != 0:
        if first_frame_rule == 2:
            c, s = math.cos(float(theta[0])), math.sin(float(theta[0]))
        else:
            c, s = 0.0, -1.0
        M[::frames_per_video] = torch.tensor([[0.0, 0.0, c], [0.0, 0.0, s], [-c, -s, 0.0]])
    X = _pixel_centres(feat, F_mat_size)
    L = torch.einsum("rij,nj->rni", M, X)
    nrm = (L[..., 0] ** 2 + L[..., 1] ** 2).sqrt() + 1e-6
    lo, hi = float(X[0, 0]), float(X[-1, 0])
    corners = torch.tensor([[lo, lo, 1.0], [hi, lo, 1.0], [lo, hi, 1.0], [hi, hi, 1.0]])
    dmax = (torch.einsum("rni,ci->rnc", L, corners).abs().amax(dim=-1) / nrm).amax(dim=-1)                 # [R]
    band = pixel_band / math.floor(F_mat_size * 0.5) * dmax
    rowp = torch.zeros(M.shape[0], 12)
    rowp[:, :9] = M.reshape(-1, 9)
    rowp[:, 9] = band
    rowp[:, 10] = decay_alpha / (band + 1e-6)
    bad = ~torch.isfinite(rowp).all(dim=1)
    rowp[bad] = 0.0
    return rowp


def epi_bias_dense(rowp, feat, F_mat_size):
    X = _pixel_centres(feat, F_mat_size)
    L = torch.einsum("rij,nj->rni", rowp[:, :9].reshape(-1, 3, 3), X)
    nrm = (L[..., 0] ** 2 + L[..., 1] ** 2).sqrt()[..., None] + 1e-6
    d = torch.einsum("rqi,ki->rqk", L, X).abs() / nrm
    b = -(d - rowp[:, 9, None, None]).clamp_min(0) * rowp[:, 10, None, None]
    return torch.nan_to_num(b, nan=0.0, posinf=0.0, neginf=0.0)


def dv_rows(head_dim):
    return (head_dim // 16 + 1) * 16


def _attend(q, k_all, vt_all, rowp, partner, heads, F_mat_size):
    R, N, C = q.shape
    d = C // heads
    kp = partner.shape[0] // R
    feat = int(round(N ** 0.5))
    _strict((d, feat) in _ATTN_TC or ((d, feat) in _ATTN_SMALL and kp in (1, 2)), f"epi attention (head_dim={d}, feat={feat}, partners={kp})")
    qh = q.float().view(R, N, heads, d).transpose(1, 2)
    ss, vs = [], []
    for p in range(kp):
        idx = partner[p * R:(p + 1) * R].long()
        kh = k_all.float()[idx].view(R, N, heads, d).transpose(1, 2)
        s = torch.matmul(qh, kh.transpose(-1, -2)) * d ** -0.5
        if rowp is not None:
            s = s + epi_bias_dense(rowp[p * R:(p + 1) * R], feat, F_mat_size)[:, None]
        ss.append(s)
        vs.append(vt_all.float()[idx][:, :, :d].transpose(-1, -2))                       # [R,H,N,d]
    pr = torch.softmax(torch.cat(ss, dim=-1), dim=-1)
    o = torch.matmul(pr, torch.cat(vs, dim=-2))
    return o.transpose(1, 2).reshape(R, N, C).contiguous()                 # fp32; the tensor-core kernels write bf16


def epi_attn_fwd(q, k, vt, rowp, partner, heads, F_mat_size, dbg_s=None):
    _strict((q.shape[2] // heads, int(round(q.shape[1] ** 0.5))) in _ATTN_TC | _ATTN_SMALL,
            f"epi attention (head_dim={q.shape[2] // heads}, feat={int(round(q.shape[1] ** 0.5))})")
    assert q.dtype == BF and k.dtype == BF and vt.dtype == BF and partner.dtype == torch.int32
    assert vt.shape == (k.shape[0], heads, dv_rows(q.shape[2] // heads), q.shape[1])
    assert bool((vt[:, :, q.shape[2] // heads] == 1).all()), "row d of every head of v^T must be ones"
    return _attend(q, k, vt, rowp, partner, heads, F_mat_size).to(BF)


def epi_attn_fwd_simt(q, k, vt, rowp, partner, heads, F_mat_size, out_dtype=torch.float32):
    STRICT_was, STRICT["on"] = STRICT["on"], False          # the CUDA-core cross-check kernel takes any (head_dim, feat)
    try:
        return _attend(q, k, vt, rowp, partner, heads, F_mat_size).to(out_dtype)
    finally:
        STRICT["on"] = STRICT_was


def peer_gather_rows(src_ptrs, partner, dst, rows_per_src):
    raise NotImplementedError("peer pulls need two GPUs: not part of the emulation")


def epi_attn_fwd_sharded(q, k_srcs, vt_srcs, rows_per_src, rowp, partner, heads, F_mat_size):
    assert all(t.shape[0] == rows_per_src for t in k_srcs)
    return _attend(q, torch.cat(list(k_srcs)), torch.cat(list(vt_srcs)), rowp, partner, heads, F_mat_size).to(BF)


def _logits_and_values(q, k_all, v_all, rowp, partner, heads, F_mat_size):
    """fp32 [R,H,N,M] logits (natural log units) and [R,H,M,d] values of the single-partner training path."""
    R, N, C = q.shape
    d = C // heads
    idx = partner.long()
    qh = q.float().view(R, N, heads, d).transpose(1, 2)
    kh = k_all.float()[idx].view(R, N, heads, d).transpose(1, 2)
    vh = v_all.float()[idx].view(R, N, heads, d).transpose(1, 2)
    s = torch.matmul(qh, kh.transpose(-1, -2)) * d ** -0.5
    if rowp is not None:
        s = s + epi_bias_dense(rowp, int(round(N ** 0.5)), F_mat_size)[:, None]
    return s, vh


def epi_attn_fwd_lse(q, k, vt, rowp, partner, heads, F_mat_size):
    out = epi_attn_fwd(q, k, vt, rowp, partner, heads, F_mat_size)
    d = q.shape[2] // heads
    v = vt[:, :, :d].transpose(-1, -2).reshape(vt.shape[0], heads, q.shape[1], d).transpose(1, 2).reshape(vt.shape[0], q.shape[1], -1)
    s, _ = _logits_and_values(q, k, v, rowp, partner, heads, F_mat_size)
    return out, (torch.logsumexp(s, dim=-1) * 1.4426950408889634).contiguous()        # base-2 log-sum-exp [R,H,N]


def epi_attn_bwd(q, k, v, rowp, partner, out, d_out, lse, heads, F_mat_size):
    _strict(q.shape[2] // heads in (40, 80, 160) and q.shape[1] % 16 == 0 and partner.shape[0] == q.shape[0],
            f"attention backward head_dim={q.shape[2] // heads} N={q.shape[1]}")
    """fp32 (dq, dk, dv) of out = softmax(q k[partner]^T / sqrt(d) + bias) v[partner]: torch autograd through the fp32
    restatement (the kernel recomputes P from lse; the contract is the gradient)."""
    R, N, C = q.shape
    qf, kf, vf = (t_.detach().float().requires_grad_(True) for t_ in (q, k, v))
    with torch.enable_grad():
        s, vh = _logits_and_values(qf, kf, vf, rowp, partner, heads, F_mat_size)
        o = torch.matmul(torch.softmax(s, dim=-1), vh).transpose(1, 2).reshape(R, N, C)
        o.backward(d_out.float())
    return qf.grad, kf.grad, vf.grad


def groupnorm_tokens(x, gamma, beta, groups=32, eps=1e-6, layout=0):
    B, C, Fr, h, w = x.shape
    _strict(C % groups == 0 and C % 2 == 0 and B * Fr <= 65535, f"GroupNorm tokens C={C} groups={groups} B*F={B * Fr}")
    xt = x.float().permute(0, 2, 1, 3, 4).reshape(B * Fr, C, h, w)
    g = F.group_norm(xt, groups, gamma, beta, eps).view(B, Fr, C, h * w)
    if layout == 0:                                                     # (b f) (h w) c
        return g.permute(0, 1, 3, 2).reshape(B * Fr, h * w, C).to(BF).contiguous()
    return g.permute(0, 3, 1, 2).reshape(B * h * w, Fr, C).to(BF).contiguous()      # (b h w) f c


def _ln(x, gamma, beta, eps, pe, seq_len):
    _strict(x.shape[-1] in (64, 320, 640, 1280), f"LayerNorm width {x.shape[-1]} (64, 320, 640, 1280)")
    n = F.layer_norm(x.float(), (x.shape[-1],), gamma, beta, eps)
    if pe is not None:
        n = (n.view(-1, seq_len, x.shape[-1]) + pe[:seq_len]).view(x.shape)
    return n.to(BF)


def layernorm(x, gamma, beta, eps=1e-5, pe=None, seq_len=1):
    return _ln(x, gamma, beta, eps, pe, seq_len)


def residual_layernorm_(x, y, bias, scale, gamma, beta, eps=1e-5, pe=None, seq_len=1):
    assert x.dtype == torch.float32 and x.is_contiguous()
    upd = torch.zeros_like(x) if y is None else y.float()
    if bias is not None:
        upd = upd + bias
    if y is not None or bias is not None:
        x.add_(upd * scale)
    return _ln(x, gamma, beta, eps, pe, seq_len)


def geglu(hg):
    _strict((hg.shape[-1] // 2) % 8 == 0, f"GEGLU inner width {hg.shape[-1] // 2} (multiple of 8)")
    h, g = hg.float().chunk(2, dim=-1)
    return (h * F.gelu(g)).to(BF).contiguous()


def residual_add_(x, y, bias=None, scale=1.0, want_bf16=False):
    assert x.dtype == torch.float32 and x.is_contiguous()
    x.add_((y.float() + (0 if bias is None else bias)) * scale)
    return x.to(BF) if want_bf16 else None


def tokens_to_video(tok, bias, residual, layout=0):
    B, C, Fr, h, w = residual.shape
    t = tok.float() + (0 if bias is None else bias)
    if layout == 0:
        v = t.view(B, Fr, h, w, C).permute(0, 4, 1, 2, 3)
    else:
        v = t.view(B, h, w, Fr, C).permute(0, 4, 3, 1, 2)
    return (v + residual.float()).to(residual.dtype).contiguous()


def temporal_attn(q, k, v, heads):
    T, Fr, C = q.shape
    d = C // heads
    _strict(Fr <= 32 and d % 2 == 0 and C % 8 == 0, f"temporal attention F={Fr} head_dim={d} C={C}")

    def hd(t):
        return t.float().view(T, Fr, heads, d).transpose(1, 2)
    o = F.scaled_dot_product_attention(hd(q), hd(k), hd(v))
    return o.transpose(1, 2).reshape(T, Fr, C).to(BF).contiguous()


def add_pose_tokens(x_tok, pose):
    B, C, Fr, h, w = pose.shape
    p = pose.float().permute(0, 3, 4, 2, 1).reshape(B * h * w, Fr, C)
    return (x_tok.float() + p).to(BF).contiguous()


def scale_add(a, b, scale):
    return (a.float() * scale + b.float()).to(BF)


def pairwise_fundamental(c2w, K, partner):
    from oracle import epi_oracle as O
    rows = c2w.shape[0]
    out = torch.zeros(rows, 3, 3)
    c, k = c2w.double().numpy(), K.double().numpy()
    for s in range(rows):
        t = int(partner[s])
        import numpy as np
        out[s] = torch.from_numpy(O.fundamental_matrix(np.linalg.inv(c[t]) @ c[s], k[s], k[t])).float()
    return out


def plucker_embedding(intrinsics, c2w, H, W, flip_flag=None, channel_major=True):
    from oracle import epi_oracle as O
    out = O.plucker_embedding(intrinsics.numpy(), c2w.numpy(), H, W, None if flip_flag is None else flip_flag.numpy())
    return out.permute(0, 3, 1, 2).contiguous() if channel_major else out


EMULATED = ["epi_row_params", "epi_bias_dense", "epi_attn_fwd", "epi_attn_fwd_sharded", "epi_attn_fwd_lse", "epi_attn_bwd",
            "epi_attn_fwd_simt", "peer_gather_rows", "groupnorm_tokens", "layernorm",
            "residual_layernorm_", "geglu", "residual_add_", "tokens_to_video", "temporal_attn", "add_pose_tokens", "scale_add",
            "pairwise_fundamental", "plucker_embedding"]


def install(monkeypatch, strict_shapes=False):
    """Replace the C-ABI wrappers of ``cvd_b200.ops`` by the emulations above for the duration of one test.  Also gives
    the CPU ``torch.mm`` the ``out_dtype`` keyword the product uses on CUDA (bf16 operands, fp32 result straight out of the
    GEMM): on the CPU it is emulated as an fp32 product of the bf16-rounded operands."""
    from cvd_b200 import ops
    g = globals()
    if hasattr(monkeypatch, "setitem"):
        monkeypatch.setitem(STRICT, "on", bool(strict_shapes))          # restored when the test ends
    else:
        STRICT["on"] = bool(strict_shapes)

    def counted(fn):
        def call(*a, **k):
            ops.LAUNCHES += 1                      # the real wrappers count their launches (bench.py: gpu_launches)
            return fn(*a, **k)
        return call
    for name in EMULATED:
        monkeypatch.setattr(ops, name, counted(g[name]))
    real_mm = torch.mm

    def mm(a, b, *, out_dtype=None, out=None):
        if out_dtype is None:
            return real_mm(a, b, out=out) if out is not None else real_mm(a, b)
        return real_mm(a.to(out_dtype), b.to(out_dtype))
    monkeypatch.setattr(torch, "mm", mm)

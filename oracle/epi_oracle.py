"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the CVD cross-video synchronisation hot path.

A plain fp32 (geometry: fp64 numpy) restatement of what the reference computes on the
path SURVEY.md section 8(a) lists.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module; the
product (``cvd_b200``) never does and has no CPU fallback.

Parity status: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4), and its arithmetic partly lives in ``diffusers==0.24.0``
(requirements.txt:2, absent here; restated in ``oracle/diffusers_shim.py``).  The oracle
is therefore pinned against outputs of the *reference itself* executed in the build
container (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``), see
``tests/test_oracle_golden.py``.

Every function cites the reference lines it follows (paths relative to /root/reference).
Modules are addressed functionally: weights come in as a ``state_dict`` with the
reference's parameter names so the same dict drives the
reference, the oracle and the CUDA path.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

FRAMES_PER_VIDEO = 16  # hard-coded by the reference: epi_module.py:316, attention_processor.py:631


# ----------------------------------------------------------------------------------------
# a11  pose -> fundamental matrices        animatediff/data/dataset_validation.py:28-64
# ----------------------------------------------------------------------------------------
def cross_product_matrix(v):
    """[v]x for a 3-vector (dataset_validation.py:28-39)."""
    x, y, z = (float(t) for t in v)
    return np.array([[0.0, -z, y], [z, 0.0, -x], [-y, x, 0.0]], dtype=np.float64)


def essential_matrix(T):
    """E = R [(-R^T t)]x for x2 = R x1 + t (dataset_validation.py:45-50)."""
    R = np.asarray(T, dtype=np.float64)[:3, :3]
    t = np.asarray(T, dtype=np.float64)[:3, 3]
    return R @ cross_product_matrix(-(R.T @ t))


def fundamental_matrix(T, K1, K2):
    """F = K2^-T E K1^-1 (dataset_validation.py:57-64)."""
    return np.linalg.inv(K2).T @ essential_matrix(T) @ np.linalg.inv(K1)


def read_pose_file(path):
    """RealEstate10K pose txt: URL line, then rows of 19 floats
    (id fx fy cx cy 0 0 w2c[3x4]).  Returns (intrinsics[n,4], c2w[n,4,4]) following
    ``Camera`` (dataset_validation.py:79-91), which inverts the stored 3x4."""
    with open(path, "r") as fh:
        rows = [[float(x) for x in ln.strip().split(" ")] for ln in fh.readlines()[1:] if ln.strip()]
    intr, c2w = [], []
    for e in rows:
        w2c = np.eye(4)
        w2c[:3, :] = np.array(e[7:]).reshape(3, 4)
        intr.append(e[1:5])
        c2w.append(np.linalg.inv(w2c))
    return np.array(intr), np.array(c2w)


def folded_pair_geometry(pose_file_0, pose_file_1, n_frames=16, sample_size=256):
    """c2w / K / F_mats of the 'folded' two-trajectory layout used by inference_epi.py.

    Follows ValRealEstate10KPoseFolded.load_cameras_specific + get_batch
    (dataset_validation.py:180-271), including its quirks: H,W = 1280,720 swapped
    (:203), trajectory 1 reversed (:194), K of trajectory 1 replaced by trajectory 0's
    (:238), relative poses rounded through float32 (:178).
    Returns F_mats float32 [2*n_frames,3,3] with F[n:] = F[:n]^T, c2w [2n-1,4,4], K [2n-1,3,3].
    """
    intr0, c2w0 = read_pose_file(pose_file_0)
    _, c2w1 = read_pose_file(pose_file_1)
    c2w1 = c2w1[::-1]
    H, W = 1280, 720
    crop = min(H, W)
    rescale = sample_size / crop
    dH, dW = (H - crop) / 2, (W - crop) / 2
    K0 = np.array([[[W * rescale * fx, 0, (W * cx - dW) * rescale],
                    [0, H * rescale * fy, (H * cy - dH) * rescale],
                    [0, 0, 1]] for fx, fy, cx, cy in intr0])

    def relative(c2w):
        inv0 = np.linalg.inv(c2w[0])
        return np.array([inv0 @ m for m in c2w], dtype=np.float32)

    r0, r1 = relative(c2w0), relative(c2w1)
    c2w = np.concatenate([r0[1:][::-1], r1], axis=0)
    K = np.concatenate([K0[1:][::-1], K0], axis=0)
    Fs = []
    for i in range(n_frames):
        sid, tid = n_frames - 1 - i, n_frames - 1 + i
        s2t = np.linalg.inv(c2w[tid]) @ c2w[sid]
        Fs.append(fundamental_matrix(s2t, K[sid], K[tid]))
    Fm = torch.as_tensor(np.array(Fs)).float()
    return torch.cat([Fm, Fm.permute(0, 2, 1)], dim=0).contiguous(), c2w, K


def pairwise_F_mats(c2w, K, partner, n_frames=FRAMES_PER_VIDEO):
    """Per-UNet-call F recompute of the V-view sampler
    (pipelines/pipeline_animation_epi_advanced.py:636-647).
    c2w [V*n,4,4], K [V*n,3,3] (numpy), partner[v] = partner view of view v.
    Row (v,f) gets F for source (v,f) -> target (partner[v], f).  Returns f32 [V*n,3,3]."""
    V = len(partner)
    out = np.zeros((V * n_frames, 3, 3), dtype=np.float64)
    for v in range(V):
        for f in range(n_frames):
            s, t = v * n_frames + f, int(partner[v]) * n_frames + f
            s2t = np.linalg.inv(c2w[t]) @ c2w[s]
            out[s] = fundamental_matrix(s2t, K[s], K[t])
    return torch.as_tensor(out).float()


# ----------------------------------------------------------------------------------------
# f2  Pluecker ray embedding          animatediff/data/dataset_validation.py:102-144
# ----------------------------------------------------------------------------------------
def plucker_embedding(intrinsics, c2w, H, W, flip_flag=None):
    """ray_condition (dataset_validation.py:102-144) for V cameras: intrinsics [V,4] = (fx, fy, cx, cy), c2w [V,4,4]
    -> fp32 [V,H,W,6] = (o x d, d), d = the world-space unit ray through each pixel centre (i + .5, j + .5);
    ``flip_flag`` [V] bool mirrors the pixel grid of those cameras horizontally (:117-126)."""
    K = np.asarray(intrinsics, dtype=np.float32)
    M = np.asarray(c2w, dtype=np.float32)
    V = M.shape[0]
    jj, ii = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    i = np.broadcast_to(ii.reshape(1, H * W), (V, H * W)).copy()
    if flip_flag is not None:
        flip = np.asarray(flip_flag, dtype=bool)
        i[flip] = (W - 1) - i[flip]
    i = i + np.float32(0.5)
    j = np.broadcast_to(jj.reshape(1, H * W), (V, H * W)) + np.float32(0.5)
    xs = (i - K[:, 2:3]) / K[:, 0:1]
    ys = (j - K[:, 3:4]) / K[:, 1:2]
    d = np.stack([xs, ys, np.ones_like(xs)], axis=-1)
    d = d / np.linalg.norm(d, axis=-1, keepdims=True)
    rays_d = d @ np.swapaxes(M[:, :3, :3], -1, -2)
    rays_o = np.broadcast_to(M[:, None, :3, 3], rays_d.shape)
    out = np.concatenate([np.cross(rays_o, rays_d), rays_d], axis=-1)
    return torch.from_numpy(out.reshape(V, H, W, 6).astype(np.float32))


# ----------------------------------------------------------------------------------------
# a1/a2  epipolar additive bias            animatediff/models/epi_module.py:293-346,398-407
# ----------------------------------------------------------------------------------------
def pixel_centres(feat, S):
    """Homogeneous pixel centres in S-pixel space, token n = row*feat + col
    (coords buffer epi_module.py:284-288, rescale :296-299)."""
    idx = torch.arange(feat, dtype=torch.float32)
    step = S / feat
    x = (step * idx + (step - 1) / 2)
    xs = x.repeat(feat)                      # col varies fastest
    ys = x.repeat_interleave(feat)
    return torch.stack([xs, ys, torch.ones_like(xs)], dim=-1)  # [N,3]


def effective_F(F_mats, theta=None):
    """Fold the first-frame rule into the matrices: rows r % 16 == 0 do not use F_r but a
    pseudo line through the query pixel (epi_module.py:316,331-346):
      theta given  -> (cos t, sin t, -(cos t x + sin t y))   [rand_slope_on_first_frame]
      theta None   -> (0, -1, y)                               [horizontal line]
    Both are linear in (x,y,1), i.e. a matrix M with l_q = M x_q."""
    Fe = F_mats.float().clone()
    if theta is None:
        M = torch.tensor([[0.0, 0.0, 0.0], [0.0, 0.0, -1.0], [0.0, 1.0, 0.0]])
    else:
        c, s = math.cos(float(theta)), math.sin(float(theta))
        c, s = float(np.float32(c)), float(np.float32(s))
        M = torch.tensor([[0.0, 0.0, c], [0.0, 0.0, s], [-c, -s, 0.0]])
    Fe[::FRAMES_PER_VIDEO] = M
    return Fe


def pseudo_line_matrices(slopes):
    """get_pseudo_F_coords(random_slope=True) (epi_module.py:331-339) as matrices: the line through the query pixel with
    direction angle ``slopes[r]`` is (cos, sin, -(cos x + sin y)) = M_r (x, y, 1)^T.  slopes [R] -> [R,3,3].
    This is the WHOLE bias geometry of the F_mats=None branch (:317-320: random pseudo lines on every row)."""
    c, s = torch.cos(slopes.float()), torch.sin(slopes.float())
    M = torch.zeros(slopes.shape[0], 3, 3)
    M[:, 0, 2], M[:, 1, 2], M[:, 2, 0], M[:, 2, 1] = c, s, -c, -s
    return M


def homography_line_matrices(H_mats, slopes, S=256):
    """The H_mats branch (epi_module.py:301-310): every query pixel is warped by the row's homography (about the image
    centre, perspective divide with +1e-6) and gets the pseudo line with angle ``slopes[r]`` through the WARPED point:
        x'' = (x - c0, y - c0, 1), c0 = (S - 1) / 2;  (U, V, W) = H x'' (+1e-6 on W);  (u, v) = (U / W + c0, V / W + c0)
        l_q = (cos, sin, -(cos u + sin v))
    Lines are homogeneous and the distance |l . x_k| / |(a, b)| does not change under l -> W l, and
        W l_q = (cos W, sin W, -(cos (U + c0 W) + sin (V + c0 W)))
    is LINEAR in x'', hence in (x, y, 1): l_q ~ M_r x_q with
        M_r = [cos h3'; sin h3'; -(cos h1 + sin h2 + (cos + sin) c0 h3')] T,   h3' = h3 + (0, 0, 1e-6),  T = shift by -c0.
    (The reference divides by |(a,b)| + 1e-6 = 1 + 1e-6 where this form divides by |W| + 1e-6: a relative difference of
    1e-6 (1 - 1/|W|), far below fp32 resolution for any usable homography.)  H_mats [R,3,3], slopes [R] -> [R,3,3]."""
    H = H_mats.float()
    c0 = (S - 1) / 2
    c, s = torch.cos(slopes.float())[:, None], torch.sin(slopes.float())[:, None]
    h1, h2 = H[:, 0], H[:, 1]
    h3 = H[:, 2].clone()
    h3[:, 2] += 1e-6
    rows = torch.stack([c * h3, s * h3, -(c * h1 + s * h2 + (c + s) * c0 * h3)], dim=1)       # acts on x''
    T = torch.tensor([[1.0, 0.0, -c0], [0.0, 1.0, -c0], [0.0, 0.0, 1.0]])
    return rows @ T


def epi_bias(F_mats, feat, S=256, pixel_band=3, decay_alpha=3, theta=None, first_frame=True):
    """Dense additive attention bias [R, N, N] fp32 (EpiEncoding.get_attn_map, F_mats branch,
    epi_module.py:311-329) followed by the NaN/inf scrub (epi_module.py:404-407).

      l_q   = F_r x_q ;  d[r,q,k] = |l_q . x_k| / (sqrt(a^2+b^2) + 1e-6)
      band  = pixel_band / (S // 2) * max_{q,k} d[r]
      bias  = -max(d - band, 0) * decay_alpha / (band + 1e-6)

    ``first_frame=False``: the matrices already ARE the effective line matrices of every row (H_mats / F_mats=None
    branches through ``homography_line_matrices`` / ``pseudo_line_matrices``): no first-frame override.
    """
    X = pixel_centres(feat, S)                                   # [N,3]
    Fe = effective_F(F_mats, theta) if first_frame else F_mats.float()   # [R,3,3]
    L = torch.einsum("rij,nj->rni", Fe, X)                       # [R,N,3]  lines per query
    nrm = (L[..., 0] ** 2 + L[..., 1] ** 2).sqrt()[..., None]
    d = torch.einsum("rqi,ki->rqk", L, X).abs() / (nrm + 1e-6)   # [R,N,N]
    band = (pixel_band / (S // 2) * d.reshape(d.shape[0], -1).max(dim=-1)[0])[:, None, None]
    bias = -(d - band).clip(0) * (decay_alpha / (band + 1e-6))
    return torch.nan_to_num(bias, nan=0.0, posinf=0.0, neginf=0.0)


def interleave_partners(bias, R):
    """[k*R,N,N] -> [R,N,N*k] with key index n*k + p (epi_module.py:398-402)."""
    if bias.shape[0] == R:
        return bias
    kR, N, _ = bias.shape
    return bias.reshape(-1, R, N, N).permute(1, 2, 3, 0).reshape(R, N, -1)


# ----------------------------------------------------------------------------------------
# a4/a5  EpiAttnProcessor                  animatediff/models/attention_processor.py:548-650
# ----------------------------------------------------------------------------------------
def partner_rows(R, kv_index=None):
    """Row r reads K/V of row partner[r]: swap halves (:575-576) or the given index (:578)."""
    if kv_index is None:
        return torch.cat([torch.arange(R // 2, R), torch.arange(0, R // 2)])
    return torch.as_tensor(np.asarray(kv_index)).long()


def epi_attention(x, sd, prefix, heads, bias=None, kv_index=None, fix_firstframe=False,
                  rescale_output_factor=1.0):
    """x [R,N,C] -> (out [R,N,C], {"query","key"}) -- EpiAttnProcessor.forward.

    q = to_q(x); k,v = to_k/to_v(x[partner]) (multi-partner: keys interleaved, :579-583);
    softmax(q k^T / sqrt(d) + bias) v per head (bias shared by the heads, :585-587,617);
    optional fix_firstframe (:629-635); to_out Linear+bias (:641); / rescale (:648)."""
    R, N, C = x.shape
    idx = partner_rows(R, kv_index)
    enc = x[idx]
    if idx.shape[0] != R:
        enc = enc.reshape(-1, R, N, C).permute(1, 2, 0, 3).reshape(R, -1, C)
    q = F.linear(x, sd[prefix + "to_q.weight"])
    k = F.linear(enc, sd[prefix + "to_k.weight"])
    v = F.linear(enc, sd[prefix + "to_v.weight"])
    d = C // heads
    qh = q.view(R, -1, heads, d).transpose(1, 2)
    kh = k.view(R, -1, heads, d).transpose(1, 2)
    vh = v.view(R, -1, heads, d).transpose(1, 2)
    s = torch.matmul(qh, kh.transpose(-1, -2)) * (d ** -0.5)
    if bias is not None:
        s = s + bias[:, None]
    o = torch.matmul(torch.softmax(s, dim=-1), vh)               # [R,H,N,d]
    if fix_firstframe:
        v_self = F.linear(x, sd[prefix + "to_v.weight"]).view(R, -1, heads, d).transpose(1, 2)
        vr = v_self.reshape(-1, 2, FRAMES_PER_VIDEO, heads, N, d)
        ff = vr[:, :, 0:1].mean(dim=0, keepdim=True).expand(vr.shape[0], -1, -1, -1, -1, -1)
        o = o.clone()
        o[::FRAMES_PER_VIDEO] = ff.reshape(-1, heads, N, d)
    o = o.transpose(1, 2).reshape(R, -1, C)
    o = F.linear(o, sd[prefix + "to_out.0.weight"], sd[prefix + "to_out.0.bias"])
    return o / rescale_output_factor, {"query": q, "key": k}


def epi_self_attention(x, sd, prefix, heads, F_mats, S=256, theta=None, use_mask=True,
                       fix_firstframe=False, rescale_output_factor=1.0):
    """EpiSelfAttention.forward (epi_module.py:386-462), F_mats branch.

    ``F_mats`` is a tensor [R,3,3] or a list [tensor[k*R,3,3], kv_index].  The reference's
    memory-relief branch (R > 200 or M > 2048, :424-451) is reproduced *including* its
    quirks: with a tensor F_mats every view attends to itself, rows beyond a multiple of
    the chunk size are dropped, aux is None."""
    R, N, C = x.shape
    feat = int(N ** 0.5)
    Fm, kv_index = (F_mats[0], F_mats[1]) if isinstance(F_mats, (list, tuple)) else (F_mats, None)
    bias = None
    if use_mask:
        bias = interleave_partners(epi_bias(Fm, feat, S, theta=theta), R)
    if bias is not None and (bias.shape[0] > 200 or bias.shape[-1] > 2048):
        chunk = 128 * 1024 // max(1024, bias.shape[-1])
        src = torch.arange(R) if kv_index is None else torch.as_tensor(np.asarray(kv_index)).long()
        outs = []
        for i in range(R // chunk):
            sl = slice(i * chunk, (i + 1) * chunk)
            outs.append(_attend_rows(x[sl], x[src][sl], sd, prefix, heads, bias[sl], rescale_output_factor))
        return torch.cat(outs, dim=0), None
    return epi_attention(x, sd, prefix, heads, bias, kv_index, fix_firstframe, rescale_output_factor)


def _attend_rows(xq, xkv, sd, prefix, heads, bias, rescale):
    R, N, C = xq.shape
    d = C // heads
    q = F.linear(xq, sd[prefix + "to_q.weight"]).view(R, -1, heads, d).transpose(1, 2)
    k = F.linear(xkv, sd[prefix + "to_k.weight"]).view(R, -1, heads, d).transpose(1, 2)
    v = F.linear(xkv, sd[prefix + "to_v.weight"]).view(R, -1, heads, d).transpose(1, 2)
    p = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) * d ** -0.5 + bias[:, None], dim=-1)
    o = torch.matmul(p, v).transpose(1, 2).reshape(R, -1, C)
    return F.linear(o, sd[prefix + "to_out.0.weight"], sd[prefix + "to_out.0.bias"]) / rescale


# ----------------------------------------------------------------------------------------
# a6-a8, a10  block / transformer / module    epi_module.py:84-89,157-188,251-269; resnet.py:30-40
# ----------------------------------------------------------------------------------------
def geglu_ff(x, sd, prefix):
    """FeedForward(geglu): Linear(C->8C) -> h * gelu_erf(g) -> Linear(4C->C)."""
    hg = F.linear(x, sd[prefix + "net.0.proj.weight"], sd[prefix + "net.0.proj.bias"])
    h, g = hg.chunk(2, dim=-1)
    return F.linear(h * F.gelu(g), sd[prefix + "net.2.weight"], sd[prefix + "net.2.bias"])


def inflated_group_norm(x, weight, bias, groups=32, eps=1e-6):
    """GroupNorm applied per frame of [B,C,F,h,w] (resnet.py:30-40)."""
    B, C, Fr, h, w = x.shape
    y = x.permute(0, 2, 1, 3, 4).reshape(B * Fr, C, h, w)
    y = F.group_norm(y, groups, weight, bias, eps)
    return y.reshape(B, Fr, C, h, w).permute(0, 2, 1, 3, 4)


def epi_transformer_block(x, sd, prefix, heads, F_mats, n_attn=2, S=256, thetas=None, **kw):
    """EpiTransformerBlock.forward (epi_module.py:251-269)."""
    aux = []
    for i in range(n_attn):
        nx = F.layer_norm(x, x.shape[-1:], sd[f"{prefix}norms.{i}.weight"], sd[f"{prefix}norms.{i}.bias"])
        res, a = epi_self_attention(nx, sd, f"{prefix}attention_blocks.{i}.", heads, F_mats, S,
                                    theta=None if thetas is None else thetas[i], **kw)
        x = x + res
        aux.append(a)
    nx = F.layer_norm(x, x.shape[-1:], sd[prefix + "ff_norm.weight"], sd[prefix + "ff_norm.bias"])
    return geglu_ff(nx, sd, prefix + "ff.") + x, aux


def epi_module(x, sd, heads, F_mats, S=256, thetas=None, n_blocks=1, n_attn=2, prefix="epi_transformer.", **kw):
    """EpiModule.forward == EpiTransformer3DModel.forward (epi_module.py:84-89,157-188).
    x [B,C,F,h,w]; F_mats [B,F,3,3] or [tensor[k*B,F,3,3], kv_index].
    ``thetas``: per-attention first-frame slopes in call order (None -> horizontal rule)."""
    B, C, Fr, h, w = x.shape
    res = x
    y = inflated_group_norm(x, sd[prefix + "norm.weight"], sd[prefix + "norm.bias"])
    y = y.permute(0, 2, 3, 4, 1).reshape(B * Fr, h * w, C)
    if isinstance(F_mats, (list, tuple)):
        Fm = [F_mats[0].reshape(-1, 3, 3), F_mats[1]]
    else:
        Fm = F_mats.reshape(-1, 3, 3)
    y = F.linear(y, sd[prefix + "proj_in.weight"], sd[prefix + "proj_in.bias"])
    aux = []
    for b in range(n_blocks):
        th = None if thetas is None else thetas[b * n_attn:(b + 1) * n_attn]
        y, a = epi_transformer_block(y, sd, f"{prefix}transformer_blocks.{b}.", heads, Fm, n_attn, S, th, **kw)
        aux += a
    y = F.linear(y, sd[prefix + "proj_out.weight"], sd[prefix + "proj_out.bias"])
    y = y.reshape(B, Fr, h, w, C).permute(0, 4, 1, 2, 3)
    return y + res, aux


# ----------------------------------------------------------------------------------------
# a9  temporal ("motion") transformer       motion_module.py:370-394,447-460,463-481,509-549
#     processors attention_processor.py:25-110 (SDPA) and :272-371 (CameraCtrl pose adaptor)
# ----------------------------------------------------------------------------------------
def sinusoidal_pe(d_model, max_len=32):
    """PositionalEncoding buffer (motion_module.py:472-476)."""
    pos = torch.arange(max_len).unsqueeze(1)
    div = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(max_len, d_model)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def temporal_attention(x, sd, prefix, heads, pose_feature=None, merge=None, pose_scale=1.0):
    """TemporalSelfAttention.forward + its processor.  x [(B h w), F, C] already LayerNorm-ed.
    x += PE (motion_module.py:515-516); if ``merge`` = (W,b) of ``qkv_merge`` and a pose
    feature is given: x = qkv_merge(x + pose) * scale + x (attention_processor.py:324-326);
    then plain self-attention over the F frames with 1/sqrt(d) scaling and to_out."""
    T, Fr, C = x.shape
    x = x + sinusoidal_pe(C)[:Fr]
    if merge is not None and pose_feature is not None:
        x = F.linear(x + pose_feature, merge[0], merge[1]) * pose_scale + x
    d = C // heads
    q = F.linear(x, sd[prefix + "to_q.weight"]).view(T, Fr, heads, d).transpose(1, 2)
    k = F.linear(x, sd[prefix + "to_k.weight"]).view(T, Fr, heads, d).transpose(1, 2)
    v = F.linear(x, sd[prefix + "to_v.weight"]).view(T, Fr, heads, d).transpose(1, 2)
    p = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) * d ** -0.5, dim=-1)
    o = torch.matmul(p, v).transpose(1, 2).reshape(T, Fr, C)
    return F.linear(o, sd[prefix + "to_out.0.weight"], sd[prefix + "to_out.0.bias"])


def temporal_module(x, sd, heads, pose_feature=None, merges=None, pose_scale=1.0, n_attn=2,
                    prefix="temporal_transformer."):
    """VanillaTemporalModule.forward (motion_module.py:245-250,370-394), one block.
    x [B,C,F,h,w]; pose_feature [B,C,F,h,w] or None; merges[i] = (W,b) or None per
    attention block (the config installs the pose adaptor on attention_blocks.0 only,
    configs/inference_config.yaml:48-56)."""
    B, C, Fr, h, w = x.shape
    res = x
    y = inflated_group_norm(x, sd[prefix + "norm.weight"], sd[prefix + "norm.bias"])
    y = y.permute(0, 3, 4, 2, 1).reshape(B * h * w, Fr, C)
    pf = None if pose_feature is None else pose_feature.permute(0, 3, 4, 2, 1).reshape(B * h * w, Fr, C)
    y = F.linear(y, sd[prefix + "proj_in.weight"], sd[prefix + "proj_in.bias"])
    bp = prefix + "transformer_blocks.0."
    for i in range(n_attn):
        ny = F.layer_norm(y, (C,), sd[f"{bp}norms.{i}.weight"], sd[f"{bp}norms.{i}.bias"])
        m = None if merges is None else merges[i]
        y = temporal_attention(ny, sd, f"{bp}attention_blocks.{i}.", heads, pf, m, pose_scale) + y
    ny = F.layer_norm(y, (C,), sd[bp + "ff_norm.weight"], sd[bp + "ff_norm.bias"])
    y = geglu_ff(ny, sd, bp + "ff.") + y
    y = F.linear(y, sd[prefix + "proj_out.weight"], sd[prefix + "proj_out.bias"])
    y = y.reshape(B, h, w, Fr, C).permute(0, 4, 3, 1, 2)
    return y + res

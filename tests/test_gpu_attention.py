"""-m gpu: the fused tcgen05 epipolar attention kernel through the C ABI.

Parity chain:  reference (fixtures) == CPU oracle (tests/test_oracle_golden.py)
               CPU oracle / fp32 torch-on-device restatement  ~  CUDA-core kernel   (fp32 math, tight)
               CUDA-core kernel / oracle                       ~  tcgen05 kernel    (bf16 operands)
Tolerances (stated, bf16 tensor-core path vs fp32): relative L2 <= 1e-2 on the attention output,
largest element error <= 8e-2 of the output RMS.
"""
import numpy as np
import pytest
import torch

from gpu_util import assert_close, attention_ref, make_vt, rel_l2, t

pytestmark = pytest.mark.gpu

REL_TC = 1e-2      # tcgen05 (bf16 P, ex2.approx) vs fp32 restatement
REL_SIMT = 2e-3    # CUDA-core fp32 kernel vs fp32 restatement (output stored as fp32)


@pytest.fixture(scope="module")
def dev():
    from gpu_util import test_device
    return test_device()


def make_case(dev, R, feat, heads, d, seed, golden, kp=1, pair=None, qk_gain=1.0, theta=None, S=256, use_bias=True):
    from cvd_b200 import _lib, ops
    import golden_inputs as gi
    N, C = feat * feat, heads * d
    g = torch.Generator(device="cpu").manual_seed(seed)
    q = (torch.randn(R, N, C, generator=g) * qk_gain).to(dev).to(torch.bfloat16)
    k = (torch.randn(R, N, C, generator=g) * qk_gain).to(dev).to(torch.bfloat16)
    v = torch.randn(R, N, C, generator=g).to(dev).to(torch.bfloat16)
    Fm = gi.tile_F(t(golden["geometry"]["F_mats"]), kp * R).to(dev)
    if pair == "views":
        partner = torch.from_numpy(gi.view_pair_index(R // 16).astype(np.int32)).to(dev)
    elif kp == 2:
        half = torch.cat([torch.arange(R // 2, R), torch.arange(0, R // 2)])
        partner = torch.cat([half, torch.arange(R)]).to(torch.int32).to(dev)
    else:
        partner = torch.cat([torch.arange(R // 2, R), torch.arange(0, R // 2)]).to(torch.int32).to(dev)
    rowp, bias = None, None
    if use_bias:
        rule = _lib.FF_HORIZONTAL if theta is None else _lib.FF_RANDOM_SLOPE
        th = None if theta is None else torch.tensor([theta], device=dev, dtype=torch.float32)
        rowp = ops.epi_row_params(Fm.contiguous(), feat, S, th, rule)
        bias = ops.epi_bias_dense(rowp, feat, S)
    return q, k, v, make_vt(v, heads), rowp, bias, partner


@pytest.mark.parametrize("R,feat,d,kp", [(32, 4, 40, 1), (32, 8, 40, 1), (32, 8, 160, 1), (32, 16, 80, 1), (32, 4, 40, 2), (16, 32, 40, 1)])
def test_simt_kernel_vs_fp32(golden, dev, R, feat, d, kp):
    from cvd_b200 import ops
    q, k, v, vt, rowp, bias, partner = make_case(dev, R, feat, 8, d, 11, golden, kp=kp)
    out = ops.epi_attn_fwd_simt(q, k, vt, rowp, partner, 8, 256.0)
    assert_close(out, attention_ref(q, k, v, 8, bias, partner), rel=REL_SIMT, what="simt")


def test_qk_accumulator_dump(golden, dev):
    """Bring-up check of the first MMA alone: raw S = Q K^T of key tile 0 (TMA + UMMA descriptors)."""
    from cvd_b200 import ops
    R, feat, d, H = 32, 32, 40, 8
    N = feat * feat
    q, k, v, vt, rowp, bias, partner = make_case(dev, R, feat, H, d, 5, golden)
    grid = R * H * (N // 128)
    dbg_all = torch.zeros(grid * 128 * 64 + 2 * grid * 32 * 8 * 2, device=dev, dtype=torch.float32)   # S dump + int64 stamps
    ops.epi_attn_fwd(q, k, vt, rowp, partner, H, 256.0, dbg_s=dbg_all)
    dbg = dbg_all[:grid * 128 * 64]
    torch.cuda.synchronize()
    qh = q.float().view(R, N, H, d).permute(0, 2, 1, 3)                       # [R,H,N,d]
    kh = k.float()[partner.long()].view(R, N, H, d).permute(0, 2, 1, 3)
    ref = torch.matmul(qh, kh[:, :, :64].transpose(-1, -2))                    # [R,H,N,64]
    got = dbg.view(R, H, N // 128, 128, 64).reshape(R, H, N, 64)
    assert_close(got, ref, rel=1e-5, max_abs=1e-3, what="S tile 0")


@pytest.mark.parametrize("R,feat,d,kp,pair,theta", [
    (32, 32, 40, 1, None, None),       # BASELINE config 1 shape
    (32, 32, 40, 1, None, 0.7),
    (64, 32, 40, 1, "views", None),    # 4 views paired by kv_index
    (32, 16, 80, 1, None, None),
    (32, 8, 160, 1, None, None),
    (32, 4, 160, 1, None, None),
    (32, 4, 40, 1, None, None),
    (32, 4, 80, 1, None, None),
    (32, 8, 80, 1, None, None),
    (32, 32, 40, 2, None, None),       # two partners, one softmax over 2N keys
    (32, 8, 160, 2, None, None),       # small-grid (warp-level) kernel, two partners
    (32, 4, 40, 2, None, 0.3),
    (32, 8, 40, 1, None, 1.1),
    (16, 64, 40, 1, None, None),       # 512x512 images
])
def test_tcgen05_kernel_vs_fp32(golden, dev, R, feat, d, kp, pair, theta):
    from cvd_b200 import ops
    q, k, v, vt, rowp, bias, partner = make_case(dev, R, feat, 8, d, 21, golden, kp=kp, pair=pair, theta=theta)
    out = ops.epi_attn_fwd(q, k, vt, rowp, partner, 8, 256.0)
    ref = attention_ref(q, k, v, 8, bias, partner)
    assert_close(out, ref, rel=REL_TC, what=f"tcgen05 R={R} feat={feat} d={d}")
    simt = ops.epi_attn_fwd_simt(q, k, vt, rowp, partner, 8, 256.0)
    assert rel_l2(out.float(), simt) < REL_TC


def test_tcgen05_large_logits_rescale(golden, dev):
    """Sharp softmax (|logit| ~ 30): exercises the lazy O-rescale path."""
    from cvd_b200 import ops
    q, k, v, vt, rowp, bias, partner = make_case(dev, 32, 32, 8, 40, 31, golden, qk_gain=2.5)
    out = ops.epi_attn_fwd(q, k, vt, rowp, partner, 8, 256.0)
    assert_close(out, attention_ref(q, k, v, 8, bias, partner), rel=2e-2, what="sharp")


@pytest.mark.parametrize("feat,d", [(32, 40), (16, 80), (8, 160)])
def test_tcgen05_extreme_logits_do_not_overflow(golden, dev, feat, d):
    """|logit| in the hundreds (2^x overflows fp32 beyond x = 128): the exponentials of a quarter are computed BEFORE the
    reference-maximum vote, so a quarter whose scores jump far above the reference yields inf speculatively -- the vote
    must then recompute it against the new reference, and nothing non-finite may survive."""
    from cvd_b200 import ops
    q, k, v, vt, rowp, bias, partner = make_case(dev, 32, feat, 8, d, 61, golden, qk_gain=14.0)
    out = ops.epi_attn_fwd(q, k, vt, rowp, partner, 8, 256.0)
    assert torch.isfinite(out.float()).all()
    assert_close(out, attention_ref(q, k, v, 8, bias, partner), rel=3e-2, max_abs=0.5, what="extreme logits")


def test_tcgen05_no_bias_matches_sdpa(golden, dev):
    """rowp = NULL (epi_no_attention_mask): plain cross attention == torch SDPA."""
    from cvd_b200 import ops
    q, k, v, vt, _, _, partner = make_case(dev, 32, 32, 8, 40, 41, golden, use_bias=False)
    out = ops.epi_attn_fwd(q, k, vt, None, partner, 8, 256.0)
    assert_close(out, attention_ref(q, k, v, 8, None, partner), rel=REL_TC, what="no-bias")


def test_properties_full_size(golden, dev):
    """Size-independent properties at the bench shape (R=64, N=1024, C=320):
    rows of P sum to 1 (constant V -> constant output), linearity in V, and independence of
    unrelated rows (changing the K/V of a row only changes the rows partnered with it)."""
    from cvd_b200 import ops
    R, feat, H, d = 64, 32, 8, 40
    q, k, v, vt, rowp, bias, partner = make_case(dev, R, feat, H, d, 51, golden)
    ones = torch.ones_like(v)
    o1 = ops.epi_attn_fwd(q, k, make_vt(ones, H), rowp, partner, H, 256.0).float()
    assert float((o1 - 1).abs().max()) < 1e-2
    v2 = torch.randn_like(v.float()).to(torch.bfloat16)
    oa = ops.epi_attn_fwd(q, k, vt, rowp, partner, H, 256.0).float()
    ob = ops.epi_attn_fwd(q, k, make_vt(v2, H), rowp, partner, H, 256.0).float()
    oab = ops.epi_attn_fwd(q, k, make_vt((v.float() + v2.float()).to(torch.bfloat16), H), rowp, partner, H, 256.0).float()
    assert rel_l2(oab, oa + ob) < 1.5e-2
    k3 = k.clone()
    k3[5] = torch.randn_like(k3[5].float()).to(torch.bfloat16)
    oc = ops.epi_attn_fwd(q, k3, vt, rowp, partner, H, 256.0).float()
    changed = (oc - oa).abs().flatten(1).max(dim=1)[0] > 0
    expect = partner.long().cpu() == 5
    assert torch.equal(changed.cpu(), expect)
    # determinism
    assert torch.equal(ops.epi_attn_fwd(q, k, vt, rowp, partner, H, 256.0), ops.epi_attn_fwd(q, k, vt, rowp, partner, H, 256.0))


def test_unsupported_shape_is_an_error(dev):
    from cvd_b200 import CvdError, ops
    q = torch.zeros(2, 16, 8 * 24, device=dev, dtype=torch.bfloat16)
    vt = torch.zeros(2, 8, 32, 16, device=dev, dtype=torch.bfloat16)
    partner = torch.tensor([1, 0], device=dev, dtype=torch.int32)
    with pytest.raises(CvdError):
        ops.epi_attn_fwd(q, q, vt, None, partner, 8, 256.0)


@pytest.mark.parametrize("theta", [None, 0.9])
def test_tcgen05_kernel_vs_oracle_bias_directly(golden, dev, theta):
    """The in-kernel bias against the ORACLE's dense bias with no product code in between: oracle.epi_bias (pinned to
    the reference's get_attn_map by tests/test_oracle_golden.py) -> fp32 softmax(q k^T / sqrt(d) + bias) v on the device
    -> the tcgen05 kernel, at the bench shape's latent (32x32, d = 40), both first-frame rules."""
    from cvd_b200 import ops
    from oracle import epi_oracle as O
    import golden_inputs as gi
    R, feat, d = 32, 32, 40
    q, k, v, vt, rowp, _bias_prod, partner = make_case(dev, R, feat, 8, d, 77, golden, theta=theta)
    Fm = gi.tile_F(t(golden["geometry"]["F_mats"]), R)
    bias = O.epi_bias(Fm, feat, 256, theta=theta)                      # CPU oracle, [R,N,N] fp32
    bias = torch.nan_to_num(bias, nan=0.0, posinf=0.0, neginf=0.0).to(dev)
    out = ops.epi_attn_fwd(q, k, vt, rowp, partner, 8, 256.0)
    assert_close(out, attention_ref(q, k, v, 8, bias, partner), rel=REL_TC, max_abs=0.08, what="tcgen05 vs oracle bias")

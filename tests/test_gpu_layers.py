"""-m gpu: the drop-in layers (EpiSelfAttention / EpiModule / VanillaTemporalModule) against the
reference fixtures (tests/golden/*.npz = outputs of the unmodified reference) and the CPU oracle.

Tolerance (stated): the product computes projections and attention with bf16 operands and fp32
accumulation while the reference runs fp32 end to end, so layer outputs are compared with a
relative L2 bound of 2e-2 and a largest-element bound of 0.16 x RMS.
"""
import numpy as np
import pytest
import torch

import golden_inputs as gi
from gpu_util import assert_close, t

pytestmark = pytest.mark.gpu
REL = 2e-2


@pytest.fixture(scope="module")
def dev():
    from gpu_util import test_device
    return test_device()


def build_attn(C, dev, sd, **kw):
    import cvd_b200
    attn = cvd_b200.EpiSelfAttention(attention_mode="Epi_Self", cross_attention_dim=None, query_dim=C, heads=8,
                                     dim_head=C // 8, epi_position_encoding=True,
                                     epi_position_encoding_feat_max_size=64, epi_position_encoding_F_mat_size=256, **kw)
    attn.set_processor(cvd_b200.EpiAttnProcessor(hidden_size=C))
    missing, unexpected = attn.load_state_dict(sd, strict=False)
    assert not unexpected and all("pos_encoder" in m for m in missing)
    return attn.to(dev).eval()


@pytest.mark.parametrize("case", [c for c in gi.ATTN_CASES if c["name"] != "chunk"], ids=lambda c: c["name"])
def test_epi_self_attention_vs_reference_fixture(golden, dev, case):
    g = golden["attention"]
    Fm = t(golden["geometry"]["F_mats"])
    name, C, feat, R = case["name"], case["C"], case["feat"], case["R"]
    attn = build_attn(C, dev, gi.attn_state_dict(C, case["seed"]), epi_rand_slope_ff=case.get("rand_slope", False),
                      epi_fix_firstframe=case.get("fix_firstframe", False))
    if case.get("rand_slope"):
        attn.pos_encoder.theta_override = torch.tensor([float(golden["bias"]["theta"])], device=dev)
    x = gi.hidden_tokens(R, feat * feat, C, case["seed"] + 1).to(dev)
    F_in = gi.attn_F_input(case, Fm)
    F_in = [F_in[0].to(dev), F_in[1]] if isinstance(F_in, list) else F_in.to(dev)
    with torch.no_grad():
        y, aux = attn(x, F_mats=F_in, pose_feature=None, scale=None)
    assert y.dtype == x.dtype and y.shape == x.shape
    assert_close(gi.stored(case, y), t(g[name + "_out"]), rel=REL, what=name)
    assert set(aux.keys()) == {"query", "key"} and aux["query"].shape == (R, feat * feat, C)
    assert abs(float(aux["query"].double().abs().sum()) / float(g[name + "_qsum"]) - 1) < 5e-3
    assert abs(float(aux["key"].double().abs().sum()) / float(g[name + "_ksum"]) - 1) < 5e-3
    assert_close(aux["key"][0, ::case.get("queries", 1)], t(g[name + "_key0"]), rel=1e-2, what=name + " key")


def test_chunk_branch_quirk_matches_oracle(golden, dev):
    """R = 208 > 200 with a tensor F_mats: the reference's memory-relief branch attends every view to
    ITSELF, returns 128 rows and aux=None (epi_module.py:424-451; pinned by the 'chunk' fixture)."""
    from oracle import epi_oracle as O
    Fm = gi.tile_F(t(golden["geometry"]["F_mats"]), 208)
    C, feat = 320, 4
    sd = gi.attn_state_dict(C, 700)
    x = gi.hidden_tokens(208, feat * feat, C, 701)
    ref, aux_ref = O.epi_self_attention(x, sd, "", 8, Fm, 256)
    attn = build_attn(C, dev, sd)
    with torch.no_grad():
        y, aux = attn(x.to(dev), F_mats=Fm.to(dev))
    assert aux is None and aux_ref is None and y.shape == ref.shape == (128, 16, C)
    assert_close(y, ref, rel=REL, what="chunk quirk")
    attn.reference_chunk_quirk = False
    with torch.no_grad():
        y2, aux2 = attn(x.to(dev), F_mats=Fm.to(dev))
    assert y2.shape[0] == 208 and aux2 is not None


def test_config1_layer_vs_oracle(golden, dev):
    """BASELINE config 1: R=32 (2 views x 16 frames), 32x32 latent, C=320, real pose-pair geometry."""
    from oracle import epi_oracle as O
    Fm = t(golden["geometry"]["F_mats"])
    C, feat, R = 320, 32, 32
    sd = gi.attn_state_dict(C, 900)
    x = gi.hidden_tokens(R, feat * feat, C, 901)
    ref, _ = O.epi_self_attention(x, sd, "", 8, Fm, 256, theta=0.9)
    attn = build_attn(C, dev, sd, epi_rand_slope_ff=True)
    attn.pos_encoder.theta_override = torch.tensor([0.9], device=dev)
    with torch.no_grad():
        y, _ = attn(x.to(dev), F_mats=Fm.to(dev))
    assert_close(y, ref, rel=REL, what="config 1")
    for dt in (torch.bfloat16, torch.float16):          # other activation dtypes keep their dtype
        with torch.no_grad():
            yd, _ = attn(x.to(dev).to(dt), F_mats=Fm.to(dev))
        assert yd.dtype == dt
        assert_close(yd, ref, rel=3e-2, what=f"config 1 {dt}")


def test_rng_consumption_matches_reference(golden, dev):
    """One torch.rand([1], device) per attention call when rand_slope is on, none otherwise
    (epi_module.py:316,335): the global CUDA generator advances exactly like the reference's."""
    Fm = t(golden["geometry"]["F_mats"]).to(dev)
    sd = gi.attn_state_dict(320, 100)
    x = gi.hidden_tokens(32, 16, 320, 101).to(dev)
    for rand, n_draws in ((True, 1), (False, 0)):
        attn = build_attn(320, dev, sd, epi_rand_slope_ff=rand)
        torch.manual_seed(77)
        with torch.no_grad():
            attn(x, F_mats=Fm)
        after = torch.rand([1], device=dev)
        torch.manual_seed(77)
        for _ in range(n_draws):
            torch.rand([1], device=dev)
        assert torch.equal(after, torch.rand([1], device=dev))


def build_module(cls, C, dev, sd, **kw):
    m = cls(C, num_attention_heads=8, num_transformer_block=1, **kw)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all("pos_encoder" in k for k in missing), (missing, unexpected)
    return m.to(dev).eval()


@pytest.mark.parametrize("case", gi.MODULE_CASES, ids=lambda c: c["name"])
def test_epi_module_vs_reference_fixture(golden, dev, case):
    import cvd_b200
    g = golden["module"]
    Fm = t(golden["geometry"]["F_mats"])
    name, C, feat, B = case["name"], case["C"], case["feat"], case["B"]
    m = build_module(cvd_b200.EpiModule, C, dev, gi.module_state_dict(C, case["seed"]),
                     attention_block_types=["Epi_Self", "Epi_Self"], epi_position_encoding=True,
                     epi_position_encoding_feat_max_size=64, epi_position_encoding_F_mat_size=256,
                     epi_rand_slope_ff=case.get("rand_slope", False), zero_initialize=True)
    if case.get("rand_slope"):
        for i, blk in enumerate(m.epi_transformer.transformer_blocks[0].attention_blocks):
            blk.pos_encoder.theta_override = torch.tensor([float(g[name + "_thetas"][i])], device=dev)
    x = gi.hidden_video(B, C, 16, feat, feat, case["seed"] + 1).to(dev)
    F_in = gi.module_F_input(case, Fm)
    F_in = [F_in[0].to(dev), F_in[1]] if isinstance(F_in, list) else F_in.to(dev)
    with torch.no_grad():
        y, aux = m(x, F_mats=F_in, temb=None, encoder_hidden_states=None, cross_attention_kwargs={"pose_feature": None})
    assert len(aux) == int(g[name + "_naux"])
    ref = t(g[name + "_out"])
    assert_close(y, ref, rel=REL, what=name)
    assert_close(y.cpu() - x.cpu(), ref - x.cpu(), rel=3e-2, what=name + " (module delta)")


def test_epi_module_zero_init_is_identity(golden, dev):
    """zero_initialize=True zeroes proj_out (epi_module.py:81-82): a fresh module is an exact identity."""
    import cvd_b200
    m = cvd_b200.get_epi_module(320, dict(num_attention_heads=8, num_transformer_block=1,
                                          attention_block_types=["Epi_Self", "Epi_Self"], epi_position_encoding=True,
                                          epi_position_encoding_feat_max_size=64, epi_position_encoding_F_mat_size=256,
                                          epi_rand_slope_ff=True, zero_initialize=True)).to(dev).eval()
    x = gi.hidden_video(2, 320, 16, 8, 8, 5).to(dev)
    Fm = t(golden["geometry"]["F_mats"]).reshape(2, 16, 3, 3).to(dev)
    with torch.no_grad():
        y, _ = m(x, F_mats=Fm)
    assert torch.equal(y, x)


def test_epi_module_res1_vs_oracle(golden, dev):
    """Resolution-1 module (C=320, 32x32 latent, 2 views): full EpiModule forward vs the CPU oracle."""
    from oracle import epi_oracle as O
    import cvd_b200
    C, feat, B = 320, 32, 2
    sd = gi.module_state_dict(C, 3100)
    x = gi.hidden_video(B, C, 16, feat, feat, 3101)
    Fm = t(golden["geometry"]["F_mats"]).reshape(2, 16, 3, 3)
    ref, _ = O.epi_module(x, sd, 8, Fm, 256)
    m = build_module(cvd_b200.EpiModule, C, dev, sd, attention_block_types=["Epi_Self", "Epi_Self"],
                     epi_position_encoding=True, epi_position_encoding_feat_max_size=64,
                     epi_position_encoding_F_mat_size=256, epi_rand_slope_ff=False, zero_initialize=True)
    with torch.no_grad():
        y, _ = m(x.to(dev), F_mats=Fm.to(dev))
    assert_close(y.cpu() - x, ref - x, rel=3e-2, what="res-1 module delta")


@pytest.mark.parametrize("case", gi.TEMPORAL_CASES, ids=lambda c: c["name"])
def test_temporal_module_vs_reference_fixture(golden, dev, case):
    import cvd_b200
    g = golden["temporal"]
    name, C, feat, B = case["name"], case["C"], case["feat"], case["B"]
    m = cvd_b200.VanillaTemporalModule(C, num_attention_heads=8, num_transformer_block=1,
                                       attention_block_types=["Temporal_Self", "Temporal_Self"],
                                       temporal_position_encoding=True, temporal_position_encoding_max_len=32,
                                       zero_initialize=False)
    blk = m.temporal_transformer.transformer_blocks[0]
    blk.attention_blocks[0].set_processor(cvd_b200.PoseAdaptorAttnProcessor(
        hidden_size=C, pose_feature_dim=C, query_condition=True, key_value_condition=True, scale=1.0))
    blk.attention_blocks[1].set_processor(cvd_b200.AttnProcessor(attn_type="temporal"))
    missing, unexpected = m.load_state_dict(gi.temporal_state_dict(C, case["seed"]), strict=False)
    assert not unexpected and all("pos_encoder" in k for k in missing), (missing, unexpected)
    m = m.to(dev).eval()
    x = gi.hidden_video(B, C, 16, feat, feat, case["seed"] + 1).to(dev)
    pose = gi.hidden_video(B, C, 16, feat, feat, case["seed"] + 2).to(dev) if case["pose"] else None
    with torch.no_grad():
        y = m(x, cross_attention_kwargs={"pose_feature": pose})
    ref = t(g[name + "_out"])
    assert_close(y, ref, rel=REL, what=name)
    assert_close(y.cpu() - x.cpu(), ref - x.cpu(), rel=3e-2, what=name + " (module delta)")


def test_temporal_module_train_twin_vs_oracle_autograd(golden, dev):
    """The motion module is frozen in train_epi_control.py (only the epi parameters train, :270-283) but the gradients of
    the epi modules in front of it flow THROUGH it (:589-630), so in train() mode it records an autograd graph
    (``forward_train``) instead of refusing.  Forward, d/dx and d/dpose against torch autograd through the fp32 CPU
    oracle (pinned to the reference's temporal fixtures) on the fixture's own state dict: module delta rel-L2 3e-2,
    gradients 5e-2 (bf16 GEMM operands vs fp32), and the twin equals the CUDA inference path to bf16 rounding."""
    import cvd_b200
    from gpu_util import rel_l2
    from oracle import epi_oracle as O
    case = gi.TEMPORAL_CASES[0]
    assert case["pose"]
    C, feat, B = case["C"], case["feat"], case["B"]
    sd = gi.temporal_state_dict(C, case["seed"])
    m = cvd_b200.VanillaTemporalModule(C, num_attention_heads=8, num_transformer_block=1,
                                       attention_block_types=["Temporal_Self", "Temporal_Self"],
                                       temporal_position_encoding=True, temporal_position_encoding_max_len=32,
                                       zero_initialize=False)
    blk = m.temporal_transformer.transformer_blocks[0]
    blk.attention_blocks[0].set_processor(cvd_b200.PoseAdaptorAttnProcessor(
        hidden_size=C, pose_feature_dim=C, query_condition=True, key_value_condition=True, scale=1.0))
    blk.attention_blocks[1].set_processor(cvd_b200.AttnProcessor(attn_type="temporal"))
    m.load_state_dict(sd, strict=False)
    m = m.to(dev)
    x = gi.hidden_video(B, C, 16, feat, feat, case["seed"] + 1)
    pose = gi.hidden_video(B, C, 16, feat, feat, case["seed"] + 2)
    wgt = gi.hidden_video(B, C, 16, feat, feat, case["seed"] + 3)
    # oracle, fp32 CPU autograd
    x_ref, p_ref = x.clone().requires_grad_(True), pose.clone().requires_grad_(True)
    mk = "temporal_transformer.transformer_blocks.0.attention_blocks.0.processor.qkv_merge."
    y_ref = O.temporal_module(x_ref, sd, 8, p_ref, merges=[(sd[mk + "weight"], sd[mk + "bias"]), None])
    (y_ref * wgt).sum().backward()
    # product
    with torch.no_grad():
        y_inf = m.eval()(x.to(dev), cross_attention_kwargs={"pose_feature": pose.to(dev)})
    xp, pp = x.to(dev).requires_grad_(True), pose.to(dev).requires_grad_(True)
    y = m.train()(xp, cross_attention_kwargs={"pose_feature": pp})
    assert y.shape == x.shape and y.dtype == x.dtype
    (y * wgt.to(dev)).sum().backward()
    e_same = rel_l2(y.detach().cpu() - x, y_inf.cpu() - x)
    e_fwd = rel_l2(y.detach().cpu() - x, y_ref.detach() - x)
    e_x, e_p = rel_l2(xp.grad.cpu(), x_ref.grad), rel_l2(pp.grad.cpu(), p_ref.grad)
    print(f"temporal train twin: vs inference path {e_same:.2e}, vs oracle {e_fwd:.2e}, d/dx {e_x:.2e}, d/dpose {e_p:.2e}")
    assert e_same < 1e-2 and e_fwd < 3e-2 and e_x < 5e-2 and e_p < 5e-2

"""-m gpu: tests written in round 2 AFTER GPU access for this repository was closed (a 2-GPU experiment with multicast
TMA from peer memory wedged the box): the kernels they exercise were validated on a B200 earlier in
the round (profiles/r02_summary.md), the host logic is pinned on the CPU (tests/test_host_logic.py,
tests/test_oracle_golden.py), but THESE test functions have not run on hardware yet -- the file name sorts last so that
everything validated runs before them.

* multi-source K / V^T (cvd_epi_attn_fwd_sharded) and the opt-in cluster-multicast kernels on LOCAL memory;
* the H_mats / F_mats=None branches of get_attn_map through the device kernels (dense bias and layer level).
"""
import numpy as np
import pytest
import torch

import golden_inputs as gi
from gpu_util import assert_close, attention_ref
from test_gpu_attention import REL_TC, make_case
from test_gpu_layers import REL, build_attn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from gpu_util import test_device
    return test_device()


@pytest.fixture(autouse=True)
def fp32_convolutions(monkeypatch):
    """The "SD blocks fp32" comparisons mean fp32: cuDNN's default TF32 convolutions are switched off for this file."""
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)


@pytest.mark.parametrize("R,feat,d,n_src,kp", [(32, 32, 40, 1, 1), (32, 32, 40, 2, 1), (32, 16, 80, 2, 1), (32, 16, 40, 4, 2),
                                               (16, 64, 40, 1, 1), (32, 8, 160, 2, 1), (32, 4, 160, 2, 1)])
def test_multi_source_and_cluster_multicast_kernels_equal_default(golden, dev, R, feat, d, n_src, kp, monkeypatch):
    """cvd_epi_attn_fwd_sharded on ONE GPU: K / V^T split into n_src buffers ("sources", here all local), partner rows
    numbered globally -- first with the default (unicast) kernels, then with CVD_ATTN_CLUSTER=1: for latents > 8x8 the
    query tiles of a (row, head) form a thread-block cluster and share every K / V^T tile through one multicast TMA
    load.  Both must be BIT-identical to the default kernel on the un-split buffers (same arithmetic, only the
    delivery of the tiles differs).  (Local memory only: multicast TMA from peer memory wedges the GPU, DESIGN.md.)"""
    from cvd_b200 import ops
    monkeypatch.delenv("CVD_ATTN_CLUSTER", raising=False)
    q, k, v, vt, rowp, bias, partner = make_case(dev, R, feat, 8, d, 31 + feat, golden, kp=kp)
    ref = ops.epi_attn_fwd(q, k, vt, rowp, partner, 8, 256.0)
    rps = R // n_src
    ks = [k[i * rps:(i + 1) * rps].clone() for i in range(n_src)]
    vs = [vt[i * rps:(i + 1) * rps].clone() for i in range(n_src)]
    out = ops.epi_attn_fwd_sharded(q, ks, vs, rps, rowp, partner, 8, 256.0)
    assert torch.equal(out, ref)
    monkeypatch.setenv("CVD_ATTN_CLUSTER", "1")
    out_mc = ops.epi_attn_fwd_sharded(q, ks, vs, rps, rowp, partner, 8, 256.0)
    out_mc1 = ops.epi_attn_fwd(q, k, vt, rowp, partner, 8, 256.0)
    torch.cuda.synchronize()
    monkeypatch.delenv("CVD_ATTN_CLUSTER")
    assert torch.equal(out_mc, ref) and torch.equal(out_mc1, ref)
    assert_close(out, attention_ref(q, k, v, 8, bias, partner), rel=REL_TC, max_abs=0.08, what="multi-source")


@pytest.mark.parametrize("feat", [8, 16])
@pytest.mark.parametrize("tag", ["H", "none"])
def test_pseudo_line_branches_vs_reference_fixture(dev, feat, tag):
    """get_attn_map's H_mats branch (epi_module.py:301-310: homography-warped pseudo lines, the WebVid batches of
    train_epi_control.py:532-545) and F_mats=None branch (:317-320: random pseudo lines on every row) on the device:
    effective line matrices (torch, on the GPU) -> cvd_epi_row_params (no first-frame rule) -> dense writer, against the
    REFERENCE's outputs for the same H matrices and slopes (tests/golden/bias_pseudo.npz)."""
    import os
    import cvd_b200
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "bias_pseudo.npz"))
    enc = cvd_b200.EpiEncoding(320, max_feat_size=64, F_mat_size=256).to(dev)
    enc.slope_override = torch.from_numpy(g["slopes"]).to(dev)
    H = torch.from_numpy(g["H"]).to(dev)
    x = torch.zeros(H.shape[0], feat * feat, 1, device=dev)
    b = (enc.get_attn_map(x, H_mats=H) if tag == "H" else enc.get_attn_map(x)).cpu()
    k = f"f{feat}_{tag}_"
    N = feat * feat
    np.testing.assert_allclose(b[:, 0].numpy(), g[k + "q0"], atol=3e-3, rtol=1e-4)
    np.testing.assert_allclose(b[:, N // 2 + 3].numpy(), g[k + "qmid"], atol=3e-3, rtol=1e-4)
    np.testing.assert_allclose(torch.diagonal(b, dim1=1, dim2=2).numpy(), g[k + "diag"], atol=3e-3, rtol=1e-4)
    assert abs(float(b.min()) - float(g[k + "min"])) < 2e-3
    if feat == 8:
        np.testing.assert_allclose(b.numpy(), g[k + "full"], atol=3e-3, rtol=1e-4)
    # RNG consumption of these branches: ONE torch.rand([rows], device) per call (epi_module.py:310,320)
    enc.slope_override = None
    torch.manual_seed(5)
    enc.row_params(x, None, H if tag == "H" else None)
    after = torch.rand([1], device=dev)
    torch.manual_seed(5)
    torch.rand([H.shape[0]], device=dev)
    assert torch.equal(after, torch.rand([1], device=dev))


@pytest.mark.parametrize("branch", ["H", "none"])
def test_layer_homography_and_random_line_branches(golden, dev, branch):
    """EpiSelfAttention.forward with H_mats (train_epi_control.py:532-545) or with neither F_mats nor H_mats
    (epi_module.py:317-320), inference and training mode, against the fp32 oracle layer with the reference-pinned dense
    bias of the same slopes injected."""
    import os
    from gpu_util import rel_l2
    from oracle import epi_oracle as O
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "bias_pseudo.npz"))
    C, feat, R = 320, 8, 32
    sd = gi.attn_state_dict(C, 321)
    x = gi.hidden_tokens(R, feat * feat, C, 322)
    H = torch.from_numpy(g["H"])
    slopes = torch.from_numpy(g["slopes"])
    M = O.homography_line_matrices(H, slopes) if branch == "H" else O.pseudo_line_matrices(slopes)
    bias = O.epi_bias(M, feat, 256, first_frame=False)
    ref, _ = O.epi_attention(x, sd, "", 8, bias=bias)
    attn = build_attn(C, dev, sd)
    attn.pos_encoder.slope_override = slopes.to(dev)
    kw = dict(H_mats=H.to(dev)) if branch == "H" else {}
    with torch.no_grad():
        y, aux = attn(x.to(dev), **kw)
    assert_close(y, ref, rel=REL, what=f"{branch} branch")
    xt = x.to(dev).requires_grad_(True)
    yt, _ = attn.train()(xt, **kw)
    yt.float().pow(2).sum().backward()
    assert rel_l2(yt.detach(), y) < 1e-2 and torch.isfinite(xt.grad).all() and float(xt.grad.abs().max()) > 0


@pytest.mark.parametrize("case", [c for c in gi.POSE_ENCODER_CASES if c.get("device")], ids=lambda c: c["name"])
def test_pose_encoder_inference_path_vs_reference_fixture(dev, case):
    """cvd_b200.CameraPoseEncoder.forward on the device (cuDNN convolutions + this package's temporal transformer
    block kernels) against the reference's CameraPoseEncoder (tests/golden/pose_encoder.npz); the torch twin of the same
    module is pinned to the same fixture on the CPU (tests/test_host_logic.py)."""
    import os
    import cvd_b200
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "pose_encoder.npz"))
    n = case["name"]
    m = cvd_b200.CameraPoseEncoder(**case["kwargs"]).eval()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if not k.endswith("pos_encoder.pe")}
    m.load_state_dict(gi.fill_state_dict(shapes, case["seed"]), strict=False)
    m = m.to(dev)
    x = gi.plucker_like(case["B"], case["H"], case["W"], case["seed"] + 1).to(dev)
    with torch.no_grad():
        feats = m(x)
    assert len(feats) == int(g[n + "_n"])
    for i, f in enumerate(feats):
        assert_close(f, torch.from_numpy(g[f"{n}_feat{i}"]), rel=REL, what=f"{n} level {i}")


def _fixture_unet(dev, amp):
    """The product UNet at widths (320, 640): the smallest configuration whose sync / temporal modules have shapes the
    device kernels are instantiated for (head sizes 40 / 80 at 8x8 / 4x4 latents)."""
    from cvd_b200.unet import UNet3DConditionModelPoseCond
    m = UNet3DConditionModelPoseCond(**gi.UNET_GPU_KWARGS, amp=amp)
    m.set_all_attn_processor(add_spatial_lora=False, add_motion_lora=False, lora_kwargs={"lora_rank": 2, "lora_scale": 1.0},
                             motion_lora_kwargs={"lora_rank": -1, "lora_scale": 1.0},
                             sync_lora_kwargs={"sync_lora_rank": 0, "sync_lora_scale": 0}, spatial_extended_attention=False,
                             **gi.UNET_GPU_PROCESSOR_KWARGS)
    skip = ("pos_encoder.pe", "pos_encoder.coords")
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if not k.endswith(skip)}
    m.load_state_dict(gi.fill_state_dict(shapes, gi.UNET_GPU_SEED), strict=False)
    return m.to(dev).eval()


@pytest.mark.parametrize("amp", [False, True])
def test_unet_forward_vs_reference_unet_fixture(golden, dev, amp):
    """cvd_b200.unet.UNet3DConditionModelPoseCond on the device against the REFERENCE's UNet (models/unet.py +
    unet_blocks.py executed verbatim, tests/golden/unet_gpu.npz): SD blocks fp32 (2e-2) and under bf16 autocast (4e-2)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "unet_gpu.npz"))
    m = _fixture_unet(dev, amp)
    lat, text, poses = gi.unet_inputs(gi.UNET_GPU_SEED, channels=(320, 640), text_dim=64)
    F2 = torch.from_numpy(golden["geometry"]["F_mats"]).reshape(2, 16, 3, 3).to(dev)
    with torch.no_grad():
        for ts in (481, 21):
            out = m(lat.to(dev), ts, encoder_hidden_states=text.to(dev), F_mats=F2,
                    pose_embedding_features=[p.to(dev) for p in poses]).sample
            assert_close(out, torch.from_numpy(g[f"sample_t{ts}"]), rel=4e-2 if amp else REL, what=f"unet t={ts} amp={amp}")


def test_fixed_seed_ddim_final_latents_vs_reference_unet_on_device(golden, dev):
    """North star: final latents of a fixed-seed DDIM run with random-init weights -- sampler.sample_pair + the product
    UNet on the B200 (SD blocks fp32, sync / temporal kernels through the C ABI) against the oracle loop driving the
    REFERENCE UNet (fixture ``ddim_final`` of unet_gpu.npz: 4 steps, guidance 4.0).  Stated tolerance 5e-2."""
    import os
    from cvd_b200 import sampler
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "unet_gpu.npz"))
    m = _fixture_unet(dev, False)
    lat, _text, poses = gi.unet_inputs(gi.UNET_GPU_SEED, channels=(320, 640), text_dim=64)
    F2 = torch.from_numpy(golden["geometry"]["F_mats"]).reshape(2, 16, 3, 3).to(dev)
    with torch.no_grad():
        final = sampler.sample_pair(m, lat.to(dev), F2, torch.from_numpy(g["ddim_text"]).to(dev), sampler.DDIMSchedule(device=dev),
                                    4, 4.0, pose_embedding_features=[p.to(dev) for p in poses])
    assert_close(final, torch.from_numpy(g["ddim_final"]), rel=5e-2, what="ddim final latents")


def test_multiview_ddim_final_latents_vs_reference_unet_on_device(dev):
    """The V-view loop end to end on the B200 (sampler.sample_multiview + the product UNet: per-call pairing, device F
    recompute, kv_index gather, injected re-noise; 4 views, 3 DDIM steps, multistep 2) against the oracle loop driving the
    REFERENCE UNet at widths 320 / 640 (fixture ``mv_final`` of unet_gpu.npz).  Tolerance 5e-2; the reduced-width twin of
    this test runs with emulated kernels on the CPU (3.4e-3)."""
    import os
    from cvd_b200 import sampler
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "unet_gpu.npz"))
    m = _fixture_unet(dev, False)
    lat4, text4, poses4, noises = gi.unet_multiview_inputs(seed=gi.UNET_GPU_SEED + 50, channels=(320, 640), text_dim=64)
    c2w, K = sampler.multiview_cameras(4, 16, "circle", 0.5, 256)
    it = iter(noises)
    np.random.seed(2024)
    with torch.no_grad():
        final = sampler.sample_multiview(m, lat4.to(dev), text4.to(dev), sampler.DDIMSchedule(device=dev), 3, 4.0, aux_c2w=c2w,
                                         aux_K_mats=K, pose_embedding_features=[p.to(dev) for p in poses4], multistep=2,
                                         noise_fn=lambda shape: next(it).to(dev))
    assert_close(final, torch.from_numpy(g["mv_final"]), rel=5e-2, what="multiview ddim final latents")


def test_unet_training_step_vs_reference_training_step_on_device(golden, dev):
    """BASELINE config 5 at the UNet level on the B200: one optimisation step of cvd_b200.training (product UNet at widths
    320 / 640 + pose encoder; tcgen05 forward with lse, mma.sync backward, glue through torch autograd twins, AdamW) against
    the reference's own loop on the reference's modules (tests/golden/train_step_gpu.npz) -- loss, epsilon prediction,
    gradient norms per parameter, clipping, update.  The same case runs with emulated kernels on the CPU
    (tests/test_emulated_layers.py)."""
    from train_case import run_training_step_case
    run_training_step_case(dev, torch.from_numpy(golden["geometry"]["F_mats"]).reshape(2, 16, 3, 3), True, REL)

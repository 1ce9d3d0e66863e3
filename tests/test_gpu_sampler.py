"""-m gpu: the sampler loops driving the CUDA sync path against the CPU
oracle: fixed-seed DDIM runs with random-init weights on the hot-path skeleton UNet
(cvd_b200/sync_unet.py), per-call noise predictions and final latents."""
import os

import numpy as np
import pytest
import torch

from gpu_util import rel_l2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dev():
    from gpu_util import test_device
    return test_device()


def oracle_factories():
    """Factories for a CPU twin of SyncPathUNet: same parameter containers (product classes, never run), forward =
    the CPU oracle's functional restatement on this module's own state_dict (oracle/module_twins.py)."""
    from cvd_b200 import sync_unet
    from oracle import module_twins
    return module_twins.sync_unet_factories(), sync_unet


def build_pair(dev, levels, seed):
    """(product UNet on the GPU, oracle twin on the CPU) with identical random-init weights; epi proj_out and the
    pose adaptor's qkv_merge re-initialised non-zero; first-frame rule = horizontal (slope RNG is device-specific)."""
    kw, sync_unet = oracle_factories()
    ek = dict(sync_unet.EPI_KWARGS, epi_rand_slope_ff=False)
    torch.manual_seed(seed)
    prod = sync_unet.SyncPathUNet(levels=levels, epi_kwargs=ek).randomize_zero_init(seed)
    twin = sync_unet.SyncPathUNet(levels=levels, epi_kwargs=ek, **kw)
    missing, unexpected = twin.load_state_dict(prod.state_dict(), strict=False)
    assert not unexpected and all("pos_encoder" in k or "coords" in k for k in missing), (missing, unexpected)
    return prod.to(dev).eval(), twin.eval()


def test_multiview_ddim_latents_vs_oracle(dev):
    """4-view 'circle' rig, CFG, multistep 2: per-call pairing from numpy's global RNG, F recomputed on the device
    per call, kv_index gather, re-noise -- final latents and the first noise prediction against the CPU oracle."""
    from cvd_b200 import sampler
    from oracle import sampler_oracle as SO
    V, feat, steps, guidance = 4, 8, 3, 4.0
    prod, twin = build_pair(dev, levels=2, seed=7)
    c2w, K = sampler.multiview_cameras(V, 16, "circle", 0.5, 256)
    g = torch.Generator().manual_seed(42)
    lat = torch.randn(V, 4, 16, feat, feat, generator=g)
    text = (torch.randn(2, 77, 768, generator=g) * 0.5).repeat(V, 1, 1)
    poses = [torch.randn(V, c, 16, feat >> i, feat >> i, generator=g) * 0.5 for i, c in enumerate((320, 640))]
    noises = [torch.randn(V, 4, 16, feat, feat, generator=g) for _ in range(steps)]

    first = {}

    def tap(module, tag):
        def run(x, t, **kw):
            y = module(x, t, **kw)
            first.setdefault(tag, y.detach().float().cpu())
            return y
        return run

    it = iter(noises)
    np.random.seed(123)
    with torch.no_grad():
        ref = SO.sample_multiview(tap(twin, "ref"), lat.clone(), text, steps, guidance, c2w, K, poses, multistep=2,
                                  accumulate=1, noise_fn=lambda shape: next(it))
    it = iter(noises)
    np.random.seed(123)
    with torch.no_grad():
        out = sampler.sample_multiview(tap(prod, "out"), lat.to(dev), text.to(dev), sampler.DDIMSchedule(device=dev), steps,
                                       guidance, aux_c2w=c2w, aux_K_mats=K,
                                       pose_embedding_features=[p.to(dev) for p in poses], multistep=2,
                                       noise_fn=lambda shape: next(it).to(dev))
    torch.cuda.synchronize()
    r0 = rel_l2(first["out"], first["ref"])
    r = rel_l2(out.cpu(), ref)
    print(f"first noise_pred rel-L2 {r0:.3e}; final latents rel-L2 {r:.3e}")
    # tolerance: bf16 projections / attention against the fp32 oracle, amplified by guidance 4 over 5 UNet calls
    assert torch.isfinite(out).all() and r0 < 2e-2 and r < 5e-2, (r0, r)
    # the sync path must matter in this configuration (otherwise the comparison would be vacuous)
    with torch.no_grad():
        base = twin(lat.repeat_interleave(2, dim=0), 961, F_mats=None, encoder_hidden_states=text,
                    pose_embedding_features=[p.repeat_interleave(2, dim=0) for p in poses])
    assert rel_l2(base, first["ref"]) > 5e-2


def test_six_view_upper_hemi_accumulate_vs_oracle(dev):
    """BASELINE config 4 in miniature: 6 views on the 'upper_hemi' rig, accumulate_step 2 (two pairings averaged per
    step), multistep 2, one UNet level (4x4 latent: the warp-level attention kernel) -- final latents against the CPU oracle."""
    import cvd_b200
    from cvd_b200 import sampler
    from oracle import sampler_oracle as SO
    V, feat, steps, guidance = 6, 4, 2, 3.0
    prod, twin = build_pair(dev, levels=1, seed=13)
    c2w, K = sampler.multiview_cameras(V, 16, "upper_hemi", 0.5, 256)
    g = torch.Generator().manual_seed(4)
    lat = torch.randn(V, 4, 16, feat, feat, generator=g)
    text = (torch.randn(2, 77, 768, generator=g) * 0.5).repeat(V, 1, 1)
    poses = [torch.randn(V, 320, 16, feat, feat, generator=g) * 0.5]
    noises = [torch.randn(V, 4, 16, feat, feat, generator=g) for _ in range(steps)]
    it = iter(noises)
    np.random.seed(77)
    with torch.no_grad():
        ref = SO.sample_multiview(twin, lat.clone(), text, steps, guidance, c2w, K, poses, multistep=2, accumulate=2,
                                  noise_fn=lambda shape: next(it))
    it = iter(noises)
    np.random.seed(77)
    calls = {"n": 0}

    def unet(x, t, **kw):
        calls["n"] += 1
        return prod(x, t, **kw)

    with torch.no_grad():
        out = sampler.sample_multiview(unet, lat.to(dev), text.to(dev), sampler.DDIMSchedule(device=dev), steps, guidance,
                                       aux_c2w=c2w, aux_K_mats=K, pose_embedding_features=[p.to(dev) for p in poses],
                                       multistep=2, accumulate_step=2, noise_fn=lambda shape: next(it).to(dev))
    r = rel_l2(out.cpu(), ref)
    print(f"6-view upper_hemi, accumulate 2: final latents rel-L2 {r:.3e} after {calls['n']} UNet calls")
    assert calls["n"] == (2 + 1) * 2 and torch.isfinite(out).all() and r < 5e-2, r


def test_graphed_unet_equals_eager(dev):
    """One CUDA graph per UNet forward (what bench.py's denoise loop replays): same result as the eager modules,
    timestep as a device scalar so that every step reuses the graph."""
    import cvd_b200
    prod, _ = build_pair(dev, levels=2, seed=21)
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(4, 4, 16, 8, 8, device=dev, generator=g)
    text = torch.randn(4, 77, 768, device=dev, generator=g)
    Fm = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "geometry.npz"))["F_mats"]).reshape(2, 16, 3, 3)
    Fm = torch.cat([Fm[:1], Fm[:1], Fm[1:], Fm[1:]]).to(dev)
    poses = [torch.randn(4, c, 16, 8 >> i, 8 >> i, device=dev, generator=g) for i, c in enumerate((320, 640))]
    gnet = cvd_b200.GraphedModule(prod)
    with torch.no_grad():
        for t in (961.0, 41.0):
            tt = torch.full((), t, device=dev)
            y_e = prod(x, tt, F_mats=Fm, encoder_hidden_states=text, pose_embedding_features=poses)
            y_g = gnet(x, tt, F_mats=Fm, encoder_hidden_states=text, pose_embedding_features=poses).clone()
            assert torch.equal(y_e, y_g), float((y_e - y_g).abs().max())
    assert len(gnet._graphs) == 1


def test_graphed_multiview_loop_equals_eager(dev):
    """The V-view loop on a CUDA-graphed UNet with kv_index handed over as a device tensor (one graph for every pairing)
    gives the same latents as the eager loop with the reference's numpy kv_index (same pairing / re-noise draws)."""
    import cvd_b200
    from cvd_b200 import sampler
    prod, _ = build_pair(dev, levels=2, seed=23)
    V = 4
    g = torch.Generator(device=dev).manual_seed(3)
    lat = torch.randn(V, 4, 16, 8, 8, device=dev, generator=g)
    text = torch.randn(2, 77, 768, device=dev, generator=g).repeat(V, 1, 1)
    poses = [torch.randn(V, c, 16, 8 >> i, 8 >> i, device=dev, generator=g) for i, c in enumerate((320, 640))]
    c2w, K = sampler.multiview_cameras(V, 16, "circle", 0.5, 256)
    sch = sampler.DDIMSchedule(device=dev)
    gnet = cvd_b200.GraphedModule(prod)

    def graphed(x, t, **kw):
        return gnet(x, torch.full((), float(t), device=dev), **kw)

    outs = []
    for unet, on_dev in ((prod, False), (graphed, True)):
        np.random.seed(11)
        torch.manual_seed(11)
        with torch.no_grad():
            outs.append(sampler.sample_multiview(unet, lat, text, sch, 3, 4.0, aux_c2w=c2w, aux_K_mats=K,
                                                 pose_embedding_features=poses, multistep=2, kv_on_device=on_dev).clone())
    assert torch.isfinite(outs[0]).all()
    assert float((outs[0] - outs[1]).abs().max()) < 1e-5, float((outs[0] - outs[1]).abs().max())
    assert len(gnet._graphs) == 1          # five UNet calls, five pairings, one graph


def test_pair_ddim_latents_vs_oracle(dev):
    """Two-video loop (pipeline_animation_epi.py), real pose-pair geometry, swap-halves partner."""
    from cvd_b200 import sampler
    from oracle import sampler_oracle as SO
    feat, steps, guidance = 8, 3, 4.0
    prod, twin = build_pair(dev, levels=2, seed=9)
    Fm = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "geometry.npz"))["F_mats"]).reshape(2, 16, 3, 3)
    g = torch.Generator().manual_seed(42)
    lat = torch.randn(2, 4, 16, feat, feat, generator=g)
    text = torch.randn(2, 77, 768, generator=g) * 0.5
    with torch.no_grad():
        ref = SO.sample_pair(twin, lat.clone(), Fm, text, steps, guidance)
        out = sampler.sample_pair(prod, lat.to(dev), Fm.to(dev), text.to(dev), sampler.DDIMSchedule(device=dev), steps, guidance)
    r = rel_l2(out.cpu(), ref)
    print(f"two-video final latents rel-L2 {r:.3e}")
    assert torch.isfinite(out).all() and r < 5e-2, r


def test_device_F_recompute_matches_reference_fixture(dev):
    """cvd_pairwise_fundamental (f2) against F matrices produced by the reference's own per-frame loop."""
    from cvd_b200 import ops, sampler
    gold = np.load(os.path.join(ROOT, "tests", "golden", "sampler.npz"))
    for V in (4, 6):
        c2w = torch.from_numpy(gold[f"v{V}_c2w"]).double().to(dev)
        K = torch.from_numpy(gold[f"v{V}_K"]).double().to(dev)
        np.random.seed(1000 + V)
        for call in range(3):
            vid, _ = sampler.draw_view_pairing(V, 16)
            Fm = ops.pairwise_fundamental(c2w, K, torch.from_numpy(vid.astype(np.int32)).to(dev)).cpu()
            ref = torch.from_numpy(gold[f"v{V}_call{call}_F"]).reshape(-1, 3, 3)
            assert float((Fm - ref).abs().max() / ref.abs().max()) < 2e-5


def test_sync_unet_full_size_runs_and_counts_launches(dev):
    """Full four-level skeleton at the BASELINE config-2 shape (R = 64 rows, 32x32 latent): 40 epi-attention calls per
    forward, finite output, and the epi modules change the result."""
    from cvd_b200 import ops, sync_unet
    torch.manual_seed(0)
    net = sync_unet.SyncPathUNet().randomize_zero_init(1).to(dev).eval()
    Fm = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "geometry.npz"))["F_mats"]).reshape(2, 16, 3, 3)
    Fm = torch.cat([Fm[:1], Fm[:1], Fm[1:], Fm[1:]]).to(dev)
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(4, 4, 16, 32, 32, device=dev, generator=g)
    text = torch.randn(4, 77, 768, device=dev, generator=g)
    poses = [torch.randn(4, c, 16, 32 >> i, 32 >> i, device=dev, generator=g) for i, c in enumerate(sync_unet.CHANNELS)]
    calls = {"n": 0}
    orig = ops.epi_attn_fwd

    def counting(*a, **k):
        calls["n"] += 1
        return orig(*a, **k)

    ops.epi_attn_fwd = counting
    try:
        with torch.no_grad():
            y = net(x, 961, F_mats=Fm, encoder_hidden_states=text, pose_embedding_features=poses)
            y0 = net(x, 961, F_mats=None, encoder_hidden_states=text, pose_embedding_features=poses)
    finally:
        ops.epi_attn_fwd = orig
    assert calls["n"] == 40 and y.shape == x.shape and torch.isfinite(y).all()
    assert rel_l2(y, y0) > 1e-2

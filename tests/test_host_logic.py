"""CPU: host-side logic of the drop-in layers -- parameter names, partner-row resolution,
geometry, error behaviour.  No kernels run."""
import numpy as np
import pytest
import torch

import golden_inputs as gi


def test_state_dict_keys_match_the_reference():
    """the keys cvd.ckpt['unet_trainable_dict'] carries."""
    import cvd_b200
    m = cvd_b200.get_epi_module(320, dict(num_attention_heads=8, num_transformer_block=1,
                                          attention_block_types=["Epi_Self", "Epi_Self"],
                                          epi_position_encoding_feat_max_size=64))
    keys = set(m.state_dict().keys())
    blk = "epi_transformer.transformer_blocks.0."
    expect = {"epi_transformer.norm.weight", "epi_transformer.norm.bias", "epi_transformer.proj_in.weight",
              "epi_transformer.proj_in.bias", "epi_transformer.proj_out.weight", "epi_transformer.proj_out.bias",
              blk + "ff.net.0.proj.weight", blk + "ff.net.0.proj.bias", blk + "ff.net.2.weight", blk + "ff.net.2.bias",
              blk + "ff_norm.weight", blk + "ff_norm.bias"}
    for i in (0, 1):
        expect |= {f"{blk}attention_blocks.{i}.{n}" for n in
                   ("to_q.weight", "to_k.weight", "to_v.weight", "to_out.0.weight", "to_out.0.bias", "pos_encoder.coords")}
        expect |= {f"{blk}norms.{i}.weight", f"{blk}norms.{i}.bias"}
    assert keys == expect
    assert sum(p.numel() for p in m.parameters()) == 2259520
    assert m.state_dict()[blk + "attention_blocks.0.pos_encoder.coords"].shape == (64, 64, 3)
    assert float(m.epi_transformer.proj_out.weight.abs().sum()) == 0.0      # zero_initialize
    # the fixture state dicts (reference names) load with nothing unexpected
    missing, unexpected = m.load_state_dict(gi.module_state_dict(320, 1), strict=False)
    assert not unexpected and all("pos_encoder" in k for k in missing)


def test_temporal_state_dict_keys():
    import cvd_b200
    m = cvd_b200.get_motion_module(320, "Vanilla", dict(num_attention_heads=8, num_transformer_block=1,
                                                        attention_block_types=["Temporal_Self", "Temporal_Self"],
                                                        temporal_position_encoding=True,
                                                        temporal_position_encoding_max_len=32, zero_initialize=False))
    blk = m.temporal_transformer.transformer_blocks[0]
    blk.attention_blocks[0].set_processor(cvd_b200.PoseAdaptorAttnProcessor(
        hidden_size=320, pose_feature_dim=320, query_condition=True, key_value_condition=True))
    missing, unexpected = m.load_state_dict(gi.temporal_state_dict(320, 1), strict=False)
    assert not unexpected and all("pos_encoder" in k for k in missing)
    p = "temporal_transformer.transformer_blocks.0.attention_blocks.0."
    assert p + "processor.qkv_merge.weight" in m.state_dict() and p + "pos_encoder.pe" in m.state_dict()
    from oracle import epi_oracle as O
    assert torch.allclose(m.state_dict()[p + "pos_encoder.pe"][0], O.sinusoidal_pe(320, 32))


def test_partner_index_resolution():
    import cvd_b200
    proc = cvd_b200.EpiAttnProcessor(hidden_size=320)
    cpu = torch.device("cpu")
    assert proc.partner_index(8, None, cpu).tolist() == [4, 5, 6, 7, 0, 1, 2, 3]       # swap halves
    kv = gi.view_pair_index(4)                                                          # numpy int64, like the sampler's
    idx = proc.partner_index(64, kv, cpu)
    assert idx.dtype == torch.int32 and idx.tolist() == kv.tolist()
    assert proc.partner_index(64, kv, cpu) is idx                                       # cached
    assert proc.partner_index(64, torch.from_numpy(kv), cpu).tolist() == kv.tolist()
    multi = np.concatenate([kv, np.arange(64)])
    assert proc.partner_index(64, multi, cpu).shape[0] == 128
    with pytest.raises(AssertionError):
        proc.partner_index(64, np.arange(65), cpu)
    with pytest.raises(IndexError):
        proc.partner_index(64, np.arange(64) + 1, cpu)


def test_reference_error_behaviour_without_gpu():
    import cvd_b200
    attn = cvd_b200.EpiSelfAttention(attention_mode="Epi_Self", query_dim=320, heads=8, dim_head=40,
                                     epi_position_encoding=True, epi_mono_direction=True).eval()
    x = torch.zeros(32, 16, 320)
    with torch.no_grad():
        with pytest.raises(cvd_b200.CvdError):          # no CPU path: loud failure, not a fallback
            attn(x, F_mats=torch.zeros(32, 3, 3))
    with pytest.raises(AssertionError):
        cvd_b200.EpiSelfAttention(attention_mode="Temporal_Self", query_dim=320)
    with pytest.raises(ValueError):                     # non-square latent (epi_module.py:294)
        with torch.no_grad():
            attn(torch.zeros(32, 15, 320), F_mats=torch.zeros(32, 3, 3))


def test_geometry_matches_reference_fixture(golden):
    from cvd_b200 import geometry as G
    g = golden["geometry"]
    Fm = G.pairwise_fundamental_matrices(g["c2w"][0], g["K_mats"], [1, 0])
    np.testing.assert_allclose(Fm.numpy(), g["F_mats"], rtol=2e-4, atol=1e-9)
    # batched implementation == the oracle's frame-by-frame loop for a 4-view pairing
    from oracle import epi_oracle as O
    rng = np.random.RandomState(3)
    from scipy.spatial.transform import Rotation
    c2w = np.tile(np.eye(4), (64, 1, 1))
    c2w[:, :3, :3] = Rotation.from_rotvec(rng.randn(64, 3) * 0.2).as_matrix()
    c2w[:, :3, 3] = rng.randn(64, 3)
    K = np.tile(np.array([[223.578, 0, 128], [0, 223.578, 128], [0, 0, 1.0]]), (64, 1, 1))
    pairing = [2, 3, 0, 1]
    np.testing.assert_allclose(G.pairwise_fundamental_matrices(c2w, K, pairing).numpy(),
                               O.pairwise_F_mats(c2w, K, pairing).numpy(), rtol=1e-5, atol=1e-9)
    assert np.allclose(G.calc_cross_product_matrix(np.array([1.0, 2.0, 3.0])), O.cross_product_matrix([1, 2, 3]))


def test_training_path_has_no_cpu_fallback_either():
    """train() mode routes EpiModule through the differentiable path; like the inference
    path it refuses CPU tensors instead of falling back to torch, and the (frozen) temporal module refuses to build a
    graph at all."""
    import cvd_b200
    m = cvd_b200.EpiModule(320, num_attention_heads=8, num_transformer_block=1, attention_block_types=["Epi_Self"]).train()
    with pytest.raises(cvd_b200.CvdError):
        m(torch.zeros(2, 320, 16, 4, 4), F_mats=torch.zeros(2, 16, 3, 3))
    t = cvd_b200.VanillaTemporalModule(320, num_attention_heads=8, num_transformer_block=1,
                                       attention_block_types=["Temporal_Self"], temporal_position_encoding=True,
                                       temporal_position_encoding_max_len=32, zero_initialize=False).train()
    with pytest.raises((NotImplementedError, cvd_b200.CvdError)):
        t(torch.zeros(2, 320, 16, 4, 4), cross_attention_kwargs={"pose_feature": None})


def test_pseudo_line_branches_match_reference_fixture():
    """The H_mats (epi_module.py:301-310) and F_mats=None (:317-320) branches of get_attn_map reduce to effective per-row
    line matrices (cvd_b200.epi_module.homography_line_matrices / pseudo_line_matrices: plain torch, run here on the
    CPU); the dense bias of those matrices (oracle.epi_bias, no first-frame override) must equal what the REFERENCE
    computed for the same H matrices and the same slopes (tests/golden/bias_pseudo.npz, make_golden.py: bias_pseudo()),
    and the product consumes the torch global RNG like the reference: one rand([rows]) per call."""
    import math
    import os
    import numpy as np
    import torch
    from cvd_b200 import epi_module as em
    from oracle import epi_oracle as O
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "bias_pseudo.npz"))
    H = torch.from_numpy(g["H"])
    torch.manual_seed(int(g["seed"]))
    slopes = torch.rand([H.shape[0]]) * math.pi                 # the draw row_params makes (and the reference made)
    assert torch.equal(slopes, torch.from_numpy(g["slopes"]))
    for feat in (8, 16):
        for tag, M in (("H", em.homography_line_matrices(H, slopes, 256)), ("none", em.pseudo_line_matrices(slopes))):
            b = O.epi_bias(M, feat, 256, first_frame=False)
            k = f"f{feat}_{tag}_"
            assert float((b[:, 0, :] - torch.from_numpy(g[k + "q0"])).abs().max()) < 5e-4
            assert float((b[:, feat * feat // 2 + 3, :] - torch.from_numpy(g[k + "qmid"])).abs().max()) < 5e-4
            assert float((torch.diagonal(b, dim1=1, dim2=2) - torch.from_numpy(g[k + "diag"])).abs().max()) < 5e-4
            assert abs(float(b.double().sum()) / float(g[k + "sum"]) - 1) < 1e-6 and float(b.min()) == float(g[k + "min"])
            if feat == 8:
                assert float((b - torch.from_numpy(g[k + "full"])).abs().max()) < 5e-4
    # the oracle's own restatement of the two branches is the same function
    assert torch.allclose(em.homography_line_matrices(H, slopes, 256), O.homography_line_matrices(H, slopes, 256), atol=1e-6)
    assert torch.equal(em.pseudo_line_matrices(slopes), O.pseudo_line_matrices(slopes))


def test_pose_encoder_matches_reference_fixture():
    """cvd_b200.CameraPoseEncoder against the REFERENCE's CameraPoseEncoder (models/pose_adaptor.py:248-329, fixture
    tests/golden/pose_encoder.npz from make_golden.py: pose_encoder()): identical parameter names and shapes (the
    state_dict contract: a CameraCtrl checkpoint loads unchanged), and the feature maps of the autograd twin (plain
    torch ops, run here on the CPU by calling ``forward_train`` directly; bf16 GEMM operands -> rel-L2 2e-2).  The
    public ``forward`` refuses CPU tensors like the rest of the package."""
    import os
    import numpy as np
    import pytest
    import torch
    import cvd_b200
    import golden_inputs as gi
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "pose_encoder.npz"))
    for case in gi.POSE_ENCODER_CASES:
        n = case["name"]
        m = cvd_b200.CameraPoseEncoder(**case["kwargs"]).eval()
        mine = {k: tuple(v.shape) for k, v in m.state_dict().items() if not k.endswith("pos_encoder.pe")}
        assert sorted(mine) == g[n + "_names"].tolist()
        assert [str(mine[k]) for k in sorted(mine)] == g[n + "_shapes"].tolist()
        missing, unexpected = m.load_state_dict(gi.fill_state_dict(mine, case["seed"]), strict=False)
        assert not unexpected and all(k.endswith("pos_encoder.pe") for k in missing)
        x = gi.plucker_like(case["B"], case["H"], case["W"], case["seed"] + 1)
        with pytest.raises(cvd_b200.CvdError):
            m(x)
        with torch.no_grad():
            feats = m.forward_train(x)
        assert len(feats) == int(g[n + "_n"])
        for i, f in enumerate(feats):
            ref = torch.from_numpy(g[f"{n}_feat{i}"])
            assert f.shape == ref.shape
            err = float((f.double() - ref.double()).norm() / ref.double().norm())
            assert err < 2e-2, (n, i, err)


def test_folded_pose_conditioning_matches_reference_dataset(golden, monkeypatch):
    """geometry.folded_pose_conditioning (pose files -> F matrices, folded cameras, intrinsics) against what the REFERENCE
    dataset derived from the shipped pose pair (tests/golden/plucker.npz: load_cameras_specific's c2w / intrinsics;
    geometry.npz: F_mats and the checksums of the folded Pluecker embedding the pipeline consumes), with the Pluecker
    kernel replaced by its test double."""
    import os
    import numpy as np
    import torch
    import emulated_ops
    import golden_inputs as gi
    from cvd_b200 import geometry, ops
    emulated_ops.install(monkeypatch)
    ref = "/root/reference" if os.path.exists("/root/reference") else None
    if ref is None:
        import pytest
        pytest.skip("the pose files live in the reference checkout (absent on the GPU box)")
    cond = geometry.folded_pose_conditioning(os.path.join(ref, gi.POSE_FILE_0), os.path.join(ref, gi.POSE_FILE_1))
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "plucker.npz"))
    assert np.allclose(cond["c2w_all"], g["c2w"], atol=1e-6) and np.allclose(cond["intrinsics_all"], g["intrinsics"], rtol=1e-6)
    assert torch.allclose(cond["F_mats"], torch.from_numpy(golden["geometry"]["F_mats"]), rtol=1e-5, atol=1e-9)
    assert np.allclose(cond["c2w"].numpy()[None], golden["geometry"]["c2w"], atol=1e-6)
    pl = ops.plucker_embedding(cond["intrinsics"], cond["c2w"], 256, 256).double()          # [32, 6, 256, 256], folded
    assert abs(float(pl.sum()) / float(golden["geometry"]["plucker_sum"]) - 1) < 1e-5
    assert abs(float(pl.abs().sum()) / float(golden["geometry"]["plucker_abs_sum"]) - 1) < 1e-6


def test_gpu_local_cpu_binding_helpers(tmp_path, monkeypatch):
    """sharding.bind_host_thread_near_gpu: sysfs cpulist parsing, intersection with the allowed set, no-ops (unknown GPU,
    empty intersection, already inside) and the per-thread restore."""
    import os
    import types
    from cvd_b200 import sharding
    assert sharding.parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11] and sharding.parse_cpulist("") == []
    props = types.SimpleNamespace(pci_domain_id=0, pci_bus_id=0x1b, pci_device_id=0)
    monkeypatch.setattr(torch.cuda, "get_device_properties", lambda i: props)
    d = tmp_path / "bus" / "pci" / "devices" / "0000:1b:00.0"
    d.mkdir(parents=True)
    allowed = sorted(os.sched_getaffinity(0))
    (d / "local_cpulist").write_text(f"{allowed[0]}\n")
    assert sharding.gpu_local_cpus(0, str(tmp_path)) == [allowed[0]]
    prev = sharding.bind_host_thread_near_gpu(0, str(tmp_path))
    try:
        if len(allowed) > 1:
            assert prev == set(allowed) and os.sched_getaffinity(0) == {allowed[0]}
            assert sharding.bind_host_thread_near_gpu(0, str(tmp_path)) is None          # already inside: nothing to do
        else:
            assert prev is None
    finally:
        sharding.restore_host_affinity(prev)
    assert os.sched_getaffinity(0) == set(allowed)
    (d / "local_cpulist").write_text("100000-100001\n")                                  # no allowed CPU on that node
    assert sharding.bind_host_thread_near_gpu(0, str(tmp_path)) is None and os.sched_getaffinity(0) == set(allowed)
    props.pci_bus_id = 0x2c                                                              # unknown device: no sysfs entry
    assert sharding.gpu_local_cpus(0, str(tmp_path)) is None and sharding.bind_host_thread_near_gpu(0, str(tmp_path)) is None
    sharding.restore_host_affinity(None)


def test_training_batch_conditioning_matches_reference_dataset(tmp_path):
    """geometry.sample_clip_frames + folded_clip_conditioning against ONE batch of the reference's training dataset
    (RealEstate10KPoseFolded.get_batch of dataset_train_realestate10k.py executed verbatim, tests/golden/train_batch.npz):
    the same frames from the same ``random`` seed, F matrices (and their transposes for the way back), folded cameras,
    intrinsics of the centre-cropped non-square image, and the Pluecker rays (oracle restatement of ray_condition)."""
    import os
    import random
    from cvd_b200 import geometry
    from oracle import epi_oracle as O
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "train_batch.npz"))
    cfg = gi.TRAIN_BATCH
    pose = str(tmp_path / "clip.txt")
    gi.write_synthetic_pose_file(pose, cfg["n_cams"], cfg["seed"])
    random.seed(cfg["random_seed"])
    frames = geometry.sample_clip_frames(cfg["n_cams"], 16, cfg["sample_stride"])
    assert len(frames) == 31 and frames[0] >= 0 and frames[-1] < cfg["n_cams"]
    cond = geometry.folded_clip_conditioning(pose, frames, cfg["image_hw"])
    Fm, Fr = cond["F_mats"].numpy(), g["F_mats"]
    assert Fm.shape == Fr.shape == (32, 3, 3)
    scale = np.abs(Fr).max(axis=(1, 2), keepdims=True)
    assert np.abs(Fm - Fr).max() / scale.max() < 2e-4 and np.all(np.abs(Fm - Fr) <= 2e-4 * scale + 1e-7)
    np.testing.assert_allclose(Fm[16:], np.transpose(Fm[:16], (0, 2, 1)))
    np.testing.assert_allclose(cond["c2w"].numpy(), g["c2w"], atol=1e-6)
    np.testing.assert_allclose(cond["K"], g["K"], rtol=1e-6)
    pl = np.transpose(np.asarray(O.plucker_embedding(cond["intrinsics"], cond["c2w"], 256, 256)), (0, 3, 1, 2))   # [32,6,H,W]
    assert tuple(pl.shape) == tuple(g["plucker_shape"])
    np.testing.assert_allclose(pl[::3, :, ::17, ::19], g["plucker_sub"], atol=2e-5)
    # the same clip as 4 interpolated views (validation_video_split = 4, :466-478)
    mv = geometry.multiview_clip_conditioning(cond, 4)
    np.testing.assert_allclose(mv["c2w"].numpy(), g["mv_c2w"], atol=2e-6)
    np.testing.assert_allclose(mv["K"], g["mv_K"], rtol=1e-6)
    np.testing.assert_allclose(mv["c2w"][:16].numpy(), cond["c2w"][:16].numpy(), atol=2e-6)        # view 0 = video 0
    np.testing.assert_allclose(mv["c2w"][48:].numpy(), cond["c2w"][16:].numpy(), atol=2e-6)        # view 3 = video 1
    pl4 = np.transpose(np.asarray(O.plucker_embedding(mv["intrinsics"], mv["c2w"], 256, 256)), (0, 3, 1, 2))
    assert tuple(pl4.shape) == tuple(g["mv_plucker_shape"])
    np.testing.assert_allclose(pl4[::5, :, ::17, ::19], g["mv_plucker_sub"], atol=3e-5)


def test_fold_batch_matches_the_training_script_reshapes():
    """training.fold_batch against the statements of train_epi_control.py:518-545,577-582 written out with einops."""
    from einops import rearrange
    from cvd_b200 import training
    g = torch.Generator().manual_seed(3)
    batch = {"F_mats": torch.randn(2, 32, 3, 3, generator=g), "H_mats": torch.randn(2, 32, 3, 3, generator=g),
             "warped_masks": (torch.rand(2, 32, 256, 256, generator=g) > 0.02).double(),
             "plucker_embedding": torch.randn(2, 32, 6, 16, 24, generator=g)}
    lat = torch.zeros(1, dtype=torch.float32)
    out = training.fold_batch(batch, lat)
    F_ref = torch.cat(batch["F_mats"].chunk(2, dim=1), dim=0).type_as(lat)
    wm = torch.cat(batch["warped_masks"].chunk(2, dim=1), dim=0).type_as(lat)
    wm = rearrange(wm, "b f (h h1) (w w1) -> b f h w (h1 w1)", h1=8, w1=8).min(dim=-1)[0][:, None].float()
    pl = rearrange(torch.cat(batch["plucker_embedding"].chunk(2, dim=1), dim=0), "b f c h w -> b c f h w")
    assert torch.equal(out["F_mats"], F_ref) and out["F_mats"].shape == (4, 16, 3, 3)
    assert torch.equal(out["warped_masks"], wm) and wm.shape == (4, 1, 16, 32, 32) and 0 < float(wm.mean()) < 1
    assert torch.equal(out["plucker_embedding"], pl) and pl.shape == (4, 6, 16, 16, 24)
    only_F = training.fold_batch({"F_mats": batch["F_mats"]})
    assert only_F["H_mats"] is None and only_F["warped_masks"] is None and only_F["plucker_embedding"] is None


def test_unet_from_pretrained_2d_loads_a_stable_diffusion_folder(tmp_path):
    """UNet3DConditionModelPoseCond.from_pretrained_2d (unet.py:862-912): a Stable-Diffusion-style 2-D UNet folder
    (config.json with the 2-D block names and diffusers' bookkeeping keys + diffusion_pytorch_model.safetensors) becomes the
    3-D UNet with motion / epi modules; the 2-D weights land on the parameters of the same name, only motion / epi keys are
    missing, unexpected keys and missing files are errors."""
    import json
    import safetensors.torch
    from cvd_b200.unet import UNet3DConditionModelPoseCond
    extra = dict(use_motion_module=True, motion_module_resolutions=[1, 2, 4, 8], motion_module_mid_block=False,
                 motion_module_type="Vanilla", motion_module_kwargs=gi.UNET_MOTION_KWARGS, use_epi_module=True,
                 epi_module_resolutions=[1, 2, 4, 8], epi_module_mid_block=False, epi_module_kwargs=gi.UNET_EPI_KWARGS)
    folder = tmp_path / "sd" / "unet"
    folder.mkdir(parents=True)
    config = {"_class_name": "UNet2DConditionModel", "_diffusers_version": "0.6.0", "act_fn": "silu", "attention_head_dim": 8,
              "block_out_channels": [32, 64, 64, 64], "center_input_sample": False, "cross_attention_dim": 48,
              "down_block_types": ["CrossAttnDownBlock2D"] * 3 + ["DownBlock2D"], "downsample_padding": 1,
              "flip_sin_to_cos": True, "freq_shift": 0, "in_channels": 4, "layers_per_block": 1, "mid_block_scale_factor": 1,
              "norm_eps": 1e-05, "norm_num_groups": 8, "out_channels": 4, "sample_size": 8,
              "up_block_types": ["UpBlock2D"] + ["CrossAttnUpBlock2D"] * 3}
    (folder / "config.json").write_text(json.dumps(config))
    with pytest.raises(RuntimeError, match="safetensors does not exist"):
        UNet3DConditionModelPoseCond.from_pretrained_2d(str(tmp_path / "sd"), subfolder="unet", unet_additional_kwargs=extra)
    with pytest.raises(RuntimeError, match="config.json does not exist"):
        UNet3DConditionModelPoseCond.from_pretrained_2d(str(tmp_path / "nowhere"), unet_additional_kwargs=extra)
    # the 2-D checkpoint: every parameter of a plain (no motion / epi) UNet of that configuration
    plain = UNet3DConditionModelPoseCond(**{k: v for k, v in config.items() if not k.startswith("_") and "block_types" not in k})
    sd2d = {k: torch.randn(v.shape, generator=torch.Generator().manual_seed(i)) for i, (k, v) in enumerate(plain.state_dict().items())}
    safetensors.torch.save_file(sd2d, str(folder / "diffusion_pytorch_model.safetensors"))
    m = UNet3DConditionModelPoseCond.from_pretrained_2d(str(tmp_path / "sd"), subfolder="unet", unet_additional_kwargs=extra,
                                                        amp=False)
    assert m.amp is False and len(m.down_blocks) == 4 and m.config["cross_attention_dim"] == 48
    # the surface the reference's scripts / pipelines read (pipeline_animation_epi.py:91-110,512,549; inference_epi.py:93)
    assert m.config.sample_size == 8 and m.in_channels == 4 and m.device == torch.device("cpu") and m.dtype == torch.float32
    assert not hasattr(m.config, "_diffusers_version") and dict(m.config)["in_channels"] == 4
    sd3d = m.state_dict()
    assert all(torch.equal(sd3d[k], v) for k, v in sd2d.items())
    rest = sorted(set(sd3d) - set(sd2d))
    assert rest and all(("motion_modules." in k) or ("epi_modules." in k) for k in rest)
    safetensors.torch.save_file(dict(sd2d, stray=torch.zeros(1)), str(folder / "diffusion_pytorch_model.safetensors"))
    with pytest.raises(AssertionError):
        UNet3DConditionModelPoseCond.from_pretrained_2d(str(tmp_path / "sd"), subfolder="unet", unet_additional_kwargs=extra)
    # the training script's unconditional LoRA-scale calls (train_epi_control.py:208,582,642)
    m.set_image_layer_lora_scale(0.0)
    m.set_motion_module_lora_scale(0.5)
    assert all(b.lora_scale == 0.0 and b.motion_lora_scale == 0.5 for b in list(m.down_blocks) + list(m.up_blocks) + [m.mid_block])


def test_strict_state_dict_contract_including_buffers():
    """Parameters AND persistent buffers (positional tables) of the product UNet and pose encoder carry the reference's names
    and shapes (tests/golden/state_keys.npz), so ``load_state_dict(..., strict=True)`` of a reference checkpoint works
    (inference_epi.py:118-119); the sinusoidal table itself equals the reference's."""
    import os
    import cvd_b200
    from cvd_b200.unet import UNet3DConditionModelPoseCond
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "state_keys.npz"))
    sd = UNet3DConditionModelPoseCond(**gi.UNET_KWARGS).state_dict()
    assert sorted(sd) == g["unet_keys"].tolist()
    assert [str(tuple(sd[k].shape)) for k in sorted(sd)] == g["unet_shapes"].tolist()
    for case in gi.POSE_ENCODER_CASES:
        n = case["name"]
        sd = cvd_b200.CameraPoseEncoder(**case["kwargs"]).state_dict()
        assert sorted(sd) == g[n + "_keys"].tolist()
        assert [str(tuple(sd[k].shape)) for k in sorted(sd)] == g[n + "_shapes"].tolist()
        pe = [k for k in sorted(sd) if k.endswith(".pe")]
        if pe:
            np.testing.assert_allclose(sd[pe[0]].numpy(), g[n + "_pe0"], atol=1e-6)

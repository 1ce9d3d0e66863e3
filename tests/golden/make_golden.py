#!/usr/bin/env python
"""Generate tests/golden/*.npz by executing the UNMODIFIED reference in the build container.

    python tests/golden/make_golden.py            # needs /root/reference (absent on the GPU box)

The reference's hot-path files are imported verbatim from /root/reference behind
``oracle/diffusers_shim.py`` (diffusers 0.24.0 is pinned by the reference but is not
installable here).  Inputs and weights are NOT stored: they are regenerated from
``golden_inputs.py`` (numpy Generator streams, platform-stable), so the fixtures hold only
reference OUTPUTS plus a few input checksums.  The committed fixtures pin the oracle
(tests/test_oracle_golden.py, CPU) and the CUDA path (tests/test_gpu_*.py).
"""
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("CVD_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.path.insert(0, HERE)

from oracle import diffusers_shim  # noqa: E402

diffusers_shim.install()

import golden_inputs as gi  # noqa: E402
from animatediff.data.dataset_validation import ValRealEstate10KPoseFolded  # noqa: E402
from animatediff.models.attention_processor import (AttnProcessor, EpiAttnProcessor,  # noqa: E402
                                                    PoseAdaptorAttnProcessor)
from animatediff.models.epi_module import EpiEncoding, EpiModule, EpiSelfAttention  # noqa: E402
from animatediff.models.motion_module import VanillaTemporalModule  # noqa: E402

torch.set_grad_enabled(False)


def save(name, **arrays):
    out = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrays.items()}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"wrote {name}.npz  ({sum(a.nbytes for a in out.values()) / 1e6:.2f} MB raw)")


def geometry():
    ds = ValRealEstate10KPoseFolded(sample_n_frames=16, relative_pose=True, sample_size=256,
                                    validation_prompts=["x"],
                                    pose_file_0=os.path.join(REF, gi.POSE_FILE_0),
                                    pose_file_1=os.path.join(REF, gi.POSE_FILE_1))
    s = ds[0]
    pl = s["plucker_embedding"].double()
    save("geometry", F_mats=s["F_mats"], K_mats=s["ret_K_mats"], c2w=s["ret_c2w"],
         plucker_sum=pl.sum().item(), plucker_abs_sum=pl.abs().sum().item())
    return s["F_mats"]


def bias(Fm):
    out = {}
    for feat in (32, 16, 8, 4):
        N = feat * feat
        for tag, rand in (("h", False), ("t", True)):
            enc = EpiEncoding(320, max_feat_size=64, F_mat_size=256, rand_slope_on_first_frame=rand)
            torch.manual_seed(gi.THETA_SEED)
            b = enc.get_attn_map(torch.zeros(32, N, 1), Fm)
            torch.nan_to_num(b, nan=0.0, posinf=0.0, neginf=0.0, out=b)
            k = f"f{feat}{tag}_"
            out[k + "sum"] = b.double().sum().item()
            out[k + "min"] = b.min().item()
            out[k + "zeros"] = int((b == 0).sum())
            out[k + "rowsum"] = b.double().sum(dim=(1, 2))
            out[k + "q0"] = b[:, 0, :].clone()               # [32,N] query 0 of every row
            out[k + "qmid"] = b[:, N // 2 + 3, :].clone()
            out[k + "diag"] = torch.diagonal(b, dim1=1, dim2=2).clone()
            if feat <= 8:
                out[k + "full"] = b.clone()
    torch.manual_seed(gi.THETA_SEED)
    out["theta"] = (torch.rand([1]) * math.pi).item()
    # degenerate inputs: zero F on a non-first-frame row, non-finite F
    Fbad = Fm.clone()
    Fbad[3] = 0.0
    Fbad[5, 0, 0] = float("nan")
    Fbad[7, 1, 2] = float("inf")
    enc = EpiEncoding(320, max_feat_size=64, F_mat_size=256, rand_slope_on_first_frame=False)
    b = enc.get_attn_map(torch.zeros(32, 64, 1), Fbad)
    torch.nan_to_num(b, nan=0.0, posinf=0.0, neginf=0.0, out=b)
    out["bad_full"] = b
    # a different F_mat_size (image height 512 -> S=512) at feat 8
    enc = EpiEncoding(320, max_feat_size=64, F_mat_size=512, rand_slope_on_first_frame=False)
    out["s512_f8_full"] = enc.get_attn_map(torch.zeros(32, 64, 1), Fm)
    save("bias", **out)


def make_attn(C, **kw):
    attn = EpiSelfAttention(attention_mode="Epi_Self", cross_attention_dim=None, query_dim=C, heads=8,
                            dim_head=C // 8, dropout=0.0, bias=False, upcast_attention=False,
                            epi_position_encoding=True, epi_position_encoding_feat_max_size=64,
                            epi_position_encoding_F_mat_size=256, **kw)
    attn.set_processor(EpiAttnProcessor(hidden_size=C))
    return attn


def attention(Fm):
    out = {}
    for case in gi.ATTN_CASES:
        name, C, feat, R = case["name"], case["C"], case["feat"], case["R"]
        attn = make_attn(C, epi_rand_slope_ff=case.get("rand_slope", False),
                         epi_fix_firstframe=case.get("fix_firstframe", False))
        sd = gi.attn_state_dict(C, case["seed"])
        missing, unexpected = attn.load_state_dict(sd, strict=False)
        assert not unexpected and all("pos_encoder" in m for m in missing), (missing, unexpected)
        x = gi.hidden_tokens(R, feat * feat, C, case["seed"] + 1)
        F_in = gi.attn_F_input(case, Fm)
        torch.manual_seed(gi.THETA_SEED)
        y, aux = attn(x, F_mats=F_in, pose_feature=None, scale=None)
        out[name + "_out"] = gi.stored(case, y).clone()
        out[name + "_xsum"] = x.double().sum().item()
        if aux is not None:
            out[name + "_qsum"] = aux["query"].double().abs().sum().item()
            out[name + "_ksum"] = aux["key"].double().abs().sum().item()
            out[name + "_key0"] = aux["key"][0, ::case.get("queries", 1)].clone()
        else:
            out[name + "_aux_none"] = 1
    save("attention", **out)


def module(Fm):
    out = {}
    for case in gi.MODULE_CASES:
        name, C, feat, B = case["name"], case["C"], case["feat"], case["B"]
        m = EpiModule(C, num_attention_heads=8, num_transformer_block=1,
                      attention_block_types=["Epi_Self", "Epi_Self"], epi_position_encoding=True,
                      epi_position_encoding_feat_max_size=64, epi_position_encoding_F_mat_size=256,
                      epi_rand_slope_ff=case.get("rand_slope", False), zero_initialize=True)
        for mod in m.modules():
            if isinstance(mod, EpiSelfAttention):
                mod.set_processor(EpiAttnProcessor(hidden_size=C))
        sd = gi.module_state_dict(C, case["seed"])
        missing, unexpected = m.load_state_dict(sd, strict=False)
        assert not unexpected and all("pos_encoder" in k for k in missing), (missing, unexpected)
        x = gi.hidden_video(B, C, 16, feat, feat, case["seed"] + 1)
        F_in = gi.module_F_input(case, Fm)
        torch.manual_seed(gi.THETA_SEED)
        y, aux = m(x, F_mats=F_in)
        out[name + "_out"] = y
        out[name + "_naux"] = len(aux)
        if case.get("rand_slope", False):
            torch.manual_seed(gi.THETA_SEED)
            out[name + "_thetas"] = np.array([(torch.rand([1]) * math.pi).item() for _ in range(2)])
    save("module", **out)


def temporal():
    out = {}
    for case in gi.TEMPORAL_CASES:
        name, C, feat, B = case["name"], case["C"], case["feat"], case["B"]
        m = VanillaTemporalModule(C, num_attention_heads=8, num_transformer_block=1,
                                  attention_block_types=["Temporal_Self", "Temporal_Self"],
                                  temporal_position_encoding=True, temporal_position_encoding_max_len=32,
                                  zero_initialize=False)
        blk = m.temporal_transformer.transformer_blocks[0]
        p0 = PoseAdaptorAttnProcessor(hidden_size=C, pose_feature_dim=C, query_condition=True,
                                      key_value_condition=True, scale=1.0)
        blk.attention_blocks[0].set_processor(p0)
        blk.attention_blocks[1].set_processor(AttnProcessor(attn_type="temporal"))
        sd = gi.temporal_state_dict(C, case["seed"])
        missing, unexpected = m.load_state_dict(sd, strict=False)
        assert not unexpected and all("pos_encoder" in k for k in missing), (missing, unexpected)
        x = gi.hidden_video(B, C, 16, feat, feat, case["seed"] + 1)
        pose = gi.hidden_video(B, C, 16, feat, feat, case["seed"] + 2) if case["pose"] else None
        y = m(x, cross_attention_kwargs={"pose_feature": pose})
        out[name + "_out"] = y
    save("temporal", **out)


def _ref_lines(relpath, first, last):
    """Source lines [first, last] (1-based, inclusive) of a reference file, dedented -- executed, never stored."""
    import textwrap
    with open(os.path.join(REF, relpath)) as fh:
        lines = fh.readlines()[first - 1:last]
    return textwrap.dedent("".join(lines))


def sampler():
    """Pins the sampler-side logic of the advanced pipeline by EXECUTING the reference's own lines:
    camera rigs (inference_epi_advanced.py:110-128,297-345), the random view pairing and its CFG expansion
    (pipelines/pipeline_animation_epi_advanced.py:622-633,651-652) and the per-call F loop (:639-645)."""
    import math as _math
    import types
    from scipy.spatial.transform import Rotation as R, Slerp
    from animatediff.data.dataset_validation import calc_fundamental_matrix
    cam_fn_src = _ref_lines("inference_epi_advanced.py", 110, 128)
    cam_src = _ref_lines("inference_epi_advanced.py", 297, 345)
    pair_src = _ref_lines("animatediff/pipelines/pipeline_animation_epi_advanced.py", 622, 633)
    floop_src = _ref_lines("animatediff/pipelines/pipeline_animation_epi_advanced.py", 639, 645)
    cfg_src = _ref_lines("animatediff/pipelines/pipeline_animation_epi_advanced.py", 651, 652)
    out = {}
    for V, pattern in ((4, "circle"), (6, "upper_hemi")):
        ns = {"np": np, "math": _math, "R": R, "Slerp": Slerp}
        exec(cam_fn_src, ns)
        ns["args"] = types.SimpleNamespace(view_num=V, video_length=16, image_width=256, image_height=256,
                                           cam_pattern=pattern, camera_dist=0.5, cam_perturb_traj=0.0)
        ns["c2ws_list"] = []
        exec(cam_src, ns)
        c2w = ns["c2ws_list"].astype(np.float32)          # the pipeline receives float32 tensors (:351-353)
        Kf = ns["K_mats"].astype(np.float32)
        out[f"v{V}_c2w"], out[f"v{V}_K"] = c2w, Kf
        np.random.seed(1000 + V)
        for call in range(3):
            ps = {"np": np, "torch": torch, "video_split_num": V, "single_model_length": 16,
                  "calc_fundamental_matrix": calc_fundamental_matrix, "aux_c2w": c2w, "aux_K_mats": Kf}
            exec(pair_src, ps)
            exec(floop_src, ps)
            Fm = torch.cat(ps["F_mat_list"], dim=0).reshape(V, 16, 3, 3).float()
            exec(cfg_src, ps)
            out[f"v{V}_call{call}_video_id"] = ps["video_id"]
            out[f"v{V}_call{call}_video_id_input"] = ps["video_id_input"]
            out[f"v{V}_call{call}_F"] = Fm
    save("sampler", **out)


def bias_pseudo():
    """The two training-only branches of get_attn_map (epi_module.py:301-310 H_mats, :317-320 F_mats=None): the reference
    draws one slope per row from the torch global RNG; the fixture stores the seed, the H matrices and the reference's
    dense outputs (full at 8x8, sampled queries at 16x16)."""
    out = {}
    R = 32
    rs = np.random.RandomState(909)
    H = np.tile(np.eye(3), (R, 1, 1))
    H[:, :2, :2] += 0.08 * rs.standard_normal((R, 2, 2))
    H[:, :2, 2] += 12.0 * rs.standard_normal((R, 2))
    H[:, 2, :2] += 4e-4 * rs.standard_normal((R, 2))
    H = torch.from_numpy(H).float()
    out["H"] = H
    out["seed"] = 4321
    for feat in (8, 16):
        N = feat * feat
        enc = EpiEncoding(320, max_feat_size=64, F_mat_size=256, rand_slope_on_first_frame=False)
        for tag, kw in (("H", dict(H_mats=H)), ("none", dict())):
            torch.manual_seed(4321)
            b = enc.get_attn_map(torch.zeros(R, N, 1), **kw)
            torch.nan_to_num(b, nan=0.0, posinf=0.0, neginf=0.0, out=b)
            k = f"f{feat}_{tag}_"
            out[k + "sum"] = b.double().sum().item()
            out[k + "min"] = b.min().item()
            out[k + "q0"] = b[:, 0, :].clone()
            out[k + "qmid"] = b[:, N // 2 + 3, :].clone()
            out[k + "diag"] = torch.diagonal(b, dim1=1, dim2=2).clone()
            if feat == 8:
                out[k + "full"] = b.clone()
    torch.manual_seed(4321)
    out["slopes"] = torch.rand([R]) * math.pi
    save("bias_pseudo", **out)


def pose_encoder():
    """CameraPoseEncoder of the reference (models/pose_adaptor.py:248-329) on two small configurations; the fixture holds
    the parameter table (names + shapes: the state_dict contract) and the feature maps."""
    from animatediff.models.pose_adaptor import CameraPoseEncoder
    out = {}
    for case in gi.POSE_ENCODER_CASES:
        m = CameraPoseEncoder(**case["kwargs"]).eval()
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if not k.endswith("pos_encoder.pe")}
        sd = gi.fill_state_dict(shapes, case["seed"])
        missing, unexpected = m.load_state_dict(sd, strict=False)
        assert not unexpected and all(k.endswith("pos_encoder.pe") for k in missing), (missing, unexpected)
        x = gi.plucker_like(case["B"], case["H"], case["W"], case["seed"] + 1)
        feats = m(x)
        n = case["name"]
        out[n + "_names"] = np.array(sorted(shapes))
        out[n + "_shapes"] = np.array([str(shapes[k]) for k in sorted(shapes)])
        out[n + "_n"] = len(feats)
        for i, f in enumerate(feats):
            out[f"{n}_feat{i}"] = f
    save("pose_encoder", **out)


def unet(Fm):
    """The reference's UNet3DConditionModelPoseCond (models/unet.py, unet_blocks.py run verbatim on the restated diffusers
    blocks of oracle/diffusers_unet_shim.py) at a reduced width: parameter table, two forward passes, and the final latents
    of a 4-step two-video DDIM run (the oracle's restatement of pipeline_animation_epi.py:563-656 driving the reference
    UNet, guidance 4.0)."""
    from animatediff.models.unet import UNet3DConditionModelPoseCond
    from oracle import sampler_oracle as SO
    m = UNet3DConditionModelPoseCond(**gi.UNET_KWARGS)
    m.set_all_attn_processor(add_spatial_lora=False, add_motion_lora=False, lora_kwargs={"lora_rank": 2, "lora_scale": 1.0},
                             motion_lora_kwargs={"lora_rank": -1, "lora_scale": 1.0},
                             sync_lora_kwargs={"sync_lora_rank": 0, "sync_lora_scale": 0}, spatial_extended_attention=False,
                             **gi.UNET_PROCESSOR_KWARGS)
    m.eval()
    skip = ("pos_encoder.pe", "pos_encoder.coords")
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if not k.endswith(skip)}
    missing, unexpected = m.load_state_dict(gi.fill_state_dict(shapes, gi.UNET_SEED), strict=False)
    assert not unexpected and all(k.endswith(skip) for k in missing), (missing, unexpected)
    lat, text, poses = gi.unet_inputs()
    F2 = Fm.reshape(2, 16, 3, 3)
    out = {"names": np.array(sorted(shapes)), "shapes": np.array([str(shapes[k]) for k in sorted(shapes)])}
    for t in (481, 21):
        out[f"sample_t{t}"] = m(lat, t, encoder_hidden_states=text, F_mats=F2, pose_embedding_features=poses).sample
    # without F matrices the epi modules fall back to random pseudo lines (epi_module.py:317-320): one torch.rand([rows]) per
    # epi attention call from the global generator -- seeded here so that the product's RNG consumption can be checked
    torch.manual_seed(1234)
    out["sample_nopose_noF"] = m(lat, 481, encoder_hidden_states=text).sample

    def call(x, t, F_mats=None, encoder_hidden_states=None, pose_embedding_features=None):
        return m(x, t, encoder_hidden_states=encoder_hidden_states, F_mats=F_mats,
                 pose_embedding_features=pose_embedding_features).sample
    text2 = torch.cat([text * 0.0, text], dim=0)[[0, 2]]          # (uncond, cond) of one prompt, as the pipeline batches it
    out["ddim_text"] = text2
    out["ddim_final"] = SO.sample_pair(lambda x, t, **kw: call(x, t, kw.get("F_mats"), kw.get("encoder_hidden_states"),
                                                               kw.get("pose_embedding_features")),
                                       lat.clone(), F2, text2, 4, 4.0, pose_features=poses)
    # the V-view loop of the advanced pipeline (oracle restatement of pipeline_animation_epi_advanced.py:596-705) on the
    # reference UNet: 4 views on the 'circle' rig, 3 DDIM steps, multistep 2 (re-noise), guidance 4.0, numpy seed 2024
    from cvd_b200 import sampler as _s      # camera rig only (pinned to the reference's rig code by tests/golden/sampler.npz)
    c2w, K = _s.multiview_cameras(4, 16, "circle", 0.5, 256)
    lat4, text4, poses4, noises = gi.unet_multiview_inputs()
    it = iter(noises)
    np.random.seed(2024)
    out["mv_final"] = SO.sample_multiview(
        lambda x, t, **kw: call(x, t, kw.get("F_mats"), kw.get("encoder_hidden_states"), kw.get("pose_embedding_features")),
        lat4.clone(), text4, 3, 4.0, c2w, K, poses4, multistep=2, accumulate=1, noise_fn=lambda shape: next(it))
    save("unet", **out)


def unet_gpu(Fm):
    """The reference UNet at widths (320, 640) -- the shapes the device kernels are instantiated for -- for the GPU parity
    tests: two forward passes and the final latents of a 4-step two-video DDIM run (guidance 4.0)."""
    from animatediff.models.unet import UNet3DConditionModelPoseCond
    from oracle import sampler_oracle as SO
    m = UNet3DConditionModelPoseCond(**gi.UNET_GPU_KWARGS)
    m.set_all_attn_processor(add_spatial_lora=False, add_motion_lora=False, lora_kwargs={"lora_rank": 2, "lora_scale": 1.0},
                             motion_lora_kwargs={"lora_rank": -1, "lora_scale": 1.0},
                             sync_lora_kwargs={"sync_lora_rank": 0, "sync_lora_scale": 0}, spatial_extended_attention=False,
                             **gi.UNET_GPU_PROCESSOR_KWARGS)
    m.eval()
    skip = ("pos_encoder.pe", "pos_encoder.coords")
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if not k.endswith(skip)}
    missing, unexpected = m.load_state_dict(gi.fill_state_dict(shapes, gi.UNET_GPU_SEED), strict=False)
    assert not unexpected and all(k.endswith(skip) for k in missing), (missing, unexpected)
    lat, text, poses = gi.unet_inputs(gi.UNET_GPU_SEED, channels=(320, 640), text_dim=64)
    F2 = Fm.reshape(2, 16, 3, 3)
    out = {"n_params": sum(int(np.prod(v)) for v in shapes.values())}
    for t in (481, 21):
        out[f"sample_t{t}"] = m(lat, t, encoder_hidden_states=text, F_mats=F2, pose_embedding_features=poses).sample
    text2 = torch.cat([text * 0.0, text], dim=0)[[0, 2]]
    out["ddim_text"] = text2
    out["ddim_final"] = SO.sample_pair(
        lambda x, t, **kw: m(x, t, encoder_hidden_states=kw.get("encoder_hidden_states"), F_mats=kw.get("F_mats"),
                             pose_embedding_features=kw.get("pose_embedding_features")).sample,
        lat.clone(), F2, text2, 4, 4.0, pose_features=poses)
    from cvd_b200 import sampler as _s
    c2w, K = _s.multiview_cameras(4, 16, "circle", 0.5, 256)
    lat4, text4, poses4, noises = gi.unet_multiview_inputs(seed=gi.UNET_GPU_SEED + 50, channels=(320, 640), text_dim=64)
    it = iter(noises)
    np.random.seed(2024)
    out["mv_final"] = SO.sample_multiview(
        lambda x, t, **kw: m(x, t, encoder_hidden_states=kw.get("encoder_hidden_states"), F_mats=kw.get("F_mats"),
                             pose_embedding_features=kw.get("pose_embedding_features")).sample,
        lat4.clone(), text4, 3, 4.0, c2w, K, poses4, multistep=2, accumulate=1, noise_fn=lambda shape: next(it))
    save("unet_gpu", **out)


def train_step(Fm, gpu=False):
    """(``gpu``: the same at widths 320 / 640, the shapes the device kernels are instantiated for -> train_step_gpu.npz.)
    One optimisation step of the reference's training loop on the reference's own modules: the freeze of
    train_epi_control.py:245-259, PoseAdaptor(unet, pose_encoder) (pose_adaptor.py:81-104), and the statements of
    :548-630 from the latents on (mixed_precision_training=False branch; forward diffusion by the published add_noise
    formula -- diffusers' scheduler is not installable here).  The fixture holds the loss, the total gradient norm,
    per-parameter gradient norms and update norms, and three updated parameters."""
    import torch.nn.functional as F
    from animatediff.models.pose_adaptor import CameraPoseEncoder, PoseAdaptor
    from animatediff.models.unet import UNet3DConditionModelPoseCond
    ukw, pkw, useed = (gi.UNET_GPU_KWARGS, gi.UNET_GPU_PROCESSOR_KWARGS, gi.UNET_GPU_SEED) if gpu else \
        (gi.UNET_KWARGS, gi.UNET_PROCESSOR_KWARGS, gi.UNET_SEED)
    unet_ = UNet3DConditionModelPoseCond(**ukw)
    unet_.set_all_attn_processor(add_spatial_lora=False, add_motion_lora=False, lora_kwargs={"lora_rank": 2, "lora_scale": 1.0},
                                 motion_lora_kwargs={"lora_rank": -1, "lora_scale": 1.0},
                                 sync_lora_kwargs={"sync_lora_rank": 0, "sync_lora_scale": 0}, spatial_extended_attention=False,
                                 **pkw)
    skip = ("pos_encoder.pe", "pos_encoder.coords")
    shapes = {k: tuple(v.shape) for k, v in unet_.state_dict().items() if not k.endswith(skip)}
    unet_.load_state_dict(gi.fill_state_dict(shapes, useed), strict=False)
    enc = CameraPoseEncoder(**(gi.TRAIN_GPU_POSE_ENCODER_KWARGS if gpu else gi.TRAIN_POSE_ENCODER_KWARGS))
    eshapes = {k: tuple(v.shape) for k, v in enc.state_dict().items() if not k.endswith("pos_encoder.pe")}
    enc.load_state_dict(gi.fill_state_dict(eshapes, gi.TRAIN_SEED + 1), strict=False)
    # train_epi_control.py:245-259
    unet_.requires_grad_(False)
    enc.requires_grad_(False)
    for name, param in unet_.named_parameters():
        if "epi_modules" in name or "sync" in name or "auxiliary" in name:
            param.requires_grad = True
    pose_adaptor = PoseAdaptor(unet_, enc)
    names = [k for k, v in unet_.named_parameters() if v.requires_grad]
    params = [v for k, v in unet_.named_parameters() if v.requires_grad]
    optimizer = torch.optim.AdamW(params, **gi.TRAIN_OPT)
    pose_adaptor.train()
    lat, text, _poses = gi.unet_inputs(gi.UNET_GPU_SEED, channels=(320, 640), text_dim=64) if gpu else gi.unet_inputs()
    noise, timesteps, plucker = gi.train_inputs()
    F2 = Fm.reshape(2, 16, 3, 3)
    betas = torch.linspace(0.00085, 0.012, 1000, dtype=torch.float32)
    ac = torch.cumprod(1.0 - betas, dim=0)[timesteps].reshape(-1, 1, 1, 1, 1)
    noisy = ac ** 0.5 * lat + (1.0 - ac) ** 0.5 * noise
    before = [p.detach().clone() for p in params]
    with torch.enable_grad():
        model_pred, auxiliary = pose_adaptor(noisy, timesteps, encoder_hidden_states=text, pose_embedding=plucker,
                                             F_mats=F2, H_mats=None)
        assert auxiliary is None
        warped_masks = torch.ones_like(lat)
        loss = F.mse_loss(model_pred.float() * warped_masks, noise.float() * warped_masks, reduction="mean")
        loss.backward()
    grad_norms = torch.stack([p.grad.norm() for p in params])
    total = torch.nn.utils.clip_grad_norm_(filter(lambda p: p.requires_grad, pose_adaptor.parameters()), gi.TRAIN_MAX_GRAD_NORM)
    optimizer.step()
    deltas = torch.stack([(p.detach() - b).norm() for p, b in zip(params, before)])
    small = [i for i, p in enumerate(params) if p.numel() <= 4096]
    picks = [small[0], small[len(small) // 2], small[-1]]
    out = {"names": np.array(names), "loss": loss.detach(), "total_norm": total, "grad_norms": grad_norms, "delta_norms": deltas,
           "model_pred": model_pred.detach(), "picks": np.array(picks)}
    for j, i in enumerate(picks):
        out[f"after{j}"] = params[i].detach()
    print(f"train_step: {len(names)} trainable tensors, loss {float(loss):.5f}, grad norm {float(total):.4f}")
    save("train_step_gpu" if gpu else "train_step", **out)


def pipeline_mv():
    """The V-view generation path end to end on the reference's own code: camera rig, intrinsics and Pluecker rays by
    EXECUTING inference_epi_advanced.py:110-128,294-354 (rig 'interpolate', 4 views, 64x64 images; and a non-square
    64x32 rig for the per-axis intrinsics scaling), the reference's CameraPoseEncoder, and the oracle restatement of
    pipeline_animation_epi_advanced.py:596-705 (multistep 2, accumulate 2, guidance 4) driving the reference's UNet."""
    import math as _math
    import types
    from einops import rearrange
    from scipy.spatial.transform import Rotation as R, Slerp
    from animatediff.data.dataset_validation import ray_condition
    from animatediff.models.pose_adaptor import CameraPoseEncoder
    from animatediff.models.unet import UNet3DConditionModelPoseCond
    from oracle import sampler_oracle as SO
    cam_fn_src = _ref_lines("inference_epi_advanced.py", 110, 128)
    rig_src = _ref_lines("inference_epi_advanced.py", 294, 354)
    cfg = gi.PIPE_MV

    def rig(width, height):
        ns = {"np": np, "math": _math, "R": R, "Slerp": Slerp, "torch": torch, "ray_condition": ray_condition,
              "rearrange": rearrange, "device": "cpu",
              "pipeline": types.SimpleNamespace(unet=types.SimpleNamespace(device="cpu"))}
        exec(cam_fn_src, ns)
        ns["args"] = types.SimpleNamespace(view_num=cfg["view_num"], video_length=16, image_width=width, image_height=height,
                                           cam_pattern=cfg["cam_pattern"], camera_dist=cfg["camera_dist"], cam_perturb_traj=0.0)
        exec(rig_src, ns)
        return ns["c2ws"][0], ns["K_mats"][0], ns["plucker_embedding"]        # [Vf,4,4], [Vf,3,3], [V,6,f,H,W]
    out = {}
    c2w_ns, K_ns, pl_ns = rig(64, 32)
    out["ns_c2w"], out["ns_K"], out["ns_plucker_sub"] = c2w_ns, K_ns, pl_ns[:, :, ::5, ::7, ::9]
    c2w, K, pl = rig(cfg["width"], cfg["height"])
    out["c2w"], out["K"], out["plucker_sub"] = c2w, K, pl[:, :, ::5, ::9, ::9]
    m = UNet3DConditionModelPoseCond(**gi.UNET_KWARGS)
    m.set_all_attn_processor(add_spatial_lora=False, add_motion_lora=False, lora_kwargs={"lora_rank": 2, "lora_scale": 1.0},
                             motion_lora_kwargs={"lora_rank": -1, "lora_scale": 1.0},
                             sync_lora_kwargs={"sync_lora_rank": 0, "sync_lora_scale": 0}, spatial_extended_attention=False,
                             **gi.UNET_PROCESSOR_KWARGS)
    m.eval()
    skip = ("pos_encoder.pe", "pos_encoder.coords")
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if not k.endswith(skip)}
    m.load_state_dict(gi.fill_state_dict(shapes, gi.UNET_SEED), strict=False)
    enc = CameraPoseEncoder(**gi.PIPE_POSE_ENCODER_KWARGS).eval()
    eshapes = {k: tuple(v.shape) for k, v in enc.state_dict().items() if not k.endswith("pos_encoder.pe")}
    enc.load_state_dict(gi.fill_state_dict(eshapes, gi.PIPE_SEED), strict=False)
    V = cfg["view_num"]
    feats = [rearrange(x, "(b f) c h w -> b c f h w", b=V) for x in enc(pl)]          # pipeline :583-586
    out["feat0_sub"], out["feat1"] = feats[0][:, ::4, ::5], feats[1][:, ::8, ::5]
    lat4, text4, _poses4, noises = gi.unet_multiview_inputs(V, seed=gi.PIPE_SEED + 1)
    it = iter(noises)
    np.random.seed(77)

    def call(x, t, **kw):
        return m(x, t, encoder_hidden_states=kw.get("encoder_hidden_states"), F_mats=kw.get("F_mats"),
                 pose_embedding_features=kw.get("pose_embedding_features")).sample
    out["final"] = SO.sample_multiview(call, lat4.clone(), text4, cfg["num_inference_steps"], cfg["guidance_scale"],
                                       c2w.numpy(), K.numpy(), feats, multistep=cfg["multistep"], accumulate=cfg["accumulate_step"],
                                       noise_fn=lambda shape: next(it))
    save("pipeline_mv", **out)


def train_batch():
    """One batch of the reference's TRAINING dataset (dataset_train_realestate10k.py: RealEstate10KPoseFolded.get_batch
    executed verbatim; image decoding stubbed to a blank 360x640 frame, the annotation index replaced by one entry) on a
    synthetic 40-frame camera file: F matrices, folded cameras / intrinsics and Pluecker rays of the two videos."""
    import random
    import tempfile
    import types
    from animatediff.data import dataset_train_realestate10k as D
    cfg = gi.TRAIN_BATCH
    with tempfile.TemporaryDirectory() as tmp:
        pose = os.path.join(tmp, "clip.txt")
        gi.write_synthetic_pose_file(pose, cfg["n_cams"], cfg["seed"])
        ds = object.__new__(D.RealEstate10KPoseFolded)
        ds.root_path, ds.sample_stride, ds.sample_n_frames, ds.sample_size = tmp, cfg["sample_stride"], 16, (256, 256)
        ds.validation_prompts = ds.validation_negative_prompts = ds.validation_video_split = None
        ds.pose_file_0 = ds.pose_file_1 = None
        ds.dataset = [{"clip_name": "clip", "clip_path": tmp, "caption": "a room", "pose_file": "clip.txt"}]
        ds.pixel_transforms = lambda x: x
        H, W = cfg["image_hw"]
        D.imageio = types.SimpleNamespace(imread=lambda path: np.zeros((H, W, 3), dtype=np.uint8))
        random.seed(cfg["random_seed"])
        _pix, _cap, plucker_embedding, F_mats, _name, _vp, _vn, ret_c2w, ret_K = ds.get_batch(0, 0)
        # the same clip as 4 interpolated views (validation_video_split = 4, :466-478)
        ds.validation_video_split = 4
        random.seed(cfg["random_seed"])
        _pix, _cap, plucker4, _F4, _name, _vp, _vn, c2w4, K4 = ds.get_batch(0, 0)
    save("train_batch", F_mats=F_mats, c2w=ret_c2w[0], K=ret_K, plucker_sub=plucker_embedding[::3, :, ::17, ::19],
         plucker_shape=np.array(plucker_embedding.shape), mv_c2w=c2w4, mv_K=K4, mv_plucker_sub=plucker4[::5, :, ::17, ::19],
         mv_plucker_shape=np.array(plucker4.shape))


def unet_extended(Fm):
    """The reference UNet with ``spatial_extended_attention=True`` (AttnProcessor of attention_processor.py:67-79 on the SD
    spatial attention): one forward with the 7-token text of the UNet fixture (the quirk: any context that is not 77 tokens
    long is extended too), one with a 77-token text (cross-attention untouched), one on a 4-video batch (> 32 frames: the
    32-frame chunking)."""
    from animatediff.models.unet import UNet3DConditionModelPoseCond
    m = UNet3DConditionModelPoseCond(**gi.UNET_KWARGS)
    m.set_all_attn_processor(add_spatial_lora=False, add_motion_lora=False, lora_kwargs={"lora_rank": 2, "lora_scale": 1.0},
                             motion_lora_kwargs={"lora_rank": -1, "lora_scale": 1.0},
                             sync_lora_kwargs={"sync_lora_rank": 0, "sync_lora_scale": 0}, spatial_extended_attention=True,
                             **gi.UNET_PROCESSOR_KWARGS)
    m.eval()
    skip = ("pos_encoder.pe", "pos_encoder.coords")
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if not k.endswith(skip)}
    m.load_state_dict(gi.fill_state_dict(shapes, gi.UNET_SEED), strict=False)
    lat, text, poses = gi.unet_inputs()
    F2 = Fm.reshape(2, 16, 3, 3)
    out = {"t7": m(lat, 481, encoder_hidden_states=text, F_mats=F2, pose_embedding_features=poses).sample}
    text77 = gi.text77()
    out["t77"] = m(lat, 481, encoder_hidden_states=text77, F_mats=F2, pose_embedding_features=poses).sample
    lat4 = torch.cat([lat, lat.flip(0) * 0.5])
    out["b4"] = m(lat4, 481, encoder_hidden_states=torch.cat([text77, text77 * 0.5]), F_mats=torch.cat([F2, F2]),
                  pose_embedding_features=[torch.cat([p, p.flip(0)]) for p in poses]).sample
    save("unet_ext", **out)


def state_keys():
    """Full state_dict key tables (parameters AND persistent buffers, with shapes) of the reference's UNet and pose encoder:
    the strict-load contract (inference_epi.py:118-119 loads the pose encoder with strict=True)."""
    from animatediff.models.pose_adaptor import CameraPoseEncoder
    from animatediff.models.unet import UNet3DConditionModelPoseCond
    out = {}
    sd = UNet3DConditionModelPoseCond(**gi.UNET_KWARGS).state_dict()
    out["unet_keys"] = np.array(sorted(sd))
    out["unet_shapes"] = np.array([str(tuple(sd[k].shape)) for k in sorted(sd)])
    for case in gi.POSE_ENCODER_CASES:
        sd = CameraPoseEncoder(**case["kwargs"]).state_dict()
        out[case["name"] + "_keys"] = np.array(sorted(sd))
        out[case["name"] + "_shapes"] = np.array([str(tuple(sd[k].shape)) for k in sorted(sd)])
        pe = [k for k in sorted(sd) if k.endswith(".pe")]
        if pe:
            out[case["name"] + "_pe0"] = sd[pe[0]]
    save("state_keys", **out)


def plucker():
    """ray_condition of the reference (data/dataset_validation.py:102-144) on the shipped pose pair's cameras at the
    256 x 256 sample size and at a small non-square 5 x 7 grid, plus one horizontally flipped camera."""
    from animatediff.data.dataset_validation import ray_condition
    ds = ValRealEstate10KPoseFolded(sample_n_frames=16, relative_pose=True, sample_size=256, validation_prompts=["x"],
                                    pose_file_0=os.path.join(REF, gi.POSE_FILE_0),
                                    pose_file_1=os.path.join(REF, gi.POSE_FILE_1))
    c2w_list, _K, intr_list = ds.load_cameras_specific()
    intr = torch.as_tensor(np.array(intr_list)).float()[None]            # [1, 31, 4], as get_batch builds it (:253)
    c2w = torch.as_tensor(c2w_list)[None]                                # [1, 31, 4, 4] float32
    full = ray_condition(intr, c2w, 256, 256, device="cpu")[0].permute(0, 3, 1, 2)   # [31, 6, 256, 256]
    small_K = torch.tensor([[[6.1, 4.3, 3.4, 2.6]] * 4])
    small = ray_condition(small_K, c2w[:, :4], 5, 7, device="cpu")[0]                 # [4, 5, 7, 6]
    flip = torch.tensor([False, True, False, True])
    flipped = ray_condition(small_K, c2w[:, :4], 5, 7, device="cpu", flip_flag=flip)[0]
    save("plucker", intrinsics=intr[0], c2w=c2w[0], frame17=full[17], sub8=full[:, :, 3::8, 5::8],
         chan_sum=full.double().sum(dim=(2, 3)), abs_sum=full.double().abs().sum().item(),
         small_K=small_K[0], small=small, flipped=flipped, flip=flip)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] in ("plucker", "bias_pseudo", "pose_encoder", "unet", "train_step", "pipeline_mv", "train_batch", "state_keys", "unet_ext"):
        if sys.argv[1] == "unet_ext":
            unet_extended(torch.from_numpy(np.load(os.path.join(HERE, "geometry.npz"))["F_mats"]))
            sys.exit(0)
        if sys.argv[1] == "state_keys":
            state_keys()
            sys.exit(0)
        if sys.argv[1] == "train_batch":
            train_batch()
            sys.exit(0)
        if sys.argv[1] == "pipeline_mv":
            pipeline_mv()
            sys.exit(0)
        if sys.argv[1] == "train_step":
            train_step(torch.from_numpy(np.load(os.path.join(HERE, "geometry.npz"))["F_mats"]))
            train_step(torch.from_numpy(np.load(os.path.join(HERE, "geometry.npz"))["F_mats"]), gpu=True)
            sys.exit(0)
        if sys.argv[1] == "unet":
            unet(torch.from_numpy(np.load(os.path.join(HERE, "geometry.npz"))["F_mats"]))
            unet_gpu(torch.from_numpy(np.load(os.path.join(HERE, "geometry.npz"))["F_mats"]))
            sys.exit(0)
        {"plucker": plucker, "bias_pseudo": bias_pseudo, "pose_encoder": pose_encoder}[sys.argv[1]]()
        sys.exit(0)
    Fm = geometry()
    bias(Fm)
    attention(Fm)
    module(Fm)
    temporal()
    sampler()
    plucker()
    bias_pseudo()
    pose_encoder()
    unet(Fm)
    unet_gpu(Fm)
    train_step(Fm)
    train_step(Fm, gpu=True)
    pipeline_mv()
    train_batch()
    state_keys()
    unet_extended(Fm)

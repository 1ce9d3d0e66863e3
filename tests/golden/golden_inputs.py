"""Deterministic inputs / weights shared by ``make_golden.py`` (reference side, build
container) and the tests (oracle + CUDA side, anywhere).  ``numpy.random.RandomState`` is a
frozen legacy stream, so the same seed gives the same bits on every platform; nothing here
reads /root/reference.
"""
import numpy as np
import torch

POSE_FILE_0 = "assets/pose_files/2f25826f0d0ef09a.txt"
POSE_FILE_1 = "assets/pose_files/2c80f9eb0d3b2bb4.txt"
THETA_SEED = 1234  # torch.manual_seed before each reference call that draws the first-frame slope


def _randn(seed, *shape, scale=1.0):
    return torch.from_numpy((np.random.RandomState(seed).standard_normal(shape) * scale).astype(np.float32))


def hidden_tokens(R, N, C, seed):
    return _randn(seed, R, N, C)


def hidden_video(B, C, Fr, h, w, seed):
    return _randn(seed, B, C, Fr, h, w)


def _linear(seed, out_f, in_f, gain=1.0):
    return _randn(seed, out_f, in_f, scale=gain / np.sqrt(in_f))


def attn_state_dict(C, seed, prefix=""):
    """to_q/to_k/to_v (no bias), to_out.0 (+bias) -- names as in SURVEY.md section 4.3."""
    return {
        prefix + "to_q.weight": _linear(seed + 11, C, C, 1.5),
        prefix + "to_k.weight": _linear(seed + 12, C, C, 1.5),
        prefix + "to_v.weight": _linear(seed + 13, C, C, 1.0),
        prefix + "to_out.0.weight": _linear(seed + 14, C, C, 1.0),
        prefix + "to_out.0.bias": _randn(seed + 15, C, scale=0.1),
    }


def _block_state_dict(C, seed, prefix):
    sd = {}
    for i in range(2):
        sd.update(attn_state_dict(C, seed + 100 * (i + 1), f"{prefix}attention_blocks.{i}."))
        sd[f"{prefix}norms.{i}.weight"] = 1.0 + _randn(seed + 31 + i, C, scale=0.1)
        sd[f"{prefix}norms.{i}.bias"] = _randn(seed + 41 + i, C, scale=0.1)
    sd[prefix + "ff.net.0.proj.weight"] = _linear(seed + 51, 8 * C, C)
    sd[prefix + "ff.net.0.proj.bias"] = _randn(seed + 52, 8 * C, scale=0.1)
    sd[prefix + "ff.net.2.weight"] = _linear(seed + 53, C, 4 * C)
    sd[prefix + "ff.net.2.bias"] = _randn(seed + 54, C, scale=0.1)
    sd[prefix + "ff_norm.weight"] = 1.0 + _randn(seed + 55, C, scale=0.1)
    sd[prefix + "ff_norm.bias"] = _randn(seed + 56, C, scale=0.1)
    return sd


def _transformer_state_dict(C, seed, prefix):
    sd = {
        prefix + "norm.weight": 1.0 + _randn(seed + 1, C, scale=0.1),
        prefix + "norm.bias": _randn(seed + 2, C, scale=0.1),
        prefix + "proj_in.weight": _linear(seed + 3, C, C),
        prefix + "proj_in.bias": _randn(seed + 4, C, scale=0.1),
        # proj_out is zero-initialised by the reference (epi_module.py:81-82) which would make
        # the module an identity: the fixtures use a NON-zero proj_out (SURVEY.md hard part 5).
        prefix + "proj_out.weight": _linear(seed + 5, C, C, 0.5),
        prefix + "proj_out.bias": _randn(seed + 6, C, scale=0.1),
    }
    sd.update(_block_state_dict(C, seed + 1000, prefix + "transformer_blocks.0."))
    return sd


def module_state_dict(C, seed):
    return _transformer_state_dict(C, seed, "epi_transformer.")


def temporal_state_dict(C, seed):
    sd = _transformer_state_dict(C, seed, "temporal_transformer.")
    p = "temporal_transformer.transformer_blocks.0.attention_blocks.0.processor.qkv_merge."
    # zero-initialised in the reference (attention_processor.py:249-251); non-zero here
    sd[p + "weight"] = _linear(seed + 7, C, C, 0.5)
    sd[p + "bias"] = _randn(seed + 8, C, scale=0.1)
    return sd


def view_pair_index(V, n_frames=16):
    """kv_index pairing view v with V-1-v (a perfect matching like
    pipeline_animation_epi_advanced.py:625-633 draws), as numpy int64 like the pipeline's."""
    idx = np.concatenate([np.arange(n_frames) + (V - 1 - v) * n_frames for v in range(V)])
    return idx.astype(np.int64)


def tile_F(Fm, R):
    """[32,3,3] golden matrices -> [R,3,3]: alternate the matrices and their frame-flipped copy."""
    reps = [Fm if i % 2 == 0 else Fm.flip(0) for i in range((R + 31) // 32)]
    return torch.cat(reps, dim=0)[:R].contiguous()


ATTN_CASES = [
    dict(name="swap", C=320, feat=4, R=32, seed=100),
    dict(name="swap_theta", C=320, feat=4, R=32, seed=200, rand_slope=True, rows=2),
    dict(name="swap_f8", C=320, feat=8, R=32, seed=300, queries=4),
    dict(name="kvidx", C=320, feat=4, R=64, seed=400, kv="pair", rows=2),
    dict(name="multi", C=320, feat=4, R=32, seed=500, kv="multi", rows=2),
    dict(name="fixff", C=320, feat=4, R=64, seed=600, fix_firstframe=True, rows=2),
    dict(name="chunk", C=64, feat=4, R=208, seed=700),
    dict(name="c640", C=640, feat=4, R=32, seed=800, rows=2),
]


def stored(case, out):
    """The fixtures keep every ``rows``-th row and every ``queries``-th query of an output."""
    return out[::case.get("rows", 1), ::case.get("queries", 1)]


def attn_F_input(case, Fm):
    R = case["R"]
    if case.get("kv") == "pair":
        return [tile_F(Fm, R), view_pair_index(R // 16)]
    if case.get("kv") == "multi":
        half = np.concatenate([np.arange(R // 2, R), np.arange(0, R // 2)])
        return [tile_F(Fm, 2 * R), np.concatenate([half, np.arange(R)]).astype(np.int64)]
    return tile_F(Fm, R)


MODULE_CASES = [
    dict(name="mod", C=320, feat=4, B=2, seed=1100),
    dict(name="mod_theta", C=320, feat=4, B=2, seed=1200, rand_slope=True),
    dict(name="mod_kvidx", C=320, feat=4, B=4, seed=1300, kv="pair"),
]


def module_F_input(case, Fm):
    B = case["B"]
    F4 = tile_F(Fm, B * 16).reshape(B, 16, 3, 3)
    if case.get("kv") == "pair":
        return [F4, view_pair_index(B)]
    return F4


TEMPORAL_CASES = [
    dict(name="tmp_pose", C=320, feat=4, B=2, seed=2100, pose=True),
    dict(name="tmp_nopose", C=320, feat=4, B=2, seed=2200, pose=False),
]


# ---- CameraCtrl pose encoder (pose_adaptor.py:248-329): two small configurations that take every structural branch -----
POSE_ENCODER_CASES = [
    # the shipped configuration's structure (configs/inference_config.yaml:34-46) at reduced width
    dict(name="pe_cfg", kwargs=dict(downscale_factor=8, channels=[64, 128, 128], nums_rb=2, cin=384, ksize=1, sk=True,
                                    use_conv=False, compression_factor=1, temporal_attention_nhead=8,
                                    attention_block_types=["Temporal_Self"], temporal_position_encoding=True,
                                    temporal_position_encoding_max_len=16), B=2, H=32, W=32, seed=7100),
    # 3x3 convolutions, strided-convolution downsampling, three ResNet blocks per level, no positional encoding
    dict(name="pe_conv", kwargs=dict(downscale_factor=4, channels=[32, 64], nums_rb=3, cin=96, ksize=3, sk=True,
                                     use_conv=True, compression_factor=1, temporal_attention_nhead=4,
                                     attention_block_types=["Temporal_Self"], temporal_position_encoding=False,
                                     temporal_position_encoding_max_len=16), B=1, H=16, W=24, seed=7200),
    # the shipped configuration at the two narrowest SD-1.5 widths: the shapes the DEVICE kernels are instantiated for
    # (LayerNorm widths 320 / 640, temporal attention head sizes 40 / 80) -- the case the GPU parity test runs
    dict(name="pe_dev", kwargs=dict(downscale_factor=8, channels=[320, 640], nums_rb=2, cin=384, ksize=1, sk=True,
                                    use_conv=False, compression_factor=1, temporal_attention_nhead=8,
                                    attention_block_types=["Temporal_Self"], temporal_position_encoding=True,
                                    temporal_position_encoding_max_len=16), B=2, H=32, W=32, seed=7300, device=True),
]


def fill_state_dict(shapes, seed):
    """Deterministic weights for a {name: shape} table (sorted by name, one numpy stream per tensor): matrices and
    convolutions ~ N(0, 0.5^2 / fan_in), 1-D ``weight`` (norm scales) ~ 1 + 0.1 N(0,1), biases ~ 0.1 N(0,1)."""
    sd = {}
    for i, name in enumerate(sorted(shapes)):
        shape = tuple(shapes[name])
        rs = np.random.RandomState(seed + i)
        if len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            w = rs.standard_normal(shape) * (0.5 / np.sqrt(fan_in))
        elif name.endswith("weight"):
            w = 1.0 + 0.1 * rs.standard_normal(shape)
        else:
            w = 0.1 * rs.standard_normal(shape)
        sd[name] = torch.from_numpy(w.astype(np.float32))
    return sd


def plucker_like(B, H, W, seed, frames=16):
    rs = np.random.RandomState(seed)
    return torch.from_numpy(rs.standard_normal((B, 6, frames, H, W)).astype(np.float32))


# ---- full UNet wiring (unet.py / unet_blocks.py) at a reduced width: every block kind, two resolutions --------------------
UNET_MOTION_KWARGS = dict(num_attention_heads=8, num_transformer_block=1, attention_block_types=["Temporal_Self", "Temporal_Self"],
                          temporal_position_encoding=True, temporal_position_encoding_max_len=32, temporal_attention_dim_div=1,
                          zero_initialize=False)
UNET_EPI_KWARGS = dict(num_attention_heads=8, num_transformer_block=1, attention_block_types=["Epi_Self", "Epi_Self"],
                       epi_position_encoding=True, epi_position_encoding_feat_max_size=64, epi_position_encoding_F_mat_size=256,
                       epi_rand_slope_ff=False, zero_initialize=True)
UNET_KWARGS = dict(sample_size=8, in_channels=4, out_channels=4, block_out_channels=(32, 64),
                   down_block_types=("CrossAttnDownBlock3D", "DownBlock3D"), up_block_types=("UpBlock3D", "CrossAttnUpBlock3D"),
                   layers_per_block=1, cross_attention_dim=48, attention_head_dim=8, norm_num_groups=8,
                   use_motion_module=True, motion_module_resolutions=(1, 2, 4, 8), motion_module_mid_block=False,
                   motion_module_type="Vanilla", motion_module_kwargs=UNET_MOTION_KWARGS,
                   use_epi_module=True, epi_module_resolutions=(1, 2, 4, 8), epi_module_mid_block=False,
                   epi_module_kwargs=UNET_EPI_KWARGS)
UNET_PROCESSOR_KWARGS = dict(add_spatial=False, spatial_attn_names="attn1", add_temporal=True, temporal_attn_names="0",
                             pose_feature_dimensions=[32, 64], query_condition=True, key_value_condition=True, scale=1.0)
UNET_SEED = 8800


# The same layout at the two narrowest widths of SD-1.5 (320, 640 -> head sizes 40 / 80 at 8x8 / 4x4 latents): the smallest
# UNet whose sync / temporal modules have shapes the DEVICE kernels are instantiated for (the GPU parity tests use it; the
# reduced-width configuration above only runs against the plain-torch test double).
UNET_GPU_KWARGS = dict(UNET_KWARGS, block_out_channels=(320, 640), norm_num_groups=32, cross_attention_dim=64)
UNET_GPU_PROCESSOR_KWARGS = dict(UNET_PROCESSOR_KWARGS, pose_feature_dimensions=[320, 640])
UNET_GPU_SEED = 9900


def unet_inputs(seed=UNET_SEED, feat=8, channels=(32, 64), text_dim=48):
    rs = np.random.RandomState(seed + 1)
    lat = torch.from_numpy(rs.standard_normal((2, 4, 16, feat, feat)).astype(np.float32))
    text = torch.from_numpy((0.5 * rs.standard_normal((2, 7, text_dim))).astype(np.float32))
    poses = [torch.from_numpy((0.5 * rs.standard_normal((2, c, 16, feat >> i, feat >> i))).astype(np.float32))
             for i, c in enumerate(channels)]
    return lat, text, poses


def unet_multiview_inputs(V=4, seed=UNET_SEED + 50, feat=8, channels=(32, 64), text_dim=48):
    """V-view inputs of the advanced pipeline at the UNet fixture's width: latents, CFG-interleaved text (U, C per view),
    pose features of all views, and the re-noise tensors of the multistep repeats (drawn here, injected on both sides)."""
    rs = np.random.RandomState(seed)
    lat = torch.from_numpy(rs.standard_normal((V, 4, 16, feat, feat)).astype(np.float32))
    text = torch.from_numpy((0.5 * rs.standard_normal((2, 7, text_dim))).astype(np.float32)).repeat(V, 1, 1)
    poses = [torch.from_numpy((0.5 * rs.standard_normal((V, c, 16, feat >> i, feat >> i))).astype(np.float32))
             for i, c in enumerate(channels)]
    noises = [torch.from_numpy(rs.standard_normal((V, 4, 16, feat, feat)).astype(np.float32)) for _ in range(8)]
    return lat, text, poses, noises


# ---- one optimisation step of the sync modules (train_epi_control.py:245-259, :548-630) at the reduced UNet width -----------
TRAIN_POSE_ENCODER_KWARGS = dict(downscale_factor=2, channels=[32, 64], nums_rb=2, cin=24, ksize=1, sk=True, use_conv=False,
                                 compression_factor=1, temporal_attention_nhead=4, attention_block_types=["Temporal_Self"],
                                 temporal_position_encoding=True, temporal_position_encoding_max_len=16)
TRAIN_GPU_POSE_ENCODER_KWARGS = dict(TRAIN_POSE_ENCODER_KWARGS, channels=[320, 640], temporal_attention_nhead=8)      # device widths
TRAIN_SEED = 8900
TRAIN_OPT = dict(lr=1e-3, betas=(0.9, 0.999), weight_decay=1e-2, eps=1e-8)      # the script's AdamW (:275-281), larger lr
TRAIN_MAX_GRAD_NORM = 0.1
TRAIN_TIMESTEPS = (481, 37)


def train_inputs(seed=TRAIN_SEED, feat=8):
    """noise [2,4,16,f,f], one timestep per video, Pluecker-like embedding [2,6,16,2f,2f] (pose encoder downscale 2)."""
    rs = np.random.RandomState(seed)
    noise = torch.from_numpy(rs.standard_normal((2, 4, 16, feat, feat)).astype(np.float32))
    plucker = torch.from_numpy(rs.standard_normal((2, 6, 16, 2 * feat, 2 * feat)).astype(np.float32))
    return noise, torch.tensor(TRAIN_TIMESTEPS, dtype=torch.long), plucker


# ---- the V-view generation path from the camera rig to the final latents (inference_epi_advanced.py:293-354 + pipeline) ------
PIPE_POSE_ENCODER_KWARGS = dict(downscale_factor=8, channels=[32, 64], nums_rb=2, cin=384, ksize=1, sk=True, use_conv=False,
                                compression_factor=1, temporal_attention_nhead=8, attention_block_types=["Temporal_Self"],
                                temporal_position_encoding=True, temporal_position_encoding_max_len=16)
PIPE_MV = dict(view_num=4, cam_pattern="interpolate", camera_dist=0.5, height=64, width=64, num_inference_steps=3,
               guidance_scale=4.0, multistep=2, accumulate_step=2)
PIPE_SEED = 9100


# ---- one batch of the TRAINING dataset (dataset_train_realestate10k.py: get_batch) on a synthetic 40-frame camera file -------
TRAIN_BATCH = dict(n_cams=40, image_hw=(360, 640), sample_stride=8, random_seed=5, seed=9300)


def write_synthetic_pose_file(path, n=40, seed=9300):
    """A RealEstate10K-format camera file (header line, then ``id fx fy cx cy 0 0 w2c[3x4]`` per frame) of a smooth
    trajectory: the shipped pose files hold 16 frames, the training dataset needs at least 31."""
    rs = np.random.RandomState(seed)
    axis = rs.standard_normal(3)
    axis /= np.linalg.norm(axis)
    drift = 0.03 * rs.standard_normal(3)
    with open(path, "w") as fh:
        fh.write("https://example.invalid/synthetic\n")
        for i in range(n):
            ang = 0.012 * i + 0.002 * np.sin(0.7 * i)
            Kx = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
            Rm = np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * (Kx @ Kx)
            tvec = drift * i + 0.01 * np.array([np.sin(0.3 * i), np.cos(0.2 * i), 0.5 * np.sin(0.1 * i)])
            w2c = np.concatenate([Rm, tvec[:, None]], axis=1).reshape(-1)
            vals = [1000 + 33 * i, 0.49 + 0.001 * i, 0.87, 0.5, 0.5, 0.0, 0.0] + [float(v) for v in w2c]
            fh.write(" ".join(repr(float(v)) if k else str(int(v)) for k, v in enumerate(vals)) + "\n")


def text77(seed=UNET_SEED + 7, text_dim=48):
    """A CLIP-length (77-token) text context for the UNet fixture (the reference tells text from image tokens by that length)."""
    rs = np.random.RandomState(seed)
    return torch.from_numpy((0.5 * rs.standard_normal((2, 77, text_dim))).astype(np.float32))

"""Sampler loops that drive the cross-video synchronisation path.

Mirrors the two denoising loops of the reference, restricted to what they do *around* the UNet
call (the UNet itself is any callable with the reference's signature):

* :func:`sample_pair`       -- ``pipelines/pipeline_animation_epi.py:563-656`` (two-video inference,
  CFG batch laid out ``[U-src, C-src, U-tgt, C-tgt]``, fixed pose pair, swap-halves partner);
* :func:`sample_multiview`  -- ``pipelines/pipeline_animation_epi_advanced.py:596-705`` (V views:
  a fresh random view pairing per UNet call drawn from the *numpy global* generator, fundamental
  matrices recomputed per call, CFG batch interleaved ``[v0-U, v0-C, v1-U, ...]``, ``multistep``
  re-noising between repeats of one timestep, ``accumulate_step`` averaging).

B200-first differences (results identical, RNG consumption identical):
  * the per-call pose -> F recompute is one kernel launch (``cvd_pairwise_fundamental``) instead of
    a Python loop of V*16 numpy inversions + an H2D copy (``:636-647``);
  * the partner index stays an int32 device tensor; nothing on the path synchronises the host;
  * ``shard=ViewShard``: every rank owns V/world views (all frames, both CFG branches) and runs the
    same loop on them; the pairing is drawn identically on every rank (same numpy seed), re-noise
    is drawn full-size from the identically seeded torch generator and sliced, and the only
    exchange is the partner-row pull inside the epi attention (``cvd_b200/sharding.py``).

``DDIMSchedule`` restates ``diffusers==0.24.0`` ``DDIMScheduler`` for the configuration of
``configs/inference_config.yaml:58-64`` (linear betas, ``steps_offset=1``, ``clip_sample=False``,
"leading" spacing, ``set_alpha_to_one=True``, ``eta=0``).
"""
import numpy as np
import torch

FRAMES = 16   # hard-coded in the reference (epi_module.py:316, attention_processor.py:631)


class DDIMSchedule:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="linear",
                 steps_offset=1, clip_sample=False, set_alpha_to_one=True, device="cpu"):
        if beta_schedule != "linear":
            raise NotImplementedError("only the reference configuration (beta_schedule='linear') is restated")
        if clip_sample:
            raise NotImplementedError("clip_sample=True is not the reference configuration")
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset
        betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0).to(device)
        self.final_alpha_cumprod = torch.tensor(1.0, device=device) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = None

    def set_timesteps(self, num_inference_steps):
        self.num_inference_steps = num_inference_steps
        ratio = self.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64) + self.steps_offset
        self.timesteps = [int(t) for t in ts]
        return self.timesteps

    def scale_model_input(self, sample, t):
        return sample

    def step(self, model_output, t, sample):
        """x_t -> x_{t-1}, eta = 0, epsilon prediction."""
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        pred_x0 = (sample - (1.0 - a_t) ** 0.5 * model_output) / a_t ** 0.5
        return a_prev ** 0.5 * pred_x0 + (1.0 - a_prev) ** 0.5 * model_output

    def add_noise(self, original_samples, noise, timesteps):
        """The forward diffusion the training step applies (``train_epi_control.py:555`` -> diffusers ``add_noise``):
        ``sqrt(a_t) x + sqrt(1 - a_t) noise`` with one timestep per sample (``timesteps`` int [B], broadcast over the
        remaining dimensions)."""
        a = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)[timesteps.to(original_samples.device)]
        a = a.reshape(-1, *([1] * (original_samples.ndim - 1)))
        return a ** 0.5 * original_samples + (1.0 - a) ** 0.5 * noise

    def renoise(self, latents, t, noise):
        """pipeline_animation_epi_advanced.py:700-705: back from t_prev to t between multistep repeats.
        (``alphas_cumprod[prev_t]`` with Python indexing, negative ``prev_t`` included, like the reference.)"""
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        r = self.alphas_cumprod[t] / self.alphas_cumprod[prev_t]
        return latents * r ** 0.5 + (1.0 - r) ** 0.5 * noise


def _sample_of(out):
    return out.sample if hasattr(out, "sample") else (out[0] if isinstance(out, (tuple, list)) else out)


def sample_pair(unet, latents, F_mats, text_embeddings, schedule, num_inference_steps=25, guidance_scale=8.5,
                pose_embedding_features=None, callback=None):
    """pipeline_animation_epi.py:563-656.  latents [2b,c,f,h,w] (first half source videos, second half
    target videos), F_mats [2b,f,3,3], text_embeddings [2b',l,c] = (uncond, cond).  Returns the final
    latents."""
    cfg = guidance_scale > 1.0
    timesteps = schedule.set_timesteps(num_inference_steps)
    if F_mats is not None:
        assert F_mats.shape[0] == latents.shape[0]
        F_mats = F_mats.type_as(latents)
        if cfg:                       # order: Uncond-src, Cond-src, Uncond-tgt, Cond-tgt (:567-571)
            a, b = F_mats.chunk(2)
            F_mats = torch.cat([a, a, b, b], dim=0)
    text_embeddings = text_embeddings.repeat(2, 1, 1)
    if pose_embedding_features is not None and cfg:
        bs = pose_embedding_features[0].shape[0]
        pose_embedding_features = [torch.cat([x[:bs // 2], x[:bs // 2], x[bs // 2:], x[bs // 2:]], dim=0)
                                   for x in pose_embedding_features]
    for i, t in enumerate(timesteps):
        if cfg:
            a, b = latents.chunk(2)
            x_in = torch.cat([a, a, b, b], dim=0)
        else:
            x_in = latents
        x_in = schedule.scale_model_input(x_in, t)
        noise_pred = _sample_of(unet(x_in, t, F_mats=F_mats, encoder_hidden_states=text_embeddings,
                                     pose_embedding_features=pose_embedding_features)).to(latents.dtype)
        if cfg:
            p = noise_pred.chunk(4)
            uncond, text = torch.cat([p[0], p[2]], dim=0), torch.cat([p[1], p[3]], dim=0)
            noise_pred = uncond + guidance_scale * (text - uncond)
        latents = schedule.step(noise_pred, t, latents)
        if callback is not None:
            callback(i, t, latents)
    return latents


def draw_view_pairing(V, n_frames=FRAMES, rng=np.random):
    """pipeline_animation_epi_advanced.py:621-633: one ``rng.permutation(V)`` split into V/2 disjoint
    pairs.  Returns (video_id [V*f] partner row per (view, frame) row, offset [V*f] in rows)."""
    bf = V * n_frames
    perm = rng.permutation(V).reshape(2, V // 2)
    off = np.zeros(V, dtype=np.int64)
    off[perm[0, :]] = perm[1, :] - perm[0, :]
    off[perm[1, :]] = perm[0, :] - perm[1, :]
    off = (off * n_frames).repeat(n_frames)
    return (np.arange(bf) + off + bf) % bf, off


def cfg_pair_index(offset_rows, n_frames=FRAMES):
    """:653-656: the pairing in the CFG-interleaved batch (rows [v0-U, v0-C, v1-U, ...] x f)."""
    bf2 = 2 * offset_rows.shape[0]
    return (np.arange(bf2) + offset_rows.repeat(2, axis=0) * 2) % bf2


def sample_multiview(unet, latents, text_embeddings, schedule, num_inference_steps=25, guidance_scale=8.5,
                     aux_c2w=None, aux_K_mats=None, F_mats=None, pose_embedding_features=None, multistep=1,
                     accumulate_step=1, shard=None, rng=np.random, noise_fn=None, callback=None, kv_on_device=False):
    """pipeline_animation_epi_advanced.py:596-705.

    latents [V_local,c,f,h,w]; text_embeddings [2*V_local,l,c] already CFG-interleaved per view
    (``:556-560``) or [V_local,l,c] without guidance; aux_c2w [V*f,4,4] and aux_K_mats [V*f,3,3] of ALL views
    (device tensors or arrays); ``F_mats`` [2,f,3,3] only for V == 2 (then no kv_index is passed and
    the layer swaps halves, exactly as ``:637,657``).  ``pose_embedding_features``: per-level
    [V_local,C,f,h,w] (CameraCtrl features), required like in the reference for the sync path to
    be active (``:619``).  ``shard``: :class:`cvd_b200.sharding.ViewShard` or None.  ``noise_fn(shape)``: source of the
    re-noise tensor (default: ``torch.randn`` on the global generator of the latents' device, like ``:701``).
    ``kv_on_device``: hand kv_index to the UNet as a device int32 tensor instead of the reference's numpy array, so that
    a CUDA-graphed UNet (:class:`cvd_b200.GraphedModule`) replays ONE graph for every pairing (a numpy kv_index is part
    of the captured control flow: one graph per distinct pairing).
    """
    from . import ops
    cfg = guidance_scale > 1.0
    world = 1 if shard is None else shard.world
    rank = 0 if shard is None else shard.rank
    V_loc = latents.shape[0]
    V = V_loc * world
    f = latents.shape[2]
    assert f == FRAMES, "the sync path hard-codes 16 frames per video (epi_module.py:316)"
    assert V % 2 == 0, "views are paired: V must be even (pipeline_animation_epi_advanced.py:625)"
    dev = latents.device
    timesteps = schedule.set_timesteps(num_inference_steps)
    if V != 2:
        assert aux_c2w is not None and aux_K_mats is not None
        c2w = torch.as_tensor(np.asarray(aux_c2w.cpu() if torch.is_tensor(aux_c2w) else aux_c2w), dtype=torch.float64).reshape(-1, 4, 4).to(dev)
        Kd = torch.as_tensor(np.asarray(aux_K_mats.cpu() if torch.is_tensor(aux_K_mats) else aux_K_mats), dtype=torch.float64).reshape(-1, 3, 3).to(dev)
        assert c2w.shape[0] == V * f
    if pose_embedding_features is not None and cfg:
        pose_embedding_features = [x.repeat_interleave(2, dim=0) for x in pose_embedding_features]
    loc = slice(rank * V_loc, (rank + 1) * V_loc)

    for i, t in enumerate(timesteps):
        n_rep = multistep if i != len(timesteps) - 1 else 1
        for mt in range(n_rep):
            noise_pred_full = torch.zeros_like(latents)
            for _acc in range(accumulate_step):
                F_in = None
                if pose_embedding_features is not None:
                    video_id, off = draw_view_pairing(V, f, rng)          # same draw on every rank
                    if V != 2:
                        partner = torch.from_numpy(video_id.astype(np.int32)).to(dev)
                        Fm = ops.pairwise_fundamental(c2w, Kd, partner).reshape(V, f, 3, 3)[loc]
                    else:
                        Fm = F_mats[loc]
                    Fm = Fm.type_as(latents)
                    if cfg:
                        Fm = Fm.repeat_interleave(2, dim=0)
                        kv = cfg_pair_index(off, f)
                    else:
                        kv = video_id
                    if shard is not None:
                        kv = shard.local_kv_index(kv, kv.shape[0])
                    if kv_on_device:
                        kv = torch.from_numpy(np.ascontiguousarray(kv).astype(np.int32)).to(dev, non_blocking=True)
                    F_in = [Fm, kv] if (V != 2 or shard is not None) else Fm
                x_in = latents.repeat_interleave(2, dim=0) if cfg else latents
                x_in = schedule.scale_model_input(x_in, t)
                noise_pred = _sample_of(unet(x_in, t, F_mats=F_in, encoder_hidden_states=text_embeddings,
                                             pose_embedding_features=pose_embedding_features)).to(latents.dtype)
                if cfg:
                    uncond, text = noise_pred[0::2], noise_pred[1::2]
                    noise_pred = uncond + guidance_scale * (text - uncond)
                noise_pred_full += noise_pred
            latents = schedule.step(noise_pred_full / accumulate_step, t, latents)
            if mt != n_rep - 1:
                # torch.randn_like on the GLOBAL generator (:701); sharded ranks draw the full tensor and slice
                full = (V,) + tuple(latents.shape[1:])
                noise = (torch.randn(full, device=dev, dtype=latents.dtype) if noise_fn is None else noise_fn(full))[loc]
                latents = schedule.renoise(latents, t, noise)
        if callback is not None:
            callback(i, t, latents)
    return latents


def unit_table(V, cfg, accumulate_step):
    """The independent UNet evaluations of ONE denoising update of the V-view loop
    (pipeline_animation_epi_advanced.py:603-699): unit = (accumulate step a, view v, CFG branch c) -- every accumulate
    step runs the UNet on the same latents under its own view pairing and the predictions are summed (:605,686-699).
    Unit u = (a * V + v) * n_cfg + c.  Returns int arrays (a, v, c), each [A * V * n_cfg]."""
    n_cfg = 2 if cfg else 1
    u = np.arange(accumulate_step * V * n_cfg)
    return u // (V * n_cfg), (u // n_cfg) % V, u % n_cfg


def sample_multiview_units(unet, latents, text_embeddings, schedule, num_inference_steps=25, guidance_scale=8.5,
                           aux_c2w=None, aux_K_mats=None, pose_embedding_features=None, multistep=1, accumulate_step=1,
                           shard=None, rng=np.random, noise_fn=None, callback=None, kv_on_device=False, fundamental=None):
    """The V-view loop of pipeline_animation_epi_advanced.py:596-705 partitioned by (view, CFG branch, accumulate step)
    UNITS instead of whole views (BASELINE config 4: 6 views x 2 CFG x 2 accumulate steps = 24
    units -> 3 per GPU on 8 GPUs; 6 views do not split over 8 ranks).

    Every rank holds the latents of ALL V views (a few hundred KB) and evaluates the UNet on its own
    ``U / world`` units in one batch: a unit is one video (16 rows) with one text embedding, the pose features of its
    view and the F matrices of its accumulate step's pairing; its epi attention reads the rows of the unit
    (a, partner_a(v), c), wherever that unit lives (global ``kv_index`` rows, exchanged by the layers exactly as for
    view sharding).  The guidance combination and the accumulate sum are linear in the unit predictions, so every rank
    adds ``w_c / A`` times its predictions into a [V, ...] buffer (w = 1 - g for the unconditional branch, g for the
    conditional one) and ONE small all-reduce per update gives every rank the full noise prediction; DDIM update and
    re-noise (drawn full-size from the identically seeded generator) then run replicated.

    latents [V,c,f,h,w] (all views, identical on every rank); text_embeddings [2V,l,c] CFG-interleaved per view
    (:556-560) or [V,l,c]; aux_c2w [V*f,4,4], aux_K_mats [V*f,3,3]; pose_embedding_features per level [V,C,f,h,w] (all
    views).  ``fundamental(c2w, K, partner_rows) -> [V*f,3,3]``: defaults to the device kernel
    (``ops.pairwise_fundamental``); the CPU tests pass the host fp64 path of ``cvd_b200.geometry``.
    Returns the final latents of all V views."""
    import torch.distributed as dist
    cfg = guidance_scale > 1.0
    world = 1 if shard is None else shard.world
    rank = 0 if shard is None else shard.rank
    V, f = latents.shape[0], latents.shape[2]
    assert f == FRAMES, "the sync path hard-codes 16 frames per video (epi_module.py:316)"
    assert V % 2 == 0, "views are paired: V must be even (pipeline_animation_epi_advanced.py:625)"
    assert pose_embedding_features is not None, "the sync path is only active with pose features (:619)"
    dev = latents.device
    A = accumulate_step
    ua, uv, uc = unit_table(V, cfg, A)
    U = ua.shape[0]
    if U % world != 0:
        raise ValueError(f"{U} (accumulate, view, cfg) units do not split over {world} ranks")
    U_loc = U // world
    mine = slice(rank * U_loc, (rank + 1) * U_loc)
    la, lv, lc = ua[mine], uv[mine], uc[mine]
    n_cfg = 2 if cfg else 1
    timesteps = schedule.set_timesteps(num_inference_steps)
    c2w = torch.as_tensor(np.asarray(aux_c2w.cpu() if torch.is_tensor(aux_c2w) else aux_c2w), dtype=torch.float64).reshape(-1, 4, 4)
    Kd = torch.as_tensor(np.asarray(aux_K_mats.cpu() if torch.is_tensor(aux_K_mats) else aux_K_mats), dtype=torch.float64).reshape(-1, 3, 3)
    assert c2w.shape[0] == V * f
    if fundamental is None:
        from . import ops
        c2w, Kd = c2w.to(dev), Kd.to(dev)

        def fundamental(c2w_, K_, partner_rows):
            return ops.pairwise_fundamental(c2w_, K_, torch.from_numpy(partner_rows.astype(np.int32)).to(dev))
    lv_t = torch.as_tensor(lv, device=dev, dtype=torch.long)
    text_loc = text_embeddings[torch.as_tensor(lv * n_cfg + lc, device=text_embeddings.device, dtype=torch.long)]
    pose_loc = [x[lv_t.to(x.device)] for x in pose_embedding_features]
    weight = torch.as_tensor(np.where(lc == 1, guidance_scale, 1.0 - guidance_scale) if cfg else np.ones(U_loc),
                             device=dev, dtype=latents.dtype).view(U_loc, 1, 1, 1, 1)
    frames = np.arange(f)

    for i, t in enumerate(timesteps):
        n_rep = multistep if i != len(timesteps) - 1 else 1
        for mt in range(n_rep):
            # one pairing per accumulate step, drawn in the reference's order (:621-633 inside the accumulate loop)
            F_rows, kv_rows = [], []
            pair = [draw_view_pairing(V, f, rng)[0] for _ in range(A)]               # video_id [V*f] per accumulate step
            F_acc = [fundamental(c2w, Kd, pair[a_]).reshape(V, f, 3, 3) for a_ in sorted(set(la.tolist()))]
            F_by_a = dict(zip(sorted(set(la.tolist())), F_acc))
            for a_, v_, c_ in zip(la, lv, lc):
                F_rows.append(F_by_a[int(a_)][int(v_)])
                pv = pair[int(a_)][int(v_) * f] // f                                   # partner view under pairing a
                pu = (int(a_) * V + int(pv)) * n_cfg + int(c_)                         # partner unit: same a, same c
                kv_rows.append(pu * f + frames)
            Fm = torch.stack(F_rows).to(device=dev).type_as(latents)
            kv = np.concatenate(kv_rows)                                             # GLOBAL rows, this rank's units
            if kv_on_device:
                kv = torch.from_numpy(np.ascontiguousarray(kv).astype(np.int32)).to(dev, non_blocking=True)
            x_in = schedule.scale_model_input(latents[lv_t], t)
            pred = _sample_of(unet(x_in, t, F_mats=[Fm, kv], encoder_hidden_states=text_loc,
                                   pose_embedding_features=pose_loc)).to(latents.dtype)
            noise_pred = torch.zeros_like(latents)
            noise_pred.index_add_(0, lv_t, pred * weight)
            if world > 1:
                dist.all_reduce(noise_pred, group=shard.group)
            latents = schedule.step(noise_pred / A, t, latents)
            if mt != n_rep - 1:
                shape = tuple(latents.shape)
                noise = torch.randn(shape, device=dev, dtype=latents.dtype) if noise_fn is None else noise_fn(shape)
                latents = schedule.renoise(latents, t, noise)
        if callback is not None:
            callback(i, t, latents)
    return latents


def install_sharding(module, shard, peer=None):
    """Point every EpiAttnProcessor under ``module`` at the view shard (and the NVLink peer exchange)."""
    from .epi_module import EpiAttnProcessor
    n = 0
    for m in module.modules():
        if isinstance(m, EpiAttnProcessor):
            m.shard, m.peer = shard, peer
            n += 1
    return n


def multiview_cameras(view_num, video_length=FRAMES, cam_pattern="circle", camera_dist=0.5, image_size=256, image_height=None):
    """Synthetic V-view camera rigs of ``inference_epi_advanced.py:293-343`` (patterns ``circle``, ``upper_hemi`` and
    ``interpolate``; every trajectory runs from the identity pose to the view's target pose, :func:`interpolate_pose`
    without perturbation).  ``image_size`` is the image width (and height unless ``image_height`` is given): the 256-pixel
    intrinsics are scaled per axis (``:299-300``).  Returns (c2w [V*f,4,4] float32, K [V*f,3,3] float32)."""
    import math
    K = np.array([[223.578, 0, 128], [0, 223.578, 128], [0, 0, 1]], dtype=np.float64)
    K = np.repeat(K[None], view_num * video_length, axis=0)
    K[:, 0] *= image_size / 256
    K[:, 1] *= (image_size if image_height is None else image_height) / 256
    out = []
    for i in range(view_num):
        if cam_pattern == "upper_hemi":
            angle = math.pi / (view_num - 1) * i + math.pi
        elif cam_pattern == "circle":
            angle = 2 * math.pi / view_num * i
        elif cam_pattern == "interpolate":
            angle = math.pi / (view_num - 1) * i
        else:
            raise ValueError(cam_pattern)
        if cam_pattern == "interpolate":                  # :302-309: a shallow arc in front of the scene
            cam_at = np.array([math.cos(angle), math.cos(angle + 0.5) * 0.3, -math.sin(angle) * 0.2]) * camera_dist
        else:
            cam_at = np.array([math.cos(angle), math.sin(angle), 0]) * camera_dist
        cam_z = np.array([0, 0, 1]) - cam_at
        cam_x = np.array([1, 0, 0])
        cam_y = np.cross(cam_z, cam_x)
        cam_y = cam_y / (np.linalg.norm(cam_y) + 1e-6)
        cam_x = np.cross(cam_y, cam_z)
        cam_x = cam_x / (np.linalg.norm(cam_x) + 1e-6)
        tgt = np.eye(4)
        tgt[:3, :3] = np.stack([cam_x, cam_y, cam_z], axis=1)
        tgt[:3, 3] = cam_at
        out.append(interpolate_pose(np.eye(4), tgt, video_length))
    return np.concatenate(out, axis=0).astype(np.float32), K.astype(np.float32)


def interpolate_pose(src_pose, tgt_pose, n):
    """``inference_epi_advanced.py:110-128`` with ``cam_perturb_traj = 0``: translation blended linearly,
    rotation by spherical interpolation between the two orientations, over the n frames."""
    from scipy.spatial.transform import Rotation, Slerp
    w = np.linspace(0.0, 1.0, n)
    poses = np.repeat(np.asarray(src_pose, dtype=np.float64)[None], n, axis=0)
    poses[:, :3, 3] = (1.0 - w)[:, None] * src_pose[:3, 3][None] + w[:, None] * tgt_pose[:3, 3][None]
    ends = Rotation.from_matrix(np.stack([src_pose[:3, :3], tgt_pose[:3, :3]]))
    poses[:, :3, :3] = Slerp([0.0, 1.0], ends)(w).as_matrix()
    return poses

"""The two-video generation path of ``inference_epi.py`` / ``AnimationEpiPipeline.__call__`` from pose files to final
latents, on this package's modules -- the part of the reference's pipeline between
the text encoder and the VAE decoder, which stay the reference's own (out of scope):

    pose files --(geometry.folded_pose_conditioning: F matrices, folded cameras, intrinsics)-->
    Pluecker rays on the device (ops.plucker_embedding) --> CameraPoseEncoder --> per-level pose features
    --> sampler.sample_pair (CFG batch [U-src, C-src, U-tgt, C-tgt], DDIM) around UNet3DConditionModelPoseCond.

Reference: ``inference_epi.py:236-262`` (batch preparation), ``pipelines/pipeline_animation_epi.py:520-660``
(latent preparation, pose-feature batching, denoising loop).

``MultiviewPipeline`` is the same for the V-view path of ``inference_epi_advanced.py`` /
``AnimationEpiAdvancedPipeline.__call__``: synthetic camera rig -> intrinsics -> Pluecker rays -> pose encoder ->
sampler.sample_multiview (random pairing, per-call F from the poses, multistep / accumulate).  Reference:
``inference_epi_advanced.py:293-354`` (rig, intrinsics, ``ray_condition``), ``pipelines/pipeline_animation_epi_advanced.py:520-705``."""
import torch

from . import geometry, ops, sampler


class PairPipeline:
    def __init__(self, unet, pose_encoder, schedule=None):
        self.unet, self.pose_encoder = unet, pose_encoder
        self.schedule = schedule

    @torch.no_grad()
    def pose_features(self, cond, height, width, device):
        """dataset_validation.py:256-268 + inference_epi.py:252-254 + pipeline_animation_epi.py:592-595: folded Pluecker
        embedding [2, 6, n, H, W] -> pose-encoder features per UNet level, each [2, C, n, h, w]."""
        n = cond["c2w"].shape[0] // 2
        pl = ops.plucker_embedding(cond["intrinsics"].to(device), cond["c2w"].to(device).contiguous(), height, width)   # [2n,6,H,W]
        pl = pl.view(2, n, 6, height, width).permute(0, 2, 1, 3, 4).contiguous()
        feats = self.pose_encoder(pl)                                                        # list of [(2 n), C, h, w]
        return [f.view(2, n, f.shape[1], f.shape[2], f.shape[3]).permute(0, 2, 1, 3, 4).contiguous() for f in feats]

    @torch.no_grad()
    def __call__(self, text_embeddings, pose_file_0, pose_file_1, height=256, width=256, video_length=16,
                 num_inference_steps=25, guidance_scale=8.5, generator=None, latents=None, sample_size=256):
        """``text_embeddings`` [2, L, C] = (negative / unconditional, prompt) of the ONE prompt both videos share
        (pipeline_animation_epi.py:556-560).  Returns the final latents [2, 4, f, height / 8, width / 8]."""
        dev = next(self.unet.parameters()).device
        cond = geometry.folded_pose_conditioning(pose_file_0, pose_file_1, video_length, sample_size)
        F_mats = cond["F_mats"].to(dev).view(2, video_length, 3, 3)          # torch.cat(F_mats.chunk(2, dim=1), dim=0)
        poses = self.pose_features(cond, height, width, dev)
        if latents is None:
            shape = (2, self.unet.config["in_channels"], video_length, height // 8, width // 8)
            latents = torch.randn(shape, generator=generator, device=dev)
        sch = self.schedule if self.schedule is not None else sampler.DDIMSchedule(device=dev)
        text = text_embeddings.to(dev)
        if guidance_scale <= 1.0:            # no guidance: the prompt embedding alone (_encode_prompt without the negative half)
            text = text[1:]
        return sampler.sample_pair(self.unet, latents * sch.init_noise_sigma, F_mats, text, sch,
                                   num_inference_steps, guidance_scale, pose_embedding_features=poses)


class MultiviewPipeline:
    def __init__(self, unet, pose_encoder, schedule=None):
        self.unet, self.pose_encoder = unet, pose_encoder
        self.schedule = schedule

    @torch.no_grad()
    def pose_features(self, c2w, K_mats, height, width, video_length, device):
        """inference_epi_advanced.py:346-354 + pipeline_animation_epi_advanced.py:583-588: intrinsics (fx, fy, cx, cy) of
        every frame, Pluecker rays [V, 6, f, H, W] on the device, pose-encoder features per UNet level [V, C, f, h, w]."""
        K = torch.as_tensor(K_mats, dtype=torch.float32)
        intr = torch.stack([K[:, 0, 0], K[:, 1, 1], K[:, 0, 2], K[:, 1, 2]], dim=1).to(device).contiguous()
        c2w = torch.as_tensor(c2w, dtype=torch.float32).to(device).contiguous()
        V = c2w.shape[0] // video_length
        pl = ops.plucker_embedding(intr, c2w, height, width)                                 # [V f, 6, H, W]
        pl = pl.view(V, video_length, 6, height, width).permute(0, 2, 1, 3, 4).contiguous()
        feats = self.pose_encoder(pl)                                                        # list of [(V f), C, h, w]
        return [f.view(V, video_length, f.shape[1], f.shape[2], f.shape[3]).permute(0, 2, 1, 3, 4).contiguous() for f in feats]

    @torch.no_grad()
    def __call__(self, text_embeddings, view_num, cam_pattern="circle", camera_dist=0.5, height=256, width=256, video_length=16,
                 num_inference_steps=25, guidance_scale=8.5, multistep=1, accumulate_step=1, generator=None, latents=None,
                 aux_c2w=None, aux_K_mats=None, shard=None, rng=None, noise_fn=None):
        """``text_embeddings`` [2, L, C] = (negative / unconditional, prompt), shared by all views
        (pipeline_animation_epi_advanced.py:527-529,566); ``aux_c2w`` [V f, 4, 4] / ``aux_K_mats`` [V f, 3, 3]: cameras of
        all views (default: the script's rig for ``cam_pattern``).  With ``shard`` (view sharding) this rank generates
        views [rank V / world, (rank + 1) V / world).  Returns the final latents of this rank's views [V_local, 4, f,
        height / 8, width / 8]."""
        import numpy as np
        dev = next(self.unet.parameters()).device
        if aux_c2w is None:
            aux_c2w, aux_K_mats = sampler.multiview_cameras(view_num, video_length, cam_pattern, camera_dist, width, height)
        poses = self.pose_features(aux_c2w, aux_K_mats, height, width, video_length, dev)
        world, rank = (1, 0) if shard is None else (shard.world, shard.rank)
        if view_num % world != 0:
            raise ValueError(f"{view_num} views do not split over {world} ranks (use sampler.sample_multiview_units)")
        loc = slice(rank * (view_num // world), (rank + 1) * (view_num // world))
        if latents is None:
            shape = (view_num, self.unet.config["in_channels"], video_length, height // 8, width // 8)
            latents = torch.randn(shape, generator=generator, device=dev)[loc]               # same draw on every rank
        sch = self.schedule if self.schedule is not None else sampler.DDIMSchedule(device=dev)
        text = text_embeddings.to(dev)
        text = text.repeat(latents.shape[0], 1, 1) if guidance_scale > 1.0 else text[1:].repeat(latents.shape[0], 1, 1)
        return sampler.sample_multiview(self.unet, latents * sch.init_noise_sigma, text, sch, num_inference_steps, guidance_scale,
                                        aux_c2w=aux_c2w, aux_K_mats=aux_K_mats, pose_embedding_features=[p[loc] for p in poses],
                                        multistep=multistep, accumulate_step=accumulate_step, shard=shard,
                                        rng=np.random if rng is None else rng, noise_fn=noise_fn)

"""The two-video generation path of ``inference_epi.py`` / ``AnimationEpiPipeline.__call__`` from pose files to final
latents, on this package's modules -- the part of the reference's pipeline between
the text encoder and the VAE decoder, which stay the reference's own (out of scope):

    pose files --(geometry.folded_pose_conditioning: F matrices, folded cameras, intrinsics)-->
    Pluecker rays on the device (ops.plucker_embedding) --> CameraPoseEncoder --> per-level pose features
    --> sampler.sample_pair (CFG batch [U-src, C-src, U-tgt, C-tgt], DDIM) around UNet3DConditionModelPoseCond.

Reference: ``inference_epi.py:236-262`` (batch preparation), ``pipelines/pipeline_animation_epi.py:520-660``
(latent preparation, pose-feature batching, denoising loop)."""
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
        return sampler.sample_pair(self.unet, latents * sch.init_noise_sigma, F_mats, text_embeddings.to(dev), sch,
                                   num_inference_steps, guidance_scale, pose_embedding_features=poses)

"""The reference's video UNet with the sync path on this package's kernels.

Mirrors ``UNet3DConditionModelPoseCond`` (``animatediff/models/unet.py:56-330,940-1452``) and its blocks
(``animatediff/models/unet_blocks.py``: ``CrossAttnDownBlock3D`` :308, ``DownBlock3D`` :493, ``UNetMidBlock3DCrossAttn``
:164, ``CrossAttnUpBlock3D`` :623, ``UpBlock3D`` :824): same constructor arguments for the configuration the shipped
configs use, same module tree and ``state_dict`` keys (an SD-1.5 + AnimateDiff + CameraCtrl + CVD checkpoint loads
unchanged), same forward signature and ``.sample`` output.  Per layer the order is the reference's: per-frame ResNet block
-> spatial transformer (per frame) -> motion module -> epi module (``unet_blocks.py:463-477``).

Execution: the SD-1.5 spatial blocks (``cvd_b200/sd_blocks.py``) are library convolutions / GEMMs under bf16 autocast;
the motion and epi modules are this package's drop-ins (``VanillaTemporalModule``, ``EpiModule``: hand-written sm_100a
kernels behind the C ABI).  ``set_all_attn_processor`` installs the CameraCtrl pose adaptor on the first temporal attention
block of every motion module like the reference (``unet.py:985-1143``, inference configuration).  ``motion_factory`` /
``epi_factory`` default to this package's modules (``get_motion_module`` / ``get_epi_module``, the reference's own factory
names, ``unet_blocks.py:9-10``); the tests and the reference-GPU leg of ``bench.py`` pass factories of oracle-backed twins
to obtain the reference's eager fp32 path on the same parameter tree.  Everything the shipped
inference / training configs leave off (first-frame fusers, image-layer LoRA, ControlNet residuals, class embeddings,
dual cross-attention, linear projections) is not built and raises.
"""
from typing import Any, Dict, List, Optional, Sequence

import torch
from torch import nn

from . import sd_blocks as sd
from .epi_module import get_epi_module
from .motion_module import AttnProcessor, PoseAdaptorAttnProcessor, get_motion_module


class UNet3DConditionOutput(dict):
    """``unet.py:44-47``: ``.sample`` (and ``.auxiliary``, None unless ``additional_channel`` > 0 -- not built)."""

    def __init__(self, sample, auxiliary=None):
        super().__init__(sample=sample, auxiliary=auxiliary)
        self.sample, self.auxiliary = sample, auxiliary


class InflatedConv3d(nn.Conv2d):
    """``resnet.py:19-27``: a 2-D convolution applied to every frame of ``[b, c, f, h, w]``."""

    def forward(self, x):
        b, c, f, h, w = x.shape
        y = super().forward(x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w))
        return y.view(b, f, y.shape[1], y.shape[2], y.shape[3]).permute(0, 2, 1, 3, 4)


def _frames(x):
    b, c, f, h, w = x.shape
    return x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w), (b, f)


def _video(y, bf):
    b, f = bf
    return y.view(b, f, y.shape[1], y.shape[2], y.shape[3]).permute(0, 2, 1, 3, 4)


class _Block3D(nn.Module):
    """One resolution level: ``layers`` x [ResNet (+ spatial transformer) + motion module + epi module] and an optional
    resampler.  ``kind``: "down" | "up" | "mid" (the mid block is resnet, [transformer, motion, epi, resnet] per layer)."""

    def __init__(self, kind, in_channels, out_channels, temb_channels, num_layers, cross_attention, heads,
                 cross_attention_dim, eps, groups, prev_output_channel=None, add_resample=False, downsample_padding=1,
                 use_motion_module=False, motion_module_type=None, motion_module_kwargs=None, use_epi_module=False,
                 epi_module_kwargs=None, motion_factory=get_motion_module, epi_factory=get_epi_module):
        super().__init__()
        self.kind = kind
        self.has_cross_attention = cross_attention
        resnets, attentions, motion, epi = [], [], [], []
        if kind == "mid":
            resnets.append(sd.ResnetBlock2D(in_channels, in_channels, temb_channels, eps, groups))
        for i in range(num_layers):
            if kind == "down":
                r_in = in_channels if i == 0 else out_channels
            elif kind == "up":      # unet_blocks.py:664-669: skip connection of the mirrored down layer is concatenated
                skip = in_channels if i == num_layers - 1 else out_channels
                r_in = (prev_output_channel if i == 0 else out_channels) + skip
            else:
                r_in = out_channels = in_channels
            if kind != "mid":
                resnets.append(sd.ResnetBlock2D(r_in, out_channels, temb_channels, eps, groups))
            if cross_attention:
                attentions.append(sd.Transformer2DModel(heads, out_channels // heads, out_channels, cross_attention_dim, groups))
            motion.append(motion_factory(out_channels, motion_module_type, motion_module_kwargs) if use_motion_module else None)
            epi.append(epi_factory(out_channels, epi_module_kwargs) if use_epi_module else None)
            if kind == "mid":
                resnets.append(sd.ResnetBlock2D(in_channels, in_channels, temb_channels, eps, groups))
        self.resnets = nn.ModuleList(resnets)
        if cross_attention:
            self.attentions = nn.ModuleList(attentions)
        # the reference keeps plain Python lists of None when a module kind is switched off (unet_blocks.py:396-397)
        self.motion_modules = nn.ModuleList(motion) if use_motion_module else motion
        self.epi_modules = nn.ModuleList(epi) if use_epi_module else epi
        self.downsamplers = self.upsamplers = None
        if kind == "down" and add_resample:
            self.downsamplers = nn.ModuleList([sd.Downsample2D(out_channels, downsample_padding)])
        if kind == "up" and add_resample:
            self.upsamplers = nn.ModuleList([sd.Upsample2D(out_channels)])

    def _layer(self, x, i, amp, temb_rep, text_rep, F_mats, H_mats, temb, encoder_hidden_states, mkw, aux, resnet_first=True):
        if resnet_first:
            fr, bf = _frames(x)
            with amp():
                fr = self.resnets[i](fr, temb_rep)
        else:
            fr, bf = _frames(x)
        if self.has_cross_attention:
            with amp():
                fr = self.attentions[i](fr, text_rep)
        x = _video(fr.float(), bf)
        if self.motion_modules[i] is not None:
            x = self.motion_modules[i](x, temb=temb, encoder_hidden_states=encoder_hidden_states, cross_attention_kwargs=mkw)
        if self.epi_modules[i] is not None:
            x, a = self.epi_modules[i](x, F_mats, H_mats, temb=temb, encoder_hidden_states=encoder_hidden_states,
                                       cross_attention_kwargs=mkw)
            aux += a
        return x

    def forward(self, x, amp, temb, text_rep, F_mats, H_mats, encoder_hidden_states, pose_feature, res_samples=None,
                upsample_size=None):
        frames = x.shape[2]
        temb_rep = temb.repeat_interleave(frames, dim=0)
        mkw = {"pose_feature": pose_feature}
        aux: List[Any] = []
        outputs = ()
        if self.kind == "mid":
            fr, bf = _frames(x)
            with amp():
                fr = self.resnets[0](fr, temb_rep)
            x = _video(fr.float(), bf)
            for i in range(len(self.resnets) - 1):
                x = self._layer(x, i, amp, temb_rep, text_rep, F_mats, H_mats, temb, encoder_hidden_states, mkw, aux,
                                resnet_first=False)
                fr, bf = _frames(x)
                with amp():
                    fr = self.resnets[i + 1](fr, temb_rep)
                x = _video(fr.float(), bf)
            return x, aux
        for i in range(len(self.resnets)):
            if self.kind == "up":
                x = torch.cat([x, res_samples[-1]], dim=1)
                res_samples = res_samples[:-1]
            x = self._layer(x, i, amp, temb_rep, text_rep, F_mats, H_mats, temb, encoder_hidden_states, mkw, aux)
            outputs += (x,)
        if self.downsamplers is not None:
            fr, bf = _frames(x)
            with amp():
                fr = self.downsamplers[0](fr)
            x = _video(fr.float(), bf)
            outputs += (x,)
        if self.upsamplers is not None:
            fr, bf = _frames(x)
            with amp():
                fr = self.upsamplers[0](fr, upsample_size)
            x = _video(fr.float(), bf)
        return (x, outputs, aux) if self.kind == "down" else (x, aux)


class _Config(dict):
    """Constructor arguments with attribute access, like diffusers' FrozenDict (``unet.config.sample_size``)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name) from None


class UNet3DConditionModelPoseCond(nn.Module):
    """``unet.py:56-330`` (constructor), ``:1145-1452`` (forward), for the block layout every shipped config uses."""

    def __init__(self, sample_size=None, in_channels=4, out_channels=4, flip_sin_to_cos=True, freq_shift=0,
                 down_block_types: Sequence[str] = ("CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "DownBlock3D"),
                 mid_block_type="UNetMidBlock3DCrossAttn",
                 up_block_types: Sequence[str] = ("UpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D"),
                 block_out_channels: Sequence[int] = (320, 640, 1280, 1280), layers_per_block=2, downsample_padding=1,
                 mid_block_scale_factor=1, act_fn="silu", norm_num_groups=32, norm_eps=1e-5, cross_attention_dim=1280,
                 attention_head_dim=8, use_motion_module=False, motion_module_resolutions=(1, 2, 4, 8),
                 motion_module_mid_block=False, motion_module_type=None, motion_module_kwargs: Optional[Dict] = None,
                 use_epi_module=False, epi_module_resolutions=(1, 2, 4, 8), epi_module_mid_block=False,
                 epi_module_kwargs: Optional[Dict] = None, decoder_add_posecond=True, amp=True,
                 motion_factory=get_motion_module, epi_factory=get_epi_module, **unsupported):
        super().__init__()
        off = {k: v for k, v in unsupported.items() if v not in (False, None, 0, "default", "silu", 1, 1.0)
               and k not in ("unet_use_cross_frame_attention", "unet_use_temporal_attention")}
        if off:
            raise NotImplementedError(f"UNet options outside the shipped CVD configuration are not built: {sorted(off)}")
        if act_fn != "silu" or mid_block_type != "UNetMidBlock3DCrossAttn" or mid_block_scale_factor != 1:
            raise NotImplementedError("only the SD-1.5 block configuration is built")
        motion_module_kwargs = dict(motion_module_kwargs or {})
        epi_module_kwargs = dict(epi_module_kwargs or {})
        # what the reference's scripts and pipelines read off the UNet (diffusers ModelMixin / ConfigMixin surface):
        # unet.config.sample_size / ["in_channels"] / dict(unet.config), unet.in_channels, unet.device, unet.dtype
        self.config = _Config(in_channels=in_channels, out_channels=out_channels, block_out_channels=tuple(block_out_channels),
                              cross_attention_dim=cross_attention_dim, sample_size=sample_size, layers_per_block=layers_per_block,
                              attention_head_dim=attention_head_dim, norm_num_groups=norm_num_groups)
        self.in_channels = in_channels
        self.decoder_add_posecond = decoder_add_posecond
        self.amp = amp                     # bf16 autocast for the SD blocks (False: fp32, parity tests against the reference)
        self.flip_sin_to_cos, self.freq_shift = flip_sin_to_cos, freq_shift
        ch = list(block_out_channels)
        n = len(ch)
        temb_ch = ch[0] * 4
        heads = [attention_head_dim] * n if isinstance(attention_head_dim, int) else list(attention_head_dim)
        self.conv_in = InflatedConv3d(in_channels, ch[0], kernel_size=3, padding=(1, 1))
        self.time_embedding = sd.TimestepEmbedding(ch[0], temb_ch)
        common = dict(temb_channels=temb_ch, cross_attention_dim=cross_attention_dim, eps=norm_eps, groups=norm_num_groups,
                      motion_module_type=motion_module_type, motion_module_kwargs=motion_module_kwargs,
                      epi_module_kwargs=epi_module_kwargs, motion_factory=motion_factory, epi_factory=epi_factory)
        self.down_blocks = nn.ModuleList()
        out_c = ch[0]
        for i, kind in enumerate(down_block_types):
            in_c, out_c = out_c, ch[i]
            res = 2 ** i
            self.down_blocks.append(_Block3D(
                "down", in_c, out_c, num_layers=layers_per_block, cross_attention=(kind == "CrossAttnDownBlock3D"),
                heads=heads[i], add_resample=(i != n - 1), downsample_padding=downsample_padding,
                use_motion_module=use_motion_module and res in motion_module_resolutions,
                use_epi_module=use_epi_module and res in epi_module_resolutions, **common))
        self.mid_block = _Block3D("mid", ch[-1], ch[-1], num_layers=1, cross_attention=True, heads=heads[-1],
                                  use_motion_module=use_motion_module and motion_module_mid_block,
                                  use_epi_module=use_epi_module and epi_module_mid_block, **common)
        self.up_blocks = nn.ModuleList()
        rev, rev_heads = ch[::-1], heads[::-1]
        out_c = rev[0]
        self.num_upsamplers = 0
        for i, kind in enumerate(up_block_types):
            res = 2 ** (3 - i)                         # unet.py:262 (hard-wired to four levels, kept as is)
            prev, out_c = out_c, rev[i]
            in_c = rev[min(i + 1, n - 1)]
            last = i == n - 1
            self.num_upsamplers += 0 if last else 1
            self.up_blocks.append(_Block3D(
                "up", in_c, out_c, num_layers=layers_per_block + 1, cross_attention=(kind == "CrossAttnUpBlock3D"),
                heads=rev_heads[i], prev_output_channel=prev, add_resample=not last,
                use_motion_module=use_motion_module and res in motion_module_resolutions,
                use_epi_module=use_epi_module and res in epi_module_resolutions, **common))
        self.conv_norm_out = nn.GroupNorm(num_channels=ch[0], num_groups=norm_num_groups, eps=norm_eps)
        self.conv_act = nn.SiLU()
        self.conv_out = InflatedConv3d(ch[0], out_channels, kernel_size=3, padding=1)

    @property
    def device(self) -> torch.device:
        return next(self.parameters()).device

    @property
    def dtype(self) -> torch.dtype:
        return next(self.parameters()).dtype

    # -- checkpoints (unet.py:862-912) ------------------------------------------------------------------------------------
    @classmethod
    def from_pretrained_2d(cls, pretrained_model_path, subfolder=None, unet_additional_kwargs=None, logger=None, **overrides):
        """Build the 3-D UNet from a Stable-Diffusion 2-D UNet folder, like the reference: ``config.json`` (block types
        replaced by their 3-D names, private ``_*`` keys dropped) + ``unet_additional_kwargs`` -> constructor; then
        ``diffusion_pytorch_model.safetensors`` loaded non-strictly (the motion / epi modules and pose-adaptor processors
        stay at their initial values until their own checkpoints are loaded, ``inference_epi.py:96-131``); unexpected keys
        are an error.  ``overrides``: constructor arguments of this package only (e.g. ``amp=False``)."""
        import json
        import os
        if subfolder is not None:
            pretrained_model_path = os.path.join(pretrained_model_path, subfolder)
        config_file = os.path.join(pretrained_model_path, "config.json")
        if not os.path.isfile(config_file):
            raise RuntimeError(f"{config_file} does not exist")
        with open(config_file, "r") as f:
            config = {k: v for k, v in json.load(f).items() if not k.startswith("_")}
        config["down_block_types"] = ["CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "CrossAttnDownBlock3D", "DownBlock3D"]
        config["up_block_types"] = ["UpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D", "CrossAttnUpBlock3D"]
        config.update(unet_additional_kwargs or {})
        config.update(overrides)
        model = cls(**config)
        model_file = os.path.join(pretrained_model_path, "diffusion_pytorch_model.safetensors")
        if not os.path.isfile(model_file):
            raise RuntimeError(f"{model_file} does not exist")
        import safetensors.torch
        state_dict = safetensors.torch.load_file(model_file, device="cpu")
        missing, unexpected = model.load_state_dict(state_dict, strict=False)
        if logger is not None:
            logger.info(f"Missing keys: {len(missing)}; Unexpected keys: {len(unexpected)};")
        assert len(unexpected) == 0, unexpected[:5]
        return model

    def set_image_layer_lora_scale(self, lora_scale: float = 1.0):
        """unet.py:338-341.  Recorded on the blocks like the reference does; it scales the image-layer LoRA processors, which
        this package does not build (``set_all_attn_processor(add_spatial_lora=True)`` raises), so it changes nothing here
        -- the training script calls it unconditionally (``train_epi_control.py:208,582,642``)."""
        for block in list(self.down_blocks) + list(self.up_blocks) + [self.mid_block]:
            block.lora_scale = lora_scale

    def set_motion_module_lora_scale(self, lora_scale: float = 1.0):
        """unet.py:343-346 (same remark: motion LoRA processors are not built)."""
        for block in list(self.down_blocks) + list(self.up_blocks) + [self.mid_block]:
            block.motion_lora_scale = lora_scale

    # -- processors (unet.py:985-1143, the branch the inference / training scripts take) ---------------------------------
    def set_all_attn_processor(self, add_spatial=False, spatial_attn_names="attn1", add_temporal=False, add_spatial_lora=False,
                               add_motion_lora=False, temporal_attn_names="0", pose_feature_dimensions=(320, 640, 1280, 1280),
                               lora_kwargs=None, motion_lora_kwargs=None, sync_lora_kwargs=None,
                               spatial_extended_attention=False, **attention_processor_kwargs):
        if add_spatial or add_spatial_lora or add_motion_lora:
            raise NotImplementedError("spatial pose adaptors and LoRA processors are not built")
        for module in self.modules():                       # unet.py:1062: AttnProcessor(attn_type="spatial", ...)
            if isinstance(module, sd.SpatialAttention):
                module.spatial_extended_attention = bool(spatial_extended_attention)
        if sync_lora_kwargs and (sync_lora_kwargs.get("sync_lora_rank", 0) or sync_lora_kwargs.get("sync_lora_scale", 0)):
            raise NotImplementedError("sync LoRA is unused at inference (inference_epi.py:84) and not built")
        names = set(temporal_attn_names.split(","))
        for name, module in self.named_modules():
            if not hasattr(module, "set_processor") or "motion_modules." not in name:
                continue                                    # epi processors are installed by EpiSelfAttention itself
            if add_temporal and name.split(".")[-1] in names:
                C = module.to_q.in_features
                assert C in tuple(pose_feature_dimensions)
                module.set_processor(PoseAdaptorAttnProcessor(hidden_size=C, pose_feature_dim=C, **attention_processor_kwargs))
            else:
                module.set_processor(AttnProcessor(attn_type="temporal"))

    # -- forward (unet.py:1145-1452) ----------------------------------------------------------------------------------
    def forward(self, sample, timestep, encoder_hidden_states=None, F_mats=None, H_mats=None, class_labels=None,
                attention_mask=None, cross_attention_kwargs=None, pose_embedding_features=None, return_dict=True,
                down_block_additional_residuals=None, mid_block_additional_residual=None, motion_module_alphas=1.0):
        if class_labels is not None or attention_mask is not None or down_block_additional_residuals is not None \
                or mid_block_additional_residual is not None:
            raise NotImplementedError("class labels, attention masks and ControlNet residuals are not on the CVD path")
        up_factor = 2 ** self.num_upsamplers
        forward_upsample_size = any(s % up_factor != 0 for s in sample.shape[-2:])
        if not torch.is_tensor(timestep):
            timestep = torch.tensor([timestep], dtype=torch.float32, device=sample.device)
        elif timestep.ndim == 0:
            timestep = timestep[None].to(sample.device)
        timestep = timestep.expand(sample.shape[0])

        def amp():
            return torch.autocast(sample.device.type, dtype=torch.bfloat16, enabled=self.amp)

        emb = self.time_embedding(sd.timestep_embedding(timestep, self.conv_in.out_channels, self.flip_sin_to_cos, self.freq_shift))
        frames = sample.shape[2]
        text_rep = encoder_hidden_states.repeat_interleave(frames, dim=0)
        poses = [None] * len(self.down_blocks) if pose_embedding_features is None else list(pose_embedding_features)
        with amp():
            x = self.conv_in(sample)
        x = x.float()
        res_samples = (x,)
        aux_all: List[Any] = []
        for block, pose in zip(self.down_blocks, poses):
            x, res, aux = block(x, amp, emb, text_rep, F_mats, H_mats, encoder_hidden_states, pose)
            res_samples += res
            aux_all += aux
        x, aux = self.mid_block(x, amp, emb, text_rep, F_mats, H_mats, encoder_hidden_states, poses[-1])
        aux_all += aux
        for i, block in enumerate(self.up_blocks):
            pose = poses[-(i + 1)] if self.decoder_add_posecond else None
            n_res = len(block.resnets)
            res, res_samples = res_samples[-n_res:], res_samples[:-n_res]
            size = res_samples[-1].shape[-2:] if (i != len(self.up_blocks) - 1 and forward_upsample_size) else None
            x, aux = block(x, amp, emb, text_rep, F_mats, H_mats, encoder_hidden_states, pose, res_samples=res,
                           upsample_size=size)
            aux_all += aux
        fr, bf = _frames(x)
        with amp():
            fr = self.conv_act(self.conv_norm_out(fr))
            out = self.conv_out(_video(fr, bf))
        out = out.to(sample.dtype)
        if not return_dict:
            return (out, None)
        return UNet3DConditionOutput(sample=out, auxiliary=None)

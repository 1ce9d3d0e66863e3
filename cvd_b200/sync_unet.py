"""Hot-path skeleton of the reference UNet: every (motion module -> epi module) site of
``animatediff/models/unet_blocks.py`` (``:286-299,434-476,584-610,760-801,900-932``) at the
reference's four resolutions, with the SD-1.5 spatial blocks (ResnetBlock2D, Transformer2DModel,
Down/Upsample2D -- out of scope) replaced by parameter-light stand-ins.

It exists so that the sampler loops (``cvd_b200/sampler.py``) and the "UNet denoise steps/s"
half of the BASELINE metric can be exercised end to end on the path this repository owns:
20 epi modules (2 per down level, 3 per up level = 5 per resolution, i.e. the 40 epi-attention
calls per forward that ``bench.py`` times) interleaved with 20 temporal modules whose first
attention block takes the CameraCtrl pose feature.  It is NOT the SD UNet and its samples are not
images: steps/s measured on it excludes the backbone, and says so wherever it is reported.

Stand-ins (plain PyTorch, fp32): ``conv_in``/``conv_out`` 1x1x1 convolutions, per-site additive
conditioning ``Linear(sin-embedding(t)) + Linear(mean-token text embedding)``, 2x2 average pooling
+ 1x1x1 convolution between down levels, nearest 2x upsampling + 1x1x1 convolution between up
levels, additive skip connections.

``forward(sample[B,4,F,h,w], timestep, F_mats=None, encoder_hidden_states=None,
pose_embedding_features=None)`` keeps the argument names of ``UNet3DConditionModel.forward``
(``unet.py:735-748``) that the pipelines use.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

CHANNELS = (320, 640, 1280, 1280)     # block_out_channels of SD-1.5
DOWN_SITES, UP_SITES = 2, 3           # layers_per_block, layers_per_block + 1

EPI_KWARGS = dict(num_attention_heads=8, num_transformer_block=1, attention_block_types=["Epi_Self", "Epi_Self"],
                  epi_position_encoding=True, epi_position_encoding_feat_max_size=64,
                  epi_position_encoding_F_mat_size=256, epi_rand_slope_ff=True, zero_initialize=True)
MOTION_KWARGS = dict(num_attention_heads=8, num_transformer_block=1,
                     attention_block_types=["Temporal_Self", "Temporal_Self"], temporal_position_encoding=True,
                     temporal_position_encoding_max_len=32, temporal_attention_dim_div=1, zero_initialize=False)


def timestep_embedding(t, dim, device):
    half = dim // 2
    freq = torch.exp(-math.log(10000.0) * torch.arange(half, device=device, dtype=torch.float32) / half)
    arg = (t.to(device=device, dtype=torch.float32) if torch.is_tensor(t) else float(t)) * freq   # tensor: graph-capturable
    return torch.cat([torch.sin(arg), torch.cos(arg)])


def default_epi_factory(channels, epi_kwargs):
    from .epi_module import get_epi_module
    return get_epi_module(channels, epi_kwargs)


def default_motion_factory(channels, motion_kwargs):
    """Temporal module with the CameraCtrl processors of configs/inference_config.yaml:48-56: the pose adaptor on
    attention block 0, the plain temporal processor on block 1."""
    from .motion_module import AttnProcessor, PoseAdaptorAttnProcessor, get_motion_module
    m = get_motion_module(channels, "Vanilla", motion_kwargs)
    blk = m.temporal_transformer.transformer_blocks[0]
    blk.attention_blocks[0].set_processor(PoseAdaptorAttnProcessor(
        hidden_size=channels, pose_feature_dim=channels, query_condition=True, key_value_condition=True, scale=1.0))
    blk.attention_blocks[1].set_processor(AttnProcessor(attn_type="temporal"))
    return m


class _Site(nn.Module):
    def __init__(self, channels, text_dim, epi, motion):
        super().__init__()
        self.time_proj = nn.Linear(CHANNELS[0], channels)
        self.text_proj = nn.Linear(text_dim, channels)
        self.motion_module = motion
        self.epi_module = epi


class SyncPathUNet(nn.Module):
    def __init__(self, in_channels=4, channels=CHANNELS, text_dim=768, epi_kwargs=None, motion_kwargs=None,
                 epi_factory=default_epi_factory, motion_factory=default_motion_factory, levels=None):
        super().__init__()
        self.in_channels = in_channels
        self.channels = tuple(channels[:levels] if levels else channels)
        ek = dict(EPI_KWARGS if epi_kwargs is None else epi_kwargs)
        mk = dict(MOTION_KWARGS if motion_kwargs is None else motion_kwargs)
        L = len(self.channels)
        self.conv_in = nn.Conv3d(in_channels, self.channels[0], 1)
        self.conv_out = nn.Conv3d(self.channels[0], in_channels, 1)

        def site(c):
            return _Site(c, text_dim, epi_factory(c, ek), motion_factory(c, mk))

        self.down_sites = nn.ModuleList([nn.ModuleList([site(c) for _ in range(DOWN_SITES)]) for c in self.channels])
        self.up_sites = nn.ModuleList([nn.ModuleList([site(c) for _ in range(UP_SITES)]) for c in self.channels])
        self.down_proj = nn.ModuleList([nn.Conv3d(self.channels[i], self.channels[i + 1], 1) for i in range(L - 1)])
        self.up_proj = nn.ModuleList([nn.Conv3d(self.channels[i + 1], self.channels[i], 1) for i in range(L - 1)])

    # -- one (conditioning -> motion module -> epi module) site ---------------------------------
    def _run_site(self, s, x, temb, text, F_mats, pose):
        cond = s.time_proj(temb)[None, :] + s.text_proj(text)                       # [B, C]
        x = x + cond[:, :, None, None, None]
        x = s.motion_module(x, cross_attention_kwargs={"pose_feature": pose})
        if F_mats is not None:
            x, _aux = s.epi_module(x, F_mats, None)
        return x

    def forward(self, sample, timestep, F_mats=None, encoder_hidden_states=None, pose_embedding_features=None):
        B = sample.shape[0]
        temb = timestep_embedding(timestep, CHANNELS[0], sample.device)
        if encoder_hidden_states is None:
            text = torch.zeros(B, self.down_sites[0][0].text_proj.in_features, device=sample.device)
        else:
            text = encoder_hidden_states.float().mean(dim=1)
        poses = [None] * len(self.channels) if pose_embedding_features is None else pose_embedding_features
        x = self.conv_in(sample)
        skips = []
        L = len(self.channels)
        for lvl in range(L):
            for s in self.down_sites[lvl]:
                x = self._run_site(s, x, temb, text, F_mats, poses[lvl])
                skips.append(x)
            if lvl < L - 1:
                x = F.avg_pool3d(x, (1, 2, 2))
                x = self.down_proj[lvl](x)
        for lvl in reversed(range(L)):
            for i, s in enumerate(self.up_sites[lvl]):
                if i < DOWN_SITES:
                    x = x + skips.pop()
                x = self._run_site(s, x, temb, text, F_mats, poses[lvl])
            if lvl > 0:
                x = F.interpolate(x, scale_factor=(1, 2, 2), mode="nearest")
                x = self.up_proj[lvl - 1](x)
        return self.conv_out(x)

    def randomize_zero_init(self, seed=0, gain=0.5):
        """``zero_initialize: true`` makes a freshly built epi module an exact identity (epi_module.py:81-82) and
        the pose adaptor's ``qkv_merge`` a no-op (attention_processor.py:249-251); re-initialise both so that the
        path is visible in random-init runs."""
        g = torch.Generator().manual_seed(seed)
        for name, p in self.named_parameters():
            if name.endswith("proj_out.weight") or name.endswith("qkv_merge.weight"):
                with torch.no_grad():
                    p.copy_(torch.randn(p.shape, generator=g) * (gain / math.sqrt(p.shape[1])))
        return self

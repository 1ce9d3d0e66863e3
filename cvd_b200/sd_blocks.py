"""The SD-1.5 spatial blocks the reference UNet is assembled from: per-frame ResNet blocks,
the spatial transformer (self-attention + text cross-attention + GEGLU), down / up sampling and the timestep embedding.

In the reference these are ``diffusers==0.24.0`` classes (``ResnetBlock2D``, ``Transformer2DModel`` / ``BasicTransformerBlock``,
``Downsample2D`` / ``Upsample2D``, ``Timesteps`` / ``TimestepEmbedding``) instantiated by
``animatediff/models/unet_blocks.py:6-7,205-260,350-400`` and ``unet.py:118-125``; here they are restated with the same
attribute names, so that a Stable-Diffusion / AnimateDiff checkpoint's ``state_dict`` keys load unchanged.  They are dense
convolutions and GEMM-shaped attention with no CVD-specific structure: plain library calls (cuDNN convolutions, cuBLAS
GEMMs, ``F.scaled_dot_product_attention``) run under bf16 autocast by the UNet -- the hand-written kernels of this package
are the sync / temporal modules that sit between these blocks.  Parity: the wheel is not installable in the build image,
so these restatements are UNPINNED against diffusers itself; the UNet WIRING around them is pinned to the reference's own
unet.py / unet_blocks.py executed verbatim (tests/golden/make_golden.py: unet()).
"""
import math

import torch
import torch.nn.functional as F
from torch import nn


def timestep_embedding(timesteps, dim, flip_sin_to_cos=True, freq_shift=0.0, max_period=10000):
    """diffusers ``get_timestep_embedding`` (embeddings.py): sinusoidal features of the (float) timesteps [B] -> [B, dim]."""
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device) / (half - freq_shift)
    arg = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(arg), torch.cos(arg)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return F.pad(emb, (0, 1)) if dim % 2 == 1 else emb


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels, time_embed_dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, sample):
        return self.linear_2(self.act(self.linear_1(sample)))


class ResnetBlock2D(nn.Module):
    """GroupNorm -> SiLU -> 3x3 conv -> + time embedding -> GroupNorm -> SiLU -> 3x3 conv, 1x1 shortcut when the width changes."""

    def __init__(self, in_channels, out_channels, temb_channels, eps=1e-5, groups=32, output_scale_factor=1.0):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, 1, 1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps, affine=True)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, 1, 1)
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1, 1, 0) if in_channels != out_channels else None
        self.output_scale_factor = output_scale_factor

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return (x + h) / self.output_scale_factor


class Downsample2D(nn.Module):
    def __init__(self, channels, padding=1):
        super().__init__()
        self.padding = padding
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=padding)

    def forward(self, x):
        if self.padding == 0:
            x = F.pad(x, (0, 1, 0, 1))
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x, output_size=None):
        if output_size is None:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        else:
            x = F.interpolate(x, size=output_size, mode="nearest")
        return self.conv(x)


class SpatialAttention(nn.Module):
    """diffusers ``Attention`` with its default processor: bias-free q / k / v, ``to_out = [Linear, Dropout]``."""

    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(cross_attention_dim or query_dim, inner, bias=False)
        self.to_v = nn.Linear(cross_attention_dim or query_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])
        # attention_processor.py:21-23,67-79 (AttnProcessor(attn_type="spatial", spatial_extended_attention=True)): every
        # frame also attends to the same frame of the other video(s) of its pair / group -- keys and values of the batch
        # chunks are concatenated along the token axis.  Set by UNet3DConditionModelPoseCond.set_all_attn_processor.
        self.spatial_extended_attention = False

    @staticmethod
    def _extend(t):
        """attention_processor.py:70-79, quirks kept: batches of more than 32 frames (= more than one pair of 16-frame
        videos) are cut into 32-frame chunks, otherwise into two halves; the chunks are laid side by side along the token
        axis and the result is repeated for every chunk."""
        n = t.shape[0] // 32 if t.shape[0] > 32 else 2
        return torch.cat(t.chunk(n), dim=1).repeat(n, 1, 1)

    def forward(self, x, context=None):
        B, N, _ = x.shape
        ctx = x if context is None else context
        q = self.to_q(x).view(B, N, self.heads, -1).transpose(1, 2)
        k, v = self.to_k(ctx), self.to_v(ctx)
        if self.spatial_extended_attention and k.shape[1] != 77:      # ":68": anything but a 77-token text context
            k, v = self._extend(k), self._extend(v)
        k = k.view(B, k.shape[1], self.heads, -1).transpose(1, 2)
        v = v.view(B, v.shape[1], self.heads, -1).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, N, -1)
        return self.to_out[0](o)


class _GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        h, g = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(g)


class _FeedForward(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([_GEGLU(dim, 4 * dim), nn.Dropout(0.0), nn.Linear(4 * dim, dim)])

    def forward(self, x):
        return self.net[2](self.net[0](x))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, dim_head, cross_attention_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = SpatialAttention(dim, None, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = SpatialAttention(dim, cross_attention_dim, heads, dim_head)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = _FeedForward(dim)

    def forward(self, x, context):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), context) + x
        return self.ff(self.norm3(x)) + x


class Transformer2DModel(nn.Module):
    """GroupNorm(eps 1e-6) -> 1x1 conv in -> tokens -> one BasicTransformerBlock -> 1x1 conv out -> + residual."""

    def __init__(self, heads, dim_head, in_channels, cross_attention_dim, norm_num_groups=32):
        super().__init__()
        inner = heads * dim_head
        self.norm = nn.GroupNorm(norm_num_groups, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(inner, heads, dim_head, cross_attention_dim)])
        self.proj_out = nn.Conv2d(inner, in_channels, 1)

    def forward(self, x, context):
        B, _, h, w = x.shape
        res = x
        t = self.proj_in(self.norm(x))
        inner = t.shape[1]
        t = t.permute(0, 2, 3, 1).reshape(B, h * w, inner)
        for blk in self.transformer_blocks:
            t = blk(t, context)
        t = t.reshape(B, h, w, inner).permute(0, 3, 1, 2).contiguous()
        return self.proj_out(t) + res

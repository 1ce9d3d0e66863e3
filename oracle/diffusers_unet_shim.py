"""TEST INFRASTRUCTURE ONLY -- restatement of the `diffusers==0.24.0` classes the reference UNet is assembled from
(animatediff/models/unet.py:13-22, unet_blocks.py:6-7), so that the UNMODIFIED reference ``UNet3DConditionModelPoseCond``
can be imported and executed in the build container behind ``oracle/diffusers_shim.py``:

  * ``Timesteps`` / ``TimestepEmbedding``             (diffusers.models.embeddings)
  * ``ResnetBlock2D`` / ``Downsample2D`` / ``Upsample2D`` (diffusers.models.resnet)
  * ``Transformer2DModel`` / ``BasicTransformerBlock``  (diffusers.models.transformer_2d / attention)
  * ``ModelMixin`` / ``ConfigMixin`` / ``register_to_config`` and the loader / LoRA names that are only imported.

Restated from the published source for the configuration SD-1.5 uses (GroupNorm + SiLU ResNet blocks with "default" time
embedding, convolutional ``proj_in`` / ``proj_out``, one ``BasicTransformerBlock`` with GEGLU, nearest-neighbour 2x
upsampling, stride-2 padded 3x3 downsampling).  The wheel is not installable here, so parity with diffusers itself is
UNPINNED; what these classes pin is the reference's own WIRING (unet.py / unet_blocks.py run verbatim on top of them):
the product UNet (cvd_b200/unet.py) must reproduce the reference's output when both stand on the same block arithmetic.
"""
import inspect
import math
import types

import torch
import torch.nn.functional as F
from torch import nn

from .diffusers_shim import Attention, BaseOutput, FeedForward


# ---- configuration / model mixins -----------------------------------------------------------------------------------
class _Config(dict):
    __getattr__ = dict.__getitem__


def register_to_config(init):
    """diffusers.configuration_utils.register_to_config: record the constructor arguments in ``self.config``."""
    sig = inspect.signature(init)

    def wrapped(self, *args, **kwargs):
        bound = sig.bind(self, *args, **kwargs)
        bound.apply_defaults()
        cfg = {k: v for k, v in bound.arguments.items() if k != "self"}
        init(self, *args, **kwargs)
        object.__setattr__(self, "_cfg", _Config(cfg))
    wrapped.__wrapped__ = init
    return wrapped


class ConfigMixin:
    @property
    def config(self):
        return self._cfg


class ModelMixin(nn.Module):
    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device


class _Logger:
    def info(self, *a, **k):
        pass

    warning = debug = info


logging = types.SimpleNamespace(get_logger=lambda name=None: _Logger())


class _Unused:
    """Names the reference imports but never touches on the inference path (LoRA processors, loader mixins)."""

    def __init__(self, *a, **k):
        raise NotImplementedError("not on the inference path")


class UNet2DConditionLoadersMixin:
    pass


# ---- embeddings ----------------------------------------------------------------------------------------------------
def get_timestep_embedding(timesteps, embedding_dim, flip_sin_to_cos=False, downscale_freq_shift=1.0, scale=1.0,
                           max_period=10000):
    half = embedding_dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - downscale_freq_shift)
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :] * scale
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    if embedding_dim % 2 == 1:
        emb = F.pad(emb, (0, 1, 0, 0))
    return emb


class Timesteps(nn.Module):
    def __init__(self, num_channels, flip_sin_to_cos, downscale_freq_shift):
        super().__init__()
        self.num_channels, self.flip_sin_to_cos, self.downscale_freq_shift = num_channels, flip_sin_to_cos, downscale_freq_shift

    def forward(self, timesteps):
        return get_timestep_embedding(timesteps, self.num_channels, self.flip_sin_to_cos, self.downscale_freq_shift)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels, time_embed_dim, act_fn="silu", out_dim=None):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, out_dim if out_dim is not None else time_embed_dim)

    def forward(self, sample, condition=None):
        return self.linear_2(self.act(self.linear_1(sample)))


# ---- resnet.py -----------------------------------------------------------------------------------------------------
class Upsample2D(nn.Module):
    def __init__(self, channels, use_conv=False, use_conv_transpose=False, out_channels=None, name="conv"):
        super().__init__()
        self.channels, self.out_channels, self.use_conv = channels, out_channels or channels, use_conv
        assert use_conv and not use_conv_transpose
        self.conv = nn.Conv2d(self.channels, self.out_channels, 3, padding=1)

    def forward(self, hidden_states, output_size=None, scale=1.0):
        if output_size is None:
            hidden_states = F.interpolate(hidden_states, scale_factor=2.0, mode="nearest")
        else:
            hidden_states = F.interpolate(hidden_states, size=output_size, mode="nearest")
        return self.conv(hidden_states)


class Downsample2D(nn.Module):
    def __init__(self, channels, use_conv=False, out_channels=None, padding=1, name="conv"):
        super().__init__()
        self.channels, self.out_channels, self.use_conv, self.padding = channels, out_channels or channels, use_conv, padding
        assert use_conv
        self.conv = nn.Conv2d(self.channels, self.out_channels, 3, stride=2, padding=padding)

    def forward(self, hidden_states, scale=1.0):
        if self.padding == 0:
            hidden_states = F.pad(hidden_states, (0, 1, 0, 1), mode="constant", value=0)
        return self.conv(hidden_states)


class ResnetBlock2D(nn.Module):
    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout=0.0, temb_channels=512, groups=32,
                 groups_out=None, pre_norm=True, eps=1e-6, non_linearity="swish", skip_time_act=False,
                 time_embedding_norm="default", kernel=None, output_scale_factor=1.0, use_in_shortcut=None, up=False,
                 down=False, conv_shortcut_bias=True, conv_2d_out_channels=None):
        super().__init__()
        assert time_embedding_norm == "default" and not up and not down
        out_channels = in_channels if out_channels is None else out_channels
        self.in_channels, self.out_channels, self.output_scale_factor = in_channels, out_channels, output_scale_factor
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, 1, 1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels) if temb_channels is not None else None
        self.norm2 = nn.GroupNorm(groups if groups_out is None else groups_out, out_channels, eps=eps, affine=True)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = nn.Conv2d(out_channels, conv_2d_out_channels or out_channels, 3, 1, 1)
        self.nonlinearity = nn.SiLU()
        use_sc = in_channels != (conv_2d_out_channels or out_channels) if use_in_shortcut is None else use_in_shortcut
        self.conv_shortcut = nn.Conv2d(in_channels, conv_2d_out_channels or out_channels, 1, 1, 0, bias=conv_shortcut_bias) if use_sc else None

    def forward(self, input_tensor, temb, scale=1.0):
        h = self.conv1(self.nonlinearity(self.norm1(input_tensor)))
        if self.time_emb_proj is not None:
            h = h + self.time_emb_proj(self.nonlinearity(temb))[:, :, None, None]
        h = self.conv2(self.dropout(self.nonlinearity(self.norm2(h))))
        if self.conv_shortcut is not None:
            input_tensor = self.conv_shortcut(input_tensor)
        return (input_tensor + h) / self.output_scale_factor


# ---- attention.py / transformer_2d.py ----------------------------------------------------------------------------------
class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, num_attention_heads, attention_head_dim, dropout=0.0, cross_attention_dim=None,
                 activation_fn="geglu", num_embeds_ada_norm=None, attention_bias=False, only_cross_attention=False,
                 double_self_attention=False, upcast_attention=False, norm_elementwise_affine=True, norm_type="layer_norm",
                 final_dropout=False, **unused):
        super().__init__()
        assert norm_type == "layer_norm" and num_embeds_ada_norm is None
        self.only_cross_attention = only_cross_attention
        self.norm1 = nn.LayerNorm(dim, elementwise_affine=norm_elementwise_affine)
        self.attn1 = Attention(query_dim=dim, heads=num_attention_heads, dim_head=attention_head_dim, dropout=dropout,
                               bias=attention_bias, cross_attention_dim=cross_attention_dim if only_cross_attention else None,
                               upcast_attention=upcast_attention)
        if cross_attention_dim is not None or double_self_attention:
            self.norm2 = nn.LayerNorm(dim, elementwise_affine=norm_elementwise_affine)
            self.attn2 = Attention(query_dim=dim, cross_attention_dim=cross_attention_dim if not double_self_attention else None,
                                   heads=num_attention_heads, dim_head=attention_head_dim, dropout=dropout,
                                   bias=attention_bias, upcast_attention=upcast_attention)
        else:
            self.norm2 = self.attn2 = None
        self.norm3 = nn.LayerNorm(dim, elementwise_affine=norm_elementwise_affine)
        self.ff = FeedForward(dim, dropout=dropout, activation_fn=activation_fn, final_dropout=final_dropout)

    def forward(self, hidden_states, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None,
                timestep=None, cross_attention_kwargs=None, class_labels=None):
        kw = dict(cross_attention_kwargs) if cross_attention_kwargs is not None else {}
        kw.pop("scale", None)
        n = self.norm1(hidden_states)
        hidden_states = self.attn1(n, encoder_hidden_states=encoder_hidden_states if self.only_cross_attention else None,
                                   attention_mask=attention_mask, **kw) + hidden_states
        if self.attn2 is not None:
            n = self.norm2(hidden_states)
            hidden_states = self.attn2(n, encoder_hidden_states=encoder_hidden_states, attention_mask=encoder_attention_mask,
                                       **kw) + hidden_states
        return self.ff(self.norm3(hidden_states)) + hidden_states


class Transformer2DModelOutput(BaseOutput):
    def __init__(self, sample):
        super().__init__(sample=sample)
        self.sample = sample


class Transformer2DModel(nn.Module):
    def __init__(self, num_attention_heads=16, attention_head_dim=88, in_channels=None, out_channels=None, num_layers=1,
                 dropout=0.0, norm_num_groups=32, cross_attention_dim=None, attention_bias=False, sample_size=None,
                 num_vector_embeds=None, patch_size=None, activation_fn="geglu", num_embeds_ada_norm=None,
                 use_linear_projection=False, only_cross_attention=False, double_self_attention=False, upcast_attention=False,
                 norm_type="layer_norm", norm_elementwise_affine=True, **unused):
        super().__init__()
        self.use_linear_projection = use_linear_projection
        inner = num_attention_heads * attention_head_dim
        self.in_channels = in_channels
        self.norm = nn.GroupNorm(num_groups=norm_num_groups, num_channels=in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Linear(in_channels, inner) if use_linear_projection else nn.Conv2d(in_channels, inner, 1, 1, 0)
        self.transformer_blocks = nn.ModuleList([
            BasicTransformerBlock(inner, num_attention_heads, attention_head_dim, dropout=dropout,
                                  cross_attention_dim=cross_attention_dim, activation_fn=activation_fn,
                                  attention_bias=attention_bias, only_cross_attention=only_cross_attention,
                                  double_self_attention=double_self_attention, upcast_attention=upcast_attention,
                                  norm_type=norm_type, norm_elementwise_affine=norm_elementwise_affine)
            for _ in range(num_layers)])
        self.proj_out = nn.Linear(inner, in_channels) if use_linear_projection else nn.Conv2d(inner, in_channels, 1, 1, 0)

    def forward(self, hidden_states, encoder_hidden_states=None, timestep=None, class_labels=None, cross_attention_kwargs=None,
                attention_mask=None, encoder_attention_mask=None, return_dict=True):
        batch, _, height, width = hidden_states.shape
        residual = hidden_states
        hidden_states = self.norm(hidden_states)
        if not self.use_linear_projection:
            hidden_states = self.proj_in(hidden_states)
            inner = hidden_states.shape[1]
            hidden_states = hidden_states.permute(0, 2, 3, 1).reshape(batch, height * width, inner)
        else:
            inner = hidden_states.shape[1]
            hidden_states = self.proj_in(hidden_states.permute(0, 2, 3, 1).reshape(batch, height * width, inner))
        for block in self.transformer_blocks:
            hidden_states = block(hidden_states, attention_mask=attention_mask, encoder_hidden_states=encoder_hidden_states,
                                  encoder_attention_mask=encoder_attention_mask, timestep=timestep,
                                  cross_attention_kwargs=cross_attention_kwargs, class_labels=class_labels)
        if not self.use_linear_projection:
            hidden_states = hidden_states.reshape(batch, height, width, inner).permute(0, 3, 1, 2).contiguous()
            hidden_states = self.proj_out(hidden_states)
        else:
            hidden_states = self.proj_out(hidden_states).reshape(batch, height, width, inner).permute(0, 3, 1, 2).contiguous()
        output = hidden_states + residual
        return Transformer2DModelOutput(sample=output) if return_dict else (output,)


def register(mod):
    """Called by ``diffusers_shim.install``: add the UNet-level names to the fake ``diffusers`` package."""
    mod("diffusers.configuration_utils", ConfigMixin=ConfigMixin, register_to_config=register_to_config)
    mod("diffusers.models.modeling_utils", ModelMixin=ModelMixin)
    mod("diffusers.models.embeddings", TimestepEmbedding=TimestepEmbedding, Timesteps=Timesteps)
    mod("diffusers.models.unet_2d_condition", UNet2DConditionModel=_Unused, UNet2DConditionOutput=_Unused)
    mod("diffusers.loaders", AttnProcsLayers=_Unused, UNet2DConditionLoadersMixin=UNet2DConditionLoadersMixin)
    mod("diffusers.models.resnet", ResnetBlock2D=ResnetBlock2D, Downsample2D=Downsample2D, Upsample2D=Upsample2D)
    mod("diffusers.models.transformer_2d", Transformer2DModel=Transformer2DModel)

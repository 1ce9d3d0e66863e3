"""TEST INFRASTRUCTURE ONLY -- not part of the shipped product path.

Restatement of the few `diffusers==0.24.0` classes that the reference hot path
subclasses or calls (the wheel is pinned at /root/reference/requirements.txt:2 but
is not installed in this image and is not vendored under /root/reference):

  * ``Attention``    -- base class of the reference's EpiSelfAttention
                        (animatediff/models/epi_module.py:9,350) and
                        TemporalSelfAttention (animatediff/models/motion_module.py:9,484).
  * ``FeedForward``  -- GEGLU feed-forward (epi_module.py:10,248; motion_module.py:10,444).
  * ``LoRALinearLayer`` / ``USE_PEFT_BACKEND`` -- imported at
                        animatediff/models/attention_processor.py:6,8.

Only the published behaviour that those call sites exercise is restated
(SURVEY.md section 8c lists it); the reference has no tests that pin this boundary, so
parity against diffusers itself is "unpinned" -- it is anchored on the reference's
own call sites instead.

``install()`` registers the shim under the ``diffusers.*`` module names (plus empty
``decord`` / ``imageio`` modules) so that the *unmodified* reference files can be
imported from /root/reference by ``tests/golden/make_golden.py`` in the build
container.  Nothing on the GPU box imports the reference.
"""
import sys
import types

import torch
import torch.nn.functional as F
from torch import nn


class BaseOutput(dict):
    """diffusers.utils.BaseOutput placeholder (only used as a dataclass base)."""


USE_PEFT_BACKEND = False


class LoRALinearLayer(nn.Module):
    """diffusers.models.lora.LoRALinearLayer (0.24.0): down/up projection pair."""

    def __init__(self, in_features, out_features, rank=4, network_alpha=None, device=None, dtype=None):
        super().__init__()
        self.down = nn.Linear(in_features, rank, bias=False, device=device, dtype=dtype)
        self.up = nn.Linear(rank, out_features, bias=False, device=device, dtype=dtype)
        self.network_alpha = network_alpha
        self.rank = rank
        nn.init.normal_(self.down.weight, std=1 / rank)
        nn.init.zeros_(self.up.weight)

    def forward(self, hidden_states):
        orig_dtype = hidden_states.dtype
        dtype = self.down.weight.dtype
        down = self.down(hidden_states.to(dtype))
        up = self.up(down)
        if self.network_alpha is not None:
            up = up * (self.network_alpha / self.rank)
        return up.to(orig_dtype)


class _Linear(nn.Linear):
    """LoRACompatibleLinear without LoRA: ``forward(x, scale)`` == ``F.linear``."""

    def forward(self, hidden_states, scale=1.0):
        return F.linear(hidden_states, self.weight, self.bias)


class AttnProcessor2_0:
    """diffusers.models.attention_processor.AttnProcessor2_0 (0.24.0), the default processor of ``Attention`` -- what the
    TemporalSelfAttention blocks of the CameraCtrl pose encoder run (models/pose_adaptor.py:290-300 never installs
    another one): to_q / to_k / to_v, heads split, ``F.scaled_dot_product_attention`` with 1/sqrt(d) scaling, to_out
    (Linear + Dropout), optional residual, division by ``rescale_output_factor``.  Restated from the published source
    (3-D inputs, no group / spatial norm: the only configuration on this path)."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, scale=1.0):
        residual = hidden_states
        batch_size = hidden_states.shape[0]
        if attention_mask is not None:
            seq = (hidden_states if encoder_hidden_states is None else encoder_hidden_states).shape[1]
            attention_mask = attn.prepare_attention_mask(attention_mask, seq, batch_size)
            attention_mask = attention_mask.view(batch_size, attn.heads, -1, attention_mask.shape[-1])
        query = attn.to_q(hidden_states)
        if encoder_hidden_states is None:
            encoder_hidden_states = hidden_states
        key = attn.to_k(encoder_hidden_states)
        value = attn.to_v(encoder_hidden_states)
        head_dim = key.shape[-1] // attn.heads
        query = query.view(batch_size, -1, attn.heads, head_dim).transpose(1, 2)
        key = key.view(batch_size, -1, attn.heads, head_dim).transpose(1, 2)
        value = value.view(batch_size, -1, attn.heads, head_dim).transpose(1, 2)
        hidden_states = F.scaled_dot_product_attention(query, key, value, attn_mask=attention_mask, dropout_p=0.0,
                                                       is_causal=False)
        hidden_states = hidden_states.transpose(1, 2).reshape(batch_size, -1, attn.heads * head_dim).to(query.dtype)
        hidden_states = attn.to_out[1](attn.to_out[0](hidden_states))
        if attn.residual_connection:
            hidden_states = hidden_states + residual
        return hidden_states / attn.rescale_output_factor


class Attention(nn.Module):
    """diffusers.models.attention_processor.Attention (0.24.0), the subset used.

    inner_dim = heads * dim_head, scale = dim_head ** -0.5, to_q/to_k/to_v are bias-free
    unless ``bias=True``; ``to_out = [Linear(inner_dim, query_dim, bias=True), Dropout]``.
    """

    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, dropout=0.0,
                 bias=False, upcast_attention=False, upcast_softmax=False, out_bias=True,
                 scale_qk=True, residual_connection=False, rescale_output_factor=1.0,
                 processor=None, **unused):
        super().__init__()
        self.inner_dim = dim_head * heads
        self.cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.upcast_attention = upcast_attention
        self.upcast_softmax = upcast_softmax
        self.rescale_output_factor = rescale_output_factor
        self.residual_connection = residual_connection
        self.scale = dim_head ** -0.5 if scale_qk else 1.0
        self.heads = heads
        self.group_norm = None
        self.spatial_norm = None
        self.norm_cross = None
        self.to_q = _Linear(query_dim, self.inner_dim, bias=bias)
        self.to_k = _Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_v = _Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_out = nn.ModuleList([_Linear(self.inner_dim, query_dim, bias=out_bias), nn.Dropout(dropout)])
        # diffusers: ``AttnProcessor2_0() if hasattr(F, "scaled_dot_product_attention") and self.scale_qk else AttnProcessor()``
        self.processor = processor if processor is not None else AttnProcessor2_0()

    def set_processor(self, processor):
        # diffusers pops a previously registered nn.Module processor before re-assigning
        if (hasattr(self, "processor") and isinstance(self.processor, nn.Module)
                and not isinstance(processor, nn.Module)):
            self._modules.pop("processor")
        self.processor = processor

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **cross_attention_kwargs):
        """diffusers: ``Attention.forward`` hands everything to the installed processor."""
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **cross_attention_kwargs)

    def prepare_attention_mask(self, attention_mask, target_length, batch_size, out_dim=3):
        head_size = self.heads
        if attention_mask is None:
            return attention_mask
        current_length = attention_mask.shape[-1]
        if current_length != target_length:
            attention_mask = F.pad(attention_mask, (0, target_length), value=0.0)
        if out_dim == 3:
            if attention_mask.shape[0] < batch_size * head_size:
                attention_mask = attention_mask.repeat_interleave(head_size, dim=0)
        elif out_dim == 4:
            attention_mask = attention_mask.unsqueeze(1)
            attention_mask = attention_mask.repeat_interleave(head_size, dim=1)
        return attention_mask

    def head_to_batch_dim(self, tensor, out_dim=3):
        head_size = self.heads
        batch_size, seq_len, dim = tensor.shape
        tensor = tensor.reshape(batch_size, seq_len, head_size, dim // head_size)
        tensor = tensor.permute(0, 2, 1, 3)
        if out_dim == 3:
            tensor = tensor.reshape(batch_size * head_size, seq_len, dim // head_size)
        return tensor

    def batch_to_head_dim(self, tensor):
        head_size = self.heads
        batch_size, seq_len, dim = tensor.shape
        tensor = tensor.reshape(batch_size // head_size, head_size, seq_len, dim)
        tensor = tensor.permute(0, 2, 1, 3).reshape(batch_size // head_size, seq_len, dim * head_size)
        return tensor

    def get_attention_scores(self, query, key, attention_mask=None):
        dtype = query.dtype
        if self.upcast_attention:
            query = query.float()
            key = key.float()
        if attention_mask is None:
            baddbmm_input = torch.empty(query.shape[0], query.shape[1], key.shape[1],
                                        dtype=query.dtype, device=query.device)
            beta = 0
        else:
            baddbmm_input = attention_mask
            beta = 1
        attention_scores = torch.baddbmm(baddbmm_input, query, key.transpose(-1, -2), beta=beta, alpha=self.scale)
        if self.upcast_softmax:
            attention_scores = attention_scores.float()
        attention_probs = attention_scores.softmax(dim=-1)
        return attention_probs.to(dtype)


class GEGLU(nn.Module):
    """diffusers.models.activations.GEGLU: ``h, g = proj(x).chunk(2, -1); h * gelu_erf(g)``."""

    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = _Linear(dim_in, dim_out * 2)

    def forward(self, hidden_states, scale=1.0):
        hidden_states, gate = self.proj(hidden_states).chunk(2, dim=-1)
        return hidden_states * F.gelu(gate)


class FeedForward(nn.Module):
    """diffusers.models.attention.FeedForward (0.24.0) with ``activation_fn='geglu'``:
    ``net = [GEGLU(dim, 4*dim), Dropout, Linear(4*dim, dim)]``."""

    def __init__(self, dim, dim_out=None, mult=4, dropout=0.0, activation_fn="geglu", final_dropout=False):
        super().__init__()
        inner_dim = int(dim * mult)
        dim_out = dim_out if dim_out is not None else dim
        assert activation_fn == "geglu", "only the GEGLU variant is on the reference hot path"
        self.net = nn.ModuleList([GEGLU(dim, inner_dim), nn.Dropout(dropout), _Linear(inner_dim, dim_out)])

    def forward(self, hidden_states, scale=1.0):
        for module in self.net:
            hidden_states = module(hidden_states)
        return hidden_states


class _Placeholder:
    def __init__(self, *a, **k):
        raise NotImplementedError("placeholder for a diffusers class that is off the hot path")


def install():
    """Register the shim as ``diffusers`` (+ empty decord/imageio) in ``sys.modules``."""
    if "diffusers" in sys.modules and not getattr(sys.modules["diffusers"], "_cvd_b200_shim", False):
        return  # a real diffusers is importable: leave it alone

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    root = mod("diffusers", _cvd_b200_shim=True)
    root.__path__ = []
    from . import diffusers_unet_shim as U        # the UNet-level classes (imports this module: keep the import local)
    mod("diffusers.utils", BaseOutput=BaseOutput, USE_PEFT_BACKEND=USE_PEFT_BACKEND, logging=U.logging,
        SAFETENSORS_WEIGHTS_NAME="diffusion_pytorch_model.safetensors")
    models = mod("diffusers.models")
    models.__path__ = []
    mod("diffusers.models.attention_processor", Attention=Attention, SpatialNorm=_Placeholder, AttentionProcessor=object,
        LoRAAttnProcessor=_Placeholder, AttnProcessor2_0=AttnProcessor2_0)
    mod("diffusers.models.attention", Attention=Attention, FeedForward=FeedForward,
        BasicTransformerBlock=U.BasicTransformerBlock)
    mod("diffusers.models.lora", LoRALinearLayer=LoRALinearLayer)
    mod("diffusers.models.activations", get_activation=lambda name: nn.SiLU(), GEGLU=GEGLU)
    mod("diffusers.models.normalization", AdaGroupNorm=_Placeholder)
    U.register(mod)
    for name in ("decord", "imageio"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                mod(name, VideoReader=_Placeholder)

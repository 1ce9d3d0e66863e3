"""Drop-in replacement of the AnimateDiff temporal ("motion") transformer that sits next to
the epi module in every UNet layer -- the "temporal-attention and norm/reshape glue" of the
hot path.

Mirrors ``animatediff/models/motion_module.py`` (VanillaTemporalModule :204,
TemporalTransformer3DModel :253, TemporalTransformerBlock :397, PositionalEncoding :463,
TemporalSelfAttention :484) and the two processors the shipped config installs on it
(``animatediff/models/attention_processor.py``: ``PoseAdaptorAttnProcessor`` :228-371 on
``attention_blocks.0``, ``AttnProcessor(attn_type="temporal")`` :17-110 on ``attention_blocks.1``;
``configs/inference_config.yaml:48-56``).  Same names, signatures and ``state_dict`` keys.
Inference runs on the CUDA kernels of ``libcvd_b200.so``.  Under autograd (``train()`` mode or inputs that require
grad -- the reference's training script keeps these modules frozen but still back-propagates THROUGH them into the epi
modules below, train_epi_control.py:589-630) ``forward_train`` records the same arithmetic as torch ops (bf16 GEMM
operands, fp32 residual stream, SDPA over the 16 frames).  Causal-mask variants (unused by the config) are not built.
"""
import math
from typing import Any, Dict

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib, ops
from .epi_module import FeedForward, _wants_grad, _WeightCache, zero_module


def get_motion_module(in_channels, motion_module_type: str, motion_module_kwargs: dict):
    """motion_module.py: get_motion_module."""
    if motion_module_type == "Vanilla":
        return VanillaTemporalModule(in_channels=in_channels, **motion_module_kwargs)
    raise ValueError(motion_module_type)


class AttnProcessor:
    """attention_processor.py:17-110 (plain SDPA processor), temporal flavour only."""

    def __init__(self, attn_type=None, spatial_extended_attention=False):
        if spatial_extended_attention:
            raise NotImplementedError("spatial extended attention is outside the hot path")
        self.attn_type = attn_type


class PoseAdaptorAttnProcessor(nn.Module):
    """attention_processor.py:228-371 with query_condition and key_value_condition (the only
    combination the config uses): ``x' = qkv_merge(x + pose) * scale + x`` before q/k/v."""

    def __init__(self, hidden_size, pose_feature_dim=None, cross_attention_dim=None, query_condition=False,
                 key_value_condition=False, scale=1.0, sync_lora_rank=0, network_alpha=None, sync_lora_scale=0):
        super().__init__()
        assert hidden_size == pose_feature_dim
        if not (query_condition and key_value_condition):
            raise NotImplementedError("only query_condition=key_value_condition=True is on the reference path")
        if not (sync_lora_rank == 0 or sync_lora_scale == 0):
            raise NotImplementedError("sync LoRA is unused at inference (inference_epi.py:84)")
        self.hidden_size = hidden_size
        self.scale = scale
        self.qkv_merge = nn.Linear(hidden_size, hidden_size)
        nn.init.zeros_(self.qkv_merge.weight)
        nn.init.zeros_(self.qkv_merge.bias)
        self._w = _WeightCache()


class PositionalEncoding(nn.Module):
    """motion_module.py:463-481 (sinusoidal table ``pe`` [1, max_len, d_model])."""

    def __init__(self, d_model, dropout=0., max_len=32):
        super().__init__()
        position = torch.arange(max_len).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
        pe = torch.zeros(1, max_len, d_model)
        pe[0, :, 0::2] = torch.sin(position * div_term)
        pe[0, :, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe)


class TemporalSelfAttention(nn.Module):
    """motion_module.py:484-549 (diffusers ``Attention`` subclass in the reference)."""

    def __init__(self, attention_mode=None, temporal_position_encoding=False, temporal_position_encoding_max_len=32,
                 rescale_output_factor=1.0, query_dim=None, cross_attention_dim=None, heads=8, dim_head=64, dropout=0.0,
                 bias=False, upcast_attention=False, **unused):
        super().__init__()
        assert attention_mode == "Temporal_Self"
        inner = heads * dim_head
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(cross_attention_dim or query_dim, inner, bias=bias)
        self.to_v = nn.Linear(cross_attention_dim or query_dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(dropout)])
        self.pos_encoder = PositionalEncoding(query_dim, max_len=temporal_position_encoding_max_len) \
            if temporal_position_encoding else None
        self.rescale_output_factor = rescale_output_factor
        self.processor = AttnProcessor(attn_type="temporal")
        self._w = _WeightCache()

    def set_processor(self, processor):
        if isinstance(getattr(self, "processor", None), nn.Module) and not isinstance(processor, nn.Module):
            self._modules.pop("processor")
        self.processor = processor

    def set_use_memory_efficient_attention_xformers(self, *args, **kwargs):
        pass   # motion_module.py:502-507

    def pe_table(self, frames, device):
        if self.pos_encoder is None:
            return None
        return self._w.get(("pe", frames), self.pos_encoder.pe,
                           lambda t: t[0, :frames].detach().to(device=device, dtype=torch.float32).contiguous())

    def attend(self, n, pose_feature=None, scale=None):
        """n: bf16 [T, F, C], already LayerNorm-ed and position-encoded.  pose_feature: video
        layout [B,C,F,h,w] or None.  Returns bf16 [T, F, C] without the to_out bias."""
        w = self._w
        proc = self.processor
        if isinstance(proc, PoseAdaptorAttnProcessor) and pose_feature is not None:
            xp = ops.add_pose_tokens(n, pose_feature.contiguous())
            m = F.linear(xp, proc._w.bf16("mw", proc.qkv_merge.weight), proc._w.bf16("mb", proc.qkv_merge.bias))
            n = ops.scale_add(m, n, scale or proc.scale)
        q = F.linear(n, w.bf16("q", self.to_q.weight))
        k = F.linear(n, w.bf16("k", self.to_k.weight))
        v = F.linear(n, w.bf16("v", self.to_v.weight))
        o = ops.temporal_attn(q, k, v, self.heads)
        return F.linear(o, w.bf16("o", self.to_out[0].weight))

    def attend_train(self, n, pose_tokens=None, scale=None):
        """Autograd twin of :meth:`attend` + the to_out bias / rescale: n bf16 [T, F, C] -> fp32 [T, F, C]."""
        bf = torch.bfloat16
        proc = self.processor
        if isinstance(proc, PoseAdaptorAttnProcessor) and pose_tokens is not None:
            m = F.linear(n + pose_tokens, proc.qkv_merge.weight.to(bf), proc.qkv_merge.bias.to(bf))
            n = m * (scale or proc.scale) + n
        T, Fr, C = n.shape
        H = self.heads

        def heads(t):
            return t.view(T, Fr, H, C // H).transpose(1, 2)
        q, k, v = (heads(F.linear(n, p.weight.to(bf))) for p in (self.to_q, self.to_k, self.to_v))
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(T, Fr, C)
        y = F.linear(o, self.to_out[0].weight.to(bf)).float() + self.to_out[0].bias.float()
        return y / self.rescale_output_factor


class TemporalTransformerBlock(nn.Module):
    """motion_module.py:397-460."""

    def __init__(self, dim, num_attention_heads, attention_head_dim, attention_block_types=("Temporal_Self", "Temporal_Self"),
                 dropout=0.0, norm_num_groups=32, cross_attention_dim=768, activation_fn="geglu", attention_bias=False,
                 upcast_attention=False, temporal_position_encoding=False, temporal_position_encoding_max_len=32,
                 encoder_hidden_states_query=(False, False), attention_activation_scale=1.0,
                 attention_processor_kwargs: Dict = {}, rescale_output_factor=1.0):
        super().__init__()
        self.attention_block_types = attention_block_types
        self.attention_blocks = nn.ModuleList([
            TemporalSelfAttention(attention_mode=name, cross_attention_dim=None, query_dim=dim, heads=num_attention_heads,
                                  dim_head=attention_head_dim, dropout=dropout, bias=attention_bias,
                                  upcast_attention=upcast_attention,
                                  temporal_position_encoding=temporal_position_encoding,
                                  temporal_position_encoding_max_len=temporal_position_encoding_max_len,
                                  rescale_output_factor=rescale_output_factor)
            for name in attention_block_types])
        self.norms = nn.ModuleList([nn.LayerNorm(dim) for _ in attention_block_types])
        self.ff = FeedForward(dim, dropout=dropout, activation_fn=activation_fn)
        self.ff_norm = nn.LayerNorm(dim)
        self._w = _WeightCache()

    def forward_stream(self, hs, pose_feature=None, scale=None, want_bf16=False, pending=None):
        """hs: fp32 residual stream [T, F, C], updated in place.  Each residual update is fused with the LayerNorm
        that follows it (``cvd_residual_layernorm_fwd``); ``pending`` = (y | None, bias | None, scale) still owed to
        the stream by the caller (the deferred bias of ``proj_in``)."""
        w = self._w
        frames = hs.shape[1]
        y_prev, b_prev, s_prev = pending if pending is not None else (None, None, 1.0)
        for i, (attn, norm) in enumerate(zip(self.attention_blocks, self.norms)):
            pe = attn.pe_table(frames, hs.device)
            n = ops.residual_layernorm_(hs, y_prev, b_prev, s_prev, w.f32(f"g{i}", norm.weight), w.f32(f"b{i}", norm.bias),
                                        norm.eps, pe=pe, seq_len=frames)
            y = attn.attend(n, pose_feature, scale)
            y_prev, b_prev, s_prev = y, attn._w.f32("bo", attn.to_out[0].bias), 1.0 / attn.rescale_output_factor
        n = ops.residual_layernorm_(hs, y_prev, b_prev, s_prev, w.f32("gf", self.ff_norm.weight),
                                    w.f32("bf", self.ff_norm.bias), self.ff_norm.eps)
        y = self.ff.forward_tokens(n)
        hb = ops.residual_add_(hs, y, self.ff._w.f32("b2", self.ff.net[2].bias), 1.0, want_bf16=want_bf16)
        return hs, hb

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None,
                cross_attention_kwargs: Dict[str, Any] = {}):
        """motion_module.py:447-460 on tokens [T, F, C] (how the CameraCtrl pose encoder calls the block,
        pose_adaptor.py:322-324): CUDA kernels at inference, the autograd twin when a graph is wanted."""
        if attention_mask is not None:
            raise _lib.CvdError("temporal attention masks are not supported (the reference passes None)")
        pose = cross_attention_kwargs.get("pose_feature", None)
        scale = cross_attention_kwargs.get("scale", None)
        if _wants_grad(self, hidden_states, pose):
            if not hidden_states.is_cuda:
                raise _lib.CvdError("cvd_b200 has no CPU path: hidden_states must be a CUDA tensor")
            return self.forward_train(hidden_states.float(), pose, scale).to(hidden_states.dtype)
        if pose is not None:
            raise _lib.CvdError("token-level calls take the pose feature in video layout through VanillaTemporalModule")
        hs, _ = self.forward_stream(hidden_states.float().contiguous().clone())
        return hs.to(hidden_states.dtype)

    def forward_train(self, hs, pose_tokens=None, scale=None):
        """Autograd twin of :meth:`forward_stream` (motion_module.py:447-460): fp32 stream [T, F, C] in, out."""
        bf = torch.bfloat16
        C = hs.shape[-1]
        frames = hs.shape[1]
        for attn, norm in zip(self.attention_blocks, self.norms):
            n = F.layer_norm(hs, (C,), norm.weight.float(), norm.bias.float(), norm.eps)
            if attn.pos_encoder is not None:
                n = n + attn.pos_encoder.pe[0, :frames].to(device=hs.device, dtype=torch.float32)
            hs = hs + attn.attend_train(n.to(bf), pose_tokens, scale)
        n = F.layer_norm(hs, (C,), self.ff_norm.weight.float(), self.ff_norm.bias.float(), self.ff_norm.eps).to(bf)
        proj, out = self.ff.net[0].proj, self.ff.net[2]
        hcol, gate = F.linear(n, proj.weight.to(bf), proj.bias.to(bf)).chunk(2, dim=-1)
        return hs + F.linear(hcol * F.gelu(gate), out.weight.to(bf)).float() + out.bias.float()


class TemporalTransformer3DModel(nn.Module):
    """motion_module.py:253-394."""

    def __init__(self, in_channels, num_attention_heads, attention_head_dim, num_layers,
                 attention_block_types=("Temporal_Self", "Temporal_Self"), dropout=0.0, norm_num_groups=32,
                 cross_attention_dim=320, activation_fn="geglu", attention_bias=False, upcast_attention=False,
                 temporal_position_encoding=False, temporal_position_encoding_max_len=32,
                 encoder_hidden_states_query=(False, False), attention_activation_scale=1.0,
                 attention_processor_kwargs: Dict = {}, causal_temporal_attention=None,
                 causal_temporal_attention_mask_type="", rescale_output_factor=1.0):
        super().__init__()
        assert causal_temporal_attention is not None
        if causal_temporal_attention:
            raise NotImplementedError("causal temporal masks are unused by the shipped config and not built")
        inner_dim = num_attention_heads * attention_head_dim
        self.norm = nn.GroupNorm(num_groups=norm_num_groups, num_channels=in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Linear(in_channels, inner_dim)
        self.transformer_blocks = nn.ModuleList([
            TemporalTransformerBlock(dim=inner_dim, num_attention_heads=num_attention_heads,
                                     attention_head_dim=attention_head_dim, attention_block_types=attention_block_types,
                                     dropout=dropout, norm_num_groups=norm_num_groups,
                                     cross_attention_dim=cross_attention_dim, activation_fn=activation_fn,
                                     attention_bias=attention_bias, upcast_attention=upcast_attention,
                                     temporal_position_encoding=temporal_position_encoding,
                                     temporal_position_encoding_max_len=temporal_position_encoding_max_len,
                                     encoder_hidden_states_query=encoder_hidden_states_query,
                                     attention_activation_scale=attention_activation_scale,
                                     attention_processor_kwargs=attention_processor_kwargs,
                                     rescale_output_factor=rescale_output_factor)
            for _ in range(num_layers)])
        self.proj_out = nn.Linear(inner_dim, in_channels)
        self._w = _WeightCache()

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None,
                cross_attention_kwargs: Dict[str, Any] = {}):
        assert hidden_states.dim() == 5, f"Expected hidden_states to have ndim=5, but got ndim={hidden_states.dim()}."
        if attention_mask is not None:
            raise _lib.CvdError("temporal attention masks are not supported (the reference passes None)")
        x = hidden_states.contiguous()
        B, C, Fr, h, w_ = x.shape
        w = self._w
        pose = cross_attention_kwargs.get("pose_feature", None)
        scale = cross_attention_kwargs.get("scale", None)
        if pose is not None and pose.ndim != 5:
            raise _lib.CvdError("pose_feature must be in video layout [B,C,F,h,w]")
        if _wants_grad(self, hidden_states, pose):
            return self.forward_train(x, pose, scale)
        g = ops.groupnorm_tokens(x, w.f32("gn_w", self.norm.weight), w.f32("gn_b", self.norm.bias), self.norm.num_groups,
                                 self.norm.eps, _lib.LAYOUT_BHW_F)
        # fp32 residual stream straight out of the proj_in GEMM; its bias lands in the first fused residual + LayerNorm
        hs = torch.mm(g.view(-1, g.shape[-1]), w.bf16("in_wt", self.proj_in.weight, transpose=True),
                      out_dtype=torch.float32).view(g.shape[0], g.shape[1], -1)
        pending = (None, w.f32("in_b32", self.proj_in.bias), 1.0)
        hb = None
        last = len(self.transformer_blocks) - 1
        for i, block in enumerate(self.transformer_blocks):
            hs, hb = block.forward_stream(hs, pose, scale, want_bf16=(i == last), pending=pending)
            pending = None
        if hb is None:
            hb = hs.to(torch.bfloat16)
        t = F.linear(hb, w.bf16("out_w", self.proj_out.weight))
        return ops.tokens_to_video(t, w.f32("out_b", self.proj_out.bias), x, _lib.LAYOUT_BHW_F)

    def forward_train(self, x, pose=None, scale=None):
        """Autograd-recorded twin of :meth:`forward` (motion_module.py:370-394): per-frame GroupNorm, ``b c f h w ->
        (b h w) f c``, proj_in, blocks, proj_out, inverse layout, residual.  The training script freezes this module
        but the gradients of the epi modules in front of it flow through here (train_epi_control.py:589-630)."""
        if not x.is_cuda:
            raise _lib.CvdError("cvd_b200 has no CPU path: hidden_states must be a CUDA tensor")
        bf = torch.bfloat16
        B, C, Fr, h, w_ = x.shape

        def tokens(v):      # [B,C,F,h,w] -> [(b h w), F, C]
            return v.permute(0, 3, 4, 2, 1).reshape(B * h * w_, Fr, C)
        xt = x.permute(0, 2, 1, 3, 4).reshape(B * Fr, C, h, w_)
        g = F.group_norm(xt.float(), self.norm.num_groups, self.norm.weight.float(), self.norm.bias.float(), self.norm.eps)
        tok = tokens(g.view(B, Fr, C, h, w_).permute(0, 2, 1, 3, 4)).to(bf)
        hs = F.linear(tok, self.proj_in.weight.to(bf)).float() + self.proj_in.bias.float()
        pose_tok = None if pose is None else tokens(pose).to(bf)
        for block in self.transformer_blocks:
            hs = block.forward_train(hs, pose_tok, scale)
        t = F.linear(hs.to(bf), self.proj_out.weight.to(bf)).float() + self.proj_out.bias.float()
        out = t.view(B, h, w_, Fr, C).permute(0, 4, 3, 1, 2) + x.float()
        return out.to(x.dtype)


class VanillaTemporalModule(nn.Module):
    """motion_module.py:204-250."""

    def __init__(self, in_channels, num_attention_heads=8, num_transformer_block=2, attention_block_types=("Temporal_Self",),
                 temporal_position_encoding=True, temporal_position_encoding_max_len=32, temporal_attention_dim_div=1,
                 cross_attention_dim=320, zero_initialize=True, encoder_hidden_states_query=(False, False),
                 attention_activation_scale=1.0, attention_processor_kwargs: Dict = {}, causal_temporal_attention=False,
                 causal_temporal_attention_mask_type="", rescale_output_factor=1.0):
        super().__init__()
        self.temporal_transformer = TemporalTransformer3DModel(
            in_channels=in_channels, num_attention_heads=num_attention_heads,
            attention_head_dim=in_channels // num_attention_heads // temporal_attention_dim_div,
            num_layers=num_transformer_block, attention_block_types=attention_block_types,
            cross_attention_dim=cross_attention_dim, temporal_position_encoding=temporal_position_encoding,
            temporal_position_encoding_max_len=temporal_position_encoding_max_len,
            encoder_hidden_states_query=encoder_hidden_states_query,
            attention_activation_scale=attention_activation_scale,
            attention_processor_kwargs=attention_processor_kwargs, causal_temporal_attention=causal_temporal_attention,
            causal_temporal_attention_mask_type=causal_temporal_attention_mask_type,
            rescale_output_factor=rescale_output_factor)
        if zero_initialize:
            self.temporal_transformer.proj_out = zero_module(self.temporal_transformer.proj_out)

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, attention_mask=None,
                cross_attention_kwargs: Dict[str, Any] = {}):
        return self.temporal_transformer(hidden_states, encoder_hidden_states, attention_mask,
                                         cross_attention_kwargs=cross_attention_kwargs)

"""CameraCtrl pose encoder: Pluecker embedding [b, 6, f, H, W] -> one feature map per UNet resolution, the
``pose_embedding_features`` every temporal module's first attention block mixes in.

Mirrors ``animatediff/models/pose_adaptor.py`` (``CameraPoseEncoder`` :248-329, ``ResnetBlock`` :134-167, ``Downsample``
:107-131): same constructor arguments, same ``state_dict`` keys, same output (a list of ``[(b f), c, h, w]`` maps).  It
runs ONCE per generation, in front of the denoising loop (``pipeline_animation_epi.py``: ``self.pose_encoder(pose_embedding)``),
so the 1x1 / 3x3 convolutions stay library calls (cuDNN through ``nn.Conv2d``); the temporal transformer block after every
ResNet block is this package's ``TemporalTransformerBlock`` (``cvd_residual_layernorm_fwd`` / ``cvd_temporal_attn_fwd`` /
``cvd_geglu_fwd`` kernels at inference, the autograd twin when a graph is wanted -- CameraCtrl trains this module).
There is no CPU path: ``forward`` refuses CPU tensors like the rest of the package.
"""
from typing import Sequence

import torch
from torch import nn

from . import _lib
from .epi_module import _wants_grad
from .motion_module import TemporalTransformerBlock


class Downsample(nn.Module):
    """pose_adaptor.py:107-131 (2-D only): stride-2 3x3 convolution or 2x2 average pooling."""

    def __init__(self, channels, use_conv, out_channels=None, padding=1):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        if use_conv:
            self.op = nn.Conv2d(self.channels, self.out_channels, 3, stride=2, padding=padding)
        else:
            assert self.channels == self.out_channels
            self.op = nn.AvgPool2d(kernel_size=2, stride=2)

    def forward(self, x):
        assert x.shape[1] == self.channels
        return self.op(x)


class ResnetBlock(nn.Module):
    """pose_adaptor.py:134-167."""

    def __init__(self, in_c, out_c, down, ksize=3, sk=False, use_conv=True):
        super().__init__()
        in_c, out_c = int(in_c), int(out_c)
        ps = ksize // 2
        self.in_conv = nn.Conv2d(in_c, out_c, ksize, 1, ps) if (in_c != out_c or sk is False) else None
        self.block1 = nn.Conv2d(out_c, out_c, 3, 1, 1)
        self.act = nn.ReLU()
        self.block2 = nn.Conv2d(out_c, out_c, ksize, 1, ps)
        self.skep = nn.Conv2d(in_c, out_c, ksize, 1, ps) if sk is False else None
        self.down = down
        if self.down:
            self.down_opt = Downsample(in_c, use_conv=use_conv)

    def forward(self, x):
        if self.down:
            x = self.down_opt(x)
        if self.in_conv is not None:
            x = self.in_conv(x)
        h = self.block2(self.act(self.block1(x)))
        return h + (self.skep(x) if self.skep is not None else x)


class CameraPoseEncoder(nn.Module):
    """pose_adaptor.py:248-329."""

    def __init__(self, downscale_factor, channels: Sequence[int] = (320, 640, 1280, 1280), nums_rb=3, cin=64, ksize=3, sk=False,
                 use_conv=True, compression_factor=1, temporal_attention_nhead=8, attention_block_types=("Temporal_Self",),
                 temporal_position_encoding=False, temporal_position_encoding_max_len=16, rescale_output_factor=1.0):
        super().__init__()
        self.unshuffle = nn.PixelUnshuffle(downscale_factor)
        self.channels = list(channels)
        self.nums_rb = nums_rb
        self.encoder_down_conv_blocks = nn.ModuleList()
        self.encoder_down_attention_blocks = nn.ModuleList()
        for i in range(len(self.channels)):
            convs, attns = nn.ModuleList(), nn.ModuleList()
            for j in range(nums_rb):
                mid = int(self.channels[i] / compression_factor)
                if j == 0 and i != 0:
                    in_dim, out_dim, down = self.channels[i - 1], mid, True
                elif j == 0:
                    in_dim, out_dim, down = self.channels[0], mid, False
                elif j == nums_rb - 1:
                    in_dim, out_dim, down = mid, self.channels[i], False
                else:
                    in_dim, out_dim, down = mid, mid, False
                convs.append(ResnetBlock(in_dim, out_dim, down=down, ksize=ksize, sk=sk, use_conv=use_conv))
                attns.append(TemporalTransformerBlock(
                    dim=out_dim, num_attention_heads=temporal_attention_nhead,
                    attention_head_dim=int(out_dim / temporal_attention_nhead), attention_block_types=attention_block_types,
                    dropout=0.0, cross_attention_dim=None, temporal_position_encoding=temporal_position_encoding,
                    temporal_position_encoding_max_len=temporal_position_encoding_max_len,
                    rescale_output_factor=rescale_output_factor))
            self.encoder_down_conv_blocks.append(convs)
            self.encoder_down_attention_blocks.append(attns)
        self.encoder_conv_in = nn.Conv2d(cin, self.channels[0], 3, 1, 1)

    @property
    def dtype(self) -> torch.dtype:
        return next(self.parameters()).dtype

    def _run(self, x, block_fn):
        """The layer sequence of pose_adaptor.py:311-329; ``block_fn(layer, tokens[(b h w), f, c]) -> tokens``."""
        bs, _, frames = x.shape[:3]
        x = x.permute(0, 2, 1, 3, 4).reshape(bs * frames, x.shape[1], x.shape[3], x.shape[4])      # b c f h w -> (b f) c h w
        x = self.encoder_conv_in(self.unshuffle(x))
        features = []
        for res_block, attention_block in zip(self.encoder_down_conv_blocks, self.encoder_down_attention_blocks):
            for res_layer, attention_layer in zip(res_block, attention_block):
                x = res_layer(x)
                c, h, w = x.shape[1:]
                tok = x.view(bs, frames, c, h, w).permute(0, 3, 4, 1, 2).reshape(bs * h * w, frames, c)
                tok = block_fn(attention_layer, tok)
                x = tok.view(bs, h, w, frames, c).permute(0, 3, 4, 1, 2).reshape(bs * frames, c, h, w).to(x.dtype)
            features.append(x)
        return features

    def forward_train(self, x):
        """Autograd-recorded twin (torch ops only): what a CameraCtrl training step differentiates."""
        return self._run(x, lambda layer, tok: layer.forward_train(tok.float()))

    def forward(self, x):
        if not x.is_cuda:
            raise _lib.CvdError("cvd_b200 has no CPU path: the Pluecker embedding must be a CUDA tensor")
        if _wants_grad(self, x):
            return self.forward_train(x)
        return self._run(x, lambda layer, tok: layer.forward_stream(tok.float().contiguous().clone())[0])

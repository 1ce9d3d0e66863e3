"""Data-parallel plumbing of the sync-module training step (BASELINE config 5).

The reference trains with one process per GPU under ``DistributedDataParallel`` (``train_epi_control.py:349``): both
views of a pair stay on one GPU (no K / V exchange in training), only the gradients of the trainable set -- the epi
modules, 416.8 M parameters in the shipped configuration (``:262-272``) -- are averaged across ranks after
``loss.backward()`` and before clipping / ``optimizer.step()`` (``:612-640``).  This module is that one collective,
written for NVSwitch: gradients are packed into a few large bf16 buckets (NVLS / ring all-reduce cost is per launch, not
per link, so buckets are sized for launch latency and overlap, not for link count), all-reduced asynchronously in
reverse parameter order (the order autograd finishes them in), and unpacked averaged.

    opt.zero_grad(set_to_none=True); loss.backward()
    cvd_b200.training.allreduce_gradients(trainable)            # NCCL over NVLink; gloo in the CPU tests
    torch.nn.utils.clip_grad_norm_(trainable, max_grad_norm); opt.step()

The step itself (``train_epi_control.py:245-259`` freeze, ``:548-630`` step, from the latents on -- VAE and text encoder
are out of scope) is ``freeze_for_sync_training`` + ``PoseAdaptor`` + ``training_step``; pinned on the CPU against the
reference's own UNet / PoseAdaptor executed verbatim (tests/golden/train_step.npz).
"""
from typing import Iterable, List

import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn


def trainable_epi_parameters(module) -> List[torch.nn.Parameter]:
    """The reference's trainable set (``train_epi_control.py:262-263``: every parameter that still requires grad after
    the script froze everything but the sync modules), in ``named_parameters`` order."""
    return [p for _, p in module.named_parameters() if p.requires_grad]


def gradient_buckets(params: Iterable[torch.nn.Parameter], bucket_bytes: int, elem_size: int):
    """Greedy buckets over the parameters that HAVE a gradient, in reverse order (autograd produces the gradients of the
    last layers first, so the first bucket is complete first)."""
    cap = max(1, bucket_bytes // elem_size)
    buckets, cur, n = [], [], 0
    for p in reversed([p for p in params if p.grad is not None]):
        if cur and n + p.numel() > cap:
            buckets.append(cur)
            cur, n = [], 0
        cur.append(p)
        n += p.numel()
    if cur:
        buckets.append(cur)
    return buckets


def allreduce_gradients(params: Iterable[torch.nn.Parameter], group=None, comm_dtype=torch.bfloat16,
                        bucket_bytes: int = 64 << 20):
    """Average ``p.grad`` over the ranks of ``group`` in place.  ``comm_dtype``: wire format (bf16 halves the bytes: 834 MB
    instead of 1.67 GB for the 416.8 M epi parameters; fp32 is exact).  Returns the number of buckets (collectives)."""
    world = dist.get_world_size(group)
    params = list(params)
    if world == 1:
        return 0
    buckets = gradient_buckets(params, bucket_bytes, torch.empty((), dtype=comm_dtype).element_size())
    pending = []
    for bucket in buckets:                       # pack + launch every collective, then unpack: transfers overlap packing
        flat = torch.cat([p.grad.reshape(-1).to(comm_dtype) for p in bucket])
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True)
        pending.append((bucket, flat, work))
    for bucket, flat, work in pending:
        work.wait()
        off = 0
        for p in bucket:
            n = p.numel()
            p.grad.copy_((flat[off:off + n].to(torch.float32) / world).view_as(p.grad))
            off += n
    return len(buckets)


def freeze_for_sync_training(unet, pose_encoder=None) -> List[str]:
    """``train_epi_control.py:245-259``: everything frozen except the parameters whose name contains ``epi_modules``,
    ``sync`` or ``auxiliary``.  Returns the trainable names in ``named_parameters`` order (the checkpoint's
    ``unet_trainable_dict`` keys, ``:647-653``)."""
    unet.requires_grad_(False)
    if pose_encoder is not None:
        pose_encoder.requires_grad_(False)
    names = []
    for name, param in unet.named_parameters():
        if "epi_modules" in name or "sync" in name or "auxiliary" in name:
            param.requires_grad = True
            names.append(name)
    return names


class PoseAdaptor(nn.Module):
    """``animatediff/models/pose_adaptor.py:81-104``: pose encoder in front of the UNet; returns ``(noise_pred, auxiliary)``."""

    def __init__(self, unet, pose_encoder):
        super().__init__()
        self.unet = unet
        self.pose_encoder = pose_encoder

    def forward(self, noisy_latents, timesteps, encoder_hidden_states, pose_embedding, F_mats=None, H_mats=None):
        feats = None
        if pose_embedding is not None:
            assert pose_embedding.ndim == 5
            bs = pose_embedding.shape[0]                                     # b c f h w
            feats = [x.reshape(bs, x.shape[0] // bs, *x.shape[1:]).permute(0, 2, 1, 3, 4)      # (b f) c h w -> b c f h w
                     for x in self.pose_encoder(pose_embedding)]
        out = self.unet(noisy_latents, timesteps, encoder_hidden_states, F_mats=F_mats, H_mats=H_mats,
                        pose_embedding_features=feats)
        return out.sample, out.auxiliary


def training_step(pose_adaptor, optimizer, schedule, latents, noise, timesteps, encoder_hidden_states,
                  plucker_embedding=None, F_mats=None, H_mats=None, warped_masks=None, max_grad_norm=1.0, group=None,
                  comm_dtype=torch.bfloat16):
    """One optimisation step of the sync modules, from the latents on (``train_epi_control.py:548-630``, the
    ``mixed_precision_training=False`` branch; this package's UNet applies bf16 autocast to its SD blocks itself):

    forward diffusion of ``latents`` [b, c, f, h, w] at ``timesteps`` [b] -> epsilon prediction -> MSE against ``noise``
    (both sides multiplied by ``warped_masks`` when the batch carries homographies, ``:536-545``) -> backward -> gradient
    average over ``group`` (what DistributedDataParallel does for the reference, ``:349``; skipped in a single process)
    -> clip to ``max_grad_norm`` -> ``optimizer.step()`` -> ``zero_grad(set_to_none=True)``.
    Returns ``(loss, total_grad_norm)`` as detached tensors."""
    trainable = trainable_epi_parameters(pose_adaptor)
    noisy = schedule.add_noise(latents, noise, timesteps)
    pred, auxiliary = pose_adaptor(noisy, timesteps, encoder_hidden_states, plucker_embedding, F_mats=F_mats, H_mats=H_mats)
    if auxiliary is not None:
        raise NotImplementedError("the auxiliary head (additional_channel > 0) is not built: the reference's loss for it "
                                  "(tools/consistency_losses.py, train_epi_control.py:38) is not part of the reference tree")
    mask = torch.ones_like(latents) if warped_masks is None else warped_masks
    loss = F.mse_loss(pred.float() * mask, noise.float() * mask, reduction="mean")
    loss.backward()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        allreduce_gradients(trainable, group=group, comm_dtype=comm_dtype)
    total_norm = torch.nn.utils.clip_grad_norm_(trainable, max_grad_norm)
    optimizer.step()
    optimizer.zero_grad(set_to_none=True)
    return loss.detach(), total_norm.detach()


def fold_batch(batch, latents_like=None):
    """The reshaping between the training dataset's batch and the step (``train_epi_control.py:518-545,577-582``): every
    per-frame tensor of a folded clip [b, 2n, ...] becomes the two videos of a pair stacked on the batch axis [2b, n, ...]
    (``torch.cat(x.chunk(2, dim=1), dim=0)``): ``F_mats`` / ``H_mats`` [2b, n, 3, 3]; ``plucker_embedding`` additionally to
    channels-first [2b, 6, n, H, W]; ``warped_masks`` [b, 2n, 256, 256] min-pooled 8 x 8 down to the latent grid
    [2b, 1, n, 32, 32] (a latent pixel counts only if every image pixel under it is valid).  Keys that are absent stay
    absent (None).  ``latents_like``: tensor whose dtype the matrices / masks take (``type_as(latents)``)."""
    def pair(x):
        return torch.cat(x.chunk(2, dim=1), dim=0)

    def typed(x):
        return x if latents_like is None else x.type_as(latents_like)
    out = {"F_mats": None, "H_mats": None, "warped_masks": None, "plucker_embedding": None}
    if "F_mats" in batch:
        out["F_mats"] = typed(pair(batch["F_mats"]))
    if "H_mats" in batch:
        out["H_mats"] = typed(pair(batch["H_mats"]))
        m = typed(pair(batch["warped_masks"]))                               # [2b, n, 256, 256]
        assert m.shape[-1] == 256
        b2, n, H, W = m.shape
        m = m.view(b2, n, H // 8, 8, W // 8, 8).permute(0, 1, 2, 4, 3, 5).reshape(b2, n, H // 8, W // 8, 64)
        out["warped_masks"] = m.min(dim=-1)[0][:, None].float()
    if "plucker_embedding" in batch:
        out["plucker_embedding"] = pair(batch["plucker_embedding"]).permute(0, 2, 1, 3, 4)      # b f c h w -> b c f h w
    return out

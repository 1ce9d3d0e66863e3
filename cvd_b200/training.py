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
"""
from typing import Iterable, List

import torch
import torch.distributed as dist


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

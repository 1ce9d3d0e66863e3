"""View sharding of the epi attention across the GPUs of one NVSwitch box.

Everything in the UNet is per video or per frame except the epi attention, where row r reads
the K/V of a *different view, same frame* (``attention_processor.py:573-583``;
``pipeline_animation_epi_advanced.py:621-633``).  So (view, cfg) units shard across ranks with
one exchange per epi-attention call: every rank projects K and V^T for its own rows, the
shards are all-gathered (NCCL over NVLink; row order = rank order), and the fused kernel
indexes the gathered buffer with GLOBAL partner rows.  No other collective is on the path.

The reference has no multi-GPU inference; this layout is new.
"""
import numpy as np
import torch
import torch.distributed as dist


class ViewShard:
    """Rows [rank*R_local, (rank+1)*R_local) of the global row space live on this rank."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def local_rows(self, R_global):
        if R_global % self.world != 0:
            raise ValueError(f"{R_global} rows do not split over {self.world} ranks")
        n = R_global // self.world
        return slice(self.rank * n, (self.rank + 1) * n)

    def local_kv_index(self, kv_index_global, R_global):
        """Slice a global kv_index (length k*R_global, values = global rows) to this rank's rows,
        keeping GLOBAL values: [k*R_local]."""
        idx = np.asarray(kv_index_global).reshape(-1, R_global)
        return np.ascontiguousarray(idx[:, self.local_rows(R_global)]).reshape(-1)

    def swap_halves_index(self, R_global):
        """The reference's default partner (attention_processor.py:575-576) in global rows."""
        half = np.concatenate([np.arange(R_global // 2, R_global), np.arange(0, R_global // 2)])
        return self.local_kv_index(half, R_global)

    def gather_rows(self, t):
        """All-gather along dim 0 (rank order).  NCCL: one all_gather_into_tensor; gloo (CPU tests):
        list gather + cat."""
        t = t.contiguous()
        if self.world == 1:
            return t
        if t.is_cuda:
            out = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype)
            dist.all_gather_into_tensor(out, t, group=self.group)
            return out
        parts = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(parts, t, group=self.group)
        return torch.cat(parts, dim=0)

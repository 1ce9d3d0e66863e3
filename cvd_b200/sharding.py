"""View sharding of the epi attention across the GPUs of one NVSwitch box.

Everything in the UNet is per video or per frame except the epi attention, where row r reads
the K/V of a *different view, same frame* (``attention_processor.py:573-583``;
``pipeline_animation_epi_advanced.py:621-633``).  So (view, cfg) units shard across ranks with
one exchange per epi-attention call and no other collective on the path.

Exchange mechanisms:

* ``PeerExchange`` in ``hidden`` mode (product path): the LayerNorm-ed hidden states live in symmetric memory; after one
  device-side cross-rank barrier every rank PULLS exactly the rows its queries attend to out of its peers' HBM with
  ``cvd_peer_gather_rows`` (our kernel, 16-byte loads over NVLink) and projects K / V^T of those rows locally -- half
  the bytes of exchanging K and V^T, R_local rows per call whatever the world size, no collective launch.
* ``PeerExchange`` in ``fused`` mode: K / V^T of the local rows go to symmetric memory and, at latents <= 8x8, the
  attention kernel reads the partner tiles out of the peers' HBM itself (``cvd_epi_attn_fwd_sharded``, unicast TMA over
  NVLink: no landing buffer, no copy launch).  The kernel benchmark (``bench.py --gpus N``) uses the same in-kernel reads
  for its small grids and the copy engines in row chunks under the attention for the large ones.
* ``ViewShard.gather_rows`` (baseline, and the CPU/gloo test path): NCCL all-gather of every rank's
  rows; the attention kernel then indexes the gathered buffer with GLOBAL partner rows.

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


# A device-side cross-rank barrier spins on a flag the peer writes.  With a timeout the spin ends in a trap (a CUDA error
# on the host) when a peer died or never arrives, instead of a kernel that occupies the GPU forever.
BARRIER_TIMEOUT_MS = 30000


def barrier(hdl, channel=0):
    """Symmetric-memory barrier of ``hdl``'s group on the current stream, bounded by ``BARRIER_TIMEOUT_MS``."""
    hdl.barrier(channel, BARRIER_TIMEOUT_MS)


class PeerExchange:
    """Symmetric-memory workspaces (two slots per shape, used alternately) + peer row pull.

    Protocol per call: write the local rows into the current slot (stream-ordered), ``barrier``
    (device-side, all ranks), pull the partner rows from the peers' copies of that slot.  With two
    slots one barrier per call is enough: a slot is rewritten two calls later, after a barrier that
    every rank only passes once its earlier pull (stream-ordered before it) has been issued."""

    def __init__(self, shard: ViewShard):
        import torch.distributed._symmetric_memory as symm_mem
        self.symm_mem = symm_mem
        self.shard = shard
        self.group = shard.group if shard.group is not None else dist.group.WORLD
        self._slots = {}
        self._views = {}
        self._arange = {}
        # "hidden" (default): pull the partner rows of the LayerNorm-ed hidden states with cvd_peer_gather_rows and
        # project K / V^T of those rows locally (half the bytes of K + V^T).
        # "fused": K / V^T of the local rows go to symmetric memory and, for latents <= 8x8 (one query tile per
        # (row, head): every byte crosses NVLink once), the attention kernel reads the partner tiles out of the peers'
        # HBM itself by unicast TMA; larger latents keep the hidden-row pull (several query-tile CTAs would each stream
        # the partner rows, and multicast TMA from peer memory wedges the GPU: DESIGN.md section 6).
        self.mode = "hidden"

    def alloc(self, shape, dtype, device):
        """One symmetric buffer (a collective call: every rank must allocate in the same order).
        Returns (local tensor, peer base pointers in rank order, handle)."""
        t = self.symm_mem.empty(*shape, dtype=dtype, device=device)
        hdl = self.symm_mem.rendezvous(t, self.group)
        ptrs = [hdl.get_buffer(r, tuple(shape), dtype).data_ptr() for r in range(self.shard.world)]
        return t, ptrs, hdl

    def slot(self, shape, dtype, device):
        key = (tuple(shape), dtype)
        e = self._slots.get(key)
        if e is None:
            e = {"bufs": [self.alloc(shape, dtype, device) for _ in range(2)], "next": 0}
            self._slots[key] = e
        b = e["bufs"][e["next"]]
        e["next"] ^= 1
        return b

    def slot_views(self, shape, dtype, device):
        """Next slot of ``shape`` as (local tensor, [every rank's copy of it as a tensor on THIS device, rank order],
        handle): the K / V^T sources of ``ops.epi_attn_fwd_sharded``."""
        key = (tuple(shape), dtype)
        local, ptrs, hdl = self.slot(shape, dtype, device)
        views = self._views.setdefault(key, {})
        v = views.get(local.data_ptr())
        if v is None:
            v = [hdl.get_buffer(r, tuple(shape), dtype) for r in range(self.shard.world)]
            views[local.data_ptr()] = v
        return local, v, hdl

    def pull(self, ptrs, hdl, partner, out, rows_per_src):
        """out[r] = peers' row partner[r] (GLOBAL numbering) of an already written symmetric buffer,
        after a cross-rank barrier."""
        from . import ops
        barrier(hdl)
        return ops.peer_gather_rows(ptrs, partner, out, rows_per_src=rows_per_src)

    def exchange(self, x, partner):
        """x: local rows [R_loc, ...] (any contiguous CUDA tensor).  Returns the partner rows
        [len(partner), ...] pulled from the peers (partner: int32 device tensor of GLOBAL rows)."""
        from . import ops
        local, ptrs, hdl = self.slot(x.shape, x.dtype, x.device)
        local.copy_(x)
        barrier(hdl)
        out = torch.empty((partner.shape[0],) + tuple(x.shape[1:]), device=x.device, dtype=x.dtype)
        return ops.peer_gather_rows(ptrs, partner, out, rows_per_src=x.shape[0])

    def identity_index(self, n, device):
        key = (n, device)
        if key not in self._arange:
            self._arange[key] = torch.arange(n, device=device, dtype=torch.int32)
        return self._arange[key]


# ------------------------------------------------------------------------------------------------------------------
# Host side of the per-GPU process: staging buffers next to the GPU
# ------------------------------------------------------------------------------------------------------------------
def parse_cpulist(text):
    """'0-3,8,10-11' (sysfs cpulist syntax) -> sorted list of CPU numbers."""
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return sorted(cpus)


def gpu_local_cpus(device_index, sysfs="/sys"):
    """CPUs of the NUMA node the GPU hangs off (``/sys/bus/pci/devices/<bdf>/local_cpulist``), or None when unknown."""
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"{sysfs}/bus/pci/devices/{bdf}/local_cpulist") as f:
            return parse_cpulist(f.read()) or None
    except (OSError, ValueError, AttributeError, RuntimeError, AssertionError):
        return None


def bind_host_thread_near_gpu(device_index, sysfs="/sys"):
    """Restrict the calling process to the CPUs of the GPU's NUMA node, so that the pinned staging buffers it allocates
    afterwards (first touch) and the threads that fill them sit on the socket whose PCIe root the GPU is attached to.
    With one process per GPU on a two-socket box this keeps every rank's host<->device traffic off the inter-socket link
    (round 1: eight ranks, 3 GB per step each, 5.4x slower per rank than one rank alone).  Returns the previous affinity
    (hand it to ``os.sched_setaffinity(0, prev)`` to undo) or None when nothing was changed -- never an error."""
    import os
    try:
        prev = os.sched_getaffinity(0)
        local = gpu_local_cpus(device_index, sysfs)
        want = prev & set(local or ())
        if not want or want == prev:
            return None
        os.sched_setaffinity(0, want)
        return prev
    except (OSError, AttributeError):
        return None


def restore_host_affinity(prev):
    """Undo ``bind_host_thread_near_gpu`` for EVERY thread of the process (threads created meanwhile inherited the
    restricted set; ``sched_setaffinity`` acts per thread)."""
    import os
    if prev is None:
        return
    try:
        tids = [int(t) for t in os.listdir("/proc/self/task")]
    except OSError:
        tids = [0]
    for tid in tids:
        try:
            os.sched_setaffinity(tid, prev)
        except OSError:
            pass

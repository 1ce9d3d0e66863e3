"""N-GPU check (torchrun, >= 2 ranks): the attention kernels reading the partner's K / V^T straight out of the peers' HBM
(symmetric-memory buffers passed as the K / V^T sources of cvd_epi_attn_fwd_sharded: unicast TMA over NVLink at 8x8, plain
loads at 4x4 -- the latents where a (row, head) has one query tile and the product uses in-kernel reads) give
BIT-identical results to the same call on local copies of those rows, and to the default kernel on the all-gathered
buffers.  The cluster-multicast kernels are never used with peer memory (multicast TMA from a peer mapping wedges the
GPU on driver 580.159: DESIGN.md section 6)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, torch, torch.distributed as dist
import golden_inputs as gi
from cvd_b200 import _lib, ops, sharding

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
assert world >= 2
torch.cuda.set_device(int(os.environ["LOCAL_RANK"])); dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
os.environ.pop("CVD_ATTN_CLUSTER", None)
dist.init_process_group("nccl", device_id=dev)
shard = sharding.ViewShard(); peer = sharding.PeerExchange(shard)
R, H = 32, 8
Fg = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "geometry.npz"))["F_mats"])
for feat, C, kp in ((8, 1280, 1), (4, 1280, 1), (8, 320, 1), (8, 640, 2), (4, 1280, 2)):
    N, d = feat * feat, C // H
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    q = torch.randn(R, N, C, device=dev, generator=g).to(torch.bfloat16)
    k, pk, hk = peer.alloc((R, N, C), torch.bfloat16, dev)
    vt, pv, hv = peer.alloc((R, H, ops.dv_rows(d), N), torch.bfloat16, dev)
    k.copy_(torch.randn(R, N, C, device=dev, generator=g))
    vt.zero_(); vt[:, :, :d] = torch.randn(R, H, d, N, device=dev, generator=g).to(torch.bfloat16); vt[:, :, d] = 1.0
    rowp = ops.epi_row_params(gi.tile_F(Fg, kp * R).to(dev), feat, 256.0, None, _lib.FF_HORIZONTAL)
    # partner rows (GLOBAL numbering): the next rank's rows reversed; second partner (kp = 2): the rank after that / own
    nxt = (rank + 1) % world
    idx = [nxt * R + torch.arange(R - 1, -1, -1)]
    if kp == 2:
        idx.append(((rank + 2) % world) * R + torch.arange(R))
    idx = torch.cat(idx).to(device=dev, dtype=torch.int32)
    sharding.barrier(hk)
    k_peers = [hk.get_buffer(r_, (R, N, C), torch.bfloat16) for r_ in range(world)]
    v_peers = [hv.get_buffer(r_, (R, H, ops.dv_rows(d), N), torch.bfloat16) for r_ in range(world)]
    direct = ops.epi_attn_fwd_sharded(q, k_peers, v_peers, R, rowp, idx, H, 256.0)
    local = ops.epi_attn_fwd_sharded(q, [t.clone() for t in k_peers], [t.clone() for t in v_peers], R, rowp, idx, H, 256.0)
    gathered = ops.epi_attn_fwd(q, torch.cat(k_peers), torch.cat(v_peers), rowp, idx, H, 256.0)
    torch.cuda.synchronize()
    same, same2 = torch.equal(direct, local), torch.equal(direct, gathered)
    print(f"rank {rank}: {feat}x{feat} C={C} partners={kp}: peer read == local copies: {same}; == default kernel on gathered rows: "
          f"{same2}; finite: {bool(torch.isfinite(direct.float()).all())}", flush=True)
    assert same and same2
    sharding.barrier(hk)
dist.barrier(); dist.destroy_process_group()
if rank == 0: print("DIRECT PEER READ OK")

"""2-GPU check (torchrun, 2 ranks): the attention kernels reading the partner's K / V^T straight out of the peer's HBM
(symmetric-memory buffer passed as k / vt: TMA over NVLink at 8x8, plain loads at 4x4) give bit-identical results to the
same call on a local copy of those rows."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, torch, torch.distributed as dist
import golden_inputs as gi
from cvd_b200 import _lib, ops, sharding

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
assert world == 2
torch.cuda.set_device(int(os.environ["LOCAL_RANK"])); dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=dev)
shard = sharding.ViewShard(); peer = sharding.PeerExchange(shard)
R, H = 64, 8
Fg = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "geometry.npz"))["F_mats"])
for feat, C in ((8, 1280), (4, 1280), (8, 320)):
    N, d = feat * feat, C // H
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    q = torch.randn(R, N, C, device=dev, generator=g).to(torch.bfloat16)
    k, pk, hk = peer.alloc((R, N, C), torch.bfloat16, dev)
    vt, pv, hv = peer.alloc((R, H, ops.dv_rows(d), N), torch.bfloat16, dev)
    k.copy_(torch.randn(R, N, C, device=dev, generator=g))
    vt.zero_(); vt[:, :, :d] = torch.randn(R, H, d, N, device=dev, generator=g).to(torch.bfloat16); vt[:, :, d] = 1.0
    rowp = ops.epi_row_params(gi.tile_F(Fg, R).to(dev), feat, 256.0, None, _lib.FF_HORIZONTAL)
    idx = torch.arange(R - 1, -1, -1, device=dev, dtype=torch.int32)        # partner rows on the OTHER rank, reversed
    hk.barrier()
    k_peer = hk.get_buffer(1 - rank, (R, N, C), torch.bfloat16)
    v_peer = hv.get_buffer(1 - rank, (R, H, ops.dv_rows(d), N), torch.bfloat16)
    direct = ops.epi_attn_fwd(q, k_peer, v_peer, rowp, idx, H, 256.0)
    local = ops.epi_attn_fwd(q, k_peer.clone(), v_peer.clone(), rowp, idx, H, 256.0)
    torch.cuda.synchronize()
    same = torch.equal(direct, local)
    print(f"rank {rank}: {feat}x{feat} C={C}: direct peer read == local copy: {same}; finite: {bool(torch.isfinite(direct.float()).all())}")
    assert same
    hk.barrier()
dist.barrier(); dist.destroy_process_group()
if rank == 0: print("DIRECT PEER READ OK")

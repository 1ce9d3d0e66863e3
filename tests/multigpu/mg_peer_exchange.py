"""2-GPU check of the NVLink peer exchange (run under torchrun, 2 ranks): the sharded layer output of each
rank equals the corresponding rows of the un-sharded layer computed on one GPU."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, torch, torch.distributed as dist
import cvd_b200, golden_inputs as gi
from cvd_b200 import sharding

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"])); dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
os.environ.pop("CVD_ATTN_CLUSTER", None)
dist.init_process_group("nccl", device_id=dev)
shard = sharding.ViewShard(); peer = sharding.PeerExchange(shard)
C, feat, Rg = 320, 16, 64 * world // 2 * 2 // 2 * 1   # 64 rows total for world=2 (32 per rank)
Rg = 32 * world
N = feat * feat
Fm = gi.tile_F(torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "geometry.npz"))["F_mats"]), Rg)
x = gi.hidden_tokens(Rg, N, C, 77)
kv = gi.view_pair_index(Rg // 16)          # global pairing v <-> V-1-v
def make():
    a = cvd_b200.EpiSelfAttention(attention_mode="Epi_Self", query_dim=C, heads=8, dim_head=C // 8, epi_position_encoding=True,
                                  epi_position_encoding_feat_max_size=64, epi_position_encoding_F_mat_size=256)
    a.load_state_dict(gi.attn_state_dict(C, 5), strict=False)
    return a.to(dev).eval()
ref_layer, sh_layer, ag_layer = make(), make(), make()
rows = shard.local_rows(Rg)
with torch.no_grad():
    ref, _ = ref_layer(x.to(dev), F_mats=[Fm.to(dev), kv])                    # un-sharded, every rank computes it
    sh_layer.processor.shard = shard; sh_layer.processor.peer = peer
    assert peer.mode == "hidden"                                               # product path: pull the LayerNorm-ed hidden rows
    for it in range(3):                                                        # several calls: slot ping-pong + barriers
        out, aux = sh_layer(x[rows].to(dev), F_mats=[Fm[rows].to(dev), shard.local_kv_index(kv, Rg)])
    key_ok = torch.equal(aux["key"], ref_layer(x.to(dev), F_mats=[Fm.to(dev), kv])[1]["key"][rows])
    out_h = out
    ag_layer.processor.shard = shard
    out2, _ = ag_layer(x[rows].to(dev), F_mats=[Fm[rows].to(dev), shard.local_kv_index(kv, Rg)])
torch.cuda.synchronize()
e1 = float((out - ref[rows]).abs().max()); e2 = float((out2 - ref[rows]).abs().max())
print(f"rank {rank}: max|diff| vs un-sharded: hidden-row pull = {e1:.3e}, all-gather = {e2:.3e}; aux key identical: {key_ok}")
assert e1 < 1e-6 and e2 < 1e-6 and key_ok, (e1, e2, key_ok)
# fused mode at a small latent (8x8: one query tile per (row, head)): K / V^T in symmetric memory, the attention kernel
# reads the partner tiles out of the peer's HBM itself (unicast TMA over NVLink)
feat8, N8 = 8, 64
x8 = gi.hidden_tokens(Rg, N8, C, 78)
ref8_layer, f8_layer = make(), make()
f8_layer.processor.shard = shard; f8_layer.processor.peer = peer
with torch.no_grad():
    ref8, _ = ref8_layer(x8.to(dev), F_mats=[Fm.to(dev), kv])
    peer.mode = "fused"
    for it in range(3):
        out8, aux8 = f8_layer(x8[rows].to(dev), F_mats=[Fm[rows].to(dev), shard.local_kv_index(kv, Rg)])
    peer.mode = "hidden"
torch.cuda.synchronize()
e8 = float((out8 - ref8[rows]).norm() / ref8[rows].norm())      # the K GEMM writes into symmetric memory (mm with out=): same math,
print(f"rank {rank}: fused in-kernel exchange at 8x8: rel-L2 vs un-sharded = {e8:.3e}")     # possibly another cuBLAS kernel
assert e8 < 5e-3, e8
dist.barrier(); dist.destroy_process_group()
if rank == 0: print("PEER EXCHANGE OK")

"""Per-kernel GPU time of one EpiModule training step (forward + backward), torch.profiler.
    python scripts/profile_train_kernels.py [C] [feat] [pairs]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, torch
import golden_inputs as gi, cvd_b200
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
C = int(sys.argv[1]) if len(sys.argv) > 1 else 320
feat = int(sys.argv[2]) if len(sys.argv) > 2 else 32
B = 2 * (int(sys.argv[3]) if len(sys.argv) > 3 else 8)
m = cvd_b200.EpiModule(C, num_attention_heads=8, num_transformer_block=1, attention_block_types=["Epi_Self", "Epi_Self"],
                       epi_position_encoding=True, epi_position_encoding_feat_max_size=64, epi_position_encoding_F_mat_size=256,
                       epi_rand_slope_ff=False, zero_initialize=True)
m.load_state_dict(gi.module_state_dict(C, 1), strict=False); m = m.to(dev).train()
x = gi.hidden_video(B, C, 16, feat, feat, 2).to(dev)
w = gi.hidden_video(B, C, 16, feat, feat, 3).to(dev)
Fm = gi.tile_F(torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "geometry.npz"))["F_mats"]), B * 16).reshape(B, 16, 3, 3).to(dev)
def step():
    xp = x.clone().requires_grad_(True)
    y, _ = m(xp, F_mats=Fm)
    (y * w).sum().backward()
for _ in range(2): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"C={C} feat={feat} rows={B*16}: {tot/3/1e3:.2f} ms of kernels per step")
for e in rows[:24]:
    print(f"{e.device_time_total/3/1e3:8.3f} ms  x{e.count//3:3d}  {e.key[:120]}")

import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests/golden")
import numpy as np, torch
import golden_inputs as gi, cvd_b200
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
C, feat, B = int(sys.argv[1]), int(sys.argv[2]), 4
m = cvd_b200.EpiModule(C, num_attention_heads=8, num_transformer_block=1, attention_block_types=["Epi_Self", "Epi_Self"],
                       epi_position_encoding=True, epi_position_encoding_feat_max_size=64, epi_position_encoding_F_mat_size=256,
                       epi_rand_slope_ff=True, zero_initialize=True)
m.load_state_dict(gi.module_state_dict(C, 1), strict=False); m = m.to(dev).eval()
x = gi.hidden_video(B, C, 16, feat, feat, 2).to(dev)
Fm = gi.tile_F(torch.from_numpy(np.load("/root/repo/tests/golden/geometry.npz")["F_mats"]), B * 16).reshape(B, 16, 3, 3).to(dev)
with torch.no_grad():
    for _ in range(3): m(x, F_mats=Fm)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5): m(x, F_mats=Fm)
        torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"C={C} feat={feat} R={B*16}: total {tot/5:.0f} us per forward")
for e in rows[:16]:
    print(f"{e.device_time_total/5:8.1f} us  x{e.count//5:2d}  {e.key[:110]}")

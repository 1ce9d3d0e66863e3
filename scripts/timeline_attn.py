"""Per-iteration clock64 breakdown of the softmax warp (debug build of the bench-shape kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from cvd_b200 import _lib, ops
R, feat, C, H = 64, 32, 320, 8
N, d = feat * feat, C // H
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
q = torch.randn(R, N, C, device=dev, generator=g).to(torch.bfloat16)
k = torch.randn(R, N, C, device=dev, generator=g).to(torch.bfloat16)
vt = torch.zeros(R, H, ops.dv_rows(d), N, device=dev, dtype=torch.bfloat16)
vt[:, :, :d] = torch.randn(R, H, d, N, device=dev, generator=g).to(torch.bfloat16); vt[:, :, d] = 1.0
Fm = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "geometry.npz"))["F_mats"])
Fm = torch.cat([Fm[:16], Fm[:16], Fm[16:], Fm[16:]]).contiguous().to(dev)
partner = torch.cat([torch.arange(R // 2, R), torch.arange(0, R // 2)]).to(torch.int32).to(dev)
rowp = ops.epi_row_params(Fm, feat, 256.0, None, _lib.FF_HORIZONTAL)
grid = R * H * (N // 128)
buf = torch.zeros(grid * 128 * 64 + 2 * grid * 32 * 8 * 2, device=dev, dtype=torch.float32)
for _ in range(2):
    ops.epi_attn_fwd(q, k, vt, rowp, partner, H, 256.0, dbg_s=buf)
torch.cuda.synchronize()
allst = buf[grid * 128 * 64:].view(torch.int64).view(2, grid, 32, 8).cpu().numpy()
st = allst[0][:, :16, :6]
mm = allst[1][:, :15, :5]
names = ["wait s_full", "half0 ld+score+exp", "half1 ld+score+exp", "wait pv_done", "P st+arrive"]
d_ = np.diff(st, axis=2)
gap = st[:, 1:, 0] - st[:, :-1, 5]
life = st[:, 15, 5] - st[:, 1, 0]
for sel, label in ((slice(0, 444), "first wave"), (slice(1000, 3000), "steady CTAs")):
    print(label)
    for i, nm in enumerate(names):
        print(f"  {nm:22s} t1..15 mean={d_[sel, 1:, i].mean():8.0f}  p90={np.percentile(d_[sel, 1:, i], 90):8.0f}")
    print(f"  loop gap               mean={gap[sel].mean():8.0f}")
    print(f"  tiles 1..15 span       mean={life[sel].mean():8.0f}  per-tile={life[sel].mean() / 15:8.0f}")
md = np.diff(mm, axis=2)
mnames = ["wait s_free", "issue QK(j+1) (wait k + 3 MMA + 2 commits)", "wait v_full/p_full", "issue PV (4 MMA + commits)"]
sel = slice(1000, 3000)
print("MMA warp (steady CTAs)")
for i, nm in enumerate(mnames):
    print(f"  {nm:44s} j1..13 mean={md[sel, 1:14, i].mean():8.0f}  p90={np.percentile(md[sel, 1:14, i], 90):8.0f}")
print(f"  per-tile period                              mean={(mm[sel, 14, 0] - mm[sel, 1, 0]).mean() / 13:8.0f}")

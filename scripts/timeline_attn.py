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
st = allst[0][:, :32, :6]
mm = allst[1][:, :30, :5]
names = ["wait s_full", "tmem ld", "score+max", "exp+pack(+rescale)", "P st+arrive"]
d_ = np.diff(st, axis=2)                       # [grid,16,6]
gap = st[:, 1:, 0] - st[:, :-1, 5]             # loop overhead between iterations
life = st[:, 31, 5] - st[:, 2, 0]
for sel, label in ((slice(0, 296), "first wave"), (slice(1000, 3000), "steady CTAs")):
    print(label)
    for i, nm in enumerate(names):
        print(f"  {nm:22s} it0={d_[sel, 0, i].mean():8.0f}  t2..31 mean={d_[sel, 2:, i].mean():8.0f}  p90={np.percentile(d_[sel, 2:, i], 90):8.0f}")
    print(f"  loop gap               mean={gap[sel].mean():8.0f}")
    print(f"  sub-tiles 2..31 span   mean={life[sel].mean():8.0f}  per-subtile={life[sel].mean() / 30:8.0f}")

md = np.diff(mm, axis=2)
mnames = ["wait v/p_full", "issue PV (2 MMA)", "commits", "issue QK(t+2) (wait k + 3 MMA + commits)"]
sel = slice(1000, 3000)
print("MMA thread (steady CTAs)")
for i, nm in enumerate(mnames):
    print(f"  {nm:42s} t2..29 mean={md[sel, 2:, i].mean():8.0f}  p90={np.percentile(md[sel, 2:, i], 90):8.0f}")
print(f"  per-subtile period                          mean={(mm[sel, 29, 0] - mm[sel, 2, 0]).mean() / 27:8.0f}")

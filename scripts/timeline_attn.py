"""Per-tile clock64 breakdown of the softmax and MMA warps (debug build of the bench-shape kernel)."""
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
items = R * H * (N // 128)
grid = min(items, 148 * 2)
buf = torch.zeros(items * 128 * 64 + 2 * grid * 64 * 8 * 2, device=dev, dtype=torch.float32)
for _ in range(2):
    ops.epi_attn_fwd(q, k, vt, rowp, partner, H, 256.0, dbg_s=buf)
torch.cuda.synchronize()
allst = buf[items * 128 * 64:].view(torch.int64).view(2, grid, 64, 8).cpu().numpy()
st = allst[0][:, 8:56, :6]
mm = allst[1][:, 8:56, :5]
names = ["wait s_full", "half0 ld+score+exp", "half1 ld+score+exp", "epilogue?+wait p_free", "P st+arrive"]
d_ = np.diff(st, axis=2)
gap = st[:, 1:, 0] - st[:, :-1, 5]
print("softmax thread 0, stream tiles 8..55")
for i, nm in enumerate(names):
    print(f"  {nm:24s} mean={d_[:, :, i].mean():8.0f}  p50={np.percentile(d_[:, :, i], 50):8.0f}  p90={np.percentile(d_[:, :, i], 90):8.0f}")
print(f"  loop gap                 mean={gap.mean():8.0f}")
print(f"  per-tile period          mean={(st[:, -1, 0] - st[:, 0, 0]).mean() / (st.shape[1] - 1):8.0f}")
md = np.diff(mm, axis=2)
mnames = ["QK waits (q/k/s_free)", "QK issue (3 MMA + 2 commits)", "PV waits (v_full/p_full)", "PV issue (4 MMA + 2 commits)"]
print("MMA warp, stream steps 8..55")
for i, nm in enumerate(mnames):
    print(f"  {nm:30s} mean={md[:, :, i].mean():8.0f}  p50={np.percentile(md[:, :, i], 50):8.0f}  p90={np.percentile(md[:, :, i], 90):8.0f}")
print(f"  per-step period                mean={(mm[:, -1, 0] - mm[:, 0, 0]).mean() / (mm.shape[1] - 1):8.0f}")
# lead of S over the softmax: time between QK issue of tile s and the softmax's wait for it
lead = allst[0][:, 8:56, 0] - allst[1][:, 8:56, 2]
print(f"  QK issued -> softmax asks for it: mean={lead.mean():8.0f} p10={np.percentile(lead, 10):8.0f}")

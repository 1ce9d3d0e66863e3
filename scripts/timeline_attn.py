"""Hand-off latencies of the attention pipeline from clock64 stamps (what-if build, CVD_ATTN_VARIANT=11 skeleton
or 12 full kernel; make -C cvd_b200/csrc EXTRA=-DCVD_ATTN_WHATIF)."""
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
buf = torch.zeros(items * 128 * 64 + 2 * grid * 64 * 8 * 2 + grid * 4 * 16, device=dev, dtype=torch.float32)
for _ in range(2):
    ops.epi_attn_fwd(q, k, vt, rowp, partner, H, 256.0, dbg_s=buf)
torch.cuda.synchronize()
tail = buf[items * 128 * 64:]
allst = tail[:2 * grid * 64 * 8 * 2].view(torch.int64).view(2, grid, 64, 8).cpu().numpy().astype(np.float64)
sm, mm = allst[0], allst[1]
lo, hi = 10, 50
def stat(name, x):
    x = x[np.isfinite(x)]
    print(f"  {name:58s} mean={x.mean():8.0f}  p10={np.percentile(x, 10):8.0f}  p50={np.percentile(x, 50):8.0f}  p90={np.percentile(x, 90):8.0f}")
print("softmax thread 0 (stream tiles %d..%d)" % (lo, hi))
stat("tile period", sm[:, lo + 1:hi + 1, 0] - sm[:, lo:hi, 0])
stat("start -> first block issued", sm[:, lo:hi, 1] - sm[:, lo:hi, 0])
stat("publish P(g-1) (wait p_free, st, fence, arrive)", sm[:, lo:hi, 2] - sm[:, lo:hi, 1])
stat("-> S(g) released", sm[:, lo:hi, 3] - sm[:, lo:hi, 2])
stat("second block issued", sm[:, lo:hi, 4] - sm[:, lo:hi, 3])
stat("wait s_full(g+1)", sm[:, lo:hi, 5] - sm[:, lo:hi, 4])
print("Q K^T issuer")
stat("step period", mm[:, lo + 1:hi + 1, 0] - mm[:, lo:hi, 0])
stat("waits (q/k/s_free)", mm[:, lo:hi, 1] - mm[:, lo:hi, 0])
stat("issue 3 MMA + commits", mm[:, lo:hi, 2] - mm[:, lo:hi, 1])
print("P V issuer")
stat("waits (v_full/p_full)", mm[:, lo:hi, 5] - mm[:, lo:hi, 4])
stat("issue 4 MMA + commits", mm[:, lo:hi, 6] - mm[:, lo:hi, 5])
print("hand-offs")
stat("S(g) released -> QK(g+2) waits satisfied", mm[:, lo + 2:hi + 2, 1] - sm[:, lo:hi, 3])
stat("QK(g+2) issued -> softmax could see S(g+2) full (lower bound: asks later)", sm[:, lo + 1:hi + 1, 5] - mm[:, lo + 2:hi + 2, 2])
stat("P(g-1) published -> PV(g-1) waits satisfied", mm[:, lo - 1:hi - 1, 5] - sm[:, lo:hi, 2])
stat("QK(s) issued - tile s start (lead of the score MMA, >0 = late)", mm[:, lo:hi, 2] - sm[:, lo:hi, 0])
if os.environ.get("CVD_ATTN_VARIANT") == "12":
    acc = tail[2 * grid * 64 * 8 * 2:].view(torch.int32).view(grid, 4, 16).cpu().numpy().astype(np.float64)
    tiles = acc[:, :, 8].mean()
    print(f"softmax-side waits that really blocked (per warp, {tiles:.0f} tiles per CTA):")
    for i, nm in enumerate(["s_full(g+1)", "p_free", "o_done"]):
        print(f"  {nm:12s} blocked {acc[:, :, i].mean():7.1f} times, {acc[:, :, 4 + i].mean() / tiles:7.1f} cycles per tile;  by warp: " +
              " ".join(f"{acc[:, w, i].mean():6.1f}/{acc[:, w, 4 + i].mean() / tiles:5.0f}" for w in range(4)))

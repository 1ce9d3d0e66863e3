"""Hand-off timeline of the production attention kernel (epi_attn_fwd_kernel<40,32,DBG=true>) from the clock64 stamps
of its debug build: softmax thread 0 and the producer/issuer lane of every CTA, 16 key tiles per CTA.

    python scripts/timeline_attn_v4.py
"""
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
st = buf[grid * 128 * 64:].view(torch.int64).view(2, grid, 32, 8).cpu().numpy().astype(np.float64)
sel = slice(grid // 4, 3 * grid // 4)          # CTAs of the steady part of the launch
sm, mm = st[0, sel], st[1, sel]
lo, hi = 4, 14
def stat(name, x):
    print(f"  {name:64s} mean={x.mean():7.0f}  p10={np.percentile(x, 10):7.0f}  p50={np.percentile(x, 50):7.0f}  p90={np.percentile(x, 90):7.0f}")
print("softmax thread 0, tiles %d..%d of 16" % (lo, hi))
stat("tile period", sm[:, lo + 1:hi + 1, 0] - sm[:, lo:hi, 0])
stat("wait s_full(j)", sm[:, lo:hi, 1] - sm[:, lo:hi, 0])
stat("ld S + first half of the math", sm[:, lo:hi, 2] - sm[:, lo:hi, 1])
stat("second half of the math", sm[:, lo:hi, 3] - sm[:, lo:hi, 2])
stat("wait pv_done(j-1)", sm[:, lo:hi, 4] - sm[:, lo:hi, 3])
stat("st P + fence + arrive p_full", sm[:, lo:hi, 5] - sm[:, lo:hi, 4])
print("producer / issuer warp")
stat("iteration period", mm[:, lo + 1:hi + 1, 0] - mm[:, lo:hi, 0])
stat("K refill + wait s_free(j)", mm[:, lo:hi, 1] - mm[:, lo:hi, 0])
stat("issue QK(j+1) (k_full wait, 3 MMA, 2 commits)", mm[:, lo:hi, 2] - mm[:, lo:hi, 1])
stat("V refill + wait v_full / p_full(j)", mm[:, lo:hi, 3] - mm[:, lo:hi, 2])
stat("issue PV(j) (4 MMA, 2-3 commits)", mm[:, lo:hi, 4] - mm[:, lo:hi, 3])
print("hand-offs")
stat("issuer: QK(j+1) issued -> softmax passes s_full(j+1)", sm[:, lo + 1:hi + 1, 1] - mm[:, lo:hi, 2])
stat("   of which the softmax thread was already waiting (>0)", np.maximum(0, mm[:, lo:hi, 2] - sm[:, lo + 1:hi + 1, 0]))
stat("softmax: p_full(j) arrived -> issuer passes the p_full wait", mm[:, lo:hi, 3] - sm[:, lo:hi, 5])
stat("issuer: PV(j) issued -> softmax passes pv_done(j) wait of tile j+1", sm[:, lo + 1:hi + 1, 4] - mm[:, lo:hi, 4])
stat("tile start -> issuer passes s_free(j)", mm[:, lo:hi, 1] - sm[:, lo:hi, 1])

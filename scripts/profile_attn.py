"""Launch the fused epi-attention kernel a few times at the bench shape (for ncu / quick timing).

    python scripts/profile_attn.py [feat] [C] [R] [iters]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np
import torch

from cvd_b200 import _lib, ops

feat = int(sys.argv[1]) if len(sys.argv) > 1 else 32
C = int(sys.argv[2]) if len(sys.argv) > 2 else 320
R = int(sys.argv[3]) if len(sys.argv) > 3 else 64
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 5
H, N, d = 8, feat * feat, C // 8
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
q = torch.randn(R, N, C, device=dev, generator=g).to(torch.bfloat16)
k = torch.randn(R, N, C, device=dev, generator=g).to(torch.bfloat16)
vt = torch.zeros(R, H, ops.dv_rows(d), N, device=dev, dtype=torch.bfloat16)
vt[:, :, :d] = torch.randn(R, H, d, N, device=dev, generator=g).to(torch.bfloat16)
vt[:, :, d] = 1.0
Fm = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "geometry.npz"))["F_mats"])
Fm = torch.cat([Fm[:16], Fm[:16], Fm[16:], Fm[16:]]).repeat((R + 63) // 64, 1, 1)[:R].contiguous().to(dev)
partner = torch.cat([torch.arange(R // 2, R), torch.arange(0, R // 2)]).to(torch.int32).to(dev)
rowp = ops.epi_row_params(Fm, feat, 256.0, None, _lib.FF_HORIZONTAL)
for _ in range(3):
    ops.epi_attn_fwd(q, k, vt, rowp, partner, H, 256.0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if os.environ.get("CVD_PROFILE_GRAPH", "1") == "1":
    # the launches are captured in a CUDA graph so that small shapes are not bounded by the Python launch rate
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for _ in range(iters):
                ops.epi_attn_fwd(q, k, vt, rowp, partner, H, 256.0)
    graph.replay()
    torch.cuda.synchronize()
    e0.record()
    graph.replay()
    e1.record()
else:
    e0.record()
    for _ in range(iters):
        ops.epi_attn_fwd(q, k, vt, rowp, partner, H, 256.0)
    e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"feat={feat} C={C} R={R}: {ms * 1e3:.1f} us/launch, {4.0 * R * N * N * C / ms / 1e9:.1f} TFLOP/s")

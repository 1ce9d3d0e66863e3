"""Time the attention-core forward (+lse) and backward at the training shapes (BASELINE config 5: 2-view pairs,
batch 8 per GPU -> R = 256 rows) and beside torch autograd through the fp32 dense restatement on the same GPU.

    python scripts/profile_bwd.py [R]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np
import torch

import golden_inputs as gi
from cvd_b200 import _lib, ops
from gpu_util import attention_ref

dev = torch.device("cuda:0")
R = int(sys.argv[1]) if len(sys.argv) > 1 else 256
H = 8


def timed(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for feat, C in ((32, 320), (16, 640), (8, 1280), (4, 1280)):
    N, d = feat * feat, C // H
    g = torch.Generator(device=dev).manual_seed(0)
    q, k, v, do = (torch.randn(R, N, C, device=dev, generator=g).to(torch.bfloat16) for _ in range(4))
    Fm = gi.tile_F(torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "geometry.npz"))["F_mats"]), R).to(dev)
    partner = torch.cat([torch.arange(R // 2, R), torch.arange(0, R // 2)]).to(torch.int32).to(dev)
    rowp = ops.epi_row_params(Fm, feat, 256.0, None, _lib.FF_HORIZONTAL)
    vt = ops.padded_vt(v, H)
    out, lse = ops.epi_attn_fwd_lse(q, k, vt, rowp, partner, H, 256.0)
    t_f = timed(lambda: ops.epi_attn_fwd_lse(q, k, vt, rowp, partner, H, 256.0), 10)
    t_b = timed(lambda: ops.epi_attn_bwd(q, k, v, rowp, partner, out, do, lse, H, 256.0), 5)
    flops_f = 4.0 * R * N * N * C
    line = (f"R={R} N={N:5d} C={C:5d} d={d:3d}: forward+lse {t_f * 1e3:8.1f} us ({flops_f / t_f / 1e9:6.1f} TFLOP/s)   "
            f"backward {t_b * 1e3:9.1f} us ({2.5 * flops_f / t_b / 1e9:6.1f} TFLOP/s algorithmic, 5 GEMMs)")
    if R * H * N * N * 4 <= 6e9:      # dense fp32 autograd reference fits comfortably
        bias = ops.epi_bias_dense(rowp, feat, 256.0)

        def ref():
            qf, kf, vf = (a.float().requires_grad_(True) for a in (q, k, v))
            attention_ref(qf, kf, vf, H, bias, partner).backward(do.float())
        t_r = timed(ref, 2)
        line += f"   torch fp32 autograd fwd+bwd {t_r:8.2f} ms -> {t_r / (t_f + t_b):5.1f}x"
    print(line, flush=True)

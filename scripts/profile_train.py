"""Sync-module training step (forward + backward through one EpiModule, SURVEY.md section 8 row f3 / BASELINE config 5)
timed beside the reference's own PyTorch GPU path: the oracle restatement of the reference module moved to the GPU
unchanged (fp32 eager and under bf16 autocast as train_epi_control.py:589 does with fp16), torch autograd backward.

    python scripts/profile_train.py [pairs]        (pairs per GPU: 1 = reference default, 8 = BASELINE config 5)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np
import torch

import golden_inputs as gi
import cvd_b200
from oracle import epi_oracle as O

dev = torch.device("cuda:0")
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1
B = 2 * pairs


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


Fg = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "geometry.npz"))["F_mats"])
for C, feat in ((320, 32), (640, 16), (1280, 8), (1280, 4)):
    sd = gi.module_state_dict(C, 9000 + C + feat)
    x = gi.hidden_video(B, C, 16, feat, feat, 1).to(dev)
    wgt = gi.hidden_video(B, C, 16, feat, feat, 2).to(dev)
    Fm = gi.tile_F(Fg, B * 16).reshape(B, 16, 3, 3).to(dev)
    m = cvd_b200.EpiModule(C, num_attention_heads=8, num_transformer_block=1, attention_block_types=["Epi_Self", "Epi_Self"],
                           epi_position_encoding=True, epi_position_encoding_feat_max_size=64,
                           epi_position_encoding_F_mat_size=256, epi_rand_slope_ff=False, zero_initialize=True)
    m.load_state_dict(sd, strict=False)
    m = m.to(dev).train()

    def ours():
        xp = x.clone().requires_grad_(True)
        y, _ = m(xp, F_mats=Fm)
        (y * wgt).sum().backward()
        return y, xp.grad

    sd_dev = {k: v.to(dev).requires_grad_(True) for k, v in sd.items()}

    def ref(autocast):
        xp = x.clone().requires_grad_(True)
        with torch.device(dev), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            y, _ = O.epi_module(xp, sd_dev, 8, Fm, 256)
        (y.float() * wgt).sum().backward()
        return y, xp.grad

    y_o, gx_o = ours()
    line = f"C={C:5d} {feat:2d}x{feat:<2d} rows={B * 16:4d}: cvd_b200 fwd+bwd {timed(ours, 5):8.2f} ms"
    try:
        y_r, gx_r = ref(False)
        rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
        line += (f"   reference fp32 {timed(lambda: ref(False), 2):8.2f} ms, bf16 autocast {timed(lambda: ref(True), 2):8.2f} ms"
                 f"   (delta rel-L2 {rel(y_o - x, y_r - x):.1e}, d/dx {rel(gx_o, gx_r):.1e})")
    except torch.cuda.OutOfMemoryError:
        line += "   reference: out of memory"
    print(line, flush=True)
    del sd_dev
    torch.cuda.empty_cache()

"""Helpers shared by the -m gpu tests."""
import os

import numpy as np
import torch

import golden_inputs as gi  # noqa: F401


def t(a):
    return torch.from_numpy(np.asarray(a))


def test_device():
    """``cuda:0`` on the B200 box.  With CVD_TEST_EMULATE=1 in the environment (build box, no GPU) the -m gpu tests run on
    the CPU instead, tests/conftest.py having replaced the C-ABI kernels by the plain-torch test double of
    tests/emulated_ops.py in strict device-shape mode: a dry run of the test code and of the host logic it drives, never
    a parity claim."""
    if os.environ.get("CVD_TEST_EMULATE") == "1":
        return torch.device("cpu")
    assert torch.cuda.is_available(), "these tests need the B200"
    return torch.device("cuda:0")


test_device.__test__ = False


def make_vt(v, heads):
    """v [R,N,C] -> bf16 [R,H,DV,N]: V transposed per head + a row of ones (see include/cvd_b200.h)."""
    R, N, C = v.shape
    d = C // heads
    DV = (d // 16 + 1) * 16
    vt = torch.zeros(R, heads, DV, N, device=v.device, dtype=torch.bfloat16)
    vt[:, :, :d] = v.view(R, N, heads, d).permute(0, 2, 3, 1).to(torch.bfloat16)
    vt[:, :, d] = 1.0
    return vt.contiguous()


def attention_ref(q, k, v, heads, bias, partner):
    """fp32 torch restatement on the device: softmax(q k^T/sqrt(d) + bias) v with K/V rows gathered
    by partner [k*R] and bias [k*R,N,N]; keys of all partners share one softmax."""
    R, N, C = q.shape
    d = C // heads
    kp = partner.shape[0] // R
    qh = q.float().view(R, N, heads, d).transpose(1, 2)
    ss, vs = [], []
    for p in range(kp):
        idx = partner[p * R:(p + 1) * R].long()
        kh = k.float()[idx].view(R, N, heads, d).transpose(1, 2)
        s = torch.matmul(qh, kh.transpose(-1, -2)) * d ** -0.5
        if bias is not None:
            s = s + bias[p * R:(p + 1) * R, None]
        ss.append(s)
        vs.append(v.float()[idx].view(R, N, heads, d).transpose(1, 2))
    pr = torch.softmax(torch.cat(ss, dim=-1), dim=-1)
    o = torch.matmul(pr, torch.cat(vs, dim=-2))
    return o.transpose(1, 2).reshape(R, N, C)


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def assert_close(actual, ref, rel=2e-2, max_abs=None, what=""):
    """Tolerance of the bf16 tensor-core path against an fp32 reference: relative L2 error and a
    bound on the largest element error in units of the reference's RMS."""
    actual, ref = actual.detach().float().cpu(), ref.detach().float().cpu()
    assert actual.shape == ref.shape, (actual.shape, ref.shape)
    assert torch.isfinite(actual).all(), f"{what}: non-finite output"
    r = rel_l2(actual, ref)
    rms = float(ref.double().pow(2).mean().sqrt())
    worst = float((actual - ref).abs().max())
    bound = (max_abs if max_abs is not None else 8 * rel) * max(rms, 1e-6)
    if os.environ.get("CVD_TEST_VERBOSE"):      # margins for the record (pytest -s)
        print(f"[assert_close] {what}: rel_l2={r:.3e} / {rel:.1e}, max_abs={worst:.3e} / {bound:.3e}")
    assert r <= rel and worst <= bound, f"{what}: rel_l2={r:.3e} (tol {rel:.1e}), max_abs={worst:.3e} (tol {bound:.3e}), rms={rms:.3e}"
    return r, worst

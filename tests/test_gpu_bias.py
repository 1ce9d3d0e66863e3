"""-m gpu: epipolar bias prepass + dense writer through the C ABI vs the reference fixtures."""
import numpy as np
import pytest
import torch

from gpu_util import t

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from gpu_util import test_device
    return test_device()


def dense(Fm, feat, S, dev, theta=None):
    from cvd_b200 import _lib, ops
    rule = _lib.FF_HORIZONTAL if theta is None else _lib.FF_RANDOM_SLOPE
    th = None if theta is None else torch.tensor([theta], device=dev, dtype=torch.float32)
    rowp = ops.epi_row_params(Fm.to(dev).contiguous(), feat, S, th, rule)
    return ops.epi_bias_dense(rowp, feat, S).cpu(), rowp.cpu()


@pytest.mark.parametrize("feat", [32, 16, 8, 4])
@pytest.mark.parametrize("tag", ["h", "t"])
def test_bias_vs_reference_fixture(golden, dev, feat, tag):
    g = golden["bias"]
    Fm = t(golden["geometry"]["F_mats"])
    theta = float(g["theta"]) if tag == "t" else None
    b, _ = dense(Fm, feat, 256, dev, theta)
    k = f"f{feat}{tag}_"
    N = feat * feat
    np.testing.assert_allclose(b[:, 0].numpy(), g[k + "q0"], atol=3e-3, rtol=1e-4)
    np.testing.assert_allclose(b[:, N // 2 + 3].numpy(), g[k + "qmid"], atol=3e-3, rtol=1e-4)
    np.testing.assert_allclose(torch.diagonal(b, dim1=1, dim2=2).numpy(), g[k + "diag"], atol=3e-3, rtol=1e-4)
    np.testing.assert_allclose(b.double().sum(dim=(1, 2)).numpy(), g[k + "rowsum"], rtol=2e-5, atol=1.0)
    assert abs(float(b.min()) - float(g[k + "min"])) < 2e-3
    if feat <= 8:
        np.testing.assert_allclose(b.numpy(), g[k + "full"], atol=3e-3, rtol=1e-4)


def test_bias_degenerate_rows_and_other_size(golden, dev):
    g = golden["bias"]
    Fm = t(golden["geometry"]["F_mats"])
    Fbad = Fm.clone()
    Fbad[3] = 0.0
    Fbad[5, 0, 0] = float("nan")
    Fbad[7, 1, 2] = float("inf")
    b, rowp = dense(Fbad, 8, 256, dev)
    assert torch.isfinite(b).all() and torch.isfinite(rowp).all()
    np.testing.assert_allclose(b.numpy(), g["bad_full"], atol=3e-3, rtol=1e-4)
    b512, _ = dense(Fm, 8, 512, dev)
    np.testing.assert_allclose(b512.numpy(), g["s512_f8_full"], atol=3e-3, rtol=1e-4)


def test_bias_vs_oracle_random_F(dev):
    """Random fundamental-like matrices, feat 32, full dense compare with the CPU oracle."""
    from oracle import epi_oracle as O
    rng = np.random.RandomState(7)
    Fm = torch.from_numpy((rng.standard_normal((32, 3, 3)) * np.array([[1e-5, 1e-5, 1e-3], [1e-5, 1e-5, 1e-3], [1e-3, 1e-3, 1e-1]])).astype(np.float32))
    for theta in (None, 1.234):
        b, _ = dense(Fm, 32, 256, dev, theta)
        ref = O.epi_bias(Fm, 32, 256, theta=theta)
        np.testing.assert_allclose(b.numpy(), ref.numpy(), atol=5e-3, rtol=2e-4)


def test_pairwise_fundamental_on_device(golden, dev):
    """cvd_pairwise_fundamental == the reference's CPU loop (fixture F_mats of the shipped pose pair) and
    == the oracle on a random 4-view pairing."""
    from cvd_b200 import ops
    from oracle import epi_oracle as O
    g = golden["geometry"]
    c2w = torch.from_numpy(g["c2w"][0].astype(np.float64)).to(dev).contiguous()
    K = torch.from_numpy(g["K_mats"].astype(np.float64)).to(dev).contiguous()
    partner = torch.cat([torch.arange(16, 32), torch.arange(0, 16)]).to(torch.int32).to(dev)
    Fm = ops.pairwise_fundamental(c2w, K, partner).cpu()
    np.testing.assert_allclose(Fm.numpy(), g["F_mats"], rtol=2e-4, atol=1e-9)
    rng = np.random.RandomState(3)
    from scipy.spatial.transform import Rotation
    c = np.tile(np.eye(4), (64, 1, 1))
    c[:, :3, :3] = Rotation.from_rotvec(rng.randn(64, 3) * 0.2).as_matrix()
    c[:, :3, 3] = rng.randn(64, 3)
    Km = np.tile(np.array([[223.578, 0, 128], [0, 223.578, 128], [0, 0, 1.0]]), (64, 1, 1))
    pairing = [2, 3, 0, 1]
    idx = np.concatenate([np.arange(16) + p * 16 for p in pairing]).astype(np.int32)
    Fd = ops.pairwise_fundamental(torch.from_numpy(c).to(dev), torch.from_numpy(Km).to(dev), torch.from_numpy(idx).to(dev)).cpu()
    np.testing.assert_allclose(Fd.numpy(), O.pairwise_F_mats(c, Km, pairing).numpy(), rtol=1e-5, atol=1e-9)

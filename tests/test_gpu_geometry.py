"""-m gpu: camera geometry on the device: the Pluecker ray embedding of the pose-encoder
input against ray_condition of the reference itself (tests/golden/plucker.npz) and the per-call fundamental matrices
for poses that are NOT rigid (the reference inverts with np.linalg.inv, so scaled / sheared user poses must agree)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dev():
    from gpu_util import test_device
    return test_device()


def test_plucker_embedding_vs_reference(dev):
    from cvd_b200 import ops
    g = np.load(os.path.join(ROOT, "tests", "golden", "plucker.npz"))
    K = torch.from_numpy(g["intrinsics"]).to(dev)
    c2w = torch.from_numpy(g["c2w"]).to(dev)
    full = ops.plucker_embedding(K, c2w, 256, 256).cpu()                       # [31, 6, 256, 256], the dataset's layout
    assert full.shape == (31, 6, 256, 256)
    assert float((full[17] - torch.from_numpy(g["frame17"])).abs().max()) < 5e-6
    assert float((full[:, :, 3::8, 5::8] - torch.from_numpy(g["sub8"])).abs().max()) < 5e-6
    assert np.allclose(full.double().sum(dim=(2, 3)).numpy(), g["chan_sum"], rtol=1e-5, atol=5e-2)
    Ks = torch.from_numpy(g["small_K"]).to(dev)
    small = ops.plucker_embedding(Ks, c2w[:4].contiguous(), 5, 7, channel_major=False).cpu()   # ray_condition's layout
    assert float((small - torch.from_numpy(g["small"])).abs().max()) < 5e-6
    flipped = ops.plucker_embedding(Ks, c2w[:4].contiguous(), 5, 7, flip_flag=torch.from_numpy(g["flip"]),
                                    channel_major=False).cpu()
    assert float((flipped - torch.from_numpy(g["flipped"])).abs().max()) < 5e-6
    # the oracle agrees with the kernel at a size no fixture covers (ragged grid, all cameras)
    from oracle import epi_oracle as O
    ref = O.plucker_embedding(g["intrinsics"], g["c2w"], 37, 19)
    out = ops.plucker_embedding(K, c2w, 37, 19, channel_major=False).cpu()
    assert float((out - ref).abs().max()) < 5e-6


def test_pairwise_fundamental_non_rigid_poses(dev):
    """c2w with scale and shear: F must equal K2^-T E(inv(c2w_t) c2w_s) K1^-1 with a GENERAL inverse, as the reference's
    np.linalg.inv gives (pipeline_animation_epi_advanced.py:639-645), not a transpose-based rigid inverse."""
    from cvd_b200 import ops
    from oracle import epi_oracle as O
    rng = np.random.RandomState(5)
    V, n = 2, 16
    c2w = np.tile(np.eye(4), (V * n, 1, 1))
    c2w[:, :3, :3] = np.eye(3) + 0.3 * rng.standard_normal((V * n, 3, 3))       # not orthonormal
    c2w[:, :3, 3] = rng.standard_normal((V * n, 3))
    K = np.tile(np.array([[220.0, 0, 128], [0, 230.0, 120], [0, 0, 1]]), (V * n, 1, 1))
    partner_view = np.array([1, 0])
    ref = O.pairwise_F_mats(c2w, K, partner_view)
    idx = (partner_view[:, None] * n + np.arange(n)[None]).reshape(-1).astype(np.int32)
    out = ops.pairwise_fundamental(torch.from_numpy(c2w).to(dev), torch.from_numpy(K).to(dev),
                                   torch.from_numpy(idx).to(dev)).cpu()
    assert float((out - ref).abs().max()) <= 1e-5 * float(ref.abs().max())

"""-m gpu: the HBM-bound glue kernels through the C ABI vs plain fp32 torch."""
import pytest
import torch
import torch.nn.functional as F

from gpu_util import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from gpu_util import test_device
    return test_device()


@pytest.mark.parametrize("B,C,Fr,hw,dtype", [(2, 320, 16, 32, torch.float32), (2, 640, 16, 16, torch.float32),
                                             (1, 1280, 16, 8, torch.bfloat16), (2, 1280, 16, 4, torch.float16),
                                             (1, 320, 3, 5, torch.float32)])
@pytest.mark.parametrize("layout", [0, 1])
def test_groupnorm_tokens(dev, B, C, Fr, hw, dtype, layout):
    from cvd_b200 import ops
    g = torch.Generator().manual_seed(0)
    x = (torch.randn(B, C, Fr, hw, hw, generator=g) * 2 + 0.5).to(dev).to(dtype)
    gamma = (1 + 0.1 * torch.randn(C, generator=g)).to(dev)
    beta = (0.1 * torch.randn(C, generator=g)).to(dev)
    y = ops.groupnorm_tokens(x, gamma, beta, 32, 1e-6, layout)
    xf = x.float().permute(0, 2, 1, 3, 4).reshape(B * Fr, C, hw, hw)
    ref = F.group_norm(xf, 32, gamma, beta, 1e-6).reshape(B, Fr, C, hw * hw)
    ref = ref.permute(0, 1, 3, 2).reshape(B * Fr, hw * hw, C) if layout == 0 else ref.permute(0, 3, 1, 2).reshape(B * hw * hw, Fr, C)
    assert y.dtype == torch.bfloat16 and y.shape == ref.shape
    assert_close(y, ref, rel=4e-3, max_abs=2e-2, what="groupnorm_tokens")


@pytest.mark.parametrize("C", [320, 640, 1280, 64])
@pytest.mark.parametrize("with_pe", [False, True])
def test_layernorm(dev, C, with_pe):
    from cvd_b200 import ops
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(37, 16, C, generator=g) * 3 + 1).to(dev)
    gamma = (1 + 0.1 * torch.randn(C, generator=g)).to(dev)
    beta = (0.1 * torch.randn(C, generator=g)).to(dev)
    pe = torch.randn(16, C, generator=g).to(dev) if with_pe else None
    y = ops.layernorm(x, gamma, beta, 1e-5, pe=pe, seq_len=16)
    ref = F.layer_norm(x, (C,), gamma, beta, 1e-5)
    if with_pe:
        ref = ref + pe
    assert_close(y, ref, rel=4e-3, max_abs=2e-2, what="layernorm")


@pytest.mark.parametrize("C", [320, 640, 1280, 64])
@pytest.mark.parametrize("mode", ["full", "bias_only", "no_bias", "pe"])
def test_residual_layernorm_fused(dev, C, mode):
    """x += (y + bias) * scale in place, n = LN(x) (+pe): against the two separate torch ops."""
    from cvd_b200 import ops
    g = torch.Generator().manual_seed(11)
    x = (torch.randn(37, 16, C, generator=g) * 3 + 1).to(dev)
    y = None if mode == "bias_only" else torch.randn(37, 16, C, generator=g).to(dev).to(torch.bfloat16)
    bias = None if mode == "no_bias" else torch.randn(C, generator=g).to(dev)
    gamma = (1 + 0.1 * torch.randn(C, generator=g)).to(dev)
    beta = (0.1 * torch.randn(C, generator=g)).to(dev)
    pe = torch.randn(16, C, generator=g).to(dev) if mode == "pe" else None
    scale = 1.0 if mode == "bias_only" else 0.7
    upd = (0 if y is None else y.float()) + (0 if bias is None else bias)
    x_ref = x + upd * scale
    n_ref = F.layer_norm(x_ref, (C,), gamma, beta, 1e-5) + (0 if pe is None else pe)
    n = ops.residual_layernorm_(x, y, bias, scale, gamma, beta, 1e-5, pe=pe, seq_len=16)
    assert n.dtype == torch.bfloat16
    assert torch.allclose(x, x_ref, rtol=1e-6, atol=1e-5)          # the fp32 stream itself is updated exactly
    assert_close(n, n_ref, rel=4e-3, max_abs=2e-2, what="residual_layernorm")


def test_geglu_and_residual(dev):
    from cvd_b200 import ops
    g = torch.Generator().manual_seed(2)
    hg = torch.randn(50, 16, 2560, generator=g).to(dev).to(torch.bfloat16)
    y = ops.geglu(hg)
    h, gate = hg.float().chunk(2, dim=-1)
    ref = h * F.gelu(gate)
    assert_close(y, ref, rel=4e-3, max_abs=1.0, what="geglu")
    assert torch.allclose(y.float(), ref, rtol=1e-2, atol=2e-3)      # bf16 output: error scales with |value|
    x = torch.randn(50, 16, 320, generator=g).to(dev)
    yb = torch.randn(50, 16, 320, generator=g).to(dev).to(torch.bfloat16)
    bias = torch.randn(320, generator=g).to(dev)
    ref = x + (yb.float() + bias) * 0.5
    xb = ops.residual_add_(x, yb, bias, 0.5, want_bf16=True)
    assert torch.allclose(x, ref, atol=1e-6)
    assert torch.equal(xb, ref.to(torch.bfloat16))


@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_tokens_to_video(dev, layout, dtype):
    from cvd_b200 import ops
    B, C, Fr, hw = 2, 320, 16, 8
    g = torch.Generator().manual_seed(3)
    res = torch.randn(B, C, Fr, hw, hw, generator=g).to(dev).to(dtype)
    shape = (B * Fr, hw * hw, C) if layout == 0 else (B * hw * hw, Fr, C)
    tok = torch.randn(*shape, generator=g).to(dev).to(torch.bfloat16)
    bias = torch.randn(C, generator=g).to(dev)
    out = ops.tokens_to_video(tok, bias, res, layout)
    tf = tok.float() + bias
    vid = tf.reshape(B, Fr, hw, hw, C).permute(0, 4, 1, 2, 3) if layout == 0 else tf.reshape(B, hw, hw, Fr, C).permute(0, 4, 3, 1, 2)
    ref = res.float() + vid
    assert out.dtype == dtype
    assert_close(out, ref, rel=1e-6 if dtype == torch.float32 else 4e-3, max_abs=3e-2, what="tokens_to_video")


@pytest.mark.parametrize("C", [320, 640, 1280])
def test_temporal_attention(dev, C):
    from cvd_b200 import ops
    T, Fr, H = 300, 16, 8
    g = torch.Generator().manual_seed(4)
    q, k, v = (torch.randn(T, Fr, C, generator=g).to(dev).to(torch.bfloat16) for _ in range(3))
    out = ops.temporal_attn(q, k, v, H)
    d = C // H
    qh, kh, vh = (a.float().view(T, Fr, H, d).transpose(1, 2) for a in (q, k, v))
    ref = F.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(T, Fr, C)
    assert_close(out, ref, rel=4e-3, max_abs=3e-2, what="temporal_attn")


def test_pose_add_and_scale_add(dev):
    from cvd_b200 import ops
    B, C, Fr, hw = 2, 320, 16, 8
    g = torch.Generator().manual_seed(5)
    pose = torch.randn(B, C, Fr, hw, hw, generator=g).to(dev)
    x = torch.randn(B * hw * hw, Fr, C, generator=g).to(dev).to(torch.bfloat16)
    y = ops.add_pose_tokens(x, pose)
    ref = x.float() + pose.permute(0, 3, 4, 2, 1).reshape(B * hw * hw, Fr, C)
    assert_close(y, ref, rel=4e-3, max_abs=3e-2, what="add_pose_tokens")
    z = ops.scale_add(x, y, 0.75)
    assert_close(z, x.float() * 0.75 + y.float(), rel=4e-3, max_abs=3e-2, what="scale_add")


def test_cpu_tensor_is_refused():
    from cvd_b200 import CvdError, ops
    with pytest.raises(CvdError):
        ops.geglu(torch.zeros(4, 16, dtype=torch.bfloat16))


def test_peer_gather_rows_single_gpu(dev):
    """cvd_peer_gather_rows with two 'ranks' that are two local buffers: dst[r] = src[g // rows][g % rows]."""
    from cvd_b200 import ops
    g = torch.Generator().manual_seed(6)
    a = torch.randn(32, 16, 40, generator=g).to(dev).to(torch.bfloat16)
    b = torch.randn(32, 16, 40, generator=g).to(dev).to(torch.bfloat16)
    partner = torch.randperm(64, generator=g)[:48].to(torch.int32).to(dev)
    dst = torch.empty(48, 16, 40, device=dev, dtype=torch.bfloat16)
    ops.peer_gather_rows([a.data_ptr(), b.data_ptr()], partner, dst, rows_per_src=32)
    ref = torch.cat([a, b])[partner.long()]
    assert torch.equal(dst, ref)
    big = torch.randn(8, 1024, 320, generator=g).to(dev).to(torch.bfloat16)       # long rows, several chunks per row
    idx = torch.tensor([7, 0, 3, 3, 5], dtype=torch.int32, device=dev)
    out = torch.empty(5, 1024, 320, device=dev, dtype=torch.bfloat16)
    ops.peer_gather_rows([big.data_ptr()], idx, out, rows_per_src=8)
    assert torch.equal(out, big[idx.long()])

"""-m gpu: the sync-module training step (BASELINE config 5: train_epi_control.py:589-616
runs the reference EpiModule under autocast and calls loss.backward()).

Oracle: torch autograd through the fp32 CPU restatement of the reference module (oracle/epi_oracle.py: epi_module)
on the same state dict and inputs.  Tolerance (stated): the product computes every GEMM and the attention core with
bf16 operands / fp32 accumulation and returns bf16 attention gradients, so the forward is held to rel-L2 3e-2 on the
module delta and every gradient (input and each of the 26 parameter tensors) to rel-L2 5e-2.
"""
import os

import numpy as np
import pytest
import torch

import golden_inputs as gi
from gpu_util import rel_l2, t

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from gpu_util import test_device
    return test_device()


def build(C, dev, sd, **kw):
    import cvd_b200
    m = cvd_b200.EpiModule(C, num_attention_heads=8, num_transformer_block=1, attention_block_types=["Epi_Self", "Epi_Self"],
                           epi_position_encoding=True, epi_position_encoding_feat_max_size=64,
                           epi_position_encoding_F_mat_size=256, epi_rand_slope_ff=False, zero_initialize=True, **kw)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all("pos_encoder" in k for k in missing)
    return m.to(dev)


@pytest.mark.parametrize("C,feat,B", [(320, 8, 2), (640, 8, 2), (320, 16, 2), (320, 8, 4)])
def test_epi_module_train_step_vs_oracle_autograd(golden, dev, C, feat, B):
    from oracle import epi_oracle as O
    seed = 4000 + C + feat + B
    sd = gi.module_state_dict(C, seed)
    x = gi.hidden_video(B, C, 16, feat, feat, seed + 1)
    wgt = gi.hidden_video(B, C, 16, feat, feat, seed + 2)          # loss = <y, wgt>
    Fm = gi.tile_F(t(golden["geometry"]["F_mats"]), B * 16).reshape(B, 16, 3, 3)

    # oracle: fp32 CPU autograd
    sd_ref = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    x_ref = x.clone().requires_grad_(True)
    y_ref, _ = O.epi_module(x_ref, sd_ref, 8, Fm, 256)
    (y_ref * wgt).sum().backward()

    # product: train mode, CUDA attention forward / backward inside torch autograd
    m = build(C, dev, sd).train()
    xp = x.to(dev).requires_grad_(True)
    y, aux = m(xp, F_mats=Fm.to(dev), temb=None, encoder_hidden_states=None, cross_attention_kwargs={"pose_feature": None})
    (y * wgt.to(dev)).sum().backward()
    assert y.shape == x.shape and y.dtype == x.dtype and len(aux) == 2
    assert aux[0]["query"].requires_grad and aux[0]["key"].shape == (B * 16, feat * feat, C)

    e_fwd = rel_l2((y.detach().cpu() - x), (y_ref.detach() - x))
    e_x = rel_l2(xp.grad.cpu(), x_ref.grad)
    worst = ("", 0.0)
    params = dict(m.named_parameters())
    checked = 0
    for name, ref in sd_ref.items():
        g = params[name].grad
        assert g is not None, f"no gradient for {name}"
        e = rel_l2(g.cpu(), ref.grad)
        checked += 1
        if e > worst[1]:
            worst = (name, e)
    print(f"C={C} feat={feat} B={B}: module delta {e_fwd:.2e}, d/dx {e_x:.2e}, worst of {checked} parameter gradients "
          f"{worst[1]:.2e} ({worst[0]})")
    assert checked == len(sd)
    assert e_fwd < 3e-2 and e_x < 5e-2 and worst[1] < 5e-2, (e_fwd, e_x, worst)


def test_train_path_equals_inference_path(golden, dev):
    """The differentiable path and the fused inference path are the same function (both bf16 GEMMs around the same
    attention kernel): their outputs agree to bf16 rounding."""
    C, feat, B = 320, 16, 2
    sd = gi.module_state_dict(C, 4242)
    x = gi.hidden_video(B, C, 16, feat, feat, 4243).to(dev)
    Fm = t(golden["geometry"]["F_mats"]).reshape(2, 16, 3, 3).to(dev)
    m = build(C, dev, sd)
    y_train, _ = m.train()(x, F_mats=Fm)
    with torch.no_grad():
        y_inf, _ = m.eval()(x, F_mats=Fm)
    assert rel_l2(y_train.detach() - x, y_inf - x) < 1e-2


def test_optimizer_step_reduces_loss(golden, dev):
    """Three AdamW steps on the epi parameters (the trainable set of train_epi_control.py:270-283) lower an MSE loss."""
    C, feat, B = 320, 8, 2
    sd = gi.module_state_dict(C, 4300)
    x = gi.hidden_video(B, C, 16, feat, feat, 4301).to(dev)
    target = gi.hidden_video(B, C, 16, feat, feat, 4302).to(dev)
    Fm = t(golden["geometry"]["F_mats"]).reshape(2, 16, 3, 3).to(dev)
    m = build(C, dev, sd).train()
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3)
    losses = []
    for _ in range(4):
        opt.zero_grad(set_to_none=True)
        y, _ = m(x, F_mats=Fm)
        loss = torch.nn.functional.mse_loss(y.float(), target.float())
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)     # train_epi_control.py:621-631
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0], losses


def test_training_requires_cuda():
    import cvd_b200
    m = cvd_b200.EpiModule(320, num_attention_heads=8, num_transformer_block=1, attention_block_types=["Epi_Self"]).train()
    with pytest.raises(cvd_b200._lib.CvdError):
        m(torch.zeros(2, 320, 16, 4, 4), F_mats=torch.zeros(2, 16, 3, 3))


def test_epi_module_train_step_config5_shape(golden, dev):
    """BASELINE config 5 at full size: one EpiModule forward + backward at (C = 320, 32x32 latent, B = 16 videos = 8
    two-view pairs -> R = 256 rows), i.e. through the R > 200 memory-relief branch on both sides (epi_module.py:424-451,
    cvd_b200/epi_module.py attend_train).  Oracle: the fp32 restatement of the reference module with torch autograd, run
    on the GPU for this size (the CPU run of it takes minutes); same tolerances as the small cases."""
    from oracle import epi_oracle as O
    C, feat, B = 320, 32, 16
    sd = gi.module_state_dict(C, 4555)
    x = gi.hidden_video(B, C, 16, feat, feat, 4556)
    wgt = gi.hidden_video(B, C, 16, feat, feat, 4557)
    Fm = gi.tile_F(t(golden["geometry"]["F_mats"]), B * 16).reshape(B, 16, 3, 3)

    with torch.device(dev):
        sd_ref = {k: v.to(dev).requires_grad_(True) for k, v in sd.items()}
        x_ref = x.to(dev).requires_grad_(True)
        y_ref, aux_ref = O.epi_module(x_ref, sd_ref, 8, Fm.to(dev), 256)
        (y_ref * wgt.to(dev)).sum().backward()
    y_ref, gx_ref = y_ref.detach().cpu(), x_ref.grad.cpu()
    g_ref = {k: v.grad.cpu() for k, v in sd_ref.items()}
    del sd_ref, x_ref
    torch.cuda.empty_cache()

    m = build(C, dev, sd).train()
    xp = x.to(dev).requires_grad_(True)
    y, aux = m(xp, F_mats=Fm.to(dev))
    (y * wgt.to(dev)).sum().backward()
    assert aux == [None, None]                      # the memory-relief branch returns no aux (epi_module.py:450)
    e_fwd = rel_l2((y.detach().cpu() - x), (y_ref - x))
    e_x = rel_l2(xp.grad.cpu(), gx_ref)
    params = dict(m.named_parameters())
    worst = max(((rel_l2(params[n].grad.cpu(), g), n) for n, g in g_ref.items()))
    print(f"config 5 (C=320, 32x32, R=256): module delta {e_fwd:.2e}, d/dx {e_x:.2e}, worst of {len(g_ref)} parameter "
          f"gradients {worst[0]:.2e} ({worst[1]})")
    assert e_fwd < 3e-2 and e_x < 5e-2 and worst[0] < 5e-2, (e_fwd, e_x, worst)


def test_backward_through_temporal_and_epi_sites(dev):
    """The training recipe of INTEGRATION.md: unet.train(); loss.backward() with the reference's frozen motion modules
    replaced too.  Gradients of every epi module except the last flow THROUGH the temporal modules
    (train_epi_control.py:589-630), so VanillaTemporalModule records an autograd graph (forward_train) instead of
    refusing.  Two levels = 4 + 4 sites; the temporal twin equals the CUDA inference path to bf16 rounding, and every
    epi parameter receives a finite, non-zero gradient."""
    from cvd_b200 import sync_unet
    torch.manual_seed(11)
    ek = dict(sync_unet.EPI_KWARGS, epi_rand_slope_ff=False)
    net = sync_unet.SyncPathUNet(levels=2, epi_kwargs=ek).randomize_zero_init(11).to(dev)
    g = torch.Generator(device=dev).manual_seed(5)
    x = torch.randn(2, 4, 16, 16, 16, device=dev, generator=g)
    text = torch.randn(2, 77, 768, device=dev, generator=g)
    poses = [torch.randn(2, c, 16, 16 >> i, 16 >> i, device=dev, generator=g) * 0.5 for i, c in enumerate((320, 640))]
    Fm = torch.from_numpy(np.load(os.path.join(os.path.dirname(__file__), "golden", "geometry.npz"))["F_mats"])
    Fm = Fm.reshape(2, 16, 3, 3).to(dev)
    for n_, p in net.named_parameters():              # the reference trains the epi modules only
        p.requires_grad_("epi" in n_)
    net.train()
    y = net(x, 481, F_mats=Fm, encoder_hidden_states=text, pose_embedding_features=poses)
    y = y.sample if hasattr(y, "sample") else y
    y.float().pow(2).mean().backward()
    grads = [(n_, p.grad) for n_, p in net.named_parameters() if p.requires_grad]
    assert grads and all(g_ is not None and torch.isfinite(g_).all() for _, g_ in grads)
    first_epi = [g_ for n_, g_ in grads if "to_q" in n_][0]
    assert float(first_epi.abs().max()) > 0.0        # reached through the temporal modules behind it
    with torch.no_grad():
        y_inf = net.eval()(x, 481, F_mats=Fm, encoder_hidden_states=text, pose_embedding_features=poses)
    y_inf = y_inf.sample if hasattr(y_inf, "sample") else y_inf
    assert rel_l2(y.detach(), y_inf) < 2e-2

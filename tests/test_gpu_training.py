"""-m gpu: the sync-module training step (BASELINE config 5: train_epi_control.py:589-616
runs the reference EpiModule under autocast and calls loss.backward()).

Oracle: torch autograd through the fp32 CPU restatement of the reference module (oracle/epi_oracle.py: epi_module)
on the same state dict and inputs.  Tolerance (stated): the product computes every GEMM and the attention core with
bf16 operands / fp32 accumulation and returns bf16 attention gradients, so the forward is held to rel-L2 3e-2 on the
module delta and every gradient (input and each of the 26 parameter tensors) to rel-L2 5e-2.
"""
import pytest
import torch

import golden_inputs as gi
from gpu_util import rel_l2, t

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


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

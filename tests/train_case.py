"""Shared body of the UNet-level training-step parity tests (BASELINE config 5): the product's
``cvd_b200.training`` step against ONE optimisation step of the reference's loop on the reference's own modules
(tests/golden/train_step*.npz, made by tests/golden/make_golden.py: train_step()).

Callers: tests/test_emulated_layers.py (CPU, emulated kernels: the reduced-width fixture, and the device-width fixture under
the strict device-shape rules) and tests/test_gpu_zz_round2_late.py (the B200, device-width fixture)."""
import os

import numpy as np
import torch

import golden_inputs as gi
from gpu_util import rel_l2

HERE = os.path.dirname(os.path.abspath(__file__))


def run_training_step_case(dev, F_mats, device_widths, rel):
    import cvd_b200
    from cvd_b200 import sampler, training
    from cvd_b200.unet import UNet3DConditionModelPoseCond
    g = np.load(os.path.join(HERE, "golden", "train_step_gpu.npz" if device_widths else "train_step.npz"))
    ukw, pkw, useed, ekw = (gi.UNET_GPU_KWARGS, gi.UNET_GPU_PROCESSOR_KWARGS, gi.UNET_GPU_SEED, gi.TRAIN_GPU_POSE_ENCODER_KWARGS) \
        if device_widths else (gi.UNET_KWARGS, gi.UNET_PROCESSOR_KWARGS, gi.UNET_SEED, gi.TRAIN_POSE_ENCODER_KWARGS)
    unet = UNet3DConditionModelPoseCond(**ukw, amp=False)
    unet.set_all_attn_processor(add_spatial_lora=False, add_motion_lora=False, lora_kwargs={"lora_rank": 2, "lora_scale": 1.0},
                                motion_lora_kwargs={"lora_rank": -1, "lora_scale": 1.0},
                                sync_lora_kwargs={"sync_lora_rank": 0, "sync_lora_scale": 0}, spatial_extended_attention=False,
                                **pkw)
    skip = ("pos_encoder.pe", "pos_encoder.coords")
    shapes = {k: tuple(v.shape) for k, v in unet.state_dict().items() if not k.endswith(skip)}
    unet.load_state_dict(gi.fill_state_dict(shapes, useed), strict=False)
    enc = cvd_b200.CameraPoseEncoder(**ekw)
    eshapes = {k: tuple(v.shape) for k, v in enc.state_dict().items() if not k.endswith("pos_encoder.pe")}
    enc.load_state_dict(gi.fill_state_dict(eshapes, gi.TRAIN_SEED + 1), strict=False)
    unet, enc = unet.to(dev), enc.to(dev)
    names = training.freeze_for_sync_training(unet, enc)
    assert names == g["names"].tolist()                                           # the checkpoint's unet_trainable_dict keys
    assert not any(p.requires_grad for p in enc.parameters())
    adaptor = training.PoseAdaptor(unet, enc).train()
    params = training.trainable_epi_parameters(adaptor)
    assert len(params) == len(names)
    opt = torch.optim.AdamW(params, **gi.TRAIN_OPT)
    lat, text, _ = gi.unet_inputs(gi.UNET_GPU_SEED, channels=(320, 640), text_dim=64) if device_widths else gi.unet_inputs()
    noise, timesteps, plucker = gi.train_inputs()
    lat, text, noise, timesteps, plucker, F2 = (x.to(dev) for x in (lat, text, noise, timesteps, plucker, F_mats))
    sch = sampler.DDIMSchedule(device=dev)
    # forward diffusion: closed form, one timestep per video
    a = sch.alphas_cumprod[timesteps].reshape(2, 1, 1, 1, 1)
    assert torch.equal(sch.add_noise(lat, noise, timesteps), a ** 0.5 * lat + (1.0 - a) ** 0.5 * noise)
    ref = {k: torch.from_numpy(np.asarray(g[k])) for k in ("model_pred", "grad_norms", "delta_norms")}
    with torch.no_grad():                                                          # epsilon prediction (no update yet)
        pred, aux = adaptor.eval()(sch.add_noise(lat, noise, timesteps), timesteps, text, plucker, F_mats=F2)
    assert aux is None and rel_l2(pred, ref["model_pred"]) < rel, rel_l2(pred, ref["model_pred"])
    adaptor.train()
    before = [p.detach().clone() for p in params]
    grads = {}
    hooks = [p.register_hook(lambda gr, i=i: grads.__setitem__(i, gr.detach().float().norm().cpu())) for i, p in enumerate(params)]
    loss, total = training.training_step(adaptor, opt, sch, lat, noise, timesteps, text, plucker_embedding=plucker, F_mats=F2,
                                         max_grad_norm=gi.TRAIN_MAX_GRAD_NORM)
    for h in hooks:
        h.remove()
    assert abs(float(loss) - float(g["loss"])) < 2e-3 * float(g["loss"]), (float(loss), float(g["loss"]))
    assert abs(float(total) - float(g["total_norm"])) < 3e-2 * float(g["total_norm"]), (float(total), float(g["total_norm"]))
    assert float(total) > gi.TRAIN_MAX_GRAD_NORM                                   # clipping is active in this fixture
    gn = torch.stack([grads[i] for i in range(len(params))])
    assert rel_l2(gn, ref["grad_norms"]) < 3e-2, rel_l2(gn, ref["grad_norms"])
    assert all(p.grad is None for p in params)                                     # zero_grad(set_to_none=True), :640
    deltas = torch.stack([(p.detach() - b).float().norm().cpu() for p, b in zip(params, before)])
    assert rel_l2(deltas, ref["delta_norms"]) < 3e-2, rel_l2(deltas, ref["delta_norms"])
    for j, i in enumerate(g["picks"].tolist()):
        # Adam's first step moves every element by ~lr * sign(grad): an element whose gradient is ~0 may go either way, so
        # compare the direction of the MOVE element by element (and bound the distance by two steps)
        move = (params[i].detach() - before[i]).cpu()
        move_ref = torch.from_numpy(g[f"after{j}"]) - before[i].cpu()
        agree = float((torch.sign(move) == torch.sign(move_ref)).float().mean())
        assert agree >= 0.9 and float((move - move_ref).abs().max()) <= 2.1 * gi.TRAIN_OPT["lr"], (names[i], agree)
    frozen = [p for n, p in unet.named_parameters() if n not in set(names)]
    assert frozen and all(p.grad is None for p in frozen)
    return float(loss), float(total)

#!/usr/bin/env python
"""Independent fp64 restatement of the published DDIM update (Song, Meng, Ermon: "Denoising Diffusion Implicit
Models", ICLR 2021, eq. 12 with sigma_t = 0) over the scheduler configuration the reference loads
(configs/inference_config.yaml noise_scheduler_kwargs: 1000 train steps, linear betas 0.00085 .. 0.012, steps_offset 1,
clip_sample false; DDIMScheduler of diffusers 0.24.0: "leading" timestep spacing, set_alpha_to_one true) -> ddim.npz.

    python tests/golden/make_ddim_fixture.py

``diffusers`` is neither vendored in /root/reference nor installable here, so this fixture does NOT pin
``cvd_b200.sampler.DDIMSchedule`` against diffusers itself ("parity unpinned vs diffusers"); it
pins it against a second derivation that shares no code with it: the update is written here in the paper's own terms
    x_{t-1} = sqrt(a_{t-1}) * (x_t - sqrt(1 - a_t) eps) / sqrt(a_t) + sqrt(1 - a_{t-1}) eps
re-arranged into its two linear coefficients on (x_t, eps), in float64 numpy, for all 25 timesteps of the reference's
default run (including the last step, whose previous timestep is negative and takes alpha = 1), plus a 25-step
trajectory on a seeded sample with a seeded stand-in for the model output, and the re-noise step of the advanced
pipeline (pipeline_animation_epi_advanced.py:700-705)."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    T, n = 1000, 25
    betas = np.linspace(0.00085, 0.012, T, dtype=np.float64)          # "linear" schedule
    abar = np.cumprod(1.0 - betas)
    stride = T // n
    timesteps = (np.arange(n) * stride)[::-1] + 1                      # leading spacing + steps_offset
    cx, ce, a_t_all, a_prev_all = [], [], [], []
    for t in timesteps:
        a_t = abar[t]
        tp = t - stride
        a_p = abar[tp] if tp >= 0 else 1.0                             # set_alpha_to_one
        cx.append(np.sqrt(a_p / a_t))
        ce.append(np.sqrt(1.0 - a_p) - np.sqrt(a_p * (1.0 - a_t) / a_t))
        a_t_all.append(a_t)
        a_prev_all.append(a_p)
    rng = np.random.RandomState(20240)
    x = rng.standard_normal((2, 4, 3, 5, 5))
    traj, eps_all = [x.copy()], []
    for i, t in enumerate(timesteps):
        eps = 0.7 * np.tanh(x) + 0.3 * rng.standard_normal(x.shape)    # stand-in for the UNet's noise prediction
        eps_all.append(eps)
        x = cx[i] * x + ce[i] * eps
        traj.append(x.copy())
    # re-noise between multistep repeats: x_t = sqrt(a_t / a_prev) x_prev + sqrt(1 - a_t / a_prev) noise; never applied
    # at the last timestep (MULTI_STEP = 1 there, pipeline_animation_epi_advanced.py:602), so prev_t >= 1 always
    noise = rng.standard_normal(x.shape)
    ren_t = [int(timesteps[0]), int(timesteps[12]), int(timesteps[-2])]
    ren = []
    for t in ren_t:
        r = abar[t] / abar[t - stride]
        ren.append(np.sqrt(r) * traj[3] + np.sqrt(1.0 - r) * noise)
    np.savez_compressed(os.path.join(HERE, "ddim.npz"), timesteps=timesteps.astype(np.int64), coef_x=np.array(cx),
                        coef_eps=np.array(ce), alpha_t=np.array(a_t_all), alpha_prev=np.array(a_prev_all),
                        x0=traj[0], eps=np.array(eps_all), traj=np.array(traj[1:]), renoise_t=np.array(ren_t),
                        renoise_noise=noise, renoise_in=traj[3], renoise_out=np.array(ren))
    print("wrote ddim.npz; last step alpha_prev =", a_prev_all[-1], "timesteps", timesteps[:3], "...", timesteps[-2:])


if __name__ == "__main__":
    main()

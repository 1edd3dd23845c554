"""Host-side noise schedule of the MMDM (numpy float64 -> float32 buffers).

Follows MMLDM.register_schedule (reference cap4d/mmdm/mmdm.py:276-324): linear-in-sqrt betas
(util.py:21-25) -> zero-terminal-SNR rescale (cap4d/mmdm/utils.py:18-37) -> beta clip at 0.99 ->
cumulative product -> log-SNR shift by sqrt(64^2 / (image_size^2 * (n_frames-1)))
(mmdm.py:293-308, utils.py:4-14), and the DDIM sub-schedule of StochasticIOSampler.make_schedule
(cap4d/mmdm/sampler.py:32-61 with util.py:46-74).
"""
from __future__ import annotations

import numpy as np
import torch


class MMDMSchedule:
    """The three buffers the sampler reads from the model (sampler.py:26,35-41)."""

    def __init__(self, timesteps=1000, linear_start=0.00085, linear_end=0.0120, n_frames=8, image_size=64,
                 zero_snr_shift=True, shift_schedule=True, sqrt_shift=True, minus_one_shift=True):
        # torch.linspace in float64 is what the reference uses; numpy's linspace rounds differently
        root = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, timesteps, dtype=torch.float64).numpy()
        betas = root * root
        if zero_snr_shift:
            bar_sqrt = np.sqrt(np.cumprod(1.0 - betas))
            first, last = bar_sqrt[0].copy(), bar_sqrt[-1].copy()
            bar_sqrt = bar_sqrt - last
            bar_sqrt = bar_sqrt * (first / (first - last))
            bar = bar_sqrt ** 2
            ratio = bar[1:] / bar[:-1]
            betas = 1.0 - np.concatenate([bar[:1], ratio])
        betas[betas > 0.99] = 0.99
        acp = np.cumprod(1.0 - betas, axis=0)
        if shift_schedule:
            n_gen = n_frames - 1 if minus_one_shift else n_frames
            shift = (64 ** 2) / (image_size ** 2 * n_gen)
            if sqrt_shift:
                shift = np.sqrt(shift)
            log_snr = np.log(acp / (1.0 - acp)) + np.log(shift)
            shifted = np.exp(log_snr) / (1 + np.exp(log_snr))
            betas = 1 - np.concatenate([[1], shifted[1:] / shifted[:-1]])
            acp = shifted
        acp_prev = np.append(1.0, acp[:-1])
        self.num_timesteps = int(betas.shape[0])
        self.betas = torch.tensor(betas, dtype=torch.float32)
        self.alphas_cumprod = torch.tensor(acp, dtype=torch.float32)
        self.alphas_cumprod_prev = torch.tensor(acp_prev, dtype=torch.float32)


def ddim_timesteps(S: int, num_ddpm: int) -> np.ndarray:
    """'uniform' discretisation, +1 (util.py:46-60)."""
    stride = num_ddpm // S
    steps = np.asarray(list(range(0, num_ddpm, stride))) + 1
    if steps.max() >= num_ddpm:
        # the reference raises IndexError here (alphacums[ddim_timesteps], util.py:65)
        raise IndexError(f"S={S} does not give valid DDIM timesteps for {num_ddpm} DDPM steps")
    return steps


def ddim_factors(alphas_cumprod: torch.Tensor, S: int, eta: float = 0.0):
    """Per-step (timestep, x_factor, e_factor) in SAMPLING order (largest timestep first).

    x_{t-1} = x_t * x_factor + eps * e_factor with the float64 -> float32 arithmetic of
    sampler.py:215-229.  Returns (timesteps[int64], x_factor[float32], e_factor[float32])."""
    acp = alphas_cumprod.detach().cpu().to(torch.float32)
    steps = ddim_timesteps(S, acp.shape[0])
    a_t = acp[steps]                                                   # float32 tensor
    a_prev = np.asarray([acp[0]] + acp[steps[:-1]].tolist())           # float64 ndarray of float32 values
    sigma = eta * np.sqrt((1 - a_prev) / (1 - a_t) * (1 - a_t / a_prev))
    sqrt_1m = np.sqrt(1.0 - a_t)                                        # float32 tensor (np.sqrt on a tensor)
    xs, es = [], []
    n = steps.shape[0]
    for i in range(n):
        index = n - i - 1
        at = a_t.float()[index].double()
        s1 = sqrt_1m[index].double()
        sg = sigma[index]
        ap = torch.tensor(a_prev).float()[index].double()
        e_f = -ap.sqrt() * s1 / at.sqrt() + (1.0 - ap - sg ** 2).sqrt()
        x_f = ap.sqrt() / at.sqrt()
        xs.append(float(x_f.float()))
        es.append(float(e_f.float()))
    return np.flip(steps).astype(np.int64).copy(), np.asarray(xs, np.float32), np.asarray(es, np.float32)

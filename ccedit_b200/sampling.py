"""Callers of the network on the hot path - these stay PyTorch elementwise math on
[B, 4, T, h, w] latents, exactly as in the reference; the network call inside is the CUDA path.

    LegacyDDPMDiscretization   sgm/modules/diffusionmodules/discretizer.py:11-21, 42-69
    EpsScaling                 denoiser_scaling.py:16-22
    DiscreteDenoiser           denoiser.py:22-40, 43-75
    VanillaCFGTV2V             guiders.py:8-40, 56-67
    DPMPP2SAncestralSampler    sampling.py:24-77, 168-205, 370-407 ; sampling_utils.py:27-48
Constructor keywords follow the reference so the YAML / init_sampling() arguments carry over; `*_config` dicts whose
`target` names the reference's class resolve to the classes of this file.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import numpy as np
import torch
import torch.nn as nn


def append_dims(x: torch.Tensor, target_dims: int) -> torch.Tensor:
    """sgm/util.py:192-199."""
    if target_dims < x.ndim:
        raise ValueError(f"input has {x.ndim} dims but target_dims is {target_dims}, which is less")
    return x[(...,) + (None,) * (target_dims - x.ndim)]


class LegacyDDPMDiscretization:
    """SD-1.5 sigma table: betas = linspace(sqrt(0.00085), sqrt(0.012), 1000)^2 (float64), sigma = sqrt((1-a)/a)."""

    def __init__(self, linear_start=0.00085, linear_end=0.0120, num_timesteps=1000):
        self.num_timesteps = num_timesteps
        betas = (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, num_timesteps, dtype=torch.float64) ** 2).numpy()
        self.alphas_cumprod = np.cumprod(1.0 - betas, axis=0)

    def get_sigmas(self, n, device="cpu"):
        ac = self.alphas_cumprod
        if n < self.num_timesteps:
            ac = ac[np.linspace(self.num_timesteps - 1, 0, n, endpoint=False).astype(int)[::-1]]
        elif n != self.num_timesteps:
            raise ValueError
        sig = torch.tensor((1 - ac) / ac, dtype=torch.float32, device=device) ** 0.5
        return torch.flip(sig, (0,))

    def __call__(self, n, do_append_zero=True, device="cpu", flip=False):
        sig = self.get_sigmas(n, device=device)
        if do_append_zero:
            sig = torch.cat([sig, sig.new_zeros([1])])
        return sig if not flip else torch.flip(sig, (0,))


class EpsScaling:
    def __call__(self, sigma):
        return torch.ones_like(sigma), -sigma, 1 / (sigma ** 2 + 1.0) ** 0.5, sigma.clone()


class EpsWeighting:
    def __call__(self, sigma):
        return sigma ** -2.0


class VanillaCFGTV2V:
    """Parallel classifier-free guidance: uncond first, cond second; x_u + s (x_c - x_u)."""

    CAT_KEYS = ("vector", "crossattn", "concat", "cond_feat", "control_hint", "interpolate_first", "interpolate_last",
                "interpolate_first_last")

    def __init__(self, scale, dyn_thresh_config=None):
        if dyn_thresh_config is not None:
            raise NotImplementedError("ccedit_b200: dynamic thresholding is outside the hot path")
        self.scale = scale

    def __call__(self, x, sigma):
        x_u, x_c = x.chunk(2)
        return x_u + self.scale * (x_c - x_u)

    def prepare_inputs(self, x, s, c, uc):
        c_out = {}
        for k in c:
            if k in self.CAT_KEYS:
                c_out[k] = torch.cat((uc[k], c[k]), 0)
            else:
                assert c[k] == uc[k]
                c_out[k] = c[k]
        return torch.cat([x] * 2), torch.cat([s] * 2), c_out


_LOCAL = {
    "LegacyDDPMDiscretization": LegacyDDPMDiscretization, "EpsScaling": EpsScaling, "EpsWeighting": EpsWeighting,
    "VanillaCFGTV2V": VanillaCFGTV2V, "VanillaCFG": VanillaCFGTV2V,
}


def _make(config, default_cls=None):
    """instantiate_from_config for the handful of caller-side classes (matched on the class name of `target`)."""
    if config is None:
        return default_cls()
    name = config["target"].rsplit(".", 1)[-1]
    if name not in _LOCAL:
        raise NotImplementedError(f"ccedit_b200.sampling: {config['target']} is outside the hot path")
    return _LOCAL[name](**dict(config.get("params", {})))


class DiscreteDenoiser(nn.Module):
    """sigma -> nearest entry of the 1000-sigma table -> (c_skip, c_out, c_in, idx); net(x c_in, idx, c) c_out + x."""

    def __init__(self, weighting_config=None, scaling_config=None, num_idx=1000, discretization_config=None,
                 do_append_zero=False, quantize_c_noise=True, flip=True):
        super().__init__()
        self.weighting = _make(weighting_config, EpsWeighting)
        self.scaling = _make(scaling_config, EpsScaling)
        sigmas = _make(discretization_config, LegacyDDPMDiscretization)(num_idx, do_append_zero=do_append_zero, flip=flip)
        self.register_buffer("sigmas", sigmas)
        self.quantize_c_noise = quantize_c_noise

    def sigma_to_idx(self, sigma):
        return (sigma - self.sigmas[:, None]).abs().argmin(dim=0).view(sigma.shape)

    def idx_to_sigma(self, idx):
        return self.sigmas[idx]

    def w(self, sigma):
        return self.weighting(sigma)

    def forward(self, network: Callable, input: torch.Tensor, sigma: torch.Tensor, cond: Dict) -> torch.Tensor:
        sigma = self.idx_to_sigma(self.sigma_to_idx(sigma))
        shape = sigma.shape
        sigma = append_dims(sigma, input.ndim)
        c_skip, c_out, c_in, c_noise = self.scaling(sigma)
        c_noise = c_noise.reshape(shape)
        if self.quantize_c_noise:
            c_noise = self.sigma_to_idx(c_noise)
        return network(input * c_in, c_noise, cond) * c_out + input * c_skip


class DPMPP2SAncestralSampler:
    """DPM-Solver++(2S) ancestral: two network evaluations per step (one on the last), eta = 1 noise re-injection."""

    def __init__(self, discretization_config=None, num_steps: Optional[int] = None, guider_config=None,
                 verbose: bool = False, device: str = "cuda", eta: float = 1.0, s_noise: float = 1.0):
        self.num_steps = num_steps
        self.discretization = _make(discretization_config, LegacyDDPMDiscretization)
        self.guider = _make(guider_config)
        self.verbose = verbose
        self.device = device
        self.eta, self.s_noise = eta, s_noise
        self.noise_sampler = lambda x: torch.randn_like(x)

    def denoise(self, x, denoiser, sigma, cond, uc):
        return self.guider(denoiser(*self.guider.prepare_inputs(x, sigma, cond, uc)), sigma)

    def sampler_step(self, sigma, next_sigma, denoiser, x, cond, uc=None):
        nd = x.ndim
        if self.eta:
            sigma_up = torch.minimum(next_sigma, self.eta * (next_sigma ** 2 * (sigma ** 2 - next_sigma ** 2)
                                                             / sigma ** 2) ** 0.5)
            sigma_down = (next_sigma ** 2 - sigma_up ** 2) ** 0.5
        else:
            sigma_up, sigma_down = torch.zeros_like(next_sigma), next_sigma
        denoised = self.denoise(x, denoiser, sigma, cond, uc)
        x_euler = x + (x - denoised) / append_dims(sigma, nd) * append_dims(sigma_down - sigma, nd)
        if torch.sum(sigma_down) < 1e-14:          # the reference's only host sync per step (sampling.py:390)
            x = x_euler
        else:
            t, t_next = -sigma.log(), -sigma_down.log()
            h = t_next - t
            s = t + 0.5 * h
            m1, m2 = append_dims((-s).exp() / (-t).exp(), nd), append_dims((-0.5 * h).expm1(), nd)
            m3, m4 = append_dims((-t_next).exp() / (-t).exp(), nd), append_dims((-h).expm1(), nd)
            x2 = m1 * x - m2 * denoised
            denoised2 = self.denoise(x2, denoiser, (-s).exp(), cond, uc)
            x = torch.where(append_dims(sigma_down, nd) > 0.0, m3 * x - m4 * denoised2, x_euler)
        return torch.where(append_dims(next_sigma, nd) > 0.0,
                           x + self.noise_sampler(x) * self.s_noise * append_dims(sigma_up, nd), x)

    def __call__(self, denoiser, x, cond, uc=None, num_steps=None):
        sigmas = self.discretization(self.num_steps if num_steps is None else num_steps, device=self.device)
        uc = cond if uc is None else uc
        x *= torch.sqrt(1.0 + sigmas[0] ** 2.0)           # in place, as the reference (sampling.py:50)
        s_in = x.new_ones([x.shape[0]])
        for i in range(len(sigmas) - 1):
            x = self.sampler_step(s_in * sigmas[i], s_in * sigmas[i + 1], denoiser, x, cond, uc)
        return x

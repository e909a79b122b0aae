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


# ---------------------------------------------------------------------------------------------------------------------
# Fused sampler step
# ---------------------------------------------------------------------------------------------------------------------
class BoundDenoiser:
    """`denoiser` argument of the sampler with the network and the denoiser visible to it: calling it is exactly the
    reference's `lambda input, sigma, c: model.denoiser(model.model, input, sigma, c)` (sampling_tv2v.py:366-369), so any
    sampler accepts it; FusedDPMPP2SAncestralSampler recognises it and runs its fused, graph-captured step instead."""

    def __init__(self, denoiser: DiscreteDenoiser, network: Callable):
        self.denoiser, self.network = denoiser, network

    def __call__(self, input, sigma, c):
        return self.denoiser(self.network, input, sigma, c)


class FusedDPMPP2SAncestralSampler(DPMPP2SAncestralSampler):
    """DPMPP2SAncestralSampler with the per-step elementwise math (denoiser scaling, CFG combination, the DPM++2S update,
    ancestral noise: ~15 PyTorch launches and one host sync per step in the reference, sampling.py:385-407) in three CUDA
    kernels (csrc/sampler.cu), the CFG-concatenated conditioning built once per clip instead of once per network call
    (guiders.py:56-67 concatenates the 160 MB hint video every call), every per-step scalar precomputed into a device
    table with the reference's own PyTorch expressions, and a whole step - both network calls included - captured in ONE
    CUDA graph that is replayed for every step of the schedule.  Results are bit-identical to the unfused sampler.
    Same constructor and call signature; the fused path is taken when `denoiser` is a BoundDenoiser and x is fp32 CUDA."""

    def __init__(self, *args, use_cuda_graph: bool = True, cfg_dedup: bool = True, **kwargs):
        super().__init__(*args, **kwargs)
        self.use_cuda_graph = use_cuda_graph
        # CFG de-duplication: cat([x] * 2) makes the two halves of the batch identical up to the first text
        # cross-attention of each network; when the hint / reference features of cond and uncond are equal too (checked
        # once per clip) those layers run once (wrapper.forward_cfg).  Same result, ~7 % less work per call.
        self.cfg_dedup = cfg_dedup
        self._plans = {}

    # ---- per-schedule scalar table, computed with the reference's expressions on the compute device ---------------
    def _table(self, den: DiscreteDenoiser, sigmas: torch.Tensor) -> torch.Tensor:
        from . import _lib  # noqa: F401  (column indices below mirror include/ccedit_b200.h CCEDIT_SC_*)
        rows = []
        one = sigmas.new_ones([1])
        for i in range(len(sigmas) - 1):
            sigma, next_sigma = one * sigmas[i], one * sigmas[i + 1]
            if self.eta:
                sigma_up = torch.minimum(next_sigma, self.eta * (next_sigma ** 2 * (sigma ** 2 - next_sigma ** 2)
                                                                 / sigma ** 2) ** 0.5)
                sigma_down = (next_sigma ** 2 - sigma_up ** 2) ** 0.5
            else:
                sigma_up, sigma_down = torch.zeros_like(next_sigma), next_sigma

            def scal(s):                                   # DiscreteDenoiser.forward: quantise, EpsScaling
                sq = den.idx_to_sigma(den.sigma_to_idx(s))
                _, c_out, c_in, c_noise = den.scaling(sq)
                return c_in, c_out, den.sigma_to_idx(c_noise).to(torch.float32)

            c_in1, c_out1, idx1 = scal(sigma)
            euler_only = bool(torch.sum(sigma_down) < 1e-14)
            z = torch.zeros_like(sigma)
            if euler_only:
                m1 = m2 = m3 = m4 = c_in2 = c_out2 = idx2 = z
            else:
                t, t_next = -sigma.log(), -sigma_down.log()
                h = t_next - t
                s = t + 0.5 * h
                m1, m2 = (-s).exp() / (-t).exp(), (-0.5 * h).expm1()
                m3, m4 = (-t_next).exp() / (-t).exp(), (-h).expm1()
                c_in2, c_out2, idx2 = scal((-s).exp())
            rows.append(torch.cat([c_in1, c_out1, idx1, sigma, sigma_down - sigma, m1, m2, c_in2, c_out2, idx2, m3, m4,
                                   sigma_down, next_sigma, sigma_up, z + float(euler_only)]).to(torch.float32))
        return torch.stack(rows).contiguous()

    def _plan(self, den, network, x, cond, uc, num_steps):
        from . import _lib, ops
        key = (id(den), id(network), tuple(x.shape), x.device, num_steps, tuple(sorted(cond)),
               tuple((k, v.data_ptr(), tuple(v.shape)) for k, v in sorted(cond.items())),
               tuple((k, v.data_ptr(), tuple(v.shape)) for k, v in sorted(uc.items())))
        plan = self._plans.get(key)
        wver = network._weights_version() if hasattr(network, "_weights_version") else None
        if plan is not None:
            if plan["wver"] != wver:                     # weights changed (load_state_dict, LoRA merge, invalidate()):
                plan["graphs"].clear()                   # the captured step graphs replay the old packed weights
                plan["wver"] = wver
            return plan
        self._plans.clear()                                # one clip at a time: the static buffers are large
        dev, B, n = x.device, x.shape[0], x.numel()
        sigmas = self.discretization(num_steps, device=dev)
        table = self._table(den, sigmas)
        _, _, cc = self.guider.prepare_inputs(x, x.new_ones([B]), cond, uc)     # CFG concatenation, once per clip
        f32 = dict(dtype=torch.float32, device=dev)
        dedup = self.cfg_dedup and hasattr(network, "forward_cfg") and all(
            k not in cc or torch.equal(cc[k][:B], cc[k][B:]) for k in ("control_hint", "cond_feat"))
        if dedup:
            net_call = lambda p: network.forward_cfg(p["xin2"][:B], p["t2"][:B], p["cc"])
        else:
            net_call = lambda p: network(p["xin2"], p["t2"], p["cc"])
        plan = dict(net_call=net_call, dedup=dedup, wver=wver, sigmas=sigmas, table=table, euler=[bool(r) for r in table[:, 15].tolist()], cc=cc,
                    step=torch.zeros(1, dtype=torch.int32, device=dev), x=torch.empty(x.shape, **f32),
                    noise=torch.empty(x.shape, **f32), x2=torch.empty(x.shape, **f32), x_euler=torch.empty(x.shape, **f32),
                    xin2=torch.empty((2 * B,) + tuple(x.shape[1:]), **f32), t2=torch.zeros(2 * B, dtype=torch.int64, device=dev),
                    n=n, B=B, graphs={}, lib=_lib.load(), ops=ops, network=network)
        self._plans[key] = plan
        return plan

    def _step_body(self, p, euler_only: bool):
        """One step on the plan's static buffers, x updated in place (a launch sequence without host logic: capturable)."""
        lib, ops, n, B = p["lib"], p["ops"], p["n"], p["B"]
        st = torch.cuda.current_stream().cuda_stream
        scale = float(self.guider.scale)
        ptr = lambda t: t.data_ptr()
        ops._call("sampler_prepare", lib.ccedit_sampler_prepare,
                  (ptr(p["x"]), ptr(p["xin2"]), ptr(p["t2"]), ptr(p["table"]), ptr(p["step"]), n, B, st))
        eps = p["net_call"](p)
        ops._call("sampler_mid", lib.ccedit_sampler_mid,
                  (ptr(p["x"]), ptr(eps), ptr(p["x_euler"]), ptr(p["x2"]), ptr(p["xin2"]), ptr(p["t2"]), ptr(p["table"]),
                   ptr(p["step"]), scale, n, B, st))
        if not euler_only:
            eps = p["net_call"](p)
        # in place: element i of x is read and written by the same thread
        ops._call("sampler_final", lib.ccedit_sampler_final,
                  (ptr(p["x"]), ptr(p["x2"]), ptr(p["x_euler"]), ptr(eps), ptr(p["noise"]), ptr(p["x"]), ptr(p["table"]),
                   ptr(p["step"]), scale, float(self.s_noise), n, st))

    def fused_step(self, p, i: int):
        """Advance the plan's x by schedule step i; replays the captured graph of a whole step when enabled."""
        from . import ops
        p["step"].fill_(i)
        p["noise"].copy_(self.noise_sampler(p["x"]))         # one draw per step, as the reference (sampling.py:174, 186)
        euler_only = p["euler"][i]
        if not self.use_cuda_graph or torch.cuda.is_current_stream_capturing():
            self._step_body(p, euler_only)
            return
        ent = p["graphs"].get(euler_only)
        if ent is None:
            net = p["network"]
            own = getattr(net, "use_cuda_graph", None)
            if own is not None:
                net.use_cuda_graph = False                    # the step graph contains the network calls: no nested graphs
            try:
                keep = p["x"].clone()
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):                 # warm-up outside capture: packs weights, sets attributes
                    self._step_body(p, euler_only)
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                p["x"].copy_(keep)
                graph = torch.cuda.CUDAGraph()
                n0 = ops.launch_count()
                with torch.cuda.graph(graph):
                    self._step_body(p, euler_only)
                ent = p["graphs"][euler_only] = dict(graph=graph, launches=ops.launch_count() - n0)
            finally:
                if own is not None:
                    net.use_cuda_graph = own
        ent["graph"].replay()
        ops.note_graph_replay(ent["launches"])

    # ---- step-wise public API (what bench.py and a streaming caller use) ----------------------------------------
    def begin(self, denoiser: BoundDenoiser, x, cond, uc=None, num_steps=None):
        """Set up a clip: static buffers, the per-step scalar table, the CFG-concatenated conditioning (once per clip).
        Returns the plan to pass to `fused_step`; `plan["x"]` is the running latent (already scaled by sqrt(1 + sigma_0^2))."""
        uc = cond if uc is None else uc
        num_steps = self.num_steps if num_steps is None else num_steps
        with torch.no_grad():
            p = self._plan(denoiser.denoiser, denoiser.network, x, cond, uc, num_steps)
            p["x"].copy_(x)
            p["x"].mul_(torch.sqrt(1.0 + p["sigmas"][0] ** 2.0))
        return p

    def load_inputs(self, p, x=None, cond=None, uc=None):
        """Refresh the plan's device buffers from host (pinned) or device tensors without re-planning: the latent and /
        or the conditioning of the SAME clip shapes (uncond half first, as VanillaCFGTV2V.prepare_inputs orders them)."""
        B = p["B"]
        if x is not None:
            p["x"].copy_(x, non_blocking=True)
        for half, d in ((0, uc), (1, cond)):
            if d is not None:
                for k, v in d.items():
                    if k in p["cc"] and torch.is_tensor(p["cc"][k]):
                        p["cc"][k][half * B:(half + 1) * B].copy_(v, non_blocking=True)

    def __call__(self, denoiser, x, cond, uc=None, num_steps=None):
        uc = cond if uc is None else uc
        fused = (isinstance(denoiser, BoundDenoiser) and x.is_cuda and x.dtype == torch.float32
                 and isinstance(denoiser.denoiser, DiscreteDenoiser) and denoiser.denoiser.quantize_c_noise)
        if not fused:
            return super().__call__(denoiser, x, cond, uc, num_steps)
        num_steps = self.num_steps if num_steps is None else num_steps
        with torch.no_grad():
            p = self._plan(denoiser.denoiser, denoiser.network, x, cond, uc, num_steps)
            x *= torch.sqrt(1.0 + p["sigmas"][0] ** 2.0)      # in place, as the reference (sampling.py:50)
            p["x"].copy_(x)
            for i in range(len(p["sigmas"]) - 1):
                self.fused_step(p, i)
            return p["x"].clone()

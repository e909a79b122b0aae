"""ORACLE TOOLING (not product code): seeded synthetic inputs shared by make_golden.py, the tests and bench.py
(SURVEY.md section 8d "Synthetic inputs")."""
import torch


def _g(seed):
    return torch.Generator().manual_seed(seed)


def synthetic_cond(B, T, h, w, seed=3, tvi2v=False):
    """(c, uc) dicts as GeneralConditioner.get_unconditional_conditioning would return them (encoders/modules.py
    :190-204): crossattn differs, control_hint (and cond_feat) are shared between cond and uncond."""
    g = _g(seed)
    c = {"crossattn": torch.randn(B, 77, 768, generator=g)}
    uc = {"crossattn": torch.randn(B, 77, 768, generator=g)}
    c["control_hint"] = torch.rand(B, 3, T, 8 * h, 8 * w, generator=g) * 2 - 1
    uc["control_hint"] = c["control_hint"].clone()
    if tvi2v:
        c["cond_feat"] = torch.randn(B, 4, h, w, generator=g)
        uc["cond_feat"] = c["cond_feat"].clone()
    return c, uc


def synthetic_latent(B, T, h, w, seed=2):
    return torch.randn(B, 4, T, h, w, generator=_g(seed))


def cfg_batch(x, sigma_idx, c, uc):
    """What VanillaCFGTV2V.prepare_inputs + DiscreteDenoiser hand to the network: uncond first (guiders.py:56-67)."""
    cc = {k: torch.cat((uc[k], c[k]), 0) for k in c}
    return torch.cat([x] * 2), torch.cat([sigma_idx] * 2), cc


def checksum(t: torch.Tensor) -> float:
    return float(t.double().abs().sum())

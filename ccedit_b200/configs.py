"""The two inference network configs of the hot path as plain `target:` / `params:` dicts - the same nodes
`instantiate_from_config` receives from the reference's YAML files:
    tv2v    configs/inference_ccedit/keyframe_no2ndca_depthmidas.yaml:25-56
    tvi2v   configs/inference_ccedit/keyframe_ref_cp_no2ndca_add_cfca_depthzoe.yaml:32-90
(omegaconf is not installed in this image; a deployment keeps using the YAML files unchanged)."""
from __future__ import annotations

import copy

CONTROLNET_TARGET = "sgm.modules.diffusionmodules.controlmodel.ControlNet2D"
NETWORK_TARGET = "sgm.modules.diffusionmodules.controlmodel.ControlledUNetModel3DTV2V"

_BASE = dict(in_channels=4, model_channels=320, attention_resolutions=[4, 2, 1], num_res_blocks=2,
             channel_mult=[1, 2, 4, 4], num_heads=8, use_spatial_transformer=True, transformer_depth=1, context_dim=768,
             legacy=False)


def network_config(kind: str) -> dict:
    """{"target": ..., "params": {...}} of `model.params.network_config` for kind in {"tv2v", "tvi2v"}."""
    ckpt = kind == "tvi2v"
    cn = dict(_BASE, use_checkpoint=ckpt, hint_channels=3, control_scales=1.0)
    params = dict(_BASE, use_checkpoint=ckpt, out_channels=4, disable_temporal_text_ca=True,
                  controlnet_config={"target": CONTROLNET_TARGET, "params": cn})
    if kind == "tvi2v":
        params.update(enable_attention3d_crossframe=True, ST3DCA_ca_type="center_self")
        params["controlnet_img_config"] = {"target": CONTROLNET_TARGET, "params": dict(
            cn, no_add_x=True, set_input_hint_block_as_identity=True, disable_text_ca=True)}
    elif kind != "tv2v":
        raise ValueError(f"unknown config kind {kind!r}")
    return copy.deepcopy({"target": NETWORK_TARGET, "params": params})


DENOISER_CONFIG = {
    "target": "sgm.modules.diffusionmodules.denoiser.DiscreteDenoiser",
    "params": {
        "num_idx": 1000,
        "weighting_config": {"target": "sgm.modules.diffusionmodules.denoiser_weighting.EpsWeighting"},
        "scaling_config": {"target": "sgm.modules.diffusionmodules.denoiser_scaling.EpsScaling"},
        "discretization_config": {"target": "sgm.modules.diffusionmodules.discretizer.LegacyDDPMDiscretization"},
    },
}


def build_network(kind: str, device="cpu", use_cuda_graph=None, randomize_zero_init_seed=None):
    """Construct the wrapped network the way VideoDiffusionEngineTV2V does (diffusion.py:75-81): network from its config,
    wrapped by OpenAIWrapperControlLDM3DTV2V.  `randomize_zero_init_seed`: overwrite the zero-initialised tensors
    (zero_module sites) with N(0, 0.02) so that synthetic-weight runs exercise every branch (a fresh reference network
    returns eps == 0)."""
    import torch

    from .controlmodel import _instantiate_aliased
    from .wrappers import OpenAIWrapperControlLDM3DTV2V
    with torch.device(device):
        net = _instantiate_aliased(network_config(kind))
    wrap = OpenAIWrapperControlLDM3DTV2V(net, use_cuda_graph=use_cuda_graph).eval()
    if randomize_zero_init_seed is not None:
        g = torch.Generator(device=device).manual_seed(randomize_zero_init_seed)
        with torch.no_grad():
            for p in wrap.parameters():
                if p.numel() and not bool(p.any()):
                    p.copy_(torch.randn(p.shape, generator=g, device=p.device) * 0.02)
    return wrap

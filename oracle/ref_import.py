"""ORACLE TOOLING (not product code): import the UNMODIFIED reference modules from /root/reference on CPU.

Only usable where /root/reference exists (the build container) - used by oracle/make_golden.py to pin
oracle/sgm_oracle.py and to generate tests/golden/.  Recipe: SURVEY.md Appendix C.  `import sgm` itself cannot work
here (it pulls pytorch_lightning, omegaconf, kornia, open_clip, ... which are not installed), so the package
`__init__` chains are skipped with namespace stubs, three absent third-party modules are stubbed, and flash_attn /
xformers are blocked so that attention goes through CrossAttention -> F.scaled_dot_product_attention
(attention.py:444-448), the path the CPU run of the reference takes.
"""
import contextlib
import importlib.machinery
import io
import os
import sys
import types

import torch.nn as nn

REF = os.environ.get("CCEDIT_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "sgm"))


def _ns(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    m.__package__ = name
    m.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=True)
    m.__spec__.submodule_search_locations = [path]
    sys.modules[name] = m


_done = False


def install_shim():
    global _done
    if _done:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REF}")
    _ns("sgm", REF + "/sgm")
    _ns("sgm.modules", REF + "/sgm/modules")
    _ns("sgm.modules.diffusionmodules", REF + "/sgm/modules/diffusionmodules")
    sys.modules["deepspeed"] = types.ModuleType("deepspeed")
    ll = types.ModuleType("loralib")
    ll.Linear = nn.Linear
    sys.modules["loralib"] = ll
    oc = types.ModuleType("omegaconf")
    LC = type("ListConfig", (list,), {})
    oc.ListConfig, oc.OmegaConf, oc.DictConfig = LC, type("OmegaConf", (), {}), dict
    lc = types.ModuleType("omegaconf.listconfig")
    lc.ListConfig = LC
    sys.modules["omegaconf"], sys.modules["omegaconf.listconfig"] = oc, lc
    sys.modules["flash_attn"] = None
    sys.modules["xformers"] = None
    _done = True


CN_TARGET = "sgm.modules.diffusionmodules.controlmodel.ControlNet2D"


def yaml_params(kind: str):
    """network_config.params of configs/inference_ccedit/keyframe_no2ndca_depthmidas.yaml:25-56 (tv2v) and
    keyframe_ref_cp_no2ndca_add_cfca_depthzoe.yaml:32-90 (tvi2v), as plain dicts."""
    base = dict(in_channels=4, model_channels=320, attention_resolutions=[4, 2, 1], num_res_blocks=2,
                channel_mult=[1, 2, 4, 4], num_heads=8, use_spatial_transformer=True, transformer_depth=1,
                context_dim=768, legacy=False)
    cn = dict(base, use_checkpoint=False, hint_channels=3, control_scales=1.0)
    net = dict(base, use_checkpoint=False, out_channels=4, disable_temporal_text_ca=True,
               controlnet_config={"target": CN_TARGET, "params": cn})
    if kind == "tv2v":
        return net
    if kind == "tvi2v":
        net.update(enable_attention3d_crossframe=True, ST3DCA_ca_type="center_self")
        net["controlnet_img_config"] = {"target": CN_TARGET, "params": dict(
            cn, no_add_x=True, set_input_hint_block_as_identity=True, disable_text_ca=True)}
        return net
    raise ValueError(kind)


def build_reference_network(kind: str):
    """Returns the reference OpenAIWrapperControlLDM3DTV2V wrapping a freshly constructed network (CPU, fp32)."""
    install_shim()
    from sgm.modules.diffusionmodules.controlmodel import ControlledUNetModel3DTV2V
    from sgm.modules.diffusionmodules.wrappers import OpenAIWrapperControlLDM3DTV2V

    with contextlib.redirect_stdout(io.StringIO()):
        net = ControlledUNetModel3DTV2V(**yaml_params(kind))
    return OpenAIWrapperControlLDM3DTV2V(net).eval()

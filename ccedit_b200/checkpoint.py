"""Checkpoint / LoRA ingestion for the hot path: what the reference's
``scripts/sampling/util.py`` does to weights before sampling, restricted to the tensors of the denoising network.

    model_load_ckpt      scripts/sampling/util.py:45-112   checkpoint dict -> key fix-ups -> load_state_dict(strict=False)
    convert_load_lora    scripts/sampling/util.py:115-272  kohya-style ``lora_unet_*`` pairs merged into the base weights
    (callers)            scripts/sampling/sampling_tv2v.py:109, 209-234

The reference loads into the whole engine (``model.*`` = wrapper, ``first_stage_model.*`` = VAE, ``conditioner.*`` = text /
depth encoders).  Here the target is the wrapper (``ccedit_b200.wrappers.OpenAIWrapperControlLDM3DTV2V``): keys under
``model.diffusion_model.`` are routed to it, everything else is reported back untouched for the caller's other modules.
Weights stay fp32 in the reference layout inside the module; the fp16 kernel copies and captured CUDA graphs are dropped
by ``wrapper.invalidate()`` after every change made here.
"""
from __future__ import annotations

import re
from typing import Dict, Iterable, List, Mapping, Optional, Tuple

import torch

ENGINE_PREFIX = "model."                       # DiffusionEngine.model = the wrapper (diffusion.py:76-81)
NETWORK_PREFIX = "model.diffusion_model."


def read_checkpoint(path: str) -> Dict[str, torch.Tensor]:
    """File -> flat state dict, as model_load_ckpt reads it (util.py:47-60): .ckpt/.pt/.pth through torch.load (a nested
    "state_dict" entry is unwrapped; deepspeed checkpoints lose their ``_forward_module.`` prefix), .safetensors through
    the safetensors package when it is installed."""
    if path.endswith((".ckpt", ".pt", ".pth")):
        sd = torch.load(path, map_location="cpu", weights_only=False)
        if isinstance(sd, Mapping) and "state_dict" in sd:
            sd = sd["state_dict"]
        if "deepspeed" in path:
            sd = {k.replace("_forward_module.", ""): v for k, v in sd.items()}
        return dict(sd)
    if path.endswith("safetensors"):
        try:
            from safetensors.torch import load_file
        except ImportError as e:                 # not in this image; a deployment has it (requirements.txt)
            raise RuntimeError("ccedit_b200.checkpoint: reading .safetensors needs the `safetensors` package") from e
        return dict(load_file(path))
    raise NotImplementedError(f"Unknown checkpoint format: {path}")


def fix_keys(sd: Mapping[str, torch.Tensor], newbasemodel: bool = False) -> Dict[str, torch.Tensor]:
    """The two key rewrites of model_load_ckpt: a VAE nested under ``conditioner.embedders.N.`` is hoisted to
    ``first_stage_model.*`` (util.py:63-71); a plain SD-1.5 base model's ``cond_stage_model`` becomes
    ``conditioner.embedders.0`` (util.py:73-80)."""
    out = {}
    for k, v in sd.items():
        if k.startswith("conditioner.embedders.") and "first_stage_model" in k:
            k = k[k.find("first_stage_model"):]
        if newbasemodel and "cond_stage_model" in k:
            k = k.replace("cond_stage_model", "conditioner.embedders.0")
        out[k] = v
    return out


def load_network_state_dict(wrapper: torch.nn.Module, sd: Mapping[str, torch.Tensor], newbasemodel: bool = False
                            ) -> Tuple[List[str], List[str], Dict[str, torch.Tensor]]:
    """Load every ``model.diffusion_model.*`` tensor of an engine-level state dict into the wrapper (strict=False, as
    util.py:82).  Returns (missing, unexpected, rest): keys of the network the dict does not provide - with
    ``newbasemodel`` the temporal / ControlNet tensors an SD-1.5 base model cannot have are not counted (util.py:83-90) -,
    network keys the wrapper does not know, and the non-network entries (VAE, conditioner, LoRA pairs) for the caller."""
    sd = fix_keys(sd, newbasemodel)
    net = {k[len(ENGINE_PREFIX):]: v for k, v in sd.items() if k.startswith(NETWORK_PREFIX)}
    rest = {k: v for k, v in sd.items() if not k.startswith(NETWORK_PREFIX)}
    res = wrapper.load_state_dict(net, strict=False)
    missing = list(res.missing_keys)
    if newbasemodel:
        missing = [k for k in missing if "temporal" not in k and "controlnet" not in k]
    if hasattr(wrapper, "invalidate"):
        wrapper.invalidate()
    return missing, list(res.unexpected_keys), rest


def model_load_ckpt(wrapper: torch.nn.Module, path_or_sd, newbasemodel: bool = False, lora_alpha: float = 0.8):
    """Mirror of scripts/sampling/util.py:45-112 for the denoising network: read, fix keys, load, and - if the
    checkpoint carries ``lora_*`` pairs (e.g. majicmixRealistic, util.py:96-109) - merge them with alpha 0.8."""
    sd = read_checkpoint(path_or_sd) if isinstance(path_or_sd, str) else dict(path_or_sd)
    missing, unexpected, rest = load_network_state_dict(wrapper, sd, newbasemodel)
    lora = {k: v for k, v in rest.items() if k.startswith("lora")}
    if lora:
        merge_lora(wrapper, lora, alpha=lora_alpha)
    return missing, unexpected


# ---------------------------------------------------------------------------------------------------------------------
# LoRA: kohya-style key -> state-dict key of the reference network
# ---------------------------------------------------------------------------------------------------------------------
# diffusers block numbering -> SD-1.5 input/output block index (the attention layer is child 1 of the block)
_DOWN = {(0, 0): 1, (0, 1): 2, (1, 0): 4, (1, 1): 5, (2, 0): 7, (2, 1): 8}
_UP = {(1, 0): 3, (1, 1): 4, (1, 2): 5, (2, 0): 6, (2, 1): 7, (2, 2): 8, (3, 0): 9, (3, 1): 10, (3, 2): 11}
_BLOCK = re.compile(r"^lora_unet_(?:(down|up)_blocks_(\d+)_attentions_(\d+)|mid_block_attentions_0)_(.+)$")
_LEAF = (
    (re.compile(r"^proj_(in|out)$"), "proj_{0}"),
    (re.compile(r"^transformer_blocks_(\d+)_(attn[12])_to_out_(\d+)$"), "transformer_blocks.{0}.{1}.to_out.{2}"),
    (re.compile(r"^transformer_blocks_(\d+)_(attn[12])_to_([qkv])$"), "transformer_blocks.{0}.{1}.to_{2}"),
    (re.compile(r"^transformer_blocks_(\d+)_ff_net_(\d+)_proj$"), "transformer_blocks.{0}.ff.net.{1}.proj"),
    (re.compile(r"^transformer_blocks_(\d+)_ff_net_(\d+)$"), "transformer_blocks.{0}.ff.net.{1}"),
)


def lora_target_key(lora_key: str) -> Optional[str]:
    """``lora_unet_down_blocks_1_attentions_0_transformer_blocks_0_attn1_to_q.lora_down.weight`` ->
    ``model.diffusion_model.input_blocks.4.1.transformer_blocks.0.attn1.to_q.weight``; None for text-encoder pairs
    (``lora_te_*``: the CLIP encoder is outside the hot path).  Raises ValueError for a UNet key it cannot place, as
    the reference does (util.py:170, 233)."""
    stem = lora_key.split(".")[0]
    if stem.startswith("lora_te"):
        return None
    m = _BLOCK.match(stem)
    if m is None:
        raise ValueError(f"Unknown key: {lora_key}")
    side, blk, att, leaf = m.groups()
    if side is None:
        block = "middle_block.1"
    else:
        table, name = (_DOWN, "input_blocks") if side == "down" else (_UP, "output_blocks")
        idx = table.get((int(blk), int(att)))
        if idx is None:
            raise ValueError(f"Unknown key: {lora_key}")
        block = f"{name}.{idx}.1"
    for pat, fmt in _LEAF:
        lm = pat.match(leaf)
        if lm is not None:
            return f"{NETWORK_PREFIX}{block}.{fmt.format(*lm.groups())}.weight"
    raise ValueError(f"Unknown key: {lora_key}")


def lora_deltas(lora_sd: Mapping[str, torch.Tensor], alpha: float) -> Dict[str, torch.Tensor]:
    """target key -> alpha * up @ down (fp32) for every ``lora_up`` / ``lora_down`` pair; ``.alpha`` entries are ignored
    exactly as the reference ignores them (util.py:129-131: "as we have set the alpha beforehand")."""
    out: Dict[str, torch.Tensor] = {}
    for key in lora_sd:
        if ".alpha" in key or "lora_down" not in key:
            continue
        target = lora_target_key(key)
        if target is None:
            continue
        up = lora_sd[key.replace("lora_down", "lora_up")].to(torch.float32)
        down = lora_sd[key].to(torch.float32)
        if up.dim() == 4:                        # 1x1-conv LoRA (proj_in / proj_out): [r, C, 1, 1]
            delta = torch.mm(up[:, :, 0, 0], down[:, :, 0, 0])[:, :, None, None]
        else:
            delta = torch.mm(up, down)
        delta = alpha * delta
        out[target] = out[target] + delta if target in out else delta
    return out


def merge_lora(wrapper: torch.nn.Module, lora_sd: Mapping[str, torch.Tensor], alpha: float = 0.8) -> List[str]:
    """Merge LoRA pairs into the wrapper's weights in place (W += alpha * up @ down, sampling_tv2v.py:211-234) and drop
    the packed fp16 copies / captured graphs.  Returns the state-dict keys that were modified."""
    params = dict(wrapper.named_parameters())
    touched = []
    with torch.no_grad():
        for target, delta in lora_deltas(lora_sd, alpha).items():
            name = target[len(ENGINE_PREFIX):]
            p = params.get(name)
            if p is None:
                raise KeyError(f"LoRA target {target} is not a parameter of the network")
            d = delta.to(device=p.device, dtype=p.dtype)
            p.add_(d.reshape(p.shape) if d.numel() == p.numel() else d)
            touched.append(target)
    if hasattr(wrapper, "invalidate"):
        wrapper.invalidate()
    return touched


def engine_state_dict(wrapper: torch.nn.Module, extra: Optional[Iterable[Tuple[str, torch.Tensor]]] = None
                      ) -> Dict[str, torch.Tensor]:
    """The wrapper's tensors under the engine-level keys a CCEdit checkpoint uses (``model.diffusion_model.*``)."""
    sd = {ENGINE_PREFIX + k: v for k, v in wrapper.state_dict().items()}
    sd.update(dict(extra or ()))
    return sd

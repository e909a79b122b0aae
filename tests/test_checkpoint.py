"""CPU tests of checkpoint / LoRA ingestion (ccedit_b200/checkpoint.py) against the reference's
scripts/sampling/util.py:45-272.  The key mapping is pinned two ways: a table of known answers written out from the
reference's format strings (util.py:152-232), and - in the build container, where /root/reference exists - by executing
the reference's own `convert_load_lora` (extracted from its source file, which cannot be imported as a module: it pulls
cv2 / omegaconf / safetensors) on the same synthetic LoRA and comparing the merged tensors."""
import ast
import os
import sys
import types

import pytest
import torch

from ccedit_b200 import checkpoint as ck

REF_UTIL = "/root/reference/scripts/sampling/util.py"

KNOWN = {
    "lora_unet_down_blocks_0_attentions_0_proj_in.lora_down.weight": "input_blocks.1.1.proj_in.weight",
    "lora_unet_down_blocks_1_attentions_0_transformer_blocks_0_attn1_to_q.lora_down.weight":
        "input_blocks.4.1.transformer_blocks.0.attn1.to_q.weight",
    "lora_unet_down_blocks_2_attentions_1_transformer_blocks_0_attn2_to_out_0.lora_down.weight":
        "input_blocks.8.1.transformer_blocks.0.attn2.to_out.0.weight",
    "lora_unet_down_blocks_1_attentions_1_transformer_blocks_0_ff_net_0_proj.lora_down.weight":
        "input_blocks.5.1.transformer_blocks.0.ff.net.0.proj.weight",
    "lora_unet_up_blocks_1_attentions_0_transformer_blocks_0_ff_net_2.lora_down.weight":
        "output_blocks.3.1.transformer_blocks.0.ff.net.2.weight",
    "lora_unet_up_blocks_3_attentions_2_transformer_blocks_0_attn2_to_k.lora_down.weight":
        "output_blocks.11.1.transformer_blocks.0.attn2.to_k.weight",
    "lora_unet_up_blocks_2_attentions_1_proj_out.lora_down.weight": "output_blocks.7.1.proj_out.weight",
    "lora_unet_mid_block_attentions_0_proj_in.lora_down.weight": "middle_block.1.proj_in.weight",
    "lora_unet_mid_block_attentions_0_transformer_blocks_0_attn1_to_v.lora_down.weight":
        "middle_block.1.transformer_blocks.0.attn1.to_v.weight",
    "lora_unet_mid_block_attentions_0_transformer_blocks_0_attn2_to_out_0.lora_down.weight":
        "middle_block.1.transformer_blocks.0.attn2.to_out.0.weight",
    "lora_unet_mid_block_attentions_0_transformer_blocks_0_ff_net_0_proj.lora_down.weight":
        "middle_block.1.transformer_blocks.0.ff.net.0.proj.weight",
}


def test_lora_key_mapping_known_answers():
    for k, want in KNOWN.items():
        assert ck.lora_target_key(k) == "model.diffusion_model." + want
    assert ck.lora_target_key("lora_te_text_model_encoder_layers_0_self_attn_k_proj.lora_down.weight") is None
    with pytest.raises(ValueError):
        ck.lora_target_key("lora_unet_down_blocks_3_attentions_0_proj_in.lora_down.weight")   # level 3 has no attention
    with pytest.raises(ValueError):
        ck.lora_target_key("lora_unet_down_blocks_0_resnets_0_conv1.lora_down.weight")


def _shape_of(target):
    """Shape of the SD-1.5 tensor behind a mapped key (channel width from the block index)."""
    parts = target.split(".")
    blk, idx = parts[2], int(parts[3]) if parts[3].isdigit() else None
    if blk == "middle_block":
        C = 1280
    elif blk == "input_blocks":
        C = {1: 320, 2: 320, 4: 640, 5: 640, 7: 1280, 8: 1280}[idx]
    else:
        C = {3: 1280, 4: 1280, 5: 1280, 6: 640, 7: 640, 8: 640, 9: 320, 10: 320, 11: 320}[idx]
    if "proj_in" in target or "proj_out" in target:
        return (C, C, 1, 1)
    if "ff.net.0.proj" in target:
        return (8 * C, C)
    if "ff.net.2" in target:
        return (C, 4 * C)
    if "attn2.to_k" in target or "attn2.to_v" in target:
        return (C, 768)
    return (C, C)


def _synthetic_lora(rank=4):
    g = torch.Generator().manual_seed(0)
    lora, base = {}, {}
    for k, tail in KNOWN.items():
        shape = _shape_of("model.diffusion_model." + tail)
        n, kin = shape[0], shape[1]
        conv = len(shape) == 4
        down = torch.randn((rank, kin, 1, 1) if conv else (rank, kin), generator=g) * 0.1
        up = torch.randn((n, rank, 1, 1) if conv else (n, rank), generator=g) * 0.1
        lora[k] = down
        lora[k.replace("lora_down", "lora_up")] = up
        lora[k.split(".")[0] + ".alpha"] = torch.tensor(4.0)
        base["model.diffusion_model." + tail] = torch.randn(shape, generator=g)
    return lora, base


def test_lora_deltas_match_up_times_down():
    lora, base = _synthetic_lora()
    deltas = ck.lora_deltas(lora, alpha=0.8)
    assert set(deltas) == set(base)
    for k, tail in KNOWN.items():
        up, down = lora[k.replace("lora_down", "lora_up")], lora[k]
        want = 0.8 * (up.flatten(1) @ down.flatten(1))
        got = deltas["model.diffusion_model." + tail]
        assert got.shape == base["model.diffusion_model." + tail].shape
        assert torch.allclose(got.flatten(1), want)


@pytest.mark.skipif(not os.path.exists(REF_UTIL), reason="reference tree not present (build container only)")
def test_lora_merge_matches_reference_convert_load_lora():
    """Run the REFERENCE's convert_load_lora (util.py:115-272) on the synthetic LoRA and compare every merged tensor."""
    src = open(REF_UTIL).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "convert_load_lora")
    tq = types.ModuleType("tqdm")
    tq.tqdm = lambda it, *a, **k: it
    ns = {"torch": torch, "tqdm": tq, "print": lambda *a, **k: None}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), REF_UTIL, "exec"), ns)
    lora, base = _synthetic_lora()
    ref_sd = ns["convert_load_lora"]({k: v.clone() for k, v in base.items()}, dict(lora), alpha=0.6)
    deltas = ck.lora_deltas(lora, alpha=0.6)
    for k, v in base.items():
        assert torch.allclose(v + deltas[k], ref_sd[k], rtol=1e-6, atol=1e-6), k


def test_fix_keys_and_prefix_routing():
    sd = {
        "model.diffusion_model.time_embed.0.weight": torch.zeros(2),
        "conditioner.embedders.1.first_stage_model.decoder.conv_in.weight": torch.zeros(1),
        "cond_stage_model.transformer.text_model.embeddings.position_ids": torch.zeros(1),
        "lora_unet_mid_block_attentions_0_proj_in.lora_down.weight": torch.zeros(1),
    }
    out = ck.fix_keys(sd, newbasemodel=True)
    assert "first_stage_model.decoder.conv_in.weight" in out                                   # util.py:63-71
    assert "conditioner.embedders.0.transformer.text_model.embeddings.position_ids" in out      # util.py:73-80
    assert "model.diffusion_model.time_embed.0.weight" in out

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.diffusion_model = torch.nn.ModuleDict({"time_embed": torch.nn.ModuleDict({"0": torch.nn.Linear(2, 2)})})
            self.invalidated = 0

        def invalidate(self):
            self.invalidated += 1

    t = Tiny()
    w = torch.randn(2, 2)
    missing, unexpected, rest = ck.load_network_state_dict(t, {
        "model.diffusion_model.time_embed.0.weight": w, "model.diffusion_model.nope.weight": w,
        "first_stage_model.x": w})
    assert torch.equal(t.diffusion_model["time_embed"]["0"].weight, w)
    assert missing == ["diffusion_model.time_embed.0.bias"] and unexpected == ["diffusion_model.nope.weight"]
    assert list(rest) == ["first_stage_model.x"] and t.invalidated == 1


def test_read_checkpoint_formats(tmp_path):
    w = {"a": torch.arange(3.0)}
    p1, p2 = str(tmp_path / "m.ckpt"), str(tmp_path / "deepspeed_m.pt")
    torch.save({"state_dict": w}, p1)
    torch.save({"_forward_module.a": w["a"]}, p2)
    assert torch.equal(ck.read_checkpoint(p1)["a"], w["a"])
    assert torch.equal(ck.read_checkpoint(p2)["a"], w["a"])
    with pytest.raises(NotImplementedError):
        ck.read_checkpoint(str(tmp_path / "m.bin"))

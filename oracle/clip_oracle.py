"""ORACLE (test infrastructure, NOT product code): CPU fp32 restatement of the CLIP text transformer behind the reference's
FrozenCLIPEmbedder (sgm/modules/encoders/modules.py:358-420, `layer="last"`).

The algorithm is not in /root/reference: FrozenCLIPEmbedder calls HuggingFace transformers' CLIPTextModel
(requirements.txt:34 pins transformers==4.19.1; this image has 5.5.0).  Restated here from its published definition
(CLIPTextTransformer: token + learned position embeddings; 12 pre-LN encoder layers with causal multi-head attention -
q scaled by head_dim^-1/2 - and a quick_gelu MLP; final LayerNorm) and pinned against the installed transformers
implementation on seeded random weights (tests/test_oracle_golden.py::test_clip_text_oracle_matches_transformers).
No pretrained weights or tokenizer vocabulary exist in this image: parity is on seeded weights and random token ids.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

CLIP_L = dict(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
              max_position_embeddings=77)


def seeded_state_dict(shapes: Dict[str, tuple], seed: int = 0) -> Dict[str, torch.Tensor]:
    """key -> fp32 tensor for the keys / shapes of a CLIPTextModel state dict (oracle.weights.seeded_tensor rules)."""
    from oracle.weights import seeded_tensor
    out = {}
    for k, shp in shapes.items():
        if k.endswith("position_ids"):
            continue
        init = "ones" if ("layer_norm" in k and k.endswith(".weight")) else "default"
        out[k] = seeded_tensor(k, shp, init, seed)
    return out


def clip_text_forward(sd: Dict[str, torch.Tensor], ids: torch.Tensor, cfg=CLIP_L, p="text_model.") -> torch.Tensor:
    """last_hidden_state [B, L, D] of CLIPTextModel for token ids [B, L]."""
    B, L = ids.shape
    D, H = cfg["hidden_size"], cfg["num_attention_heads"]
    dh = D // H
    x = sd[p + "embeddings.token_embedding.weight"][ids] + sd[p + "embeddings.position_embedding.weight"][:L][None]
    mask = torch.full((L, L), float("-inf")).triu(1)
    for i in range(cfg["num_hidden_layers"]):
        q_ = f"{p}encoder.layers.{i}."
        lin = lambda n, t: F.linear(t, sd[q_ + n + ".weight"], sd[q_ + n + ".bias"])
        h = F.layer_norm(x, (D,), sd[q_ + "layer_norm1.weight"], sd[q_ + "layer_norm1.bias"], 1e-5)
        q = lin("self_attn.q_proj", h) * dh ** -0.5
        k, v = lin("self_attn.k_proj", h), lin("self_attn.v_proj", h)
        sp = lambda t: t.view(B, L, H, dh).transpose(1, 2)
        w = torch.softmax(sp(q) @ sp(k).transpose(-1, -2) + mask, dim=-1)
        a = (w @ sp(v)).transpose(1, 2).reshape(B, L, D)
        x = x + lin("self_attn.out_proj", a)
        h = F.layer_norm(x, (D,), sd[q_ + "layer_norm2.weight"], sd[q_ + "layer_norm2.bias"], 1e-5)
        h = lin("mlp.fc1", h)
        h = h * torch.sigmoid(1.702 * h)                                   # quick_gelu
        x = x + lin("mlp.fc2", h)
    return F.layer_norm(x, (D,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], 1e-5)


def transformers_reference(sd: Dict[str, torch.Tensor], ids: torch.Tensor, cfg=CLIP_L) -> torch.Tensor:
    """The same through the installed transformers.CLIPTextModel (the third-party implementation the reference calls)."""
    from transformers import CLIPTextConfig, CLIPTextModel
    m = CLIPTextModel(CLIPTextConfig(hidden_act="quick_gelu", layer_norm_eps=1e-5, projection_dim=cfg["hidden_size"], **cfg)).eval()
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.endswith("position_ids") for k in missing), (missing, unexpected)
    with torch.no_grad():
        return m(input_ids=ids).last_hidden_state

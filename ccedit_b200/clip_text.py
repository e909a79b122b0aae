"""Text conditioner (first half): mirror of the reference's FrozenCLIPEmbedder
(sgm/modules/encoders/modules.py:358-420), which wraps HuggingFace transformers' CLIPTextModel
("openai/clip-vit-large-patch14": 12 pre-LN layers, width 768, 12 heads, MLP 3072 with quick_gelu, causal mask, learned
positions, final LayerNorm; `layer="last"` returns last_hidden_state [B, 77, 768] = c["crossattn"]).  The algorithm lives in
that third-party package (pinned 4.19.1 in the reference's requirements.txt:34, not vendored); oracle/clip_oracle.py
restates it and is pinned against the installed transformers implementation.

Same state-dict keys as the reference module (`transformer.text_model.…`), so a CCEdit / SD-1.5 checkpoint's
`conditioner.embedders.0.*` (or `cond_stage_model.*`) tensors load unchanged.  Execution: token + position gather kernel,
every LayerNorm folded into the GEMM it feeds (tcgen05 tap-GEMM), a small causal attention kernel, quick_gelu in place,
residual adds in the GEMM epilogues.  Tokenisation stays host-side: `forward(text)` uses transformers' CLIPTokenizer when its
vocabulary files are available (they are not in this image: no network), `encode_tokens(ids)` takes token ids directly.
The depth conditioners (MiDaS / ZoeDepth, encoders/modules.py:1289-1392) import `src.controlnet11`, which is not part of
the reference tree, and are not built.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .modules import ParamHolder, linear, norm


class _Embeddings(nn.Module):
    def __init__(self, vocab, width, max_len):
        super().__init__()
        self.token_embedding = nn.Embedding(vocab, width)
        self.position_embedding = nn.Embedding(max_len, width)
        self._cache = None

    def tables(self, dev):
        ver = (dev, self.token_embedding.weight._version, self.position_embedding.weight._version,
               self.token_embedding.weight.data_ptr())
        if self._cache is None or self._cache[0] != ver:
            self._cache = (ver, self.token_embedding.weight.detach().to(device=dev, dtype=torch.float16).contiguous(),
                           self.position_embedding.weight.detach().to(device=dev, dtype=torch.float16).contiguous())
        return self._cache[1], self._cache[2]


class _Attention(nn.Module):
    def __init__(self, width, heads):
        super().__init__()
        self.heads = heads
        self.k_proj, self.v_proj, self.q_proj = linear(width, width), linear(width, width), linear(width, width)
        self.out_proj = linear(width, width)
        self._qkv = None

    def qkv_packed(self, dev, ln: ParamHolder):
        hs = (self.q_proj, self.k_proj, self.v_proj)
        ver = (dev,) + tuple(h._version() for h in hs) + (ln._version(),)
        if self._qkv is None or self._qkv[0] != ver:
            w = torch.cat([h.weight.detach().float() for h in hs], 0)
            b = torch.cat([h.bias.detach().float() for h in hs], 0)
            self._qkv = (ver, ops.pack_weight(w, b, dev, ln_gamma=ln.weight, ln_beta=ln.bias))
        return self._qkv[1]


class _MLP(nn.Module):
    def __init__(self, width, inner):
        super().__init__()
        self.fc1, self.fc2 = linear(width, inner), linear(inner, width)


class _Layer(nn.Module):
    """CLIPEncoderLayer: x += attn(LN1 x); x += fc2(quick_gelu(fc1(LN2 x)))."""

    def __init__(self, width, heads, inner):
        super().__init__()
        self.self_attn = _Attention(width, heads)
        self.layer_norm1 = norm(width)
        self.mlp = _MLP(width, inner)
        self.layer_norm2 = norm(width)

    def run(self, x: torch.Tensor, B: int, L: int) -> torch.Tensor:
        dev, (M, D) = x.device, x.shape
        a = self.self_attn
        qkv = ops.gemm(x, a.qkv_packed(dev, self.layer_norm1), torch.empty(M, 3 * D, dtype=torch.float16, device=dev),
                       rowstats=ops.layernorm_stats(x))
        q3 = qkv.view(B, L, 3 * D)
        att = ops.causal_attention_small(q3[..., :D], q3[..., D:2 * D], q3[..., 2 * D:], a.heads,
                                         torch.empty(B, L, D, dtype=torch.float16, device=dev))
        x = ops.gemm(att.view(M, D), a.out_proj.packed(dev), torch.empty_like(x), res1=x)
        h = ops.gemm(x, self.mlp.fc1.packed_ln(dev, self.layer_norm2),
                     torch.empty(M, self.mlp.fc1.weight.shape[0], dtype=torch.float16, device=dev), rowstats=ops.layernorm_stats(x))
        ops.quick_gelu_(h)
        return ops.gemm(h, self.mlp.fc2.packed(dev), torch.empty_like(x), res1=x)


class _Encoder(nn.Module):
    def __init__(self, width, heads, inner, layers):
        super().__init__()
        self.layers = nn.ModuleList([_Layer(width, heads, inner) for _ in range(layers)])


class _TextTransformer(nn.Module):
    def __init__(self, vocab, width, heads, inner, layers, max_len):
        super().__init__()
        self.embeddings = _Embeddings(vocab, width, max_len)
        self.encoder = _Encoder(width, heads, inner, layers)
        self.final_layer_norm = norm(width)


class CLIPTextModel(nn.Module):
    """State-dict compatible with transformers.CLIPTextModel (keys `text_model.…`)."""

    def __init__(self, vocab_size=49408, hidden_size=768, num_attention_heads=12, intermediate_size=3072,
                 num_hidden_layers=12, max_position_embeddings=77):
        super().__init__()
        if hidden_size // num_attention_heads != 64:
            raise NotImplementedError("ccedit_b200: the text transformer's attention kernel is built for head dim 64")
        self.text_model = _TextTransformer(vocab_size, hidden_size, num_attention_heads, intermediate_size,
                                           num_hidden_layers, max_position_embeddings)
        self.max_len = max_position_embeddings

    def encode_tokens(self, ids: torch.Tensor) -> torch.Tensor:
        """ids: int64 [B, L <= max_position_embeddings] -> last_hidden_state fp32 [B, L, width]."""
        if not ids.is_cuda:
            raise RuntimeError("ccedit_b200: the text transformer runs on CUDA (sm_100a) only; there is no CPU fallback")
        B, L = ids.shape
        if L > self.max_len:
            raise RuntimeError(f"ccedit_b200: at most {self.max_len} tokens")
        tm = self.text_model
        with torch.no_grad():
            x = ops.embed_tokens(ids, *tm.embeddings.tables(ids.device))
            for layer in tm.encoder.layers:
                x = layer.run(x, B, L)
            y = ops.layernorm(x, *tm.final_layer_norm.affine(ids.device), 1e-5)
        return y.view(B, L, -1).float()


class FrozenCLIPEmbedder(nn.Module):
    """encoders/modules.py:358-420 with `layer="last"` (the inference configs' setting)."""

    def __init__(self, version="openai/clip-vit-large-patch14", device="cuda", max_length=77, freeze=True, layer="last",
                 layer_idx=None, always_return_pooled=False):
        super().__init__()
        if layer != "last" or always_return_pooled:
            raise NotImplementedError("ccedit_b200: FrozenCLIPEmbedder is built for layer='last' without the pooled output")
        self.version, self.device, self.max_length = version, device, max_length
        self.transformer = CLIPTextModel(max_position_embeddings=max_length)
        self._tokenizer = None
        if freeze:
            self.eval()
            for p in self.parameters():
                p.requires_grad = False

    def tokenize(self, text):
        if self._tokenizer is None:
            try:
                import os
                from transformers import CLIPTokenizer
                try:                                     # cached vocabulary first; download only when explicitly allowed
                    self._tokenizer = CLIPTokenizer.from_pretrained(self.version, local_files_only=True)
                except Exception:
                    if os.environ.get("CCEDIT_HF_DOWNLOAD", "0") != "1":
                        raise
                    self._tokenizer = CLIPTokenizer.from_pretrained(self.version)
            except Exception as e:                       # offline image: the vocabulary files are not available
                raise RuntimeError("ccedit_b200.FrozenCLIPEmbedder: CLIPTokenizer vocabulary for "
                                   f"{self.version!r} is not available; pass token ids to encode_tokens()") from e
        enc = self._tokenizer(text, truncation=True, max_length=self.max_length, return_length=True,
                              return_overflowing_tokens=False, padding="max_length", return_tensors="pt")
        return enc["input_ids"]

    def encode_tokens(self, ids: torch.Tensor) -> torch.Tensor:
        return self.transformer.encode_tokens(ids.to(self.device))

    def forward(self, text):
        ids = text if torch.is_tensor(text) else self.tokenize(text)
        return self.encode_tokens(ids)

    def encode(self, text):
        return self(text)

"""Host-side mirror of the reference's building blocks (openaimodel.py / attention.py), executed by the CUDA kernels.

Every class keeps the reference's parameter names and shapes, so the reference's state-dict keys load unchanged
(SURVEY.md Appendix D), but `forward` never runs a torch op on activations: each block is a short list of C-ABI kernel
calls (ccedit_b200.ops) on channels-last fp16 buffers:

    video tensor   [B, T, H, W, C]      (the reference's "b c t h w")
    frame batch    [F=B*T, H, W, C]     (the reference's "(b t) c h w")     -- same memory, no copy
    pixel batch    [B, T, HW, C]        (the reference's "(b h w) c t")     -- same memory, no copy

Parameters live in the reference layout/dtype (for load_state_dict / LoRA merging); `pack()` derives the fp16
kernel-native copies (tap-major conv weights, fused QKV / KV, GEGLU-interleaved FF-in).
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch
import torch.nn as nn

from . import ops
from .ops import KVSegment, PackedWeight

GN_EPS_RES = 1e-5    # normalization(): util.py:296-302
GN_EPS_ATTN = 1e-6   # Normalize():     attention.py:153-156


# ---------------------------------------------------------------------------------------------------------------------
# parameter holders (names match torch.nn.{Conv2d,Conv1d,Linear,GroupNorm,LayerNorm} so state-dict keys line up)
# ---------------------------------------------------------------------------------------------------------------------
class ParamHolder(nn.Module):
    """weight (+ bias) with the reference's shape and default init; `zero=True` mirrors zero_module()."""

    def __init__(self, shape, bias=True, zero=False, norm=False):
        super().__init__()
        w = torch.empty(*shape)
        if norm:
            nn.init.ones_(w)
        elif zero:
            nn.init.zeros_(w)
        else:
            nn.init.kaiming_uniform_(w.view(shape[0], -1), a=math.sqrt(5))
        self.weight = nn.Parameter(w)
        if bias:
            b = torch.zeros(shape[0])
            if not (norm or zero):
                bound = 1.0 / math.sqrt(max(1, math.prod(shape[1:])))
                nn.init.uniform_(b, -bound, bound)
            self.bias = nn.Parameter(b)
        else:
            self.register_parameter("bias", None)
        self._cache = {}

    def forward(self, *a, **k):  # never used for compute
        raise RuntimeError("ccedit_b200 parameter holders are not callable; use the owning block's forward")

    # ---- packed views (rebuilt when the parameter is replaced or modified in place, e.g. by a LoRA merge) ----------
    def _version(self):
        b = self.bias
        return (self.weight.data_ptr(), self.weight._version, None if b is None else (b.data_ptr(), b._version))

    def _cached(self, key, make):
        ver = self._version()
        c = self._cache.get(key)
        if c is None or c[0] != ver:
            self._cache[key] = c = (ver, make())
        return c[1]

    def packed(self, device, geglu=False, scale: float = 1.0, max_bn: int = 256) -> PackedWeight:
        def make():
            w = self.weight if scale == 1.0 else self.weight.detach().float() * scale
            b = self.bias if (scale == 1.0 or self.bias is None) else self.bias.detach().float() * scale
            return ops.pack_weight(w, b, device, geglu=geglu, max_bn=max_bn)
        return self._cached((device, "w", geglu, scale, max_bn), make)

    def packed_ln(self, device, ln: "ParamHolder", geglu=False) -> PackedWeight:
        """This Linear with the LayerNorm ``ln`` that feeds it folded in (ops.pack_weight(ln_gamma=...)); rebuilt when
        either parameter set changes."""
        key = (device, "w_ln", geglu)
        ver = (self._version(), ln._version())
        c = self._cache.get(key)
        if c is None or c[0] != ver:
            self._cache[key] = c = (ver, ops.pack_weight(self.weight, self.bias, device, geglu=geglu,
                                                         ln_gamma=ln.weight, ln_beta=ln.bias))
        return c[1]

    def affine(self, device):
        return self._cached((device, "affine"), lambda: (
            self.weight.detach().to(device=device, dtype=torch.float32).contiguous(),
            self.bias.detach().to(device=device, dtype=torch.float32).contiguous()))

    def invalidate(self):
        self._cache = {}


def conv2d(cin, cout, k, zero=False):
    return ParamHolder((cout, cin, k, k), zero=zero)


def conv1d(cin, cout, k, zero=False):
    return ParamHolder((cout, cin, k), zero=zero)


def linear(cin, cout, bias=True, zero=False):
    return ParamHolder((cout, cin), bias=bias, zero=zero)


def norm(c):
    return ParamHolder((c,), norm=True)


def seq(**mods):
    """nn.ModuleDict keyed by the index the layer has inside the reference's nn.Sequential."""
    return nn.ModuleDict({k.lstrip("_"): v for k, v in mods.items()})


def _fused(device, holders: List[ParamHolder], geglu=False, ln: Optional[ParamHolder] = None) -> PackedWeight:
    """Pack the row-concatenation of several bias-free projections (QKV / KV) as one weight (optionally with the
    LayerNorm that feeds all of them folded in)."""
    w = torch.cat([h.weight.detach().float() for h in holders], 0)
    if ln is None:
        return ops.pack_weight(w, None, device, geglu=geglu)
    return ops.pack_weight(w, None, device, geglu=geglu, ln_gamma=ln.weight, ln_beta=ln.bias)


def _new(ref: torch.Tensor, *shape):
    return torch.empty(*shape, dtype=torch.float16, device=ref.device)


# ---------------------------------------------------------------------------------------------------------------------
# per-call context
# ---------------------------------------------------------------------------------------------------------------------
class Ctx:
    """What a block needs beyond its input: batch geometry, per-block time-embedding rows, text K/V."""

    def __init__(self, B, T, device):
        self.B, self.T, self.device = B, T, device
        self.emb_rows = {}   # id(resblock) -> fp32 [B, Cout] view
        self.text_kv = {}    # id(CrossAttention) -> (k [B,77,C] view, v view)


# ---------------------------------------------------------------------------------------------------------------------
# attention.py mirrors
# ---------------------------------------------------------------------------------------------------------------------
class CrossAttention(nn.Module):
    """attention.py:365-467. to_q/to_k/to_v without bias, to_out.0 with bias; SDPA with scale d^-1/2."""

    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64):
        super().__init__()
        inner = heads * dim_head
        context_dim = query_dim if context_dim is None else context_dim
        self.heads, self.dim_head, self.inner = heads, dim_head, inner
        self.to_q = linear(query_dim, inner, bias=False)
        self.to_k = linear(context_dim, inner, bias=False)
        self.to_v = linear(context_dim, inner, bias=False)
        self.to_out = seq(_0=linear(inner, query_dim))
        self._pk = {}

    def fused(self, device, which, ln: Optional[ParamHolder] = None):
        hs = {"qkv": [self.to_q, self.to_k, self.to_v], "kv": [self.to_k, self.to_v]}[which]
        ver = tuple(h._version() for h in hs) + ((ln._version(),) if ln is not None else ())
        key = (device, which, ln is not None)
        c = self._pk.get(key)
        if c is None or c[0] != ver:
            self._pk[key] = c = (ver, _fused(device, hs, ln=ln))
        return c[1]

    def invalidate(self):
        self._pk = {}


class FeedForward(nn.Module):
    """attention.py:125-141 with GEGLU (:115-122): net.0.proj [8C, C], net.2 [C, 4C]."""

    def __init__(self, dim, mult=4):
        super().__init__()
        inner = dim * mult
        self.net = seq(_0=nn.ModuleDict({"proj": linear(dim, inner * 2)}), _2=linear(inner, dim))

    def run(self, x, ln: ParamHolder, stats, out=None):
        """x + net.2(GEGLU(net.0(LN(x)))): the LayerNorm ``ln`` is folded into the GEGLU GEMM (``stats`` = row statistics)."""
        dev = x.device
        g = ops.gemm(x, self.net["0"]["proj"].packed_ln(dev, ln, geglu=True),
                     _new(x, x.shape[0], self.net["2"].weight.shape[1]), rowstats=stats)
        out = _new(x, *x.shape) if out is None else out
        return ops.gemm(g, self.net["2"].packed(dev), out, res1=x)


class BasicTransformerBlock(nn.Module):
    """attention.py:598-716: x += attn1(LN1 x); x += attn2(LN2 x, text); x += FF(LN3 x).  Tokens: [F*HW, C]."""

    def __init__(self, dim, n_heads, d_head, context_dim):
        super().__init__()
        self.attn1 = CrossAttention(dim, None, n_heads, d_head)
        self.ff = FeedForward(dim)
        self.attn2 = CrossAttention(dim, context_dim, n_heads, d_head)
        self.norm1, self.norm2, self.norm3 = norm(dim), norm(dim), norm(dim)

    def run(self, x, F, L, ctx: Ctx, sp=None, dup=False):
        """sp: (sum, sum of squares) partials of x written by the GEMM that produced it (ops.gemm(stats_out=)), or None.
        dup: x holds ONE copy of frames that the uncond and cond halves of a CFG batch share (see
        SpatialTransformer.run_spatial); everything up to the text cross-attention runs once, then the rows fan out and
        the result has 2 * F frames."""
        dev, C, M = x.device, x.shape[1], x.shape[0]
        heads = self.attn1.heads
        # The three LayerNorms are folded into the GEMMs they feed.  Their row statistics come from the epilogue of the
        # GEMM that produced x (proj_in, attn1.to_out, attn2.to_out) - no pass over the activation at all - or, when the
        # caller has none, from one statistics pass (ops.layernorm_stats).
        st = ops.layernorm_stats(x) if sp is None else sp      # partial sums are finished by the consuming epilogue
        qkv = ops.gemm(x, self.attn1.fused(dev, "qkv", ln=self.norm1), _new(x, M, 3 * C), rowstats=st)
        q3 = qkv.view(F, L, 3 * C)
        att = ops.attention(q3[..., :C], [KVSegment(q3[..., C:2 * C], q3[..., 2 * C:])], heads, _new(x, F, L, C))
        o1 = self.attn1.to_out["0"].packed(dev)
        sp = torch.empty(M, ops.stats_slots(o1), 2, dtype=torch.float32, device=dev)
        x = ops.gemm(att.view(M, C), o1, _new(x, M, C), res1=x, stats_out=sp)
        if dup:                                                # attn2 reads the text: from here the halves differ
            x, sp = ops.dup_rows(x), ops.dup_rows(sp)
            F, M = 2 * F, 2 * M
            att = _new(x, F, L, C)
        q = ops.gemm(x, self.attn2.to_q.packed_ln(dev, self.norm2), _new(x, M, C), rowstats=sp)
        k, v = ctx.text_kv[id(self.attn2)]
        att = ops.attention(q.view(F, L, C), [KVSegment(k, v, div=ctx.T)], heads, att)
        o2 = self.attn2.to_out["0"].packed(dev)
        sp2 = sp if ops.stats_slots(o2) == sp.shape[1] else torch.empty(M, ops.stats_slots(o2), 2, dtype=torch.float32, device=dev)
        x = ops.gemm(att.view(M, C), o2, _new(x, M, C), res1=x, stats_out=sp2)
        return self.ff.run(x, self.norm3, sp2)


class BasicTransformerSingleLayerBlock(nn.Module):
    """attention.py:719-761: x += attn1(LN1 x, context); x += FF(LN2 x).  K/V come from the UN-normalised context."""

    def __init__(self, dim, n_heads, d_head):
        super().__init__()
        self.attn1 = CrossAttention(dim, None, n_heads, d_head)
        self.ff = FeedForward(dim)
        self.norm1, self.norm2 = norm(dim), norm(dim)

    def run(self, x, attend, out=None, sp=None):
        """x: tokens [M, C]; attend(q [M,C], kv [M,2C]) -> [M, C] runs the attention of the calling layer; sp: statistics
        partials of x from the GEMM that produced it (see BasicTransformerBlock.run)."""
        dev, C, M = x.device, x.shape[1], x.shape[0]
        st = ops.layernorm_stats(x) if sp is None else sp      # partial sums are finished by the consuming epilogue
        q = ops.gemm(x, self.attn1.to_q.packed_ln(dev, self.norm1), _new(x, M, C), rowstats=st)
        kv = ops.gemm(x, self.attn1.fused(dev, "kv"), _new(x, M, 2 * C))
        att = attend(q, kv)
        o1 = self.attn1.to_out["0"].packed(dev)
        sp = torch.empty(M, ops.stats_slots(o1), 2, dtype=torch.float32, device=dev)
        x = ops.gemm(att, o1, _new(x, M, C), res1=x, stats_out=sp)
        return self.ff.run(x, self.norm2, sp, out)


class SpatialTransformer(nn.Module):
    """attention.py:764-889 (2-D, depth 1, 1x1-conv projections). Input/outputs: [F, H, W, C]."""

    def __init__(self, in_channels, n_heads, d_head, context_dim=None, disable_text_ca=False, **_):
        super().__init__()
        inner = n_heads * d_head
        self.heads = n_heads
        self.channels = in_channels
        self.disable_text_ca = disable_text_ca
        self.norm = norm(in_channels)
        self.proj_in = conv2d(in_channels, inner, 1)
        blk = (BasicTransformerSingleLayerBlock(inner, n_heads, d_head) if disable_text_ca
               else BasicTransformerBlock(inner, n_heads, d_head, context_dim))
        self.transformer_blocks = nn.ModuleList([blk])
        self.proj_out = conv2d(inner, in_channels, 1, zero=True)

    def text_attns(self):
        return [] if self.disable_text_ca else [self.transformer_blocks[0].attn2]

    def run_spatial(self, x4, ctx: Ctx, out=None, dup=False):
        """x4: [F, H, W, C] (contiguous). Returns x + proj_out(block(proj_in(GN(x)))) as [F, H, W, C].
        dup (CFG de-duplication): x4 holds the F frames that the uncond and the cond half of the batch have in common
        (same latent, timestep and hint; only the text differs): GN, proj_in and the self-attention run once, the rows
        fan out in front of the text cross-attention and the result has 2 * F frames (uncond half first)."""
        dev = x4.device
        F, H, W, C = x4.shape
        L, M = H * W, F * H * W
        xn = ops.groupnorm_spatial(x4, *self.norm.affine(dev), GN_EPS_ATTN, False)
        pin = self.proj_in.packed(dev)
        sp = torch.empty(M, ops.stats_slots(pin), 2, dtype=torch.float32, device=dev)   # LayerNorm statistics of h
        h = ops.gemm(xn.view(M, C), pin, _new(x4, M, C), stats_out=sp)
        blk = self.transformer_blocks[0]
        if self.disable_text_ca:
            def attend(q, kv):
                kv3 = kv.view(F, L, 2 * C)
                return ops.attention(q.view(F, L, C), [KVSegment(kv3[..., :C], kv3[..., C:])], self.heads,
                                     _new(q, F, L, C)).view(M, C)
            if dup:
                raise RuntimeError("ccedit_b200: CFG de-duplication needs a text cross-attention block to fan out at")
            h = blk.run(h, attend, sp=sp)
        else:
            h = blk.run(h, F, L, ctx, sp=sp, dup=dup)
        if dup:                                                # residual x_in for both halves, written in place
            if out is not None:
                raise RuntimeError("ccedit_b200: de-duplicated transformer cannot write into a caller buffer")
            out = ops.dup_rows(x4)
            ops.gemm(h, self.proj_out.packed(dev), out.view(2 * M, C), res1=out.view(2 * M, C))
            return out
        out = _new(x4, F, H, W, C) if out is None else out
        ops.gemm(h, self.proj_out.packed(dev), out.view(M, C) if out.is_contiguous() else out.flatten(0, 2),
                 res1=x4.view(M, C))
        return out

    def run(self, x4, ctx: Ctx, out=None, dup=False):
        return self.run_spatial(x4, ctx, out, dup)


class SpatialTransformer3D(SpatialTransformer):
    """attention.py:1000-1208 (+ SpatialTransformer3DCA :1211-1350 when ca_type is set). [B, T, H, W, C] in/out."""

    def __init__(self, in_channels, n_heads, d_head, context_dim=None, ca_type: Optional[str] = None, **_):
        super().__init__(in_channels, n_heads, d_head, context_dim)
        inner = n_heads * d_head
        self.norm_temporal = norm(in_channels)
        self.proj_in_temporal = conv1d(in_channels, inner, 1, zero=True)
        self.transformer_blocks_temporal = nn.ModuleList([BasicTransformerSingleLayerBlock(inner, n_heads, d_head)])
        self.proj_out_temporal = conv1d(inner, in_channels, 1, zero=True)
        self.ca_type = ca_type
        if ca_type is not None:
            assert ca_type in ("center", "self", "center_self")
            self.norm_temporal_ca = norm(in_channels)
            self.proj_in_temporal_ca = conv2d(in_channels, inner, 1)
            self.transformer_blocks_temporal_ca = nn.ModuleList([BasicTransformerSingleLayerBlock(inner, n_heads, d_head)])
            self.proj_out_temporal_ca = conv2d(inner, in_channels, 1, zero=True)

    def run(self, x5, ctx: Ctx, out=None, dup=False):
        dev = x5.device
        B, T, H, W, C = x5.shape
        xs = self.run_spatial(x5.view(B * T, H, W, C), ctx, dup=dup)         # [F,H,W,C] (2 F frames after a fan-out)
        if dup:
            B *= 2
        F, L, M = B * T, H * W, B * T * H * W
        heads = self.heads
        # ---- temporal attention over T per pixel (attention.py:1172-1207) ----
        xt = ops.groupnorm_temporal(xs.view(B, T, L, C), *self.norm_temporal.affine(dev), GN_EPS_ATTN, False)
        pin = self.proj_in_temporal.packed(dev)
        sp = torch.empty(M, ops.stats_slots(pin), 2, dtype=torch.float32, device=dev)
        p = ops.gemm(xt.view(M, C), pin, _new(x5, M, C), stats_out=sp)

        def attend_t(q, kv):
            kv4 = kv.view(B, T, L, 2 * C)
            return ops.temporal_attention(q.view(B, T, L, C), kv4[..., :C], kv4[..., C:], heads,
                                          _new(q, B, T, L, C)).view(M, C)

        p = self.transformer_blocks_temporal[0].run(p, attend_t, sp=sp)
        last = self.ca_type is None
        dst = (_new(x5, M, C) if out is None else _as2d(out, M, C)) if last else _new(x5, M, C)
        ops.gemm(p, self.proj_out_temporal.packed(dev), dst, res1=xs.view(M, C))
        if last:
            return dst.view(B, T, H, W, C) if out is None else out
        # ---- cross-frame attention (SpatialTransformer3DCA.forward, attention.py:1302-1350) ----
        x2 = dst                                                               # [M, C] contiguous
        xc = ops.groupnorm_spatial(x2.view(F, L, C), *self.norm_temporal_ca.affine(dev), GN_EPS_ATTN, False)
        pin = self.proj_in_temporal_ca.packed(dev)
        sp = torch.empty(M, ops.stats_slots(pin), 2, dtype=torch.float32, device=dev)
        p = ops.gemm(xc.view(M, C), pin, _new(x5, M, C), stats_out=sp)

        def attend_ca(q, kv):
            kv3 = kv.view(F, L, 2 * C)
            k, v = kv3[..., :C], kv3[..., C:]
            center = KVSegment(k, v, div=T, mul=T, add=T // 2)
            own = KVSegment(k, v)
            segs = {"center": [center], "self": [own], "center_self": [center, own]}[self.ca_type]
            return ops.attention(q.view(F, L, C), segs, heads, _new(q, F, L, C)).view(M, C)

        p = self.transformer_blocks_temporal_ca[0].run(p, attend_ca, sp=sp)
        dst = _new(x5, M, C) if out is None else _as2d(out, M, C)
        ops.gemm(p, self.proj_out_temporal_ca.packed(dev), dst, res1=x2)
        return dst.view(B, T, H, W, C) if out is None else out


def _as2d(t: torch.Tensor, M: int, C: int) -> torch.Tensor:
    """[.., C] view with collapsible leading dims (contiguous, or a channel slice of a contiguous buffer) -> [M, C]."""
    if t.dim() == 2:
        return t
    return t.as_strided((M, C), (t.stride(-2), 1), t.storage_offset())


# ---------------------------------------------------------------------------------------------------------------------
# openaimodel.py mirrors
# ---------------------------------------------------------------------------------------------------------------------
class ResBlock(nn.Module):
    """openaimodel.py:397-554 (2-D; ControlNet2D). [F, H, W, C] in/out."""

    def __init__(self, channels, emb_channels, out_channels):
        super().__init__()
        self.channels, self.out_channels = channels, out_channels
        self.in_layers = seq(_0=norm(channels), _2=conv2d(channels, out_channels, 3))
        self.emb_layers = seq(_1=linear(emb_channels, out_channels))
        self.out_layers = seq(_0=norm(out_channels), _3=conv2d(out_channels, out_channels, 3, zero=True))
        if out_channels != channels:
            self.skip_connection = conv2d(channels, out_channels, 1)
        else:
            self.skip_connection = nn.Identity()

    def run(self, x4, ctx: Ctx, out=None):
        dev = x4.device
        F, H, W, Cin = x4.shape
        Co = self.out_channels
        taps = ops.conv_taps()
        a = ops.groupnorm_spatial(x4, *self.in_layers["0"].affine(dev), GN_EPS_RES, True)
        h = ops.gemm(a, self.in_layers["2"].packed(dev), _new(x4, F, H, W, Co), taps,
                     rowbias=ctx.emb_rows[id(self)], rb_dim=2, rb_div=ctx.T)
        a = ops.groupnorm_spatial(h, *self.out_layers["0"].affine(dev), GN_EPS_RES, True, out=a if Co == Cin else None)
        if isinstance(self.skip_connection, ParamHolder):
            skip = ops.gemm(x4, self.skip_connection.packed(dev), _new(x4, F, H, W, Co))
        else:
            skip = x4
        out = h if out is None else out   # h is dead after the GN above: reuse its buffer
        return ops.gemm(a, self.out_layers["3"].packed(dev), out, taps, res1=skip)


class ResBlock3D(ResBlock):
    """openaimodel.py:557-775 (pseudo-3D). [B, T, H, W, C] in/out; every spatial conv is followed by
    `y + conv1d_k3(SiLU(GN_t(y)))` along T (spatial_temporal_forward, :129-178)."""

    def __init__(self, channels, emb_channels, out_channels):
        super().__init__(channels, emb_channels, out_channels)
        Co = out_channels
        self.in_layers_temporal = seq(_0=norm(Co), _2=conv1d(Co, Co, 3, zero=True))
        self.out_layers_temporal = seq(_0=norm(Co), _3=conv1d(Co, Co, 3, zero=True))
        if out_channels != channels:
            self.skip_connection_temporal = conv1d(Co, Co, 1, zero=True)
        else:
            self.skip_connection_temporal = None

    def run(self, x5, ctx: Ctx, out=None):
        dev = x5.device
        B, T, H, W, Cin = x5.shape
        F, L, Co = B * T, H * W, self.out_channels
        taps, ttaps = ops.conv_taps(), ops.temporal_taps(3)
        x4 = x5.view(F, H, W, Cin)
        a = ops.groupnorm_spatial(x4, *self.in_layers["0"].affine(dev), GN_EPS_RES, True)
        y = ops.gemm(a, self.in_layers["2"].packed(dev), _new(x5, F, H, W, Co), taps)
        at = ops.groupnorm_temporal(y.view(B, T, L, Co), *self.in_layers_temporal["0"].affine(dev), GN_EPS_RES, True)
        h = ops.gemm(at, self.in_layers_temporal["2"].packed(dev), _new(x5, B, T, L, Co), ttaps, res1=y.view(B, T, L, Co),
                     rowbias=ctx.emb_rows[id(self)], rb_dim=2, rb_div=1)
        a = ops.groupnorm_spatial(h.view(F, L, Co), *self.out_layers["0"].affine(dev), GN_EPS_RES, True,
                                  out=at.view(F, L, Co))
        y = ops.gemm(a.view(F, H, W, Co), self.out_layers["3"].packed(dev), y, taps)
        at = ops.groupnorm_temporal(y.view(B, T, L, Co), *self.out_layers_temporal["0"].affine(dev), GN_EPS_RES, True,
                                    out=a.view(B, T, L, Co))
        if self.skip_connection_temporal is not None:
            s = ops.gemm(x4, self.skip_connection.packed(dev), h.view(F, H, W, Co))      # h is dead: reuse
            skip = ops.gemm(s.view(B, T, L, Co), self.skip_connection_temporal.packed(dev), _new(x5, B, T, L, Co),
                            res1=s.view(B, T, L, Co))
        else:
            skip = x5.view(B, T, L, Co)
        if out is None:
            dst = _new(x5, B, T, L, Co)
        else:
            dst = out.view(B, T, L, Co) if out.is_contiguous() else out.flatten(2, 3)
        ops.gemm(at, self.out_layers_temporal["3"].packed(dev), dst, ttaps, res1=y.view(B, T, L, Co), res2=skip)
        return dst.view(B, T, H, W, Co) if out is None else out


class Downsample(nn.Module):
    """openaimodel.py:282-322: conv 3x3, stride 2, pad 1. [F, H, W, C] -> [F, H/2, W/2, C]."""

    def __init__(self, channels):
        super().__init__()
        self.channels = channels
        self.op = conv2d(channels, channels, 3)

    def run(self, x4, ctx: Ctx, out=None):
        F, H, W, C = x4.shape
        planes = ops.parity_split(x4)
        out = _new(x4, F, H // 2, W // 2, C) if out is None else out
        ops.gemm(planes, self.op.packed(x4.device), out.unsqueeze(1), ops.conv_s2_taps())
        return out


class Downsample3D(Downsample):
    """openaimodel.py:325-394: y = conv_s2(x) per frame; out = y + conv1d_k3(y) along T."""

    def __init__(self, channels):
        super().__init__(channels)
        self.conv_temporal = conv1d(channels, channels, 3, zero=True)

    def run(self, x5, ctx: Ctx, out=None):
        B, T, H, W, C = x5.shape
        y = super().run(x5.view(B * T, H, W, C), ctx)
        L = (H // 2) * (W // 2)
        dst = _new(x5, B, T, L, C) if out is None else (out.view(B, T, L, C) if out.is_contiguous() else out.flatten(2, 3))
        ops.gemm(y.view(B, T, L, C), self.conv_temporal.packed(x5.device), dst, ops.temporal_taps(3),
                 res1=y.view(B, T, L, C))
        return dst.view(B, T, H // 2, W // 2, C) if out is None else out


class Upsample3D(nn.Module):
    """openaimodel.py:220-263: nearest x2 on (H, W), conv 3x3, then y + conv1d_k3(y) along T."""

    def __init__(self, channels):
        super().__init__()
        self.channels = channels
        self.conv = conv2d(channels, channels, 3)
        self.conv_temporal = conv1d(channels, channels, 3, zero=True)

    def run(self, x5, ctx: Ctx, out=None):
        B, T, H, W, C = x5.shape
        F, L = B * T, 4 * H * W
        u = ops.upsample_nearest2x(x5.view(F, H, W, C))
        y = ops.gemm(u, self.conv.packed(x5.device), _new(x5, F, 2 * H, 2 * W, C), ops.conv_taps())
        dst = _new(x5, B, T, L, C) if out is None else (out.view(B, T, L, C) if out.is_contiguous() else out.flatten(2, 3))
        ops.gemm(y.view(B, T, L, C), self.conv_temporal.packed(x5.device), dst, ops.temporal_taps(3),
                 res1=y.view(B, T, L, C))
        return dst.view(B, T, 2 * H, 2 * W, C) if out is None else out


class TimestepEmbedSequential(nn.ModuleList):
    """openaimodel.py:85-126: children are applied in order; here every child implements run(x, ctx, out)."""

    @property
    def out_channels(self) -> int:
        for layer in reversed(list(self)):
            for name in ("out_channels", "channels"):
                if hasattr(layer, name):
                    return getattr(layer, name)
        raise AttributeError("no layer with a channel count")

    @property
    def upsamples(self) -> bool:
        return any(isinstance(layer, Upsample3D) for layer in self)

    def run(self, x, ctx: Ctx, out=None, dup=False):
        """dup: x is the de-duplicated half of a CFG batch; the block's transformer fans it out (see
        SpatialTransformer.run_spatial), layers ahead of it run once."""
        n = len(self)
        if dup and not any(isinstance(layer, SpatialTransformer) for layer in self):
            raise RuntimeError("ccedit_b200: CFG de-duplication needs a transformer in the block")
        for i, layer in enumerate(self):
            if dup and isinstance(layer, SpatialTransformer):
                x = layer.run(x, ctx, out if i == n - 1 else None, dup=True)
                dup = False
            else:
                x = layer.run(x, ctx, out if i == n - 1 else None)
        return x

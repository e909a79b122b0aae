"""Thin tensor-plumbing layer over the C-ABI kernels (include/ccedit_b200.h).

Every function takes CUDA fp16 tensors in the canonical channels-last layout, derives raw pointers / strides and makes
exactly one C-ABI call on the current CUDA stream.  PyTorch is used for memory and streams only - there is no torch
compute and no fallback here: a missing library or a CPU tensor raises.
"""
from __future__ import annotations

import ctypes as C
import functools
import itertools
import math
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import AttnDesc, GemmDesc

BLOCK_M = 128
BLOCK_K = 64


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# per-launch profiler (bench.py --breakdown / roofline): when enabled every C-ABI call is bracketed by CUDA events on
# the launching stream and recorded with its algorithmic FLOPs and bytes
_PROF = None
_PROF_SHAPES = False     # developer switch: split the gemm classes by problem shape (tools/gemm_shapes.py)


def profile_start() -> None:
    global _PROF
    _PROF = []


def profile_stop(executed: bool = False):
    """-> list of (kernel class, flops, bytes, milliseconds[, executed flops]); synchronises the device.
    flops = useful FLOPs of the launch; executed flops include what the tensor cores spend on padding (K padded to 64,
    head dim padded to a multiple of 16, partial row / key tiles)."""
    global _PROF
    recs, _PROF = _PROF or [], None
    torch.cuda.synchronize()
    if executed:
        return [(name, fl, by, e0.elapsed_time(e1), xf) for name, fl, by, e0, e1, xf in recs]
    return [(name, fl, by, e0.elapsed_time(e1)) for name, fl, by, e0, e1, _ in recs]


def _call(name: str, fn, args, flops: float = 0.0, nbytes: float = 0.0, xflops: Optional[float] = None) -> None:
    if _PROF is None:
        _lib.check(fn(*args), name)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(fn(*args), name)
    e1.record()
    _PROF.append((name, flops, nbytes, e0, e1, flops if xflops is None else xflops))


def _ceil_to(x: int, m: int) -> int:
    return -(-x // m) * m


def _nb(*ts) -> int:
    return sum(t.numel() * t.element_size() for t in ts if t is not None)


def _require(t: torch.Tensor, dtype=torch.float16, name="tensor"):
    if not t.is_cuda:
        raise RuntimeError(f"ccedit_b200: {name} must be a CUDA tensor (no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"ccedit_b200: {name} must be {dtype}, got {t.dtype}")


# ---------------------------------------------------------------------------------------------------------------------
# weights
# ---------------------------------------------------------------------------------------------------------------------
def pick_bn(n: int, geglu: bool = False, max_bn: int = 256) -> int:
    """N tile (multiple of 16, <= max_bn <= 256; multiple of 32 for GEGLU) that divides n: the largest one, unless a tile
    at least 3/4 as wide has an even number of 16-column chunks - the staged (TMA-store) epilogue of the tap-GEMM splits
    the tile's columns between two warpgroups (e.g. n = 960: 192 instead of 240)."""
    step = 32 if geglu else 16
    cands = [bn for bn in range(max_bn - max_bn % step, 0, -step) if n % bn == 0]
    if not cands:
        raise ValueError(f"output width {n} is not a multiple of {step}")
    best = cands[0]
    if not geglu and best % 32 != 0:
        for bn in cands[1:]:
            if bn % 32 == 0 and 4 * bn >= 3 * best:
                return bn
    return best


@dataclass
class PackedWeight:
    """fp16 weight in the tap-GEMM layout [N][ntaps][kpad] (+ fp32 bias)."""

    w: torch.Tensor
    bias: Optional[torch.Tensor]
    n: int
    k: int
    kpad: int
    ntaps: int
    bn: int
    geglu: bool = False
    colsum: Optional[torch.Tensor] = None   # fp32 [n]: row sums of the fp16 weight, present when a LayerNorm is folded in

    @property
    def n_out(self) -> int:
        return self.n // 2 if self.geglu else self.n


def pack_weight(w: torch.Tensor, bias: Optional[torch.Tensor], device, geglu: bool = False, n_pad_to: int = 16,
                cin_pad_to: int = 8, max_bn: int = 256, ln_gamma: Optional[torch.Tensor] = None,
                ln_beta: Optional[torch.Tensor] = None) -> PackedWeight:
    """w: [N, Cin, *taps] in the reference (PyTorch) layout: Linear [N,K], Conv2d [N,Cin,kh,kw], Conv1d [N,Cin,kt].

    Tap order is row-major over the kernel dims (kh*3+kw / kt), matching ``conv_taps``/``temporal_taps``.
    ``ln_gamma``/``ln_beta`` fold the LayerNorm that feeds this Linear into it (attention.py:667-669 -> to_q/ff.net.0):
    LN(x) W^T + b = rstd * (x (gamma o W)^T - mean * colsum(gamma o W)) + (b + W beta); ``gemm(..., rowstats=)`` applies
    the per-row part, the packed weight carries gamma o W, its fp16 row sums and the shifted bias.
    GEGLU: rows are re-ordered so that every BN tile holds BN/2 value rows followed by their BN/2 gate rows
    (reference: value = first half of the 2*inner outputs, gate = second half; attention.py:120-122).
    """
    w = w.detach().to(torch.float32)
    n, cin = w.shape[0], w.shape[1]
    ntaps = int(math.prod(w.shape[2:])) if w.dim() > 2 else 1
    fold = ln_gamma is not None
    if fold:
        if w.dim() != 2:
            raise ValueError("LayerNorm folding applies to Linear weights only")
        g32 = ln_gamma.detach().to(device=w.device, dtype=torch.float32)
        b32 = ln_beta.detach().to(device=w.device, dtype=torch.float32)
        shift = w @ b32
        bias = shift if bias is None else bias.detach().to(device=w.device, dtype=torch.float32) + shift
        w = w * g32[None, :]
    w = w.reshape(n, cin, ntaps).permute(0, 2, 1)  # [N, taps, Cin]
    b = None if bias is None else bias.detach().to(torch.float32)
    n_p = ((n + n_pad_to - 1) // n_pad_to) * n_pad_to
    if n_p != n:
        w = torch.cat([w, w.new_zeros(n_p - n, ntaps, cin)], 0)
        if b is not None:
            b = torch.cat([b, b.new_zeros(n_p - n)], 0)
    bn = pick_bn(n_p, geglu, max_bn)
    if geglu:
        half = n_p // 2
        hb = bn // 2
        idx = []
        for j in range(n_p // bn):
            idx += list(range(j * hb, (j + 1) * hb)) + list(range(half + j * hb, half + (j + 1) * hb))
        idx = torch.tensor(idx, dtype=torch.long)
        w = w[idx]
        if b is not None:
            b = b[idx]
    kpad = ((cin + BLOCK_K - 1) // BLOCK_K) * BLOCK_K
    wp = w.new_zeros(n_p, ntaps, kpad)
    wp[:, :, :cin] = w
    wp = wp.reshape(n_p, ntaps * kpad).to(device=device, dtype=torch.float16).contiguous()
    bp = None if b is None else b.to(device=device, dtype=torch.float32).contiguous()
    cs = wp.float().sum(1).contiguous() if fold else None     # sums of the ROUNDED weights: the mean term cancels exactly
    return PackedWeight(wp, bp, n_p, cin, kpad, ntaps, bn, geglu, cs)


def conv_taps():
    """3x3, stride 1, pad 1: tap (kh,kw) reads (h+kh-1, w+kw-1); offsets along (d1=W, d2=H, d3, d4)."""
    return [(kw - 1, kh - 1, 0, 0) for kh in range(3) for kw in range(3)]


def conv_s2_taps():
    """3x3, stride 2, pad 1 over the 4 parity planes (d3 = plane = (h&1)*2 + (w&1)) of ``parity_split``."""
    def split(k):  # input index 2*o + k - 1  ->  (parity, offset)
        return (1, -1) if k == 0 else ((0, 0) if k == 1 else (1, 0))
    taps = []
    for kh in range(3):
        ph, oh = split(kh)
        for kw in range(3):
            pw, ow = split(kw)
            taps.append((ow, oh, ph * 2 + pw, 0))
    return taps


def temporal_taps(k: int = 3):
    """Conv1d over T (d2), zero padded: tap kt reads t + kt - k//2."""
    return [(0, kt - k // 2, 0, 0) for kt in range(k)]


ONE_TAP = [(0, 0, 0, 0)]


@functools.lru_cache(maxsize=None)
def pick_box(dims: tuple) -> tuple:
    """Tile extents (powers of two, product 128) along d1..d4 minimising padded rows; ties -> larger inner extents."""
    best, best_cost = None, None
    for e in itertools.product(range(8), repeat=4):
        if sum(e) != 7:
            continue
        box = tuple(1 << x for x in e)
        cost = 1
        for o, b in zip(dims, box):
            cost *= -(-o // b)
        key = (cost, -box[0], -box[1], -box[2])
        if best_cost is None or key < best_cost:
            best, best_cost = box, key
    return best


def _dims_strides(t: torch.Tensor):
    """Tensor [d4, d3, d2, d1, C] (left-padded with 1s) -> (C, (d1..d4), (s1..s4))."""
    if t.stride(-1) != 1:
        raise RuntimeError("ccedit_b200: channel dimension must be contiguous")
    shape, stride = list(t.shape), list(t.stride())
    if len(shape) > 5:
        raise RuntimeError("ccedit_b200: at most 5 dims")
    while len(shape) < 5:
        stride.insert(0, stride[0] * shape[0])
        shape.insert(0, 1)
    # size-1 dims may carry arbitrary strides; give them the dense value so that TMA's 16-byte rule holds
    for i in range(3, -1, -1):
        if shape[i] == 1:
            stride[i] = stride[i + 1] * shape[i + 1]
    dims = shape[-2::-1]
    strides = stride[-2::-1]
    return shape[-1], dims, strides


def gemm(a: torch.Tensor, pw: PackedWeight, out: torch.Tensor, taps: Sequence = ONE_TAP, *,
         rowbias: Optional[torch.Tensor] = None, rb_dim: int = 0, rb_div: int = 1,
         res1: Optional[torch.Tensor] = None, res2: Optional[torch.Tensor] = None, silu: bool = False,
         box: Optional[tuple] = None, rowstats: Optional[torch.Tensor] = None,
         stats_out: Optional[torch.Tensor] = None, ln_eps: float = 1e-5) -> torch.Tensor:
    """out[p, :] = epi( sum_tap A[p + tap, :] @ W[:, tap, :]^T ).  a/out/res*: [d4, d3, d2, d1, C] views (C contiguous)."""
    _require(a, name="a")
    _require(out, name="out")
    c, adims, astr = _dims_strides(a)
    n_out, odims, ostr = _dims_strides(out)
    if n_out != pw.n_out:
        raise RuntimeError(f"ccedit_b200.gemm: out has {n_out} channels, weight produces {pw.n_out}")
    if len(taps) != pw.ntaps:
        raise RuntimeError(f"ccedit_b200.gemm: {len(taps)} taps given, weight packed for {pw.ntaps}")
    if c != pw.k and not (c <= pw.kpad and pw.k <= c):
        raise RuntimeError(f"ccedit_b200.gemm: a has {c} channels, weight expects {pw.k}")
    d = GemmDesc()
    d.a = a.data_ptr()
    d.a_dims[0] = c
    for i in range(4):
        d.a_dims[i + 1] = adims[i]
        d.a_strides[i] = astr[i]
        d.out_dims[i] = odims[i]
        d.out_strides[i] = ostr[i]
    bx = box or pick_box(tuple(odims))
    for i in range(4):
        d.box[i] = bx[i]
    d.ntaps = len(taps)
    for t, tap in enumerate(taps):
        for i in range(4):
            d.taps[t][i] = tap[i]
    d.w = pw.w.data_ptr()
    d.n = pw.n
    d.kpad = pw.kpad
    d.bn = pw.bn
    d.out = out.data_ptr()
    d.bias = _ptr(pw.bias)
    if rowbias is not None:
        _require(rowbias, torch.float32, "rowbias")
        if rowbias.dim() != 2 or rowbias.shape[-1] != n_out or rowbias.stride(1) != 1:
            raise RuntimeError("ccedit_b200.gemm: rowbias must be [R, n_out] with contiguous rows")
        d.rowbias = rowbias.data_ptr()
        d.rb_ld = rowbias.stride(0)
        d.rb_dim = rb_dim
        d.rb_div = rb_div
    for name, r in (("res1", res1), ("res2", res2)):
        if r is None:
            continue
        _require(r, name=name)
        rc, rdims, rstr = _dims_strides(r)
        if rc != n_out or list(rdims) != list(odims):
            raise RuntimeError(f"ccedit_b200.gemm: {name} shape {tuple(r.shape)} does not match out {tuple(out.shape)}")
        setattr(d, name, r.data_ptr())
        arr = getattr(d, name + "_strides")
        for i in range(4):
            arr[i] = rstr[i]
    d.flags = (_lib.GEMM_SILU if silu else 0) | (_lib.GEMM_GEGLU if pw.geglu else 0)
    if (rowstats is not None) != (pw.colsum is not None):
        raise RuntimeError("ccedit_b200.gemm: rowstats go with a LayerNorm-folded weight (pack_weight(ln_gamma=...)) and vice versa")
    if rowstats is not None:
        _require(rowstats, torch.float32, "rowstats")
        if not rowstats.is_contiguous() or rowstats.shape[0] != odims[0] or rowstats.shape[-1] != 2 or rowstats.dim() not in (2, 3):
            raise RuntimeError(f"ccedit_b200.gemm: rowstats must be contiguous [{odims[0]}, 2] (mean, rstd) or "
                               f"[{odims[0]}, P, 2] partial sums from gemm(stats_out=)")
        d.rowstats = rowstats.data_ptr()
        d.colsum = pw.colsum.data_ptr()
        d.rowstats_slots = rowstats.shape[1] if rowstats.dim() == 3 else 0
        d.ln_eps = ln_eps
    if stats_out is not None:
        _require(stats_out, torch.float32, "stats_out")
        if tuple(stats_out.shape) != (odims[0], stats_slots(pw), 2) or not stats_out.is_contiguous():
            raise RuntimeError(f"ccedit_b200.gemm: stats_out must be contiguous [{odims[0]}, {stats_slots(pw)}, 2]")
        d.stats_out = stats_out.data_ptr()
    m_rows = math.prod(odims)
    kind = {1: "gemm.linear", 3: "gemm.temporal_k3", 9: "gemm.conv3x3"}.get(len(taps), "gemm.other")
    if pw.geglu:
        kind = "gemm.linear_geglu"
    if _PROF_SHAPES:
        kind += f"[M={m_rows},K={pw.ntaps}x{min(c, pw.k)},N={pw.n},bn={pw.bn},res={int(res1 is not None) + int(res2 is not None)}]"
    m_exec = math.prod(-(-o // b) * b for o, b in zip(odims, bx))
    _call(kind, _lib.load().ccedit_gemm, (C.byref(d), _stream()), flops=2.0 * m_rows * pw.ntaps * min(c, pw.k) * pw.n,
          nbytes=_nb(a, pw.w, out, res1, res2), xflops=2.0 * m_exec * pw.ntaps * pw.kpad * pw.n)
    return out


# ---------------------------------------------------------------------------------------------------------------------
# normalisation
# ---------------------------------------------------------------------------------------------------------------------
_GN_SCRATCH = {}


def _gn_scratch(device, F: int) -> torch.Tensor:
    key = (device, torch.cuda.current_stream().cuda_stream)
    need = 8192 + F * 32 * 64             # the single-pass kernel's arrival counters (fixed 8192 slots) + partial sums [F][32][32][2]
    buf = _GN_SCRATCH.get(key)
    if buf is None or buf.numel() < need:
        buf = torch.zeros(need, dtype=torch.float32, device=device)   # the counters must start at zero (self re-arming)
        _GN_SCRATCH[key] = buf
    return buf


def groupnorm_spatial(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, silu: bool,
                      out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x: contiguous [F, HW, C] (or [F, H, W, C]); GroupNorm(32, C) per frame (+ SiLU)."""
    _require(x, name="x")
    if not x.is_contiguous():
        raise RuntimeError("ccedit_b200.groupnorm_spatial: x must be contiguous")
    F, Cc = x.shape[0], x.shape[-1]
    HW = x.numel() // (F * Cc)
    out = torch.empty_like(x) if out is None else out
    _call("groupnorm_spatial", _lib.load().ccedit_groupnorm_spatial,
          (x.data_ptr(), out.data_ptr(), gamma.data_ptr(), beta.data_ptr(), _gn_scratch(x.device, F).data_ptr(), F, HW, Cc,
           eps, int(silu), _stream()), nbytes=_nb(x, out))
    return out


def groupnorm_temporal(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, silu: bool,
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x: contiguous [B, T, HW, C] (or [B, T, H, W, C]); GroupNorm(32, C) over (C/32, T) per pixel (+ SiLU)."""
    _require(x, name="x")
    if not x.is_contiguous():
        raise RuntimeError("ccedit_b200.groupnorm_temporal: x must be contiguous")
    B, T, Cc = x.shape[0], x.shape[1], x.shape[-1]
    HW = x.numel() // (B * T * Cc)
    out = torch.empty_like(x) if out is None else out
    _call("groupnorm_temporal", _lib.load().ccedit_groupnorm_temporal,
          (x.data_ptr(), out.data_ptr(), gamma.data_ptr(), beta.data_ptr(), B, T, HW, Cc, eps, int(silu), _stream()),
          nbytes=_nb(x, out))
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x: [..., C] with a uniform row stride (contiguous or a channel slice); per-row LayerNorm -> contiguous."""
    _require(x, name="x")
    Cc = x.shape[-1]
    M = x.numel() // Cc
    if x.dim() == 2:
        x2 = x
    elif x.is_contiguous():
        x2 = x.view(M, Cc)
    else:
        raise RuntimeError("ccedit_b200.layernorm: x must be 2-D (row stride free) or contiguous")
    if x2.stride(1) != 1:
        raise RuntimeError("ccedit_b200.layernorm: channel dimension must be contiguous")
    ldx = x2.stride(0)
    out = torch.empty(x.shape, dtype=torch.float16, device=x.device) if out is None else out
    _call("layernorm" + (f"[M={M},C={Cc}]" if _PROF_SHAPES else ""), _lib.load().ccedit_layernorm,
          (x2.data_ptr(), ldx, out.data_ptr(), gamma.data_ptr(), beta.data_ptr(), M, Cc, eps, _stream()), nbytes=2 * _nb(out))
    return out


def stats_slots(pw: PackedWeight) -> int:
    """Partial-statistics slots per row a GEMM with this weight writes (``gemm(stats_out=)``): 2 per N tile."""
    return 2 * (pw.n // pw.bn)


def layernorm_stats_combine(partial: torch.Tensor, C: int, eps: float = 1e-5, out: Optional[torch.Tensor] = None
                            ) -> torch.Tensor:
    """partial: fp32 [M, P, 2] (sum, sum of squares) from ``gemm(stats_out=)`` over C channels -> fp32 [M, 2] (mean, rstd)."""
    _require(partial, torch.float32, "partial")
    M, P, _ = partial.shape
    out = torch.empty(M, 2, dtype=torch.float32, device=partial.device) if out is None else out
    _call("layernorm_stats_combine", _lib.load().ccedit_layernorm_stats_combine,
          (partial.data_ptr(), P, out.data_ptr(), M, C, eps, _stream()), nbytes=_nb(partial, out))
    return out


def layernorm_stats(x: torch.Tensor, eps: float = 1e-5, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x: [M, C] (row stride free) -> fp32 [M, 2] = (mean, 1/sqrt(var + eps)) per row, for a LayerNorm folded into the
    consuming GEMM (``pack_weight(ln_gamma=...)`` + ``gemm(rowstats=...)``): one read of x instead of read + write + read."""
    _require(x, name="x")
    Cc = x.shape[-1]
    x2 = x.reshape(-1, Cc) if x.is_contiguous() else x
    if x2.dim() != 2 or x2.stride(1) != 1:
        raise RuntimeError("ccedit_b200.layernorm_stats: x must be [M, C] with contiguous channels")
    M = x2.shape[0]
    out = torch.empty(M, 2, dtype=torch.float32, device=x.device) if out is None else out
    _call("layernorm_stats" + (f"[M={M},C={Cc}]" if _PROF_SHAPES else ""), _lib.load().ccedit_layernorm_stats,
          (x2.data_ptr(), x2.stride(0), out.data_ptr(), M, Cc, eps, _stream()), nbytes=_nb(x2) + out.numel() * 4)
    return out


# ---------------------------------------------------------------------------------------------------------------------
# attention
# ---------------------------------------------------------------------------------------------------------------------
@dataclass
class KVSegment:
    """k, v: [Fkv, Lkv, heads*d] views (row stride free); query frame f reads kv frame (f // div) * mul + add."""

    k: torch.Tensor
    v: torch.Tensor
    div: int = 1
    mul: int = 1
    add: int = 0


def attention(q: torch.Tensor, segments: Sequence[KVSegment], heads: int, out: torch.Tensor,
              scale: Optional[float] = None) -> torch.Tensor:
    """q, out: [F, L, heads*d] views.  softmax(q k^T * scale) v over the concatenation of the key segments."""
    _require(q, name="q")
    _require(out, name="out")
    F, L, Cc = q.shape
    dh = Cc // heads
    a = AttnDesc()
    a.q, a.ldq, a.q_frame_stride = q.data_ptr(), q.stride(1), q.stride(0)
    a.o, a.ldo, a.o_frame_stride = out.data_ptr(), out.stride(1), out.stride(0)
    a.nseg = len(segments)
    for s, seg in enumerate(segments):
        _require(seg.k, name="k")
        _require(seg.v, name="v")
        a.k[s], a.v[s] = seg.k.data_ptr(), seg.v.data_ptr()
        a.ldk[s], a.ldv[s] = seg.k.stride(1), seg.v.stride(1)
        if seg.k.stride(0) != seg.v.stride(0):
            raise RuntimeError("ccedit_b200.attention: k and v must share the frame stride")
        a.kv_frame_stride[s] = seg.k.stride(0)
        a.lkv[s] = seg.k.shape[1]
        a.kv_div[s], a.kv_mul[s], a.kv_add[s] = seg.div, seg.mul, seg.add
    a.frames, a.lq, a.heads, a.d = F, L, heads, dh
    a.scale = float(dh) ** -0.5 if scale is None else scale
    lkv = sum(seg.k.shape[1] for seg in segments)
    kvb = sum(2.0 * seg.k.shape[0] * seg.k.shape[1] * Cc * 2 for seg in segments)   # each K/V row read once (ideal)
    _call("attention" + (f"[F={F},L={L},Lkv={lkv},d={dh}]" if _PROF_SHAPES else ""), _lib.load().ccedit_attention,
          (C.byref(a), _stream()), flops=4.0 * F * L * lkv * Cc,
          nbytes=2.0 * F * L * Cc * 2 + kvb, xflops=4.0 * F * _ceil_to(L, 128) * heads * _ceil_to(dh, 16) *
          sum(_ceil_to(seg.k.shape[1], 128 if dh <= 64 else 64) for seg in segments))
    return out


def temporal_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, out: torch.Tensor,
                       scale: Optional[float] = None) -> torch.Tensor:
    """q, k, v, out: [B, T, HW, heads*d] views with uniform row stride; attention over T for every (b, pixel, head)."""
    for t, nm in ((q, "q"), (k, "k"), (v, "v"), (out, "out")):
        _require(t, name=nm)
    B, T, HW, Cc = q.shape
    dh = Cc // heads
    sc = float(dh) ** -0.5 if scale is None else scale
    _call("temporal_attention", _lib.load().ccedit_temporal_attention,
          (q.data_ptr(), q.stride(2), k.data_ptr(), k.stride(2), v.data_ptr(), v.stride(2), out.data_ptr(), out.stride(2),
           B, T, HW, heads, dh, sc, _stream()), flops=4.0 * B * HW * T * T * Cc, nbytes=4.0 * B * T * HW * Cc * 2,
          xflops=4.0 * B * HW * _ceil_to(T, 16) ** 2 * heads * _ceil_to(dh, 16))
    return out


# ---------------------------------------------------------------------------------------------------------------------
# small kernels
# ---------------------------------------------------------------------------------------------------------------------
def ncthw_to_cl(src: torch.Tensor, cpad: int, mul: float = 1.0, add: float = 0.0, pre: float = 0.0) -> torch.Tensor:
    """[B, C, T, H, W] fp32/fp16 -> [B, T, H, W, cpad] fp16 ((v + pre) * mul + add, zero padded channels)."""
    if not src.is_cuda:
        raise RuntimeError("ccedit_b200: src must be a CUDA tensor (no CPU fallback)")
    if src.dtype not in (torch.float32, torch.float16):
        src = src.float()
    src = src.contiguous()
    B, Cin, T, H, W = src.shape
    dst = torch.empty(B, T, H, W, cpad, dtype=torch.float16, device=src.device)
    _call("ncthw_to_cl", _lib.load().ccedit_ncthw_to_cl,
          (src.data_ptr(), int(src.dtype == torch.float32), dst.data_ptr(), B, Cin, T, H, W, cpad, pre, mul, add, _stream()),
          nbytes=_nb(src, dst))
    return dst


def out_temporal(y: torch.Tensor, wt: torch.Tensor, bias_t: torch.Tensor, cout: int, out_dtype) -> torch.Tensor:
    """y: [B, T, H, W, ld] fp16 -> [B, cout, T, H, W] (y + conv1d_k3(silu(y)) over T)."""
    _require(y, name="y")
    B, T, H, W, ld = y.shape
    dst = torch.empty(B, cout, T, H, W, dtype=out_dtype, device=y.device)
    _call("out_temporal", _lib.load().ccedit_out_temporal,
          (y.data_ptr(), ld, wt.data_ptr(), bias_t.data_ptr(), dst.data_ptr(), int(out_dtype == torch.float32), B, cout, T,
           H * W, _stream()), nbytes=_nb(y, dst))
    return dst


def timestep_embedding(t: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    tf = t.to(torch.float32).contiguous()
    out = torch.empty(tf.shape[0], dim, dtype=torch.float32, device=tf.device)
    _call("timestep_embedding", _lib.load().ccedit_timestep_embedding,
          (tf.data_ptr(), out.data_ptr(), tf.shape[0], dim, max_period, _stream()), nbytes=_nb(out))
    return out


def linear_small(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], act_in: bool = False,
                 act_out: bool = False) -> torch.Tensor:
    """fp32 x [M<=8, K] @ fp16 w [N, K]^T + b -> fp32 [M, N]; optional SiLU on the input and/or output."""
    _require(x, torch.float32, "x")
    _require(w, torch.float16, "w")
    M, K = x.shape
    N = w.shape[0]
    out = torch.empty(M, N, dtype=torch.float32, device=x.device)
    _call("linear_small", _lib.load().ccedit_linear_small,
          (x.data_ptr(), w.data_ptr(), _ptr(b), out.data_ptr(), M, N, K, int(act_in), int(act_out), _stream()),
          flops=2.0 * M * N * K, nbytes=_nb(x, w, out))
    return out


def parity_split(x: torch.Tensor) -> torch.Tensor:
    """[F, H, W, C] -> [F, 4, H/2, W/2, C]."""
    _require(x, name="x")
    F, H, W, Cc = x.shape
    y = torch.empty(F, 4, H // 2, W // 2, Cc, dtype=torch.float16, device=x.device)
    _call("parity_split", _lib.load().ccedit_parity_split, (x.data_ptr(), y.data_ptr(), F, H, W, Cc, _stream()),
          nbytes=_nb(x, y))
    return y


def upsample_nearest2x(x: torch.Tensor) -> torch.Tensor:
    """[F, H, W, C] -> [F, 2H, 2W, C]."""
    _require(x, name="x")
    F, H, W, Cc = x.shape
    y = torch.empty(F, 2 * H, 2 * W, Cc, dtype=torch.float16, device=x.device)
    _call("upsample_nearest2x", _lib.load().ccedit_upsample_nearest2x, (x.data_ptr(), y.data_ptr(), F, H, W, Cc, _stream()),
          nbytes=_nb(x, y))
    return y


def add_rows(a: torch.Tensor, b: Optional[torch.Tensor], dst: torch.Tensor) -> torch.Tensor:
    """dst[..., :] = a + b for [M, C] views with uniform row strides (dst may be a channel slice of a wider buffer)."""
    Cc = a.shape[-1]
    M = a.numel() // Cc
    a2, d2 = a.reshape(M, Cc) if a.is_contiguous() else a.view(M, Cc), dst.view(M, Cc)
    b2 = None if b is None else (b.reshape(M, Cc) if b.is_contiguous() else b.view(M, Cc))
    _call("add_rows", _lib.load().ccedit_add_rows,
          (a2.data_ptr(), a2.stride(0), _ptr(b2), 0 if b2 is None else b2.stride(0), d2.data_ptr(), d2.stride(0), M, Cc,
           _stream()), nbytes=_nb(a, b) + M * Cc * 2)
    return dst


def dup_rows(t: torch.Tensor) -> torch.Tensor:
    """Contiguous tensor [d0, ...] -> [2 * d0, ...] holding t twice (raw 16-byte copies; any dtype).  Used by the CFG
    de-duplication: layers whose inputs are identical in the uncond and cond halves run once and fan out here."""
    if not t.is_cuda or not t.is_contiguous():
        raise RuntimeError("ccedit_b200.dup_rows: t must be a contiguous CUDA tensor")
    dst = torch.empty((2 * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    nbytes = t.numel() * t.element_size()
    if nbytes % 16:
        raise RuntimeError("ccedit_b200.dup_rows: size must be a multiple of 16 bytes")
    n16 = nbytes // 2                                   # size in fp16 elements
    C = next(c for c in (4096, 512, 64, 8) if n16 % c == 0)
    M = n16 // C
    for half in range(2):
        _call("dup_rows", _lib.load().ccedit_add_rows,
              (t.data_ptr(), C, None, 0, dst.data_ptr() + half * nbytes, C, M, C, _stream()), nbytes=2 * nbytes)
    return dst


def add_center_frame(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """x: [B, T, H, W, C] contiguous; y: [B, H, W, C]; x[:, T//2] += y in place."""
    _require(x, name="x")
    _require(y, name="y")
    B, T = x.shape[0], x.shape[1]
    Cc = x.shape[-1]
    HW = x.numel() // (B * T * Cc)
    _call("add_center_frame", _lib.load().ccedit_add_center_frame, (x.data_ptr(), y.data_ptr(), B, T, HW, Cc, _stream()),
          nbytes=3 * _nb(y))
    return x


def to_half(src: torch.Tensor) -> torch.Tensor:
    """fp32 (or already fp16) tensor -> contiguous fp16 tensor of the same shape."""
    if not src.is_cuda:
        raise RuntimeError("ccedit_b200: src must be a CUDA tensor (no CPU fallback)")
    if src.dtype == torch.float16:
        return src.contiguous()
    src = src.to(torch.float32).contiguous()
    dst = torch.empty(src.shape, dtype=torch.float16, device=src.device)
    _call("to_half", _lib.load().ccedit_to_half, (src.data_ptr(), dst.data_ptr(), src.numel(), _stream()),
          nbytes=_nb(src, dst))
    return dst


def pack_hint_stem_weight(w: torch.Tensor, bias: torch.Tensor, device, cin_pad: int, kpad: int):
    """Conv2d weight [N, Cin, 3, 3] (N = 16 or 32) -> fp16 [N][kpad] with k = tap*cin_pad + channel (tap = kh*3 + kw), fp32 bias."""
    n, cin = w.shape[0], w.shape[1]
    if n not in (16, 32) or cin > cin_pad or 9 * cin_pad > kpad:
        raise RuntimeError(f"ccedit_b200.pack_hint_stem_weight: unsupported conv shape {tuple(w.shape)}")
    wp = torch.zeros(n, 9, cin_pad, dtype=torch.float32)
    wp[:, :, :cin] = w.detach().float().cpu().reshape(n, cin, 9).permute(0, 2, 1)
    out = torch.zeros(n, kpad, dtype=torch.float32)
    out[:, :9 * cin_pad] = wp.reshape(n, 9 * cin_pad)
    return (out.to(device=device, dtype=torch.float16).contiguous(),
            bias.detach().to(device=device, dtype=torch.float32).contiguous())


def hint_stem01(x: torch.Tensor, w0: torch.Tensor, b0: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor) -> torch.Tensor:
    """x: [F, H, W, 8] fp16 -> SiLU(conv3x3(SiLU(conv3x3(x)))) [F, H, W, 16]: the two full-resolution layers of the
    ControlNet hint stem in one pass (weights from ``pack_hint_stem_weight`` with (cin_pad, kpad) = (8, 80), (16, 144))."""
    _require(x, name="x")
    if not x.is_contiguous() or x.shape[-1] != 8:
        raise RuntimeError("ccedit_b200.hint_stem01: x must be contiguous [F, H, W, 8]")
    F, H, W, _ = x.shape
    y = torch.empty(F, H, W, 16, dtype=torch.float16, device=x.device)
    _call("hint_stem01", _lib.load().ccedit_hint_stem01,
          (x.data_ptr(), y.data_ptr(), w0.data_ptr(), b0.data_ptr(), w1.data_ptr(), b1.data_ptr(), F, H, W, _stream()),
          flops=2.0 * F * H * W * 16 * 9 * (3 + 16), nbytes=_nb(x, y))
    return y


def hint_stem23(x: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor, w3: torch.Tensor, b3: torch.Tensor) -> torch.Tensor:
    """x: [F, H, W, 16] fp16 (H, W even) -> SiLU(conv3x3(SiLU(conv3x3_stride2(x)))) [F, H/2, W/2, 32]: layers 2 and 3 of the
    ControlNet hint stem in one pass (weights from ``pack_hint_stem_weight`` with (cin_pad, kpad) = (16, 144), (32, 288))."""
    _require(x, name="x")
    if not x.is_contiguous() or x.shape[-1] != 16 or x.shape[1] % 2 or x.shape[2] % 2:
        raise RuntimeError("ccedit_b200.hint_stem23: x must be contiguous [F, H, W, 16] with even H and W")
    F, H, W, _ = x.shape
    y = torch.empty(F, H // 2, W // 2, 32, dtype=torch.float16, device=x.device)
    _call("hint_stem23", _lib.load().ccedit_hint_stem23,
          (x.data_ptr(), y.data_ptr(), w2.data_ptr(), b2.data_ptr(), w3.data_ptr(), b3.data_ptr(), F, H, W, _stream()),
          flops=2.0 * F * (H // 2) * (W // 2) * 32 * 9 * (16 + 32), nbytes=_nb(x, y))
    return y


# ---------------------------------------------------------------------------------------------------------------------
# launch accounting (bench.py's gpu_launches): kernels launched by the library + kernels replayed from CUDA graphs
# ---------------------------------------------------------------------------------------------------------------------
_GRAPH_LAUNCHES = 0


def note_graph_replay(n: int) -> None:
    global _GRAPH_LAUNCHES
    _GRAPH_LAUNCHES += int(n)


def launch_count() -> int:
    """Kernels of this library executed so far in this process (direct launches + graph-replayed launches)."""
    return _lib.launch_count() + _GRAPH_LAUNCHES


# ---------------------------------------------------------------------------------------------------------------------
# first-stage (VAE) helpers
# ---------------------------------------------------------------------------------------------------------------------
def softmax_rows(x: torch.Tensor) -> torch.Tensor:
    """x: fp16 [M, N] (row stride free, N contiguous): row softmax in place, fp32 arithmetic."""
    _require(x, name="x")
    if x.dim() != 2 or x.stride(1) != 1:
        raise RuntimeError("ccedit_b200.softmax_rows: x must be [M, N] with contiguous columns")
    M, N = x.shape
    _call("softmax_rows", _lib.load().ccedit_softmax_rows, (x.data_ptr(), x.stride(0), M, N, _stream()), nbytes=2 * _nb(x))
    return x


def cl_to_ncthw(y: torch.Tensor, cout: int, out_dtype=torch.float32) -> torch.Tensor:
    """y: channels-last fp16 [B, T, H, W, ld] -> [B, cout, T, H, W] (first cout channels)."""
    _require(y, name="y")
    if not y.is_contiguous():
        raise RuntimeError("ccedit_b200.cl_to_ncthw: y must be contiguous")
    B, T, H, W, ld = y.shape
    dst = torch.empty(B, cout, T, H, W, dtype=out_dtype, device=y.device)
    _call("cl_to_ncthw", _lib.load().ccedit_cl_to_ncthw,
          (y.data_ptr(), ld, dst.data_ptr(), int(out_dtype == torch.float32), B, cout, T, H * W, _stream()), nbytes=_nb(y, dst))
    return dst


def activation_as_weight(w: torch.Tensor, bias: Optional[torch.Tensor] = None) -> PackedWeight:
    """A contiguous fp16 activation matrix [N, K] (K % 64 == 0) used as the weight operand of ``gemm`` - the K / V^T
    operands of a GEMM-formulated attention."""
    _require(w, name="w")
    if w.dim() != 2 or not w.is_contiguous() or w.shape[1] % BLOCK_K:
        raise RuntimeError("ccedit_b200.activation_as_weight: w must be contiguous [N, K] with K a multiple of 64")
    n, k = w.shape
    return PackedWeight(w, bias, n, k, k, 1, pick_bn(n))


def conv_s2_taps_asym():
    """3x3, stride 2, NO padding on the top/left and one zero row/column on the bottom/right (the first stage's
    Downsample, model.py:83-91: F.pad(x, (0, 1, 0, 1)) + conv stride 2 padding 0) over the 4 parity planes of
    ``parity_split``: input index 2*o + k -> (parity, offset) = (0, 0), (1, 0), (0, +1)."""
    split = lambda k: (0, 0) if k == 0 else ((1, 0) if k == 1 else (0, 1))
    taps = []
    for kh in range(3):
        ph, oh = split(kh)
        for kw in range(3):
            pw, ow = split(kw)
            taps.append((ow, oh, ph * 2 + pw, 0))
    return taps


# ---------------------------------------------------------------------------------------------------------------------
# text conditioner helpers (CLIP text transformer)
# ---------------------------------------------------------------------------------------------------------------------
def embed_tokens(ids: torch.Tensor, token_emb: torch.Tensor, pos_emb: torch.Tensor) -> torch.Tensor:
    """ids int64 [B, L]; token_emb fp16 [V, D]; pos_emb fp16 [>= L, D] -> fp16 [B * L, D]."""
    _require(token_emb, name="token_emb")
    _require(pos_emb, name="pos_emb")
    if not ids.is_cuda or ids.dtype != torch.int64:
        raise RuntimeError("ccedit_b200.embed_tokens: ids must be a CUDA int64 tensor")
    B, L = ids.shape
    V, D = token_emb.shape
    out = torch.empty(B * L, D, dtype=torch.float16, device=ids.device)
    _call("embed_tokens", _lib.load().ccedit_embed_tokens,
          (ids.contiguous().data_ptr(), token_emb.data_ptr(), pos_emb.data_ptr(), out.data_ptr(), B, L, D, V, _stream()),
          nbytes=2 * _nb(out))
    return out


def quick_gelu_(x: torch.Tensor) -> torch.Tensor:
    """x * sigmoid(1.702 x) in place on a contiguous fp16 tensor."""
    _require(x, name="x")
    if not x.is_contiguous():
        raise RuntimeError("ccedit_b200.quick_gelu_: x must be contiguous")
    _call("quick_gelu", _lib.load().ccedit_quick_gelu, (x.data_ptr(), x.numel(), _stream()), nbytes=2 * _nb(x))
    return x


def causal_attention_small(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, out: torch.Tensor,
                           scale: Optional[float] = None) -> torch.Tensor:
    """q, k, v: [B, L, heads*64] views sharing one row stride; causal softmax(q k^T * scale) v -> out [B, L, heads*64]."""
    for t, nm in ((q, "q"), (k, "k"), (v, "v"), (out, "out")):
        _require(t, name=nm)
    B, L, Cc = q.shape
    dh = Cc // heads
    if not (q.stride(1) == k.stride(1) == v.stride(1)) or q.stride(0) != L * q.stride(1):
        raise RuntimeError("ccedit_b200.causal_attention_small: q, k, v must share a uniform row stride")
    _call("causal_attention_small", _lib.load().ccedit_causal_attention_small,
          (q.data_ptr(), k.data_ptr(), v.data_ptr(), q.stride(1), out.data_ptr(), out.stride(1), B, L, heads, dh,
           float(dh) ** -0.5 if scale is None else scale, _stream()), flops=4.0 * B * L * L * Cc, nbytes=4.0 * B * L * Cc * 2)
    return out

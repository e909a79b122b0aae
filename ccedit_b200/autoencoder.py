"""First stage: mirror of the reference's KL autoencoder on the CUDA kernels of this package.

    AutoencoderKLInferenceWrapper     sgm/models/autoencoder.py:322-343 (AutoencoderKL :283-319)
    Decoder / Encoder                 sgm/modules/diffusionmodules/model.py:617-761 / 498-614
    ResnetBlock, AttnBlock, Upsample, Downsample     model.py:94-151, 161-201, 56-71, 74-91
    decode_first_stage / encode_first_stage          sgm/models/diffusion.py:152-163

Same constructor keywords (`ddconfig`, `embed_dim`), same state-dict keys (`decoder.up.1.block.0.norm1.weight`,
`post_quant_conv.weight`, ...: tests/test_host_logic.py checks them against the reference's), same `decode(z)` /
`encode(x)` call for 4-D and 5-D ("b c t h w") tensors.  Execution: channels-last fp16 buffers [F, H, W, C]; every
conv is the tcgen05 tap-GEMM (3x3: 9 taps with TMA zero fill; the encoder's asymmetric stride-2 conv: 9 taps over
parity planes), GroupNorm(32, C, eps 1e-6) + SiLU is the spatial GroupNorm kernel, nearest x2 the upsample kernel.
The mid block's single-head attention of width 512 over h*w tokens does not fit the flash kernel's TMEM budget (O alone
would take all 512 columns), so it runs per frame as  S = (q * C^-1/2) k^T  (tap-GEMM with the K tokens as the weight
operand) -> row softmax (csrc/elementwise.cu) -> O = P v  (tap-GEMM with V^T as the weight operand; V^T comes straight
out of a GEMM whose activation operand is the v projection weight, and v's bias is added after P v because the rows of P
sum to one).  The reference runs the first stage in fp32 (disable_first_stage_autocast: True); here activations are
stored in fp16 with fp32 accumulation, like the rest of the path - tests/test_vae_gpu.py holds the tolerance.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import ops
from .modules import ParamHolder, conv2d, norm

GN_EPS = 1e-6     # Normalize(): model.py:50-53


def _new(ref: torch.Tensor, *shape):
    return torch.empty(*shape, dtype=torch.float16, device=ref.device)


class ResnetBlock(nn.Module):
    """model.py:94-151 with temb_channels = 0: GN+SiLU -> conv3x3 -> GN+SiLU -> conv3x3, + x (1x1 nin_shortcut if widths differ)."""

    def __init__(self, in_channels: int, out_channels: Optional[int] = None):
        super().__init__()
        out_channels = in_channels if out_channels is None else out_channels
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1 = norm(in_channels)
        self.conv1 = conv2d(in_channels, out_channels, 3)
        self.norm2 = norm(out_channels)
        self.conv2 = conv2d(out_channels, out_channels, 3)
        if in_channels != out_channels:
            self.nin_shortcut = conv2d(in_channels, out_channels, 1)

    def run(self, x4: torch.Tensor) -> torch.Tensor:
        dev = x4.device
        F, H, W, _ = x4.shape
        Co, taps = self.out_channels, ops.conv_taps()
        a = ops.groupnorm_spatial(x4, *self.norm1.affine(dev), GN_EPS, True)
        h = ops.gemm(a, self.conv1.packed(dev), _new(x4, F, H, W, Co), taps)
        a = ops.groupnorm_spatial(h, *self.norm2.affine(dev), GN_EPS, True, out=a if Co == self.in_channels else None)
        skip = x4
        if self.in_channels != Co:
            skip = ops.gemm(x4, self.nin_shortcut.packed(dev), _new(x4, F, H, W, Co))
        return ops.gemm(a, self.conv2.packed(dev), h, taps, res1=skip)      # h is dead after the GroupNorm: reuse it


class AttnBlock(nn.Module):
    """model.py:161-201: GN -> q, k, v (1x1 convs) -> one head of width C over h*w tokens -> proj_out -> + x."""

    def __init__(self, in_channels: int):
        super().__init__()
        self.in_channels = in_channels
        self.norm = norm(in_channels)
        self.q = conv2d(in_channels, in_channels, 1)
        self.k = conv2d(in_channels, in_channels, 1)
        self.v = conv2d(in_channels, in_channels, 1)
        self.proj_out = conv2d(in_channels, in_channels, 1)

    def run(self, x4: torch.Tensor) -> torch.Tensor:
        dev = x4.device
        F, H, W, C = x4.shape
        L = H * W
        Lp = -(-L // 64) * 64                              # key count padded to the GEMM's K granularity
        hn = ops.groupnorm_spatial(x4, *self.norm.affine(dev), GN_EPS, False).view(F, L, C)
        # q carries the softmax scale C^-1/2 (SDPA default, model.py:191-193)
        q = ops.gemm(hn.view(F * L, C), self.q.packed(dev, scale=float(C) ** -0.5), _new(x4, F * L, C)).view(F, L, C)
        k = ops.gemm(hn.view(F * L, C), self.k.packed(dev), _new(x4, F * L, C)).view(F, L, C)
        wv = self.v._cached((dev, "v_as_activation"), lambda: self.v.weight.detach().reshape(C, C).to(
            device=dev, dtype=torch.float16).contiguous())
        bv = self.v.affine(dev)[1]
        o = _new(x4, F, L, C)
        s = _new(x4, L, Lp)
        vt = torch.zeros(C, Lp, dtype=torch.float16, device=dev) if Lp != L else _new(x4, C, Lp)
        for f in range(F):
            # S = q k^T: the frame's K tokens are the weight operand [N = L][K = C]
            ops.gemm(q[f], ops.activation_as_weight(k[f]), s[:, :L])
            ops.softmax_rows(s[:, :L])
            if Lp != L:
                s[:, L:].zero_()                                   # padded keys carry no probability
            # V^T [C][L] = Wv h^T: the projection weight as the activation operand, the tokens as the weight operand
            ops.gemm(wv, ops.activation_as_weight(hn[f]), vt[:, :L])
            # O = P V + b_v (rows of P sum to one)
            ops.gemm(s, ops.activation_as_weight(vt, bv), o[f])
        out = _new(x4, F, H, W, C)
        ops.gemm(o.view(F * L, C), self.proj_out.packed(dev), out.view(F * L, C), res1=x4.view(F * L, C))
        return out


class Upsample(nn.Module):
    """model.py:56-71: nearest x2, conv3x3."""

    def __init__(self, in_channels: int, with_conv: bool = True):
        super().__init__()
        if not with_conv:
            raise NotImplementedError("ccedit_b200: resamp_with_conv=False is outside the hot path")
        self.conv = conv2d(in_channels, in_channels, 3)

    def run(self, x4: torch.Tensor) -> torch.Tensor:
        F, H, W, C = x4.shape
        u = ops.upsample_nearest2x(x4)
        return ops.gemm(u, self.conv.packed(x4.device), _new(x4, F, 2 * H, 2 * W, C), ops.conv_taps())


class Downsample(nn.Module):
    """model.py:74-91: zero pad (right, bottom) by one, conv3x3 stride 2 without padding."""

    def __init__(self, in_channels: int, with_conv: bool = True):
        super().__init__()
        if not with_conv:
            raise NotImplementedError("ccedit_b200: resamp_with_conv=False is outside the hot path")
        self.conv = conv2d(in_channels, in_channels, 3)

    def run(self, x4: torch.Tensor) -> torch.Tensor:
        F, H, W, C = x4.shape
        if H % 2 or W % 2:
            raise RuntimeError("ccedit_b200: the first-stage encoder needs even spatial sizes at every level")
        planes = ops.parity_split(x4)
        out = _new(x4, F, H // 2, W // 2, C)
        ops.gemm(planes, self.conv.packed(x4.device), out.unsqueeze(1), ops.conv_s2_taps_asym())
        return out


class _Level(nn.Module):
    """`up[i]` / `down[i]` of the reference: .block (ModuleList), .attn (empty: attn_resolutions = []), .upsample / .downsample."""

    def __init__(self):
        super().__init__()
        self.block = nn.ModuleList()
        self.attn = nn.ModuleList()


class _Mid(nn.Module):
    def __init__(self, ch: int):
        super().__init__()
        self.block_1 = ResnetBlock(ch, ch)
        self.attn_1 = AttnBlock(ch)
        self.block_2 = ResnetBlock(ch, ch)

    def run(self, h):
        return self.block_2.run(self.attn_1.run(self.block_1.run(h)))


def _check_ddconfig(attn_resolutions, dropout, resamp_with_conv, attn_type, use_linear_attn):
    if list(attn_resolutions) or dropout or not resamp_with_conv or attn_type != "vanilla" or use_linear_attn:
        raise NotImplementedError("ccedit_b200: first-stage ddconfig outside the inference configs "
                                  "(attn_resolutions [], dropout 0, resamp_with_conv, vanilla attention)")


class Decoder(nn.Module):
    """model.py:617-761."""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, give_pre_end=False, tanh_out=False,
                 use_linear_attn=False, attn_type="vanilla", **ignorekwargs):
        super().__init__()
        _check_ddconfig(attn_resolutions, dropout, resamp_with_conv, attn_type, use_linear_attn)
        if give_pre_end or tanh_out:
            raise NotImplementedError("ccedit_b200: give_pre_end / tanh_out are outside the hot path")
        self.ch, self.num_resolutions, self.num_res_blocks, self.out_ch = ch, len(ch_mult), num_res_blocks, out_ch
        block_in = ch * ch_mult[-1]
        self.conv_in = conv2d(z_channels, block_in, 3)
        self.mid = _Mid(block_in)
        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            up = _Level()
            block_out = ch * ch_mult[i_level]
            for _ in range(num_res_blocks + 1):
                up.block.append(ResnetBlock(block_in, block_out))
                block_in = block_out
            if i_level != 0:
                up.upsample = Upsample(block_in, resamp_with_conv)
            self.up.insert(0, up)
        self.norm_out = norm(block_in)
        self.conv_out = conv2d(block_in, out_ch, 3)

    def run(self, z4: torch.Tensor) -> torch.Tensor:
        """z4: channels-last fp16 [F, h, w, >= z_channels] -> [F, 8h, 8w, 16] (first out_ch channels valid)."""
        dev = z4.device
        F, H, W, _ = z4.shape
        pw = self.conv_in.packed(dev)
        h = self.mid.run(ops.gemm(z4, pw, _new(z4, F, H, W, pw.n), ops.conv_taps()))
        for i_level in reversed(range(self.num_resolutions)):
            for blk in self.up[i_level].block:
                h = blk.run(h)
            if i_level != 0:
                h = self.up[i_level].upsample.run(h)
        a = ops.groupnorm_spatial(h, *self.norm_out.affine(dev), GN_EPS, True)
        pw = self.conv_out.packed(dev)
        return ops.gemm(a, pw, _new(z4, h.shape[0], h.shape[1], h.shape[2], pw.n), ops.conv_taps())


class Encoder(nn.Module):
    """model.py:498-614."""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, double_z=True, use_linear_attn=False,
                 attn_type="vanilla", **ignore_kwargs):
        super().__init__()
        _check_ddconfig(attn_resolutions, dropout, resamp_with_conv, attn_type, use_linear_attn)
        self.ch, self.num_resolutions, self.num_res_blocks = ch, len(ch_mult), num_res_blocks
        self.conv_in = conv2d(in_channels, ch, 3)
        in_ch_mult = (1,) + tuple(ch_mult)
        self.down = nn.ModuleList()
        block_in = ch
        for i_level in range(self.num_resolutions):
            down = _Level()
            block_in, block_out = ch * in_ch_mult[i_level], ch * ch_mult[i_level]
            for _ in range(num_res_blocks):
                down.block.append(ResnetBlock(block_in, block_out))
                block_in = block_out
            if i_level != self.num_resolutions - 1:
                down.downsample = Downsample(block_in, resamp_with_conv)
            self.down.append(down)
        self.mid = _Mid(block_in)
        self.norm_out = norm(block_in)
        self.conv_out = conv2d(block_in, 2 * z_channels if double_z else z_channels, 3)

    def run(self, x4: torch.Tensor) -> torch.Tensor:
        """x4: channels-last fp16 [F, H, W, 8] (3 image channels, zero padded) -> [F, H/8, W/8, 16] (first 2 z_channels valid)."""
        dev = x4.device
        F, H, W, _ = x4.shape
        pw = self.conv_in.packed(dev)
        h = ops.gemm(x4, pw, _new(x4, F, H, W, pw.n), ops.conv_taps())
        for i_level in range(self.num_resolutions):
            for blk in self.down[i_level].block:
                h = blk.run(h)
            if i_level != self.num_resolutions - 1:
                h = self.down[i_level].downsample.run(h)
        h = self.mid.run(h)
        a = ops.groupnorm_spatial(h, *self.norm_out.affine(dev), GN_EPS, True)
        pw = self.conv_out.packed(dev)
        return ops.gemm(a, pw, _new(x4, h.shape[0], h.shape[1], h.shape[2], pw.n), ops.conv_taps())


class AutoencoderKLInferenceWrapper(nn.Module):
    """autoencoder.py:283-343.  `decode(z)` takes [B, C, h, w] or [B, C, T, h, w] latents (already divided by the scale
    factor, as decode_first_stage does, diffusion.py:152-156) and returns the image / video in the same layout, fp32.
    `encode_moments(x)` returns quant_conv(encoder(x)) = [mean | logvar]; `encode(x)` samples from it like the reference
    (DiagonalGaussianDistribution.sample: mean + exp(0.5 * clamp(logvar, -30, 20)) * randn)."""

    def __init__(self, ddconfig, embed_dim: int, lossconfig=None, ckpt_path=None, ignore_keys=(), monitor=None, **kwargs):
        super().__init__()
        if ckpt_path is not None:
            raise NotImplementedError("ccedit_b200: load the first stage through load_state_dict / ccedit_b200.checkpoint")
        self.encoder = Encoder(**ddconfig)
        self.decoder = Decoder(**ddconfig)
        self.quant_conv = conv2d(2 * ddconfig["z_channels"], 2 * embed_dim, 1)
        self.post_quant_conv = conv2d(embed_dim, ddconfig["z_channels"], 1)
        self.embed_dim = embed_dim
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate())

    def invalidate(self):
        """Drop the packed fp16 weights (after an edit the parameter version counters cannot see)."""
        for m in self.modules():
            if isinstance(m, ParamHolder):
                m.invalidate()

    @staticmethod
    def _frames(t: torch.Tensor):
        if t.dim() == 5:
            return t, True
        if t.dim() != 4:
            raise RuntimeError("ccedit_b200: first-stage tensors are [B, C, h, w] or [B, C, T, h, w]")
        return t.unsqueeze(2), False

    def decode(self, z: torch.Tensor, scale: float = 1.0, **decoder_kwargs) -> torch.Tensor:
        """`scale` folds decode_first_stage's `1 / scale_factor * z` into the layout change."""
        if not z.is_cuda:
            raise RuntimeError("ccedit_b200: the first stage runs on CUDA (sm_100a) only; there is no CPU fallback")
        with torch.no_grad():
            z5, is_video = self._frames(z)
            B, Cz, T, h, w = z5.shape
            dev = z.device
            z_cl = ops.ncthw_to_cl(z5, 8, mul=scale).view(B * T, h, w, 8)
            pq = self.post_quant_conv.packed(dev)
            zq = ops.gemm(z_cl.view(B * T * h * w, 8), pq, _new(z_cl, B * T * h * w, pq.n)).view(B * T, h, w, pq.n)
            img = self.decoder.run(zq)                                                   # [F, 8h, 8w, 16]
            out = ops.cl_to_ncthw(img.view(B, T, img.shape[1], img.shape[2], img.shape[3]), self.decoder.out_ch)
            return out if is_video else out[:, :, 0]

    def encode_moments(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("ccedit_b200: the first stage runs on CUDA (sm_100a) only; there is no CPU fallback")
        with torch.no_grad():
            x5, is_video = self._frames(x)
            B, _, T, H, W = x5.shape
            x_cl = ops.ncthw_to_cl(x5, 8).view(B * T, H, W, 8)
            hm = self.encoder.run(x_cl)                                                   # [F, H/8, W/8, 16]
            F, hh, ww, C = hm.shape
            qc = self.quant_conv.packed(x.device)
            m = ops.gemm(hm.view(F * hh * ww, C), qc, _new(hm, F * hh * ww, qc.n)).view(B, T, hh, ww, qc.n)
            out = ops.cl_to_ncthw(m, 2 * self.embed_dim)
            return out if is_video else out[:, :, 0]

    def encode(self, x: torch.Tensor) -> torch.Tensor:
        mean, logvar = torch.chunk(self.encode_moments(x), 2, dim=1)
        std = torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0))
        return mean + std * torch.randn(mean.shape, device=mean.device)

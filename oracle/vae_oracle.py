"""ORACLE (test infrastructure, NOT product code): CPU fp32 restatement of the first-stage decoder / encoder
(SURVEY.md section 8 row f1), state-dict driven, in plain F.* ops.

    AutoencoderKLInferenceWrapper.decode / encode     sgm/models/autoencoder.py:322-343 (AutoencoderKL :306-319)
    Decoder.forward                                   sgm/modules/diffusionmodules/model.py:728-761 (__init__ :617-718)
    Encoder.forward                                   sgm/modules/diffusionmodules/model.py:587-614 (__init__ :498-585)
    ResnetBlock.forward / AttnBlock / Upsample / Downsample   model.py:131-151, 161-201, 56-71, 74-91
    decode_first_stage / encode_first_stage           sgm/models/diffusion.py:152-163  (1 / scale_factor, scale_factor)

Only tests/ (and tools/ reports) may import this; pinned against the unmodified reference classes by
oracle/make_golden.py vae -> tests/golden/vae.pt, checked in tests/test_oracle_golden.py.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F
from einops import rearrange

SD = Dict[str, torch.Tensor]

# first_stage_config.params.ddconfig of configs/inference_ccedit/keyframe_no2ndca_depthmidas.yaml:80-90
DDCONFIG = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
                num_res_blocks=2, attn_resolutions=[], dropout=0.0)
SCALE_FACTOR = 0.18215          # keyframe_no2ndca_depthmidas.yaml:7


def _gn(sd, p, x):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], 1e-6)       # Normalize(): model.py:50-53


def _conv(sd, p, x, stride=1, padding=1):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=padding)


def _swish(x):
    return x * torch.sigmoid(x)                                               # nonlinearity(): model.py:45-47


def resnet_block(sd: SD, p: str, x):
    """ResnetBlock.forward with temb = None, model.py:131-151."""
    h = _conv(sd, p + ".conv1", _swish(_gn(sd, p + ".norm1", x)))
    h = _conv(sd, p + ".conv2", _swish(_gn(sd, p + ".norm2", h)))
    if (p + ".nin_shortcut.weight") in sd:
        x = _conv(sd, p + ".nin_shortcut", x, padding=0)
    return x + h


def attn_block(sd: SD, p: str, x):
    """AttnBlock.forward, model.py:179-201: single-head attention over h*w tokens of width C, scale C^-1/2."""
    h_ = _gn(sd, p + ".norm", x)
    q, k, v = (_conv(sd, f"{p}.{n}", h_, padding=0) for n in ("q", "k", "v"))
    b, c, h, w = q.shape
    q, k, v = (rearrange(t, "b c h w -> b 1 (h w) c").contiguous() for t in (q, k, v))
    o = F.scaled_dot_product_attention(q, k, v)
    o = rearrange(o, "b 1 (h w) c -> b c h w", h=h, w=w, c=c, b=b)
    return x + _conv(sd, p + ".proj_out", o, padding=0)


def upsample(sd: SD, p: str, x):
    """Upsample.forward (with_conv), model.py:65-71."""
    x = F.interpolate(x.to(torch.float32), scale_factor=2.0, mode="nearest").to(x.dtype)
    return _conv(sd, p + ".conv", x)


def downsample(sd: SD, p: str, x):
    """Downsample.forward (with_conv): pad right/bottom by one, conv 3x3 stride 2 without padding, model.py:83-91."""
    return _conv(sd, p + ".conv", F.pad(x, (0, 1, 0, 1), mode="constant", value=0), stride=2, padding=0)


def decoder_forward(sd: SD, z, cfg=DDCONFIG, p="decoder"):
    """Decoder.forward, model.py:728-761 (give_pre_end / tanh_out False, temb None)."""
    nres, nblk = len(cfg["ch_mult"]), cfg["num_res_blocks"]
    h = _conv(sd, p + ".conv_in", z)
    h = resnet_block(sd, p + ".mid.block_1", h)
    h = attn_block(sd, p + ".mid.attn_1", h)
    h = resnet_block(sd, p + ".mid.block_2", h)
    for i_level in reversed(range(nres)):
        for i_block in range(nblk + 1):
            h = resnet_block(sd, f"{p}.up.{i_level}.block.{i_block}", h)
        if i_level != 0:
            h = upsample(sd, f"{p}.up.{i_level}.upsample", h)
    return _conv(sd, p + ".conv_out", _swish(_gn(sd, p + ".norm_out", h)))


def encoder_forward(sd: SD, x, cfg=DDCONFIG, p="encoder"):
    """Encoder.forward, model.py:587-614."""
    nres, nblk = len(cfg["ch_mult"]), cfg["num_res_blocks"]
    h = _conv(sd, p + ".conv_in", x)
    for i_level in range(nres):
        for i_block in range(nblk):
            h = resnet_block(sd, f"{p}.down.{i_level}.block.{i_block}", h)
        if i_level != nres - 1:
            h = downsample(sd, f"{p}.down.{i_level}.downsample", h)
    h = resnet_block(sd, p + ".mid.block_1", h)
    h = attn_block(sd, p + ".mid.attn_1", h)
    h = resnet_block(sd, p + ".mid.block_2", h)
    return _conv(sd, p + ".conv_out", _swish(_gn(sd, p + ".norm_out", h)))


def decode_first_stage(sd: SD, z, scale_factor=SCALE_FACTOR):
    """diffusion.py:152-156 + AutoencoderKLInferenceWrapper.decode, autoencoder.py:333-343."""
    z = 1.0 / scale_factor * z
    is_video = z.dim() == 5
    if is_video:
        b, _, t, _, _ = z.shape
        z = rearrange(z, "b c t h w -> (b t) c h w")
    dec = decoder_forward(sd, _conv(sd, "post_quant_conv", z, padding=0))
    return rearrange(dec, "(b t) c h w -> b c t h w", b=b, t=t) if is_video else dec


def encode_first_stage_moments(sd: SD, x):
    """AutoencoderKL.encode up to the posterior parameters, autoencoder.py:306-313: [mean | logvar] on dim 1
    (the sample adds seeded noise, DiagonalGaussianDistribution; parity is checked on the moments)."""
    is_video = x.dim() == 5
    if is_video:
        b, _, t, _, _ = x.shape
        x = rearrange(x, "b c t h w -> (b t) c h w")
    m = _conv(sd, "quant_conv", encoder_forward(sd, x), padding=0)
    return rearrange(m, "(b t) c h w -> b c t h w", b=b, t=t) if is_video else m

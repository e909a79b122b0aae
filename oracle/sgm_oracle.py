"""ORACLE (test infrastructure, NOT product code): CPU fp32 restatement of CCEdit's denoising hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module, and
only as the checker / CPU baseline.  The product (ccedit_b200/) never imports it.

What it is: a functional, state-dict driven restatement in plain PyTorch fp32 ops (F.conv2d, F.group_norm,
F.scaled_dot_product_attention, ...) of
    OpenAIWrapperControlLDM3DTV2V.forward      sgm/modules/diffusionmodules/wrappers.py:155-207
    ControlNet2D.forward                       sgm/modules/diffusionmodules/controlmodel.py:252-317
    ControlledUNetModel3DTV2V.forward          sgm/modules/diffusionmodules/controlmodel.py:471-550
and of the sampler-side callers (DiscreteDenoiser, EpsScaling, VanillaCFGTV2V, DPMPP2SAncestralSampler).
It keeps the reference's own layouts and rearranges ("b c t h w", "(b t) c h w", "(b h w) c t") so that every line
can be compared with the file:line it cites.

Pinning: the reference has no tests or golden vectors of its own.  This restatement is pinned
against the reference itself: oracle/make_golden.py imports the unmodified reference modules from /root/reference on
CPU, runs them on seeded weights/inputs (oracle/weights.py) and commits inputs+outputs under tests/golden/;
tests/test_oracle_golden.py checks this file against those fixtures.
"""
from __future__ import annotations

import contextlib
import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F
from einops import rearrange, repeat

SD = Dict[str, torch.Tensor]

TV2V_UNET_CFG = dict(
    in_channels=4, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1], num_res_blocks=2,
    channel_mult=[1, 2, 4, 4], num_heads=8, transformer_depth=1, context_dim=768,
    enable_attention3d_crossframe=False, ST3DCA_ca_type=None,
)
TV2V_CONTROLNET_CFG = dict(
    in_channels=4, hint_channels=3, model_channels=320, attention_resolutions=[4, 2, 1], num_res_blocks=2,
    channel_mult=[1, 2, 4, 4], num_heads=8, transformer_depth=1, context_dim=768, control_scales=1.0,
    no_add_x=False, set_input_hint_block_as_identity=False, disable_text_ca=False,
)


# ---------------------------------------------------------------------------------------------------------------------
# structure (UNetModel.__init__, openaimodel.py:1254-1527): which sub-blocks each input/middle/output block holds
# ---------------------------------------------------------------------------------------------------------------------
def unet_plan(cfg) -> dict:
    mc, mult = cfg["model_channels"], list(cfg["channel_mult"])
    nrb = cfg["num_res_blocks"]
    nrb = [nrb] * len(mult) if isinstance(nrb, int) else list(nrb)
    att = list(cfg["attention_resolutions"])
    heads = cfg["num_heads"]
    inp = [[("conv_in", cfg["in_channels"], mc)]]
    chans = [mc]
    ch, ds = mc, 1
    for level, m in enumerate(mult):
        for _ in range(nrb[level]):
            layers = [("res", ch, m * mc)]
            ch = m * mc
            if ds in att:
                layers.append(("attn", ch, ch // heads))  # legacy=False: dim_head = ch // num_heads (:1283-1284)
            inp.append(layers)
            chans.append(ch)
        if level != len(mult) - 1:
            inp.append([("down", ch, ch)])
            chans.append(ch)
            ds *= 2
    mid = [("res", ch, ch), ("attn", ch, ch // heads), ("res", ch, ch)]
    out = []
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(nrb[level] + 1):
            ich = chans.pop()
            layers = [("res", ch + ich, mc * m)]
            ch = mc * m
            if ds in att:
                layers.append(("attn", ch, ch // heads))
            if level and i == nrb[level]:
                layers.append(("up", ch, ch))
                ds //= 2
            out.append(layers)
    return dict(input=inp, middle=mid, output=out, heads=heads)


# ---------------------------------------------------------------------------------------------------------------------
# fp16-storage emulation (test-side only).  The reference runs this path on a GPU under torch.cuda.amp.autocast()
# (scripts/sampling/sampling_tv2v.py:361-362): conv / linear / SDPA take fp16 operands and return fp16 tensors, the
# norms and softmax run in fp32.  `with emulate_half_storage():` makes the oracle round exactly those operands and
# results to fp16 (arithmetic stays fp32 on the CPU, i.e. fp32 accumulation as on tensor cores).  The distance between
# this variant and the plain fp32 oracle is the error ANY fp16-storage implementation of the path carries; the parity
# tests compare the CUDA path's own distance with it (tests/test_network_gpu.py, profiles/r02_parity.md).
# ---------------------------------------------------------------------------------------------------------------------
_HALF = False


@contextlib.contextmanager
def emulate_half_storage(enabled: bool = True):
    global _HALF
    old, _HALF = _HALF, enabled
    try:
        yield
    finally:
        _HALF = old


def _r(t):
    return t.half().float() if (_HALF and t is not None) else t


# ---------------------------------------------------------------------------------------------------------------------
# leaf ops
# ---------------------------------------------------------------------------------------------------------------------
def _gn(sd: SD, p: str, x, eps):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps)


def _conv2d(sd: SD, p: str, x, stride=1, padding=1):
    return _r(F.conv2d(_r(x), _r(sd[p + ".weight"]), sd.get(p + ".bias"), stride=stride, padding=padding))


def _conv1d(sd: SD, p: str, x, padding):
    return _r(F.conv1d(_r(x), _r(sd[p + ".weight"]), sd.get(p + ".bias"), padding=padding))


def _linear(sd: SD, p: str, x):
    return _r(F.linear(_r(x), _r(sd[p + ".weight"]), sd.get(p + ".bias")))


def timestep_embedding(timesteps, dim, max_period=10000):
    """util.py:244-268."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device) / half)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def time_embed(sd: SD, p: str, t_emb):
    """openaimodel.py:1216-1223: Linear -> SiLU -> Linear."""
    return _linear(sd, p + ".2", F.silu(_linear(sd, p + ".0", t_emb)))


def spatial_temporal_forward(x, spatial, temporal):
    """openaimodel.py:129-178: y = S(x) per frame; out = y + Tm(y) per pixel over T (Tm None => + 0)."""
    b = x.shape[0]
    x = rearrange(x, "b c t h w -> (b t) c h w")
    x = spatial(x)
    _, _, h, w = x.shape
    x = rearrange(x, "(b t) c h w -> (b h w) c t", b=b)
    identity = x
    x = temporal(x) if temporal is not None else torch.zeros_like(identity)
    x = x + identity
    return rearrange(x, "(b h w) c t -> b c t h w", h=h, w=w)


# ---------------------------------------------------------------------------------------------------------------------
# attention.py blocks
# ---------------------------------------------------------------------------------------------------------------------
def cross_attention(sd: SD, p: str, x, context, heads):
    """CrossAttention.forward, attention.py:392-467 (SDPA, scale = dim_head^-0.5)."""
    q = _linear(sd, p + ".to_q", x)
    context = x if context is None else context
    k = _linear(sd, p + ".to_k", context)
    v = _linear(sd, p + ".to_v", context)
    q, k, v = (rearrange(t, "b n (h d) -> b h n d", h=heads) for t in (q, k, v))
    out = _r(F.scaled_dot_product_attention(q, k, v))
    out = rearrange(out, "b h n d -> b n (h d)")
    return _linear(sd, p + ".to_out.0", out)


def feed_forward(sd: SD, p: str, x):
    """FeedForward with GEGLU, attention.py:115-141: value first, gate second, exact-erf gelu."""
    a, gate = _linear(sd, p + ".net.0.proj", x).chunk(2, dim=-1)
    return _linear(sd, p + ".net.2", a * F.gelu(gate))


def _ln(sd: SD, p: str, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def basic_transformer_block(sd: SD, p: str, x, context, heads):
    """BasicTransformerBlock._forward, attention.py:695-716 (disable_self_attn False)."""
    x = cross_attention(sd, p + ".attn1", _ln(sd, p + ".norm1", x), None, heads) + x
    x = cross_attention(sd, p + ".attn2", _ln(sd, p + ".norm2", x), context, heads) + x
    x = feed_forward(sd, p + ".ff", _ln(sd, p + ".norm3", x)) + x
    return x


def single_layer_block(sd: SD, p: str, x, context, heads):
    """BasicTransformerSingleLayerBlock._forward, attention.py:758-761: context is NOT normalised."""
    x = cross_attention(sd, p + ".attn1", _ln(sd, p + ".norm1", x), context, heads) + x
    x = feed_forward(sd, p + ".ff", _ln(sd, p + ".norm2", x)) + x
    return x


def spatial_transformer(sd: SD, p: str, x, context, heads, disable_text_ca=False):
    """SpatialTransformer.forward (2-D), attention.py:865-889; depth 1; 1x1-conv projections (use_linear False)."""
    b, c, h, w = x.shape
    x_in = x
    x = _gn(sd, p + ".norm", x, 1e-6)
    x = _conv2d(sd, p + ".proj_in", x, padding=0)
    x = rearrange(x, "b c h w -> b (h w) c")
    if disable_text_ca:
        x = single_layer_block(sd, p + ".transformer_blocks.0", x, x, heads)
    else:
        x = basic_transformer_block(sd, p + ".transformer_blocks.0", x, context, heads)
    x = rearrange(x, "b (h w) c -> b c h w", h=h, w=w)
    x = _conv2d(sd, p + ".proj_out", x, padding=0)
    return x + x_in


def spatial_transformer_3d(sd: SD, p: str, x, context, heads, ca_type: Optional[str] = None):
    """SpatialTransformer3D.forward attention.py:1141-1208 (+ SpatialTransformer3DCA.forward :1302-1350 if ca_type)."""
    b, c, t, h, w = x.shape
    x = rearrange(x, "b c t h w -> (b t) c h w")
    x_in = x
    x = _gn(sd, p + ".norm", x, 1e-6)
    x = _conv2d(sd, p + ".proj_in", x, padding=0)
    x = rearrange(x, "bt c h w -> bt (h w) c")
    ctx = repeat(context, "b l c -> (b t) l c", t=t) if context is not None else None
    x = basic_transformer_block(sd, p + ".transformer_blocks.0", x, ctx, heads)
    x = rearrange(x, "bt (h w) c -> bt c h w", h=h, w=w)
    x = _conv2d(sd, p + ".proj_out", x, padding=0)
    x = x + x_in
    # temporal attention (disable_temporal_text_ca: block(x, context=x), :1191-1192)
    x = rearrange(x, "(b t) c h w -> (b h w) c t", t=t)
    x_in = x
    x = _gn(sd, p + ".norm_temporal", x, 1e-6)
    x = _conv1d(sd, p + ".proj_in_temporal", x, 0)
    x = rearrange(x, "bhw c t -> bhw t c")
    x = single_layer_block(sd, p + ".transformer_blocks_temporal.0", x, x, heads)
    x = rearrange(x, "bhw t c -> bhw c t")
    x = _conv1d(sd, p + ".proj_out_temporal", x, 0)
    x = x_in + x
    x = rearrange(x, "(b h w) c t -> b c t h w", h=h, w=w)
    if ca_type is None:
        return x
    # cross-frame attention, attention.py:1302-1350
    x = rearrange(x, "b c t h w -> (b t) c h w")
    x_in = x
    x = _gn(sd, p + ".norm_temporal_ca", x, 1e-6)
    x = _conv2d(sd, p + ".proj_in_temporal_ca", x, padding=0)
    x = rearrange(x, "bt c h w -> bt (h w) c")
    xb = rearrange(x, "(b t) hw c -> b t hw c", b=b)
    anchor = repeat(xb[:, t // 2], "b hw c -> (b t) hw c", t=t)
    if ca_type == "center":
        ctx_tex = anchor
    elif ca_type == "self":
        ctx_tex = x
    elif ca_type == "center_self":
        ctx_tex = torch.cat([anchor, x], dim=1)
    else:
        raise NotImplementedError(ca_type)
    x = single_layer_block(sd, p + ".transformer_blocks_temporal_ca.0", x, ctx_tex, heads)
    x = rearrange(x, "bt (h w) c -> bt c h w", h=h, w=w)
    x = _conv2d(sd, p + ".proj_out_temporal_ca", x, padding=0)
    x = x + x_in
    return rearrange(x, "(b t) c h w -> b c t h w", b=b, t=t)


# ---------------------------------------------------------------------------------------------------------------------
# openaimodel.py blocks
# ---------------------------------------------------------------------------------------------------------------------
def resblock(sd: SD, p: str, x, emb):
    """ResBlock._forward (2-D), openaimodel.py:528-554."""
    h = _conv2d(sd, p + ".in_layers.2", F.silu(_gn(sd, p + ".in_layers.0", x, 1e-5)))
    emb_out = _linear(sd, p + ".emb_layers.1", F.silu(emb))[..., None, None]
    h = h + emb_out
    h = _conv2d(sd, p + ".out_layers.3", F.silu(_gn(sd, p + ".out_layers.0", h, 1e-5)))
    if (p + ".skip_connection.weight") in sd:
        x = _conv2d(sd, p + ".skip_connection", x, padding=0)
    return x + h


def resblock3d(sd: SD, p: str, x, emb):
    """ResBlock3D._forward, openaimodel.py:730-775 (no up/down, no scale-shift norm)."""
    identity = x
    x = spatial_temporal_forward(
        x,
        lambda z: _conv2d(sd, p + ".in_layers.2", F.silu(_gn(sd, p + ".in_layers.0", z, 1e-5))),
        lambda z: _conv1d(sd, p + ".in_layers_temporal.2", F.silu(_gn(sd, p + ".in_layers_temporal.0", z, 1e-5)), 1),
    )
    emb_out = _linear(sd, p + ".emb_layers.1", F.silu(emb))[..., None, None, None]
    x = x + emb_out
    x = spatial_temporal_forward(
        x,
        lambda z: _conv2d(sd, p + ".out_layers.3", F.silu(_gn(sd, p + ".out_layers.0", z, 1e-5))),
        lambda z: _conv1d(sd, p + ".out_layers_temporal.3", F.silu(_gn(sd, p + ".out_layers_temporal.0", z, 1e-5)), 1),
    )
    if (p + ".skip_connection.weight") in sd:
        identity = spatial_temporal_forward(
            identity,
            lambda z: _conv2d(sd, p + ".skip_connection", z, padding=0),
            lambda z: _conv1d(sd, p + ".skip_connection_temporal", z, 0),
        )
    return identity + x


def downsample(sd: SD, p: str, x):
    """Downsample.forward, openaimodel.py:320-322 (conv 3x3 stride 2 pad 1)."""
    return _conv2d(sd, p + ".op", x, stride=2, padding=1)


def downsample3d(sd: SD, p: str, x):
    """Downsample3D.forward, openaimodel.py:388-394."""
    return spatial_temporal_forward(x, lambda z: _conv2d(sd, p + ".op", z, stride=2, padding=1),
                                    lambda z: _conv1d(sd, p + ".conv_temporal", z, 1))


def upsample3d(sd: SD, p: str, x):
    """Upsample3D.forward, openaimodel.py:254-263 (nearest x2 on H, W only)."""
    x = F.interpolate(x.float(), (x.shape[2], x.shape[3] * 2, x.shape[4] * 2), mode="nearest")
    return spatial_temporal_forward(x, lambda z: _conv2d(sd, p + ".conv", z),
                                    lambda z: _conv1d(sd, p + ".conv_temporal", z, 1))


# ---------------------------------------------------------------------------------------------------------------------
# networks
# ---------------------------------------------------------------------------------------------------------------------
def _run_layers_2d(sd, p, layers, h, emb, context, heads, disable_text_ca):
    for j, (kind, _cin, _cout) in enumerate(layers):
        q = f"{p}.{j}"
        if kind == "conv_in":
            h = _conv2d(sd, q, h)
        elif kind == "res":
            h = resblock(sd, q, h, emb)
        elif kind == "attn":
            h = spatial_transformer(sd, q, h, context, heads, disable_text_ca)
        elif kind == "down":
            h = downsample(sd, q, h)
        else:
            raise ValueError(kind)
    return h


def controlnet2d_forward(sd: SD, cfg, x, hint, timesteps, context, prefix="") -> List[torch.Tensor]:
    """ControlNet2D.forward, controlmodel.py:252-317. Returns the list of 13 control tensors."""
    p = prefix
    plan = unet_plan(cfg)
    heads = plan["heads"]
    dtc = cfg.get("disable_text_ca", False)
    emb = time_embed(sd, p + "time_embed", timestep_embedding(timesteps, cfg["model_channels"]))
    is_video = x.dim() == 5
    if is_video:
        n_frames = x.shape[2]
        x = rearrange(x, "b c t h w -> (b t) c h w")
        hint = rearrange(hint, "b c t h w -> (b t) c h w")
        emb = repeat(emb, "b d -> (b t) d", t=n_frames)
        context = repeat(context, "b n d -> (b t) n d", t=n_frames) if context is not None else None
    if cfg.get("set_input_hint_block_as_identity", False):
        guided_hint = _run_layers_2d(sd, p + "input_blocks.0", plan["input"][0], hint, emb, context, heads, dtc)
    else:
        g = hint
        strides = [1, 1, 2, 1, 2, 1, 2, 1]
        for i, s in enumerate(strides):  # controlmodel.py:215-231
            g = _conv2d(sd, f"{p}input_hint_block.{2 * i}", g, stride=s, padding=1)
            if i < 7:
                g = F.silu(g)
        guided_hint = g
    outs = []
    h = x
    for i, layers in enumerate(plan["input"]):
        if guided_hint is not None:
            if cfg.get("no_add_x", False):
                h = guided_hint
            else:
                h = _run_layers_2d(sd, f"{p}input_blocks.{i}", layers, h, emb, context, heads, dtc)
                h = h + guided_hint
            guided_hint = None
        else:
            h = _run_layers_2d(sd, f"{p}input_blocks.{i}", layers, h, emb, context, heads, dtc)
        outs.append(_conv2d(sd, f"{p}zero_convs.{i}.0", h, padding=0))
    h = _run_layers_2d(sd, p + "middle_block", plan["middle"], h, emb, context, heads, dtc)
    outs.append(_conv2d(sd, p + "middle_block_out.0", h, padding=0))
    control = [c * cfg.get("control_scales", 1.0) for c in outs]
    if is_video:
        control = [rearrange(c, "(b t) c h w -> b c t h w", t=n_frames) for c in control]
    return control


def _run_layers_3d(sd, p, layers, h, emb, context, heads, ca_type):
    for j, (kind, _cin, _cout) in enumerate(layers):
        q = f"{p}.{j}"
        if kind == "res":
            h = resblock3d(sd, q, h, emb)
        elif kind == "attn":
            h = spatial_transformer_3d(sd, q, h, context, heads, ca_type)
        elif kind == "down":
            h = downsample3d(sd, q, h)
        elif kind == "up":
            h = upsample3d(sd, q, h)
        else:
            raise ValueError(kind)
    return h


def unet3d_forward(sd: SD, cfg, x, timesteps, context, control=None, img_control=None, prefix=""):
    """ControlledUNetModel3DTV2V.forward, controlmodel.py:471-550."""
    p = prefix
    plan = unet_plan(cfg)
    heads = plan["heads"]
    ca_type = cfg.get("ST3DCA_ca_type") if cfg.get("enable_attention3d_crossframe", False) else None
    control = None if control is None else list(control)
    img_control = None if img_control is None else list(img_control)
    emb = time_embed(sd, p + "time_embed", timestep_embedding(timesteps, cfg["model_channels"]))
    hs = []
    h = x
    for i, layers in enumerate(plan["input"]):
        if i == 0:
            h = spatial_temporal_forward(h, lambda z: _conv2d(sd, p + "input_blocks.0.0", z),
                                         lambda z: _conv1d(sd, p + "input_blocks_temporal.0", z, 1))
        else:
            h = _run_layers_3d(sd, f"{p}input_blocks.{i}", layers, h, emb, context, heads, ca_type)
        if img_control is not None:
            h = h.clone()
            h[:, :, h.shape[2] // 2] += img_control.pop(0)
        hs.append(h)
    h = _run_layers_3d(sd, p + "middle_block", plan["middle"], h, emb, context, heads, ca_type)
    if img_control is not None:
        h = h.clone()
        h[:, :, h.shape[2] // 2] += img_control.pop(0)
    if control is not None:
        h = h + control.pop()
    for i, layers in enumerate(plan["output"]):
        if control is None:
            h = torch.cat([h, hs.pop()], dim=1)
        else:
            h = torch.cat([h, hs.pop() + control.pop()], dim=1)
        h = _run_layers_3d(sd, f"{p}output_blocks.{i}", layers, h, emb, context, heads, ca_type)
    return spatial_temporal_forward(
        h,
        lambda z: _conv2d(sd, p + "out.2", F.silu(_gn(sd, p + "out.0", z, 1e-5))),
        lambda z: _conv1d(sd, p + "out_temporal.1", F.silu(z), 1),
    )


def wrapper_forward(sd: SD, unet_cfg, cn_cfg, x, t, c: dict, cn_img_cfg=None, prefix="diffusion_model."):
    """OpenAIWrapperControlLDM3DTV2V.forward, wrappers.py:155-207."""
    hint = 1.0 - (c["control_hint"] + 1) / 2.0
    ctx = c.get("crossattn")
    control = controlnet2d_forward(sd, cn_cfg, x, hint, t, ctx, prefix + "controlnet.")
    img_control = None
    if c.get("cond_feat") is not None:
        img_control = controlnet2d_forward(sd, cn_img_cfg, x[:, :, x.shape[2] // 2], c["cond_feat"], t, ctx,
                                           prefix + "controlnet_img.")
    return unet3d_forward(sd, unet_cfg, x, t, ctx, control, img_control, prefix)


# ---------------------------------------------------------------------------------------------------------------------
# sampler-side callers (a20): discretisation, denoiser, CFG, DPM++2S ancestral
# ---------------------------------------------------------------------------------------------------------------------
def legacy_ddpm_alphas_cumprod(num_timesteps=1000, linear_start=0.00085, linear_end=0.0120):
    """LegacyDDPMDiscretization.__init__, discretizer.py:42-56 + make_beta_schedule('linear'), util.py:25-37
    (float64 torch.linspace -> numpy -> np.cumprod, exactly as the reference)."""
    import numpy as np

    betas = (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, num_timesteps, dtype=torch.float64) ** 2).numpy()
    return np.cumprod(1.0 - betas, axis=0)


def legacy_ddpm_sigmas(n: int, num_timesteps=1000, do_append_zero=True, flip=False):
    """LegacyDDPMDiscretization.get_sigmas + Discretization.__call__, discretizer.py:11-21,58-69."""
    import numpy as np

    ac = legacy_ddpm_alphas_cumprod(num_timesteps)
    if n < num_timesteps:
        ts = np.linspace(num_timesteps - 1, 0, n, endpoint=False).astype(int)[::-1]
        ac = ac[ts]
    elif n != num_timesteps:
        raise ValueError
    sig = torch.tensor((1 - ac) / ac, dtype=torch.float32) ** 0.5
    sig = torch.flip(sig, (0,))
    if do_append_zero:
        sig = torch.cat([sig, sig.new_zeros([1])])
    return sig if not flip else torch.flip(sig, (0,))


class DiscreteDenoiserOracle:
    """DiscreteDenoiser + EpsScaling: denoiser.py:22-40,43-75; denoiser_scaling.py:16-22."""

    def __init__(self, num_idx=1000):
        # DiscreteDenoiser.__init__ (denoiser.py:54-57): discretization(num_idx, do_append_zero=False, flip=True)
        self.sigmas = legacy_ddpm_sigmas(num_idx, num_idx, do_append_zero=False, flip=True)

    def sigma_to_idx(self, sigma):
        dists = sigma - self.sigmas[:, None]
        return dists.abs().argmin(dim=0).view(sigma.shape)

    def idx_to_sigma(self, idx):
        return self.sigmas[idx]

    def __call__(self, network, x, sigma, cond):
        sigma = self.idx_to_sigma(self.sigma_to_idx(sigma))
        sigma_shape = sigma.shape
        sigma = sigma.reshape(sigma.shape + (1,) * (x.dim() - 1))
        c_skip, c_out, c_in, c_noise = torch.ones_like(sigma), -sigma, 1 / (sigma ** 2 + 1.0) ** 0.5, sigma.clone()
        c_noise = self.sigma_to_idx(c_noise.reshape(sigma_shape))
        return network(x * c_in, c_noise, cond) * c_out + x * c_skip


def cfg_prepare_inputs(x, s, c: dict, uc: dict):
    """VanillaCFGTV2V.prepare_inputs, guiders.py:56-67: uncond first."""
    c_out = {}
    for k in c:
        if k in ("vector", "crossattn", "concat", "control_hint", "cond_feat"):
            c_out[k] = torch.cat((uc[k], c[k]), 0)
        else:
            assert c[k] == uc[k]
            c_out[k] = c[k]
    return torch.cat([x] * 2), torch.cat([s] * 2), c_out


def cfg_combine(x, scale):
    """VanillaCFG.__call__, guiders.py:25-29."""
    x_u, x_c = x.chunk(2)
    return x_u + scale * (x_c - x_u)


def dpmpp2s_ancestral_sample(denoiser_fn, x, cond, uc, num_steps, scale, noises: List[torch.Tensor], eta=1.0,
                             s_noise=1.0):
    """AncestralSampler.__call__ + DPMPP2SAncestralSampler.sampler_step, sampling.py:190-205, 370-407.

    denoiser_fn(x, sigma, c) -> denoised; `noises` replaces torch.randn_like (one tensor per step, consumed in order).
    """
    sigmas = legacy_ddpm_sigmas(num_steps)
    x = x * torch.sqrt(1.0 + sigmas[0] ** 2.0)
    s_in = x.new_ones([x.shape[0]])

    def denoise(xx, sigma):
        return cfg_combine(denoiser_fn(*cfg_prepare_inputs(xx, sigma, cond, uc)), scale)

    for i in range(num_steps):
        sigma, next_sigma = s_in * sigmas[i], s_in * sigmas[i + 1]
        # get_ancestral_step, sampling_utils.py:27-36
        sigma_up = torch.minimum(next_sigma, eta * (next_sigma ** 2 * (sigma ** 2 - next_sigma ** 2) / sigma ** 2) ** 0.5)
        sigma_down = (next_sigma ** 2 - sigma_up ** 2) ** 0.5
        denoised = denoise(x, sigma)
        ad = lambda v: v.reshape(v.shape + (1,) * (x.dim() - 1))
        # ancestral_euler_step (:176-180)
        d = (x - denoised) / ad(sigma)
        x_euler = x + d * ad(sigma_down - sigma)
        if torch.sum(sigma_down) < 1e-14:
            x = x_euler
        else:
            t, t_next = -sigma.log(), -sigma_down.log()
            h = t_next - t
            s = t + 0.5 * h
            # get_mult (:371-383)
            mult1 = (-s).exp() / (-t).exp()          # to_sigma(s)/to_sigma(t)
            mult2 = (-0.5 * h).expm1()
            mult3 = (-t_next).exp() / (-t).exp()
            mult4 = (-h).expm1()
            x2 = ad(mult1) * x - ad(mult2) * denoised
            denoised2 = denoise(x2, (-s).exp())
            x_dpmpp2s = ad(mult3) * x - ad(mult4) * denoised2
            x = torch.where(ad(sigma_down) > 0.0, x_dpmpp2s, x_euler)
        # ancestral_step (:182-188)
        x = torch.where(ad(next_sigma) > 0.0, x + noises[i] * s_noise * ad(sigma_up), x)
    return x

"""Network classes of the hot path: drop-in mirrors of the reference's
    ControlNet2D                  sgm/modules/diffusionmodules/controlmodel.py:195-317
    ControlledUNetModel3DTV2V     sgm/modules/diffusionmodules/controlmodel.py:320-553   (UNetModel3D openaimodel.py:1581-1639,
                                                                                          UNetModel.__init__ :1033-1527)
Same constructor kwargs as the YAML `params` (configs/inference_ccedit/*.yaml), same forward signatures, same
state-dict keys, same attributes other code touches (`.controlnet`, `.controlnet_img`,
`.input_hint_block[0].weight`, `.input_blocks_temporal[0].weight`).

Execution is a list of C-ABI kernel calls on channels-last fp16 buffers (see modules.py).  At the public boundary
tensors keep the reference's shapes: inputs are "b c t h w" (any strides, fp32 or fp16); the control tensors returned by
ControlNet2D.forward are [B, C, T, h, w]-shaped *views* of channels-last memory (torch's channels_last_3d), so the
reference's `control.pop()` / `c * scale` code keeps working while ControlledUNetModel3DTV2V consumes them without a
copy; the network output is a fresh contiguous [B, 4, T, h, w] tensor of x.dtype (controlmodel.py:546-550).
"""
from __future__ import annotations

import importlib
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from . import ops
from .modules import (GN_EPS_RES, Ctx, Downsample, Downsample3D, ParamHolder, ResBlock, ResBlock3D, SpatialTransformer,
                      SpatialTransformer3D, TimestepEmbedSequential, Upsample3D, conv1d, conv2d, linear, norm, seq)


def instantiate_from_config(config):
    """sgm/util.py:168-185 - the reference's plugin loader (YAML `target:` + `params:`)."""
    if "target" not in config:
        raise KeyError("Expected key `target` to instantiate.")
    module, cls = config["target"].rsplit(".", 1)
    return getattr(importlib.import_module(module), cls)(**dict(config.get("params", dict())))


# the reference's dotted paths resolve to this package's classes (a deployment may instead edit the YAML `target:`)
_TARGET_ALIASES = {
    "sgm.modules.diffusionmodules.controlmodel.ControlNet2D": "ccedit_b200.controlmodel.ControlNet2D",
    "sgm.modules.diffusionmodules.controlmodel.ControlledUNetModel3DTV2V":
        "ccedit_b200.controlmodel.ControlledUNetModel3DTV2V",
}


def _instantiate_aliased(config):
    cfg = dict(config)
    cfg["target"] = _TARGET_ALIASES.get(cfg["target"], cfg["target"])
    return instantiate_from_config(cfg)


class IndexedModules(nn.ModuleDict):
    """ModuleDict keyed by the integer positions the layers have in the reference's nn.Sequential; `m[0]` works."""

    def __getitem__(self, key):
        return super().__getitem__(str(key))


def _to_cl(t: torch.Tensor, cpad: int, mul: float = 1.0, add: float = 0.0) -> torch.Tensor:
    """[B, C, T, H, W] (any dtype/strides) -> channels-last fp16 [B, T, H, W, cpad] without a copy when possible."""
    if t.dtype == torch.float16 and t.shape[1] == cpad and mul == 1.0 and add == 0.0:
        cl = t.permute(0, 2, 3, 4, 1)
        if cl.is_contiguous():
            return cl
    return ops.ncthw_to_cl(t, cpad, mul, add)


def _as_ncthw(cl: torch.Tensor) -> torch.Tensor:
    """[B, T, H, W, C] channels-last buffer -> [B, C, T, H, W]-shaped view (or [F, C, H, W] for 4-D)."""
    return cl.permute(0, 4, 1, 2, 3) if cl.dim() == 5 else cl.permute(0, 3, 1, 2)


class UNetModel(nn.Module):
    """Encoder (+ optional decoder) structure of UNetModel.__init__, openaimodel.py:1033-1527, restricted to what
    the two inference configs use: use_spatial_transformer, transformer_depth 1, conv_resample, no class labels."""

    def __init__(self, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions, dropout=0,
                 channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None, use_checkpoint=False,
                 use_fp16=False, num_heads=-1, num_head_channels=-1, num_heads_upsample=-1, use_scale_shift_norm=False,
                 resblock_updown=False, use_new_attention_order=False, use_spatial_transformer=False,
                 transformer_depth=1, context_dim=None, n_embed=None, legacy=True, disable_self_attentions=None,
                 num_attention_blocks=None, disable_middle_self_attn=False, use_linear_in_transformer=False,
                 spatial_transformer_attn_type="softmax", adm_in_channels=None, use_fairscale_checkpoint=False,
                 offload_to_cpu=False, transformer_depth_middle=None, unet_type=None,
                 enable_attention3d_crossframe=False, disable_text_ca=False, disable_temporal_text_ca=False,
                 build_decoder=True, **kwargs):
        super().__init__()
        unsupported = dict(dims=(dims, 2), num_classes=(num_classes, None), use_scale_shift_norm=(use_scale_shift_norm, False),
                           resblock_updown=(resblock_updown, False), n_embed=(n_embed, None),
                           use_linear_in_transformer=(use_linear_in_transformer, False),
                           disable_self_attentions=(disable_self_attentions, None),
                           num_attention_blocks=(num_attention_blocks, None),
                           disable_middle_self_attn=(disable_middle_self_attn, False), dropout=(dropout, 0),
                           conv_resample=(conv_resample, True))
        for k, (got, want) in unsupported.items():
            if got != want:
                raise NotImplementedError(f"ccedit_b200: {k}={got!r} is outside the hot path (supported: {want!r})")
        if not use_spatial_transformer or context_dim is None:
            raise NotImplementedError("ccedit_b200: use_spatial_transformer=True with a context_dim is required")
        depth = transformer_depth if isinstance(transformer_depth, int) else max(transformer_depth)
        if depth != 1 or (transformer_depth_middle not in (None, 1)):
            raise NotImplementedError("ccedit_b200: transformer_depth must be 1")
        if num_heads == -1 and num_head_channels == -1:
            raise ValueError("Either num_heads or num_head_channels has to be set")
        if isinstance(context_dim, (list, tuple)) or type(context_dim).__name__ == "ListConfig":
            context_dim = list(context_dim)[0]
        pseudo3d = unet_type == "pseudo-3d"
        if unet_type not in (None, "2d", "pseudo-3d"):
            raise NotImplementedError(f"unet_type={unet_type}")
        if pseudo3d and not disable_temporal_text_ca:
            raise NotImplementedError("ccedit_b200: temporal blocks with text cross-attention are outside the hot path "
                                      "(both inference configs set disable_temporal_text_ca: True)")
        self.in_channels, self.model_channels, self.out_channels = in_channels, model_channels, out_channels
        self.num_res_blocks = ([num_res_blocks] * len(channel_mult) if isinstance(num_res_blocks, int)
                               else list(num_res_blocks))
        self.attention_resolutions = list(attention_resolutions)
        self.channel_mult = list(channel_mult)
        self.num_classes = None
        self.num_heads, self.num_head_channels = num_heads, num_head_channels
        self.context_dim = context_dim
        self.disable_text_ca = disable_text_ca
        self.pseudo3d = pseudo3d
        self.predict_codebook_ids = False
        self.dtype = torch.float16
        ca_type = kwargs.get("ST3DCA_ca_type", None) if (pseudo3d and enable_attention3d_crossframe) else None
        if pseudo3d and enable_attention3d_crossframe and ca_type is None:
            ca_type = "center"        # SpatialTransformer3DCA default (attention.py:1299)
        self.ca_type = ca_type

        def res(cin, cout):
            return (ResBlock3D if pseudo3d else ResBlock)(cin, ted, cout)

        def attn(ch, heads, dh):
            if pseudo3d:
                return SpatialTransformer3D(ch, heads, dh, context_dim, ca_type=ca_type)
            return SpatialTransformer(ch, heads, dh, context_dim, disable_text_ca=disable_text_ca)

        def head_split(ch):
            if num_head_channels == -1:
                heads, dh = num_heads, ch // num_heads
            else:
                heads, dh = ch // num_head_channels, num_head_channels
            if legacy:
                dh = ch // heads
            return heads, dh

        ted = model_channels * 4
        self.time_embed = seq(_0=linear(model_channels, ted), _2=linear(ted, ted))
        self.input_blocks = nn.ModuleList([TimestepEmbedSequential([conv2d(in_channels, model_channels, 3)])])
        chans = [model_channels]
        ch, ds = model_channels, 1
        for level, mult in enumerate(self.channel_mult):
            for _ in range(self.num_res_blocks[level]):
                layers = [res(ch, mult * model_channels)]
                ch = mult * model_channels
                if ds in self.attention_resolutions:
                    layers.append(attn(ch, *head_split(ch)))
                self.input_blocks.append(TimestepEmbedSequential(layers))
                chans.append(ch)
            if level != len(self.channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential([(Downsample3D if pseudo3d else Downsample)(ch)]))
                chans.append(ch)
                ds *= 2
        self.input_block_chans = list(chans)
        self.middle_block = TimestepEmbedSequential([res(ch, ch), attn(ch, *head_split(ch)), res(ch, ch)])
        self.middle_channels = ch
        if build_decoder:
            self.output_blocks = nn.ModuleList()
            for level, mult in list(enumerate(self.channel_mult))[::-1]:
                for i in range(self.num_res_blocks[level] + 1):
                    ich = chans.pop()
                    layers = [res(ch + ich, model_channels * mult)]
                    ch = model_channels * mult
                    if ds in self.attention_resolutions:
                        layers.append(attn(ch, *head_split(ch)))
                    if level and i == self.num_res_blocks[level]:
                        if not pseudo3d:
                            raise NotImplementedError("ccedit_b200: the 2-D decoder is outside the hot path")
                        layers.append(Upsample3D(ch))
                        ds //= 2
                    self.output_blocks.append(TimestepEmbedSequential(layers))
            self.out = seq(_0=norm(ch), _2=conv2d(model_channels, out_channels, 3, zero=True))
        self._emb_pack = None
        self._kv_pack = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_packed())

    # ---- packed, per-network fused small weights -----------------------------------------------------------------
    def _own_modules(self):
        """Sub-modules of this network, excluding nested networks (controlnet, controlnet_img)."""
        skip = [m for n, m in self.named_children() if isinstance(m, UNetModel)]
        skipped = set()
        for s in skip:
            skipped.update(id(x) for x in s.modules())
        return [m for m in self.modules() if id(m) not in skipped]

    def invalidate_packed(self):
        self._emb_pack = None
        self._kv_pack = None
        for m in self.modules():
            if m is not self and hasattr(m, "invalidate") and not isinstance(m, UNetModel):
                m.invalidate()
            if isinstance(m, UNetModel) and m is not self:
                m._emb_pack = m._kv_pack = None

    def _param_version(self, holders):
        """Version key over EVERY tensor a fused pack caches (weights and biases)."""
        return tuple((h.weight._version, h.weight.data_ptr()) + (() if h.bias is None else (h.bias._version, h.bias.data_ptr()))
                     for h in holders)

    def _prepare_ctx(self, timesteps: torch.Tensor, context: Optional[torch.Tensor], B: int, T: int, dev) -> Ctx:
        """Everything that depends only on (t, text): time embedding MLP (openaimodel.py:1216-1223), the SiLU+Linear of
        every ResBlock (:471-477) as ONE small linear over the row-concatenated weights, and the text K/V of every
        attn2 (attention.py:383-385) as ONE GEMM over the row-concatenated to_k/to_v."""
        if B > 8:
            raise RuntimeError("ccedit_b200: at most 8 batch entries (CFG included) per network call")
        ctx = Ctx(B, T, dev)
        mods = self._own_modules()
        resblocks = [m for m in mods if isinstance(m, ResBlock)]
        holders = [r.emb_layers["1"] for r in resblocks]
        ver = (dev, self._param_version(holders + [self.time_embed["0"], self.time_embed["2"]]))
        if self._emb_pack is None or self._emb_pack[0] != ver:
            w = torch.cat([h.weight.detach().float() for h in holders], 0).to(device=dev, dtype=torch.float16).contiguous()
            b = torch.cat([h.bias.detach().float() for h in holders], 0).to(device=dev).contiguous()
            te = self.time_embed
            te_w = [te[k].weight.detach().to(device=dev, dtype=torch.float16).contiguous() for k in ("0", "2")]
            te_b = [te[k].bias.detach().to(device=dev, dtype=torch.float32).contiguous() for k in ("0", "2")]
            self._emb_pack = (ver, w, b, te_w, te_b)
        _, w, b, te_w, te_b = self._emb_pack
        t_emb = ops.timestep_embedding(timesteps.to(dev), self.model_channels)
        e = ops.linear_small(t_emb, te_w[0], te_b[0], act_out=True)
        emb = ops.linear_small(e, te_w[1], te_b[1])
        rows = ops.linear_small(emb, w, b, act_in=True)        # [B, sum Cout] fp32
        off = 0
        for r in resblocks:
            ctx.emb_rows[id(r)] = rows[:, off:off + r.out_channels]
            off += r.out_channels
        attns = [a for m in mods if isinstance(m, SpatialTransformer) for a in m.text_attns()]
        if attns:
            if context is None:
                raise RuntimeError("ccedit_b200: context (crossattn) is required by the text cross-attention blocks")
            kvh = [h for a in attns for h in (a.to_k, a.to_v)]
            ver = (dev, self._param_version(kvh))
            if self._kv_pack is None or self._kv_pack[0] != ver:
                wkv = torch.cat([h.weight.detach().float() for h in kvh], 0)
                self._kv_pack = (ver, ops.pack_weight(wkv, None, dev))
            pw = self._kv_pack[1]
            c16 = ops.to_half(context.to(dev))                    # [B, 77, 768]
            Bc, Lc, Dc = c16.shape
            if Bc != B:
                raise RuntimeError(f"ccedit_b200: context batch {Bc} != x batch {B}")
            kv = ops.gemm(c16.view(Bc * Lc, Dc), pw, torch.empty(Bc * Lc, pw.n, dtype=torch.float16, device=dev))
            kv = kv.view(Bc, Lc, pw.n)
            off = 0
            for a in attns:
                ctx.text_kv[id(a)] = (kv[..., off:off + a.inner], kv[..., off + a.inner:off + 2 * a.inner])
                off += 2 * a.inner
        return ctx


# ---------------------------------------------------------------------------------------------------------------------
class ControlNet2D(UNetModel):
    """controlmodel.py:195-317: the frozen 2-D SD-1.5 encoder run on every frame + hint stem + 13 zero convs."""

    HINT_STRIDES = (1, 1, 2, 1, 2, 1, 2, 1)          # controlmodel.py:215-231
    HINT_WIDTHS = (16, 16, 32, 32, 96, 96, 256)

    def __init__(self, hint_channels, control_scales, no_add_x=False, set_input_hint_block_as_identity=False, *args,
                 **kwargs):
        kwargs["out_channels"] = kwargs["in_channels"]
        kwargs.pop("unet_type", None)
        super().__init__(*args, build_decoder=False, **kwargs)
        self.control_scales = control_scales
        self.no_add_x = no_add_x
        self.set_input_hint_block_as_identity = set_input_hint_block_as_identity
        mc = self.model_channels
        if set_input_hint_block_as_identity:
            self.input_hint_block = IndexedModules({"0": nn.Identity()})
        else:
            widths = (hint_channels,) + self.HINT_WIDTHS + (mc,)
            self.input_hint_block = IndexedModules({
                str(2 * i): conv2d(widths[i], widths[i + 1], 3, zero=(i == 7)) for i in range(8)})
        self.zero_convs = nn.ModuleList([TimestepEmbedSequential([conv2d(c, c, 1, zero=True)])
                                         for c in self.input_block_chans])
        self.middle_block_out = TimestepEmbedSequential([conv2d(self.middle_channels, self.middle_channels, 1, zero=True)])

    # ---- hint stem: 8 conv3x3 (+SiLU), three of them stride 2 ---------------------------------------------------
    def _hint_stem(self, hint_cl: torch.Tensor) -> torch.Tensor:
        g = hint_cl                                               # [F, H, W, 8]
        dev = g.device
        first = 0
        h0, h1 = self.input_hint_block[0], self.input_hint_block[2]
        if (g.shape[-1] == 8 and tuple(h0.weight.shape[:1]) == (16,) and h0.weight.shape[1] <= 8
                and tuple(h1.weight.shape[:2]) == (16, 16)):
            # the two full-resolution layers (3 -> 16 -> 16) in one pass over the hint video (csrc/hint_stem.cu)
            w0, b0 = h0._cached((dev, "hs0"), lambda: ops.pack_hint_stem_weight(h0.weight, h0.bias, dev, 8, 80))
            w1, b1 = h1._cached((dev, "hs1"), lambda: ops.pack_hint_stem_weight(h1.weight, h1.bias, dev, 16, 144))
            g = ops.hint_stem01(g, w0, b0, w1, b1)
            first = 2
            h2, h3 = self.input_hint_block[4], self.input_hint_block[6]
            if (tuple(self.HINT_STRIDES[2:4]) == (2, 1) and tuple(h2.weight.shape[:2]) == (32, 16)
                    and tuple(h3.weight.shape[:2]) == (32, 32) and g.shape[1] % 2 == 0 and g.shape[2] % 2 == 0):
                # layers 2 + 3 (16 -> 32 stride 2, 32 -> 32) the same way: no parity split, no 32-channel round trip
                w2, b2 = h2._cached((dev, "hs2"), lambda: ops.pack_hint_stem_weight(h2.weight, h2.bias, dev, 16, 144))
                w3, b3 = h3._cached((dev, "hs3"), lambda: ops.pack_hint_stem_weight(h3.weight, h3.bias, dev, 32, 288))
                g = ops.hint_stem23(g, w2, b2, w3, b3)
                first = 4
        for i, s in enumerate(self.HINT_STRIDES):
            if i < first:
                continue
            pw = self.input_hint_block[2 * i].packed(dev)
            F, H, W, _ = g.shape
            if s == 2:
                planes = ops.parity_split(g)
                out = torch.empty(F, H // 2, W // 2, pw.n, dtype=torch.float16, device=dev)
                ops.gemm(planes, pw, out.unsqueeze(1), ops.conv_s2_taps(), silu=(i < 7))
            else:
                out = torch.empty(F, H, W, pw.n, dtype=torch.float16, device=dev)
                ops.gemm(g, pw, out, ops.conv_taps(), silu=(i < 7))
            g = out
        return g

    def forward_cl(self, x_cl: Optional[torch.Tensor], hint_cl: torch.Tensor, timesteps, context, B: int, T: int,
                   cfg_dedup: bool = False, sinks=None) -> Optional[List[torch.Tensor]]:
        """x_cl: [B*T, h, w, 8] (ignored if no_add_x); hint_cl: [B*T, 8h, 8w, 8] (or [B*T, h, w, 8] latent features when
        the hint block is the identity).  Returns 13 channels-last tensors [B*T, h_l, w_l, C_l] (already * scale).
        cfg_dedup: B counts BOTH halves of a CFG batch (timesteps / context have B entries, uncond half first) but x_cl
        and hint_cl hold only the B/2 * T frames the halves share; the layers ahead of the first text cross-attention
        run once and fan out there (modules.SpatialTransformer.run_spatial).
        sinks (ControlledUNetModel3DTV2V.control_sinks): the ControlNet residual add fused into the skip write
        (controlmodel.py:536-543): zero conv j does not materialise control[j] but writes  zero_conv(h) * scale + skip
        straight into the UNet decoder's concat buffer (epilogue residual); nothing is returned."""
        dev = hint_cl.device
        ctx = self._prepare_ctx(timesteps, context, B, T, dev)
        if cfg_dedup and (self.disable_text_ca or len(self.input_blocks[1]) < 2):
            raise RuntimeError("ccedit_b200: CFG de-duplication needs a text cross-attention in input block 1")
        conv_in = self.input_blocks[0][0].packed(dev)
        if self.set_input_hint_block_as_identity:
            F, H, W, _ = hint_cl.shape
            guided = ops.gemm(hint_cl, conv_in, torch.empty(F, H, W, conv_in.n, dtype=torch.float16, device=dev),
                              ops.conv_taps())
        else:
            guided = self._hint_stem(hint_cl)
        outs = []
        h = None
        for i, (module, zc) in enumerate(zip(self.input_blocks, self.zero_convs)):
            if i == 0:
                if self.no_add_x:
                    h = guided
                else:
                    h = ops.gemm(x_cl, conv_in, torch.empty_like(guided), ops.conv_taps(), res1=guided)
            elif i == 1 and cfg_dedup:
                h = module.run(h, ctx, dup=True)
            else:
                h = module.run(h, ctx)
            if sinks is not None:
                self._zero_conv_into(zc[0], h, sinks[i])
                continue
            o = self._zero_conv(zc[0], h)
            outs.append(ops.dup_rows(o) if (i == 0 and cfg_dedup) else o)
        h = self.middle_block.run(h, ctx)
        if sinks is not None:
            self._zero_conv_into(self.middle_block_out[0], h, sinks[-1])
            return None
        outs.append(self._zero_conv(self.middle_block_out[0], h))
        return outs

    def _zero_conv_into(self, holder: ParamHolder, h: torch.Tensor, sink) -> None:
        """sink: [(dst, res)] - dst a channel slice of a concat buffer, res the UNet skip (same rows as h)."""
        from .modules import _as2d
        pw = holder.packed(h.device, scale=float(self.control_scales))
        F, H, W, C = h.shape
        M = F * H * W
        for dst, res in sink:
            if dst.numel() != M * pw.n or res.numel() != M * pw.n:
                raise RuntimeError("ccedit_b200: control sink does not match the ControlNet output "
                                   f"({tuple(dst.shape)} / {tuple(res.shape)} vs {(F, H, W, pw.n)})")
            ops.gemm(h.view(M, C), pw, _as2d(dst, M, pw.n), res1=res.reshape(M, pw.n))

    def _zero_conv(self, holder: ParamHolder, h: torch.Tensor) -> torch.Tensor:
        pw = holder.packed(h.device, scale=float(self.control_scales))
        F, H, W, C = h.shape
        return ops.gemm(h.view(F * H * W, C), pw, torch.empty(F * H * W, pw.n, dtype=torch.float16, device=h.device)
                        ).view(F, H, W, pw.n)

    def forward(self, x, hint, timesteps=None, context=None, y=None, **kwargs):
        assert y is None, "must specify y if and only if the model is class-conditional"
        is_video = x.dim() == 5
        if not is_video:
            x, hint = x.unsqueeze(2), hint.unsqueeze(2)
        B, _, T = x.shape[:3]
        dev = x.device
        x_cl = None if self.no_add_x else _to_cl(x, 8).view(B * T, x.shape[3], x.shape[4], 8)
        hint_cl = _to_cl(hint, 8).view(B * T, hint.shape[3], hint.shape[4], 8)
        outs = self.forward_cl(x_cl, hint_cl, timesteps, context, B, T)
        if is_video:
            return [_as_ncthw(o.view(B, T, *o.shape[1:])) for o in outs]
        return [_as_ncthw(o) for o in outs]


# ---------------------------------------------------------------------------------------------------------------------
class ControlledUNetModel3DTV2V(UNetModel):
    """controlmodel.py:320-553 on top of UNetModel3D (openaimodel.py:1581-1639): the pseudo-3D UNet whose skips and
    middle receive the ControlNet residuals; (tvi2v) the centre frame receives the reference-image residuals."""

    def __init__(self, controlnet_config, *args, temporal_kernel_size=None, offload_to_cpu=False, n_embed=None,
                 use_learnable_alpha=False, **kwargs):
        if temporal_kernel_size not in (None, 3) or use_learnable_alpha or kwargs.get("crossframe_type") is not None:
            raise NotImplementedError("ccedit_b200: temporal_kernel_size != 3 / learnable alpha / crossframe_type "
                                      "are outside the hot path")
        kwargs["unet_type"] = kwargs.get("unet_type", "pseudo-3d")
        controlnet_img_config = kwargs.pop("controlnet_img_config", None)
        super().__init__(*args, **kwargs)
        mc, oc = self.model_channels, self.out_channels
        self.temporal_kernel_size = 3
        self.input_blocks_temporal = TimestepEmbedSequential([conv1d(mc, mc, 3, zero=True)])
        self.out_temporal = seq(_1=conv1d(oc, oc, 3, zero=True))
        self.controlnet = _instantiate_aliased(controlnet_config)
        if controlnet_img_config is not None:
            self.controlnet_img = _instantiate_aliased(controlnet_img_config)
        self._tail = None

    def _tail_params(self, dev):
        h = self.out_temporal["1"]
        ver = (dev, h.weight._version, h.bias._version)
        if self._tail is None or self._tail[0] != ver:
            self._tail = (ver, h.weight.detach().to(device=dev, dtype=torch.float32).contiguous(),
                          h.bias.detach().to(device=dev, dtype=torch.float32).contiguous())
        return self._tail[1], self._tail[2]

    def invalidate_packed(self):
        super().invalidate_packed()
        self._tail = None

    def forward_cl(self, x_cl: torch.Tensor, timesteps, context, control: Optional[List[torch.Tensor]],
                   img_control: Optional[List[torch.Tensor]], only_mid_control: bool, out_dtype,
                   cfg_dedup: bool = False) -> torch.Tensor:
        """x_cl: [B, T, h, w, 8]; control: 13 x [B*T, h_l, w_l, C_l] or None; img_control: 13 x [B, h_l, w_l, C_l].
        cfg_dedup: x_cl (and img_control) hold ONE copy of the B latents that the uncond and cond halves of a CFG batch
        share, timesteps / context / control have 2 B entries (uncond first); input blocks 0 and 1 run once up to the
        first text cross-attention and fan out there; the result has 2 B entries."""
        hs, h, ctx = self.encode_cl(x_cl, timesteps, context, img_control, only_mid_control, cfg_dedup)
        if cfg_dedup:
            hs[0] = ops.dup_rows(hs[0])
        cats = self.decoder_buffers(hs, h)
        # decoder input: h = cat([h, hs.pop() + control.pop()], 1) (controlmodel.py:536-544); control = None: plain copies
        control = None if control is None else list(control)
        ops.add_rows(h, control.pop() if control is not None else None, cats[0][..., :h.shape[-1]])
        for i in range(len(self.output_blocks)):
            skip = hs[len(hs) - 1 - i]
            sc = control.pop() if (control is not None and not only_mid_control) else None
            ops.add_rows(skip, sc, cats[i][..., cats[i].shape[-1] - skip.shape[-1]:])
        return self.decode_cl(cats, ctx, out_dtype)

    def encode_cl(self, x_cl, timesteps, context, img_control, only_mid_control: bool, cfg_dedup: bool):
        """Input blocks + middle block.  Returns (hs: 12 skip tensors [B, T, h_l, w_l, C_l], h: middle output, ctx).  With
        cfg_dedup hs[0] (the output of input block 0, ahead of the fan-out) has B entries, everything else 2 B."""
        dev = x_cl.device
        B, T, H, W, _ = x_cl.shape
        ctx = self._prepare_ctx(timesteps, context, 2 * B if cfg_dedup else B, T, dev)
        img_control = None if img_control is None else list(img_control)
        mc = self.model_channels
        hs = []
        h = None
        for i, module in enumerate(self.input_blocks):
            if i == 0:
                # spatial_temporal_forward(conv3x3, input_blocks_temporal), controlmodel.py:523-526
                y = ops.gemm(x_cl.view(B * T, H, W, 8), module[0].packed(dev),
                             torch.empty(B * T, H, W, mc, dtype=torch.float16, device=dev), ops.conv_taps())
                y4 = y.view(B, T, H * W, mc)
                h = ops.gemm(y4, self.input_blocks_temporal[0].packed(dev), torch.empty_like(y4), ops.temporal_taps(3),
                             res1=y4).view(B, T, H, W, mc)
            elif i == 1 and cfg_dedup:
                h = module.run(h, ctx, dup=True)                    # [2B, T, H, W, C] from here on
                if img_control is not None:
                    img_control = [ops.dup_rows(ic.contiguous()) for ic in img_control]
            else:
                h = module.run(h, ctx)
            if img_control is not None and not only_mid_control:
                ops.add_center_frame(h, img_control.pop(0))
            hs.append(h)
        h = self.middle_block.run(h, ctx)
        if img_control is not None:
            ops.add_center_frame(h, img_control.pop(0))
        return hs, h, ctx

    def decoder_buffers(self, hs, h_mid) -> List[torch.Tensor]:
        """The 12 concatenated decoder inputs cat([h, skip + control], 1) (controlmodel.py:541-543), allocated up front:
        cats[i] = [B, T, h_i, w_i, Ch_i + Cs_i] is the input of output block i; its first Ch_i channels are written by the
        producer of h (the previous output block, or middle + control for i = 0), the last Cs_i by whoever adds skip and
        ControlNet residual (ops.add_rows, or the ControlNet's zero-conv GEMM itself: ControlNet2D.forward_cl(sinks=))."""
        cats = []
        B, T = h_mid.shape[0], h_mid.shape[1]
        ch, hh, ww = h_mid.shape[-1], h_mid.shape[2], h_mid.shape[3]
        for i, module in enumerate(self.output_blocks):
            cs = hs[len(hs) - 1 - i].shape[-1]
            cats.append(torch.empty(B, T, hh, ww, ch + cs, dtype=torch.float16, device=h_mid.device))
            ch = module.out_channels
            if module.upsamples:
                hh, ww = 2 * hh, 2 * ww
        return cats

    def control_sinks(self, hs, h_mid, cats):
        """(destination, residual) per ControlNet output, in the ControlNet's order (12 zero convs, then middle_block_out):
        zero conv j lands in the skip half of cats[11 - j] on top of hs[j]; the middle one in the h half of cats[0] on top of
        the middle block's output.  With a de-duplicated hs[0] (B entries against 2 B in cats[11]) the first sink is a pair."""
        n = len(hs)
        sinks = []
        for j in range(n):
            cat = cats[n - 1 - j]
            dst = cat[..., cat.shape[-1] - hs[j].shape[-1]:]
            if hs[j].shape[0] != cat.shape[0]:                      # fan-out of the shared block-0 output
                Bh = hs[j].shape[0]
                sinks.append([(dst[:Bh], hs[j]), (dst[Bh:], hs[j])])
            else:
                sinks.append([(dst, hs[j])])
        sinks.append([(cats[0][..., :h_mid.shape[-1]], h_mid)])
        return sinks

    def decode_cl(self, cats, ctx, out_dtype) -> torch.Tensor:
        """Output blocks on the filled concat buffers + out head.  The concatenation is never a copy: every block writes
        its result straight into channels [0, Ch) of the next block's input buffer."""
        dev = cats[0].device
        B, T = cats[0].shape[0], cats[0].shape[1]
        n_out = len(self.output_blocks)
        h = None
        for i, module in enumerate(self.output_blocks):
            if i + 1 < n_out:
                module.run(cats[i], ctx, out=cats[i + 1][..., :module.out_channels])
            else:
                h = module.run(cats[i], ctx)
        # out: GN + SiLU + conv3x3 then y + conv1d_k3(SiLU(y)) over T (openaimodel.py:1513-1519, 1627-1632)
        Hh, Ww = h.shape[2], h.shape[3]
        a = ops.groupnorm_spatial(h.view(B * T, Hh, Ww, h.shape[-1]), *self.out["0"].affine(dev), GN_EPS_RES, True)
        pw = self.out["2"].packed(dev)
        yo = ops.gemm(a, pw, torch.empty(B * T, Hh, Ww, pw.n, dtype=torch.float16, device=dev), ops.conv_taps())
        wt, bt = self._tail_params(dev)
        return ops.out_temporal(yo.view(B, T, Hh, Ww, pw.n), wt, bt, self.out_channels, out_dtype)

    @staticmethod
    def _control_to_cl(c: torch.Tensor, frames_5d: bool) -> torch.Tensor:
        """API tensor [B, C, T, h, w] (or [B, C, h, w]) -> channels-last fp16, zero-copy for ControlNet2D's own outputs."""
        if c.dim() == 5:
            cl = _to_cl(c, c.shape[1])
            return cl.reshape(cl.shape[0] * cl.shape[1], *cl.shape[2:]) if frames_5d else cl
        cl = c.permute(0, 2, 3, 1)
        if c.dtype == torch.float16 and cl.is_contiguous():
            return cl
        return _to_cl(c.unsqueeze(2), c.shape[1])[:, 0]

    def forward(self, x, timesteps=None, context=None, y=None, control=None, img_control=None,
                only_mid_control=False, **kwargs):
        assert y is None, "must specify y if and only if the model is class-conditional"
        if x.dim() != 5:
            raise RuntimeError("ccedit_b200: ControlledUNetModel3DTV2V expects x as [B, C, T, h, w]")
        x_cl = _to_cl(x, 8)
        ctrl = None
        if control is not None:
            ctrl = [self._control_to_cl(c, True) for c in control]
            del control[:]        # the reference consumes the list with pop() (controlmodel.py:536-543)
        ictrl = None
        if img_control is not None:
            ictrl = [self._control_to_cl(c, False) for c in img_control]
            del img_control[:]
        return self.forward_cl(x_cl, timesteps, context, ctrl, ictrl, only_mid_control, x.dtype)

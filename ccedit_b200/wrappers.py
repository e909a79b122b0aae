"""Drop-in boundary of the hot path: mirror of OpenAIWrapperControlLDM3DTV2V (sgm/modules/diffusionmodules/
wrappers.py:13-25, 155-207).

    wrapper(x [B',4,T,h,w], t int64 [B'], c {"crossattn", "control_hint", ("cond_feat")}) -> eps [B',4,T,h,w]

Same call as the reference (`model.denoiser(model.model, input, sigma, c)` reaches it through DiscreteDenoiser,
scripts/sampling/sampling_tv2v.py:366-369).  Internally the ControlNet(s) and the UNet exchange channels-last fp16
buffers directly; with `use_cuda_graph` the whole network call (~1.6 k kernel launches, static shapes per clip) is
captured once per input signature and replayed, which removes the per-launch host cost from the step loop.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import ops
from .controlmodel import _to_cl


class IdentityWrapper(nn.Module):
    """wrappers.py:13-25 (compile_model is accepted and ignored: there is no tracing compiler on this path)."""

    def __init__(self, diffusion_model, compile_model: bool = False):
        super().__init__()
        self.diffusion_model = diffusion_model

    def forward(self, *args, **kwargs):
        return self.diffusion_model(*args, **kwargs)


class OpenAIWrapperControlLDM3DTV2V(IdentityWrapper):
    def __init__(self, diffusion_model, compile_model: bool = False, use_cuda_graph: Optional[bool] = None):
        super().__init__(diffusion_model, compile_model)
        if use_cuda_graph is None:
            use_cuda_graph = os.environ.get("CCEDIT_CUDA_GRAPH", "1") != "0"
        self.use_cuda_graph = use_cuda_graph
        self._graphs: Dict[tuple, dict] = {}
        self._params = None
        self._generation = 0            # bumped by invalidate(): step graphs captured by a sampler watch it too

    # ---- one network call, eager: a flat list of C-ABI kernel launches on the current stream ---------------------
    def _network(self, x, t, crossattn, control_hint, cond_feat):
        net = self.diffusion_model
        B, _, T, h, w = x.shape
        x_cl = _to_cl(x, 8)                                                       # [B, T, h, w, 8]
        # hint: 1 - (hint + 1) / 2  (wrappers.py:160-162) folded into the layout change
        hint_cl = ops.ncthw_to_cl(control_hint, 8, pre=1.0, mul=-0.5, add=1.0)    # [B, T, 8h, 8w, 8]
        img_control = None
        if cond_feat is not None:
            feat_cl = _to_cl(cond_feat.unsqueeze(2), 8)                           # [B, 1, h, w, 8]
            # wrappers.py:180-186: controlnet_img sees the centre frame of x (ignored when it was built with no_add_x)
            xc_cl = None if net.controlnet_img.no_add_x else x_cl[:, T // 2].contiguous()
            img_control = net.controlnet_img.forward_cl(xc_cl, feat_cl.view(B, h, w, 8), t, crossattn, B, 1)
        # UNet encoder first, then the ControlNet: its zero convs add their residual to the encoder's skips and write the
        # sums straight into the decoder's concat buffers (controlmodel.py:536-543 without the 13 control tensors)
        hs, hm, ctx = net.encode_cl(x_cl, t, crossattn, img_control, False, False)
        cats = net.decoder_buffers(hs, hm)
        net.controlnet.forward_cl(x_cl.view(B * T, h, w, 8), hint_cl.view(B * T, hint_cl.shape[2], hint_cl.shape[3], 8),
                                  t, crossattn, B, T, sinks=net.control_sinks(hs, hm, cats))
        return net.decode_cl(cats, ctx, x.dtype)

    def forward_cfg(self, x: torch.Tensor, t: torch.Tensor, c: dict) -> torch.Tensor:
        """One network call for a classifier-free-guidance batch whose two halves share everything but the text:
        x [B, 4, T, h, w] and t [B] are given ONCE, c["crossattn"] has 2 B entries (uncond first, guiders.py:56-67),
        c["control_hint"] / c["cond_feat"] may have B or 2 B entries (the first B are used: the caller guarantees that
        the halves are equal, as GeneralConditioner.get_unconditional_conditioning makes them).  Equivalent to
        forward(cat([x] * 2), cat([t] * 2), c) -> eps [2 B, 4, T, h, w], but the layers ahead of the first text
        cross-attention of the ControlNet and of the UNet (hint stem, input blocks 0 and 1 up to the self-attention:
        ~7 % of a call's time at the headline shape) are computed once instead of twice.  Launches on the current
        stream; capturable (the fused sampler step records it inside its step graph)."""
        if not x.is_cuda:
            raise RuntimeError("ccedit_b200: the network runs on CUDA (sm_100a) only; there is no CPU fallback")
        net = self.diffusion_model
        B, _, T, h, w = x.shape
        crossattn, hint, feat = c["crossattn"], c["control_hint"][:B], c.get("cond_feat", None)
        if crossattn.shape[0] != 2 * B:
            raise RuntimeError(f"ccedit_b200.forward_cfg: crossattn must have 2 * {B} entries (uncond first)")
        with torch.no_grad():
            t2 = torch.cat([t] * 2)
            x_cl = _to_cl(x, 8)
            hint_cl = ops.ncthw_to_cl(hint, 8, pre=1.0, mul=-0.5, add=1.0)
            img_control = None
            if feat is not None:
                if not net.controlnet_img.disable_text_ca:
                    raise NotImplementedError("ccedit_b200.forward_cfg: controlnet_img with text cross-attention")
                # controlnet_img has no text cross-attention (disable_text_ca): identical for both halves, computed once
                feat_cl = _to_cl(feat[:B].unsqueeze(2), 8)
                xc_cl = None if net.controlnet_img.no_add_x else x_cl[:, T // 2].contiguous()
                img_control = net.controlnet_img.forward_cl(xc_cl, feat_cl.view(B, h, w, 8), t, None, B, 1)
            hs, hm, ctx = net.encode_cl(x_cl, t2, crossattn, img_control, False, True)
            cats = net.decoder_buffers(hs, hm)
            net.controlnet.forward_cl(x_cl.view(B * T, h, w, 8), hint_cl.view(B * T, hint_cl.shape[2], hint_cl.shape[3], 8),
                                      t2, crossattn, 2 * B, T, cfg_dedup=True, sinks=net.control_sinks(hs, hm, cats))
            return net.decode_cl(cats, ctx, x.dtype)

    def _graphed(self, x, t, crossattn, control_hint, cond_feat):
        ins = dict(x=x, t=t, crossattn=crossattn, control_hint=control_hint, cond_feat=cond_feat)
        ins = {k: v for k, v in ins.items() if v is not None}
        key = tuple((k, tuple(v.shape), v.dtype, v.device) for k, v in ins.items())
        ent = self._graphs.get(key)
        wver = self._weights_version()
        if ent is not None and ent["wver"] != wver:      # weights changed (load_state_dict / in-place LoRA merge)
            self._graphs.clear()
            ent = None
        if ent is None:
            static = {k: torch.empty_like(v, memory_format=torch.contiguous_format) for k, v in ins.items()}
            for k, v in ins.items():
                static[k].copy_(v)
            args = (static["x"], static["t"], static.get("crossattn"), static["control_hint"], static.get("cond_feat"))
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                # warm-up: packs weights, sets kernel attributes
                self._network(*args)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            n0 = ops.launch_count()
            with torch.cuda.graph(graph):
                out = self._network(*args)
            ent = dict(graph=graph, static=static, out=out, launches=ops.launch_count() - n0, wver=wver)
            self._graphs[key] = ent
        else:
            for k, v in ins.items():
                ent["static"][k].copy_(v)
        ent["graph"].replay()
        ops.note_graph_replay(ent["launches"])
        return ent["out"].clone()

    def forward(self, x: torch.Tensor, t: torch.Tensor, c: dict, **kwargs) -> torch.Tensor:
        concat = c.get("concat", None)
        if concat is not None and concat.numel():
            x = torch.cat((x, concat), dim=1)
        if c.get("vector", None) is not None:
            raise NotImplementedError("ccedit_b200: class/vector conditioning is outside the hot path")
        if not x.is_cuda:
            raise RuntimeError("ccedit_b200: the network runs on CUDA (sm_100a) only; there is no CPU fallback")
        args = (x, t, c.get("crossattn", None), c["control_hint"], c.get("cond_feat", None))
        with torch.no_grad():
            if self.use_cuda_graph and not torch.cuda.is_current_stream_capturing():
                return self._graphed(*args)
            return self._network(*args)

    def _weights_version(self):
        """Cheap staleness probe for the captured graphs: the autograd version counters of all parameters (bumped by
        load_state_dict, `p.add_()`, `state_dict()[k] += ...` - the reference's LoRA merge, sampling_tv2v.py:211-234) plus
        the storage of the first few.  Edits through `p.data` (`p.data.copy_()`, EMA swaps) do NOT bump the counters:
        call `invalidate()` after those."""
        if self._params is None:
            self._params = list(self.diffusion_model.parameters())
        return (sum(p._version for p in self._params) + sum(p.data_ptr() for p in self._params[:8]), self._generation)

    def reset_graphs(self):
        """Drop captured graphs (keeps the packed fp16 weights)."""
        self._graphs.clear()

    def invalidate(self):
        """Call after ANY weight edit the version counters cannot see (`p.data` writes, EMA swaps, re-assigned
        parameters): drops every packed fp16 kernel copy, the fused embedding / text-K/V packs and the captured CUDA
        graphs; the next call re-packs and re-captures.  ccedit_b200.checkpoint.merge_lora / load_network_state_dict
        call it themselves."""
        self._graphs.clear()
        self._params = None
        self._generation += 1
        for m in self.diffusion_model.modules():
            if hasattr(m, "invalidate_packed"):
                m.invalidate_packed()

"""GPU tests of the drop-in boundary through the PUBLIC interface an unmodified reference wrapper would use
(sgm/modules/diffusionmodules/wrappers.py:155-207): ControlNet2D.forward -> python list of [B, C, T, h, w] tensors ->
ControlledUNetModel3DTV2V.forward(x, timesteps, context, control=list, img_control=list), the pop()-style consumption of
those lists (controlmodel.py:529-543), checkpoint ingestion under engine-level keys (scripts/sampling/util.py:45-112) and
weight edits (LoRA merge, sampling_tv2v.py:211-234) under CUDA-graph replay."""
import pytest
import torch

from conftest import NET_TOL, load_golden, rel_err
from oracle import inputs as oin

pytestmark = pytest.mark.gpu


def _cuda(d):
    return {k: v.cuda() for k, v in d.items()}


def _inputs(kind, g):
    B, T, h, w = g["shape"]
    c, uc = oin.synthetic_cond(B, T, h, w, seed=3, tvi2v=(kind == "tvi2v"))
    xin, tin, cc = oin.cfg_batch(oin.synthetic_latent(B, T, h, w, seed=2), torch.tensor([g["t"]]), c, uc)
    return xin.cuda(), tin.cuda(), _cuda(cc)


@pytest.mark.parametrize("kind", ["tv2v", "tvi2v"])
def test_public_forward_chain_equals_wrapper(kind, gpu_wrappers):
    """The reference's own wrapper body (wrappers.py:160-207), line by line, on this package's public classes."""
    g = load_golden(f"network_{kind}.pt")
    x, t, cc = _inputs(kind, g)
    wrap = gpu_wrappers(kind, graph=False)
    net = wrap.diffusion_model
    with torch.no_grad():
        fast = wrap(x, t, cc)
        hint = 1 - (cc["control_hint"] + 1) / 2.0                                    # wrappers.py:160-162
        control = net.controlnet(x=x, hint=hint, timesteps=t, context=cc["crossattn"])
        assert isinstance(control, list) and len(control) == 13
        assert all(c.dim() == 5 and c.shape[0] == x.shape[0] and c.shape[2] == x.shape[2] for c in control)
        img_control = None
        if kind == "tvi2v":
            xc = x[:, :, x.shape[2] // 2, :, :]                                      # wrappers.py:181
            img_control = net.controlnet_img(x=xc, hint=cc["cond_feat"], timesteps=t, context=cc["crossattn"])
            assert len(img_control) == 13 and all(c.dim() == 4 for c in img_control)
        out = net(x=x, timesteps=t, context=cc["crossattn"], control=control, img_control=img_control,
                  only_mid_control=False)
    assert control == [] and (img_control is None or img_control == [])            # consumed like control.pop()
    assert out.dtype == x.dtype and out.is_contiguous()
    # Same kernels on the same data up to one point: the public path materialises the 13 control tensors in fp16 and adds
    # them to the skips afterwards (two roundings), the wrapper's fast path adds the skip inside the zero conv's fp32
    # epilogue (one rounding).  A few fp16 ulps on 13 tensors, carried through the decoder: measured 1.7e-3 of max|out|,
    # the same size as either path's distance to the fp32 reference.
    assert rel_err(out, fast) < NET_TOL
    assert rel_err(out, g["output"]) < NET_TOL


def test_public_forward_accepts_scaled_and_fp32_control(gpu_wrappers):
    """The reference multiplies / re-materialises the control tensors in Python (`[c * scale for c in control]`,
    controlmodel.py:316); fresh fp32 [B, C, T, h, w] tensors must be accepted as well as the zero-copy views."""
    g = load_golden("network_tv2v.pt")
    x, t, cc = _inputs("tv2v", g)
    net = gpu_wrappers("tv2v", graph=False).diffusion_model
    with torch.no_grad():
        hint = 1 - (cc["control_hint"] + 1) / 2.0
        control = net.controlnet(x=x, hint=hint, timesteps=t, context=cc["crossattn"])
        ref = net(x=x, timesteps=t, context=cc["crossattn"], control=list(control))
        fresh = [c.float().contiguous() * 1.0 for c in control]
        out = net(x=x, timesteps=t, context=cc["crossattn"], control=fresh)
        mid_only = net(x=x, timesteps=t, context=cc["crossattn"], control=list(control), only_mid_control=True)
    assert torch.equal(out, ref)
    assert not torch.equal(mid_only, ref)


def test_checkpoint_with_engine_prefix_loads_and_matches(gpu_wrappers, state_dicts):
    """A CCEdit checkpoint stores the network under `model.diffusion_model.*` next to VAE / conditioner tensors."""
    from ccedit_b200 import checkpoint as ck
    from ccedit_b200.configs import build_network
    g = load_golden("network_tv2v.pt")
    x, t, cc = _inputs("tv2v", g)
    sd = {"model." + k: v for k, v in state_dicts("tv2v").items()}
    sd["first_stage_model.decoder.conv_in.weight"] = torch.zeros(1)
    sd["conditioner.embedders.0.transformer.text_model.embeddings.position_ids"] = torch.zeros(1)
    wrap = build_network("tv2v", device="cpu", use_cuda_graph=False)
    missing, unexpected = ck.model_load_ckpt(wrap, {"state_dict": sd}["state_dict"])
    assert missing == [] and unexpected == []
    out = wrap.cuda()(x, t, cc)
    assert torch.equal(out, gpu_wrappers("tv2v", graph=False)(x, t, cc))


def test_weight_edits_take_effect_under_graph_replay(gpu_wrappers):
    """LoRA-style in-place merges must change the next output even though the call is a CUDA-graph replay of kernels
    reading packed fp16 copies: tracked edits (`p.add_`, `state_dict()[k] += d`) are picked up automatically, `p.data`
    edits after `invalidate()`; undoing the edit restores the original output bit for bit."""
    from ccedit_b200 import checkpoint as ck
    g = load_golden("network_tv2v.pt")
    x, t, cc = _inputs("tv2v", g)
    wrap = gpu_wrappers("tv2v", graph=True)
    base = wrap(x, t, cc)
    assert torch.equal(wrap(x, t, cc), base)                                       # replay
    key = "diffusion_model.input_blocks.1.1.transformer_blocks.0.attn1.to_q.weight"
    p = dict(wrap.named_parameters())[key]
    orig = p.detach().clone()
    delta = 0.05 * torch.randn(p.shape, generator=torch.Generator().manual_seed(5)).to(p.device)
    try:
        # (1) the reference's merge idiom: in place on the state-dict view, then load_state_dict (sampling_tv2v.py:213-234)
        sd = wrap.state_dict()
        sd[key] += delta
        wrap.load_state_dict(sd)
        out1 = wrap(x, t, cc)
        assert not torch.equal(out1, base)
        # (2) an edit the version counters cannot see, followed by the documented invalidate()
        p.data.copy_(orig)
        wrap.invalidate()
        assert torch.equal(wrap(x, t, cc), base)
        # (3) kohya-style LoRA pair through the merge helper; the merged weight equals W + alpha * up @ down
        r = 4
        gen = torch.Generator().manual_seed(6)
        lora = {"lora_unet_down_blocks_0_attentions_0_transformer_blocks_0_attn1_to_q.lora_down.weight":
                torch.randn(r, p.shape[1], generator=gen) * 0.05,
                "lora_unet_down_blocks_0_attentions_0_transformer_blocks_0_attn1_to_q.lora_up.weight":
                torch.randn(p.shape[0], r, generator=gen) * 0.05}
        before = p.detach().clone()
        touched = ck.merge_lora(wrap, lora, alpha=0.8)
        assert touched == ["model." + key]
        up, down = [v.to(p.device) for k, v in sorted(lora.items(), reverse=True)]
        assert torch.allclose(p.detach(), before + 0.8 * up @ down, atol=1e-6)
        out3 = wrap(x, t, cc)
        assert not torch.equal(out3, base)
        wrap.use_cuda_graph = False                                                  # eager == graph on the merged weights
        assert torch.equal(wrap(x, t, cc), out3)
    finally:
        p.data.copy_(orig)
        wrap.use_cuda_graph = True
        wrap.invalidate()
    assert torch.equal(wrap(x, t, cc), base)


def test_weight_edits_reach_the_fused_sampler_step_graph(gpu_wrappers):
    """The fused sampler replays ONE captured graph per step with the network calls inside; a LoRA merge between two
    clips must drop it (the plan watches the wrapper's weight version / invalidate() generation)."""
    from ccedit_b200 import checkpoint as ck
    from ccedit_b200.sampling import BoundDenoiser, DiscreteDenoiser, FusedDPMPP2SAncestralSampler
    g = load_golden("sampler_tv2v.pt")
    B, T, h, w = g["shape"]
    c, uc = oin.synthetic_cond(B, T, h, w, seed=7)
    x0 = oin.synthetic_latent(B, T, h, w, seed=6)
    wrap = gpu_wrappers("tv2v", graph=False)
    sampler = FusedDPMPP2SAncestralSampler(num_steps=2, device="cuda", eta=0.0, s_noise=1.0, guider_config={
        "target": "sgm.modules.diffusionmodules.guiders.VanillaCFGTV2V", "params": {"scale": 7.5}}, use_cuda_graph=True)
    den = BoundDenoiser(DiscreteDenoiser().cuda(), wrap)
    cd, ucd = _cuda(c), _cuda(uc)
    base = sampler(den, x0.clone().cuda(), cd, uc=ucd)
    assert torch.equal(sampler(den, x0.clone().cuda(), cd, uc=ucd), base)         # same plan, graph replay
    key = "diffusion_model.output_blocks.5.1.transformer_blocks.0.attn2.to_q.weight"
    p = dict(wrap.named_parameters())[key]
    orig = p.detach().clone()
    try:
        lora = {"lora_unet_up_blocks_1_attentions_2_transformer_blocks_0_attn2_to_q.lora_down.weight": torch.randn(4, p.shape[1]) * 0.1,
                "lora_unet_up_blocks_1_attentions_2_transformer_blocks_0_attn2_to_q.lora_up.weight": torch.randn(p.shape[0], 4) * 0.1}
        assert ck.merge_lora(wrap, lora, alpha=0.8) == ["model." + key]
        merged = sampler(den, x0.clone().cuda(), cd, uc=ucd)
        assert not torch.equal(merged, base)
    finally:
        p.data.copy_(orig)
        wrap.invalidate()
    assert torch.equal(sampler(den, x0.clone().cuda(), cd, uc=ucd), base)

"""CPU tests of the host side: state-dict layout vs the reference, weight packing layouts (emulating the tap-GEMM in
torch against the oracle's convs), tap tables, tile-box picking, the caller-side sampler classes against the oracle,
the FLOP census against FlopCounterMode on the oracle, and loud failure of the product path without CUDA."""
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, rel_err
from ccedit_b200 import ops
from oracle import sgm_oracle as so


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["tv2v", "tvi2v"])
def test_state_dict_layout_matches_reference(kind, manifests):
    """Same keys, shapes and zero/ones/default initialisation classes as the reference network (Appendix D)."""
    from oracle.ref_import import yaml_params
    from ccedit_b200.controlmodel import ControlledUNetModel3DTV2V
    from ccedit_b200.wrappers import OpenAIWrapperControlLDM3DTV2V
    net = ControlledUNetModel3DTV2V(**yaml_params(kind))
    sd = OpenAIWrapperControlLDM3DTV2V(net).state_dict()
    man = manifests[kind]
    assert set(sd) == set(man)
    for k, (shape, init) in man.items():
        v = sd[k]
        assert list(v.shape) == shape, k
        got = "zeros" if bool((v == 0).all()) else ("ones" if bool((v == 1).all()) else "default")
        assert got == init, (k, got, init)
    # attributes other reference code touches (wrappers.py:164, controlmodel.py:510)
    assert net.input_blocks_temporal[0].weight.dtype == torch.float32
    if kind == "tv2v":
        assert net.controlnet.input_hint_block[0].weight.shape == (16, 3, 3, 3)
    else:
        assert hasattr(net, "controlnet_img") and net.ca_type == "center_self"


def test_first_stage_state_dict_layout_matches_reference():
    """ccedit_b200.autoencoder keeps the reference's first-stage keys and shapes (tests/golden/manifest_vae.json was
    written from the reference's Encoder / Decoder / quant convs by oracle/make_golden.py vae)."""
    from oracle.vae_oracle import DDCONFIG
    from oracle.weights import load_manifest
    from ccedit_b200.autoencoder import AutoencoderKLInferenceWrapper
    sd = AutoencoderKLInferenceWrapper(ddconfig=dict(DDCONFIG), embed_dim=4).state_dict()
    man = load_manifest("vae")
    assert set(sd) == set(man)
    for k, (shape, _) in man.items():
        assert list(sd[k].shape) == shape, k
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        AutoencoderKLInferenceWrapper(ddconfig=dict(DDCONFIG), embed_dim=4).decode(torch.zeros(1, 4, 8, 8))


def test_unsupported_configs_fail_loudly():
    from oracle.ref_import import yaml_params
    from ccedit_b200.controlmodel import ControlledUNetModel3DTV2V
    p = yaml_params("tv2v")
    with pytest.raises(NotImplementedError):
        ControlledUNetModel3DTV2V(**dict(p, use_scale_shift_norm=True))
    with pytest.raises(NotImplementedError):
        ControlledUNetModel3DTV2V(**dict(p, transformer_depth=2))


def test_network_refuses_cpu_tensors():
    """No CPU fallback: the product path raises instead of computing on the host."""
    from ccedit_b200.wrappers import OpenAIWrapperControlLDM3DTV2V
    wrap = OpenAIWrapperControlLDM3DTV2V(torch.nn.Identity())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        wrap(torch.zeros(1, 4, 1, 8, 8), torch.zeros(1, dtype=torch.long),
             {"crossattn": torch.zeros(1, 77, 768), "control_hint": torch.zeros(1, 3, 1, 64, 64)})
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.groupnorm_spatial(torch.zeros(1, 4, 64, dtype=torch.float16), torch.ones(64), torch.zeros(64), 1e-5, False)


# ---------------------------------------------------------------------------------------------------------------------
def _tap_gemm_emulated(a, pw, taps, out_dims):
    """Torch emulation of ccedit_gemm's contraction (include/ccedit_b200.h): a [d4,d3,d2,d1,C] fp32, zero outside."""
    d1, d2, d3, d4 = out_dims
    W = pw.w.float().view(pw.n, pw.ntaps, pw.kpad)
    C = a.shape[-1]
    out = torch.zeros(d4, d3, d2, d1, pw.n)
    pad = 2
    ap = F.pad(a, (0, 0, pad, pad, pad, pad, pad, pad, 0, 0))        # pad d1, d2, d3 (taps never move along d4 here)
    for t, (o1, o2, o3, o4) in enumerate(taps):
        assert o4 == 0
        win = ap[:, pad + o3:pad + o3 + d3, pad + o2:pad + o2 + d2, pad + o1:pad + o1 + d1, :] if o3 + d3 + pad <= ap.shape[1] \
            else None
        out += win @ W[:, t, :C].t()
    if pw.bias is not None:
        out += pw.bias
    return out


def test_pack_weight_conv3x3_layout():
    g = torch.Generator().manual_seed(0)
    x, w, b = torch.randn(2, 24, 5, 7, generator=g), torch.randn(16, 24, 3, 3, generator=g), torch.randn(16, generator=g)
    pw = ops.pack_weight(w, b, "cpu")
    assert (pw.n, pw.k, pw.kpad, pw.ntaps) == (16, 24, 64, 9)
    a = x.permute(0, 2, 3, 1).half().float()[None]                    # [1, F, H, W, C]
    out = _tap_gemm_emulated(a, pw, ops.conv_taps(), (7, 5, 2, 1))[0]
    ref = F.conv2d(x.half().float(), w.half().float(), b, padding=1).permute(0, 2, 3, 1)
    assert rel_err(out, ref) < 1e-5


def test_pack_weight_stride2_parity_planes():
    g = torch.Generator().manual_seed(1)
    x, w = torch.randn(2, 8, 6, 8, generator=g), torch.randn(16, 8, 3, 3, generator=g)
    pw = ops.pack_weight(w, None, "cpu")
    cl = x.permute(0, 2, 3, 1)
    planes = torch.stack([cl[:, ph::2, pwd::2] for ph in (0, 1) for pwd in (0, 1)], 1)   # [F, 4, H/2, W/2, C]
    a = planes.half().float()                                                            # d4=F d3=plane d2=H/2 d1=W/2
    W = pw.w.float().view(pw.n, 9, pw.kpad)
    out = torch.zeros(2, 3, 4, 16)
    ap = F.pad(a, (0, 0, 1, 1, 1, 1))
    for t, (o1, o2, o3, _) in enumerate(ops.conv_s2_taps()):
        out += ap[:, o3, 1 + o2:1 + o2 + 3, 1 + o1:1 + o1 + 4, :] @ W[:, t, :8].t()
    ref = F.conv2d(x.half().float(), w.half().float(), None, stride=2, padding=1).permute(0, 2, 3, 1)
    assert rel_err(out, ref) < 1e-5


def test_pack_weight_temporal_and_geglu():
    g = torch.Generator().manual_seed(2)
    w = torch.randn(32, 16, 3, generator=g)
    pw = ops.pack_weight(w, None, "cpu")
    W = pw.w.float().view(32, 3, pw.kpad)
    assert torch.equal(W[:, 0, :16], w[:, :, 0].half().float()) and float(W[:, :, 16:].abs().max()) == 0
    assert ops.temporal_taps(3) == [(0, -1, 0, 0), (0, 0, 0, 0), (0, 1, 0, 0)]
    # GEGLU: every BN tile = BN/2 value rows then the matching BN/2 gate rows (attention.py:120-122)
    wl, bl = torch.randn(2560, 320, generator=g), torch.randn(2560, generator=g)
    pg = ops.pack_weight(wl, bl, "cpu", geglu=True)
    hb = pg.bn // 2
    Wg = pg.w.float()
    for j in (0, 3):
        assert torch.equal(Wg[j * pg.bn:j * pg.bn + hb, :320], wl[j * hb:(j + 1) * hb].half().float())
        assert torch.equal(Wg[j * pg.bn + hb:(j + 1) * pg.bn, :320], wl[1280 + j * hb:1280 + (j + 1) * hb].half().float())
        assert torch.equal(pg.bias[j * pg.bn + hb:(j + 1) * pg.bn], bl[1280 + j * hb:1280 + (j + 1) * hb])
    assert pg.n_out == 1280


def test_pick_box_and_bn():
    for dims in [(96, 64, 34, 1), (6144, 17, 2, 1), (12, 8, 34, 1), (1, 1, 1, 1), (208896, 1, 1, 1), (48, 32, 4, 34)]:
        box = ops.pick_box(dims)
        assert math.prod(box) == 128 and all(b & (b - 1) == 0 for b in box)
    assert ops.pick_box((96, 64, 34, 1))[:2] in ((32, 4), (16, 8), (64, 2), (128, 1)[:2])
    assert ops.pick_bn(320) == 160 and ops.pick_bn(1280) == 256 and ops.pick_bn(2560, geglu=True) == 256
    assert ops.pick_bn(16) == 16 and ops.pick_bn(96) == 96


def test_pick_bn_prefers_tiles_the_staged_epilogue_can_split():
    """An even number of 16-column chunks lets the TMA-store epilogue split the tile between its two warpgroups."""
    assert ops.pick_bn(960) == 192 and ops.pick_bn(1920) == 192 and ops.pick_bn(640) == 160 and ops.pick_bn(3840) == 256
    assert ops.pick_bn(336) == 112       # no even-chunk tile within 3/4 of the largest divisor: keep the largest


def test_pack_weight_with_folded_layernorm():
    """attention.py:667-669 feeding a Linear: LN(x) W^T + b == rstd * (x (gamma o W)^T - mean * colsum) + (b + W beta),
    with colsum taken over the fp16-rounded weight so that the mean term cancels exactly (host-side algebra only)."""
    g = torch.Generator().manual_seed(5)
    K, N, M = 64, 48, 37
    w, b = torch.randn(N, K, generator=g) / 8, torch.randn(N, generator=g)
    gamma, beta = 1 + 0.3 * torch.randn(K, generator=g), 0.2 * torch.randn(K, generator=g)
    pw = ops.pack_weight(w, b, "cpu", ln_gamma=gamma, ln_beta=beta)
    Wp = pw.w.float()[:, :K]
    assert torch.equal(Wp, (w * gamma[None, :]).half().float())
    assert torch.allclose(pw.colsum, Wp.sum(1), atol=1e-6) and torch.allclose(pw.bias, b + w @ beta, atol=1e-6)
    x = (torch.randn(M, K, generator=g) * 1.5 + 4.0).half().float()
    mean, rstd = x.mean(1, keepdim=True), (x.var(1, unbiased=False, keepdim=True) + 1e-5).rsqrt()
    folded = rstd * (x @ Wp.t() - mean * pw.colsum[None, :]) + pw.bias[None, :]
    ref = F.linear(F.layer_norm(x, (K,), gamma, beta), w, b)
    assert rel_err(folded, ref) < 2e-3   # only the fp16 rounding of gamma o W separates the two
    with pytest.raises(ValueError):
        ops.pack_weight(torch.randn(8, 8, 3, 3), None, "cpu", ln_gamma=torch.ones(8), ln_beta=torch.zeros(8))


def test_pack_hint_stem_weight_layout():
    """csrc/hint_stem.cu operand layout: [N][kpad] with k = tap * cin_pad + channel, tap = kh * 3 + kw, zero padded."""
    g = torch.Generator().manual_seed(6)
    w, b = torch.randn(16, 3, 3, 3, generator=g), torch.randn(16, generator=g)
    wp, bp = ops.pack_hint_stem_weight(w, b, "cpu", 8, 80)
    assert wp.shape == (16, 80) and wp.dtype == torch.float16 and torch.equal(bp, b)
    W = wp.float().view(16, 10, 8)
    for kh in range(3):
        for kw in range(3):
            assert torch.equal(W[:, kh * 3 + kw, :3], w[:, :, kh, kw].half().float())
    assert float(W[:, :9, 3:].abs().max()) == 0 and float(W[:, 9].abs().max()) == 0
    w3 = torch.randn(32, 32, 3, 3, generator=g)                      # layers 2 / 3: 32 output channels
    wp3, _ = ops.pack_hint_stem_weight(w3, torch.zeros(32), "cpu", 32, 288)
    assert wp3.shape == (32, 288) and torch.equal(wp3.float().view(32, 9, 32)[:, 5], w3[:, :, 1, 2].half().float())
    with pytest.raises(RuntimeError):
        ops.pack_hint_stem_weight(torch.randn(48, 3, 3, 3), torch.randn(48), "cpu", 8, 80)
    with pytest.raises(RuntimeError):
        ops.pack_hint_stem_weight(torch.randn(16, 16, 3, 3), torch.randn(16), "cpu", 8, 80)   # more channels than cin_pad


# ---------------------------------------------------------------------------------------------------------------------
def test_sampler_callers_match_oracle():
    """ccedit_b200.sampling (DiscreteDenoiser / VanillaCFGTV2V / DPMPP2SAncestralSampler) vs the oracle's restatement,
    with a cheap stand-in network, on CPU; sigma tables against the reference-generated fixture."""
    from ccedit_b200.sampling import DiscreteDenoiser, DPMPP2SAncestralSampler, LegacyDDPMDiscretization
    g = load_golden("sampler_tv2v.pt")
    den = DiscreteDenoiser()
    assert torch.equal(den.sigmas, g["denoiser_sigmas"])
    assert torch.equal(LegacyDDPMDiscretization()(g["steps"]), g["sampler_sigmas"])
    steps, scale = 6, 7.5
    gen = torch.Generator().manual_seed(0)
    x0 = torch.randn(1, 4, 3, 8, 8, generator=gen)
    c = {"crossattn": torch.randn(1, 77, 768, generator=gen), "control_hint": torch.randn(1, 3, 3, 64, 64, generator=gen)}
    uc = {"crossattn": torch.randn(1, 77, 768, generator=gen), "control_hint": c["control_hint"].clone()}
    noises = [torch.randn(x0.shape, generator=gen) for _ in range(steps)]
    calls = []

    def network(x, t, cond):                 # any deterministic function of all three inputs
        calls.append(t.clone())
        return torch.tanh(x) * (1 + t.view(-1, 1, 1, 1, 1) / 1000.0) + cond["crossattn"].mean(dim=(1, 2)).view(-1, 1, 1, 1, 1)

    sampler = DPMPP2SAncestralSampler(num_steps=steps, device="cpu", guider_config={
        "target": "sgm.modules.diffusionmodules.guiders.VanillaCFGTV2V", "params": {"scale": scale}})
    it = iter(noises)
    sampler.noise_sampler = lambda x: next(it)
    out = sampler(lambda x, s, cond: den(network, x, s, cond), x0.clone(), c, uc=uc)
    n_mine, calls[:] = len(calls), []
    oden = so.DiscreteDenoiserOracle()
    ref = so.dpmpp2s_ancestral_sample(lambda x, s, cond: oden(network, x, s, cond), x0.clone(), c, uc, steps, scale, noises)
    assert n_mine == len(calls) == 2 * steps - 1
    assert torch.allclose(out, ref, rtol=1e-6, atol=1e-6)


# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["tv2v", "tvi2v"])
def test_flop_census_matches_flop_counter_on_oracle(kind, manifests):
    """bench.py's roofline numerator: ccedit_b200.census vs FlopCounterMode over the oracle (meta device, no compute)."""
    from torch.utils.flop_counter import FlopCounterMode
    from ccedit_b200.census import network_flops
    B, T, h, w = 2, 3, 16, 24
    with torch.device("meta"):
        sd = {k: torch.empty(shp) for k, (shp, _) in manifests[kind].items()}
        x, t = torch.empty(B, 4, T, h, w), torch.empty(B, dtype=torch.long)
        c = {"crossattn": torch.empty(B, 77, 768), "control_hint": torch.empty(B, 3, T, 8 * h, 8 * w)}
        img_cfg = None
        ucfg = so.TV2V_UNET_CFG
        if kind == "tvi2v":
            c["cond_feat"] = torch.empty(B, 4, h, w)
            img_cfg = dict(so.TV2V_CONTROLNET_CFG, no_add_x=True, set_input_hint_block_as_identity=True, disable_text_ca=True)
            ucfg = dict(so.TV2V_UNET_CFG, enable_attention3d_crossframe=True, ST3DCA_ca_type="center_self")
        with FlopCounterMode(display=False) as fc:
            so.wrapper_forward(sd, ucfg, so.TV2V_CONTROLNET_CFG, x, t, c, img_cfg)
    assert abs(fc.get_total_flops() - network_flops(kind, B, T, h, w)["total"]) <= 1e-6 * fc.get_total_flops()


def test_flop_census_headline_totals():
    """77.68 TF (tv2v) / 110.31 TF (tvi2v) per call at CFG batch 2 x 17 x 64 x 96; 1.44 TF config 1."""
    from ccedit_b200.census import network_flops
    assert abs(network_flops("tv2v", 2, 17, 64, 96)["total"] / 1e12 - 77.68) < 0.01
    assert abs(network_flops("tvi2v", 2, 17, 64, 96)["total"] / 1e12 - 110.31) < 0.01
    assert abs(network_flops("tv2v", 1, 1, 64, 64)["total"] / 1e12 - 1.44) < 0.01

"""Pins the oracle (oracle/sgm_oracle.py, the CPU restatement) against fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py imports /root/reference on CPU; the fixtures and the script are committed).  The reference has
no tests or golden vectors of its own, so these reference-generated fixtures are the pin.
fp32 on both sides; the restatement uses the same ATen ops, so the tolerance is tight (1e-5 relative to max|ref|)."""
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import inputs as oin
from oracle import sgm_oracle as so
from oracle import vae_oracle as vo

TOL = 2e-5

BLOCK_FN = {
    "rb3d_same": lambda sd, p, a: so.resblock3d(sd, p, *a),
    "rb3d_skip": lambda sd, p, a: so.resblock3d(sd, p, *a),
    "rb3d_cat": lambda sd, p, a: so.resblock3d(sd, p, *a),
    "st3d_320": lambda sd, p, a: so.spatial_transformer_3d(sd, p, a[0], a[1], 8),
    "st3d_1280": lambda sd, p, a: so.spatial_transformer_3d(sd, p, a[0], a[1], 8),
    "down3d": lambda sd, p, a: so.downsample3d(sd, p, a[0]),
    "up3d": lambda sd, p, a: so.upsample3d(sd, p, a[0]),
    "rb2d_skip": lambda sd, p, a: so.resblock(sd, p, *a),
    "st2d_640": lambda sd, p, a: so.spatial_transformer(sd, p, a[0], a[1], 8),
    "down2d": lambda sd, p, a: so.downsample(sd, p, a[0]),
    "st3dca_320": lambda sd, p, a: so.spatial_transformer_3d(sd, p, a[0], a[1], 8, "center_self"),
    "st3dca_1280": lambda sd, p, a: so.spatial_transformer_3d(sd, p, a[0], a[1], 8, "center_self"),
    "st2d_notext": lambda sd, p, a: so.spatial_transformer(sd, p, a[0], a[1], 8, disable_text_ca=True),
}

CN_IMG_CFG = dict(so.TV2V_CONTROLNET_CFG, no_add_x=True, set_input_hint_block_as_identity=True, disable_text_ca=True)
TVI2V_UNET_CFG = dict(so.TV2V_UNET_CFG, enable_attention3d_crossframe=True, ST3DCA_ca_type="center_self")


@pytest.mark.parametrize("kind,name", [("tv2v", n) for n in ("rb3d_same", "rb3d_skip", "rb3d_cat", "st3d_320", "st3d_1280",
                                                             "down3d", "up3d", "rb2d_skip", "st2d_640", "down2d")] +
                         [("tvi2v", n) for n in ("st3dca_320", "st3dca_1280", "st2d_notext")])
def test_block_matches_reference(kind, name, state_dicts):
    g = load_golden(f"blocks_{kind}.pt")[name]
    with torch.no_grad():
        out = BLOCK_FN[name](state_dicts(kind), g["prefix"], g["inputs"])
    assert rel_err(out, g["output"]) < TOL


def test_hint_block_matches_reference(state_dicts):
    g = load_golden("blocks_tv2v.pt")["hint_block"]
    sd = state_dicts("tv2v")
    x = g["inputs"][0]
    with torch.no_grad():
        for i, s in enumerate([1, 1, 2, 1, 2, 1, 2, 1]):
            x = so._conv2d(sd, f"{g['prefix']}.{2 * i}", x, stride=s, padding=1)
            if i < 7:
                x = torch.nn.functional.silu(x)
    assert rel_err(x, g["output"]) < TOL


def test_controlnet2d_matches_reference(state_dicts):
    g = load_golden("blocks_tv2v.pt")["controlnet2d"]
    x, hint, t, ctx = g["inputs"]
    with torch.no_grad():
        outs = so.controlnet2d_forward(state_dicts("tv2v"), so.TV2V_CONTROLNET_CFG, x, hint, t, ctx, g["prefix"])
    assert len(outs) == len(g["output"]) == 13
    for o, r in zip(outs, g["output"]):
        assert rel_err(o, r) < TOL


def test_controlnet_img_matches_reference(state_dicts):
    g = load_golden("blocks_tvi2v.pt")["controlnet_img"]
    x, feat, t, ctx = g["inputs"]
    with torch.no_grad():
        outs = so.controlnet2d_forward(state_dicts("tvi2v"), CN_IMG_CFG, x, feat, t, ctx, g["prefix"])
    for o, r in zip(outs, g["output"]):
        assert rel_err(o, r) < TOL


def test_unet_without_control_matches_reference(state_dicts):
    g = load_golden("blocks_tv2v.pt")["unet_nocontrol"]
    x, t, ctx = g["inputs"]
    with torch.no_grad():
        out = so.unet3d_forward(state_dicts("tv2v"), so.TV2V_UNET_CFG, x, t, ctx, None, None, g["prefix"])
    assert rel_err(out, g["output"]) < TOL


@pytest.mark.parametrize("kind", ["tv2v", "tvi2v"])
def test_network_call_matches_reference(kind, state_dicts):
    g = load_golden(f"network_{kind}.pt")
    B, T, h, w = g["shape"]
    c, uc = oin.synthetic_cond(B, T, h, w, seed=3, tvi2v=(kind == "tvi2v"))
    x0 = oin.synthetic_latent(B, T, h, w, seed=2)
    xin, tin, cc = oin.cfg_batch(x0, torch.tensor([g["t"]]), c, uc)
    assert abs(oin.checksum(xin) - g["x_checksum"]) < 1e-6 * g["x_checksum"]       # same seeded inputs as the fixture
    assert abs(oin.checksum(cc["control_hint"]) - g["hint_checksum"]) < 1e-6 * g["hint_checksum"]
    with torch.no_grad():
        out = so.wrapper_forward(state_dicts(kind), TVI2V_UNET_CFG if kind == "tvi2v" else so.TV2V_UNET_CFG,
                                 so.TV2V_CONTROLNET_CFG, xin, tin, cc, CN_IMG_CFG if kind == "tvi2v" else None)
    assert rel_err(out, g["output"]) < TOL


def test_config1_single_keyframe_matches_reference(state_dicts):
    """BASELINE config 1: single UNet forward, 1 keyframe, 64x64 latent, fp32 on CPU."""
    g = load_golden("config1_tv2v.pt")
    B, T, h, w = g["shape"]
    c, _ = oin.synthetic_cond(B, T, h, w, seed=5)
    x0 = oin.synthetic_latent(B, T, h, w, seed=4)
    with torch.no_grad():
        out = so.wrapper_forward(state_dicts("tv2v"), so.TV2V_UNET_CFG, so.TV2V_CONTROLNET_CFG, x0, torch.tensor([g["t"]]), c)
    assert rel_err(out, g["output"]) < TOL


def test_sigma_tables_match_reference():
    g = load_golden("sampler_tv2v.pt")
    assert torch.equal(so.DiscreteDenoiserOracle().sigmas, g["denoiser_sigmas"])
    assert torch.equal(so.legacy_ddpm_sigmas(g["steps"]), g["sampler_sigmas"])


def test_sampler_matches_reference(state_dicts):
    """3 DPM++2S-ancestral steps with CFG 7.5 through DiscreteDenoiser, against the reference's sampler output."""
    g = load_golden("sampler_tv2v.pt")
    B, T, h, w = g["shape"]
    sd = state_dicts("tv2v")
    c, uc = oin.synthetic_cond(B, T, h, w, seed=7)
    x0 = oin.synthetic_latent(B, T, h, w, seed=6)
    gn = torch.Generator().manual_seed(8)
    noises = [torch.randn(x0.shape, generator=gn) for _ in range(g["steps"])]
    den = so.DiscreteDenoiserOracle()
    sigmas_seen = []

    def network(x, t, cond):
        return so.wrapper_forward(sd, so.TV2V_UNET_CFG, so.TV2V_CONTROLNET_CFG, x, t, cond)

    def denoiser(x, sigma, cond):
        sigmas_seen.append(sigma[0].clone())
        return den(network, x, sigma, cond)

    with torch.no_grad():
        out = so.dpmpp2s_ancestral_sample(denoiser, x0.clone(), c, uc, g["steps"], g["scale"], noises)
    assert len(sigmas_seen) == g["n_calls"] == 2 * g["steps"] - 1
    assert torch.allclose(torch.stack(sigmas_seen), g["call_sigmas"], rtol=1e-6, atol=0)
    assert rel_err(out, g["output"]) < 1e-4


# ---------------------------------------------------------------------------------------------------------------------
# first stage (oracle/vae_oracle.py) against the reference's Decoder / Encoder fixtures (tests/golden/vae.pt)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def vae_sd():
    from oracle.weights import load_manifest, seeded_state_dict
    return seeded_state_dict(load_manifest("vae"), seed=0)


VAE_BLOCK_FN = {
    "res_512": lambda sd, p, x: vo.resnet_block(sd, p, x), "res_256_128": lambda sd, p, x: vo.resnet_block(sd, p, x),
    "attn_512": lambda sd, p, x: vo.attn_block(sd, p, x), "attn_512_big": lambda sd, p, x: vo.attn_block(sd, p, x),
    "up_512": lambda sd, p, x: vo.upsample(sd, p, x), "down_128": lambda sd, p, x: vo.downsample(sd, p, x),
}


@pytest.mark.parametrize("name", sorted(VAE_BLOCK_FN))
def test_vae_block_matches_reference(name, vae_sd):
    g = load_golden("vae.pt")[name]
    with torch.no_grad():
        out = VAE_BLOCK_FN[name](vae_sd, g["prefix"], g["inputs"][0])
    assert rel_err(out, g["output"]) < TOL


def test_vae_decode_and_encode_match_reference(vae_sd):
    gold = load_golden("vae.pt")
    with torch.no_grad():
        for name in ("decode_video", "decode_frame"):
            assert rel_err(vo.decode_first_stage(vae_sd, gold[name]["inputs"][0]), gold[name]["output"]) < TOL
        g = gold["encode_moments"]
        assert rel_err(vo.encode_first_stage_moments(vae_sd, g["inputs"][0]), g["output"]) < TOL


# ---------------------------------------------------------------------------------------------------------------------
# text conditioner: the restatement of transformers' CLIPTextModel against the installed implementation
# ---------------------------------------------------------------------------------------------------------------------
def test_clip_text_oracle_matches_transformers():
    pytest.importorskip("transformers")
    from oracle import clip_oracle as co
    from ccedit_b200.clip_text import CLIPTextModel
    shapes = {k: tuple(v.shape) for k, v in CLIPTextModel().state_dict().items()}
    sd = co.seeded_state_dict(shapes, seed=0)
    ids = torch.randint(0, 49408, (2, 77), generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        mine = co.clip_text_forward(sd, ids)
    ref = co.transformers_reference(sd, ids)
    assert rel_err(mine, ref) < TOL

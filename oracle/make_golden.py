"""ORACLE TOOLING (not product code): generate tests/golden/ from the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only:  python -m oracle.make_golden
Writes   tests/golden/manifest_{tv2v,tvi2v}.json   state-dict key -> [shape, init kind] of the reference networks
         tests/golden/blocks_tv2v.pt, blocks_tvi2v.pt   per-block inputs/outputs (real widths, tiny spatial sizes)
         tests/golden/network_{tv2v,tvi2v}.pt           whole network call outputs (+ input checksums)
         tests/golden/config1_tv2v.pt                   BASELINE config 1: B=1, T=1, 64x64 latent
         tests/golden/sampler_tv2v.pt                   3-step DPM++2S-ancestral + CFG through the reference sampler
Everything the reference computes here goes through its own classes; the oracle restatement is NOT involved.
"""
import contextlib
import io
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import inputs as oin  # noqa: E402
from oracle import ref_import  # noqa: E402
from oracle.weights import GOLDEN_DIR, seeded_state_dict  # noqa: E402


def manifest_of(module):
    man = {}
    for k, v in module.state_dict().items():
        if v.numel() and bool((v == 0).all()):
            init = "zeros"
        elif v.numel() and bool((v == 1).all()):
            init = "ones"
        else:
            init = "default"
        man[k] = [list(v.shape), init]
    return man


def rnd(*shape, seed):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def run(kind):
    t0 = time.time()
    wrap = ref_import.build_reference_network(kind)
    man = manifest_of(wrap)
    with open(os.path.join(GOLDEN_DIR, f"manifest_{kind}.json"), "w") as f:
        json.dump(man, f)
    sd = seeded_state_dict(man, seed=0)
    wrap.load_state_dict(sd, strict=True)
    net = wrap.diffusion_model
    cn = net.controlnet
    print(kind, "built + loaded", len(man), "tensors in", round(time.time() - t0, 1), "s", flush=True)
    blocks = {}
    with torch.no_grad():
        emb1, emb2 = rnd(1, 1280, seed=11), rnd(2, 1280, seed=12)
        ctx2 = rnd(2, 77, 768, seed=13)

        def rec(name, mod, args, prefix):
            out = mod(*args)
            blocks[name] = dict(prefix=prefix, inputs=[a.clone() for a in args], output=out.clone())

        if kind == "tv2v":
            rec("rb3d_same", net.input_blocks[1][0], (rnd(1, 320, 3, 4, 6, seed=21), emb1), "diffusion_model.input_blocks.1.0")
            rec("rb3d_skip", net.input_blocks[4][0], (rnd(2, 320, 2, 4, 6, seed=22), emb2), "diffusion_model.input_blocks.4.0")
            rec("rb3d_cat", net.output_blocks[11][0], (rnd(1, 640, 3, 4, 6, seed=23), emb1), "diffusion_model.output_blocks.11.0")
            rec("st3d_320", net.input_blocks[1][1], (rnd(2, 320, 3, 4, 6, seed=24), ctx2), "diffusion_model.input_blocks.1.1")
            rec("st3d_1280", net.input_blocks[7][1], (rnd(2, 1280, 2, 4, 4, seed=25), ctx2), "diffusion_model.input_blocks.7.1")
            rec("down3d", net.input_blocks[3][0], (rnd(1, 320, 3, 8, 8, seed=26),), "diffusion_model.input_blocks.3.0")
            rec("up3d", net.output_blocks[2][1], (rnd(1, 1280, 2, 3, 4, seed=27),), "diffusion_model.output_blocks.2.1")
            rec("rb2d_skip", cn.input_blocks[4][0], (rnd(2, 320, 4, 6, seed=28), emb2), "diffusion_model.controlnet.input_blocks.4.0")
            rec("st2d_640", cn.input_blocks[4][1], (rnd(2, 640, 4, 6, seed=29), ctx2), "diffusion_model.controlnet.input_blocks.4.1")
            rec("down2d", cn.input_blocks[3][0], (rnd(2, 320, 8, 8, seed=30),), "diffusion_model.controlnet.input_blocks.3.0")
            hint = torch.rand(2, 3, 32, 48, generator=torch.Generator().manual_seed(31))
            out = cn.input_hint_block(hint, emb2, ctx2)
            blocks["hint_block"] = dict(prefix="diffusion_model.controlnet.input_hint_block", inputs=[hint], output=out.clone())
            # whole ControlNet2D on a video batch
            x = rnd(1, 4, 2, 16, 16, seed=32)
            hint5 = torch.rand(1, 3, 2, 128, 128, generator=torch.Generator().manual_seed(33))
            t = torch.tensor([417])
            ctx1 = rnd(1, 77, 768, seed=34)
            outs = cn(x=x, hint=hint5, timesteps=t, context=ctx1)
            blocks["controlnet2d"] = dict(prefix="diffusion_model.controlnet.", inputs=[x, hint5, t, ctx1],
                                          output=[o.clone() for o in outs])
            # UNet without control
            out = net(x, timesteps=t, context=ctx1, control=None, img_control=None)
            blocks["unet_nocontrol"] = dict(prefix="diffusion_model.", inputs=[x, t, ctx1], output=out.clone())
        else:
            rec("st3dca_320", net.input_blocks[1][1], (rnd(2, 320, 3, 4, 6, seed=41), ctx2), "diffusion_model.input_blocks.1.1")
            rec("st3dca_1280", net.input_blocks[7][1], (rnd(1, 1280, 3, 2, 4, seed=42), ctx2[:1]), "diffusion_model.input_blocks.7.1")
            cni = net.controlnet_img
            rec("st2d_notext", cni.input_blocks[1][1], (rnd(2, 320, 4, 6, seed=43), ctx2), "diffusion_model.controlnet_img.input_blocks.1.1")
            x4 = rnd(2, 4, 16, 16, seed=44)
            feat = rnd(2, 4, 16, 16, seed=45)
            t = torch.tensor([900, 33])
            outs = cni(x=x4, hint=feat, timesteps=t, context=ctx2)
            blocks["controlnet_img"] = dict(prefix="diffusion_model.controlnet_img.", inputs=[x4, feat, t, ctx2],
                                            output=[o.clone() for o in outs])
        torch.save(blocks, os.path.join(GOLDEN_DIR, f"blocks_{kind}.pt"))
        print(kind, "blocks done", round(time.time() - t0, 1), "s", flush=True)

        # whole network call at CFG batch 2 (B=1), T=3, 16x16 latent
        B, T, h, w = 1, 3, 16, 16
        c, uc = oin.synthetic_cond(B, T, h, w, seed=3, tvi2v=(kind == "tvi2v"))
        x0 = oin.synthetic_latent(B, T, h, w, seed=2)
        xin, tin, cc = oin.cfg_batch(x0, torch.tensor([640]), c, uc)
        out = wrap(xin, tin, cc)
        torch.save(dict(shape=(B, T, h, w), t=640, x_checksum=oin.checksum(xin),
                        hint_checksum=oin.checksum(cc["control_hint"]), ctx_checksum=oin.checksum(cc["crossattn"]),
                        output=out.clone()), os.path.join(GOLDEN_DIR, f"network_{kind}.pt"))
        print(kind, "network done", round(time.time() - t0, 1), "s", flush=True)

        if kind == "tv2v":
            # BASELINE config 1: single UNet forward, 1 keyframe, 64x64 latent, fp32 CPU, no CFG
            B, T, h, w = 1, 1, 64, 64
            c, _ = oin.synthetic_cond(B, T, h, w, seed=5)
            x0 = oin.synthetic_latent(B, T, h, w, seed=4)
            t1 = time.time()
            out = wrap(x0, torch.tensor([500]), c)
            dt = time.time() - t1
            torch.save(dict(shape=(B, T, h, w), t=500, x_checksum=oin.checksum(x0),
                            hint_checksum=oin.checksum(c["control_hint"]), output=out.clone(), ref_seconds=dt,
                            threads=torch.get_num_threads()), os.path.join(GOLDEN_DIR, "config1_tv2v.pt"))
            print("config1 done: reference CPU call took", round(dt, 2), "s", flush=True)

            # sampler: DiscreteDenoiser + VanillaCFGTV2V + DPMPP2SAncestralSampler, 3 steps, pre-drawn noise
            from sgm.modules.diffusionmodules.denoiser import DiscreteDenoiser
            from sgm.modules.diffusionmodules.sampling import DPMPP2SAncestralSampler

            P = "sgm.modules.diffusionmodules."
            den = DiscreteDenoiser(weighting_config={"target": P + "denoiser_weighting.EpsWeighting"},
                                   scaling_config={"target": P + "denoiser_scaling.EpsScaling"}, num_idx=1000,
                                   discretization_config={"target": P + "discretizer.LegacyDDPMDiscretization"})
            steps, scale = 3, 7.5
            sampler = DPMPP2SAncestralSampler(
                num_steps=steps, discretization_config={"target": P + "discretizer.LegacyDDPMDiscretization"},
                guider_config={"target": P + "guiders.VanillaCFGTV2V", "params": {"scale": scale}}, eta=1.0,
                s_noise=1.0, verbose=False, device="cpu")
            B, T, h, w = 1, 2, 16, 16
            c, uc = oin.synthetic_cond(B, T, h, w, seed=7)
            x0 = oin.synthetic_latent(B, T, h, w, seed=6)
            gn = torch.Generator().manual_seed(8)
            noises = [torch.randn(x0.shape, generator=gn) for _ in range(steps)]
            it = iter(noises)
            sampler.noise_sampler = lambda x: next(it)
            calls = []

            def denoiser(inp, sigma, cond):
                calls.append(sigma.clone())
                return den(wrap, inp, sigma, cond)

            out = sampler(denoiser, x0.clone(), c, uc=uc)
            torch.save(dict(shape=(B, T, h, w), steps=steps, scale=scale, output=out.clone(), n_calls=len(calls),
                            call_sigmas=torch.stack([s[0] for s in calls]), denoiser_sigmas=den.sigmas.clone(),
                            sampler_sigmas=sampler.discretization(steps)), os.path.join(GOLDEN_DIR, "sampler_tv2v.pt"))
            print("sampler done:", len(calls), "network calls", round(time.time() - t0, 1), "s", flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# full-size fixtures: ONE network call of the unmodified reference at the headline shape (BASELINE configs[1] / [2]:
# CFG batch 2 x 17 keyframes x latent 64x96) and at two corners of the configs[4] sweep, plus the same call through the
# oracle's fp16-storage emulation (sgm_oracle.emulate_half_storage) - the error floor of any fp16-storage implementation.
# Inputs and weights are regenerated from seeds by the tests; only the outputs (0.8 MB each) are committed.
#   python -m oracle.make_golden full tv2v headline sweep33 sweep9        (about 4 min per reference call on 8 cores)
# ---------------------------------------------------------------------------------------------------------------------
FULL_CASES = {
    #  name      (B, T, h, w)      t    cond seed, latent seed
    "headline": ((1, 17, 64, 96), 500, 31, 32),
    "sweep33": ((1, 33, 24, 32), 250, 41, 42),
    "sweep9": ((1, 9, 48, 72), 250, 41, 42),
}


def run_full(kind, names):
    from oracle import sgm_oracle as so
    path = os.path.join(GOLDEN_DIR, f"full_{kind}.pt")
    store = torch.load(path, weights_only=False) if os.path.exists(path) else {}
    t0 = time.time()
    wrap = ref_import.build_reference_network(kind)
    from oracle.weights import load_manifest
    sd = seeded_state_dict(load_manifest(kind), seed=0)
    wrap.load_state_dict(sd, strict=True)
    print(kind, "reference built", round(time.time() - t0, 1), "s", flush=True)
    ucfg, icfg = so.TV2V_UNET_CFG, None
    if kind == "tvi2v":
        ucfg = dict(so.TV2V_UNET_CFG, enable_attention3d_crossframe=True, ST3DCA_ca_type="center_self")
        icfg = dict(so.TV2V_CONTROLNET_CFG, no_add_x=True, set_input_hint_block_as_identity=True, disable_text_ca=True)
    for name in names:
        (B, T, h, w), t, cs, ls = FULL_CASES[name]
        c, uc = oin.synthetic_cond(B, T, h, w, seed=cs, tvi2v=(kind == "tvi2v"))
        xin, tin, cc = oin.cfg_batch(oin.synthetic_latent(B, T, h, w, seed=ls), torch.tensor([t]), c, uc)
        with torch.no_grad():
            t1 = time.time()
            ref = wrap(xin, tin, cc)
            dt_ref = time.time() - t1
            print(kind, name, "reference call", round(dt_ref, 1), "s", flush=True)
            t1 = time.time()
            with so.emulate_half_storage():
                emu = so.wrapper_forward(sd, ucfg, so.TV2V_CONTROLNET_CFG, xin, tin, cc, icfg)
            dt_emu = time.time() - t1
        err = float((emu - ref).abs().max() / ref.abs().max())
        print(kind, name, "fp16-storage emulation", round(dt_emu, 1), "s; max|emu-ref|/max|ref| =", f"{err:.3e}", flush=True)
        store[name] = dict(shape=(B, T, h, w), t=t, cond_seed=cs, latent_seed=ls, x_checksum=oin.checksum(xin),
                           hint_checksum=oin.checksum(cc["control_hint"]), output=ref.clone(), output_half_emulated=emu.clone(),
                           ref_seconds=dt_ref, threads=torch.get_num_threads())
        torch.save(store, path)


# ---------------------------------------------------------------------------------------------------------------------
# first stage: the reference's Decoder / Encoder classes (sgm/modules/diffusionmodules/model.py) with
# the ddconfig of the inference YAML, wrapped the way AutoencoderKL does (post_quant_conv / quant_conv, autoencoder.py
# :296-319; the class itself cannot be imported: it derives from pytorch_lightning.LightningModule).
#   python -m oracle.make_golden vae
# ---------------------------------------------------------------------------------------------------------------------
def run_vae():
    import torch.nn as nn
    ref_import.install_shim()
    from sgm.modules.diffusionmodules.model import Decoder, Encoder
    from oracle.vae_oracle import DDCONFIG, SCALE_FACTOR
    t0 = time.time()

    class FirstStage(nn.Module):                        # AutoencoderKL.__init__, autoencoder.py:283-303
        def __init__(self):
            super().__init__()
            with contextlib.redirect_stdout(io.StringIO()):
                self.encoder = Encoder(**DDCONFIG)
                self.decoder = Decoder(**DDCONFIG)
            self.quant_conv = nn.Conv2d(2 * DDCONFIG["z_channels"], 2 * 4, 1)
            self.post_quant_conv = nn.Conv2d(4, DDCONFIG["z_channels"], 1)

    fs = FirstStage().eval()
    man = manifest_of(fs)
    with open(os.path.join(GOLDEN_DIR, "manifest_vae.json"), "w") as f:
        json.dump(man, f)
    fs.load_state_dict(seeded_state_dict(man, seed=0), strict=True)
    dec, enc = fs.decoder, fs.encoder
    out = {}
    with torch.no_grad():
        def rec(name, mod, x, prefix, *extra):
            out[name] = dict(prefix=prefix, inputs=[x.clone()], output=mod(x, *extra).clone())

        rec("res_512", dec.mid.block_1, rnd(2, 512, 8, 12, seed=61), "decoder.mid.block_1", None)
        rec("res_256_128", dec.up[0].block[0], rnd(1, 256, 16, 24, seed=62), "decoder.up.0.block.0", None)
        rec("attn_512", dec.mid.attn_1, rnd(2, 512, 16, 12, seed=63), "decoder.mid.attn_1")
        rec("attn_512_big", dec.mid.attn_1, rnd(1, 512, 32, 40, seed=64) * 2.0, "decoder.mid.attn_1")
        rec("up_512", dec.up[3].upsample, rnd(1, 512, 6, 8, seed=65), "decoder.up.3.upsample")
        rec("down_128", enc.down[0].downsample, rnd(2, 128, 16, 24, seed=66), "encoder.down.0.downsample")
        # decode_first_stage (diffusion.py:152-156) on a video latent: 2 frames of 8x12 -> 64x96 pixels
        z = rnd(1, 4, 2, 8, 12, seed=67)
        zz = (1.0 / SCALE_FACTOR * z).permute(0, 2, 1, 3, 4).reshape(2, 4, 8, 12)
        d = dec(fs.post_quant_conv(zz))
        out["decode_video"] = dict(inputs=[z], output=d.reshape(1, 2, 3, 64, 96).permute(0, 2, 1, 3, 4).clone())
        # one frame at a latent whose token count is not a multiple of 128 (mid attention tails)
        z1 = rnd(1, 4, 1, 16, 24, seed=68)
        d1 = dec(fs.post_quant_conv((1.0 / SCALE_FACTOR * z1)[:, :, 0]))
        out["decode_frame"] = dict(inputs=[z1], output=d1[:, :, None].clone())
        # encoder moments (quant_conv(encoder(x))) on 2 frames of 64x96 pixels
        x = torch.rand(2, 3, 64, 96, generator=torch.Generator().manual_seed(69)) * 2 - 1
        out["encode_moments"] = dict(inputs=[x], output=fs.quant_conv(enc(x)).clone())
    torch.save(out, os.path.join(GOLDEN_DIR, "vae.pt"))
    print("vae fixtures done", round(time.time() - t0, 1), "s", {k: tuple(v["output"].shape) for k, v in out.items()}, flush=True)


if __name__ == "__main__":
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.manual_seed(0)
    if len(sys.argv) > 1 and sys.argv[1] == "vae":
        run_vae()
    elif len(sys.argv) > 1 and sys.argv[1] == "full":
        run_full(sys.argv[2], sys.argv[3:] or list(FULL_CASES))
    else:
        for kind in (sys.argv[1:] or ["tv2v", "tvi2v"]):
            run(kind)

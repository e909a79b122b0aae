"""Parity report of the CUDA path against every committed fixture (outputs of the UNMODIFIED reference, CPU fp32).

For each fixture: the CUDA path's max|err|/max|ref| and mean|err|/mean|ref|, the fraction of elements inside north_star's
elementwise bound |err| <= 1e-4 + 1e-3 |ref|, and the same three numbers for the oracle's fp16-storage emulation
(sgm_oracle.emulate_half_storage: conv / linear / SDPA operands and results rounded to fp16, fp32 arithmetic - what
the reference's own GPU run under autocast stores) - the floor of any fp16-storage implementation of the path.
Run on a GPU box:  python tools/parity_report.py > gpurun_out/parity.md   (committed as profiles/rNN_parity.md)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import GOLDEN, load_golden  # noqa: E402
import test_blocks_gpu as tb  # noqa: E402
import test_oracle_golden as tg  # noqa: E402
from oracle import inputs as oin  # noqa: E402
from oracle import sgm_oracle as so  # noqa: E402
from oracle.weights import load_manifest, seeded_state_dict  # noqa: E402
from ccedit_b200.configs import build_network  # noqa: E402

RTOL, ATOL = 1e-3, 1e-4


def stats(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    err = (got - ref).abs()
    return (float(err.max() / ref.abs().max()), float(err.mean() / ref.abs().mean()),
            float((err <= ATOL + RTOL * ref.abs()).float().mean()))


def main():
    print("| fixture | shape | CUDA max-norm err | CUDA mean err | CUDA inside rtol 1e-3 / atol 1e-4 | fp16-storage floor: "
          "max-norm | mean | inside |")
    print("|---|---|---|---|---|---|---|---|")

    def row(name, got, ref, emu=None):
        a = stats(got, ref)
        b = stats(emu, ref) if emu is not None else None
        tail = f"{b[0]:.2e} | {b[1]:.2e} | {100 * b[2]:.1f} %" if b else "- | - | -"
        print(f"| {name} | {tuple(ref.shape)} | {a[0]:.2e} | {a[1]:.2e} | {100 * a[2]:.1f} % | {tail} |", flush=True)

    for kind, names in (("tv2v", tb.TV2V_BLOCKS), ("tvi2v", tb.TVI2V_BLOCKS)):
        sd = seeded_state_dict(load_manifest(kind), seed=0)
        wrap = build_network(kind, device="cpu", use_cuda_graph=False)
        wrap.load_state_dict(sd, strict=True)
        wrap = wrap.cuda()
        ucfg = tg.TVI2V_UNET_CFG if kind == "tvi2v" else so.TV2V_UNET_CFG
        icfg = tg.CN_IMG_CFG if kind == "tvi2v" else None
        blocks = load_golden(f"blocks_{kind}.pt")
        with torch.no_grad():
            for name in names:
                g = blocks[name]
                block = tb._module(wrap, g["prefix"])
                x = g["inputs"][0]
                B, T = x.shape[0], (x.shape[2] if x.dim() == 5 else 1)
                emb = g["inputs"][1] if name.startswith("rb") else None
                context = g["inputs"][1] if name.startswith("st") else None
                out = block.run(tb._cl(x), tb._ctx_for(block, emb, context, B, T))
                with so.emulate_half_storage():
                    emu = tg.BLOCK_FN[name](sd, g["prefix"], g["inputs"])
                row(f"{kind}/{name}", tb._back(out), g["output"], emu)
            if kind == "tv2v":
                g = blocks["controlnet2d"]
                outs = wrap.diffusion_model.controlnet(*[a.cuda() for a in g["inputs"][:2]], timesteps=g["inputs"][2].cuda(),
                                                       context=g["inputs"][3].cuda())
                with so.emulate_half_storage():
                    emus = so.controlnet2d_forward(sd, so.TV2V_CONTROLNET_CFG, *g["inputs"], g["prefix"])
                for i, (o, r, e) in enumerate(zip(outs, g["output"], emus)):
                    row(f"tv2v/controlnet2d[{i}]", o, r, e)
                g = blocks["unet_nocontrol"]
                with so.emulate_half_storage():
                    emu = so.unet3d_forward(sd, so.TV2V_UNET_CFG, *g["inputs"], None, None, g["prefix"])
                row("tv2v/unet_nocontrol", wrap.diffusion_model(g["inputs"][0].cuda(), timesteps=g["inputs"][1].cuda(),
                                                                context=g["inputs"][2].cuda()), g["output"], emu)
            g = load_golden(f"network_{kind}.pt")
            B, T, h, w = g["shape"]
            c, uc = oin.synthetic_cond(B, T, h, w, seed=3, tvi2v=(kind == "tvi2v"))
            xin, tin, cc = oin.cfg_batch(oin.synthetic_latent(B, T, h, w, seed=2), torch.tensor([g["t"]]), c, uc)
            with so.emulate_half_storage():
                emu = so.wrapper_forward(sd, ucfg, so.TV2V_CONTROLNET_CFG, xin, tin, cc, icfg)
            row(f"{kind}/network call", wrap(xin.cuda(), tin.cuda(), {k: v.cuda() for k, v in cc.items()}), g["output"], emu)
            if kind == "tv2v":
                g = load_golden("config1_tv2v.pt")
                B, T, h, w = g["shape"]
                c, _ = oin.synthetic_cond(B, T, h, w, seed=5)
                x0 = oin.synthetic_latent(B, T, h, w, seed=4)
                with so.emulate_half_storage():
                    emu = so.wrapper_forward(sd, ucfg, so.TV2V_CONTROLNET_CFG, x0, torch.tensor([g["t"]]), c)
                row("tv2v/config1 (1x1x64x64)", wrap(x0.cuda(), torch.tensor([g["t"]]).cuda(), {k: v.cuda() for k, v in c.items()}),
                    g["output"], emu)
            # full-size calls: reference and emulation outputs were generated in the build container (make_golden.py full)
            if os.path.exists(os.path.join(GOLDEN, f"full_{kind}.pt")):
                for name, g in load_golden(f"full_{kind}.pt").items():
                    B, T, h, w = g["shape"]
                    c, uc = oin.synthetic_cond(B, T, h, w, seed=g["cond_seed"], tvi2v=(kind == "tvi2v"))
                    xin, tin, cc = oin.cfg_batch(oin.synthetic_latent(B, T, h, w, seed=g["latent_seed"]),
                                                 torch.tensor([g["t"]]), c, uc)
                    out = wrap(xin.cuda(), tin.cuda(), {k: v.cuda() for k, v in cc.items()})
                    row(f"{kind}/FULL {name} (CFG 2 x {T} x {h}x{w})", out, g["output"], g["output_half_emulated"])
        del wrap
        torch.cuda.empty_cache()

    # ---- sampler (3 DPM++2S-ancestral steps, CFG 7.5, 5 network calls) through the fused step ----
    from ccedit_b200.sampling import BoundDenoiser, DiscreteDenoiser, FusedDPMPP2SAncestralSampler
    sd = seeded_state_dict(load_manifest("tv2v"), seed=0)
    wrap = build_network("tv2v", device="cpu", use_cuda_graph=False)
    wrap.load_state_dict(sd, strict=True)
    wrap = wrap.cuda()
    g = load_golden("sampler_tv2v.pt")
    B, T, h, w = g["shape"]
    c, uc = oin.synthetic_cond(B, T, h, w, seed=7)
    x0 = oin.synthetic_latent(B, T, h, w, seed=6)
    gn = torch.Generator().manual_seed(8)
    noises = iter([torch.randn(x0.shape, generator=gn).cuda() for _ in range(g["steps"])])
    sampler = FusedDPMPP2SAncestralSampler(num_steps=g["steps"], device="cuda", eta=1.0, s_noise=1.0, guider_config={
        "target": "sgm.modules.diffusionmodules.guiders.VanillaCFGTV2V", "params": {"scale": g["scale"]}})
    sampler.noise_sampler = lambda x: next(noises)
    cu = lambda d: {k: v.cuda() for k, v in d.items()}
    out = sampler(BoundDenoiser(DiscreteDenoiser().cuda(), wrap), x0.clone().cuda(), cu(c), uc=cu(uc))
    row("tv2v/sampler 3 steps, cfg 7.5 (fused step, CFG de-duplication)", out, g["output"])
    del wrap
    torch.cuda.empty_cache()

    # ---- first stage (reference runs it in fp32: the floor columns do not apply) ----
    from oracle.vae_oracle import DDCONFIG, SCALE_FACTOR
    from ccedit_b200.autoencoder import AutoencoderKLInferenceWrapper
    import test_vae_gpu as tv
    vae = AutoencoderKLInferenceWrapper(ddconfig=dict(DDCONFIG), embed_dim=4)
    vae.load_state_dict(seeded_state_dict(load_manifest("vae"), seed=0), strict=True)
    vae = vae.cuda().eval()
    gold = load_golden("vae.pt")
    with torch.no_grad():
        for name in ("res_512", "res_256_128", "attn_512", "attn_512_big", "up_512", "down_128"):
            gg = gold[name]
            row(f"vae/{name}", tv._mod(vae, gg["prefix"]).run(tv._cl(gg["inputs"][0])).permute(0, 3, 1, 2), gg["output"])
        for name in ("decode_video", "decode_frame"):
            row(f"vae/{name}", vae.decode(gold[name]["inputs"][0].cuda(), scale=1.0 / SCALE_FACTOR), gold[name]["output"])
        row("vae/encode_moments", vae.encode_moments(gold["encode_moments"]["inputs"][0].cuda()), gold["encode_moments"]["output"])


if __name__ == "__main__":
    main()

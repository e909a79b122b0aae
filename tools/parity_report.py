"""Print the measured parity errors (max|err| / max|ref|) of every committed fixture through the CUDA path.
Run on a GPU box:  python tools/parity_report.py > gpurun_out/parity.md   (summarised in profiles/rNN_parity.md)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import load_golden, rel_err  # noqa: E402
import test_blocks_gpu as tb  # noqa: E402
from oracle import inputs as oin  # noqa: E402
from oracle.weights import load_manifest, seeded_state_dict  # noqa: E402
from ccedit_b200.configs import build_network  # noqa: E402


def wrapper(kind):
    w = build_network(kind, device="cpu", use_cuda_graph=False)
    w.load_state_dict(seeded_state_dict(load_manifest(kind), seed=0), strict=True)
    return w.cuda()


def main():
    print("| fixture | shape | max|err|/max|ref| | mean|err|/mean|ref| |")
    print("|---|---|---|---|")

    def row(name, got, ref):
        got, ref = got.float().cpu(), ref.float().cpu()
        print(f"| {name} | {tuple(ref.shape)} | {rel_err(got, ref):.3e} | "
              f"{float((got - ref).abs().mean() / ref.abs().mean()):.3e} |", flush=True)

    for kind, names in (("tv2v", tb.TV2V_BLOCKS), ("tvi2v", tb.TVI2V_BLOCKS)):
        wrap = wrapper(kind)
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
                row(f"{kind}/{name}", tb._back(out), g["output"])
            if kind == "tv2v":
                g = blocks["controlnet2d"]
                outs = wrap.diffusion_model.controlnet(*[a.cuda() for a in g["inputs"][:2]], timesteps=g["inputs"][2].cuda(),
                                                       context=g["inputs"][3].cuda())
                for i, (o, r) in enumerate(zip(outs, g["output"])):
                    row(f"tv2v/controlnet2d[{i}]", o, r)
                g = blocks["unet_nocontrol"]
                row("tv2v/unet_nocontrol", wrap.diffusion_model(g["inputs"][0].cuda(), timesteps=g["inputs"][1].cuda(),
                                                                context=g["inputs"][2].cuda()), g["output"])
            g = load_golden(f"network_{kind}.pt")
            B, T, h, w = g["shape"]
            c, uc = oin.synthetic_cond(B, T, h, w, seed=3, tvi2v=(kind == "tvi2v"))
            xin, tin, cc = oin.cfg_batch(oin.synthetic_latent(B, T, h, w, seed=2), torch.tensor([g["t"]]), c, uc)
            row(f"{kind}/network call", wrap(xin.cuda(), tin.cuda(), {k: v.cuda() for k, v in cc.items()}), g["output"])
            if kind == "tv2v":
                g = load_golden("config1_tv2v.pt")
                B, T, h, w = g["shape"]
                c, _ = oin.synthetic_cond(B, T, h, w, seed=5)
                x0 = oin.synthetic_latent(B, T, h, w, seed=4)
                row("tv2v/config1 (1x1x64x64)", wrap(x0.cuda(), torch.tensor([g["t"]]).cuda(), {k: v.cuda() for k, v in c.items()}),
                    g["output"])
        del wrap
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

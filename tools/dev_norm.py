"""Developer timing of the HBM-bound kernels at the headline network's shapes (NOT the parity suite): spatial / temporal
GroupNorm and the temporal attention, CUDA events over rotating buffers larger than the L2.  Also checks that the
kernels agree with torch fp32 on the same inputs.  Env toggles select the older schedules (see norm.cu / attention.cu)."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccedit_b200 import ops  # noqa: E402

torch.manual_seed(0)
dev = "cuda"


def timed(run, nbuf, iters=12):
    for i in range(nbuf):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        run(i % nbuf)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def gn_spatial(Fr, HW, C):
    nb = max(2, int(400e6 // (Fr * HW * C * 2)) + 1)
    xs = [torch.randn(Fr, HW, C, device=dev).half() for _ in range(nb)]
    ys = [torch.empty_like(x) for x in xs]
    g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
    ms = timed(lambda i: ops.groupnorm_spatial(xs[i], g, b, 1e-5, True, out=ys[i]), nb)
    ref = F.silu(F.group_norm(xs[0].float().transpose(1, 2), 32, g, b, 1e-5)).transpose(1, 2)
    err = (ys[0].float() - ref).abs().max().item()
    gb = 2.0 * Fr * HW * C * 2 / 1e9
    print(f"gn_spatial  F={Fr:3d} HW={HW:5d} C={C:4d}: {ms * 1e3:8.1f} us {gb / ms * 1e3:8.0f} GB/s  err {err:.2e}", flush=True)


def gn_temporal(B, T, HW, C):
    nb = max(2, int(400e6 // (B * T * HW * C * 2)) + 1)
    xs = [torch.randn(B, T, HW, C, device=dev).half() for _ in range(nb)]
    ys = [torch.empty_like(x) for x in xs]
    g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
    ms = timed(lambda i: ops.groupnorm_temporal(xs[i], g, b, 1e-5, True, out=ys[i]), nb)
    xr = xs[0].float().permute(0, 2, 3, 1).reshape(B * HW, C, T)
    ref = F.silu(F.group_norm(xr, 32, g, b, 1e-5)).view(B, HW, C, T).permute(0, 3, 1, 2)
    err = (ys[0].float() - ref).abs().max().item()
    gb = 2.0 * B * T * HW * C * 2 / 1e9
    print(f"gn_temporal B={B} T={T:2d} HW={HW:5d} C={C:4d}: {ms * 1e3:8.1f} us {gb / ms * 1e3:8.0f} GB/s  err {err:.2e}", flush=True)


def t_attn(B, T, HW, heads, d):
    C = heads * d
    nb = max(2, int(400e6 // (B * T * HW * 4 * C * 2)) + 1)
    qkvs = [torch.randn(B, T, HW, 3 * C, device=dev).half() for _ in range(nb)]
    outs = [torch.empty(B, T, HW, C, dtype=torch.float16, device=dev) for _ in range(nb)]
    ms = timed(lambda i: ops.temporal_attention(qkvs[i][..., :C], qkvs[i][..., C:2 * C], qkvs[i][..., 2 * C:], heads, outs[i]), nb)
    sp = lambda t: t.float().permute(0, 2, 1, 3).reshape(B * HW, T, heads, d).transpose(1, 2)
    q, k, v = qkvs[0][..., :C], qkvs[0][..., C:2 * C], qkvs[0][..., 2 * C:]
    ref = F.scaled_dot_product_attention(sp(q), sp(k), sp(v)).transpose(1, 2).reshape(B, HW, T, C).permute(0, 2, 1, 3)
    err = (outs[0].float() - ref).abs().max().item()
    gb = 4.0 * B * T * HW * C * 2 / 1e9
    print(f"t_attn      B={B} T={T:2d} HW={HW:5d} h={heads} d={d:3d}: {ms * 1e3:8.1f} us {gb / ms * 1e3:8.0f} GB/s  err {err:.2e}", flush=True)


if __name__ == "__main__":
    print({k: v for k, v in os.environ.items() if k.startswith("CCEDIT_")}, flush=True)
    for Fr, HW, C in ((34, 6144, 320), (34, 6144, 640), (34, 6144, 960), (34, 1536, 640), (34, 1536, 1280), (34, 384, 1280),
                      (34, 384, 2560), (17, 6144, 320), (34, 96, 1280)):
        gn_spatial(Fr, HW, C)
    for T in (17, 9, 33):
        for HW, C in ((6144, 320), (1536, 640), (384, 1280), (96, 1280)):
            gn_temporal(2, T, HW, C)
    for T in (17, 9, 33):
        for HW, h, d in ((6144, 8, 40), (1536, 8, 80), (384, 8, 160), (96, 8, 160)):
            t_attn(2, T, HW, h, d)

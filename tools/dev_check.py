"""Developer smoke/perf check of every kernel against PyTorch on the GPU (NOT the parity suite - that is tests/ with
the oracle).  Usage on a GPU box: python tools/dev_check.py [--perf]"""
import math
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccedit_b200 import ops  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
FAILS = []


def report(name, got, ref, tol=2e-3):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    ok = err <= tol * scale + 1e-3
    print(f"{'OK ' if ok else 'BAD'} {name:60s} max|err|={err:.4e} ref_max={scale:.3e}", flush=True)
    if not ok:
        FAILS.append(name)


def rnd(*shape, s=1.0):
    return (torch.randn(*shape, device=dev) * s).half()


def check_linear(M, K, N, geglu=False, res=False, bias=True):
    a = rnd(M, K)
    w = torch.randn(N, K, device=dev) / math.sqrt(K)
    b = torch.randn(N, device=dev) if bias else None
    pw = ops.pack_weight(w, b, dev, geglu=geglu)
    n_out = N // 2 if geglu else N
    out = torch.empty(M, n_out, dtype=torch.float16, device=dev)
    r = rnd(M, n_out) if res else None
    ops.gemm(a, pw, out, res1=r)
    ref = a.float() @ w.half().float().t()
    if bias:
        ref = ref + b
    if geglu:
        ref = ref[:, :n_out] * F.gelu(ref[:, n_out:])
    if res:
        ref = ref + r.float()
    report(f"linear M={M} K={K} N={N} geglu={geglu} res={res}", out, ref)


def check_conv3(Fr, H, W, Cin, Cout, emb=False, silu=False):
    x = rnd(Fr, H, W, Cin)
    w = torch.randn(Cout, Cin, 3, 3, device=dev) / math.sqrt(9 * Cin)
    b = torch.randn(Cout, device=dev)
    pw = ops.pack_weight(w, b, dev)
    out = torch.empty(Fr, H, W, pw.n, dtype=torch.float16, device=dev)
    T = 2 if Fr % 2 == 0 else 1
    rb = torch.randn(Fr // T, pw.n, device=dev) if emb else None
    ops.gemm(x, pw, out, ops.conv_taps(), rowbias=rb, rb_dim=2, rb_div=T, silu=silu)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.half().float(), b, padding=1).permute(0, 2, 3, 1)
    if emb:
        ref = ref + rb[:, :Cout].repeat_interleave(T, 0)[:, None, None, :]
    if silu:
        ref = F.silu(ref)
    report(f"conv3x3 F={Fr} {H}x{W} {Cin}->{Cout} emb={emb} silu={silu}", out[..., :Cout], ref)


def check_conv_s2(Fr, H, W, Cin, Cout):
    x = rnd(Fr, H, W, Cin)
    w = torch.randn(Cout, Cin, 3, 3, device=dev) / math.sqrt(9 * Cin)
    b = torch.randn(Cout, device=dev)
    pw = ops.pack_weight(w, b, dev)
    planes = ops.parity_split(x)  # [F,4,H2,W2,C]
    out = torch.empty(Fr, 1, H // 2, W // 2, Cout, dtype=torch.float16, device=dev)
    ops.gemm(planes, pw, out, ops.conv_s2_taps())
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.half().float(), b, stride=2, padding=1).permute(0, 2, 3, 1)
    report(f"conv3x3/s2 F={Fr} {H}x{W} {Cin}->{Cout}", out[:, 0], ref)


def check_tconv(B, T, HW, C, k=3):
    x = rnd(B, T, HW, C)
    w = torch.randn(C, C, k, device=dev) / math.sqrt(k * C)
    b = torch.randn(C, device=dev)
    pw = ops.pack_weight(w, b, dev)
    out = torch.empty_like(x)
    res = rnd(B, T, HW, C)
    ops.gemm(x, pw, out, ops.temporal_taps(k), res1=res)
    xr = x.float().permute(0, 2, 3, 1).reshape(B * HW, C, T)
    ref = F.conv1d(xr, w.half().float(), b, padding=k // 2).reshape(B, HW, C, T).permute(0, 3, 1, 2) + res.float()
    report(f"tconv k={k} B={B} T={T} HW={HW} C={C}", out, ref)


def check_norms():
    for (Fr, HW, C) in [(4, 96, 320), (3, 1536, 640), (2, 384, 1280), (2, 96, 2560), (2, 256, 64), (2, 600, 960)]:
        x = rnd(Fr, HW, C) * 2 + 0.5
        g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
        for silu in (False, True):
            y = ops.groupnorm_spatial(x, g, b, 1e-5, silu)
            ref = F.group_norm(x.float().permute(0, 2, 1), 32, g, b, 1e-5).permute(0, 2, 1)
            ref = F.silu(ref) if silu else ref
            report(f"gn_spatial F={Fr} HW={HW} C={C} silu={silu}", y, ref)
    for (B, T, HW, C) in [(2, 17, 96, 320), (1, 9, 40, 1280), (2, 33, 24, 640), (2, 3, 50, 64)]:
        x = rnd(B, T, HW, C) * 2 + 0.5
        g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
        y = ops.groupnorm_temporal(x, g, b, 1e-6, True)
        xr = x.float().permute(0, 2, 3, 1).reshape(B * HW, C, T)
        ref = F.silu(F.group_norm(xr, 32, g, b, 1e-6)).reshape(B, HW, C, T).permute(0, 3, 1, 2)
        report(f"gn_temporal B={B} T={T} HW={HW} C={C}", y, ref)
    for (M, C) in [(1000, 320), (77, 640), (513, 1280), (64, 64)]:
        x = rnd(M, C) * 3 + 1
        g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
        y = ops.layernorm(x, g, b)
        report(f"layernorm M={M} C={C}", y, F.layer_norm(x.float(), (C,), g, b))
    big = rnd(200, 960)
    y = ops.layernorm(big[:, 320:640], torch.ones(320, device=dev), torch.zeros(320, device=dev))
    report("layernorm strided slice", y, F.layer_norm(big[:, 320:640].float(), (320,)))


def check_attention():
    for (Fr, L, heads, d, Lkv) in [(3, 384, 8, 40, 384), (2, 200, 8, 80, 200), (2, 96, 8, 160, 96), (4, 300, 8, 40, 77),
                                   (2, 128, 4, 16, 128), (2, 1536, 8, 80, 1536)]:
        C = heads * d
        qkv = rnd(Fr, L, 3 * C)
        if Lkv == L:
            q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
            seg = ops.KVSegment(k, v)
            kk, vv = k, v
        else:
            q = qkv[..., :C]
            kv = rnd(2, Lkv, 2 * C)
            seg = ops.KVSegment(kv[..., :C], kv[..., C:], div=Fr // 2)
            kk = kv[..., :C].repeat_interleave(Fr // 2, 0)
            vv = kv[..., C:].repeat_interleave(Fr // 2, 0)
        out = torch.empty(Fr, L, C, dtype=torch.float16, device=dev)
        ops.attention(q, [seg], heads, out)
        sh = lambda t: t.float().reshape(t.shape[0], t.shape[1], heads, d).transpose(1, 2)
        ref = F.scaled_dot_product_attention(sh(q), sh(kk), sh(vv)).transpose(1, 2).reshape(Fr, L, C)
        report(f"attention F={Fr} L={L} Lkv={Lkv} heads={heads} d={d}", out, ref)
    # two segments (center_self)
    B, T, L, heads, d = 2, 3, 96, 8, 40
    C = heads * d
    q = rnd(B * T, L, C)
    kv = rnd(B * T, L, 2 * C)
    k, v = kv[..., :C], kv[..., C:]
    out = torch.empty_like(q)
    ops.attention(q, [ops.KVSegment(k, v, div=T, mul=T, add=T // 2), ops.KVSegment(k, v)], heads, out)
    kc = k.reshape(B, T, L, C)[:, T // 2].repeat_interleave(T, 0)
    vc = v.reshape(B, T, L, C)[:, T // 2].repeat_interleave(T, 0)
    kk, vv = torch.cat([kc, k], 1), torch.cat([vc, v], 1)
    sh = lambda t: t.float().reshape(t.shape[0], t.shape[1], heads, d).transpose(1, 2)
    ref = F.scaled_dot_product_attention(sh(q), sh(kk), sh(vv)).transpose(1, 2).reshape(B * T, L, C)
    report("attention center_self 2 segments", out, ref)
    for (B, T, HW, heads, d) in [(2, 17, 50, 8, 40), (1, 9, 30, 8, 160), (2, 33, 20, 8, 80), (1, 40, 10, 8, 40)]:
        C = heads * d
        q = rnd(B, T, HW, C)
        kv = rnd(B, T, HW, 2 * C)
        out = torch.empty_like(q)
        ops.temporal_attention(q, kv[..., :C], kv[..., C:], heads, out)
        sh = lambda t: t.float().permute(0, 2, 1, 3).reshape(B * HW, T, heads, d).transpose(1, 2)
        ref = F.scaled_dot_product_attention(sh(q), sh(kv[..., :C]), sh(kv[..., C:]))
        ref = ref.transpose(1, 2).reshape(B, HW, T, C).permute(0, 2, 1, 3)
        report(f"temporal_attention B={B} T={T} HW={HW} d={d}", out, ref)


def check_small():
    x = torch.randn(2, 4, 3, 8, 12, device=dev)
    y = ops.ncthw_to_cl(x, 8)
    report("ncthw_to_cl", y[..., :4], x.permute(0, 2, 3, 4, 1))
    assert y[..., 4:].abs().max().item() == 0
    h = torch.rand(2, 3, 3, 16, 24, device=dev) * 2 - 1
    report("hint transform", ops.ncthw_to_cl(h, 8, -0.5, 1.0, 1.0)[..., :3], (1 - (h + 1) / 2).permute(0, 2, 3, 4, 1))
    t = torch.tensor([999.0, 17.0, 0.0], device=dev)
    half = 160
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32, device=dev) / half)
    args = t[:, None] * freqs[None]
    report("timestep_embedding", ops.timestep_embedding(t, 320), torch.cat([torch.cos(args), torch.sin(args)], -1), 1e-4)
    xs = torch.randn(2, 1280, device=dev)
    w = (torch.randn(960, 1280, device=dev) / 36).half()
    b = torch.randn(960, device=dev)
    report("linear_small silu-in", ops.linear_small(xs, w, b, act_in=True), F.silu(xs) @ w.float().t() + b, 1e-4)
    report("linear_small silu-out", ops.linear_small(xs, w, b, act_out=True), F.silu(xs @ w.float().t() + b), 1e-4)
    x = rnd(3, 8, 12, 64)
    report("upsample", ops.upsample_nearest2x(x),
           F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1))
    a, b2 = rnd(100, 320), rnd(100, 320)
    dst = torch.zeros(100, 960, dtype=torch.float16, device=dev)
    ops.add_rows(a, b2, dst[:, 640:])
    report("add_rows into slice", dst[:, 640:], a.float() + b2.float())
    assert dst[:, :640].abs().max().item() == 0
    xx, yy = rnd(2, 5, 4, 6, 64), rnd(2, 4, 6, 64)
    ref = xx.clone().float()
    ref[:, 2] += yy.float()
    report("add_center_frame", ops.add_center_frame(xx, yy), ref)
    B, T, H, W = 2, 5, 6, 8
    y16 = rnd(B, T, H, W, 16)
    wt, bt = torch.randn(4, 4, 3, device=dev), torch.randn(4, device=dev)
    o = ops.out_temporal(y16, wt, bt, 4, torch.float32)
    yr = y16[..., :4].float().permute(0, 2, 3, 4, 1).reshape(B * H * W, 4, T)
    ref = (yr + F.conv1d(F.silu(yr), wt, bt, padding=1)).reshape(B, H, W, 4, T).permute(0, 3, 4, 1, 2)
    report("out_temporal", o, ref)


def perf():
    print("\n--- GEMM perf (CUDA events, 10 iters) ---")
    cases = [
        ("L0 FF-in geglu", "lin", (208896, 320, 2560, True)),
        ("L0 FF-out", "lin", (208896, 1280, 320, False)),
        ("L0 qkv", "lin", (208896, 320, 960, False)),
        ("L1 FF-in geglu", "lin", (52224, 640, 5120, True)),
        ("L2 FF-in geglu", "lin", (13056, 1280, 10240, True)),
        ("L0 conv3 320", "conv", (34, 64, 96, 320, 320)),
        ("L1 conv3 640", "conv", (34, 32, 48, 640, 640)),
        ("L2 conv3 1280", "conv", (34, 16, 24, 1280, 1280)),
        ("L3 conv3 1280", "conv", (34, 8, 12, 1280, 1280)),
        ("L2 conv3 2560->1280", "conv", (34, 16, 24, 2560, 1280)),
    ]
    for name, kind, shp in cases:
        if kind == "lin":
            M, K, N, geglu = shp
            a = rnd(M, K)
            pw = ops.pack_weight(torch.randn(N, K, device=dev) / math.sqrt(K), torch.randn(N, device=dev), dev, geglu=geglu)
            out = torch.empty(M, pw.n_out, dtype=torch.float16, device=dev)
            fn = lambda: ops.gemm(a, pw, out)
            flops = 2.0 * M * K * N
        else:
            Fr, H, W, Ci, Co = shp
            a = rnd(Fr, H, W, Ci)
            pw = ops.pack_weight(torch.randn(Co, Ci, 3, 3, device=dev) / math.sqrt(9 * Ci), torch.randn(Co, device=dev), dev)
            out = torch.empty(Fr, H, W, Co, dtype=torch.float16, device=dev)
            taps = ops.conv_taps()
            fn = lambda: ops.gemm(a, pw, out, taps)
            flops = 2.0 * Fr * H * W * 9 * Ci * Co
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{name:24s} {ms:8.3f} ms  {flops / ms / 1e9:8.1f} TFLOP/s", flush=True)
    print("\n--- attention perf ---")
    for (Fr, L, heads, d) in [(34, 6144, 8, 40), (34, 1536, 8, 80), (34, 384, 8, 160)]:
        C = heads * d
        qkv = rnd(Fr, L, 3 * C)
        out = torch.empty(Fr, L, C, dtype=torch.float16, device=dev)
        fn = lambda: ops.attention(qkv[..., :C], [ops.KVSegment(qkv[..., C:2 * C], qkv[..., 2 * C:])], heads, out)
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        flops = 4.0 * Fr * heads * L * L * d
        print(f"attn F={Fr} L={L} d={d}: {ms:8.3f} ms  {flops / ms / 1e9:8.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    t0 = time.time()
    check_small()
    check_norms()
    torch.cuda.synchronize()
    check_linear(128, 64, 16)
    check_linear(256, 320, 320)
    check_linear(1000, 320, 960)
    check_linear(3264, 1280, 1280, res=True)
    check_linear(500, 320, 2560, geglu=True)
    check_linear(154, 768, 640, bias=False)
    check_linear(4096, 1280, 320, res=True)
    torch.cuda.synchronize()
    check_conv3(2, 16, 24, 64, 64)
    check_conv3(4, 8, 12, 320, 320, emb=True)
    check_conv3(2, 32, 48, 8, 320)
    check_conv3(2, 64, 96, 320, 4)
    check_conv3(2, 32, 32, 16, 32, silu=True)
    check_conv3(3, 16, 24, 960, 640)
    check_conv_s2(2, 16, 24, 64, 64)
    check_conv_s2(3, 64, 96, 320, 320)
    check_tconv(2, 17, 96, 320)
    check_tconv(1, 9, 200, 640)
    check_tconv(2, 5, 1536, 320, k=1)
    torch.cuda.synchronize()
    check_attention()
    torch.cuda.synchronize()
    print(f"checks done in {time.time() - t0:.1f}s; failures: {FAILS}")
    if "--perf" in sys.argv:
        perf()
    sys.exit(1 if FAILS else 0)

"""Developer check + timing of ccedit_attention against torch SDPA on the GPU (NOT the parity suite)."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccedit_b200 import ops  # noqa: E402

torch.manual_seed(0)
dev = "cuda"


def ref(q, k, v, heads):
    Fq, L, C = q.shape
    sp = lambda t: t.float().view(t.shape[0], t.shape[1], heads, C // heads).transpose(1, 2)
    return F.scaled_dot_product_attention(sp(q), sp(k), sp(v)).transpose(1, 2).reshape(Fq, L, C)


def check(Fr, L, Lkv, heads, d, scale=1.0):
    C = heads * d
    q, k, v = [(torch.randn(Fr, n, C, device=dev) * scale).half() for n in (L, Lkv, Lkv)]
    out = torch.empty(Fr, L, C, dtype=torch.float16, device=dev)
    ops.attention(q, [ops.KVSegment(k, v)], heads, out)
    torch.cuda.synchronize()
    r = ref(q, k, v, heads)
    err = (out.float() - r).abs().max().item()
    print(f"F={Fr} L={L} Lkv={Lkv} h={heads} d={d} scale={scale}: max|err|={err:.3e} max|ref|={r.abs().max().item():.3f} "
          f"{'OK' if err < 2e-3 * max(1.0, r.abs().max().item()) else 'BAD'}", flush=True)


def bench(Fr, L, heads, d, iters=5, scale=1.0):
    C = heads * d
    qkv = (torch.randn(Fr, L, 3 * C, device=dev) * scale).half()
    out = torch.empty(Fr, L, C, dtype=torch.float16, device=dev)
    run = lambda: ops.attention(qkv[..., :C], [ops.KVSegment(qkv[..., C:2 * C], qkv[..., 2 * C:])], heads, out)
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"attn F={Fr} L={L} d={d} scale={scale}: {ms:8.3f} ms  {4.0 * Fr * L * L * C / ms / 1e9:8.1f} TFLOP/s", flush=True)


def bench_cross(Fr, L, Lkv, heads, d, iters=10):
    C = heads * d
    q = torch.randn(Fr, L, C, device=dev).half()
    kv = torch.randn(2, Lkv, 2 * C, device=dev).half()
    out = torch.empty(Fr, L, C, dtype=torch.float16, device=dev)
    run = lambda: ops.attention(q, [ops.KVSegment(kv[..., :C], kv[..., C:], div=Fr // 2)], heads, out)
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"cross-attn F={Fr} L={L} Lkv={Lkv} d={d}: {ms * 1e3:8.1f} us  {2.0 * Fr * L * C * 2 / ms / 1e6:7.0f} GB/s (q + o)", flush=True)


def trace():
    from ccedit_b200 import _lib
    buf = torch.zeros(64, 16, dtype=torch.int64, device=dev)
    _lib.load().ccedit_gemm_trace(buf.data_ptr())
    bench(34, 6144, 8, 40, iters=1)
    _lib.load().ccedit_gemm_trace(None)
    t = buf.cpu().view(32, 32)
    t0 = int(t[0, 0])
    print("per key tile (clocks): q-tile 0 | q-tile 1 : start, waitS, ld, max, wait PV + turn, exp, arrive ; MMA warp: loop start, QK0, QK1 issued, P0 seen, PV0 issued, P1 seen, PV1 issued")
    for j in range(2, 20):
        row = []
        for qt in range(2):
            r = [int(x) - t0 for x in t[j, 8 * qt:8 * qt + 7]]
            row.append(f"{r[0]:7d} wS={r[1]-r[0]:5d} ld={r[2]-r[1]:4d} mx={r[3]-r[2]:4d} pv+turn={r[4]-r[3]:5d} exp={r[5]-r[4]:5d} ar={r[6]-r[5]:4d} tile={r[6]-r[0]:5d}")
        mm = []
        for qt in range(2):
            m = [int(x) - t0 for x in t[j, 16 + 8 * qt:16 + 8 * qt + 4]]
            mm.append(f"mma{qt} {m[0]:7d} qk=+{m[1]-m[0]:5d} p=+{m[2]-m[0]:5d} pv=+{m[3]-m[0]:5d}")
        print(f"{j:3d}: " + " | ".join(row) + " || " + " | ".join(mm))


if __name__ == "__main__":
    if os.environ.get("CCEDIT_ATTN_SCALES"):
        for sc in (0.5, 1.0, 2.0, 3.0, 4.0, 6.0):
            bench(34, 6144, 8, 40, scale=sc)
        sys.exit(0)
    if os.environ.get("CCEDIT_ATTN_TRACE"):
        trace()
        sys.exit(0)
    print("legacy" if os.environ.get("CCEDIT_ATTN_LEGACY") == "1" else "tcgen05", flush=True)
    for args in [(1, 128, 128, 1, 40), (1, 256, 256, 2, 40), (2, 300, 77, 8, 40), (3, 384, 384, 8, 40), (2, 1000, 1000, 8, 40),
                 (1, 6144, 6144, 8, 40), (2, 512, 512, 4, 64), (2, 200, 200, 4, 16), (1, 1, 1, 8, 40), (2, 640, 640, 8, 32)]:
        check(*args)
    check(2, 1024, 1024, 8, 40, scale=4.0)        # large logits: exercises the lazy rescaling
    for args in [(2, 300, 300, 8, 80), (2, 200, 77, 8, 80), (2, 96, 96, 8, 160), (1, 384, 384, 8, 160), (2, 130, 70, 4, 96),
                 (2, 500, 500, 2, 128), (1, 64, 200, 8, 24), (2, 1536, 1536, 8, 80)]:
        check(*args)
    bench(34, 6144, 8, 40)
    bench(34, 1536, 8, 40)
    bench(34, 1536, 8, 80)
    bench(34, 384, 8, 160)
    bench_cross(34, 6144, 77, 8, 40)
    bench_cross(34, 1536, 77, 8, 80)
    bench_cross(34, 384, 77, 8, 160)

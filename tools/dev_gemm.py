"""Developer timing of single GEMM shapes (CUDA events), for ncu captures of the tap-GEMM kernel."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccedit_b200 import ops  # noqa: E402

dev = "cuda"
torch.manual_seed(0)


def run(M, K, N, geglu=False, res=False, iters=5, taps=None, shape=None):
    a = (torch.randn(M, K, device=dev)).half() if shape is None else torch.randn(*shape, K, device=dev).half()
    nt = 1 if taps is None else len(taps)
    w = torch.randn(N, K, *( (3, 3) if nt == 9 else ((3,) if nt == 3 else ())), device=dev) / math.sqrt(K * nt)
    pw = ops.pack_weight(w, torch.randn(N, device=dev), dev, geglu=geglu)
    out = torch.empty(*a.shape[:-1], pw.n_out, dtype=torch.float16, device=dev)
    r = torch.randn_like(out) if res else None
    f = lambda: ops.gemm(a, pw, out, taps or ops.ONE_TAP, res1=r)
    f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    rows = a.numel() // K
    by = (a.numel() + out.numel() * (2 if res else 1)) * 2
    print(f"M={rows} K={nt}x{K} N={N} geglu={geglu} res={res} bn={pw.bn}: {ms * 1e3:8.1f} us {2.0 * rows * K * nt * N / ms / 1e9:8.1f} TF/s "
          f"{by / ms / 1e6:7.0f} GB/s", flush=True)


def trace(M, K, N, **kw):
    from ccedit_b200 import _lib
    buf = torch.zeros(64, 16, dtype=torch.int64, device=dev)
    _lib.load().ccedit_gemm_trace(buf.data_ptr())
    run(M, K, N, iters=1, **kw)
    _lib.load().ccedit_gemm_trace(None)
    t = buf.cpu()
    t0 = int(t[0, 0])
    print("tile: epi_start reads_issued acc_ready stores_done released | mma_owns first_operands mmas_issued   (clocks since start)")
    for i in range(12):
        r = [int(x) - t0 for x in t[i]]
        print(f"{i:3d}: {r[0]:8d} {r[1]:8d} {r[2]:8d} {r[3]:8d} {r[4]:8d} | {r[5]:8d} {r[7]:8d} {r[6]:8d} | blk0: ld_issue {r[8]:8d} ld_done {r[9]:8d} math_done {r[10]:8d} stored {r[11]:8d}")


if __name__ == "__main__":
    if os.environ.get("CCEDIT_GEMM_TRACE"):
        trace(208896, 320, 320)
        trace(208896, 320, 2560, geglu=True)
        sys.exit(0)
    if os.environ.get("CCEDIT_GEMM_DEV"):
        run(208896, 320, 320)
        run(208896, 320, 960)
        run(208896, 320, 2560, geglu=True)
        sys.exit(0)
    run(208896, 320, 320, res=True)
    run(208896, 320, 320)
    run(208896, 320, 2560, geglu=True)
    run(208896, 320, 960)
    run(208896, 1280, 320, res=True)
    run(52224, 640, 640, res=True)
    run(52224, 640, 5120, geglu=True)
    run(13056, 1280, 1280, res=True)
    run(0, 320, 320, taps=ops.conv_taps(), shape=(34, 64, 96))
    run(0, 320, 320, taps=ops.temporal_taps(3), res=True, shape=(2, 17, 6144))
    run(0, 16, 16, taps=ops.conv_taps(), shape=(34, 512, 768))

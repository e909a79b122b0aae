"""GPU parity of every C-ABI kernel against the oracle's leaf ops (oracle/sgm_oracle.py: F.conv2d / F.linear /
F.group_norm / F.layer_norm / SDPA in fp32) on identical fp16-rounded operands.

Tolerance (north_star): rtol = 1e-3, atol = 1e-4 - scaled by max|ref| for the absolute part, because the outputs are
stored in fp16 (relative rounding 4.9e-4) and the operands are not normalised to unit scale."""
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import assert_close

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-3, 1e-4
# attention: the probabilities are rounded to fp16 for the P.V tensor-core product (as in every flash-attention kernel,
# including the SDPA backend the reference hits on a GPU), which adds up to ~1.5e-4 absolute on outputs of magnitude 1
ATOL_ATTN = 2e-4


def close(got, ref, what, rtol=RTOL, atol=ATOL):
    assert_close(got, ref, rtol, atol * max(1.0, float(ref.abs().max())), what)


@pytest.fixture(scope="module")
def ops():
    from ccedit_b200 import ops
    return ops


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).half()


@pytest.mark.parametrize("M,K,N,geglu,res", [(128, 64, 16, False, False), (256, 320, 320, False, False),
                                              (1000, 320, 960, False, False), (3264, 1280, 1280, False, True),
                                              (500, 320, 2560, True, False), (154, 768, 640, False, False),
                                              (1, 320, 320, False, False), (4096, 1280, 320, False, True)])
def test_linear(ops, M, K, N, geglu, res):
    a, w = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=1 / math.sqrt(K))
    b = rnd(N, seed=3).float()
    r = rnd(M, N // 2 if geglu else N, seed=4) if res else None
    pw = ops.pack_weight(w.float(), b, "cuda", geglu=geglu)
    out = torch.empty(M, pw.n_out, dtype=torch.float16, device="cuda")
    ops.gemm(a.cuda(), pw, out, res1=None if r is None else r.cuda())
    ref = F.linear(a.float(), w.float(), b)
    if geglu:
        v, g = ref.chunk(2, dim=-1)
        ref = v * F.gelu(g)
    if res:
        ref = ref + r.float()
    close(out, ref, f"linear {M}x{K}x{N}")


@pytest.mark.parametrize("M,K,N,nres,geglu", [(40000, 320, 320, 1, False), (40000, 320, 320, 0, False),
                                               (30000, 320, 1280, 1, False), (20000, 640, 640, 2, False),
                                               (33000, 320, 2560, 0, True), (1000, 1280, 320, 1, False),
                                               (20011, 320, 960, 0, False), (5000, 64, 64, 1, False),
                                               (5000, 64, 32, 1, False)])
def test_linear_staged_epilogue(ops, M, K, N, nres, geglu):
    """Small-K GEMMs take the shared-memory + TMA-store epilogue (gemm_tc.cu: epilogue_staged): several tiles per CTA
    (both tile buffers, both barrier phases), ragged last tile, strided output / residual views of wider buffers."""
    a, w = rnd(M, K, seed=21), rnd(N, K, seed=22, scale=1 / math.sqrt(K))
    b = rnd(N, seed=23).float()
    pw = ops.pack_weight(w.float(), b, "cuda", geglu=geglu)
    no = pw.n_out
    wide = torch.full((M, no + 64), 7.0, dtype=torch.float16, device="cuda")
    out = wide[:, 32:32 + no]
    r1w = rnd(M, no + 16, seed=24).cuda()
    r1 = r1w[:, 8:8 + no] if nres >= 1 else None
    r2 = rnd(M, no, seed=25).cuda() if nres >= 2 else None
    ops.gemm(a.cuda(), pw, out, res1=r1, res2=r2)
    ref = F.linear(a.float(), w.float(), b)
    if geglu:
        v, g = ref.chunk(2, dim=-1)
        ref = v * F.gelu(g)
    if r1 is not None:
        ref = ref + r1.float().cpu()
    if r2 is not None:
        ref = ref + r2.float().cpu()
    close(out, ref, f"staged linear {M}x{K}x{N} nres={nres}")
    assert bool((wide[:, :32] == 7).all()) and bool((wide[:, 32 + no:] == 7).all()), "wrote outside the output view"


def test_linear_staged_epilogue_rowbias_per_row_group(ops):
    """Time-embedding style row bias whose row changes inside a warp's 32 rows (non-uniform path) and per tile."""
    M, K, N, div = 3000, 320, 320, 20
    a, w = rnd(M, K, seed=31), rnd(N, K, seed=32, scale=1 / math.sqrt(K))
    b = rnd(N, seed=33).float()
    rb = rnd((M + div - 1) // div, N, seed=34).float()
    pw = ops.pack_weight(w.float(), b, "cuda")
    out = torch.empty(M, N, dtype=torch.float16, device="cuda")
    ops.gemm(a.cuda(), pw, out, rowbias=rb.cuda(), rb_dim=0, rb_div=div, silu=True)
    ref = F.silu(F.linear(a.float(), w.float(), b) + rb.repeat_interleave(div, 0)[:M])
    close(out, ref, "staged linear + rowbias + SiLU")


@pytest.mark.parametrize("Fr,H,W,Cin,Cout,emb,silu", [(2, 16, 24, 64, 64, False, False), (4, 8, 12, 320, 320, True, False),
                                                      (2, 32, 48, 8, 320, False, False), (2, 64, 96, 320, 4, False, False),
                                                      (2, 32, 32, 16, 32, False, True), (3, 16, 24, 960, 640, False, False),
                                                      (1, 5, 7, 64, 64, False, False)])
def test_conv3x3(ops, Fr, H, W, Cin, Cout, emb, silu):
    x, w = rnd(Fr, Cin, H, W, seed=5), rnd(Cout, Cin, 3, 3, seed=6, scale=1 / math.sqrt(9 * Cin))
    b = rnd(Cout, seed=7).float()
    T = 2 if Fr % 2 == 0 else 1
    pw = ops.pack_weight(w.float(), b, "cuda")
    rb = rnd(Fr // T, pw.n, seed=8).float() if emb else None
    out = torch.empty(Fr, H, W, pw.n, dtype=torch.float16, device="cuda")
    ops.gemm(x.permute(0, 2, 3, 1).contiguous().cuda(), pw, out, ops.conv_taps(),
             rowbias=None if rb is None else rb.cuda(), rb_dim=2, rb_div=T, silu=silu)
    ref = F.conv2d(x.float(), w.float(), b, padding=1)
    if emb:
        ref = ref + rb[:, :Cout].repeat_interleave(T, 0)[:, :, None, None]
    if silu:
        ref = F.silu(ref)
    close(out[..., :Cout].permute(0, 3, 1, 2), ref, f"conv3x3 {Cin}->{Cout}")


@pytest.mark.parametrize("M,K,N,geglu,bias", [(5000, 320, 320, False, False), (3001, 320, 960, False, False),
                                               (4000, 320, 2560, True, True), (700, 1280, 10240, True, True),
                                               (2000, 640, 640, False, True), (130, 1280, 3840, False, False)])
def test_linear_with_folded_layernorm(ops, M, K, N, geglu, bias):
    """attention.py:667-669 + to_q / to_k / to_v / ff.net.0: LayerNorm folded into the consuming GEMM - statistics pass
    (ccedit_layernorm_stats) + gamma-scaled weight + per-row epilogue - against F.layer_norm + F.linear in fp32.  The rows
    carry a large common offset so that the mean term really has to cancel."""
    g = torch.Generator().manual_seed(70)
    x = (torch.randn(M, K, generator=g) * 1.5 + 3.0 * torch.randn(M, 1, generator=g)).half()
    w = rnd(N, K, seed=71, scale=1 / math.sqrt(K))
    b = rnd(N, seed=72).float() if bias else None
    gamma, beta = 1 + 0.3 * rnd(K, seed=73).float(), 0.2 * rnd(K, seed=74).float()
    pw = ops.pack_weight(w.float(), b, "cuda", geglu=geglu, ln_gamma=gamma, ln_beta=beta)
    xc = x.cuda()
    st = ops.layernorm_stats(xc)
    ref_mean, ref_var = x.float().mean(1), x.float().var(1, unbiased=False)
    close(st[:, 0], ref_mean, "row mean", rtol=1e-5, atol=1e-5)
    close(st[:, 1], (ref_var + 1e-5).rsqrt(), "row rstd", rtol=1e-5, atol=1e-5)
    out = ops.gemm(xc, pw, torch.empty(M, pw.n_out, dtype=torch.float16, device="cuda"), rowstats=st)
    # identical fp16-rounded operands (see the module docstring): the kernel's weight operand is fp16(gamma o W)
    xhat = F.layer_norm(x.float(), (K,))
    ref = F.linear(xhat, (w.float() * gamma[None, :]).half().float(), (w.float() @ beta) + (0 if b is None else b))
    if geglu:
        v, gt = ref.chunk(2, dim=-1)
        ref = v * F.gelu(gt)
    close(out, ref, f"LN-folded linear {M}x{K}x{N} geglu={geglu}")
    # and against the un-fused fp32 formulation: only the rounding of gamma o W to fp16 separates the two
    full = F.linear(F.layer_norm(x.float(), (K,), gamma, beta), w.float(), b)
    if geglu:
        v, gt = full.chunk(2, dim=-1)
        full = v * F.gelu(gt)
    assert float((out.float().cpu() - full).abs().max()) <= 4e-3 * max(1.0, float(full.abs().max()))


@pytest.mark.parametrize("M,K,N,res", [(5000, 320, 320, True), (3001, 640, 640, False), (2000, 1280, 1280, True),
                                        (777, 2560, 640, True), (300, 320, 96, False)])
def test_linear_emits_layernorm_statistics_of_its_output(ops, M, K, N, res):
    """ccedit_gemm_desc.stats_out: the epilogue's per-row partial (sum, sum of squares) + ccedit_layernorm_stats_combine
    equal the LayerNorm statistics of the GEMM's output (staged and direct epilogues, with and without residual)."""
    a, w = rnd(M, K, seed=81), rnd(N, K, seed=82, scale=1 / math.sqrt(K))
    b = (rnd(N, seed=83).float() + 0.5)
    r = rnd(M, N, seed=84).cuda() if res else None
    pw = ops.pack_weight(w.float(), b, "cuda")
    sp = torch.empty(M, ops.stats_slots(pw), 2, dtype=torch.float32, device="cuda")
    out = ops.gemm(a.cuda(), pw, torch.empty(M, N, dtype=torch.float16, device="cuda"), res1=r, stats_out=sp)
    st = ops.layernorm_stats_combine(sp, N)
    y = out.float().cpu()
    close(st[:, 0], y.mean(1), "mean from the producing epilogue", rtol=1e-3, atol=1e-3)
    close(st[:, 1], (y.var(1, unbiased=False) + 1e-5).rsqrt(), "rstd from the producing epilogue", rtol=2e-3, atol=1e-3)
    close(st, ops.layernorm_stats(out), "vs. the statistics pass", rtol=2e-3, atol=1e-3)
    # the consuming GEMM can finish the partial sums itself (rowstats = [M, P, 2]): same result as with combined statistics
    w2 = rnd(320, N, seed=85, scale=1 / math.sqrt(N))
    gamma, beta = 1 + 0.3 * rnd(N, seed=86).float(), 0.2 * rnd(N, seed=87).float()
    pw2 = ops.pack_weight(w2.float(), None, "cuda", ln_gamma=gamma, ln_beta=beta)
    y1 = ops.gemm(out, pw2, torch.empty(M, 320, dtype=torch.float16, device="cuda"), rowstats=st)
    y2 = ops.gemm(out, pw2, torch.empty(M, 320, dtype=torch.float16, device="cuda"), rowstats=sp)
    close(y2, y1.float().cpu(), "partial statistics finished in the consuming epilogue", rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("Fr,H,W", [(2, 32, 128), (1, 40, 70), (3, 16, 64), (1, 5, 9)])
def test_hint_stem_first_two_layers_fused(ops, Fr, H, W):
    """controlmodel.py:215-219: conv3x3(3->16)+SiLU+conv3x3(16->16)+SiLU in one kernel (csrc/hint_stem.cu), ragged tiles."""
    x = rnd(Fr, 3, H, W, seed=60)
    w0, b0 = rnd(16, 3, 3, 3, seed=61, scale=1 / math.sqrt(27)), rnd(16, seed=62).float()
    w1, b1 = rnd(16, 16, 3, 3, seed=63, scale=1 / math.sqrt(144)), rnd(16, seed=64).float()
    x_cl = torch.zeros(Fr, H, W, 8, dtype=torch.float16)
    x_cl[..., :3] = x.permute(0, 2, 3, 1)
    p0 = ops.pack_hint_stem_weight(w0.float(), b0, "cuda", 8, 80)
    p1 = ops.pack_hint_stem_weight(w1.float(), b1, "cuda", 16, 144)
    out = ops.hint_stem01(x_cl.cuda(), *p0, *p1)
    mid = F.silu(F.conv2d(x.float(), w0.float(), b0, padding=1)).half().float()   # the kernel keeps layer 0 in fp16
    ref = F.silu(F.conv2d(mid, w1.float(), b1, padding=1))
    close(out.permute(0, 3, 1, 2), ref, f"hint stem layers 0+1 {Fr}x{H}x{W}")


@pytest.mark.parametrize("Fr,H,W", [(2, 32, 128), (1, 40, 70), (3, 16, 64), (1, 6, 10), (1, 2, 2)])
def test_hint_stem_layers_two_and_three_fused(ops, Fr, H, W):
    """controlmodel.py:220-223: conv3x3(16->32, stride 2)+SiLU+conv3x3(32->32)+SiLU in one kernel (csrc/hint_stem.cu),
    ragged tiles, image borders inside and outside the halo."""
    x = rnd(Fr, 16, H, W, seed=65)
    w2, b2 = rnd(32, 16, 3, 3, seed=66, scale=1 / math.sqrt(144)), rnd(32, seed=67).float()
    w3, b3 = rnd(32, 32, 3, 3, seed=68, scale=1 / math.sqrt(288)), rnd(32, seed=69).float()
    p2 = ops.pack_hint_stem_weight(w2.float(), b2, "cuda", 16, 144)
    p3 = ops.pack_hint_stem_weight(w3.float(), b3, "cuda", 32, 288)
    out = ops.hint_stem23(x.permute(0, 2, 3, 1).contiguous().cuda(), *p2, *p3)
    mid = F.silu(F.conv2d(x.float(), w2.float(), b2, stride=2, padding=1)).half().float()   # the kernel keeps layer 2 in fp16
    ref = F.silu(F.conv2d(mid, w3.float(), b3, padding=1))
    assert tuple(out.shape) == (Fr, H // 2, W // 2, 32)
    close(out.permute(0, 3, 1, 2), ref, f"hint stem layers 2+3 {Fr}x{H}x{W}")


def test_conv3x3_time_embedding_rows_with_padded_tiles(ops):
    """34 frames of 2x2 pixels: the 128-row tiles pad the frame axis to 64, and the padded rows must not index past the
    [B, Cout] time-embedding rows (regression: an out-of-bounds read in the epilogue's bias staging)."""
    Fr, T, H, W, C = 34, 17, 2, 2, 64
    x, w = rnd(Fr, C, H, W, seed=50), rnd(C, C, 3, 3, seed=51, scale=1 / math.sqrt(9 * C))
    b = rnd(C, seed=52).float()
    pw = ops.pack_weight(w.float(), b, "cuda")
    rb = rnd(Fr // T, C, seed=53).float()
    out = torch.empty(Fr, H, W, C, dtype=torch.float16, device="cuda")
    ops.gemm(x.permute(0, 2, 3, 1).contiguous().cuda(), pw, out, ops.conv_taps(), rowbias=rb.cuda(), rb_dim=2, rb_div=T)
    ref = F.conv2d(x.float(), w.float(), b, padding=1) + rb.repeat_interleave(T, 0)[:, :, None, None]
    close(out.permute(0, 3, 1, 2), ref, "conv3x3 + time-embedding rows, padded frame axis")


@pytest.mark.parametrize("Fr,H,W,C", [(2, 16, 24, 64), (3, 64, 96, 320), (2, 8, 12, 1280), (2, 32, 48, 16)])
def test_conv3x3_stride2(ops, Fr, H, W, C):
    x, w = rnd(Fr, C, H, W, seed=9), rnd(C, C, 3, 3, seed=10, scale=1 / math.sqrt(9 * C))
    b = rnd(C, seed=11).float()
    pw = ops.pack_weight(w.float(), b, "cuda")
    planes = ops.parity_split(x.permute(0, 2, 3, 1).contiguous().cuda())
    out = torch.empty(Fr, 1, H // 2, W // 2, C, dtype=torch.float16, device="cuda")
    ops.gemm(planes, pw, out, ops.conv_s2_taps())
    ref = F.conv2d(x.float(), w.float(), b, stride=2, padding=1)
    close(out[:, 0].permute(0, 3, 1, 2), ref, f"conv3x3/s2 C={C}")


@pytest.mark.parametrize("B,T,HW,C", [(2, 17, 50, 320), (1, 3, 24, 1280), (2, 1, 16, 64), (1, 33, 20, 640)])
def test_temporal_conv_k3_with_identity(ops, B, T, HW, C):
    """spatial_temporal_forward's temporal half: y + conv1d_k3(y) over T per pixel, zero padded at the clip ends."""
    y, w = rnd(B, T, HW, C, seed=12), rnd(C, C, 3, seed=13, scale=1 / math.sqrt(3 * C))
    b = rnd(C, seed=14).float()
    pw = ops.pack_weight(w.float(), b, "cuda")
    yc = y.cuda()
    out = ops.gemm(yc, pw, torch.empty_like(yc), ops.temporal_taps(3), res1=yc)
    z = y.float().permute(0, 2, 3, 1).reshape(B * HW, C, T)                      # (b hw) c t
    ref = (z + F.conv1d(z, w.float(), b, padding=1)).reshape(B, HW, C, T).permute(0, 3, 1, 2)
    close(out, ref, f"temporal conv T={T} C={C}")


@pytest.mark.parametrize("Fr,HW,C,eps,silu", [(2, 96, 320, 1e-5, True), (3, 6144, 320, 1e-6, False),
                                              (2, 24, 2560, 1e-5, True), (1, 1, 64, 1e-5, False), (2, 1536, 960, 1e-5, True)])
def test_groupnorm_spatial(ops, Fr, HW, C, eps, silu):
    x = rnd(Fr, HW, C, seed=15) + 0.5
    g, b = rnd(C, seed=16).float() * 0.1 + 1, rnd(C, seed=17).float() * 0.1
    out = ops.groupnorm_spatial(x.cuda(), g.cuda(), b.cuda(), eps, silu)
    ref = F.group_norm(x.float().permute(0, 2, 1), 32, g, b, eps)
    if silu:
        ref = F.silu(ref)
    close(out, ref.permute(0, 2, 1), f"GN spatial C={C}")


def test_groupnorm_spatial_scratch_is_shape_independent(ops):
    """The single-pass kernel's arrival counters live at a fixed place of the per-stream scratch: launches with different
    frame counts / sizes back to back must not see each other's partial sums as counters (regression)."""
    for Fr, HW, C in [(34, 384, 320), (2, 1536, 640), (5, 96, 1280), (34, 96, 320), (1, 6144, 320), (3, 200, 64)]:
        x = rnd(Fr, HW, C, seed=Fr + HW) + 0.25
        gamma, beta = rnd(C, seed=16).float(), rnd(C, seed=17).float()
        for _ in range(2):
            y = ops.groupnorm_spatial(x.cuda(), gamma.cuda(), beta.cuda(), 1e-5, True)
        ref = F.silu(F.group_norm(x.float().permute(0, 2, 1), 32, gamma, beta, 1e-5)).permute(0, 2, 1)
        close(y, ref, f"GN spatial {Fr}x{HW}x{C}")


@pytest.mark.parametrize("B,T,HW,C,silu", [(2, 17, 40, 320, True), (1, 3, 24, 1280, False), (2, 33, 10, 640, True),
                                           (1, 1, 7, 64, False)])
def test_groupnorm_temporal(ops, B, T, HW, C, silu):
    x = rnd(B, T, HW, C, seed=18) + 0.25
    g, b = rnd(C, seed=19).float() * 0.1 + 1, rnd(C, seed=20).float() * 0.1
    out = ops.groupnorm_temporal(x.cuda(), g.cuda(), b.cuda(), 1e-5, silu)
    z = x.float().permute(0, 2, 3, 1).reshape(B * HW, C, T)
    ref = F.group_norm(z, 32, g, b, 1e-5)
    if silu:
        ref = F.silu(ref)
    close(out, ref.reshape(B, HW, C, T).permute(0, 3, 1, 2), f"GN temporal C={C} T={T}")


@pytest.mark.parametrize("M,C", [(513, 1280), (64, 64), (1000, 320), (7, 640)])
def test_layernorm(ops, M, C):
    x = rnd(M, C, seed=21) * 2 + 0.3
    g, b = rnd(C, seed=22).float() * 0.1 + 1, rnd(C, seed=23).float() * 0.1
    out = ops.layernorm(x.cuda(), g.cuda(), b.cuda())
    close(out, F.layer_norm(x.float(), (C,), g, b, 1e-5), f"LN C={C}")


def _sdpa(q, k, v, heads):
    Fq, L, C = q.shape
    sp = lambda t: t.float().view(t.shape[0], t.shape[1], heads, C // heads).transpose(1, 2)
    o = F.scaled_dot_product_attention(sp(q), sp(k), sp(v))
    return o.transpose(1, 2).reshape(Fq, L, C)


@pytest.mark.parametrize("Fr,L,Lkv,heads,d", [(3, 384, 384, 8, 40), (2, 200, 200, 8, 80), (2, 96, 96, 8, 160),
                                              (4, 300, 77, 8, 40), (2, 128, 128, 4, 16), (2, 1536, 1536, 8, 80),
                                              (1, 1, 1, 8, 40), (1, 6144, 6144, 8, 40)])
def test_attention(ops, Fr, L, Lkv, heads, d):
    C = heads * d
    q, k, v = rnd(Fr, L, C, seed=24), rnd(Fr, Lkv, C, seed=25), rnd(Fr, Lkv, C, seed=26)
    out = torch.empty(Fr, L, C, dtype=torch.float16, device="cuda")
    ops.attention(q.cuda(), [ops.KVSegment(k.cuda(), v.cuda())], heads, out)
    close(out, _sdpa(q, k, v, heads), f"attention L={L} Lkv={Lkv} d={d}", atol=ATOL_ATTN)


def test_attention_text_keys_shared_by_frames(ops):
    """Text cross-attention: all T frames of a batch entry read the same 77 keys (attention.py:1159-1163)."""
    B, T, L, heads, d = 2, 3, 100, 8, 40
    C = heads * d
    q, k, v = rnd(B * T, L, C, seed=27), rnd(B, 77, C, seed=28), rnd(B, 77, C, seed=29)
    out = torch.empty(B * T, L, C, dtype=torch.float16, device="cuda")
    ops.attention(q.cuda(), [ops.KVSegment(k.cuda(), v.cuda(), div=T)], heads, out)
    ref = _sdpa(q, k.repeat_interleave(T, 0), v.repeat_interleave(T, 0), heads)
    close(out, ref, "text attention", atol=ATOL_ATTN)


@pytest.mark.parametrize("B,T,L,Lkv,heads,d", [(2, 17, 6144, 77, 8, 40), (2, 17, 1536, 77, 8, 80), (2, 17, 384, 77, 8, 160),
                                               (2, 17, 96, 77, 8, 160), (2, 9, 2001, 77, 8, 40), (3, 5, 333, 128, 8, 80),
                                               (2, 33, 50, 1, 8, 40), (1, 64, 100, 16, 4, 40), (2, 8, 130, 100, 2, 160)])
def test_attention_short_keys_shared_by_frames(ops, B, T, L, Lkv, heads, d):
    """Text cross-attention at the sizes of the network call (and ragged ones): <= 128 keys shared by the T frames of a
    batch entry take short_kv_attn_kernel - K / V of a head group resident in shared memory, per-warp streaming of the
    query rows (attention.cu); partial last row block, 1 / 16 / 128 keys, 1 / 2 / 4 heads per group."""
    C = heads * d
    q, k, v = rnd(B * T, L, C, seed=57), rnd(B, Lkv, C, seed=58), rnd(B, Lkv, C, seed=59)
    out = torch.empty(B * T, L, C, dtype=torch.float16, device="cuda")
    ops.attention(q.cuda(), [ops.KVSegment(k.cuda(), v.cuda(), div=T)], heads, out)
    ref = _sdpa(q, k.repeat_interleave(T, 0), v.repeat_interleave(T, 0), heads)
    close(out, ref, f"text attention, short keys (L={L}, Lkv={Lkv}, d={d})", atol=ATOL_ATTN)


def test_attention_short_keys_strided_views_and_ragged_frames(ops):
    """q and the output as column slices of wider buffers (fused q|k|v projection layout), k / v as slices of one [.., 2C]
    projection, and a frame count that is not a multiple of the frames per key set (the last set serves fewer frames)."""
    Fr, div, L, Lkv, heads, d = 7, 3, 400, 77, 8, 40
    C = heads * d
    qkv, kv = rnd(Fr, L, 3 * C, seed=71).cuda(), rnd(3, Lkv, 2 * C, seed=72).cuda()
    buf = torch.zeros(Fr, L, 2 * C, dtype=torch.float16, device="cuda")
    q, k, v, out = qkv[..., C:2 * C], kv[..., :C], kv[..., C:], buf[..., C:]
    ops.attention(q, [ops.KVSegment(k, v, div=div)], heads, out)
    idx = torch.arange(Fr) // div
    ref = _sdpa(q.cpu().contiguous(), k.cpu()[idx].contiguous(), v.cpu()[idx].contiguous(), heads)
    close(out, ref, "text attention on strided views", atol=ATOL_ATTN)
    assert torch.count_nonzero(buf[..., :C]) == 0, "columns outside the output view were written"


@pytest.mark.parametrize("Fr,L,heads,d", [(120, 200, 8, 40), (450, 200, 8, 40)])
def test_attention_many_heads_per_cta(ops, Fr, L, heads, d):
    """One or two key tiles per head and enough query tiles: flash_attn_tc2_kernel takes several heads per CTA in a row
    (MH = true: next-head Q prefetch, per-head statistics reset, O re-initialised by the first P.V) - 2 and 8 heads per
    CTA here (per-frame keys, so the short-key kernel does not apply)."""
    C = heads * d
    q, k, v = rnd(Fr, L, C, seed=61), rnd(Fr, L, C, seed=62), rnd(Fr, L, C, seed=63)
    out = torch.empty(Fr, L, C, dtype=torch.float16, device="cuda")
    ops.attention(q.cuda(), [ops.KVSegment(k.cuda(), v.cuda())], heads, out)
    close(out, _sdpa(q, k, v, heads), f"attention, several heads per CTA (F={Fr}, L={L})", atol=ATOL_ATTN)


def test_attention_center_self_two_segments(ops):
    """cfca: K/V = cat([centre-frame tokens (repeated over t), own tokens]) (attention.py:1323-1336)."""
    B, T, L, heads, d = 2, 3, 70, 8, 40
    C = heads * d
    q, k, v = rnd(B * T, L, C, seed=30), rnd(B * T, L, C, seed=31), rnd(B * T, L, C, seed=32)
    out = torch.empty(B * T, L, C, dtype=torch.float16, device="cuda")
    kc, vc = k.cuda(), v.cuda()
    ops.attention(q.cuda(), [ops.KVSegment(kc, vc, div=T, mul=T, add=T // 2), ops.KVSegment(kc, vc)], heads, out)
    ctr = lambda t: t.view(B, T, L, C)[:, T // 2].repeat_interleave(T, 0)
    ref = _sdpa(q, torch.cat([ctr(k), k], 1), torch.cat([ctr(v), v], 1), heads)
    close(out, ref, "center_self attention", atol=ATOL_ATTN)


@pytest.mark.parametrize("B,T,HW,heads,d", [(2, 17, 50, 8, 40), (1, 9, 30, 8, 160), (2, 33, 20, 8, 80), (1, 1, 5, 8, 40),
                                           (1, 33, 6, 8, 160), (1, 40, 3, 4, 16), (1, 17, 9, 4, 96), (2, 16, 11, 8, 24), (1, 64, 2, 8, 40),
                                           # the nine (d, T) pairs of the network and its sweep: temporal_attn_fixed_kernel
                                           (2, 9, 37, 8, 40), (1, 33, 21, 8, 40), (2, 9, 13, 8, 80), (2, 17, 29, 8, 80),
                                           (2, 17, 19, 8, 160)])
def test_temporal_attention(ops, B, T, HW, heads, d):
    C = heads * d
    q, k, v = rnd(B, T, HW, C, seed=33), rnd(B, T, HW, C, seed=34), rnd(B, T, HW, C, seed=35)
    out = torch.empty(B, T, HW, C, dtype=torch.float16, device="cuda")
    ops.temporal_attention(q.cuda(), k.cuda(), v.cuda(), heads, out)
    tok = lambda t: t.permute(0, 2, 1, 3).reshape(B * HW, T, C)                   # (b hw) t c
    ref = _sdpa(tok(q), tok(k), tok(v), heads).view(B, HW, T, C).permute(0, 2, 1, 3)
    close(out, ref, f"temporal attention T={T} d={d}", atol=ATOL_ATTN)


def test_layout_and_small_kernels(ops):
    x = torch.randn(2, 3, 4, 16, 24, generator=torch.Generator().manual_seed(36))
    cl = ops.ncthw_to_cl(x.cuda(), 8, pre=1.0, mul=-0.5, add=1.0)
    ref = torch.zeros(2, 4, 16, 24, 8)
    ref[..., :3] = (1 - (x + 1) / 2.0).permute(0, 2, 3, 4, 1)                      # wrappers.py:160-162, as torch computes it
    assert torch.equal(cl.cpu(), ref.half()), "fused hint transform must equal transform-in-torch + fp16 conversion bit for bit"
    t = torch.tensor([0.0, 417.0, 999.0])
    from oracle.sgm_oracle import timestep_embedding
    assert_close(ops.timestep_embedding(t.cuda(), 320), timestep_embedding(t, 320), 1e-4, 2e-4, "timestep embedding")
    xs, w, b = torch.randn(3, 320), rnd(1280, 320, seed=37, scale=0.05), torch.randn(1280) * 0.1
    got = ops.linear_small(xs.cuda(), w.cuda(), b.cuda(), act_in=True, act_out=True)
    assert_close(got, F.silu(F.linear(F.silu(xs), w.float(), b)), 1e-4, 1e-5, "linear_small")
    u = rnd(2, 6, 8, 64, seed=38)
    assert torch.equal(ops.upsample_nearest2x(u.cuda()).cpu(),
                       F.interpolate(u.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1).half())
    a, bb = rnd(2, 3, 10, 64, seed=39), rnd(2, 3, 10, 64, seed=40)
    dst = torch.zeros(2, 3, 10, 128, dtype=torch.float16, device="cuda")
    ops.add_rows(a.cuda(), bb.cuda(), dst[..., 64:])
    close(dst[..., 64:], a.float() + bb.float(), "add_rows")
    assert float(dst[..., :64].abs().max()) == 0.0
    x5, y4 = rnd(2, 5, 12, 64, seed=41), rnd(2, 12, 64, seed=42)
    x5c = x5.cuda()
    ops.add_center_frame(x5c, y4.cuda())
    ref = x5.float().clone()
    ref[:, 2] += y4.float()
    close(x5c, ref, "add_center_frame")
    c32 = torch.randn(2, 77, 768)
    assert torch.equal(ops.to_half(c32.cuda()).cpu(), c32.half())


def test_errors_are_loud(ops):
    """Bad arguments raise with the library's message instead of computing something else."""
    a = torch.zeros(128, 64, dtype=torch.float16, device="cuda")
    pw = ops.pack_weight(torch.zeros(16, 64), None, "cuda")
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.gemm(a.cpu(), pw, torch.empty(128, 16, dtype=torch.float16))
    with pytest.raises(RuntimeError):
        ops.groupnorm_spatial(torch.zeros(1, 4, 48, dtype=torch.float16, device="cuda"), torch.ones(48).cuda(),
                              torch.zeros(48).cuda(), 1e-5, False)     # 48 channels: not a multiple of 32

// Flash attention on tcgen05 tensor cores with TMEM-resident S / P / O (sm_100a), head dims up to 64 (d = 40 is 88 % of
// the attention FLOPs of the CCEdit UNet: 7 x (34 frames x 8 heads x 6144^2) per network call).
//
// One CTA = 256 query rows (two 128-row tiles, each with its own softmax warpgroup) of one (frame, head):
//   warp 0      TMA producer: K and V tiles [128 keys x 64 channels] through 3-D tensor maps (SWIZZLE_128B; rows past the
//               end of the segment are zero-filled by TMA), 3-stage mbarrier ring
//   warp 1      TMEM allocator + tcgen05.mma issue (the whole converged warp runs the issue code with warp-uniform
//               operands and one elected lane issues: from inside `if (lane == 0)` every tcgen05 instruction costs an
//               ELECT / R2UR.BROADCAST / branch loop of ~100 clocks, which made this warp the bottleneck):
//                 S_q  = Q_q . K_j^T      (A, B from shared memory, K-major, fp32 accumulator in TMEM)
//                 O_q += P_q . V_j        (A = P from TMEM, B = V from shared memory, MN-major)
//   warps 2-5   softmax of query tile 0 (thread = one query row = one TMEM lane): tcgen05.ld S -> running max with
//   warps 6-9   lazy rescaling -> exp2 -> fp16 P written to TMEM with tcgen05.st -> row sums; final O / l epilogue
// The two query tiles share every K/V tile and interleave on the tensor pipe.  The kernel is bound by the MUFU exp2 rate
// (160 tensor FLOPs per exponential at d = 40), so everything is arranged to keep the softmax warps busy:
//   * S_q(j+1) = Q_q K_{j+1}^T is issued as soon as the softmax warpgroup has pulled S_q(j) into registers (s_free), not
//     after P_q(j): the MMA round trip is off the softmax -> softmax critical path (it used to be ~40 % of a tile);
//   * the wait for P_q(j-1).V_{j-1} (o_done: P and O are free again) sits right before the first P store of tile j;
//   * row maxima with 3-input max, and EMU of every 4 exponentials evaluated on the FMA pipe (Cody-Waite + degree-3
//     polynomial, rel. error 8.8e-5, below the fp16 rounding of P) instead of MUFU.EX2.
//
// Replaces F.scaled_dot_product_attention (attention.py:444-448) for spatial self-attention, text cross-attention and
// the two-segment centre+self context of SpatialTransformer3DCA (attention.py:1323-1336).
#include "common.cuh"
#include "../../include/ccedit_b200.h"

#include <atomic>
#include <cstdlib>
#include <mutex>

namespace ccedit {
extern std::atomic<long long> g_launch_count;
extern long long* g_trace_buf;

constexpr int kTcThreads = 352;
constexpr int kTcMmaWarp1 = 10;              // second MMA-issuing warp (query tile 1)
constexpr int kTcTile = 128;                 // query rows per tile / keys per tile
constexpr int kTcStages = 3;
constexpr int kTcDefaultStagger = 2;
constexpr int kTcDefaultEmu = 0;             // exponentials per group of 4 evaluated on the FMA pipe
constexpr int kTcTileBytes = 128 * 128;      // [128 rows][64 halves], SWIZZLE_128B
// TMEM columns (all 512): S_q fp32 scores (128 keys), P_q the probabilities as packed fp16 (64 columns), O_q fp32 output.
// P does not alias S: QK^T of the next key tile is issued into S_q right behind P.V, and a later MMA writing columns that
// an earlier MMA still reads as its A operand is not a documented interlock.
constexpr uint32_t kTcColS0 = 0, kTcColS1 = 128, kTcColP0 = 256, kTcColP1 = 320, kTcColO0 = 384, kTcColO1 = 448;
// Head dims <= 48 leave 32 columns free: O_1 moves down to 432 and the first 32 keys of P (16 columns) get a second
// buffer per query tile, used on odd key tiles - the softmax warps can then start writing P(j+1) while P(j).V is still
// running (the wait for it moves from the first to the second 32-key block; clock trace: ~400 clocks per tile).
constexpr uint32_t kTcColO1Small = 432, kTcColX0 = 480, kTcColX1 = 496;

struct FaTcParams {
  const __half* q;
  long long ldq, q_fs;
  __half* o;
  long long ldo, o_fs;
  int nseg;
  int lkv[2], kv_div[2], kv_mul[2], kv_add[2];
  int ntile[2];
  int lq, d;
  float scale_log2;
  int stagger;
  long long* trace;   // diagnostics (ccedit_gemm_trace): per-tile phase clocks of CTA 0, or nullptr
};

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_warp(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n\t}"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// named barriers (ids 1, 2) hand the "exponential phase" back and forth between the two softmax warpgroups
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float max3_f(float a, float b, float c) {
  float y;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
  return y;
}
// 2^x on the FMA pipe: x = n + f (n = floor(x) by a round-down add of 1.5 * 2^23, f in [0, 1)), 2^f by a degree-3
// minimax polynomial (max rel. error 8.8e-5), n added into the exponent field.  x is clamped to >= -126 (result ~1e-38,
// which rounds to a zero probability); masked scores (-inf) take this path safely.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.f);
  float xr;
  asm("add.rm.ftz.f32 %0, %1, %2;" : "=f"(xr) : "f"(x), "f"(12582912.f));
  const float f = x - (xr - 12582912.f);
  float p = fmaf(0.077119089663028717041015625f, f, 0.227564394474029541015625f);
  p = fmaf(p, f, 0.695146143436431884765625f);
  p = fmaf(p, f, 1.0f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(xr) << 23));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// MN-major operand tile stored as [rows = K index][64 contiguous MN elements], SWIZZLE_128B, 8-row groups 1024 B apart
// (cute::UMMA canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units): SBO = 1024 B, LBO = atom stride along MN.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

template <int KSTEPS, int EMU>  // head dim d <= 16*KSTEPS (KSTEPS in 1..4); output tile width NO = 16*KSTEPS columns
__global__ void __launch_bounds__(kTcThreads, 1)
flash_attn_tc_kernel(const __grid_constant__ CUtensorMap tmK0, const __grid_constant__ CUtensorMap tmV0,
                     const __grid_constant__ CUtensorMap tmK1, const __grid_constant__ CUtensorMap tmV1,
                     const __grid_constant__ FaTcParams p) {
  constexpr int NO = 16 * KSTEPS;
  constexpr bool kPx = NO <= 48;             // second buffer for the first 32-key block of P (see kTcColX0)
  extern __shared__ uint8_t fa_smem_raw[];
  const uint32_t raw_addr = smem_u32(fa_smem_raw);
  uint8_t* smem = fa_smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sQ = smem;                                    // 2 tiles
  uint8_t* sKV = smem + 2 * kTcTileBytes;                // stages x (K tile, V tile)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + kTcStages * 2 * kTcTileBytes);
  uint64_t* kv_full = bars;                              // [stages]
  uint64_t* kv_empty = bars + kTcStages;                 // [stages]
  uint64_t* s_full = bars + 2 * kTcStages;               // [2]  S_q ready (also: all earlier MMAs of tile q done)
  uint64_t* p_full = s_full + 2;                         // [2]  P_q written
  uint64_t* q_full = p_full + 2;                         // [1]
  uint64_t* s_free = q_full + 1;                         // [2]  S_q has been read into registers
  uint64_t* o_done = s_free + 2;                         // [2]  P_q.V done: O_q consistent, P_q may be overwritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * (2 * kTcTile), head = blockIdx.y, f = blockIdx.z;
  const int d = p.d;
  const int nq = (q0 + kTcTile < p.lq) ? 2 : 1;          // valid query tiles of this CTA
  const int ntiles = p.ntile[0] + p.ntile[1];

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmK0);
    tma_prefetch_desc(&tmV0);
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], nq);
    }
    for (int q = 0; q < 2; ++q) {
      mbar_init(&s_full[q], 1);
      mbar_init(&p_full[q], 128);
      mbar_init(&s_free[q], 128);
      mbar_init(&o_done[q], 1);
    }
    mbar_init(q_full, 256);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (whole warp, one elected lane issues) =====================
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < ntiles; ++j) {
        const int seg = j < p.ntile[0] ? 0 : 1;
        const int k0 = (seg == 0 ? j : j - p.ntile[0]) * kTcTile;
        const int kvf = (f / p.kv_div[seg]) * p.kv_mul[seg] + p.kv_add[seg];
        mbar_wait(&kv_empty[stage], phase ^ 1u);
        uint8_t* sK = sKV + stage * 2 * kTcTileBytes;
        mbar_arrive_expect_tx_warp(&kv_full[stage], 2u * kTcTileBytes);
        tma_load_3d_warp(sK, seg == 0 ? &tmK0 : &tmK1, &kv_full[stage], head * d, k0, kvf);
        tma_load_3d_warp(sK + kTcTileBytes, seg == 0 ? &tmV0 : &tmV1, &kv_full[stage], head * d, k0, kvf);
        if (++stage == kTcStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1 || warp == kTcMmaWarp1) {
    // ===================== MMA issuers: warp 1 serves query tile 0, warp 10 query tile 1 =====================
    // (whole converged warp; one elected lane issues, see umma_*_warp).  One issuer per query tile: with a single warp
    // serving both tiles in a fixed order, a tile's P.V had to wait behind the other tile's barriers (clock trace).
    const int q = warp == 1 ? 0 : 1;
    if (q < nq) {
      const uint32_t idesc_s = umma_idesc_f16(128, 128);
      const uint32_t idesc_o = umma_idesc_f16(128, NO) | (1u << 16);       // B (= V) is MN-major
      const uint32_t tS = tmem_base + (q ? kTcColS1 : kTcColS0);
      const uint32_t tO = tmem_base + (q ? (kPx ? kTcColO1Small : kTcColO1) : kTcColO0);
      const uint32_t tP = tmem_base + (q ? kTcColP1 : kTcColP0);
      const uint32_t tPx = tmem_base + (q ? kTcColX1 : kTcColX0);
      const uint64_t dQ = umma_desc_k_sw128(smem_u32(sQ + q * kTcTileBytes));
      mbar_wait(q_full, 0);
      fence_proxy_async_smem();
      tcgen05_fence_after();
      mbar_wait(&kv_full[0], 0);
      tcgen05_fence_after();
      {
        const uint64_t dK = umma_desc_k_sw128(smem_u32(sKV));
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) umma_f16_ss_warp(tS, dQ + 2u * ks, dK + 2u * ks, idesc_s, ks ? 1u : 0u);
        umma_commit_warp(&s_full[q]);
      }
      int stage = 0;
      uint32_t phase = 0;
      long long* trm = (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0) ? p.trace + 16 + 8 * q : nullptr;
      for (int j = 0; j < ntiles; ++j) {
        if (trm && j < 32) trm[32 * j + 0] = clock64();
        int nstage = stage + 1;
        uint32_t nphase = phase;
        if (nstage == kTcStages) {
          nstage = 0;
          nphase ^= 1u;
        }
        if (j + 1 < ntiles) {                            // next scores: S_q is free once it sits in registers
          mbar_wait(&kv_full[nstage], nphase);
          mbar_wait(&s_free[q], static_cast<uint32_t>(j & 1));
          tcgen05_fence_after();
          const uint64_t dKn = umma_desc_k_sw128(smem_u32(sKV + nstage * 2 * kTcTileBytes));
#pragma unroll
          for (int ks = 0; ks < KSTEPS; ++ks) umma_f16_ss_warp(tS, dQ + 2u * ks, dKn + 2u * ks, idesc_s, ks ? 1u : 0u);
          umma_commit_warp(&s_full[q]);
          if (trm && j < 32) trm[32 * j + 1] = clock64();            // QK(j+1) issued
        }
        const uint64_t dV = umma_desc_mn_sw128(smem_u32(sKV + stage * 2 * kTcTileBytes + kTcTileBytes), kTcTileBytes);
        mbar_wait(&p_full[q], static_cast<uint32_t>(j & 1));
        tcgen05_fence_after();
        if (trm && j < 32) trm[32 * j + 2] = clock64();              // P(j) seen
#pragma unroll
        for (int kk = 0; kk < kTcTile / 16; ++kk)       // 16 keys per step: P columns +8, V rows +16 (2 KiB)
          umma_f16_ts_warp(tO, ((kPx && kk < 2 && (j & 1)) ? tPx : tP) + 8u * kk, dV + 128u * kk, idesc_o, (j | kk) ? 1u : 0u);
        umma_commit_warp(&o_done[q]);
        umma_commit_warp(&kv_empty[stage]);                // this tile's K_j, V_j reads are done (count = nq)
        if (trm && j < 32) trm[32 * j + 3] = clock64();              // P.V issued
        stage = nstage;
        phase = nphase;
      }
    }
  } else {
    // ===================== softmax warpgroups =====================
    const int qt = (warp - 2) >> 2;                      // query tile of this warpgroup
    const int wq = warp & 3;                             // TMEM lane quarter of this warp
    const int row = wq * 32 + lane;                      // row inside the tile == TMEM lane
    const int t = (warp - 2) * 32 + lane;                // 0..255: row inside the CTA, for the Q load
    // ---- Q tile load: thread t copies query row t (zero-filled beyond lq and in the padded channels) ----
    {
      const int chunks = d >> 3;
      const int qrow = q0 + t;
      const bool ok = qrow < p.lq;
      const __half* src = p.q + static_cast<long long>(f) * p.q_fs + static_cast<long long>(ok ? qrow : 0) * p.ldq +
                          static_cast<long long>(head) * d;
      uint8_t* dst = sQ + (t >> 7) * kTcTileBytes + (t & 127) * 128;
      const int sw = t & 7;
      for (int c = 0; c < chunks; ++c) cp_async_16(smem_u32(dst + ((c ^ sw) << 4)), src + c * 8, ok);
      for (int c = chunks; c < 2 * KSTEPS; ++c) *reinterpret_cast<uint4*>(dst + ((c ^ sw) << 4)) = make_uint4(0, 0, 0, 0);
      cp_async_commit();
      cp_async_wait<0>();
      fence_proxy_async_smem();
      mbar_arrive(q_full);
    }
    if (qt < nq) {
      const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
      const uint32_t tS = tmem_base + lane_off + (qt ? kTcColS1 : kTcColS0);
      const uint32_t tO = tmem_base + lane_off + (qt ? (kPx ? kTcColO1Small : kTcColO1) : kTcColO0);
      const uint32_t tP = tmem_base + lane_off + (qt ? kTcColP1 : kTcColP0);
      const uint32_t tPx = tmem_base + lane_off + (qt ? kTcColX1 : kTcColX0);
      const float c = p.scale_log2;
      float mref = -INFINITY, l = 0.f;
      // trace rows: tile j -> [32*j + 8*qt + k]: 0 tile start, 1 S ready, 2 S in registers, 3 max done, 4 P.V of the previous tile done, 5 P written, 6 p_full arrived; [32*j + 16 + ...]: MMA warp
      long long* tr = (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && wq == 0 && lane == 0)
                          ? p.trace + 8 * qt : nullptr;
      // Stagger (p.stagger): left alone the two warpgroups run in lockstep - clock trace: their S load / row max / P.V
      // wait phases coincide and the MUFU pipe idles a third of every tile.  Query tile 1 therefore starts its first
      // key tile only after tile 0 is 1 = past its row maximum, 2 = half way through its exponentials.  (Strict
      // alternation of the exponential phases was measured slower: one warp per sub-partition cannot saturate MUFU.)
      const int stagger = nq == 2 ? p.stagger : 0;
      if (stagger && qt == 1) named_bar_sync(1, 256);
      for (int j = 0; j < ntiles; ++j) {
        if (tr && j < 32) tr[32 * j + 0] = clock64();
        const int seg = j < p.ntile[0] ? 0 : 1;
        const int valid = p.lkv[seg] - (seg == 0 ? j : j - p.ntile[0]) * kTcTile;   // keys of this tile that exist
        mbar_wait(&s_full[qt], static_cast<uint32_t>(j & 1));
        tcgen05_fence_after();
        if (tr && j < 32) tr[32 * j + 1] = clock64();
        uint32_t r[128];
        tmem_ld_32x32b_x32(tS, r);
        tmem_ld_32x32b_x32(tS + 32, r + 32);
        tmem_ld_32x32b_x32(tS + 64, r + 64);
        tmem_ld_32x32b_x32(tS + 96, r + 96);
        tmem_ld_wait();
        tcgen05_fence_before();
        mbar_arrive(&s_free[qt]);                            // the MMA warp may overwrite S_q with the next scores
        if (tr && j < 32) tr[32 * j + 2] = clock64();
        if (valid < kTcTile) {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (i >= valid) r[i] = 0xff800000u;          // -inf
        }
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 128; i += 8) {
          m0 = max3_f(m0, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
          m1 = max3_f(m1, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
          m2 = max3_f(m2, __uint_as_float(r[i + 4]), __uint_as_float(r[i + 5]));
          m3 = max3_f(m3, __uint_as_float(r[i + 6]), __uint_as_float(r[i + 7]));
        }
        const float mt = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * c;
        if (tr && j < 32) tr[32 * j + 3] = clock64();
        // lazy rescaling: keep the reference maximum unless the new one exceeds it by more than 2^8
        const bool need = mt > mref + 8.f;
        float alpha = 1.f;
        if (need) {
          alpha = ex2_approx(mref - mt);                 // 0 on the first tile (mref = -inf)
          mref = mt;
        }
        bool owait = j > 0;                              // P_q / O_q still belong to P.V of the previous tile
        if (j > 0 && __any_sync(0xffffffffu, need)) {     // rare after the first tiles: 16 columns at a time
          mbar_wait(&o_done[qt], static_cast<uint32_t>((j - 1) & 1));
          tcgen05_fence_after();
          owait = false;
#pragma unroll 1
          for (int cc = 0; cc < NO; cc += 16) {
            uint32_t o[16];
            tmem_ld_x16(tO + cc, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_x16(tO + cc, o);
          }
        }
        if (stagger == 1 && qt == 0 && j == 0) named_bar_arrive(1, 256);
        l *= alpha;
        const float nm = -mref;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int cc = 0; cc < 128; cc += 32) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float x0 = fmaf(__uint_as_float(r[cc + i]), c, nm), x1 = fmaf(__uint_as_float(r[cc + i + 1]), c, nm);
            const float x2 = fmaf(__uint_as_float(r[cc + i + 2]), c, nm), x3 = fmaf(__uint_as_float(r[cc + i + 3]), c, nm);
            const float p0 = ex2_approx(x0);
            const float p1 = ex2_approx(x1);
            const float p2 = EMU >= 2 ? ex2_poly(x2) : ex2_approx(x2);
            const float p3 = EMU >= 1 ? ex2_poly(x3) : ex2_approx(x3);
            s0 += p0;
            s1 += p1;
            s2 += p2;
            s3 += p3;
            pk[i >> 1] = pack_h2(p0, p1);
            pk[(i >> 1) + 1] = pack_h2(p2, p3);
          }
          if (cc == (kPx ? 32 : 0) && owait) {
            mbar_wait(&o_done[qt], static_cast<uint32_t>((j - 1) & 1));
            tcgen05_fence_after();
          }
          if (cc == (kPx ? 32 : 0) && tr && j < 32) tr[32 * j + 4] = clock64();
          tmem_st_x16(((kPx && cc == 0 && (j & 1)) ? tPx : tP) + (cc >> 1), pk);   // P as packed fp16 pairs
          if (cc == 32 && stagger == 2 && qt == 0 && j == 0) named_bar_arrive(1, 256);
        }
        l += (s0 + s1) + (s2 + s3);
        tmem_st_wait();
        if (tr && j < 32) tr[32 * j + 5] = clock64();
        tcgen05_fence_before();
        mbar_arrive(&p_full[qt]);
        if (tr && j < 32) tr[32 * j + 6] = clock64();
      }
      // ---- epilogue: O / l -> fp16 -> global ----
      mbar_wait(&o_done[qt], static_cast<uint32_t>((ntiles - 1) & 1));
      tcgen05_fence_after();
      uint32_t o[NO];
#pragma unroll
      for (int cc = 0; cc < NO; cc += 16) tmem_ld_x16(tO + cc, o + cc);
      tmem_ld_wait();
      const int qrow = q0 + qt * kTcTile + row;
      if (qrow < p.lq) {
        const float inv = 1.f / l;
        __half* dst = p.o + static_cast<long long>(f) * p.o_fs + static_cast<long long>(qrow) * p.ldo +
                      static_cast<long long>(head) * d;
#pragma unroll
        for (int cc = 0; cc < NO; cc += 8) {
          if (cc < d) {
            uint4 u;
            u.x = pack_h2(__uint_as_float(o[cc]) * inv, __uint_as_float(o[cc + 1]) * inv);
            u.y = pack_h2(__uint_as_float(o[cc + 2]) * inv, __uint_as_float(o[cc + 3]) * inv);
            u.z = pack_h2(__uint_as_float(o[cc + 4]) * inv, __uint_as_float(o[cc + 5]) * inv);
            u.w = pack_h2(__uint_as_float(o[cc + 6]) * inv, __uint_as_float(o[cc + 7]) * inv);
            *reinterpret_cast<uint4*>(dst + cc) = u;
          }
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiledA)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiledA attn_encode_fn() {
  static PFN_encodeTiledA fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiledA>(ptr);
  });
  return fn;
}

// [frames][rows][cols] fp16 view: cols contiguous, row stride ld, frame stride fs (elements); box = 64 cols x 128 rows.
static bool make_kv_map(CUtensorMap* m, const void* base, int cols, int rows, long long ld, long long fs, int frames) {
  PFN_encodeTiledA enc = attn_encode_fn();
  if (!enc) return false;
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)frames};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(frames > 1 ? fs : ld * rows) * 2};
  cuuint32_t box[3] = {64, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int KSTEPS, int EMU>
static int launch_tc(const CUtensorMap* maps, const FaTcParams& p, int frames, int heads, cudaStream_t st) {
  const int smem = 1024 + 2 * kTcTileBytes + kTcStages * 2 * kTcTileBytes + 256;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] {
    attr_err = cudaFuncSetAttribute(flash_attn_tc_kernel<KSTEPS, EMU>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  });
  if (attr_err != cudaSuccess) {
    set_last_error("ccedit_attention(tc): cudaFuncSetAttribute: %s", cudaGetErrorString(attr_err));
    return CCEDIT_ERR_CUDA;
  }
  dim3 grid((p.lq + 2 * kTcTile - 1) / (2 * kTcTile), heads, frames);
  flash_attn_tc_kernel<KSTEPS, EMU><<<grid, kTcThreads, smem, st>>>(maps[0], maps[1], maps[2], maps[3], p);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  CCEDIT_CUDA_LAUNCH_CHECK("ccedit_attention(tc)");
  return CCEDIT_OK;
}

// Returns -1 when the problem is not eligible for the tcgen05 kernel (caller falls back to the mma.sync kernel),
// otherwise a CCEDIT_* status.
int attention_tc(const ccedit_attn_desc* a, cudaStream_t st) {
  if (a->d > 64 || a->d % 8 != 0) return -1;
  FaTcParams p;
  memset(&p, 0, sizeof(p));
  CUtensorMap maps[4];
  for (int s = 0; s < a->nseg; ++s) {
    if ((reinterpret_cast<uintptr_t>(a->k[s]) & 15) || (reinterpret_cast<uintptr_t>(a->v[s]) & 15)) return -1;
    // number of kv frames addressed: the largest frame index any query frame maps to, + 1
    const int last = ((a->frames - 1) / a->kv_div[s]) * a->kv_mul[s] + a->kv_add[s];
    const int cols = a->heads * a->d;
    if (!make_kv_map(&maps[2 * s], a->k[s], cols, a->lkv[s], a->ldk[s], a->kv_frame_stride[s], last + 1) ||
        !make_kv_map(&maps[2 * s + 1], a->v[s], cols, a->lkv[s], a->ldv[s], a->kv_frame_stride[s], last + 1)) {
      set_last_error("ccedit_attention(tc): cuTensorMapEncodeTiled failed for segment %d", s);
      return CCEDIT_ERR_CUDA;
    }
    p.lkv[s] = a->lkv[s];
    p.kv_div[s] = a->kv_div[s];
    p.kv_mul[s] = a->kv_mul[s];
    p.kv_add[s] = a->kv_add[s];
    p.ntile[s] = (a->lkv[s] + kTcTile - 1) / kTcTile;
  }
  if (a->nseg == 1) {
    maps[2] = maps[0];
    maps[3] = maps[1];
    p.kv_div[1] = 1;
  }
  p.q = static_cast<const __half*>(a->q);
  p.ldq = a->ldq;
  p.q_fs = a->q_frame_stride;
  p.o = static_cast<__half*>(a->o);
  p.ldo = a->ldo;
  p.o_fs = a->o_frame_stride;
  p.nseg = a->nseg;
  p.lq = a->lq;
  p.d = a->d;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.trace = g_trace_buf;
  static const int stagger = [] {                  // developer switch; default = the measured best
    const char* e = getenv("CCEDIT_ATTN_STAGGER");
    return e ? atoi(e) : kTcDefaultStagger;
  }();
  p.stagger = stagger;
  const int ks = (a->d + 15) / 16;
  static const int emu = [] {                      // developer switch; default = the measured best
    const char* e = getenv("CCEDIT_ATTN_EMU");
    return e ? atoi(e) : kTcDefaultEmu;
  }();
  switch (ks * 4 + (emu < 0 ? 0 : emu > 2 ? 2 : emu)) {
    case 4: return launch_tc<1, 0>(maps, p, a->frames, a->heads, st);
    case 5: return launch_tc<1, 1>(maps, p, a->frames, a->heads, st);
    case 6: return launch_tc<1, 2>(maps, p, a->frames, a->heads, st);
    case 8: return launch_tc<2, 0>(maps, p, a->frames, a->heads, st);
    case 9: return launch_tc<2, 1>(maps, p, a->frames, a->heads, st);
    case 10: return launch_tc<2, 2>(maps, p, a->frames, a->heads, st);
    case 12: return launch_tc<3, 0>(maps, p, a->frames, a->heads, st);
    case 13: return launch_tc<3, 1>(maps, p, a->frames, a->heads, st);
    case 14: return launch_tc<3, 2>(maps, p, a->frames, a->heads, st);
    case 16: return launch_tc<4, 0>(maps, p, a->frames, a->heads, st);
    case 17: return launch_tc<4, 1>(maps, p, a->frames, a->heads, st);
    default: return launch_tc<4, 2>(maps, p, a->frames, a->heads, st);
  }
}

}  // namespace ccedit

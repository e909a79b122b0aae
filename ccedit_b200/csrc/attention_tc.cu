// Flash attention on tcgen05 tensor cores with TMEM-resident S / P / O (sm_100a), head dims 8..160 (the SD-1.5 widths give
// d = 40 / 80 / 160; d = 40 at 6144 tokens is 88 % of the attention FLOPs of the CCEdit UNet: 7 launches of
// 34 frames x 8 heads x 6144^2 per network call).
//
// One CTA = one 128-query tile of one (frame, head); two CTAs per SM (256 TMEM columns, <= 97 KB shared memory each):
//   warp 0      TMA producer: K and V tiles [KT keys x 64-channel blocks] through 3-D tensor maps (SWIZZLE_128B; rows past
//               the end of the segment are zero-filled by TMA), 2-stage mbarrier ring
//   warp 1      TMEM allocator + tcgen05.mma issue (the whole converged warp runs the issue code with warp-uniform
//               operands and one elected lane issues: from inside `if (lane == 0)` every tcgen05 instruction costs an
//               ELECT / R2UR.BROADCAST / branch loop of ~100 clocks):
//                 S  = Q . K_j^T      (A, B from shared memory, K-major, fp32 accumulator in TMEM)
//                 O += P . V_j        (A = P from TMEM, B = V from shared memory, MN-major)
//   warps 2-5   softmax of the first half of the tile's keys   } thread = one query row = one TMEM lane: tcgen05.ld S ->
//   warps 6-9   softmax of the second half of the tile's keys  } row max (exchanged between the halves through shared
//               memory) with lazy rescaling -> exp2 -> fp16 P written to TMEM with tcgen05.st -> row sums; O / l epilogue
// The kernel is bound by the MUFU exp2 rate (160 tensor FLOPs per exponential at d = 40; ncu: XU pipe 70 %), so everything
// is arranged to keep the softmax warps issuing exponentials - see the notes at the kernel and profiles/r01_attention.md
// for the variants that were measured (two query tiles per CTA, strict ping-pong, staggering, one MMA warp per tile):
//   * S(j+1) = Q K_{j+1}^T is issued as soon as the softmax warps have pulled S(j) into registers (s_free), not after
//     P(j): the MMA round trip is off the softmax -> softmax critical path;
//   * the wait for P(j-1).V_{j-1} (o_done: P and O are free again) sits right before the first P store that needs it,
//     and at d <= 48 the first 32 keys of P are double-buffered so that even that store does not wait;
//   * row maxima with 3-input max; one of every 4 exponentials evaluated on the FMA pipe (Cody-Waite + degree-3
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

constexpr int kTcTile = 128;                 // query rows per tile
constexpr int kTcDefaultEmu = 1;             // exponentials per group of 4 evaluated on the FMA pipe

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
  int hpc;            // heads per CTA (> 1 for short key sequences: amortises the CTA set-up, overlaps the next Q load)
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

// ---------------------------------------------------------------------------------------------------------------
// Why split rows: ONE 128-query tile per CTA, TWO CTAs per SM, and every score row split between two softmax warpgroups
// (first / second half of the tile's keys).  An earlier version ran two 128-query tiles per CTA with one warpgroup each
// (2 softmax warps per sub-partition); this one has 4: measured (clock trace + MUFU
// microbenchmark, tools/bench_mufu.cu) one warp per sub-partition sustains one MUFU.EX2 per ~10 clocks, two reach 8.4
// (the pipe's 8), and with only two warps every S-wait / TMEM-load / row-max phase (~600 clocks per tile) is MUFU idle
// time.  The two halves of a row exchange their partial maxima through shared memory (one named barrier per key tile)
// so that they scale P identically; the row sums are merged once, in the epilogue.
//   TMEM: S [0,128) fp32 scores, P [128,192) fp16 pairs, O [192,192+NO), X [240,256) second buffer of P's first 32 keys
//   warps: 0 TMA producer, 1 TMEM allocator + MMA issuer, 2-5 softmax half 0, 6-9 softmax half 1 (lane quarter = warp & 3)
// ---------------------------------------------------------------------------------------------------------------
constexpr int kT2Threads = 320;
constexpr int kT2Stages = 2;

// Head dims above 64 take NB = ceil(d / 64) column blocks of 64 channels per operand tile ([NB][rows][128 B], each
// SWIZZLE_128B) and 64-key tiles, so that S (64) + P (32) + O (16 * KSTEPS <= 160) still fit the CTA's 256 TMEM columns.
template <int KSTEPS>
struct T2Cfg {
  static constexpr int NB = (KSTEPS + 3) / 4;
  static constexpr int KT = NB == 1 ? 128 : 64;          // keys per tile
  static constexpr int NO = 16 * KSTEPS;                 // output columns
  static constexpr uint32_t ColS = 0, ColP = KT, ColO = KT + KT / 2, ColX = 240;
  static constexpr bool kPx = NB == 1 && NO <= 48;       // second buffer for the first 32 keys of P
  static constexpr int QBytes = NB * kTcTile * 128;
  static constexpr int KBytes = NB * KT * 128;           // one K (or V) tile
  static constexpr int Smem = 1024 + QBytes + kT2Stages * 2 * KBytes + 128 + 768 * 4;
  static_assert(ColO + NO <= 256, "TMEM budget");
};

template <int KSTEPS, int EMU, bool MH>   // MH: several heads per CTA (p.hpc), for short key sequences
__global__ void __launch_bounds__(kT2Threads, 2)
flash_attn_tc2_kernel(const __grid_constant__ CUtensorMap tmK0, const __grid_constant__ CUtensorMap tmV0,
                      const __grid_constant__ CUtensorMap tmK1, const __grid_constant__ CUtensorMap tmV1,
                      const __grid_constant__ FaTcParams p) {
  using Cfg = T2Cfg<KSTEPS>;
  constexpr int NB = Cfg::NB, KT = Cfg::KT, NO = Cfg::NO, HK = KT / 2;   // HK: keys per softmax half
  constexpr bool kPx = Cfg::kPx;
  extern __shared__ uint8_t fa_smem_raw[];
  const uint32_t raw_addr = smem_u32(fa_smem_raw);
  uint8_t* smem = fa_smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sQ = smem;                                    // [NB][128][128 B]
  uint8_t* sKV = smem + Cfg::QBytes;                     // stages x (K [NB][KT][128 B], V [NB][KT][128 B])
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + kT2Stages * 2 * Cfg::KBytes);
  // K and V tiles have their own full / empty barriers: K(j) is released by Q K_j^T, a whole softmax earlier than P V_j
  // releases V(j), so that the next-but-one K tile is in flight a tile earlier.  With one barrier per stage the K load was
  // requested only ~0.1 tile periods before it was needed - hidden at d = 40 (128-key tiles, long softmax), but at
  // d = 80 / 160 (64-key tiles) the softmax warps waited ~a TMA latency per tile for S (ncu: long_scoreboard 3.3).
  uint64_t* k_full = bars;                               // [stages]
  uint64_t* v_full = bars + kT2Stages;                   // [stages]
  uint64_t* k_empty = bars + 2 * kT2Stages;              // [stages]
  uint64_t* v_empty = bars + 3 * kT2Stages;              // [stages]
  uint64_t* s_full = bars + 4 * kT2Stages;               // S ready
  uint64_t* p_full = s_full + 1;                         // P written (256 arrivals)
  uint64_t* q_full = p_full + 1;                         // Q tile in shared memory (128 arrivals)
  uint64_t* s_free = q_full + 1;                         // S is in registers (256 arrivals)
  uint64_t* o_done = s_free + 1;                         // P.V done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 1);
  float* smax = reinterpret_cast<float*>(bars + 16);     // [2 parities][2 halves][128 rows]
  float* lsum = smax + 512;                              // [128 rows]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hpc = MH ? p.hpc : 1;
  const int q0 = blockIdx.x * kTcTile, head0 = blockIdx.y * hpc, f = blockIdx.z;
  const int d = p.d;
  const int ntiles = p.ntile[0] + p.ntile[1];     // key tiles per head
  const int total = ntiles * hpc;                 // (head, key tile) items of this CTA, heads outermost

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmK0);
    tma_prefetch_desc(&tmV0);
    for (int s = 0; s < kT2Stages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 256);
    mbar_init(q_full, 128);
    mbar_init(s_free, 256);
    mbar_init(o_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();                  // PDL: barriers / TMEM above overlap the previous kernel's tail; its outputs are read below
  griddep_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer: two cursors, K runs ahead of V =====================
    int kst = 0, vst = 0, kit = 0, vit = 0;
    uint32_t kph = 0, vph = 0;
    auto issue = [&](int it, int stage, bool is_v) {
      const int hl = MH ? it / ntiles : 0, j = it - hl * ntiles;
      const int head = head0 + hl;
      const int seg = j < p.ntile[0] ? 0 : 1;
      const int k0 = (seg == 0 ? j : j - p.ntile[0]) * KT;
      const int kvf = (f / p.kv_div[seg]) * p.kv_mul[seg] + p.kv_add[seg];
      uint8_t* dst = sKV + stage * 2 * Cfg::KBytes + (is_v ? Cfg::KBytes : 0);
      uint64_t* bar = is_v ? &v_full[stage] : &k_full[stage];
      const CUtensorMap* tm = is_v ? (seg == 0 ? &tmV0 : &tmV1) : (seg == 0 ? &tmK0 : &tmK1);
      mbar_arrive_expect_tx_warp(bar, static_cast<uint32_t>(Cfg::KBytes));
#pragma unroll
      for (int b = 0; b < NB; ++b) tma_load_3d_warp(dst + b * KT * 128, tm, bar, head * d + 64 * b, k0, kvf);
    };
    while (kit < total || vit < total) {
      bool moved = false;
      if (kit < total && mbar_test_wait(&k_empty[kst], kph ^ 1u)) {
        issue(kit, kst, false);
        ++kit;
        if (++kst == kT2Stages) { kst = 0; kph ^= 1u; }
        moved = true;
      }
      if (vit < total && mbar_test_wait(&v_empty[vst], vph ^ 1u)) {
        issue(vit, vst, true);
        ++vit;
        if (++vst == kT2Stages) { vst = 0; vph ^= 1u; }
        moved = true;
      }
      if (!moved) __nanosleep(32);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc_s = umma_idesc_f16(128, KT);
    const uint32_t idesc_o = umma_idesc_f16(128, NO) | (1u << 16);       // B (= V) is MN-major
    const uint32_t tS = tmem_base + Cfg::ColS, tO = tmem_base + Cfg::ColO, tP = tmem_base + Cfg::ColP, tPx = tmem_base + Cfg::ColX;
    const uint32_t sQa = smem_u32(sQ);
    mbar_wait(q_full, 0);
    fence_proxy_async_smem();
    tcgen05_fence_after();
    mbar_wait(&k_full[0], 0);
    tcgen05_fence_after();
    {
      const uint32_t sKa = smem_u32(sKV);
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks)                // channel block ks / 4, 32 bytes per k-step inside the 128 B row
        umma_f16_ss_warp(tS, umma_desc_k_sw128(sQa + (ks >> 2) * (kTcTile * 128)) + 2u * (ks & 3),
                         umma_desc_k_sw128(sKa + (ks >> 2) * (KT * 128)) + 2u * (ks & 3), idesc_s, ks ? 1u : 0u);
      umma_commit_warp(s_full);
      umma_commit_warp(&k_empty[0]);                      // K(0) is free as soon as these MMAs have read it
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < total; ++it) {
      const int hl = MH ? it / ntiles : 0, j = it - hl * ntiles;
      const bool last_tile = j + 1 == ntiles;
      int nstage = stage + 1;
      uint32_t nphase = phase;
      if (nstage == kT2Stages) {
        nstage = 0;
        nphase ^= 1u;
      }
      auto next_scores = [&]() {                         // S(it+1) = Q K^T: S is free once it sits in registers
        if (it + 1 >= total) return;
        mbar_wait(&k_full[nstage], nphase);
        mbar_wait(s_free, static_cast<uint32_t>(it & 1));
        if (MH && last_tile) {                           // first key tile of the next head: its Q tile must have landed
          mbar_wait(q_full, static_cast<uint32_t>((hl + 1) & 1));
          fence_proxy_async_smem();
        }
        tcgen05_fence_after();
        const uint32_t sKa = smem_u32(sKV + nstage * 2 * Cfg::KBytes);
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks)
          umma_f16_ss_warp(tS, umma_desc_k_sw128(sQa + (ks >> 2) * (kTcTile * 128)) + 2u * (ks & 3),
                           umma_desc_k_sw128(sKa + (ks >> 2) * (KT * 128)) + 2u * (ks & 3), idesc_s, ks ? 1u : 0u);
        umma_commit_warp(s_full);
        umma_commit_warp(&k_empty[nstage]);
      };
      // within a head the next scores go first (they are ready long before P); across heads P.V goes first, because the
      // next head's Q tile is published only after this tile's softmax
      if (!(MH && last_tile)) next_scores();
      // V tile: MN-major, 64-channel atoms KT * 128 bytes apart (LBO), 16 keys (2 KiB) per k-step
      const uint64_t dV = umma_desc_mn_sw128(smem_u32(sKV + stage * 2 * Cfg::KBytes + Cfg::KBytes), KT * 128);
      mbar_wait(&v_full[stage], phase);
      mbar_wait(p_full, static_cast<uint32_t>(it & 1));
      tcgen05_fence_after();
#pragma unroll
      for (int kk = 0; kk < KT / 16; ++kk)
        umma_f16_ts_warp(tO, ((kPx && kk < 2 && (it & 1)) ? tPx : tP) + 8u * kk, dV + 128u * kk, idesc_o, (j | kk) ? 1u : 0u);
      umma_commit_warp(o_done);
      if (MH && last_tile) next_scores();
      umma_commit_warp(&v_empty[stage]);
      stage = nstage;
      phase = nphase;
    }
  } else {
    // ===================== softmax: warps 2-5 first half of the tile's keys, warps 6-9 second half =====================
    const int half = (warp - 2) >> 2;
    const int wq = warp & 3;
    const int row = wq * 32 + lane;
    // ---- Q tile load (half 0): thread copies query row `row` of head h (zero-filled beyond lq and in the padded channels).
    // The first head's tile is waited for right away; the next head's tile is requested while the last key tile of the
    // current head is still being processed (its QK^T has completed by then, so the buffer is free) and waited for at
    // the end of the head. ----
    auto q_issue = [&](int head) {
      const int chunks = d >> 3;
      const int qrow = q0 + row;
      const bool ok = qrow < p.lq;
      const __half* src = p.q + static_cast<long long>(f) * p.q_fs + static_cast<long long>(ok ? qrow : 0) * p.ldq +
                          static_cast<long long>(head) * d;
      uint8_t* dst = sQ + row * 128;
      const int sw = row & 7;
      for (int c = 0; c < chunks; ++c)
        cp_async_16(smem_u32(dst + (c >> 3) * (kTcTile * 128) + (((c & 7) ^ sw) << 4)), src + c * 8, ok);
      for (int c = chunks; c < 2 * KSTEPS; ++c)
        *reinterpret_cast<uint4*>(dst + (c >> 3) * (kTcTile * 128) + (((c & 7) ^ sw) << 4)) = make_uint4(0, 0, 0, 0);
      cp_async_commit();
    };
    auto q_publish = [&]() {
      cp_async_wait<0>();
      fence_proxy_async_smem();
      mbar_arrive(q_full);
    };
    if (half == 0) {
      q_issue(head0);
      q_publish();
    }
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t tS = tmem_base + lane_off + Cfg::ColS + static_cast<uint32_t>(HK * half);
    const uint32_t tO = tmem_base + lane_off + Cfg::ColO;
    const uint32_t tP = tmem_base + lane_off + Cfg::ColP + static_cast<uint32_t>((HK / 2) * half);
    const uint32_t tPx = tmem_base + lane_off + Cfg::ColX;
    const float c = p.scale_log2;
    float mref = -INFINITY, l = 0.f;
    // ---- end of a head: merge the row sums; O / l -> fp16 -> global, the 16-column chunks split between the halves ----
    auto head_epilogue = [&](int it, int head, float l) {
      lsum[half * 128 + row] = l;
      named_bar_sync(1, 256);
      const float lt = l + lsum[(half ^ 1) * 128 + row];
      mbar_wait(o_done, static_cast<uint32_t>(it & 1));
      tcgen05_fence_after();
      const int qrow = q0 + row;
      const float inv = 1.f / lt;
      __half* dst = p.o + static_cast<long long>(f) * p.o_fs + static_cast<long long>(qrow < p.lq ? qrow : 0) * p.ldo +
                    static_cast<long long>(head) * d;
#pragma unroll 1
      for (int cc = 16 * half; cc < NO; cc += 32) {
        uint32_t o[16];
        tmem_ld_x16(tO + cc, o);
        tmem_ld_wait();
        if (qrow < p.lq) {
#pragma unroll
          for (int h8 = 0; h8 < 16; h8 += 8) {
            if (cc + h8 < d) {
              uint4 u;
              u.x = pack_h2(__uint_as_float(o[h8]) * inv, __uint_as_float(o[h8 + 1]) * inv);
              u.y = pack_h2(__uint_as_float(o[h8 + 2]) * inv, __uint_as_float(o[h8 + 3]) * inv);
              u.z = pack_h2(__uint_as_float(o[h8 + 4]) * inv, __uint_as_float(o[h8 + 5]) * inv);
              u.w = pack_h2(__uint_as_float(o[h8 + 6]) * inv, __uint_as_float(o[h8 + 7]) * inv);
              *reinterpret_cast<uint4*>(dst + cc + h8) = u;
            }
          }
        }
      }
    };
    // An mbarrier wait costs ~100-200 clocks even when its phase completed long ago (TRYWAIT latency).  S(it + 1) and
    // P.V(it - 1) are therefore probed with non-blocking tests issued under other work (the last P stores of a tile /
    // the row maximum) and the blocking wait is only the fallback.
    bool s_ready = false;
    for (int it = 0; it < total; ++it) {
      const int hl = MH ? it / ntiles : 0, j = it - hl * ntiles;   // head (local), key tile
      const int head = head0 + hl;
      const bool last_tile = j + 1 == ntiles;
      if (MH && j == 0) {
        mref = -INFINITY;
        l = 0.f;
      }
      const int seg = j < p.ntile[0] ? 0 : 1;
      const int valid = p.lkv[seg] - (seg == 0 ? j : j - p.ntile[0]) * KT - HK * half;   // my keys that exist
      if (!s_ready) mbar_wait(s_full, static_cast<uint32_t>(it & 1));
      tcgen05_fence_after();
      uint32_t r[HK];
#pragma unroll
      for (int cc = 0; cc < HK; cc += 32) tmem_ld_32x32b_x32(tS + cc, r + cc);
      tmem_ld_wait();
      tcgen05_fence_before();
      mbar_arrive(s_free);
      const bool o_ready = it > 0 && mbar_test_wait(o_done, static_cast<uint32_t>((it - 1) & 1));
      if (MH && half == 0 && last_tile && hl + 1 < hpc) q_issue(head + 1);   // Q of the next head (this head's QK^T are done)
      if (valid < HK) {
#pragma unroll
        for (int i = 0; i < HK; ++i)
          if (i >= valid) r[i] = 0xff800000u;            // -inf
      }
      float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < HK; i += 8) {
        m0 = max3_f(m0, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
        m1 = max3_f(m1, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
        m2 = max3_f(m2, __uint_as_float(r[i + 4]), __uint_as_float(r[i + 5]));
        m3 = max3_f(m3, __uint_as_float(r[i + 6]), __uint_as_float(r[i + 7]));
      }
      const float mloc = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      float* sm = smax + (it & 1) * 256 + row;
      sm[half * 128] = mloc;
      named_bar_sync(1, 256);                             // both halves of every row have published their maximum
      const float mt = fmaxf(mloc, sm[(half ^ 1) * 128]) * c;
      const bool need = mt > mref + 8.f;                  // identical in both halves of the row
      float alpha = 1.f;
      if (need) {
        alpha = ex2_approx(mref - mt);
        mref = mt;
      }
      bool owait = it > 0;                               // P / O still belong to P.V of the previous tile
      if (j > 0 && __any_sync(0xffffffffu, need)) {       // rare after the first tiles; the 16-column chunks alternate
        if (!o_ready) mbar_wait(o_done, static_cast<uint32_t>((it - 1) & 1));
        tcgen05_fence_after();
        owait = false;
#pragma unroll 1
        for (int cc = 16 * half; cc < NO; cc += 32) {
          uint32_t o[16];
          tmem_ld_x16(tO + cc, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_x16(tO + cc, o);
        }
      }
      l *= alpha;
      const float nm = -mref;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int cc = 0; cc < HK; cc += 32) {
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
        const bool alt = kPx && half == 0 && cc == 0;     // the first 32 keys of P have a second buffer (odd tiles)
        if (owait && !alt) {
          if (!o_ready) mbar_wait(o_done, static_cast<uint32_t>((it - 1) & 1));
          tcgen05_fence_after();
          owait = false;
        }
        tmem_st_x16(((alt && (it & 1)) ? tPx : tP) + (cc >> 1), pk);
      }
      l += (s0 + s1) + (s2 + s3);
      s_ready = it + 1 < total && mbar_test_wait(s_full, static_cast<uint32_t>((it + 1) & 1));
      tmem_st_wait();
      tcgen05_fence_before();
      mbar_arrive(p_full);
      if (MH && last_tile) {
        if (half == 0 && hl + 1 < hpc) q_publish();        // before the epilogue: the MMA warp needs it for the next QK^T
        head_epilogue(it, head, l);
        tcgen05_fence_before();                            // O has been read before the next head's first P.V overwrites it
        named_bar_sync(1, 256);                            // lsum / smax are free for the next head
      }
    }
    if (!MH) head_epilogue(total - 1, head0, l);
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 256);
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

// [frames][rows][cols] fp16 view: cols contiguous, row stride ld, frame stride fs (elements); box = 64 cols x kt rows.
static bool make_kv_map(CUtensorMap* m, const void* base, int cols, int rows, long long ld, long long fs, int frames,
                        int kt) {
  PFN_encodeTiledA enc = attn_encode_fn();
  if (!enc) return false;
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)frames};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(frames > 1 ? fs : ld * rows) * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)kt, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int KSTEPS, int EMU, bool MH>
static int launch_tc2_t(const CUtensorMap* maps, const FaTcParams& p, int frames, int heads, cudaStream_t st) {
  const int smem = T2Cfg<KSTEPS>::Smem;
  static std::atomic<bool> attr_set[kMaxDevices];
  const int dev = current_device();
  CCEDIT_CHECK_ARG(dev >= 0, "ccedit_attention(tc2): no current CUDA device");
  if (!attr_set[dev].load(std::memory_order_acquire)) {
    const cudaError_t e = cudaFuncSetAttribute(flash_attn_tc2_kernel<KSTEPS, EMU, MH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_last_error("ccedit_attention(tc2): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return CCEDIT_ERR_CUDA;
    }
    attr_set[dev].store(true, std::memory_order_release);
  }
  dim3 grid((p.lq + kTcTile - 1) / kTcTile, heads / p.hpc, frames);
  (void)launch_pdl(2, flash_attn_tc2_kernel<KSTEPS, EMU, MH>, grid, dim3(kT2Threads), smem, st, maps[0], maps[1], maps[2], maps[3], p);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  CCEDIT_CUDA_LAUNCH_CHECK("ccedit_attention(tc2)");
  return CCEDIT_OK;
}

template <int KSTEPS, int EMU>
static int launch_tc2(const CUtensorMap* maps, const FaTcParams& p, int frames, int heads, cudaStream_t st) {
  return p.hpc > 1 ? launch_tc2_t<KSTEPS, EMU, true>(maps, p, frames, heads, st)
                   : launch_tc2_t<KSTEPS, EMU, false>(maps, p, frames, heads, st);
}

// Returns -1 when the problem is not eligible for the tcgen05 kernel (caller falls back to the mma.sync kernel),
// otherwise a CCEDIT_* status.
int attention_tc(const ccedit_attn_desc* a, cudaStream_t st) {
  if (a->d > 160 || a->d % 8 != 0) return -1;
  int ks = (a->d + 15) / 16;
  ks = ks <= 6 ? ks : ks <= 8 ? 8 : 10;                  // instantiated widths: 16..96, 128, 160 channels
  const int kt = ks <= 4 ? 128 : 64;                     // T2Cfg<ks>::KT
  FaTcParams p;
  memset(&p, 0, sizeof(p));
  CUtensorMap maps[4];
  for (int s = 0; s < a->nseg; ++s) {
    if ((reinterpret_cast<uintptr_t>(a->k[s]) & 15) || (reinterpret_cast<uintptr_t>(a->v[s]) & 15)) return -1;
    if ((a->ldk[s] & 7) || (a->ldv[s] & 7) || (a->kv_frame_stride[s] & 7)) return -1;   // TMA: 16-byte strides
    // number of kv frames addressed: the largest frame index any query frame maps to, + 1
    const int last = ((a->frames - 1) / a->kv_div[s]) * a->kv_mul[s] + a->kv_add[s];
    const int cols = a->heads * a->d;
    if (!make_kv_map(&maps[2 * s], a->k[s], cols, a->lkv[s], a->ldk[s], a->kv_frame_stride[s], last + 1, kt) ||
        !make_kv_map(&maps[2 * s + 1], a->v[s], cols, a->lkv[s], a->ldv[s], a->kv_frame_stride[s], last + 1, kt)) {
      set_last_error("ccedit_attention(tc): cuTensorMapEncodeTiled failed for segment %d", s);
      return CCEDIT_ERR_CUDA;
    }
    p.lkv[s] = a->lkv[s];
    p.kv_div[s] = a->kv_div[s];
    p.kv_mul[s] = a->kv_mul[s];
    p.kv_add[s] = a->kv_add[s];
    p.ntile[s] = (a->lkv[s] + kt - 1) / kt;
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
  // Heads per CTA: with one or two key tiles per head (text cross-attention: 77 keys) a CTA's life is all set-up
  // (TMEM allocation, barrier init, descriptor fetch, Q / K / V latency: ~6 us for ~1.5 us of work), so it takes several
  // heads in a row as long as enough CTAs remain to fill the 2 x 148 slots a few times over.
  p.hpc = 1;
  {
    static const int force = [] { const char* e = getenv("CCEDIT_ATTN_HPC"); return e ? atoi(e) : 0; }();
    const int sms = device_sm_count();
    const long long items = static_cast<long long>((a->lq + kTcTile - 1) / kTcTile) * a->frames;
    if (p.ntile[0] + p.ntile[1] <= 2 && sms > 0) {
      for (int h = 8; h >= 2; h >>= 1)
        if (a->heads % h == 0 && items * (a->heads / h) >= 3ll * 2 * sms) {
          p.hpc = h;
          break;
        }
    }
    if (force > 0 && a->heads % force == 0) p.hpc = force;
  }
  static const int emu = [] {                      // developer switch (A/B of the FMA-pipe exponentials at d = 40)
    const char* e = getenv("CCEDIT_ATTN_EMU");
    return e ? atoi(e) : kTcDefaultEmu;
  }();
  static const bool emu_wide = [] { const char* e = getenv("CCEDIT_ATTN_EMU"); return e && atoi(e) != 0; }();
  switch (ks) {
    case 1: return launch_tc2<1, kTcDefaultEmu>(maps, p, a->frames, a->heads, st);
    case 2: return launch_tc2<2, kTcDefaultEmu>(maps, p, a->frames, a->heads, st);
    case 3: return emu == 0 ? launch_tc2<3, 0>(maps, p, a->frames, a->heads, st)
                            : launch_tc2<3, kTcDefaultEmu>(maps, p, a->frames, a->heads, st);
    case 4: return launch_tc2<4, kTcDefaultEmu>(maps, p, a->frames, a->heads, st);
    // d = 80 / 160 (the network's other two widths): 64-key tiles, half the exponentials per tile - the MUFU pipe is not
    // what limits them, the FMA-pipe exponentials only add instructions (CCEDIT_ATTN_EMU=1 keeps them: A/B)
    case 5: return emu_wide ? launch_tc2<5, kTcDefaultEmu>(maps, p, a->frames, a->heads, st)
                            : launch_tc2<5, 0>(maps, p, a->frames, a->heads, st);
    case 6: return launch_tc2<6, kTcDefaultEmu>(maps, p, a->frames, a->heads, st);
    case 8: return launch_tc2<8, kTcDefaultEmu>(maps, p, a->frames, a->heads, st);
    default: return emu_wide ? launch_tc2<10, kTcDefaultEmu>(maps, p, a->frames, a->heads, st)
                             : launch_tc2<10, 0>(maps, p, a->frames, a->heads, st);
  }
}

}  // namespace ccedit

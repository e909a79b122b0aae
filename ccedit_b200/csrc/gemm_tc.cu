// Tap-GEMM on tcgen05 (sm_100a): persistent, warp-specialised, TMA-staged, TMEM double-buffered accumulators.
//
//   out[p, n] = epilogue( sum_tap sum_c A[p + tap, c] * W[n, tap, c] )
//
// One kernel covers every dense contraction of the CCEdit UNet / ControlNet forward (see include/ccedit_b200.h):
// nn.Linear and 1x1 convs (1 tap), 3x3 convs (9 taps, zero padding = TMA out-of-bounds fill), stride-2 convs
// (9 taps over parity planes), temporal Conv1d k=3 (3 taps on the T axis).  The A operand is addressed through a
// 5-D tensor map (C, d1..d4) so the same channels-last [B][T][H][W][C] buffer serves spatial and temporal layers
// without any transposition (the reference makes three full-tensor copies per spatial_temporal_forward,
// openaimodel.py:147,157,177).
//
// Roles (320 threads): warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + MMA issuer (one lane),
// warps 2..9 = epilogue, two warpgroups that split the 16-column chunks of a tile between them
// (TMEM -> registers -> bias/emb/SiLU/GEGLU/residual -> fp16 global stores).  Everything the epilogue reads from
// global memory for a tile (bias row into per-warp shared memory, this thread's residual segments into registers) is
// issued BEFORE it waits for the accumulator, so those latencies overlap the tile's MMAs instead of serialising
// behind them (measured: the un-prefetched epilogue held K=320 GEMMs at 10 % tensor-pipe activity).
#include "common.cuh"
#include "../../include/ccedit_b200.h"

#include <atomic>
#include <cstdlib>
#include <mutex>

namespace ccedit {

extern std::atomic<long long> g_launch_count;

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                  // 64 fp16 = 128 B = one SWIZZLE_128B row
constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KiB
constexpr int kMaxStages = 8;
constexpr int kGemmThreads = 320;
constexpr int kEpiWarps = 8;                 // two epilogue warpgroups
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;              // columns between the two accumulator buffers

struct GemmKParams {
  int box[4];
  int odim[4];
  int tiles[4];
  int n_tiles;
  int total_tiles;
  int total_pairs;    // CL = 2: ceil(m_tiles / 2) * n_tiles work items, each a pair of M-adjacent tiles sharing one W tile
  int ntaps;
  int kchunks;
  int bn;
  int stages;
  int taps[CCEDIT_MAX_TAPS][4];
  __half* out;
  long long ostr[4];
  const float* bias;
  const float* rowbias;
  int rb_dim, rb_div, n_out_total;
  const __half* res1;
  long long r1[4];
  const __half* res2;
  long long r2[4];
  int flags;
  uint32_t idesc;
  long long* trace;   // diagnostics: per-tile phase clocks of CTA 0 (ccedit_gemm_trace), or nullptr
  // staged (shared memory + TMA store) epilogue, see epilogue_staged
  int wcols;          // output columns per epilogue warp (half of the tile's output columns)
  int cb;             // columns per staging block (one TMA box): cb == wcols, or 64 with wcols == 128
  int nbuf;           // tile buffers (1 or 2)
  int tbuf_bytes;     // bytes of one tile buffer = 128 rows * 2 * wcols * 2
  uint32_t swz_mask;  // TMA swizzle of a staging block as an XOR mask on the 16-byte chunk index (7 / 3 / 1 / 0)
  // LayerNorm folded into the epilogue (template flag LNF): out = rstd[m] * (acc - mean[m] * colsum[n]) + bias[n]
  const float2* rowstats;
  const float* colsum;
  int rs_slots;       // > 1: rowstats holds rs_slots (sum, sum of squares) partials per row (another GEMM's stats_out)
  float rs_invc, rs_eps;
  // Row statistics of THIS GEMM's output for the LayerNorm that follows it: per row and per (n-tile, column half) the
  // epilogue thread writes (sum, sum of squares) of its final fp32 values to stats_out[(m * 2 * n_tiles + slot)].
  float2* stats_out;
};

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void add_res16(float (&v)[16], const __half* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    __half2 h = *reinterpret_cast<const __half2*>(&w[j]);
    float2 f = __half22float2(h);
    v[2 * j] += f.x;
    v[2 * j + 1] += f.y;
  }
}
__device__ __forceinline__ void add_res16(float (&v)[16], const uint4& a, const uint4& b) {
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
    v[2 * j] += f.x;
    v[2 * j + 1] += f.y;
  }
}
__device__ __forceinline__ void add_f32x16(float (&v)[16], const float* p) {
  const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float4 f = __ldg(q + j);
    v[4 * j] += f.x;
    v[4 * j + 1] += f.y;
    v[4 * j + 2] += f.z;
    v[4 * j + 3] += f.w;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Epilogue.  Measured with ccedit_gemm_trace: the epilogue is a serial instruction stream per warp (two warps per SM
// sub-partition), its cost is instruction count and - above all - instruction-cache footprint (a fully unrolled
// multi-mode epilogue spent > 60 % of every tile in fetch misses after taken branches).  Hence:
//   * the kernel is a template on (MODE, NRES): each launch contains only the code of its own epilogue;
//   * one compact ROLLED loop over 16-column chunks; the residual rows of chunk i+1 are requested while chunk i is
//     being processed, those of the first chunk (and the bias row, staged in per-warp shared memory) before the
//     wait on the MMAs of the tile;
//   * tile coordinates advance by a mixed-radix add (no divisions), per-thread offsets are split into a
//     thread-constant and a tile-uniform part.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kModePlain = 0, kModeGeglu = 1, kModeSilu = 2;
constexpr int kBiasFloats = 512;             // per epilogue warp: 256 bias values + 256 column sums (LNF)

__device__ __forceinline__ void ld_res(uint4 (&dst)[2], const __half* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  dst[0] = __ldg(q);
  dst[1] = __ldg(q + 1);
}
__device__ __forceinline__ void add_res(float (&v)[16], const uint4 (&r)[2]) {
  const uint32_t w[8] = {r[0].x, r[0].y, r[0].z, r[0].w, r[1].x, r[1].y, r[1].z, r[1].w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
    v[2 * j] += f.x;
    v[2 * j + 1] += f.y;
  }
}

// Tile sequence of a CTA.  CL = 1: tiles blockIdx.x, blockIdx.x + gridDim.x, ... (n_tile fastest), advanced as mixed-radix
// digits without divisions.  CL = 2 (cluster of two CTAs sharing every W tile by TMA multicast): the CLUSTER walks the
// pair sequence cid, cid + nclusters, ...; pair (n_tile, mp) = M tiles 2 mp and 2 mp + 1 of the same N tile, this CTA takes
// 2 mp + rank.  Digits are recomputed by division (long-K tiles only); an M tile past the end (odd tile count) has
// dig[4] >= tiles[3]: all its rows are out of range, TMA zero-fills its loads and clips its stores.
__device__ __forceinline__ void pair_digits(const GemmKParams& p, int pt, int crank, int (&dig)[5]) {
  dig[0] = pt % p.n_tiles;
  int m = 2 * (pt / p.n_tiles) + crank;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    dig[i + 1] = m % p.tiles[i];
    m /= p.tiles[i];
  }
  dig[4] = m;
}

template <int MODE, int NRES, bool LNF, int CL>
__device__ __forceinline__ void epilogue_loop(const GemmKParams& p, uint32_t tmem_base, uint64_t* tfull_bar,
                                              uint64_t* tempty_bar, float* sbias_all, int warp, int lane, int crank) {
  constexpr bool GEGLU = MODE == kModeGeglu;
  const int wq = warp & 3;            // TMEM lane quarter this warp may access
  const int eg = (warp - 2) >> 2;     // epilogue group: takes the 16-column chunks with (chunk index & 1) == eg
  const int row = wq * 32 + lane;
  float* sbias = sbias_all + (warp - 2) * kBiasFloats;
  float* scs = sbias + 256;             // column sums of the LayerNorm-folded weight (LNF)
  const int ncols_out = GEGLU ? p.bn / 2 : p.bn;
  const int nchunks = ncols_out >> 4;

  // thread-constant part of the coordinates / offsets
  int l[4];
  long long lo_o = 0, lo_r1 = 0, lo_r2 = 0;
  {
    int r = row;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      l[i] = r % p.box[i];
      r /= p.box[i];
      lo_o += static_cast<long long>(l[i]) * p.ostr[i];
      if (NRES >= 1) lo_r1 += static_cast<long long>(l[i]) * p.r1[i];
      if (NRES >= 2) lo_r2 += static_cast<long long>(l[i]) * p.r2[i];
    }
  }
  // tile counter as mixed-radix digits (n_tile, t0..t3); the per-iteration increment gridDim.x likewise
  int dig[5], inc[5], radix[5];
  {
    int t = blockIdx.x, g = gridDim.x;
    radix[0] = p.n_tiles;
#pragma unroll
    for (int i = 0; i < 4; ++i) radix[i + 1] = p.tiles[i];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      dig[i] = t % radix[i]; t /= radix[i];
      inc[i] = g % radix[i]; g /= radix[i];
    }
  }
  int as = 0;
  uint32_t aphase = 0;
  const bool tracing = p.trace != nullptr && blockIdx.x == 0 && warp == 2 && lane == 0;
  int titer = 0;
  int staged_ntile = -1;                // N tile whose bias row (and column sums) sit in sbias / scs
  const int it0 = CL >= 2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int itstep = CL >= 2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int ittotal = CL >= 2 ? p.total_pairs : p.total_tiles;
  for (int tile = it0; tile < ittotal; tile += itstep) {
    if (tracing && titer < 64) p.trace[16 * titer + 0] = clock64();
    if (CL >= 2) pair_digits(p, tile, crank, dig);
    const int n_tile = dig[0];
    bool valid = true;
    long long off_o = lo_o, off_r1 = lo_r1, off_r2 = lo_r2;
    int rb_c = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int base = dig[i + 1] * p.box[i];
      valid = valid && (base + l[i] < p.odim[i]);
      off_o += static_cast<long long>(base) * p.ostr[i];
      if (NRES >= 1) off_r1 += static_cast<long long>(base) * p.r1[i];
      if (NRES >= 2) off_r2 += static_cast<long long>(base) * p.r2[i];
      if (i == p.rb_dim) rb_c = min(base + l[i], p.odim[i] - 1);   // clamp: rows of the tile padding must not index past rowbias
    }
    const int col0_out = n_tile * ncols_out;
    float rstd = 1.f, nmr = 0.f;                              // LNF: 1/std and -mean/std of this thread's row
    if (LNF) {
      const int m = min(dig[1] * p.box[0] + l[0], p.odim[0] - 1);
      if (p.rs_slots > 1) {                                   // partial sums from the producing GEMM's epilogue
        float ss = 0.f, qq = 0.f;
        for (int k = 0; k < p.rs_slots; ++k) {
          const float2 pv = __ldg(p.rowstats + static_cast<long long>(m) * p.rs_slots + k);
          ss += pv.x;
          qq += pv.y;
        }
        const float mean = ss * p.rs_invc;
        rstd = rsqrtf(fmaxf(fmaf(-mean, mean, qq * p.rs_invc), 0.f) + p.rs_eps);
        nmr = -mean * rstd;
      } else {
        const float2 st = __ldg(p.rowstats + m);
        rstd = st.y;
        nmr = -st.x * st.y;
      }
    }
    __half* optr = p.out + off_o + col0_out;
    const __half* r1ptr = p.res1 + off_r1 + col0_out;   // dereferenced only if NRES >= 1 and valid
    const __half* r2ptr = p.res2 + off_r2 + col0_out;

    // ---- global reads of this tile, issued before the wait on its MMAs ----
    uint4 c1[2], c2[2], n1[2], n2[2];
    if (NRES >= 1 && valid && eg < nchunks) ld_res(c1, r1ptr + eg * 16);
    if (NRES >= 2 && valid && eg < nchunks) ld_res(c2, r2ptr + eg * 16);
    const float* rbptr = nullptr;
    {
      bool rb_uniform = false;
      int rb_row = 0;
      if (!GEGLU && p.rowbias) {
        rb_row = p.rb_div == 1 ? rb_c : rb_c / p.rb_div;
        rb_uniform = __all_sync(0xffffffffu, rb_row == __shfl_sync(0xffffffffu, rb_row, 0));
        if (!rb_uniform && valid) rbptr = p.rowbias + static_cast<long long>(rb_row) * p.n_out_total + col0_out;
      }
      // the staged row only depends on the N tile (and on the time-embedding row): a CTA whose tiles all share their N
      // tile - n_tiles divides the grid, e.g. every N = 320 / 640 GEMM - stages it once (it cost 400-1000 clocks per tile)
      if (n_tile != staged_ntile || p.rowbias) {
        staged_ntile = n_tile;
        __syncwarp();
        float bv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {                      // bn <= 256: all loads in flight together
          const int i = lane + 32 * k;
          bv[k] = (p.bias && i < p.bn) ? __ldg(p.bias + n_tile * p.bn + i) : 0.f;
          if (rb_uniform && i < p.bn) bv[k] += __ldg(p.rowbias + static_cast<long long>(rb_row) * p.n_out_total + col0_out + i);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) sbias[lane + 32 * k] = bv[k];
        if (LNF) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int i = lane + 32 * k;
            scs[i] = i < p.bn ? __ldg(p.colsum + n_tile * p.bn + i) : 0.f;
          }
        }
      }
    }
    __syncwarp();
    if (tracing && titer < 64) p.trace[16 * titer + 1] = clock64();

    mbar_wait(&tfull_bar[as], aphase);
    tcgen05_fence_after();
    if (tracing && titer < 64) p.trace[16 * titer + 2] = clock64();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + static_cast<uint32_t>(as * kAccStride);
    float st_s = 0.f, st_q = 0.f;                             // stats_out: this thread's partial row sums

#pragma unroll 1
    for (int ci = eg; ci < nchunks; ci += 2) {
      const int c = ci * 16;
      uint32_t r[16];
      float v[16];
      tmem_ld_32x32b_x16(taddr + c, r);
      if (GEGLU) {
        uint32_t g[16];
        tmem_ld_32x32b_x16(taddr + ncols_out + c, g);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 bv = *reinterpret_cast<const float4*>(sbias + c + j);
          float4 bg = *reinterpret_cast<const float4*>(sbias + ncols_out + c + j);
          float a[4] = {__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])};
          float e[4] = {__uint_as_float(g[j]), __uint_as_float(g[j + 1]), __uint_as_float(g[j + 2]), __uint_as_float(g[j + 3])};
          if (LNF) {
            const float4 cv = *reinterpret_cast<const float4*>(scs + c + j);
            const float4 cg = *reinterpret_cast<const float4*>(scs + ncols_out + c + j);
            bv = make_float4(fmaf(nmr, cv.x, bv.x), fmaf(nmr, cv.y, bv.y), fmaf(nmr, cv.z, bv.z), fmaf(nmr, cv.w, bv.w));
            bg = make_float4(fmaf(nmr, cg.x, bg.x), fmaf(nmr, cg.y, bg.y), fmaf(nmr, cg.z, bg.z), fmaf(nmr, cg.w, bg.w));
#pragma unroll
            for (int q = 0; q < 4; ++q) { a[q] *= rstd; e[q] *= rstd; }
          }
          v[j] = (a[0] + bv.x) * gelu_erf_f(e[0] + bg.x);
          v[j + 1] = (a[1] + bv.y) * gelu_erf_f(e[1] + bg.y);
          v[j + 2] = (a[2] + bv.z) * gelu_erf_f(e[2] + bg.z);
          v[j + 3] = (a[3] + bv.w) * gelu_erf_f(e[3] + bg.w);
        }
      } else {
        // request the residual rows of this thread's next chunk while this one is processed
        if (NRES >= 1 && valid && ci + 2 < nchunks) ld_res(n1, r1ptr + c + 32);
        if (NRES >= 2 && valid && ci + 2 < nchunks) ld_res(n2, r2ptr + c + 32);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 bv = *reinterpret_cast<const float4*>(sbias + c + j);
          if (LNF) {
            const float4 cv = *reinterpret_cast<const float4*>(scs + c + j);
            bv = make_float4(fmaf(nmr, cv.x, bv.x), fmaf(nmr, cv.y, bv.y), fmaf(nmr, cv.z, bv.z), fmaf(nmr, cv.w, bv.w));
          }
          v[j] = fmaf(__uint_as_float(r[j]), rstd, bv.x);          // rstd == 1 unless LNF
          v[j + 1] = fmaf(__uint_as_float(r[j + 1]), rstd, bv.y);
          v[j + 2] = fmaf(__uint_as_float(r[j + 2]), rstd, bv.z);
          v[j + 3] = fmaf(__uint_as_float(r[j + 3]), rstd, bv.w);
        }
      }
      if (valid) {
        if (rbptr) add_f32x16(v, rbptr + c);
        if (MODE == kModeSilu) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = silu_f(v[j]);
        }
        if (NRES >= 1) add_res(v, c1);
        if (NRES >= 2) add_res(v, c2);
        if (MODE == kModePlain && p.stats_out) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            st_s += v[j];
            st_q = fmaf(v[j], v[j], st_q);
          }
        }
        uint4 s0, s1;
        s0.x = pack_half2(v[0], v[1]);
        s0.y = pack_half2(v[2], v[3]);
        s0.z = pack_half2(v[4], v[5]);
        s0.w = pack_half2(v[6], v[7]);
        s1.x = pack_half2(v[8], v[9]);
        s1.y = pack_half2(v[10], v[11]);
        s1.z = pack_half2(v[12], v[13]);
        s1.w = pack_half2(v[14], v[15]);
        uint4* o4 = reinterpret_cast<uint4*>(optr + c);
        o4[0] = s0;
        o4[1] = s1;
      }
      if (NRES >= 1) { c1[0] = n1[0]; c1[1] = n1[1]; }
      if (NRES >= 2) { c2[0] = n2[0]; c2[1] = n2[1]; }
    }
    if (MODE == kModePlain && p.stats_out && valid)
      p.stats_out[static_cast<long long>(dig[1] * p.box[0] + l[0]) * (2 * p.n_tiles) + 2 * n_tile + eg] = make_float2(st_s, st_q);
    if (tracing && titer < 64) p.trace[16 * titer + 3] = clock64();
    tcgen05_fence_before();
    __syncwarp();
    if (lane == 0) {                    // CL = 3: the accumulators of BOTH CTAs are released on the leader's barrier
      if (CL == 3) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[as]), 0));
      else mbar_arrive(&tempty_bar[as]);
    }
    if (tracing && titer < 64) p.trace[16 * titer + 4] = clock64();
    ++titer;
    as ^= 1;
    if (as == 0) aphase ^= 1u;
    // next tile: mixed-radix add with carry
    if (CL == 1) {
      int carry = 0;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        int d = dig[i] + inc[i] + carry;
        carry = d >= radix[i] ? 1 : 0;
        dig[i] = d - (carry ? radix[i] : 0);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Staged epilogue (small-K GEMMs, where the epilogue - not the MMAs - paces the tile; ccedit_gemm_trace: the direct
// epilogue above spends ~940 clocks per 16-column chunk because thread = row makes every 16-byte global access of a
// warp touch 32 different lines).  Here the only global traffic of the epilogue is TMA:
//   * each epilogue warp owns 32 rows x W columns of the tile (warpgroup eg takes columns [eg*W, eg*W + W)) and a
//     private region [32][W] fp16 of a tile buffer in shared memory (TMA-swizzled when a row is 32/64/128 bytes, so the
//     row-per-thread 16-byte accesses are bank-conflict free);
//   * res1 of tile i+1 is TMA-loaded into the region by the warp's own lane 0 while tile i is processed (two tile
//     buffers; with one buffer - BN = 256 - right after the store of tile i has left shared memory);
//   * the thread adds bias / time-embedding row / activation / residual to its TMEM row and overwrites the region in
//     place; lane 0 stores the 32 x W box with one TMA store (rows past the end of the tensor are clipped by TMA).
// ---------------------------------------------------------------------------------------------------------------
template <int MODE, int NRES, bool LNF, int CL>
__device__ __forceinline__ void epilogue_staged(const GemmKParams& p, const CUtensorMap* tmOut, const CUtensorMap* tmRes,
                                                uint32_t tmem_base, uint64_t* tfull_bar, uint64_t* tempty_bar,
                                                uint64_t* res_bar_all, float* sbias_all, uint8_t* tbuf, int warp,
                                                int lane, int crank) {
  constexpr bool GEGLU = MODE == kModeGeglu;
  const int wq = warp & 3;            // TMEM lane quarter this warp may access
  const int eg = (warp - 2) >> 2;     // column half of the tile
  const int ew = warp - 2;            // 0..7
  const int row = wq * 32 + lane;
  const int W = p.wcols;
  const int ncols_out = 2 * W;
  const int nch = W >> 4;
  const int nblk = W / p.cb;
  const int blk_bytes = 32 * p.cb * 2;
  const uint32_t row_bytes = static_cast<uint32_t>(p.cb) * 2u;
  float* sbias = sbias_all + ew * kBiasFloats;
  float* scs = sbias + 256;             // column sums of the LayerNorm-folded weight (LNF)
  uint64_t* rbar = res_bar_all + ew * 2;
  uint8_t* region0 = tbuf + ew * (32 * W * 2);

  int l[4], qoff[4];
  long long lo_r2 = 0;
  {
    int r = row, r0 = wq * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      l[i] = r % p.box[i];
      r /= p.box[i];
      qoff[i] = r0 % p.box[i];
      r0 /= p.box[i];
      if (NRES >= 2) lo_r2 += static_cast<long long>(l[i]) * p.r2[i];
    }
  }
  int dig[5], inc[5], radix[5];
  {
    int t = blockIdx.x, g = gridDim.x;
    radix[0] = p.n_tiles;
#pragma unroll
    for (int i = 0; i < 4; ++i) radix[i + 1] = p.tiles[i];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      dig[i] = t % radix[i]; t /= radix[i];
      inc[i] = g % radix[i]; g /= radix[i];
    }
  }
  const int it0 = CL >= 2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int itstep = CL >= 2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int ittotal = CL >= 2 ? p.total_pairs : p.total_tiles;
  if (CL >= 2) pair_digits(p, it0, crank, dig);
  // residual of the first tile
  // (all TMA / bulk-group / barrier operations of this epilogue are warp-collective: converged warp, elected lane)
  if (NRES >= 1 && it0 < ittotal) {
    mbar_arrive_expect_tx_warp(&rbar[0], static_cast<uint32_t>(32 * W * 2));
    for (int bk = 0; bk < nblk; ++bk)
      tma_load_5d_warp(region0 + bk * blk_bytes, tmRes, &rbar[0], dig[0] * ncols_out + eg * W + bk * p.cb,
                       dig[1] * p.box[0] + qoff[0], dig[2] * p.box[1] + qoff[1], dig[3] * p.box[2] + qoff[2],
                       dig[4] * p.box[3] + qoff[3]);
  }
  __syncwarp();

  int as = 0, it = 0;
  int staged_ntile = -1;                // N tile whose bias row (and column sums) sit in sbias / scs
  uint32_t aphase = 0, rph0 = 0, rph1 = 0;
  const bool tracing = p.trace != nullptr && blockIdx.x == 0 && warp == 2 && lane == 0;
  for (int tile = it0; tile < ittotal; tile += itstep, ++it) {
    if (tracing && it < 64) p.trace[16 * it + 0] = clock64();
    const int b = (p.nbuf == 2) ? (it & 1) : 0;
    const int n_tile = dig[0];
    bool valid = true;
    long long off_r2 = lo_r2;
    int rb_c = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int base = dig[i + 1] * p.box[i];
      valid = valid && (base + l[i] < p.odim[i]);
      if (NRES >= 2) off_r2 += static_cast<long long>(base) * p.r2[i];
      if (i == p.rb_dim) rb_c = min(base + l[i], p.odim[i] - 1);
    }
    const int col0_out = n_tile * ncols_out + eg * W;        // first output column of this warp
    float rstd = 1.f, nmr = 0.f;                              // LNF: 1/std and -mean/std of this thread's row
    if (LNF) {
      const int m = min(dig[1] * p.box[0] + l[0], p.odim[0] - 1);
      if (p.rs_slots > 1) {                                   // partial sums from the producing GEMM's epilogue
        float ss = 0.f, qq = 0.f;
        for (int k = 0; k < p.rs_slots; ++k) {
          const float2 pv = __ldg(p.rowstats + static_cast<long long>(m) * p.rs_slots + k);
          ss += pv.x;
          qq += pv.y;
        }
        const float mean = ss * p.rs_invc;
        rstd = rsqrtf(fmaxf(fmaf(-mean, mean, qq * p.rs_invc), 0.f) + p.rs_eps);
        nmr = -mean * rstd;
      } else {
        const float2 st = __ldg(p.rowstats + m);
        rstd = st.y;
        nmr = -st.x * st.y;
      }
    }
    const __half* r2ptr = p.res2 + off_r2 + col0_out;        // dereferenced only if NRES >= 2 and valid
    // next tile (mixed-radix add with carry)
    int ndig[5];
    if (CL >= 2) {
      pair_digits(p, tile + itstep, crank, ndig);
    } else {
      int carry = 0;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        int d = dig[i] + inc[i] + carry;
        carry = d >= radix[i] ? 1 : 0;
        ndig[i] = d - (carry ? radix[i] : 0);
      }
    }
    const bool has_next = tile + itstep < ittotal;

    // ---- bias row of this warp's columns -> per-warp shared memory (GEGLU: W value columns, then W gate columns) ----
    uint4 c2[2], n2[2];
    if (NRES >= 2 && valid) ld_res(c2, r2ptr);
    const float* rbptr = nullptr;
    {
      bool rb_uniform = false;
      int rb_row = 0;
      if (!GEGLU && p.rowbias) {
        rb_row = p.rb_div == 1 ? rb_c : rb_c / p.rb_div;
        rb_uniform = __all_sync(0xffffffffu, rb_row == __shfl_sync(0xffffffffu, rb_row, 0));
        if (!rb_uniform && valid) rbptr = p.rowbias + static_cast<long long>(rb_row) * p.n_out_total + col0_out;
      }
      if (n_tile != staged_ntile || p.rowbias) {          // see epilogue_loop: once per N tile, not once per tile
        staged_ntile = n_tile;
        __syncwarp();
        const int nb = GEGLU ? 2 * W : W;
        float bv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int i = lane + 32 * k;
          const int src = n_tile * p.bn + (GEGLU ? (i < W ? eg * W + i : ncols_out + eg * W + (i - W)) : eg * W + i);
          bv[k] = (p.bias && i < nb) ? __ldg(p.bias + src) : 0.f;
          if (rb_uniform && i < nb) bv[k] += __ldg(p.rowbias + static_cast<long long>(rb_row) * p.n_out_total + col0_out + i);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) sbias[lane + 32 * k] = bv[k];
        if (LNF) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int i = lane + 32 * k;
            const int src = n_tile * p.bn + (GEGLU ? (i < W ? eg * W + i : ncols_out + eg * W + (i - W)) : eg * W + i);
            scs[i] = i < nb ? __ldg(p.colsum + src) : 0.f;
          }
        }
      }
    }
    __syncwarp();
    if (tracing && it < 64) p.trace[16 * it + 1] = clock64();

    mbar_wait(&tfull_bar[as], aphase);
    tcgen05_fence_after();
    if (tracing && it < 64) p.trace[16 * it + 2] = clock64();
    uint8_t* region = region0 + b * p.tbuf_bytes;
    if (NRES >= 1) {
      if (p.nbuf == 2 && has_next) {
        bulk_wait_group_read_warp<0>();                    // the store of tile it-1 has left the other buffer
        uint8_t* other = region0 + (b ^ 1) * p.tbuf_bytes;
        mbar_arrive_expect_tx_warp(&rbar[b ^ 1], static_cast<uint32_t>(32 * W * 2));
        for (int bk = 0; bk < nblk; ++bk)
          tma_load_5d_warp(other + bk * blk_bytes, tmRes, &rbar[b ^ 1], ndig[0] * ncols_out + eg * W + bk * p.cb,
                           ndig[1] * p.box[0] + qoff[0], ndig[2] * p.box[1] + qoff[1], ndig[3] * p.box[2] + qoff[2],
                           ndig[4] * p.box[3] + qoff[3]);
      }
      mbar_wait(&rbar[b], b ? rph1 : rph0);
      if (b) rph1 ^= 1u; else rph0 ^= 1u;
    } else {
      if (p.nbuf == 2) bulk_wait_group_read_warp<1>(); else bulk_wait_group_read_warp<0>();   // this buffer's last store has been read
    }
    __syncwarp();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + static_cast<uint32_t>(as * kAccStride) +
                           static_cast<uint32_t>(eg * W);
    float st_s = 0.f, st_q = 0.f;                             // stats_out: this thread's partial row sums

#pragma unroll 1
    for (int ci = 0; ci < nch; ++ci) {
      const int c = ci * 16;
      uint32_t r[16];
      float v[16];
      tmem_ld_32x32b_x16(taddr + c, r);
      // this thread's 32 bytes of the staging row: block (c / cb), 16-byte chunks k, k+1 (TMA swizzle = XOR on the chunk)
      const int bk = c / p.cb;
      const uint32_t inrow = static_cast<uint32_t>(c - bk * p.cb) * 2u;
      const uint32_t o0 = static_cast<uint32_t>(lane) * row_bytes + inrow;
      const uint32_t sw = ((o0 >> 7) & p.swz_mask) << 4;
      uint8_t* blk = region + bk * blk_bytes;
      uint4* s0 = reinterpret_cast<uint4*>(blk + (o0 ^ sw));
      uint4* s1 = reinterpret_cast<uint4*>(blk + ((o0 + 16u) ^ sw));
      uint4 c1[2];
      if (NRES >= 1) { c1[0] = *s0; c1[1] = *s1; }
      if (GEGLU) {
        uint32_t g[16];
        tmem_ld_32x32b_x16(taddr + ncols_out + c, g);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 bv = *reinterpret_cast<const float4*>(sbias + c + j);
          float4 bg = *reinterpret_cast<const float4*>(sbias + W + c + j);
          float a[4] = {__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])};
          float e[4] = {__uint_as_float(g[j]), __uint_as_float(g[j + 1]), __uint_as_float(g[j + 2]), __uint_as_float(g[j + 3])};
          if (LNF) {
            const float4 cv = *reinterpret_cast<const float4*>(scs + c + j);
            const float4 cg = *reinterpret_cast<const float4*>(scs + W + c + j);
            bv = make_float4(fmaf(nmr, cv.x, bv.x), fmaf(nmr, cv.y, bv.y), fmaf(nmr, cv.z, bv.z), fmaf(nmr, cv.w, bv.w));
            bg = make_float4(fmaf(nmr, cg.x, bg.x), fmaf(nmr, cg.y, bg.y), fmaf(nmr, cg.z, bg.z), fmaf(nmr, cg.w, bg.w));
#pragma unroll
            for (int q = 0; q < 4; ++q) { a[q] *= rstd; e[q] *= rstd; }
          }
          v[j] = (a[0] + bv.x) * gelu_erf_f(e[0] + bg.x);
          v[j + 1] = (a[1] + bv.y) * gelu_erf_f(e[1] + bg.y);
          v[j + 2] = (a[2] + bv.z) * gelu_erf_f(e[2] + bg.z);
          v[j + 3] = (a[3] + bv.w) * gelu_erf_f(e[3] + bg.w);
        }
      } else {
        if (NRES >= 2 && valid && ci + 1 < nch) ld_res(n2, r2ptr + c + 16);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 bv = *reinterpret_cast<const float4*>(sbias + c + j);
          if (LNF) {
            const float4 cv = *reinterpret_cast<const float4*>(scs + c + j);
            bv = make_float4(fmaf(nmr, cv.x, bv.x), fmaf(nmr, cv.y, bv.y), fmaf(nmr, cv.z, bv.z), fmaf(nmr, cv.w, bv.w));
          }
          v[j] = fmaf(__uint_as_float(r[j]), rstd, bv.x);          // rstd == 1 unless LNF
          v[j + 1] = fmaf(__uint_as_float(r[j + 1]), rstd, bv.y);
          v[j + 2] = fmaf(__uint_as_float(r[j + 2]), rstd, bv.z);
          v[j + 3] = fmaf(__uint_as_float(r[j + 3]), rstd, bv.w);
        }
      }
      if (rbptr) add_f32x16(v, rbptr + c);
      if (MODE == kModeSilu) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = silu_f(v[j]);
      }
      if (NRES >= 1) add_res(v, c1);
      if (NRES >= 2 && valid) add_res(v, c2);
      if (MODE == kModePlain && p.stats_out) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          st_s += v[j];
          st_q = fmaf(v[j], v[j], st_q);
        }
      }
      uint4 q0, q1;
      q0.x = pack_half2(v[0], v[1]);
      q0.y = pack_half2(v[2], v[3]);
      q0.z = pack_half2(v[4], v[5]);
      q0.w = pack_half2(v[6], v[7]);
      q1.x = pack_half2(v[8], v[9]);
      q1.y = pack_half2(v[10], v[11]);
      q1.z = pack_half2(v[12], v[13]);
      q1.w = pack_half2(v[14], v[15]);
      *s0 = q0;
      *s1 = q1;
      if (NRES >= 2) { c2[0] = n2[0]; c2[1] = n2[1]; }
    }
    if (MODE == kModePlain && p.stats_out && valid)
      p.stats_out[static_cast<long long>(dig[1] * p.box[0] + l[0]) * (2 * p.n_tiles) + 2 * n_tile + eg] = make_float2(st_s, st_q);
    if (tracing && it < 64) p.trace[16 * it + 3] = clock64();
    tcgen05_fence_before();
    fence_proxy_async_smem();                                // this thread's staging writes -> visible to the TMA store
    __syncwarp();
    if (CL == 3) mbar_arrive_cluster_warp(mapa_shared(smem_u32(&tempty_bar[as]), 0));
    else mbar_arrive_warp(&tempty_bar[as]);
    for (int bk = 0; bk < nblk; ++bk)
      tma_store_5d_warp(tmOut, region + bk * blk_bytes, n_tile * ncols_out + eg * W + bk * p.cb, dig[1] * p.box[0] + qoff[0],
                        dig[2] * p.box[1] + qoff[1], dig[3] * p.box[2] + qoff[2], dig[4] * p.box[3] + qoff[3]);
    bulk_commit_group_warp();
    if (NRES >= 1 && p.nbuf == 1 && has_next) {
      bulk_wait_group_read_warp<0>();
      mbar_arrive_expect_tx_warp(&rbar[0], static_cast<uint32_t>(32 * W * 2));
      for (int bk = 0; bk < nblk; ++bk)
        tma_load_5d_warp(region0 + bk * blk_bytes, tmRes, &rbar[0], ndig[0] * ncols_out + eg * W + bk * p.cb,
                         ndig[1] * p.box[0] + qoff[0], ndig[2] * p.box[1] + qoff[1], ndig[3] * p.box[2] + qoff[2],
                         ndig[4] * p.box[3] + qoff[3]);
    }
    __syncwarp();
    if (tracing && it < 64) p.trace[16 * it + 4] = clock64();
    as ^= 1;
    if (as == 0) aphase ^= 1u;
#pragma unroll
    for (int i = 0; i < 5; ++i) dig[i] = ndig[i];
  }
  bulk_wait_group_read_warp<0>();                            // shared memory must outlive the last store's read
  __syncwarp();
}

template <int MODE, int NRES, bool STAGED, bool LNF, int CL>
__global__ void __launch_bounds__(kGemmThreads, 1)
tap_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes,
                const __grid_constant__ GemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);  // SWIZZLE_128B wants 1024 B alignment

  const int stage_bytes = kABytes + (CL == 3 ? p.bn / 2 : p.bn) * kBlockK * 2;   // CL = 3: this CTA holds half of the W tile
  // [stages x (A, B)] [STAGED: nbuf tile buffers] [barriers, 512 B] [per-warp bias rows]; every part a multiple of 1 KiB
  uint8_t* tbuf = smem + p.stages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tbuf + (STAGED ? p.nbuf * p.tbuf_bytes : 0));
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint64_t* res_bar = tempty_bar + 4;                                  // [kEpiWarps][2] (STAGED with a residual)
  float* sbias_all = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 512);  // [kEpiWarps][kBiasFloats], 16 B aligned

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // CL = 2: the two CTAs of a cluster work on M-adjacent tiles of the same N tile; each loads half of the W tile and
  // multicasts it to both, so a W tile crosses the L2 -> SM path once per 256 rows instead of once per 128 (the big-K
  // GEMMs run at the 64 B/clk/SM limit of that path).  A slot may be refilled once BOTH CTAs have consumed it: the MMA
  // commits arrive on the `empty` barrier of both CTAs (count 2).
  // CL = 3 (experimental, CCEDIT_GEMM_CLUSTER=3): a CTA PAIR on tcgen05 cta_group::2 computes a 256 x BN tile; each CTA
  // loads its own A tile and half of the W tile (no multicast: the pair's MMA reads W from both shared memories), the
  // leader (rank 0) issues one M = 256 instruction per k-step.  `full` lives in the leader: its arrive.expect_tx covers
  // both CTAs' bytes and both CTAs' TMA loads (cp.async.bulk.tensor.cta_group::2) signal it; `empty` / `tfull` are
  // signalled in both CTAs by multicast commits; `tempty` collects the epilogue warps of both CTAs in the leader.
  const int crank = CL >= 2 ? static_cast<int>(cluster_ctarank()) : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], CL == 2 ? 2 : 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], CL == 3 ? 2 * kEpiWarps : kEpiWarps);
    }
    if (STAGED) {
      tma_prefetch_desc(&tmOut);
      if (NRES >= 1) tma_prefetch_desc(&tmRes);
      for (int s = 0; s < 2 * kEpiWarps; ++s) mbar_init(&res_bar[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CL == 3) {
      tmem_alloc_pair(tmem_slot, kTmemCols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, kTmemCols);
      tmem_relinquish();
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (CL >= 2) cluster_sync_all();          // the peer's barriers are initialised before anything can arrive on them
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch) touches no global data and may
  // run while the previous kernel of the stream drains its last tiles; nothing below may start before that kernel has
  // completed and flushed (both instructions are no-ops for a launch without the attribute).  The dependents of THIS
  // kernel may be scheduled from here on: they too stop at their own griddepcontrol.wait.
  griddep_wait();
  griddep_launch_dependents();

  const int kblocks = p.ntaps * p.kchunks;
  const int it0 = CL >= 2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int itstep = CL >= 2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int ittotal = CL >= 2 ? p.total_pairs : p.total_tiles;

  if (warp == 0) {
    // ===================== TMA producer (whole warp, one elected lane issues) =====================
    int stage = 0;
    uint32_t phase = 0;
    // An mbarrier wait costs 100-200 clocks even when its phase completed long ago (TRYWAIT latency); with one wait per
    // 64-wide k-block the producer (and the MMA issuer below) needed ~350 clocks per block - more than the 320 clocks
    // the tensor core spends on a 128 x 160 x 64 block, which is why the BN = 160 GEMMs ran at 67-80 % pipe activity
    // while BN = 256 (512 clocks per block) did not care.  The NEXT slot's barrier is therefore probed with a
    // non-blocking test issued ahead of this block's TMA / MMA instructions; the blocking wait is only the fallback.
    bool slot_free = false;
    const int half_rows = p.bn >> 1;         // CL = 2: W rows this CTA fetches (and multicasts) per stage
    for (int tile = it0; tile < ittotal; tile += itstep) {
      const int n_tile = tile % p.n_tiles;
      int m = CL >= 2 ? 2 * (tile / p.n_tiles) + crank : tile / p.n_tiles;
      int o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        o[i] = (i < 3 ? m % p.tiles[i] : m) * p.box[i];      // past-the-end M tiles (CL = 2, odd count): TMA zero fill
        m /= p.tiles[i];
      }
      for (int tap = 0; tap < p.ntaps; ++tap) {
        const int c1 = o[0] + p.taps[tap][0], c2 = o[1] + p.taps[tap][1];
        const int c3 = o[2] + p.taps[tap][2], c4 = o[3] + p.taps[tap][3];
        for (int kc = 0; kc < p.kchunks; ++kc) {
          if (!slot_free) mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * stage_bytes;
          {
            const int ns = stage + 1 == p.stages ? 0 : stage + 1;
            slot_free = mbar_test_wait(&empty_bar[ns], (ns == 0 ? phase ^ 1u : phase) ^ 1u);
          }
          if (CL == 3) {
            // the leader's arrive.expect_tx accounts for both CTAs' bytes (count 1: the peer does not arrive at all)
            const uint32_t lead_full = mapa_shared(smem_u32(&full_bar[stage]), 0);
            if (crank == 0) mbar_arrive_expect_tx_warp(&full_bar[stage], 2u * static_cast<uint32_t>(stage_bytes));
            tma_load_5d_pair_warp(sa, &tmA, lead_full, kc * kBlockK, c1, c2, c3, c4);
            tma_load_2d_pair_warp(sa + kABytes, &tmB, lead_full, (tap * p.kchunks + kc) * kBlockK,
                                  n_tile * p.bn + crank * half_rows);
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1u;
            }
            continue;
          }
          mbar_arrive_expect_tx_warp(&full_bar[stage], static_cast<uint32_t>(stage_bytes));
          tma_load_5d_warp(sa, &tmA, &full_bar[stage], kc * kBlockK, c1, c2, c3, c4);
          if (CL == 2)
            tma_load_2d_multicast_warp(sa + kABytes + crank * half_rows * 128, &tmB, &full_bar[stage],
                                       (tap * p.kchunks + kc) * kBlockK, n_tile * p.bn + crank * half_rows, 3);
          else
            tma_load_2d_warp(sa + kABytes, &tmB, &full_bar[stage], (tap * p.kchunks + kc) * kBlockK, n_tile * p.bn);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1 && (CL != 3 || crank == 0)) {
    // ===================== MMA issuer (whole warp, one elected lane issues; CL = 3: the leader CTA only) =====================
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    int mt = 0;
    bool slot_full = false;                  // probe of the next slot's `full` barrier (see the producer)
    const bool tr = p.trace && blockIdx.x == 0 && lane == 0;
    for (int tile = it0; tile < ittotal; tile += itstep) {
      mbar_wait(&tempty_bar[as], aphase ^ 1u);
      tcgen05_fence_after();
      if (tr && mt < 64) p.trace[16 * mt + 5] = clock64();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * kAccStride);
      for (int kb = 0; kb < kblocks; ++kb) {
        if (!slot_full) mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        {
          const int ns = stage + 1 == p.stages ? 0 : stage + 1;
          slot_full = mbar_test_wait(&full_bar[ns], ns == 0 ? phase ^ 1u : phase);
        }
        if (kb == 0 && tr && mt < 64) p.trace[16 * mt + 7] = clock64();
        const uint32_t sa = smem_u32(smem + stage * stage_bytes);
        const uint64_t adesc = umma_desc_k_sw128(sa);
        const uint64_t bdesc = umma_desc_k_sw128(sa + kABytes);
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) {
          // advance 16 elements (32 B) inside the 128 B swizzle row: +2 in 16-byte units
          if (CL == 3) umma_f16_ss_pair_warp(d_tmem, adesc + 2u * k, bdesc + 2u * k, p.idesc, (kb | k) != 0 ? 1u : 0u);
          else umma_f16_ss_warp(d_tmem, adesc + 2u * k, bdesc + 2u * k, p.idesc, (kb | k) != 0 ? 1u : 0u);
        }
        if (CL == 3) umma_commit_pair_warp(&empty_bar[stage], 3);        // frees the slot in both CTAs of the pair
        else if (CL == 2) umma_commit_multicast_warp(&empty_bar[stage], 3);   // the slot is free in BOTH CTAs' books
        else umma_commit_warp(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (CL == 3) umma_commit_pair_warp(&tfull_bar[as], 3);   // both CTAs' epilogues
      else umma_commit_warp(&tfull_bar[as]);  // accumulator complete -> epilogue
      if (tr && mt < 64) p.trace[16 * mt + 6] = clock64();
      ++mt;
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
  } else if (warp >= 2) {
    // ===================== epilogue warps =====================
    if (STAGED)
      epilogue_staged<MODE, NRES, LNF, CL>(p, &tmOut, &tmRes, tmem_base, tfull_bar, tempty_bar, res_bar, sbias_all, tbuf, warp,
                                           lane, crank);
    else
      epilogue_loop<MODE, NRES, LNF, CL>(p, tmem_base, tfull_bar, tempty_bar, sbias_all, warp, lane, crank);
  }

  tcgen05_fence_before();
  __syncthreads();
  if (CL >= 2) cluster_sync_all();          // no CTA leaves while its peer may still multicast into it / arrive on its barriers
  if (warp == 1) {
    tcgen05_fence_after();
    if (CL == 3) tmem_dealloc_pair(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  });
  return fn;
}

// Per-device caches: function attributes and SM counts belong to a device, and one process may drive several.
int current_device() {
  int dev = 0;
  return cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < kMaxDevices ? dev : -1;
}
int pdl_mask() {
  // default 3: tap-GEMM + flash attention.  Measured on one box (ms per sampler step): 0: 214.56, 1: 213.2, 3: 212.8,
  // 5: 213.9, 9: 215.1 - early-launched GroupNorm / short attention CTAs cost more than their prologues save.
  static const int m = [] { const char* e = getenv("CCEDIT_PDL"); return e ? atoi(e) : 3; }();
  return m;
}
int device_sm_count() {
  static std::atomic<int> sms[kMaxDevices];
  const int dev = current_device();
  if (dev < 0) return -1;
  int v = sms[dev].load(std::memory_order_relaxed);
  if (v <= 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    sms[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

extern long long* g_trace_buf;

// Launch one instantiation; cl = 2 launches clusters of two CTAs (cudaLaunchAttributeClusterDimension).  Function
// attributes and the cluster capacity belong to a device: cached per device ordinal.
template <int MODE, int NRES, bool STAGED, bool LNF, int CL>
static cudaError_t launch_gemm_t(const CUtensorMap* tm, const GemmKParams& p, int work_items, int smem_bytes,
                                 cudaStream_t stream) {
  static std::atomic<int> cap[kMaxDevices];              // 0: not initialised; CTAs (CL = 1) / clusters (CL = 2) to launch at most
  const int dev = current_device();
  if (dev < 0) return cudaErrorInvalidDevice;
  auto kern = tap_gemm_kernel<MODE, NRES, STAGED, LNF, CL>;
  int capacity = cap[dev].load(std::memory_order_acquire);
  if (capacity == 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    const int sms = device_sm_count();
    if (sms <= 0) return cudaErrorInvalidDevice;
    capacity = sms;
    if (CL >= 2) {
      cudaLaunchConfig_t q;
      memset(&q, 0, sizeof(q));
      q.gridDim = dim3(2 * (sms / 2));
      q.blockDim = dim3(kGemmThreads);
      q.dynamicSmemBytes = 227 * 1024;
      cudaLaunchAttribute a;
      a.id = cudaLaunchAttributeClusterDimension;
      a.val.clusterDim.x = 2;
      a.val.clusterDim.y = 1;
      a.val.clusterDim.z = 1;
      q.attrs = &a;
      q.numAttrs = 1;
      int n = 0;
      e = cudaOccupancyMaxActiveClusters(&n, kern, &q);
      if (e != cudaSuccess) return e;
      capacity = n < sms / 2 ? n : sms / 2;
      if (capacity < 1) return cudaErrorLaunchOutOfResources;
    }
    cap[dev].store(capacity, std::memory_order_release);
  }
  const int units = work_items < capacity ? work_items : capacity;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(CL == 1 ? units : 2 * units);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int nattr = 0;
  if (CL >= 2) {
    attr[nattr].id = cudaLaunchAttributeClusterDimension;
    attr[nattr].val.clusterDim.x = 2;
    attr[nattr].val.clusterDim.y = 1;
    attr[nattr].val.clusterDim.z = 1;
    ++nattr;
  }
  if (pdl_enabled(1)) {
    attr[nattr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[nattr].val.programmaticStreamSerializationAllowed = 1;
    ++nattr;
  }
  cfg.attrs = attr;
  cfg.numAttrs = nattr;
  void* args[] = {const_cast<CUtensorMap*>(&tm[0]), const_cast<CUtensorMap*>(&tm[1]), const_cast<CUtensorMap*>(&tm[2]),
                  const_cast<CUtensorMap*>(&tm[3]), const_cast<GemmKParams*>(&p)};
  return cudaLaunchKernelExC(&cfg, reinterpret_cast<const void*>(kern), args);
}
template <int MODE, int NRES, bool LNF = false>
static cudaError_t launch_gemm(const CUtensorMap* tm, const GemmKParams& p, bool staged, int cl, int smem_bytes,
                               cudaStream_t stream) {
  if (cl == 3)
    return staged ? launch_gemm_t<MODE, NRES, true, LNF, 3>(tm, p, p.total_pairs, smem_bytes, stream)
                  : launch_gemm_t<MODE, NRES, false, LNF, 3>(tm, p, p.total_pairs, smem_bytes, stream);
  if (cl == 2)
    return staged ? launch_gemm_t<MODE, NRES, true, LNF, 2>(tm, p, p.total_pairs, smem_bytes, stream)
                  : launch_gemm_t<MODE, NRES, false, LNF, 2>(tm, p, p.total_pairs, smem_bytes, stream);
  return staged ? launch_gemm_t<MODE, NRES, true, LNF, 1>(tm, p, p.total_tiles, smem_bytes, stream)
                : launch_gemm_t<MODE, NRES, false, LNF, 1>(tm, p, p.total_tiles, smem_bytes, stream);
}

// Output-side tensor map of the staged epilogue: the [d4][d3][d2][d1][cols] view behind `base` with box
// (cb columns, the 32-row quarter of the tile box), TMA swizzle chosen by the row width of a staging block.
static bool make_out_map(PFN_encodeTiled encode, CUtensorMap* m, const void* base, int cols, const int* odim,
                         const long long* str, const int* sub, int cb) {
  cuuint64_t dims[5] = {(cuuint64_t)cols, (cuuint64_t)odim[0], (cuuint64_t)odim[1], (cuuint64_t)odim[2], (cuuint64_t)odim[3]};
  cuuint64_t strides[4] = {(cuuint64_t)str[0] * 2, (cuuint64_t)str[1] * 2, (cuuint64_t)str[2] * 2, (cuuint64_t)str[3] * 2};
  cuuint32_t box[5] = {(cuuint32_t)cb, (cuuint32_t)sub[0], (cuuint32_t)sub[1], (cuuint32_t)sub[2], (cuuint32_t)sub[3]};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUtensorMapSwizzle sw = cb == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : cb == 32 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : cb == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  return encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int gemm_impl(const ccedit_gemm_desc* d, cudaStream_t stream) {
  CCEDIT_CHECK_ARG(d != nullptr, "ccedit_gemm: null descriptor");
  CCEDIT_CHECK_ARG(d->a && d->w && d->out, "ccedit_gemm: null a/w/out pointer");
  CCEDIT_CHECK_ARG(d->ntaps >= 1 && d->ntaps <= CCEDIT_MAX_TAPS, "ccedit_gemm: ntaps=%d out of range", d->ntaps);
  CCEDIT_CHECK_ARG(d->bn >= 16 && d->bn <= 256 && d->bn % 16 == 0, "ccedit_gemm: bn=%d must be a multiple of 16 in [16,256]", d->bn);
  CCEDIT_CHECK_ARG(d->n > 0 && d->n % d->bn == 0, "ccedit_gemm: n=%d not a multiple of bn=%d", d->n, d->bn);
  CCEDIT_CHECK_ARG(d->kpad > 0 && d->kpad % kBlockK == 0 && d->kpad >= d->a_dims[0],
                   "ccedit_gemm: kpad=%d must be a multiple of 64 and >= C=%d", d->kpad, d->a_dims[0]);
  CCEDIT_CHECK_ARG(d->a_dims[0] % 8 == 0, "ccedit_gemm: C=%d must be a multiple of 8", d->a_dims[0]);
  const bool geglu = (d->flags & CCEDIT_GEMM_GEGLU) != 0;
  CCEDIT_CHECK_ARG(!geglu || d->bn % 32 == 0, "ccedit_gemm: GEGLU needs bn %% 32 == 0 (bn=%d)", d->bn);
  CCEDIT_CHECK_ARG(!(geglu && d->rowbias), "ccedit_gemm: GEGLU and rowbias cannot be combined");
  CCEDIT_CHECK_ARG(!(geglu && (d->flags & CCEDIT_GEMM_SILU)), "ccedit_gemm: GEGLU and SiLU cannot be combined");
  CCEDIT_CHECK_ARG(!(geglu && (d->res1 || d->res2)), "ccedit_gemm: GEGLU and residuals cannot be combined");
  long long boxprod = 1;
  for (int i = 0; i < 4; ++i) {
    CCEDIT_CHECK_ARG(d->box[i] >= 1 && d->box[i] <= 256, "ccedit_gemm: box[%d]=%d", i, d->box[i]);
    CCEDIT_CHECK_ARG(d->out_dims[i] >= 1 && d->a_dims[i + 1] >= 1, "ccedit_gemm: empty dim %d", i);
    CCEDIT_CHECK_ARG((d->a_strides[i] * 2) % 16 == 0, "ccedit_gemm: a_strides[%d]=%lld not 16-byte aligned", i,
                     (long long)d->a_strides[i]);
    boxprod *= d->box[i];
  }
  CCEDIT_CHECK_ARG(boxprod == kBlockM, "ccedit_gemm: box product %lld != 128", boxprod);
  CCEDIT_CHECK_ARG((reinterpret_cast<uintptr_t>(d->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->w) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(d->out) & 15) == 0,
                   "ccedit_gemm: a/w/out must be 16-byte aligned");

  PFN_encodeTiled encode = get_encode_fn();
  if (!encode) {
    set_last_error("ccedit_gemm: cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return CCEDIT_ERR_CUDA;
  }

  // Clusters of two CTAs on M-adjacent tiles of one N tile, for the GEMMs whose K loop is long enough that the tile is
  // paced by operand delivery (convolutions, temporal k3, K >= 640 linears: k-loops of >= 10 blocks).
  // CCEDIT_GEMM_CLUSTER: 0 = never; 1 = TMA multicast of the W tile (CL = 2) under that rule, 2 = wherever possible;
  // 3 = CTA-pair MMA (CL = 3, tcgen05 cta_group::2, M = 256; default) under that rule, 4 = wherever possible.
  static const int cluster_mode = [] { const char* e = getenv("CCEDIT_GEMM_CLUSTER"); return e ? atoi(e) : 3; }();
  long long m_tiles_all = 1;
  for (int i = 0; i < 4; ++i) m_tiles_all *= (d->out_dims[i] + d->box[i] - 1) / d->box[i];
  const int kblocks_all = d->ntaps * (d->kpad / kBlockK);
  const bool cl_rule = cluster_mode == 2 || cluster_mode == 4 || ((cluster_mode == 1 || cluster_mode == 3) && kblocks_all >= 10);
  const int cl = (cl_rule && m_tiles_all >= 2) ? (cluster_mode >= 3 ? 3 : 2) : 1;
  const int wsplit = cl >= 2 ? 2 : 1;                    // W rows per TMA box = bn / wsplit

  CUtensorMap tm[4];
  CUtensorMap& tmA = tm[0];
  CUtensorMap& tmB = tm[1];
  {
    cuuint64_t dims[5] = {(cuuint64_t)d->a_dims[0], (cuuint64_t)d->a_dims[1], (cuuint64_t)d->a_dims[2],
                          (cuuint64_t)d->a_dims[3], (cuuint64_t)d->a_dims[4]};
    cuuint64_t strides[4] = {(cuuint64_t)d->a_strides[0] * 2, (cuuint64_t)d->a_strides[1] * 2,
                             (cuuint64_t)d->a_strides[2] * 2, (cuuint64_t)d->a_strides[3] * 2};
    cuuint32_t box[5] = {(cuuint32_t)kBlockK, (cuuint32_t)d->box[0], (cuuint32_t)d->box[1], (cuuint32_t)d->box[2],
                         (cuuint32_t)d->box[3]};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(d->a), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_last_error("ccedit_gemm: cuTensorMapEncodeTiled(A) failed with CUresult %d (dims %d,%d,%d,%d,%d)", (int)r,
                     d->a_dims[0], d->a_dims[1], d->a_dims[2], d->a_dims[3], d->a_dims[4]);
      return CCEDIT_ERR_CUDA;
    }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)d->ntaps * d->kpad, (cuuint64_t)d->n};
    cuuint64_t strides[1] = {(cuuint64_t)d->ntaps * d->kpad * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)(d->bn / wsplit)};   // clusters: each CTA fetches half of the W tile
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(d->w), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_last_error("ccedit_gemm: cuTensorMapEncodeTiled(W) failed with CUresult %d", (int)r);
      return CCEDIT_ERR_CUDA;
    }
  }

  GemmKParams p;
  memset(&p, 0, sizeof(p));
  long long m_tiles = 1;
  for (int i = 0; i < 4; ++i) {
    p.box[i] = d->box[i];
    p.odim[i] = d->out_dims[i];
    p.tiles[i] = (d->out_dims[i] + d->box[i] - 1) / d->box[i];
    m_tiles *= p.tiles[i];
    p.ostr[i] = d->out_strides[i];
    p.r1[i] = d->res1 ? d->res1_strides[i] : 0;
    p.r2[i] = d->res2 ? d->res2_strides[i] : 0;
  }
  p.n_tiles = d->n / d->bn;
  CCEDIT_CHECK_ARG(m_tiles * p.n_tiles < (1ll << 31), "ccedit_gemm: too many tiles");
  p.total_tiles = static_cast<int>(m_tiles * p.n_tiles);
  p.total_pairs = static_cast<int>(((m_tiles + 1) / 2) * p.n_tiles);
  p.ntaps = d->ntaps;
  p.kchunks = d->kpad / kBlockK;
  p.bn = d->bn;
  for (int t = 0; t < d->ntaps; ++t)
    for (int i = 0; i < 4; ++i) p.taps[t][i] = d->taps[t][i];
  p.out = static_cast<__half*>(d->out);
  p.bias = d->bias;
  p.rowbias = d->rowbias;
  p.rb_dim = d->rowbias ? d->rb_dim : -1;
  p.rb_div = d->rb_div > 0 ? d->rb_div : 1;
  p.n_out_total = d->rb_ld > 0 ? d->rb_ld : (geglu ? d->n / 2 : d->n);
  p.res1 = static_cast<const __half*>(d->res1);
  p.res2 = static_cast<const __half*>(d->res2);
  p.flags = d->flags;
  static const int dev_flags = [] { const char* e = getenv("CCEDIT_GEMM_DEV"); return e ? atoi(e) << 8 : 0; }();
  p.flags |= dev_flags;                                                           // developer experiments only
  p.idesc = umma_idesc_f16(cl == 3 ? 2 * kBlockM : kBlockM, d->bn);   // pair MMA: one instruction spans both CTAs, M = 256
  p.trace = g_trace_buf;
  const int stage_bytes = kABytes + (cl == 3 ? d->bn / 2 : d->bn) * kBlockK * 2;
  const int kblocks = d->ntaps * p.kchunks;
  const int nres = (d->res1 ? 1 : 0) + (d->res2 ? 1 : 0);
  if (d->res2 && !d->res1) {
    p.res1 = p.res2;
    for (int i = 0; i < 4; ++i) p.r1[i] = p.r2[i];
    p.res2 = nullptr;
  }
  const int fixed_bytes = 1024 /* alignment slack */ + 512 /* barriers */ + kEpiWarps * kBiasFloats * 4 /* bias rows */;
  const int budget = 227 * 1024 - fixed_bytes;
  // ---- staged (TMA) epilogue: for GEMMs whose K loop is too short to hide the row-per-thread global accesses ----
  bool staged = false;
  {
    const int ncols_out = geglu ? d->bn / 2 : d->bn;
    // developer switch: 0 = always direct, 1 = staged wherever legal (read once per process)
    static const int force_epi = [] { const char* e = getenv("CCEDIT_GEMM_EPI"); return e ? (atoi(e) != 0 ? 1 : 0) : -1; }();
    bool ok = ncols_out % 32 == 0 && force_epi != 0;
    for (int i = 0; i < 4 && ok; ++i) {
      ok = d->out_strides[i] % 8 == 0 && (!p.res1 || p.r1[i] % 8 == 0);
      // a size-1 grid axis may carry any stride; TMA still wants it 16-byte aligned
    }
    ok = ok && (!p.res1 || (reinterpret_cast<uintptr_t>(p.res1) & 15) == 0);
    const int W = ncols_out / 2;
    const int cb = W % 64 == 0 ? 64 : (W == 96 ? 32 : W);    // TMA-swizzled 128 / 64-byte staging rows where rows would collide
    ok = ok && (cb <= 128) && (W % cb == 0);
    if (ok) {
      const int tbuf = kBlockM * ncols_out * 2;
      // the K loop hides a direct epilogue of ~5-8k clocks once a tile has >= ~5k clocks of MMAs (measured: K = 1280, BN = 160 is faster direct)
      const bool wanted = force_epi == 1 || kblocks * d->bn <= 15 * 160;
      int nbuf = nres >= 1 ? 2 : 1;
      int st = (budget - nbuf * tbuf) / stage_bytes;
      if (st < 3 && nbuf == 2) {
        nbuf = 1;
        st = (budget - tbuf) / stage_bytes;
      }
      if (wanted && st >= 3) {
        staged = true;
        p.wcols = W;
        p.cb = cb;
        p.nbuf = nbuf;
        p.tbuf_bytes = tbuf;
        p.swz_mask = cb == 64 ? 7u : cb == 32 ? 3u : cb == 16 ? 1u : 0u;
        int sub[4], rem = 32;
        for (int i = 0; i < 4; ++i) {
          sub[i] = d->box[i] < rem ? d->box[i] : rem;
          rem /= sub[i];
        }
        const int cols = geglu ? d->n / 2 : d->n;
        if (!make_out_map(encode, &tm[2], d->out, cols, p.odim, p.ostr, sub, cb) ||
            (p.res1 && !make_out_map(encode, &tm[3], p.res1, cols, p.odim, p.r1, sub, cb))) {
          set_last_error("ccedit_gemm: cuTensorMapEncodeTiled(out/res) failed");
          return CCEDIT_ERR_CUDA;
        }
        if (!p.res1) tm[3] = tm[2];
      }
    }
  }
  if (!staged) {
    tm[2] = tm[0];
    tm[3] = tm[0];
  }
  // direct epilogue: 192 KB of stages (5 x 36 KB at BN = 160); the CTA-pair kernel's smaller stages (26 KB) take the whole
  // budget: 8 stages = 2 560 tensor-clocks of operands in flight
  int stages = (staged ? budget - p.nbuf * p.tbuf_bytes : ((cl != 3 && 192 * 1024 < budget) ? 192 * 1024 : budget)) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  p.stages = stages;
  const int smem_bytes = stages * stage_bytes + (staged ? p.nbuf * p.tbuf_bytes : 0) + fixed_bytes;

  // (A grid rounded down to a multiple of n_tiles would keep every CTA on one N tile and save the per-tile staging for
  //  N = 960 / 2560 as well; measured: the SMs given up cost more than the staging - GEGLU 367 -> 391 us.)
  const int grid = cl;                                   // launch_gemm sizes the grid: min(work items, SMs or cluster capacity)
  if (d->stats_out) {
    CCEDIT_CHECK_ARG(!geglu && !(d->flags & CCEDIT_GEMM_SILU) && d->out_dims[1] == 1 && d->out_dims[2] == 1 && d->out_dims[3] == 1,
                     "ccedit_gemm: stats_out needs a plain 2-D [M, C] GEMM");
    CCEDIT_CHECK_ARG((reinterpret_cast<uintptr_t>(d->stats_out) & 7) == 0, "ccedit_gemm: stats_out must be 8-byte aligned");
    p.stats_out = reinterpret_cast<float2*>(d->stats_out);
  }
  cudaError_t err;
  const bool lnf = d->rowstats != nullptr;
  CCEDIT_CHECK_ARG((d->rowstats != nullptr) == (d->colsum != nullptr), "ccedit_gemm: rowstats and colsum go together");
  if (lnf) {
    CCEDIT_CHECK_ARG(nres == 0 && !d->rowbias && !(d->flags & CCEDIT_GEMM_SILU),
                     "ccedit_gemm: the LayerNorm fold cannot be combined with residuals / rowbias / SiLU");
    CCEDIT_CHECK_ARG(d->out_dims[1] == 1 && d->out_dims[2] == 1 && d->out_dims[3] == 1,
                     "ccedit_gemm: the LayerNorm fold needs a 2-D [M, C] problem (out_dims[1..3] == 1)");
    CCEDIT_CHECK_ARG((reinterpret_cast<uintptr_t>(d->rowstats) & 7) == 0, "ccedit_gemm: rowstats must be 8-byte aligned");
    p.rowstats = reinterpret_cast<const float2*>(d->rowstats);
    p.colsum = d->colsum;
    CCEDIT_CHECK_ARG(d->rowstats_slots >= 0 && d->rowstats_slots <= 64, "ccedit_gemm: rowstats_slots=%d", d->rowstats_slots);
    p.rs_slots = d->rowstats_slots;
    p.rs_invc = 1.f / static_cast<float>(d->a_dims[0]);
    p.rs_eps = d->ln_eps;
    err = geglu ? launch_gemm<kModeGeglu, 0, true>(tm, p, staged, grid, smem_bytes, stream)
                : launch_gemm<kModePlain, 0, true>(tm, p, staged, grid, smem_bytes, stream);
  } else if (geglu) err = launch_gemm<kModeGeglu, 0>(tm, p, staged, grid, smem_bytes, stream);
  else if (d->flags & CCEDIT_GEMM_SILU) {
    CCEDIT_CHECK_ARG(nres == 0, "ccedit_gemm: SiLU and residuals cannot be combined");
    err = launch_gemm<kModeSilu, 0>(tm, p, staged, grid, smem_bytes, stream);
  } else if (nres == 0) err = launch_gemm<kModePlain, 0>(tm, p, staged, grid, smem_bytes, stream);
  else if (nres == 1) err = launch_gemm<kModePlain, 1>(tm, p, staged, grid, smem_bytes, stream);
  else err = launch_gemm<kModePlain, 2>(tm, p, staged, grid, smem_bytes, stream);
  if (err != cudaSuccess) {
    set_last_error("ccedit_gemm: launch failed: %s", cudaGetErrorString(err));
    return CCEDIT_ERR_CUDA;
  }
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  CCEDIT_CUDA_LAUNCH_CHECK("ccedit_gemm");
  return CCEDIT_OK;
}

}  // namespace ccedit

extern "C" int ccedit_gemm_trace(int64_t* device_buf) {
  ccedit::g_trace_buf = reinterpret_cast<long long*>(device_buf);
  return CCEDIT_OK;
}

extern "C" int ccedit_gemm(const ccedit_gemm_desc* d, void* stream) {
  return ccedit::gemm_impl(d, static_cast<cudaStream_t>(stream));
}

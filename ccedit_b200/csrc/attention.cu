// Attention kernels.
//  * flash_attn_kernel: softmax(q k^T * scale) v with an online softmax over 64-key tiles staged in shared memory
//    (cp.async double buffering), warp-level tensor-core MMAs, fp32 accumulation.  Replaces the
//    F.scaled_dot_product_attention call at attention.py:444-448 for spatial self-attention (keys = the frame's own
//    tokens), text cross-attention (77 keys shared by all frames of a batch entry) and the cross-frame "center_self"
//    block (two key segments: centre-frame tokens then own tokens, attention.py:1323-1336).
//  * temporal_attn_kernel: one warp per (pixel, head), attends over the T frames of that pixel without ever
//    materialising the "(b h w) t c" transposition the reference performs (attention.py:1172-1194).
#include "common.cuh"
#include "../../include/ccedit_b200.h"

#include <atomic>
#include <cstdlib>

namespace ccedit {
extern std::atomic<long long> g_launch_count;
int attention_tc(const ccedit_attn_desc* a, cudaStream_t st);   // attention_tc.cu: tcgen05 path, -1 if not eligible

constexpr int kFaBM = 128;     // queries per CTA (8 warps x 16 rows)
constexpr int kFaBN = 64;      // keys per tile
constexpr int kFaThreads = 256;

struct FaParams {
  const __half* q;
  long long ldq, q_fs;
  __half* o;
  long long ldo, o_fs;
  int nseg;
  const __half* k[2];
  const __half* v[2];
  long long ldk[2], ldv[2], kv_fs[2];
  int lkv[2], kv_div[2], kv_mul[2], kv_add[2];
  int lq, d;
  float scale_log2;
};

template <int DP>
__global__ void __launch_bounds__(kFaThreads) flash_attn_kernel(const __grid_constant__ FaParams p) {
  constexpr int DS = DP + 8;        // smem row stride in halves (odd multiple of 16 B => conflict-free ldmatrix)
  constexpr int KSTEPS = DP / 16;   // k-steps of QK^T
  constexpr int NT_O = DP / 8;      // n-tiles of the output
  extern __shared__ __align__(16) uint8_t fa_smem[];
  __half* sQ = reinterpret_cast<__half*>(fa_smem);          // [128][DS]
  __half* sK = sQ + kFaBM * DS;                               // [2][64][DS]
  __half* sV = sK + 2 * kFaBN * DS;                           // [2][64][DS]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * kFaBM, head = blockIdx.y, f = blockIdx.z;
  const int d = p.d;
  const int chunks = d >> 3;        // 16-byte chunks per row actually present in global memory
  constexpr int CH = DP / 8;        // chunks per padded row

  // zero the padding chunks once (cp.async never touches them)
  if (chunks < CH) {
    const int padc = CH - chunks;
    for (int i = tid; i < (kFaBM + 4 * kFaBN) * padc; i += kFaThreads) {
      const int r = i / padc, c = chunks + i % padc;
      *reinterpret_cast<uint4*>(sQ + r * DS + c * 8) = make_uint4(0, 0, 0, 0);  // sQ,sK,sV are contiguous
    }
  }

  const __half* qbase = p.q + static_cast<long long>(f) * p.q_fs + static_cast<long long>(head) * d;
  for (int i = tid; i < kFaBM * chunks; i += kFaThreads) {
    const int r = i / chunks, c = i % chunks;
    const bool ok = (q0 + r) < p.lq;
    const __half* src = qbase + static_cast<long long>(ok ? q0 + r : 0) * p.ldq + c * 8;
    cp_async_16(smem_u32(sQ + r * DS + c * 8), src, ok);
  }

  // tile bookkeeping over the (up to two) key segments
  const int ntile0 = (p.lkv[0] + kFaBN - 1) / kFaBN;
  const int ntile1 = p.nseg > 1 ? (p.lkv[1] + kFaBN - 1) / kFaBN : 0;
  const int ntiles = ntile0 + ntile1;

  auto load_kv = [&](int it, int buf) {
    const int seg = it < ntile0 ? 0 : 1;
    const int k0 = (seg == 0 ? it : it - ntile0) * kFaBN;
    const long long kvf = static_cast<long long>(f / p.kv_div[seg]) * p.kv_mul[seg] + p.kv_add[seg];
    const __half* kb = p.k[seg] + kvf * p.kv_fs[seg] + static_cast<long long>(head) * d;
    const __half* vb = p.v[seg] + kvf * p.kv_fs[seg] + static_cast<long long>(head) * d;
    __half* dK = sK + buf * kFaBN * DS;
    __half* dV = sV + buf * kFaBN * DS;
    for (int i = tid; i < kFaBN * chunks; i += kFaThreads) {
      const int r = i / chunks, c = i % chunks;
      const bool ok = (k0 + r) < p.lkv[seg];
      const long long rr = ok ? k0 + r : 0;
      cp_async_16(smem_u32(dK + r * DS + c * 8), kb + rr * p.ldk[seg] + c * 8, ok);
      cp_async_16(smem_u32(dV + r * DS + c * 8), vb + rr * p.ldv[seg] + c * 8, ok);
    }
  };

  load_kv(0, 0);
  cp_async_commit();

  float o_acc[NT_O][4];
#pragma unroll
  for (int i = 0; i < NT_O; ++i) o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  uint32_t qf[KSTEPS][4];

  for (int it = 0; it < ntiles; ++it) {
    const int buf = it & 1;
    if (it + 1 < ntiles) {
      load_kv(it + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (it == 0) {
      // Q fragments: warp's 16 rows, A-operand layout via ldmatrix.x4
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) {
        const int r = warp * 16 + (lane & 15);
        const int c = ks * 16 + (lane >> 4) * 8;
        ldmatrix_x4(qf[ks], smem_u32(sQ + r * DS + c));
      }
    }
    const __half* tK = sK + buf * kFaBN * DS;
    const __half* tV = sV + buf * kFaBN * DS;

    // ---- S = Q K^T (16 x 64 per warp) ----
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {  // pairs of 8-key n-tiles
        uint32_t kf[4];
        const int r = np * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int c = ks * 16 + ((lane >> 3) & 1) * 8;
        ldmatrix_x4(kf, smem_u32(tK + r * DS + c));
        const uint32_t b0[2] = {kf[0], kf[1]}, b1[2] = {kf[2], kf[3]};
        mma_m16n8k16(s[2 * np], qf[ks], b0);
        mma_m16n8k16(s[2 * np + 1], qf[ks], b1);
      }
    }
    // ---- mask keys beyond the segment end, scale ----
    const int seg = it < ntile0 ? 0 : 1;
    const int k0 = (seg == 0 ? it : it - ntile0) * kFaBN;
    const int kvalid = p.lkv[seg] - k0;  // number of valid keys in this tile (may exceed 64)
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int col = nt * 8 + (lane & 3) * 2;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool ok = (col + (e & 1)) < kvalid;
        const float val = ok ? s[nt][e] * p.scale_log2 : -INFINITY;
        s[nt][e] = val;
        mx[e >> 1] = fmaxf(mx[e >> 1], val);
      }
    }
    float alpha[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
      const float m_new = fmaxf(m_run[h], mx[h]);  // finite: every tile has at least one valid key
      alpha[h] = exp2f(m_run[h] - m_new);
      m_run[h] = m_new;
    }
    float rs[2] = {0.f, 0.f};
    uint32_t pf[4][4];  // P as A fragments for 4 k-steps of 16 keys
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = exp2f(s[nt][0] - m_run[0]), p1 = exp2f(s[nt][1] - m_run[0]);
      const float p2 = exp2f(s[nt][2] - m_run[1]), p3 = exp2f(s[nt][3] - m_run[1]);
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      __half2 h01 = __floats2half2_rn(p0, p1), h23 = __floats2half2_rn(p2, p3);
      pf[nt >> 1][(nt & 1) * 2 + 0] = *reinterpret_cast<uint32_t*>(&h01);
      pf[nt >> 1][(nt & 1) * 2 + 1] = *reinterpret_cast<uint32_t*>(&h23);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) l_run[h] = l_run[h] * alpha[h] + rs[h];
#pragma unroll
    for (int i = 0; i < NT_O; ++i) {
      o_acc[i][0] *= alpha[0];
      o_acc[i][1] *= alpha[0];
      o_acc[i][2] *= alpha[1];
      o_acc[i][3] *= alpha[1];
    }
    // ---- O += P V ----
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {       // 16 keys per step
#pragma unroll
      for (int np = 0; np < NT_O / 2; ++np) {  // pairs of 8-wide d tiles
        uint32_t vf[4];
        const int r = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int c = np * 16 + (lane >> 4) * 8;
        ldmatrix_x4_trans(vf, smem_u32(tV + r * DS + c));
        const uint32_t b0[2] = {vf[0], vf[1]}, b1[2] = {vf[2], vf[3]};
        mma_m16n8k16(o_acc[2 * np], pf[kk], b0);
        mma_m16n8k16(o_acc[2 * np + 1], pf[kk], b1);
      }
    }
    __syncthreads();  // all warps done with this buffer before it is refilled
  }

  // ---- finalise: O / l, stage through this warp's own Q rows, coalesced 16-byte stores ----
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 1);
    l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 2);
  }
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
  __half* sO = sQ + warp * 16 * DS;
  __syncwarp();
#pragma unroll
  for (int nt = 0; nt < NT_O; ++nt) {
    const int col = nt * 8 + (lane & 3) * 2;
    const int r = lane >> 2;
    *reinterpret_cast<__half2*>(sO + r * DS + col) = __floats2half2_rn(o_acc[nt][0] * inv0, o_acc[nt][1] * inv0);
    *reinterpret_cast<__half2*>(sO + (r + 8) * DS + col) = __floats2half2_rn(o_acc[nt][2] * inv1, o_acc[nt][3] * inv1);
  }
  __syncwarp();
  __half* obase = p.o + static_cast<long long>(f) * p.o_fs + static_cast<long long>(head) * d;
  for (int i = lane; i < 16 * chunks; i += 32) {
    const int r = i / chunks, c = i % chunks;
    const int qr = q0 + warp * 16 + r;
    if (qr < p.lq)
      *reinterpret_cast<uint4*>(obase + static_cast<long long>(qr) * p.ldo + c * 8) =
          *reinterpret_cast<const uint4*>(sO + r * DS + c * 8);
  }
}

template <int DP>
static int launch_fa(const FaParams& p, int frames, int heads, cudaStream_t st) {
  constexpr int DS = DP + 8;
  const int smem = (kFaBM + 4 * kFaBN) * DS * 2;
  static std::atomic<bool> attr_done[kMaxDevices];        // function attributes belong to a device
  const int dev = current_device();
  CCEDIT_CHECK_ARG(dev >= 0, "ccedit_attention: no current CUDA device");
  if (!attr_done[dev].load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(flash_attn_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_last_error("ccedit_attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return CCEDIT_ERR_CUDA;
    }
    attr_done[dev].store(true, std::memory_order_release);
  }
  dim3 grid((p.lq + kFaBM - 1) / kFaBM, heads, frames);
  flash_attn_kernel<DP><<<grid, kFaThreads, smem, st>>>(p);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  CCEDIT_CUDA_LAUNCH_CHECK("ccedit_attention");
  return CCEDIT_OK;
}

__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// first MMA of an accumulator chain: C = 0 comes from the zero register instead of 4 zeroed registers per n-tile
__device__ __forceinline__ void mma_m16n8k16_first(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
      : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "f"(0.f));
}

// ---------------------------------------------------------------------------------------------------------------
// temporal attention: one CTA per pixel (x head group), one warp per head.
//   phase 1: the CTA streams the pixel's q, k, v rows ([T][C], C = heads*d contiguous) into shared memory with
//            16-byte cp.async (fully coalesced rows, everything in flight at once; rows T..TK-1 are zero-filled);
//   phase 2: warp h computes softmax(Q_h K_h^T) V_h for its head with warp-level tensor-core MMAs (m16n8k16): the T x T
//            problem is tiny (T <= 64), so the scalar version of this kernel was instruction-bound (7 TFLOP/s, 830 GB/s);
//            with ~150 warp instructions per (pixel, head) it is bound by the HBM stream of q, k, v, o again.
// Row stride in smem is C+8 halves (odd multiple of 16 B): conflict-free ldmatrix.
// ---------------------------------------------------------------------------------------------------------------
template <int DP, int TK>   // DP: head dim rounded up (multiple of 16); TK: keys rounded up to 32 or 64
__global__ void __launch_bounds__(256) temporal_attn_kernel(const __half* __restrict__ q, long long ldq,
                                                            const __half* __restrict__ k, long long ldk,
                                                            const __half* __restrict__ v, long long ldv,
                                                            __half* __restrict__ o, long long ldo, int T, int HW, int heads,
                                                            int d, float scale_log2) {
  // `heads` = heads handled by this CTA (a head group when T*C does not fit in shared memory); blockIdx.y = group
  constexpr int KSTEPS = DP / 16, NT_O = DP / 8, NTK = TK / 8;
  extern __shared__ __align__(16) uint8_t ta_smem[];
  const int C = heads * d;
  const long long col0 = static_cast<long long>(blockIdx.y) * C;
  q += col0;
  k += col0;
  v += col0;
  o += col0;
  const int RS = C + 8;                                   // smem row stride (halves)
  // Only the T real frames are staged ([T][RS] per tensor): fragment rows T..TK-1 alias frame T-1 - padded keys are
  // masked to -inf before the softmax (so their P is exactly 0 and any finite V row will do) and padded query rows are
  // never stored.  At T = 17 this halves the shared memory of a CTA (3 x 17 instead of 3 x 32 rows) and doubles the CTAs
  // per SM, which is what hides the HBM latency of this kernel.
  __half* sq = reinterpret_cast<__half*>(ta_smem);        // [T][RS]
  __half* sk = sq + T * RS;
  __half* sv = sk + T * RS;
  const int tl = T - 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long pix = blockIdx.x;                       // b*HW + hw
  const int b = static_cast<int>(pix / HW), hw = static_cast<int>(pix % HW);
  const long long row0 = (static_cast<long long>(b) * T) * HW + hw;  // row of frame 0; frame t adds t*HW
  const int cpr = C >> 3;                                 // 16-byte chunks per row
  for (int i = threadIdx.x; i < T * cpr; i += blockDim.x) {
    const int t = i / cpr, c = i - t * cpr;
    const long long row = row0 + static_cast<long long>(t) * HW;
    cp_async_16(smem_u32(sq + t * RS + c * 8), q + row * ldq + c * 8, true);
    cp_async_16(smem_u32(sk + t * RS + c * 8), k + row * ldk + c * 8, true);
    cp_async_16(smem_u32(sv + t * RS + c * 8), v + row * ldv + c * 8, true);
  }
  // the 8 padding halves at the end of every row feed the last k-step of the last head: keep them finite
  for (int i = threadIdx.x; i < 3 * T; i += blockDim.x)
    *reinterpret_cast<uint4*>(sq + i * RS + C) = make_uint4(0, 0, 0, 0);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  const bool mask_tail = (d & 15) != 0;                   // d % 16 == 8: the last k-step covers 8 foreign channels
  for (int head = warp; head < heads; head += (blockDim.x >> 5)) {
    const __half* hk = sk + head * d;
    const __half* hv = sv + head * d;
    for (int m0 = 0; m0 < T; m0 += 16) {
      // ---- S = Q K^T for 16 query frames x TK keys ----
      float sacc[NTK][4];
#pragma unroll
      for (int i = 0; i < NTK; ++i) sacc[i][0] = sacc[i][1] = sacc[i][2] = sacc[i][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) {
        if (ks * 16 >= d) break;                                   // DP may exceed d by whole k-steps (e.g. d = 96)
        uint32_t qf[4];
        // d % 16 == 8: the upper half of the last k-step lies past d - in the next head's columns, which that head's warp
        // may be overwriting with its output rows.  Those lanes read the row's zero padding instead (no race, no select).
        const __half* qrow_p = sq + min(m0 + (lane & 15), tl) * RS;
        const bool past = mask_tail && d - ks * 16 == 8 && lane >= 16;
        ldmatrix_x4(qf, smem_u32(past ? qrow_p + C : qrow_p + head * d + ks * 16 + (lane >> 4) * 8));
#pragma unroll
        for (int np = 0; np < NTK / 2; ++np) {
          uint32_t kf[4];
          ldmatrix_x4(kf, smem_u32(hk + min(np * 16 + (lane & 7) + ((lane >> 4) << 3), tl) * RS + ks * 16 + ((lane >> 3) & 1) * 8));
          const uint32_t b0[2] = {kf[0], kf[1]}, b1[2] = {kf[2], kf[3]};
          mma_m16n8k16(sacc[2 * np], qf, b0);
          mma_m16n8k16(sacc[2 * np + 1], qf, b1);
        }
      }
      // ---- softmax over the T valid keys (rows g and g+8 of the fragment) ----
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < NTK; ++nt) {
        const int col = nt * 8 + (lane & 3) * 2;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float val = (col + (e & 1)) < T ? sacc[nt][e] * scale_log2 : -INFINITY;
          sacc[nt][e] = val;
          mx[e >> 1] = fmaxf(mx[e >> 1], val);
        }
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
        mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
      }
      float rs[2] = {0.f, 0.f};
      uint32_t pf[NTK / 2][4];
#pragma unroll
      for (int nt = 0; nt < NTK; ++nt) {
        const float p0 = exp2f(sacc[nt][0] - mx[0]), p1 = exp2f(sacc[nt][1] - mx[0]);
        const float p2 = exp2f(sacc[nt][2] - mx[1]), p3 = exp2f(sacc[nt][3] - mx[1]);
        rs[0] += p0 + p1;
        rs[1] += p2 + p3;
        __half2 h01 = __floats2half2_rn(p0, p1), h23 = __floats2half2_rn(p2, p3);
        pf[nt >> 1][(nt & 1) * 2 + 0] = *reinterpret_cast<uint32_t*>(&h01);
        pf[nt >> 1][(nt & 1) * 2 + 1] = *reinterpret_cast<uint32_t*>(&h23);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        rs[h] += __shfl_xor_sync(0xffffffffu, rs[h], 1);
        rs[h] += __shfl_xor_sync(0xffffffffu, rs[h], 2);
      }
      // ---- O = P V ----
      float oacc[NT_O][4];
#pragma unroll
      for (int i = 0; i < NT_O; ++i) oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < NTK / 2; ++kk) {
#pragma unroll
        for (int np = 0; np < NT_O / 2; ++np) {
          if (np * 16 >= d) break;                                 // DP may exceed d by whole 16-column groups
          uint32_t vf[4];
          ldmatrix_x4_trans(vf, smem_u32(hv + min(kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, tl) * RS + np * 16 + (lane >> 4) * 8));
          const uint32_t b0[2] = {vf[0], vf[1]}, b1[2] = {vf[2], vf[3]};
          mma_m16n8k16(oacc[2 * np], pf[kk], b0);
          mma_m16n8k16(oacc[2 * np + 1], pf[kk], b1);
        }
      }
      const float inv0 = 1.f / rs[0], inv1 = 1.f / rs[1];
      const int r0 = m0 + (lane >> 2), r1 = r0 + 8;
      // the output rows of this head replace its query rows in shared memory (they are not needed any more: Q of block m0
      // only feeds S of block m0); the CTA writes whole [T][C] rows to global memory afterwards, 16 bytes per thread
      __half* o0 = sq + r0 * RS + head * d;
      __half* o1 = sq + r1 * RS + head * d;
      __syncwarp();
#pragma unroll
      for (int nt = 0; nt < NT_O; ++nt) {
        const int col = nt * 8 + (lane & 3) * 2;
        if (col < d) {
          if (r0 < T) *reinterpret_cast<__half2*>(o0 + col) = __floats2half2_rn(oacc[nt][0] * inv0, oacc[nt][1] * inv0);
          if (r1 < T) *reinterpret_cast<__half2*>(o1 + col) = __floats2half2_rn(oacc[nt][2] * inv1, oacc[nt][3] * inv1);
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * cpr; i += blockDim.x) {
    const int t = i / cpr, c = i - t * cpr;
    *reinterpret_cast<uint4*>(o + (row0 + static_cast<long long>(t) * HW) * ldo + c * 8) =
        *reinterpret_cast<const uint4*>(sq + t * RS + c * 8);
  }
}

template <int DP, int TK>
static int launch_ta(const __half* q, long long ldq, const __half* k, long long ldk, const __half* v, long long ldv, __half* o,
                     long long ldo, int B, int T, int HW, int heads, int d, float scale_log2, cudaStream_t st) {
  int hpb = heads;   // heads per CTA: all of them unless the pixel's q/k/v rows do not fit in shared memory
  auto smem_for = [&](int h) { return static_cast<size_t>(3) * T * (h * d + 8) * 2; };
  while (hpb > 1 && (smem_for(hpb) > 72 * 1024 || heads % hpb != 0)) --hpb;   // <= 72 KB: three CTAs per SM
  const size_t smem = smem_for(hpb);
  CCEDIT_CHECK_ARG(smem <= 227 * 1024, "ccedit_temporal_attention: T*d too large for shared memory (%zu bytes)", smem);
  static std::atomic<bool> attr_done[kMaxDevices];
  const int dev = current_device();
  CCEDIT_CHECK_ARG(dev >= 0, "ccedit_temporal_attention: no current CUDA device");
  if (!attr_done[dev].load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(temporal_attn_kernel<DP, TK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_last_error("ccedit_temporal_attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return CCEDIT_ERR_CUDA;
    }
    attr_done[dev].store(true, std::memory_order_release);
  }
  const int warps = hpb < 8 ? hpb : 8;
  const dim3 grid(static_cast<unsigned>(static_cast<long long>(B) * HW), heads / hpb);
  temporal_attn_kernel<DP, TK><<<grid, warps * 32, smem, st>>>(q, ldq, k, ldk, v, ldv, o, ldo, T, HW, hpb, d, scale_log2);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  CCEDIT_CUDA_LAUNCH_CHECK("ccedit_temporal_attention");
  return CCEDIT_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// The same operator for the shapes of the network (8 heads, d = 40 / 80 / 160, T = 9 / 17 / 33 keyframes) with everything
// a compile-time constant.  ncu of the generic kernel above at the top level (2 x 6144 pixels, T = 17, d = 40): 1 236 warp
// instructions per (pixel, head) for 48 HMMA, issue slots 74 % busy - not the HBM stream bounds it but index arithmetic
// on run-time strides, per-score masks and selects, exp2f's range handling, the zeroing of accumulators and whole MMAs on
// key tiles that hold no key.  Here: HPB heads per CTA (one warp each), row stride and head offset are immediates of the
// ldmatrix instructions, the T - 1 clamp of the fragment rows is six per-lane offsets computed once, only the key
// tile that straddles T is masked, key tiles past T are not multiplied at all, the scale lives in the exponent
// (2^(s c - max c), ex2.approx), the first MMA of a chain takes C = 0 from the zero register.
// ---------------------------------------------------------------------------------------------------------------
template <int D, int T, int HPB>
__global__ void __launch_bounds__(HPB * 32) temporal_attn_fixed_kernel(const __half* __restrict__ q, long long ldq,
                                                                       const __half* __restrict__ k, long long ldk,
                                                                       const __half* __restrict__ v, long long ldv,
                                                                       __half* __restrict__ o, long long ldo, int HW,
                                                                       float scale_log2) {
  constexpr int C = HPB * D, RS = C + 8, CPR = C / 8;
  constexpr int MB = (T + 15) / 16, NT = (T + 7) / 8, NP = (NT + 1) / 2, KK = (T + 15) / 16;
  constexpr int KSTEPS = (D + 15) / 16, NT_O = D / 8;
  constexpr bool kTail = (D % 16) != 0;
  extern __shared__ __align__(16) uint8_t ta_smem[];
  __half* sq = reinterpret_cast<__half*>(ta_smem);        // [T][RS]; k and v follow
  __half* sk = sq + T * RS;
  __half* sv = sk + T * RS;
  const long long col0 = static_cast<long long>(blockIdx.y) * C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long pix = blockIdx.x;                       // b*HW + hw
  const long long row0 = (pix / HW) * T * HW + pix % HW;  // row of frame 0; frame t adds t*HW
  griddep_wait();                  // PDL: q, k, v come from the previous kernel of the stream
  griddep_launch_dependents();
  for (int i = threadIdx.x; i < T * CPR; i += HPB * 32) {
    const int t = i / CPR, c = i - t * CPR;
    const long long row = row0 + static_cast<long long>(t) * HW;
    const uint32_t dst = static_cast<uint32_t>(t * RS + c * 8) * 2u;
    cp_async_16(smem_u32(sq) + dst, q + row * ldq + col0 + c * 8, true);
    cp_async_16(smem_u32(sk) + dst, k + row * ldk + col0 + c * 8, true);
    cp_async_16(smem_u32(sv) + dst, v + row * ldv + col0 + c * 8, true);
  }
  // the 8 padding halves at the end of every row feed the last k-step of the last head: keep them finite
  for (int i = threadIdx.x; i < 3 * T; i += HPB * 32) *reinterpret_cast<uint4*>(sq + i * RS + C) = make_uint4(0, 0, 0, 0);
  cp_async_commit();
  // fragment rows past T - 1 alias frame T - 1 (masked keys / never-stored query rows): per-lane byte offsets, once
  uint32_t qrow[MB], krow[NP], vrow[KK];
#pragma unroll
  for (int i = 0; i < MB; ++i) qrow[i] = static_cast<uint32_t>(min(i * 16 + (lane & 15), T - 1) * RS + (lane >> 4) * 8) * 2u;
#pragma unroll
  for (int i = 0; i < NP; ++i)
    krow[i] = static_cast<uint32_t>(min(i * 16 + (lane & 7) + ((lane >> 4) << 3), T - 1) * RS + ((lane >> 3) & 1) * 8) * 2u;
#pragma unroll
  for (int i = 0; i < KK; ++i)
    vrow[i] = static_cast<uint32_t>(min(i * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, T - 1) * RS + (lane >> 4) * 8) * 2u;
  const uint32_t hq = smem_u32(sq) + warp * (D * 2), hk = smem_u32(sk) + warp * (D * 2), hv = smem_u32(sv) + warp * (D * 2);
  const int ocol = (lane & 3) * 2;
  const float c = scale_log2;
  cp_async_wait<0>();
  __syncthreads();

#pragma unroll
  for (int mb = 0; mb < MB; ++mb) {
    // ---- S = Q K^T for 16 query frames x the NT key tiles that hold a key ----
    float sacc[NT][4];
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
      uint32_t qf[4];
      // d % 16 == 8: the upper half of the last k-step lies past d, in the next head's columns (which that head's warp may
      // be overwriting with its output rows): those lanes read the zero padding at the end of the row instead
      const bool past = kTail && ks == KSTEPS - 1 && lane >= 16;
      ldmatrix_x4(qf, past ? smem_u32(sq) + qrow[mb] - 16 + C * 2 : hq + qrow[mb] + ks * 32);
#pragma unroll
      for (int np = 0; np < NP; ++np) {
        uint32_t kf[4];
        ldmatrix_x4(kf, hk + krow[np] + ks * 32);
        const uint32_t b0[2] = {kf[0], kf[1]}, b1[2] = {kf[2], kf[3]};
        if (ks == 0) {
          mma_m16n8k16_first(sacc[2 * np], qf, b0);
          if (2 * np + 1 < NT) mma_m16n8k16_first(sacc[2 * np + 1], qf, b1);
        } else {
          mma_m16n8k16(sacc[2 * np], qf, b0);
          if (2 * np + 1 < NT) mma_m16n8k16(sacc[2 * np + 1], qf, b1);
        }
      }
    }
    // ---- softmax over the T keys (rows g and g + 8 of the fragment) ----
    if (T % 8 != 0) {
      const int col = (NT - 1) * 8 + ocol;
      if (col >= T) sacc[NT - 1][0] = sacc[NT - 1][2] = -INFINITY;
      if (col + 1 >= T) sacc[NT - 1][1] = sacc[NT - 1][3] = -INFINITY;
    }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      mx[0] = fmaxf(mx[0], fmaxf(sacc[nt][0], sacc[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(sacc[nt][2], sacc[nt][3]));
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 1));
      mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 2));
    }
    const float nm0 = -mx[0] * c, nm1 = -mx[1] * c;
    float rs[2] = {0.f, 0.f};
    uint32_t pf[KK][4];
#pragma unroll
    for (int nt = 0; nt < 2 * KK; ++nt) {
      if (nt < NT) {
        const float p0 = ex2_fast(fmaf(sacc[nt][0], c, nm0)), p1 = ex2_fast(fmaf(sacc[nt][1], c, nm0));
        const float p2 = ex2_fast(fmaf(sacc[nt][2], c, nm1)), p3 = ex2_fast(fmaf(sacc[nt][3], c, nm1));
        rs[0] += p0 + p1;
        rs[1] += p2 + p3;
        __half2 h01 = __floats2half2_rn(p0, p1), h23 = __floats2half2_rn(p2, p3);
        pf[nt >> 1][(nt & 1) * 2 + 0] = *reinterpret_cast<uint32_t*>(&h01);
        pf[nt >> 1][(nt & 1) * 2 + 1] = *reinterpret_cast<uint32_t*>(&h23);
      } else {                                                     // the key tile past T inside the last 16-key step
        pf[nt >> 1][(nt & 1) * 2 + 0] = 0u;
        pf[nt >> 1][(nt & 1) * 2 + 1] = 0u;
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], 1);
      rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], 2);
    }
    // ---- O = P V (the 8 columns past d of the last 16-column group are not computed) ----
    float oacc[NT_O][4];
#pragma unroll
    for (int kk = 0; kk < KK; ++kk) {
#pragma unroll
      for (int np = 0; np < (NT_O + 1) / 2; ++np) {
        uint32_t vf[4];
        ldmatrix_x4_trans(vf, hv + vrow[kk] + np * 32);
        const uint32_t b0[2] = {vf[0], vf[1]}, b1[2] = {vf[2], vf[3]};
        if (kk == 0) {
          mma_m16n8k16_first(oacc[2 * np], pf[kk], b0);
          if (2 * np + 1 < NT_O) mma_m16n8k16_first(oacc[2 * np + 1], pf[kk], b1);
        } else {
          mma_m16n8k16(oacc[2 * np], pf[kk], b0);
          if (2 * np + 1 < NT_O) mma_m16n8k16(oacc[2 * np + 1], pf[kk], b1);
        }
      }
    }
    const float inv0 = __fdividef(1.f, rs[0]), inv1 = __fdividef(1.f, rs[1]);
    // the output rows of this head replace its query rows (Q of block mb only feeds S of block mb; a later block's
    // clamped rows read frame T - 1, which belongs to the last block)
    const int r0 = mb * 16 + (lane >> 2), r1 = r0 + 8;
    __half* o0 = sq + r0 * RS + warp * D + ocol;
    __half* o1 = o0 + 8 * RS;
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < NT_O; ++nt) {
      if (r0 < T) *reinterpret_cast<__half2*>(o0 + nt * 8) = __floats2half2_rn(oacc[nt][0] * inv0, oacc[nt][1] * inv0);
      if (r1 < T) *reinterpret_cast<__half2*>(o1 + nt * 8) = __floats2half2_rn(oacc[nt][2] * inv1, oacc[nt][3] * inv1);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * CPR; i += HPB * 32) {
    const int t = i / CPR, cc = i - t * CPR;
    *reinterpret_cast<uint4*>(o + (row0 + static_cast<long long>(t) * HW) * ldo + col0 + cc * 8) =
        *reinterpret_cast<const uint4*>(sq + t * RS + cc * 8);
  }
}

template <int D, int T, int HPB>
static int launch_ta_fixed(const __half* q, long long ldq, const __half* k, long long ldk, const __half* v, long long ldv,
                           __half* o, long long ldo, int B, int HW, int heads, float scale_log2, cudaStream_t st) {
  constexpr size_t smem = static_cast<size_t>(3) * T * (HPB * D + 8) * 2;
  static_assert(smem <= 227 * 1024, "temporal_attn_fixed_kernel: shared memory");
  static std::atomic<bool> attr_done[kMaxDevices];
  const int dev = current_device();
  CCEDIT_CHECK_ARG(dev >= 0, "ccedit_temporal_attention: no current CUDA device");
  if (!attr_done[dev].load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(temporal_attn_fixed_kernel<D, T, HPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_last_error("ccedit_temporal_attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return CCEDIT_ERR_CUDA;
    }
    attr_done[dev].store(true, std::memory_order_release);
  }
  const dim3 grid(static_cast<unsigned>(static_cast<long long>(B) * HW), heads / HPB);
  (void)launch_pdl(4, temporal_attn_fixed_kernel<D, T, HPB>, grid, dim3(HPB * 32), smem, st, q, ldq, k, ldk, v, ldv, o, ldo, HW, scale_log2);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  CCEDIT_CUDA_LAUNCH_CHECK("ccedit_temporal_attention");
  return CCEDIT_OK;
}

// the network's shapes -> the specialised kernel; -1 = not one of them (generic kernel)
static int dispatch_ta_fixed(const __half* q, long long ldq, const __half* k, long long ldk, const __half* v, long long ldv,
                             __half* o, long long ldo, int B, int T, int HW, int heads, int d, float sl, cudaStream_t st) {
  static const bool off = [] { const char* e = getenv("CCEDIT_TA_FIXED"); return e && e[0] == '0'; }();
  if (off || heads != 8) return -1;
#define CCEDIT_TAF(D_, T_, H_) return launch_ta_fixed<D_, T_, H_>(q, ldq, k, ldk, v, ldv, o, ldo, B, HW, heads, sl, st);
  if (d == 40) {      // heads per CTA: all 8 while three CTAs per SM fit (<= 72 KB), else the largest group that does
    if (T == 9) CCEDIT_TAF(40, 9, 8)
    if (T == 17) CCEDIT_TAF(40, 17, 8)
    if (T == 33) CCEDIT_TAF(40, 33, 8)
  } else if (d == 80) {
    if (T == 9) CCEDIT_TAF(80, 9, 8)
    if (T == 17) CCEDIT_TAF(80, 17, 8)
    if (T == 33) CCEDIT_TAF(80, 33, 4)
  } else if (d == 160) {
    if (T == 9) CCEDIT_TAF(160, 9, 8)
    if (T == 17) CCEDIT_TAF(160, 17, 4)
    if (T == 33) CCEDIT_TAF(160, 33, 2)
  }
#undef CCEDIT_TAF
  return -1;
}

template <int TK>
static int dispatch_ta(const __half* q, long long ldq, const __half* k, long long ldk, const __half* v, long long ldv,
                       __half* o, long long ldo, int B, int T, int HW, int heads, int d, float sl, cudaStream_t st) {
  if (d <= 16) return launch_ta<16, TK>(q, ldq, k, ldk, v, ldv, o, ldo, B, T, HW, heads, d, sl, st);
  if (d <= 32) return launch_ta<32, TK>(q, ldq, k, ldk, v, ldv, o, ldo, B, T, HW, heads, d, sl, st);
  if (d <= 48) return launch_ta<48, TK>(q, ldq, k, ldk, v, ldv, o, ldo, B, T, HW, heads, d, sl, st);
  if (d <= 64) return launch_ta<64, TK>(q, ldq, k, ldk, v, ldv, o, ldo, B, T, HW, heads, d, sl, st);
  if (d <= 80) return launch_ta<80, TK>(q, ldq, k, ldk, v, ldv, o, ldo, B, T, HW, heads, d, sl, st);
  if (d <= 128) return launch_ta<128, TK>(q, ldq, k, ldk, v, ldv, o, ldo, B, T, HW, heads, d, sl, st);
  return launch_ta<160, TK>(q, ldq, k, ldk, v, ldv, o, ldo, B, T, HW, heads, d, sl, st);
}


// ---------------------------------------------------------------------------------------------------------------
// Attention over a SHORT key sequence shared by many query rows: the text cross-attention (77 keys per batch entry,
// attention.py:444-448 with context = the CLIP tokens; 21 launches per network call).  q and o are the whole cost
// (read 134 MB, write 134 MB at the top level for 0.03 TFLOP): the kernel is an HBM stream with a little tensor work
// per row, and the long-sequence kernels - one (query tile, head) per pipeline pass - ran it at 1.2-1.6 TB/s.
//   * a CTA owns one (key/value frame, head group) pair: the group's K and V columns ([lkv][G], G = heads_per_group * d
//     <= 160 channels = 320 contiguous bytes per row) are staged in shared memory ONCE and stay there;
//   * its warps walk the (query frame, 16-row block) units of that pair independently - no block-wide barrier in the
//     loop: the unit's q rows arrive with 16-byte cp.async into a per-warp double buffer (the next unit's rows are in
//     flight while this one is computed), S = Q K^T, the softmax and O = P V run on warp-level tensor-core MMAs
//     (m16n8k16, everything in registers, one key "tile": no online rescaling), the output rows replace the query rows in
//     the buffer and leave with 16-byte coalesced stores.
// Shared-memory rows are G + 8 halves (odd multiple of 16 bytes): conflict-free ldmatrix.
// ---------------------------------------------------------------------------------------------------------------
struct SkParams {
  const __half* q; long long ldq, q_fs;
  __half* o; long long ldo, o_fs;
  const __half* k; const __half* v;
  long long ldk, ldv, kv_fs;
  int lkv, kv_div, kv_mul, kv_add;
  int frames, lq, d, hpg, ngroups, nkvf;
  float scale_log2;
};

// D: head dim (40 / 80 / 160: a head group is 160 / D heads = 160 channels, so every shared-memory offset below is an
// immediate); NKT: 16-key groups (lkv in (16 NKT - 16, 16 NKT]: only the last two 8-key tiles can hold masked keys);
// NW: warps per CTA (one CTA per SM: the warps are what hides the ldmatrix / HMMA / shuffle latencies of the per-head
// chain - 12 while the accumulators leave room (14 warps cap the registers at 128: measured slower), 8 at d = 160).
template <int D, int NKT, int NW>
__global__ void __launch_bounds__(NW * 32, 1) short_kv_attn_kernel(const __grid_constant__ SkParams p) {
  constexpr int HPG = 160 / D, G = HPG * D, RS = G + 8, CPR = G / 8;        // 20 16-byte chunks per row
  constexpr int KSTEPS = (D + 15) / 16, NT_O = D / 8, NTK = 2 * NKT, KP = 16 * NKT;
  constexpr bool kTail = (D % 16) != 0;                   // d % 16 == 8: the last k-step covers 8 foreign channels
  constexpr int NJ = (16 * CPR + 31) / 32;                // 16-byte chunks of a unit per lane
  extern __shared__ __align__(16) uint8_t sk_smem[];
  __half* sK = reinterpret_cast<__half*>(sk_smem);         // [KP][RS]
  __half* sV = sK + KP * RS;                               // [KP][RS]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __half* sQ = sV + KP * RS + warp * 2 * 16 * RS;          // this warp's [2][16][RS]

  // (key/value frame, head group) pairs are dealt round-robin to the CTAs; the CTAs of a pair share its units
  const int npairs = p.nkvf * p.ngroups;
  const int pair = blockIdx.x % npairs;
  const int sub = blockIdx.x / npairs, nsub = (gridDim.x - pair + npairs - 1) / npairs;   // CTAs working on this pair
  const int kvi = pair / p.ngroups, grp = pair % p.ngroups;
  const int f_begin = kvi * p.kv_div, f_end = min(p.frames, f_begin + p.kv_div);
  const long long col0 = static_cast<long long>(grp) * G;
  griddep_wait();                  // PDL: q, k, v come from the previous kernels of the stream
  griddep_launch_dependents();
  {
    const long long kvf = static_cast<long long>(kvi) * p.kv_mul + p.kv_add;
    const __half* kb = p.k + kvf * p.kv_fs + col0;
    const __half* vb = p.v + kvf * p.kv_fs + col0;
    for (int i = threadIdx.x; i < KP * CPR; i += NW * 32) {
      const int r = i / CPR, c = i - r * CPR;
      const bool ok = r < p.lkv;                           // rows lkv..KP-1: zero-filled (masked keys, finite V)
      const long long rr = ok ? r : 0;
      cp_async_16(smem_u32(sK + r * RS + c * 8), kb + rr * p.ldk + c * 8, ok);
      cp_async_16(smem_u32(sV + r * RS + c * 8), vb + rr * p.ldv + c * 8, ok);
    }
    // the 8 padding halves of every row feed the last k-step / output group of the last head: keep them finite
    for (int i = threadIdx.x; i < 2 * KP; i += NW * 32) *reinterpret_cast<uint4*>(sK + i * RS + G) = make_uint4(0, 0, 0, 0);
    for (int i = lane; i < 2 * 16; i += 32) *reinterpret_cast<uint4*>(sQ + i * RS + G) = make_uint4(0, 0, 0, 0);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
  }

  // this lane's chunks of a unit: (row, 16-byte column) pairs, fixed for the whole kernel
  int grow[NJ];                                            // row within the unit, -1 past the end
  uint32_t soff[NJ];                                       // byte offset in the unit's buffer
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int i = lane + 32 * j, r = i / CPR, c = i - r * CPR;
    grow[j] = i < 16 * CPR ? r : -1;
    soff[j] = static_cast<uint32_t>(r * RS + c * 8) * 2u;
  }
  const int nrb = (p.lq + 15) >> 4;                        // 16-row blocks per frame
  const int nunits = (f_end - f_begin) * nrb;              // < 2^31: frames * rows / 16
  const int ustep = nsub * NW;
  int uf0 = 0, uf1 = 0, ur0 = 0, ur1 = 0;                 // frame and first row of the unit in each buffer
  const uint32_t sq_base = smem_u32(sQ);
  auto stage = [&](int u, int buf) {                       // unit u -> this warp's buffer buf (one cp.async group)
    if (u < nunits) {
      const int fi = u / nrb;
      const int f = f_begin + fi, r0 = (u - fi * nrb) << 4;
      if (buf) { uf1 = f; ur1 = r0; } else { uf0 = f; ur0 = r0; }
      const __half* qb = p.q + static_cast<long long>(f) * p.q_fs + col0;
      const uint32_t dst = sq_base + buf * (16 * RS * 2);
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        if (grow[j] >= 0) {
          const int row = min(r0 + grow[j], p.lq - 1);     // rows past the end repeat the last row (never stored)
          cp_async_16(dst + soff[j], qb + static_cast<long long>(row) * p.ldq + ((soff[j] >> 1) - grow[j] * RS), true);
        }
    }
    cp_async_commit();
  };
  int u = sub * NW + warp;
  stage(u, 0);
  stage(u + ustep, 1);
  const float c = p.scale_log2;
  // per-lane ldmatrix byte offsets (row stride and head width are compile-time: the rest are instruction immediates)
  const uint32_t q_lane = static_cast<uint32_t>((lane & 15) * RS + (lane >> 4) * 8) * 2u;
  const uint32_t k_lane = smem_u32(sK) + static_cast<uint32_t>(((lane & 7) + ((lane >> 4) << 3)) * RS + ((lane >> 3) & 1) * 8) * 2u;
  const uint32_t v_lane = smem_u32(sV) + static_cast<uint32_t>(((lane & 7) + ((lane >> 3) & 1) * 8) * RS + (lane >> 4) * 8) * 2u;
  const int ocol = (lane & 3) * 2;
  int buf = 0;
  for (; u < nunits; u += ustep, buf ^= 1) {
    cp_async_wait<1>();
    __syncwarp();
    const uint32_t bq = sq_base + buf * (16 * RS * 2);
#pragma unroll 1
    for (int h = 0; h < HPG; ++h) {
      const uint32_t hq = bq + q_lane + h * (D * 2), hk = k_lane + h * (D * 2), hv = v_lane + h * (D * 2);
      // ---- S = Q K^T: 16 query rows x KP keys ----
      float sacc[NTK][4];
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) {
        uint32_t qf[4];
        ldmatrix_x4(qf, hq + ks * 32);
        if (kTail && ks == KSTEPS - 1) qf[2] = qf[3] = 0u;        // zero the 8 channels past d
#pragma unroll
        for (int np = 0; np < NKT; ++np) {
          uint32_t kf[4];
          ldmatrix_x4(kf, hk + np * (16 * RS * 2) + ks * 32);
          const uint32_t b0[2] = {kf[0], kf[1]}, b1[2] = {kf[2], kf[3]};
          if (ks == 0) {
            mma_m16n8k16_first(sacc[2 * np], qf, b0);
            mma_m16n8k16_first(sacc[2 * np + 1], qf, b1);
          } else {
            mma_m16n8k16(sacc[2 * np], qf, b0);
            mma_m16n8k16(sacc[2 * np + 1], qf, b1);
          }
        }
      }
      // ---- softmax over the lkv valid keys (rows g and g + 8 of the fragment); the scale (> 0) is applied inside the
      //      exponential: 2^(s * c - max * c) ----
#pragma unroll
      for (int nt = NTK - 2; nt < NTK; ++nt) {
        const int col = nt * 8 + ocol;
        if (col >= p.lkv) sacc[nt][0] = sacc[nt][2] = -INFINITY;
        if (col + 1 >= p.lkv) sacc[nt][1] = sacc[nt][3] = -INFINITY;
      }
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < NTK; ++nt) {
        mx[0] = fmaxf(mx[0], fmaxf(sacc[nt][0], sacc[nt][1]));
        mx[1] = fmaxf(mx[1], fmaxf(sacc[nt][2], sacc[nt][3]));
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 1));
        mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 2));
      }
      const float nm0 = -mx[0] * c, nm1 = -mx[1] * c;
      float rs[2] = {0.f, 0.f};
      uint32_t pf[NKT][4];
#pragma unroll
      for (int nt = 0; nt < NTK; ++nt) {
        const float p0 = ex2_fast(fmaf(sacc[nt][0], c, nm0)), p1 = ex2_fast(fmaf(sacc[nt][1], c, nm0));
        const float p2 = ex2_fast(fmaf(sacc[nt][2], c, nm1)), p3 = ex2_fast(fmaf(sacc[nt][3], c, nm1));
        rs[0] += p0 + p1;
        rs[1] += p2 + p3;
        __half2 h01 = __floats2half2_rn(p0, p1), h23 = __floats2half2_rn(p2, p3);
        pf[nt >> 1][(nt & 1) * 2 + 0] = *reinterpret_cast<uint32_t*>(&h01);
        pf[nt >> 1][(nt & 1) * 2 + 1] = *reinterpret_cast<uint32_t*>(&h23);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], 1);
        rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], 2);
      }
      // ---- O = P V (the 8 columns past d of the last 16-column group are not computed) ----
      float oacc[NT_O][4];
#pragma unroll
      for (int kk = 0; kk < NKT; ++kk) {
#pragma unroll
        for (int np = 0; np < (NT_O + 1) / 2; ++np) {
          uint32_t vf[4];
          ldmatrix_x4_trans(vf, hv + kk * (16 * RS * 2) + np * 32);
          const uint32_t b0[2] = {vf[0], vf[1]}, b1[2] = {vf[2], vf[3]};
          if (kk == 0) {
            mma_m16n8k16_first(oacc[2 * np], pf[kk], b0);
            if (2 * np + 1 < NT_O) mma_m16n8k16_first(oacc[2 * np + 1], pf[kk], b1);
          } else {
            mma_m16n8k16(oacc[2 * np], pf[kk], b0);
            if (2 * np + 1 < NT_O) mma_m16n8k16(oacc[2 * np + 1], pf[kk], b1);
          }
        }
      }
      const float inv0 = __fdividef(1.f, rs[0]), inv1 = __fdividef(1.f, rs[1]);
      // the head's output rows replace its query rows in the buffer (Q of head h only feeds S of head h; the foreign
      // channels its last k-step touched were zeroed in the fragment, not in memory)
      __half* o0 = sQ + buf * (16 * RS) + (lane >> 2) * RS + h * D + ocol;
      __half* o1 = o0 + 8 * RS;
      __syncwarp();
#pragma unroll
      for (int nt = 0; nt < NT_O; ++nt) {
        *reinterpret_cast<__half2*>(o0 + nt * 8) = __floats2half2_rn(oacc[nt][0] * inv0, oacc[nt][1] * inv0);
        *reinterpret_cast<__half2*>(o1 + nt * 8) = __floats2half2_rn(oacc[nt][2] * inv1, oacc[nt][3] * inv1);
      }
      __syncwarp();
    }
    {
      const int f = buf ? uf1 : uf0, r0 = buf ? ur1 : ur0;
      __half* ob = p.o + static_cast<long long>(f) * p.o_fs + col0;
      const uint8_t* src = reinterpret_cast<const uint8_t*>(sQ) + buf * (16 * RS * 2);
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        if (grow[j] >= 0 && r0 + grow[j] < p.lq)
          *reinterpret_cast<uint4*>(ob + static_cast<long long>(r0 + grow[j]) * p.ldo + ((soff[j] >> 1) - grow[j] * RS)) =
              *reinterpret_cast<const uint4*>(src + soff[j]);
    }
    __syncwarp();                                          // the buffer has been read: the unit after next may land in it
    stage(u + 2 * ustep, buf);
  }
  cp_async_wait<0>();
}

template <int D>
constexpr int sk_warps() { return D <= 80 ? 12 : 8; }

template <int D, int NKT>
static int launch_sk(const SkParams& p, int grid, cudaStream_t st) {
  constexpr int NW = sk_warps<D>();
  constexpr int RS = (160 / D) * D + 8;
  const size_t smem = (static_cast<size_t>(2) * 16 * NKT + static_cast<size_t>(NW) * 2 * 16) * RS * 2;
  if (smem > 227 * 1024) return -1;
  static std::atomic<bool> attr_done[kMaxDevices];
  const int dev = current_device();
  CCEDIT_CHECK_ARG(dev >= 0, "ccedit_attention: no current CUDA device");
  if (!attr_done[dev].load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(short_kv_attn_kernel<D, NKT, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_last_error("ccedit_attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return CCEDIT_ERR_CUDA;
    }
    attr_done[dev].store(true, std::memory_order_release);
  }
  (void)launch_pdl(4, short_kv_attn_kernel<D, NKT, NW>, dim3(grid), dim3(NW * 32), smem, st, p);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  CCEDIT_CUDA_LAUNCH_CHECK("ccedit_attention(short kv)");
  return CCEDIT_OK;
}

// -> -1 when the problem is not of this kernel's kind (the caller goes on to the long-sequence kernels)
static int attention_short_kv(const ccedit_attn_desc* a, cudaStream_t st) {
  if (a->nseg != 1 || a->lkv[0] > 128 || a->kv_div[0] < 2) return -1;       // one short segment shared by >= 2 query frames
  const int d = a->d;
  if (d != 40 && d != 80 && d != 160) return -1;
  const int hpg = 160 / d;                                               // a head group is 160 channels wide
  if (a->heads % hpg != 0) return -1;
  SkParams p;
  memset(&p, 0, sizeof(p));
  p.hpg = hpg;
  p.ngroups = a->heads / hpg;
  p.nkvf = (a->frames + a->kv_div[0] - 1) / a->kv_div[0];
  const int sms = device_sm_count();
  const int npairs = p.nkvf * p.ngroups;
  if (sms <= 0 || npairs > sms) return -1;
  // the rows must be worth a resident K/V copy per CTA
  if (static_cast<long long>(a->frames) * a->lq < 1024) return -1;
  p.q = static_cast<const __half*>(a->q);  p.ldq = a->ldq;  p.q_fs = a->q_frame_stride;
  p.o = static_cast<__half*>(a->o);        p.ldo = a->ldo;  p.o_fs = a->o_frame_stride;
  p.k = static_cast<const __half*>(a->k[0]);
  p.v = static_cast<const __half*>(a->v[0]);
  p.ldk = a->ldk[0];  p.ldv = a->ldv[0];  p.kv_fs = a->kv_frame_stride[0];
  p.lkv = a->lkv[0];  p.kv_div = a->kv_div[0];  p.kv_mul = a->kv_mul[0];  p.kv_add = a->kv_add[0];
  p.frames = a->frames;  p.lq = a->lq;  p.d = d;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  if ((reinterpret_cast<uintptr_t>(p.q) | reinterpret_cast<uintptr_t>(p.o) | reinterpret_cast<uintptr_t>(p.k) |
       reinterpret_cast<uintptr_t>(p.v)) & 15) return -1;
  const int nkt = (p.lkv + 15) / 16;
  // every pair gets the same number of CTAs
  const int grid = (sms / npairs) * npairs;
#define CCEDIT_SK(DP_)                                                            \
  switch (nkt) {                                                                  \
    case 1: return launch_sk<DP_, 1>(p, grid, st);                          \
    case 2: return launch_sk<DP_, 2>(p, grid, st);                          \
    case 3: return launch_sk<DP_, 3>(p, grid, st);                          \
    case 4: return launch_sk<DP_, 4>(p, grid, st);                          \
    case 5: return launch_sk<DP_, 5>(p, grid, st);                          \
    case 6: return launch_sk<DP_, 6>(p, grid, st);                          \
    case 7: return launch_sk<DP_, 7>(p, grid, st);                          \
    default: return launch_sk<DP_, 8>(p, grid, st);                         \
  }
  if (d == 40) { CCEDIT_SK(40) }
  if (d == 80) { CCEDIT_SK(80) }
  CCEDIT_SK(160)
#undef CCEDIT_SK
}

}  // namespace ccedit

using namespace ccedit;

extern "C" int ccedit_attention(const ccedit_attn_desc* a, void* stream) {
  CCEDIT_CHECK_ARG(a != nullptr, "ccedit_attention: null descriptor");
  CCEDIT_CHECK_ARG(a->q && a->o && a->nseg >= 1 && a->nseg <= 2, "ccedit_attention: bad q/o/nseg");
  CCEDIT_CHECK_ARG(a->d > 0 && a->d % 8 == 0 && a->d <= 160, "ccedit_attention: head dim %d unsupported (multiple of 8, <=160)", a->d);
  CCEDIT_CHECK_ARG(a->frames > 0 && a->lq > 0 && a->heads > 0, "ccedit_attention: empty problem");
  CCEDIT_CHECK_ARG(a->ldq % 8 == 0 && a->ldo % 8 == 0 && a->q_frame_stride % 8 == 0 && a->o_frame_stride % 8 == 0,
                   "ccedit_attention: q/o strides must be multiples of 8 elements");
  FaParams p;
  memset(&p, 0, sizeof(p));
  p.q = static_cast<const __half*>(a->q);
  p.ldq = a->ldq;
  p.q_fs = a->q_frame_stride;
  p.o = static_cast<__half*>(a->o);
  p.ldo = a->ldo;
  p.o_fs = a->o_frame_stride;
  p.nseg = a->nseg;
  for (int s = 0; s < a->nseg; ++s) {
    CCEDIT_CHECK_ARG(a->k[s] && a->v[s] && a->lkv[s] > 0 && a->kv_div[s] > 0, "ccedit_attention: bad kv segment %d", s);
    CCEDIT_CHECK_ARG(a->ldk[s] % 8 == 0 && a->ldv[s] % 8 == 0 && a->kv_frame_stride[s] % 8 == 0,
                     "ccedit_attention: kv strides must be multiples of 8 elements");
    p.k[s] = static_cast<const __half*>(a->k[s]);
    p.v[s] = static_cast<const __half*>(a->v[s]);
    p.ldk[s] = a->ldk[s];
    p.ldv[s] = a->ldv[s];
    p.kv_fs[s] = a->kv_frame_stride[s];
    p.lkv[s] = a->lkv[s];
    p.kv_div[s] = a->kv_div[s];
    p.kv_mul[s] = a->kv_mul[s];
    p.kv_add[s] = a->kv_add[s];
  }
  p.lq = a->lq;
  p.d = a->d;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // head dims up to 64 (d = 40: 88 % of the attention FLOPs) run on tcgen05 / TMEM; CCEDIT_ATTN_LEGACY=1 forces the
  // mma.sync kernel (A/B measurements only)
  static const bool legacy = [] { const char* e = getenv("CCEDIT_ATTN_LEGACY"); return e && e[0] == '1'; }();
  static const bool no_short = [] { const char* e = getenv("CCEDIT_ATTN_SHORT"); return e && e[0] == '0'; }();
  if (!legacy && !no_short && a->scale > 0.f) {             // a short key sequence shared by many rows: HBM stream of q and o
    const int rc = attention_short_kv(a, st);
    if (rc >= 0) return rc;
  }
  if (!legacy && (reinterpret_cast<uintptr_t>(a->q) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->o) & 15) == 0 &&
      a->scale > 0.f) {
    const int rc = attention_tc(a, st);
    if (rc >= 0) return rc;
  }
  if (a->d <= 16) return launch_fa<16>(p, a->frames, a->heads, st);
  if (a->d <= 32) return launch_fa<32>(p, a->frames, a->heads, st);
  if (a->d <= 48) return launch_fa<48>(p, a->frames, a->heads, st);
  if (a->d <= 64) return launch_fa<64>(p, a->frames, a->heads, st);
  if (a->d <= 80) return launch_fa<80>(p, a->frames, a->heads, st);
  if (a->d <= 128) return launch_fa<128>(p, a->frames, a->heads, st);
  return launch_fa<160>(p, a->frames, a->heads, st);
}

extern "C" int ccedit_temporal_attention(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                                         int64_t ldv, void* o, int64_t ldo, int32_t B, int32_t T, int32_t HW,
                                         int32_t heads, int32_t d, float scale, void* stream) {
  CCEDIT_CHECK_ARG(q && k && v && o, "ccedit_temporal_attention: null pointer");
  CCEDIT_CHECK_ARG(B > 0 && T > 0 && T <= 64 && HW > 0 && heads > 0 && d > 0 && d % 8 == 0 && d <= 160,
                   "ccedit_temporal_attention: bad shape B=%d T=%d HW=%d heads=%d d=%d (T<=64, d%%8==0, d<=160)", B, T, HW,
                   heads, d);
  CCEDIT_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0,
                   "ccedit_temporal_attention: row strides must be multiples of 8 elements");
  CCEDIT_CHECK_ARG(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
                     reinterpret_cast<uintptr_t>(o)) & 15) == 0,
                   "ccedit_temporal_attention: q/k/v/o must be 16-byte aligned");
  const __half* qp = static_cast<const __half*>(q);
  const __half* kp = static_cast<const __half*>(k);
  const __half* vp = static_cast<const __half*>(v);
  __half* op = static_cast<__half*>(o);
  const float sl = scale * 1.4426950408889634f;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (scale > 0.f) {
    const int rc = dispatch_ta_fixed(qp, ldq, kp, ldk, vp, ldv, op, ldo, B, T, HW, heads, d, sl, st);
    if (rc >= 0) return rc;
  }
  if (T <= 32) return dispatch_ta<32>(qp, ldq, kp, ldk, vp, ldv, op, ldo, B, T, HW, heads, d, sl, st);
  return dispatch_ta<64>(qp, ldq, kp, ldk, vp, ldv, op, ldo, B, T, HW, heads, d, sl, st);
}

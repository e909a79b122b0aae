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
// Roles (192 threads): warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + MMA issuer (one lane),
// warps 2..5 = epilogue (TMEM -> registers -> bias/emb/SiLU/GEGLU/residual -> fp16 global stores).
#include "common.cuh"
#include "../../include/ccedit_b200.h"

#include <atomic>
#include <mutex>

namespace ccedit {

extern std::atomic<long long> g_launch_count;

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                  // 64 fp16 = 128 B = one SWIZZLE_128B row
constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KiB
constexpr int kMaxStages = 8;
constexpr int kGemmThreads = 192;
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;              // columns between the two accumulator buffers

struct GemmKParams {
  int box[4];
  int odim[4];
  int tiles[4];
  int n_tiles;
  int total_tiles;
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

__global__ void __launch_bounds__(kGemmThreads, 1)
tap_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ GemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);  // SWIZZLE_128B wants 1024 B alignment

  const int stage_bytes = kABytes + p.bn * kBlockK * 2;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int kblocks = p.ntaps * p.kchunks;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles;
        int m = tile / p.n_tiles;
        int o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          o[i] = (m % p.tiles[i]) * p.box[i];
          m /= p.tiles[i];
        }
        for (int tap = 0; tap < p.ntaps; ++tap) {
          const int c1 = o[0] + p.taps[tap][0], c2 = o[1] + p.taps[tap][1];
          const int c3 = o[2] + p.taps[tap][2], c4 = o[3] + p.taps[tap][3];
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            uint8_t* sa = smem + stage * stage_bytes;
            mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(stage_bytes));
            tma_load_5d(sa, &tmA, &full_bar[stage], kc * kBlockK, c1, c2, c3, c4);
            tma_load_2d(sa + kABytes, &tmB, &full_bar[stage], (tap * p.kchunks + kc) * kBlockK, n_tile * p.bn);
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[as], aphase ^ 1u);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * kAccStride);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_u32(smem + stage * stage_bytes);
          const uint64_t adesc = umma_desc_k_sw128(sa);
          const uint64_t bdesc = umma_desc_k_sw128(sa + kABytes);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // advance 16 elements (32 B) inside the 128 B swizzle row: +2 in 16-byte units
            umma_f16_ss(d_tmem, adesc + 2u * k, bdesc + 2u * k, p.idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(&tfull_bar[as]);  // accumulator complete -> epilogue
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int wq = warp & 3;  // TMEM lane quarter this warp may access
    const int row = wq * 32 + lane;
    int as = 0;
    uint32_t aphase = 0;
    const bool geglu = (p.flags & CCEDIT_GEMM_GEGLU) != 0;
    const bool do_silu = (p.flags & CCEDIT_GEMM_SILU) != 0;
    const int ncols_out = geglu ? p.bn / 2 : p.bn;
    // row -> local coordinates inside the tile box
    int l[4];
    {
      int r = row;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        l[i] = r % p.box[i];
        r /= p.box[i];
      }
    }
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.n_tiles;
      int m = tile / p.n_tiles;
      bool valid = true;
      long long off_o = 0, off_r1 = 0, off_r2 = 0;
      int rb_row = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = (m % p.tiles[i]) * p.box[i] + l[i];
        m /= p.tiles[i];
        valid = valid && (c < p.odim[i]);
        off_o += static_cast<long long>(c) * p.ostr[i];
        off_r1 += static_cast<long long>(c) * p.r1[i];
        off_r2 += static_cast<long long>(c) * p.r2[i];
        if (i == p.rb_dim) rb_row = c / p.rb_div;
      }
      const int col0_out = n_tile * ncols_out;
      __half* optr = p.out + off_o + col0_out;
      const __half* r1ptr = p.res1 ? p.res1 + off_r1 + col0_out : nullptr;
      const __half* r2ptr = p.res2 ? p.res2 + off_r2 + col0_out : nullptr;
      const float* rbptr = p.rowbias ? p.rowbias + static_cast<long long>(rb_row) * p.n_out_total + col0_out : nullptr;

      mbar_wait(&tfull_bar[as], aphase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + static_cast<uint32_t>(as * kAccStride);

      for (int c = 0; c < ncols_out; c += 16) {
        uint32_t r[16];
        float v[16];
        tmem_ld_32x32b_x16(taddr + c, r);
        if (geglu) {
          uint32_t g[16];
          tmem_ld_32x32b_x16(taddr + ncols_out + c, g);
          tmem_ld_wait();
          float gv[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            v[j] = __uint_as_float(r[j]);
            gv[j] = __uint_as_float(g[j]);
          }
          if (p.bias) {
            add_f32x16(v, p.bias + n_tile * p.bn + c);
            add_f32x16(gv, p.bias + n_tile * p.bn + ncols_out + c);
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] *= gelu_erf_f(gv[j]);
        } else {
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
          if (p.bias) add_f32x16(v, p.bias + n_tile * p.bn + c);
        }
        if (valid) {
          if (rbptr) add_f32x16(v, rbptr + c);
          if (do_silu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = silu_f(v[j]);
          }
          if (r1ptr) add_res16(v, r1ptr + c);
          if (r2ptr) add_res16(v, r2ptr + c);
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
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
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

int device_sm_count() {
  static int sms = -1;
  if (sms < 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = -1;
  }
  return sms;
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

  CUtensorMap tmA, tmB;
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
    cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)d->bn};
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
  p.idesc = umma_idesc_f16(kBlockM, d->bn);
  const int stage_bytes = kABytes + d->bn * kBlockK * 2;
  int stages = (200 * 1024) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  p.stages = stages;
  const int smem_bytes = stages * stage_bytes + 1024 + 512;

  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(tap_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  });
  if (attr_err != cudaSuccess) {
    set_last_error("ccedit_gemm: cudaFuncSetAttribute failed: %s", cudaGetErrorString(attr_err));
    return CCEDIT_ERR_CUDA;
  }
  const int sms = device_sm_count();
  if (sms <= 0) {
    set_last_error("ccedit_gemm: no CUDA device");
    return CCEDIT_ERR_CUDA;
  }
  const int grid = p.total_tiles < sms ? p.total_tiles : sms;
  tap_gemm_kernel<<<grid, kGemmThreads, smem_bytes, stream>>>(tmA, tmB, p);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  CCEDIT_CUDA_LAUNCH_CHECK("ccedit_gemm");
  return CCEDIT_OK;
}

}  // namespace ccedit

extern "C" int ccedit_gemm(const ccedit_gemm_desc* d, void* stream) {
  return ccedit::gemm_impl(d, static_cast<cudaStream_t>(stream));
}

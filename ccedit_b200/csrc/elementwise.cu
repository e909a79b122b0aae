// Small / layout kernels of the CCEdit hot path (all HBM- or latency-bound; 16-byte vectorised where the layout allows).
#include "common.cuh"
#include "../../include/ccedit_b200.h"

#include <atomic>

namespace ccedit {
extern std::atomic<long long> g_launch_count;

// [B][Cin][T][H][W] -> [B][T][H][W][Cpad] fp16, v*mul+add, zero padded channels. One thread per pixel.
template <typename SrcT>
__global__ void ncthw_to_cl_kernel(const SrcT* __restrict__ src, __half* __restrict__ dst, int Cin, int T, long long HW,
                                   int Cpad, float pre, float mul, float add, long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;  // over B*T*HW
  if (i >= total) return;
  const long long hw = i % HW;
  const long long bt = i / HW;
  const int t = static_cast<int>(bt % T);
  const long long b = bt / T;
  __half* d = dst + i * Cpad;
  for (int c0 = 0; c0 < Cpad; c0 += 8) {
    __align__(16) __half h[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      float v = 0.f;
      if (c < Cin) v = __fmaf_rn(__fadd_rn(static_cast<float>(src[((b * Cin + c) * T + t) * HW + hw]), pre), mul, add);
      h[j] = __float2half_rn(v);
    }
    *reinterpret_cast<uint4*>(d + c0) = *reinterpret_cast<const uint4*>(h);
  }
}

// UNet tail: dst[b][c][t][hw] = y + bias_t[c] + sum wt[c][c'][dt] * silu(y[t+dt-1][c'])   (Cout <= 8)
template <typename DstT>
__global__ void out_temporal_kernel(const __half* __restrict__ y, int ldy, const float* __restrict__ wt,
                                    const float* __restrict__ bias_t, DstT* __restrict__ dst, int Cout, int T,
                                    long long HW, long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;  // over B*T*HW
  if (i >= total) return;
  const long long hw = i % HW;
  const long long bt = i / HW;
  const int t = static_cast<int>(bt % T);
  const long long b = bt / T;
  float acc[8], cur[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = 0.f;
  for (int dt = 0; dt < 3; ++dt) {
    const int tt = t + dt - 1;
    if (tt < 0 || tt >= T) continue;
    const __half* yp = y + ((b * T + tt) * HW + hw) * ldy;
    float s[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float v = c < Cout ? __half2float(yp[c]) : 0.f;
      if (dt == 1) cur[c] = v;
      s[c] = silu_f(v);
    }
    for (int c = 0; c < Cout; ++c)
      for (int c2 = 0; c2 < Cout; ++c2) acc[c] += wt[(c * Cout + c2) * 3 + dt] * s[c2];
  }
  for (int c = 0; c < Cout; ++c)
    dst[((b * Cout + c) * T + t) * HW + hw] = static_cast<DstT>(cur[c] + bias_t[c] + acc[c]);
}

__global__ void timestep_embedding_kernel(const float* __restrict__ t, float* __restrict__ out, int B, int dim,
                                          float max_period) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim / 2;
  if (i >= B * half) return;
  const int b = i / half, k = i % half;
  const float freq = expf(-logf(max_period) * static_cast<float>(k) / static_cast<float>(half));
  const float arg = t[b] * freq;
  out[b * dim + k] = cosf(arg);
  out[b * dim + half + k] = sinf(arg);
  if ((dim & 1) && k == 0) out[b * dim + dim - 1] = 0.f;
}

// out[m][n] = act_out(sum_k act_in(x[m][k]) w[n][k] + b[n]); one warp per n, M <= 8.
__global__ void linear_small_kernel(const float* __restrict__ x, const __half* __restrict__ w,
                                    const float* __restrict__ bias, float* __restrict__ out, int M, int N, int K,
                                    int act_in, int act_out) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  float acc[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) acc[m] = 0.f;
  const __half* wr = w + static_cast<long long>(n) * K;
  for (int k0 = lane * 8; k0 < K; k0 += 256) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(wr + k0));
    const uint32_t ww[4] = {u.x, u.y, u.z, u.w};
    float wf[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&ww[j]));
      wf[2 * j] = f.x;
      wf[2 * j + 1] = f.y;
    }
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      if (m < M) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float xv = x[m * K + k0 + j];
          if (act_in) xv = xv / (1.f + expf(-xv));
          acc[m] = fmaf(xv, wf[j], acc[m]);
        }
      }
    }
  }
#pragma unroll
  for (int m = 0; m < 8; ++m) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[m] += __shfl_xor_sync(0xffffffffu, acc[m], o);
  }
  if (lane == 0) {
    for (int m = 0; m < M; ++m) {
      float v = acc[m] + (bias ? bias[n] : 0.f);
      if (act_out) v = v / (1.f + expf(-v));
      out[m * N + n] = v;
    }
  }
}

// [F][H][W][C] -> [F][4][H/2][W/2][C], plane = (h&1)*2 + (w&1); one thread per 16-byte vector
__global__ void parity_split_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int H, int W, int nvec,
                                    long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cv = static_cast<int>(i % nvec);
  long long r = i / nvec;
  const int w = static_cast<int>(r % W);
  r /= W;
  const int h = static_cast<int>(r % H);
  const long long f = r / H;
  const int plane = (h & 1) * 2 + (w & 1);
  const int H2 = H / 2, W2 = W / 2;
  y[(((f * 4 + plane) * H2 + (h >> 1)) * W2 + (w >> 1)) * nvec + cv] = __ldg(x + i);
}

// nearest x2: one thread per output vector
__global__ void upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int H, int W, int nvec,
                                  long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cv = static_cast<int>(i % nvec);
  long long r = i / nvec;
  const int wo = static_cast<int>(r % (2 * W));
  r /= (2 * W);
  const int ho = static_cast<int>(r % (2 * H));
  const long long f = r / (2 * H);
  y[i] = __ldg(x + ((f * H + (ho >> 1)) * W + (wo >> 1)) * nvec + cv);
}

__device__ __forceinline__ uint4 add8(const uint4& a, const uint4& b) {
  uint4 r;
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
  uint32_t rw[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&aw[j]));
    const float2 fb = __half22float2(*reinterpret_cast<const __half2*>(&bw[j]));
    __half2 h = __floats2half2_rn(fa.x + fb.x, fa.y + fb.y);
    rw[j] = *reinterpret_cast<uint32_t*>(&h);
  }
  r.x = rw[0];
  r.y = rw[1];
  r.z = rw[2];
  r.w = rw[3];
  return r;
}

__global__ void add_rows_kernel(const __half* __restrict__ a, long long lda, const __half* __restrict__ b, long long ldb,
                                __half* __restrict__ dst, long long ldd, int nvec, long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cv = static_cast<int>(i % nvec);
  const long long m = i / nvec;
  uint4 va = __ldg(reinterpret_cast<const uint4*>(a + m * lda) + cv);
  if (b) va = add8(va, __ldg(reinterpret_cast<const uint4*>(b + m * ldb) + cv));
  reinterpret_cast<uint4*>(dst + m * ldd)[cv] = va;
}

__global__ void add_center_frame_kernel(__half* __restrict__ x, const __half* __restrict__ y, int T, long long HWnvec,
                                        long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;  // over B*HW*nvec
  if (i >= total) return;
  const long long b = i / HWnvec, r = i % HWnvec;
  uint4* xp = reinterpret_cast<uint4*>(x) + (b * T + T / 2) * HWnvec + r;
  *xp = add8(*xp, __ldg(reinterpret_cast<const uint4*>(y) + i));
}

// fp32 -> fp16 cast of a flat buffer (the text context c["crossattn"], wrappers.py:164-166 casts to the model dtype)
__global__ void to_half_kernel(const float* __restrict__ src, __half* __restrict__ dst, long long n) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 f = __ldg(reinterpret_cast<const float4*>(src + i));
    __half2 a = __floats2half2_rn(f.x, f.y), b = __floats2half2_rn(f.z, f.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(dst + i) = u;
  } else {
    for (long long j = i; j < n; ++j) dst[j] = __float2half_rn(src[j]);
  }
}

static inline unsigned blocks_for(long long total, int threads) {
  return static_cast<unsigned>((total + threads - 1) / threads);
}

}  // namespace ccedit

using namespace ccedit;

extern "C" int ccedit_ncthw_to_cl(const void* src, int32_t src_f32, void* dst, int32_t B, int32_t Cin, int32_t T,
                                  int32_t H, int32_t W, int32_t Cpad, float pre, float mul, float add, void* stream) {
  CCEDIT_CHECK_ARG(src && dst, "ccedit_ncthw_to_cl: null pointer");
  CCEDIT_CHECK_ARG(B > 0 && Cin > 0 && T > 0 && H > 0 && W > 0 && Cpad >= Cin && Cpad % 8 == 0,
                   "ccedit_ncthw_to_cl: bad shape B=%d Cin=%d T=%d H=%d W=%d Cpad=%d", B, Cin, T, H, W, Cpad);
  const long long HW = static_cast<long long>(H) * W, total = static_cast<long long>(B) * T * HW;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (src_f32)
    ncthw_to_cl_kernel<float><<<blocks_for(total, 256), 256, 0, st>>>(static_cast<const float*>(src),
                                                                      static_cast<__half*>(dst), Cin, T, HW, Cpad, pre,
                                                                      mul, add, total);
  else
    ncthw_to_cl_kernel<__half><<<blocks_for(total, 256), 256, 0, st>>>(static_cast<const __half*>(src),
                                                                       static_cast<__half*>(dst), Cin, T, HW, Cpad, pre,
                                                                       mul, add, total);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  CCEDIT_CUDA_LAUNCH_CHECK("ccedit_ncthw_to_cl");
  return CCEDIT_OK;
}

extern "C" int ccedit_out_temporal(const void* y, int32_t ldy, const float* wt, const float* bias_t, void* dst,
                                   int32_t dst_f32, int32_t B, int32_t Cout, int32_t T, int32_t HW, void* stream) {
  CCEDIT_CHECK_ARG(y && wt && bias_t && dst, "ccedit_out_temporal: null pointer");
  CCEDIT_CHECK_ARG(B > 0 && Cout > 0 && Cout <= 8 && T > 0 && HW > 0 && ldy >= Cout,
                   "ccedit_out_temporal: bad shape B=%d Cout=%d T=%d HW=%d ldy=%d", B, Cout, T, HW, ldy);
  const long long total = static_cast<long long>(B) * T * HW;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dst_f32)
    out_temporal_kernel<float><<<blocks_for(total, 256), 256, 0, st>>>(static_cast<const __half*>(y), ldy, wt, bias_t,
                                                                       static_cast<float*>(dst), Cout, T, HW, total);
  else
    out_temporal_kernel<__half><<<blocks_for(total, 256), 256, 0, st>>>(static_cast<const __half*>(y), ldy, wt, bias_t,
                                                                        static_cast<__half*>(dst), Cout, T, HW, total);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  CCEDIT_CUDA_LAUNCH_CHECK("ccedit_out_temporal");
  return CCEDIT_OK;
}

extern "C" int ccedit_timestep_embedding(const float* t, float* out, int32_t B, int32_t dim, float max_period,
                                         void* stream) {
  CCEDIT_CHECK_ARG(t && out && B > 0 && dim >= 2, "ccedit_timestep_embedding: bad arguments");
  const int total = B * (dim / 2);
  timestep_embedding_kernel<<<blocks_for(total, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(t, out, B, dim,
                                                                                                   max_period);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  CCEDIT_CUDA_LAUNCH_CHECK("ccedit_timestep_embedding");
  return CCEDIT_OK;
}

extern "C" int ccedit_linear_small(const float* x, const void* w, const float* b, float* out, int32_t M, int32_t N,
                                   int32_t K, int32_t act_in, int32_t act_out, void* stream) {
  CCEDIT_CHECK_ARG(x && w && out, "ccedit_linear_small: null pointer");
  CCEDIT_CHECK_ARG(M >= 1 && M <= 8 && N > 0 && K > 0 && K % 8 == 0, "ccedit_linear_small: bad shape M=%d N=%d K=%d (M<=8, K%%8==0)", M, N, K);
  const int wpb = 4;
  linear_small_kernel<<<(N + wpb - 1) / wpb, wpb * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      x, static_cast<const __half*>(w), b, out, M, N, K, act_in, act_out);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  CCEDIT_CUDA_LAUNCH_CHECK("ccedit_linear_small");
  return CCEDIT_OK;
}

extern "C" int ccedit_parity_split(const void* x, void* y, int32_t F, int32_t H, int32_t W, int32_t C, void* stream) {
  CCEDIT_CHECK_ARG(x && y && F > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C > 0 && C % 8 == 0,
                   "ccedit_parity_split: bad shape F=%d H=%d W=%d C=%d (H, W even; C%%8==0)", F, H, W, C);
  const int nvec = C / 8;
  const long long total = static_cast<long long>(F) * H * W * nvec;
  parity_split_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), static_cast<uint4*>(y), H, W, nvec, total);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  CCEDIT_CUDA_LAUNCH_CHECK("ccedit_parity_split");
  return CCEDIT_OK;
}

extern "C" int ccedit_upsample_nearest2x(const void* x, void* y, int32_t F, int32_t H, int32_t W, int32_t C,
                                         void* stream) {
  CCEDIT_CHECK_ARG(x && y && F > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0,
                   "ccedit_upsample_nearest2x: bad shape F=%d H=%d W=%d C=%d", F, H, W, C);
  const int nvec = C / 8;
  const long long total = static_cast<long long>(F) * H * W * 4 * nvec;
  upsample2x_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), static_cast<uint4*>(y), H, W, nvec, total);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  CCEDIT_CUDA_LAUNCH_CHECK("ccedit_upsample_nearest2x");
  return CCEDIT_OK;
}

extern "C" int ccedit_add_rows(const void* a, int64_t lda, const void* b, int64_t ldb, void* dst, int64_t ldd,
                               int64_t M, int32_t C, void* stream) {
  CCEDIT_CHECK_ARG(a && dst && M > 0 && C > 0 && C % 8 == 0 && lda % 8 == 0 && ldd % 8 == 0 && (!b || ldb % 8 == 0),
                   "ccedit_add_rows: bad arguments M=%lld C=%d", (long long)M, C);
  const int nvec = C / 8;
  const long long total = M * nvec;
  add_rows_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(a), lda, static_cast<const __half*>(b), ldb, static_cast<__half*>(dst), ldd, nvec,
      total);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  CCEDIT_CUDA_LAUNCH_CHECK("ccedit_add_rows");
  return CCEDIT_OK;
}

extern "C" int ccedit_add_center_frame(void* x, const void* y, int32_t B, int32_t T, int32_t HW, int32_t C,
                                       void* stream) {
  CCEDIT_CHECK_ARG(x && y && B > 0 && T > 0 && HW > 0 && C > 0 && C % 8 == 0, "ccedit_add_center_frame: bad arguments");
  const long long HWnvec = static_cast<long long>(HW) * (C / 8), total = B * HWnvec;
  add_center_frame_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<__half*>(x), static_cast<const __half*>(y), T, HWnvec, total);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  CCEDIT_CUDA_LAUNCH_CHECK("ccedit_add_center_frame");
  return CCEDIT_OK;
}


extern "C" int ccedit_to_half(const float* src, void* dst, int64_t n, void* stream) {
  CCEDIT_CHECK_ARG(src && dst && n > 0, "ccedit_to_half: bad arguments");
  CCEDIT_CHECK_ARG((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0,
                   "ccedit_to_half: src must be 16-byte and dst 8-byte aligned");
  to_half_kernel<<<blocks_for((n + 3) / 4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, static_cast<__half*>(dst), n);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  CCEDIT_CUDA_LAUNCH_CHECK("ccedit_to_half");
  return CCEDIT_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// First-stage (VAE) helpers
// ---------------------------------------------------------------------------------------------------------------
namespace ccedit {

// In-place softmax over the N columns of every row of a row-major fp16 matrix (fp32 arithmetic): the middle step of the
// first-stage decoder's single-head d = 512 attention (model.py:161-201), which runs as GEMM -> softmax -> GEMM because
// one head of width 512 does not fit the flash kernel's TMEM budget.  One CTA per row, the row lives in registers.
template <int VPT>      // 16-byte vectors per thread
__global__ void __launch_bounds__(256) softmax_rows_kernel(__half* __restrict__ x, long long ld, int N) {
  __shared__ float red[8];
  __shared__ float bcast;
  __half* row = x + static_cast<long long>(blockIdx.x) * ld;
  const int nvec = N >> 3;
  float v[VPT][8];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int iv = threadIdx.x + 256 * i;
    if (iv < nvec) {
      const uint4 u = *reinterpret_cast<const uint4*>(row + iv * 8);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
        v[i][2 * j] = f.x;
        v[i][2 * j + 1] = f.y;
        m = fmaxf(m, fmaxf(f.x, f.y));
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = red[0];
    for (int k = 1; k < 8; ++k) t = fmaxf(t, red[k]);
    bcast = t;
  }
  __syncthreads();
  m = bcast;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    if (threadIdx.x + 256 * i < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[i][j] = __expf(v[i][j] - m);
        s += v[i][j];
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __syncthreads();
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[k];             // fixed order: deterministic
    bcast = 1.f / t;
  }
  __syncthreads();
  const float inv = bcast;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int iv = threadIdx.x + 256 * i;
    if (iv < nvec) {
      uint4 u;
      __half2 h0 = __floats2half2_rn(v[i][0] * inv, v[i][1] * inv), h1 = __floats2half2_rn(v[i][2] * inv, v[i][3] * inv);
      __half2 h2 = __floats2half2_rn(v[i][4] * inv, v[i][5] * inv), h3 = __floats2half2_rn(v[i][6] * inv, v[i][7] * inv);
      u.x = *reinterpret_cast<uint32_t*>(&h0);
      u.y = *reinterpret_cast<uint32_t*>(&h1);
      u.z = *reinterpret_cast<uint32_t*>(&h2);
      u.w = *reinterpret_cast<uint32_t*>(&h3);
      *reinterpret_cast<uint4*>(row + iv * 8) = u;
    }
  }
}

// channels-last fp16 [B][T][HW][ld] (first C channels) -> [B][C][T][HW] fp32 / fp16: the decoder's image output
template <typename DstT>
__global__ void cl_to_ncthw_kernel(const __half* __restrict__ src, int ld, DstT* __restrict__ dst, int C, int T, long long HW,
                                   long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;   // over B*T*HW
  if (i >= total) return;
  const long long hw = i % HW, bt = i / HW;
  const int t = static_cast<int>(bt % T);
  const long long b = bt / T;
  const __half* s = src + i * ld;
  for (int c = 0; c < C; ++c) dst[((b * C + c) * T + t) * HW + hw] = static_cast<DstT>(__half2float(s[c]));
}

}  // namespace ccedit

extern "C" int ccedit_softmax_rows(void* x, int64_t ld, int64_t M, int32_t N, void* stream) {
  CCEDIT_CHECK_ARG(x && M > 0 && N > 0 && N % 8 == 0 && ld % 8 == 0 && ld >= N && N <= 8 * 256 * 8,
                   "ccedit_softmax_rows: bad shape M=%lld N=%d ld=%lld (N %% 8 == 0, N <= 16384)", (long long)M, N, (long long)ld);
  CCEDIT_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0, "ccedit_softmax_rows: x must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __half* xp = static_cast<__half*>(x);
  const unsigned grid = static_cast<unsigned>(M);
  const int vpt = (N / 8 + 255) / 256;
  switch (vpt) {
    case 1: softmax_rows_kernel<1><<<grid, 256, 0, st>>>(xp, ld, N); break;
    case 2: softmax_rows_kernel<2><<<grid, 256, 0, st>>>(xp, ld, N); break;
    case 3: softmax_rows_kernel<3><<<grid, 256, 0, st>>>(xp, ld, N); break;
    case 4: softmax_rows_kernel<4><<<grid, 256, 0, st>>>(xp, ld, N); break;
    default: softmax_rows_kernel<8><<<grid, 256, 0, st>>>(xp, ld, N); break;
  }
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  CCEDIT_CUDA_LAUNCH_CHECK("ccedit_softmax_rows");
  return CCEDIT_OK;
}

extern "C" int ccedit_cl_to_ncthw(const void* src, int32_t ld, void* dst, int32_t dst_f32, int32_t B, int32_t C, int32_t T,
                                  int64_t HW, void* stream) {
  CCEDIT_CHECK_ARG(src && dst && B > 0 && C > 0 && T > 0 && HW > 0 && ld >= C, "ccedit_cl_to_ncthw: bad arguments");
  const long long total = static_cast<long long>(B) * T * HW;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dst_f32)
    cl_to_ncthw_kernel<float><<<blocks_for(total, 256), 256, 0, st>>>(static_cast<const __half*>(src), ld, static_cast<float*>(dst),
                                                                      C, T, HW, total);
  else
    cl_to_ncthw_kernel<__half><<<blocks_for(total, 256), 256, 0, st>>>(static_cast<const __half*>(src), ld,
                                                                       static_cast<__half*>(dst), C, T, HW, total);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  CCEDIT_CUDA_LAUNCH_CHECK("ccedit_cl_to_ncthw");
  return CCEDIT_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Text conditioner (CLIP-L text transformer): the three operations the tap-GEMM / LayerNorm kernels do
// not cover.  All tiny (77 tokens per prompt, once per clip): written for clarity, not for a roofline.
// ---------------------------------------------------------------------------------------------------------------
namespace ccedit {

// out[b*L + l][:] = token_emb[ids[b][l]][:] + pos_emb[l][:]   (fp16 tables, fp32 add, fp16 out); D % 8 == 0
__global__ void embed_tokens_kernel(const long long* __restrict__ ids, const __half* __restrict__ tok,
                                    const __half* __restrict__ pos, __half* __restrict__ out, int L, int D, int V, long long rows) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;   // over rows * D/8
  const int nvec = D >> 3;
  if (i >= rows * nvec) return;
  const long long r = i / nvec;
  const int cv = static_cast<int>(i % nvec);
  long long id = ids[r];
  id = id < 0 ? 0 : (id >= V ? V - 1 : id);
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(tok + id * D) + cv);
  const uint4 b = __ldg(reinterpret_cast<const uint4*>(pos + (r % L) * D) + cv);
  reinterpret_cast<uint4*>(out + r * D)[cv] = add8(a, b);
}

// x = x * sigmoid(1.702 x) in place (CLIP's quick_gelu), fp16 storage, fp32 math
__global__ void quick_gelu_kernel(__half* __restrict__ x, long long n8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  uint4 u = reinterpret_cast<uint4*>(x)[i];
  uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
    f.x = __fdividef(f.x, 1.0f + __expf(-1.702f * f.x));
    f.y = __fdividef(f.y, 1.0f + __expf(-1.702f * f.y));
    const __half2 h = __floats2half2_rn(f.x, f.y);
    w[j] = *reinterpret_cast<const uint32_t*>(&h);
  }
  reinterpret_cast<uint4*>(x)[i] = make_uint4(w[0], w[1], w[2], w[3]);
}

// Causal self-attention over a short sequence (L <= 128 tokens, head dim 64): one CTA per (batch entry, head), K and V of
// the head staged in shared memory, one warp per query row (round-robin): lane j scores keys j, j+32, j+64, j+96 (j <= i),
// warp-shuffle softmax, every lane accumulates two output channels.
__global__ void __launch_bounds__(128) causal_attn_small_kernel(const __half* __restrict__ q, const __half* __restrict__ k,
                                                                const __half* __restrict__ v, __half* __restrict__ o,
                                                                long long ld, long long ldo, int L, int heads, float scale) {
  constexpr int D = 64;
  extern __shared__ __half ca_sm[];                 // K [L][D+2], V [L][D+2] (padded rows: conflict-free column reads), P [4][128]
  __half* sK = ca_sm;
  __half* sV = ca_sm + L * (D + 2);
  float* sP = reinterpret_cast<float*>(ca_sm + 2 * L * (D + 2));
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long base = static_cast<long long>(b) * L;
  for (int idx = threadIdx.x; idx < L * (D / 2); idx += blockDim.x) {
    const int r = idx / (D / 2), c2 = idx % (D / 2);
    reinterpret_cast<__half2*>(sK + r * (D + 2))[c2] = reinterpret_cast<const __half2*>(k + (base + r) * ld + h * D)[c2];
    reinterpret_cast<__half2*>(sV + r * (D + 2))[c2] = reinterpret_cast<const __half2*>(v + (base + r) * ld + h * D)[c2];
  }
  __syncthreads();
  float* myP = sP + warp * 128;
  for (int i = warp; i < L; i += 4) {
    const __half2* qr = reinterpret_cast<const __half2*>(q + (base + i) * ld + h * D);
    float2 qv[D / 2];
#pragma unroll
    for (int c = 0; c < D / 2; ++c) qv[c] = __half22float2(__ldg(qr + c));
    float s[4];
    float m = -INFINITY;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int j = lane + 32 * t;
      s[t] = -INFINITY;
      if (j <= i && j < L) {
        const __half2* kr = reinterpret_cast<const __half2*>(sK + j * (D + 2));
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < D / 2; ++c) {
          const float2 kv = __half22float2(kr[c]);
          acc = fmaf(qv[c].x, kv.x, acc);
          acc = fmaf(qv[c].y, kv.y, acc);
        }
        s[t] = acc * scale;
      }
      m = fmaxf(m, s[t]);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float p = s[t] == -INFINITY ? 0.f : __expf(s[t] - m);
      myP[lane + 32 * t] = p;
      sum += p;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j <= i; ++j) {
      const float p = myP[j];
      const float2 vv = __half22float2(reinterpret_cast<const __half2*>(sV + j * (D + 2))[lane]);
      o0 = fmaf(p, vv.x, o0);
      o1 = fmaf(p, vv.y, o1);
    }
    const float inv = 1.f / sum;
    reinterpret_cast<__half2*>(o + (base + i) * ldo + h * D)[lane] = __floats2half2_rn(o0 * inv, o1 * inv);
    __syncwarp();
  }
}

}  // namespace ccedit

extern "C" int ccedit_embed_tokens(const int64_t* ids, const void* tok, const void* pos, void* out, int32_t B, int32_t L,
                                   int32_t D, int32_t V, void* stream) {
  CCEDIT_CHECK_ARG(ids && tok && pos && out && B > 0 && L > 0 && D > 0 && D % 8 == 0 && V > 0, "ccedit_embed_tokens: bad arguments");
  const long long rows = static_cast<long long>(B) * L, total = rows * (D / 8);
  embed_tokens_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(ids), static_cast<const __half*>(tok), static_cast<const __half*>(pos),
      static_cast<__half*>(out), L, D, V, rows);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  CCEDIT_CUDA_LAUNCH_CHECK("ccedit_embed_tokens");
  return CCEDIT_OK;
}

extern "C" int ccedit_quick_gelu(void* x, int64_t n, void* stream) {
  CCEDIT_CHECK_ARG(x && n > 0 && n % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "ccedit_quick_gelu: bad arguments");
  quick_gelu_kernel<<<blocks_for(n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<__half*>(x), n / 8);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  CCEDIT_CUDA_LAUNCH_CHECK("ccedit_quick_gelu");
  return CCEDIT_OK;
}

extern "C" int ccedit_causal_attention_small(const void* q, const void* k, const void* v, int64_t ld, void* o, int64_t ldo,
                                             int32_t B, int32_t L, int32_t heads, int32_t d, float scale, void* stream) {
  CCEDIT_CHECK_ARG(q && k && v && o && B > 0 && L > 0 && L <= 128 && heads > 0 && d == 64 && ld % 2 == 0 && ldo % 2 == 0,
                   "ccedit_causal_attention_small: bad arguments (L <= 128, d == 64)");
  const size_t smem = static_cast<size_t>(2) * L * (64 + 2) * sizeof(__half) + 4 * 128 * sizeof(float);
  causal_attn_small_kernel<<<B * heads, 128, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(q), static_cast<const __half*>(k), static_cast<const __half*>(v), static_cast<__half*>(o), ld,
      ldo, L, heads, scale);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  CCEDIT_CUDA_LAUNCH_CHECK("ccedit_causal_attention_small");
  return CCEDIT_OK;
}

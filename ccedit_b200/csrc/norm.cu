// GroupNorm (spatial: per frame; temporal: per pixel over T) and LayerNorm, fp16 in/out, fp32 statistics.
// All three are HBM-bound: one read + one write of the activation (the spatial flavour reads twice: statistics, apply).
// Reference call sites: util.py:296-302 (eps 1e-5), attention.py:153-156 (eps 1e-6), attention.py:667-669 (LayerNorm);
// the temporal reduction shape comes from openaimodel.py:157 ("(b h w) c t" => statistics over (C/32, T) per pixel).
#include "common.cuh"
#include "../../include/ccedit_b200.h"

#include <atomic>
#include <cstdlib>

namespace ccedit {
extern std::atomic<long long> g_launch_count;
int device_sm_count();

constexpr int kGroups = 32;
constexpr int kMaxSplit = 32;
constexpr int kGnCounterFloats = 8192;       // head of the spatial-GroupNorm scratch: per-frame arrival counters (zeroed once)

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  __half2 h0 = __floats2half2_rn(f[0], f[1]), h1 = __floats2half2_rn(f[2], f[3]);
  __half2 h2 = __floats2half2_rn(f[4], f[5]), h3 = __floats2half2_rn(f[6], f[7]);
  u.x = *reinterpret_cast<uint32_t*>(&h0);
  u.y = *reinterpret_cast<uint32_t*>(&h1);
  u.z = *reinterpret_cast<uint32_t*>(&h2);
  u.w = *reinterpret_cast<uint32_t*>(&h3);
  return u;
}

// Deterministic group reduction: every thread parks its 8 per-channel (sum, sumsq) in shared memory
// ([slot][C] each), then one thread per (slot-set, group) adds them in a fixed order.  (Float atomics would make the
// statistics - and through fp16 re-rounding the whole network output - differ from run to run.)
__device__ __forceinline__ void park_channels(const float (&s)[8], const float (&q)[8], float* ssum, float* ssq, int slot,
                                              int C, int c0) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ssum[slot * C + c0 + j] = s[j];
    ssq[slot * C + c0 + j] = q[j];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// spatial GroupNorm: statistics
// grid (nsplit, F); block = nvec * rpi threads (nvec = C/8 vectors per row, rpi rows per iteration)
// ---------------------------------------------------------------------------------------------------------------
__global__ void gn_spatial_stats_kernel(const __half* __restrict__ x, float* __restrict__ partial, int HW, int C,
                                        int nvec, int rpi, int nsplit) {
  extern __shared__ float gn_sm[];            // [2][rpi][C]
  float* ssum = gn_sm;
  float* ssq = gn_sm + rpi * C;
  const int f = blockIdx.y, split = blockIdx.x;
  const int cv = threadIdx.x % nvec, r0 = threadIdx.x / nvec;
  const int rows_per_split = (HW + nsplit - 1) / nsplit;
  const int row_begin = split * rows_per_split;
  const int row_end = min(HW, row_begin + rows_per_split);
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  const uint4* base = reinterpret_cast<const uint4*>(x + static_cast<size_t>(f) * HW * C) + cv;
  int r = row_begin + r0;
  for (; r + 3 * rpi < row_end; r += 4 * rpi) {          // four independent 16-byte loads in flight per thread
    uint4 u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) u[k] = __ldg(base + static_cast<size_t>(r + k * rpi) * nvec);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float v[8];
      unpack8(u[k], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += v[j];
        q[j] += v[j] * v[j];
      }
    }
  }
  for (; r < row_end; r += rpi) {
    float v[8];
    unpack8(__ldg(base + static_cast<size_t>(r) * nvec), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j] += v[j];
      q[j] += v[j] * v[j];
    }
  }
  park_channels(s, q, ssum, ssq, r0, C, cv * 8);
  __syncthreads();
  if (threadIdx.x < kGroups) {
    const int cpg = C / kGroups, g = threadIdx.x;
    float ts = 0.f, tq = 0.f;
    for (int r = 0; r < rpi; ++r)
      for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
        ts += ssum[r * C + c];
        tq += ssq[r * C + c];
      }
    float* pp = partial + (static_cast<size_t>(f) * nsplit + split) * 2 * kGroups;
    pp[2 * g] = ts;
    pp[2 * g + 1] = tq;
  }
}

// spatial GroupNorm: apply (+ optional SiLU). grid (nchunk, F)
__global__ void gn_spatial_apply_kernel(const __half* __restrict__ x, __half* __restrict__ y,
                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                        const float* __restrict__ partial, int HW, int C, int nvec, int rpi, int nsplit,
                                        float eps, int silu) {
  __shared__ float smean[kGroups], srstd[kGroups];
  const int f = blockIdx.y;
  if (threadIdx.x < kGroups) {
    float s = 0.f, q = 0.f;
    for (int k = 0; k < nsplit; ++k) {
      const float* pp = partial + (static_cast<size_t>(f) * nsplit + k) * 2 * kGroups;
      s += pp[2 * threadIdx.x];
      q += pp[2 * threadIdx.x + 1];
    }
    const float n = static_cast<float>(HW) * static_cast<float>(C / kGroups);
    const float mean = s / n;
    const float var = fmaxf(q / n - mean * mean, 0.f);
    smean[threadIdx.x] = mean;
    srstd[threadIdx.x] = rsqrtf(var + eps);
  }
  __syncthreads();
  const int cv = threadIdx.x % nvec, r0 = threadIdx.x / nvec;
  const int cpg = C / kGroups;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cv * 8 + j;
    const int g = c / cpg;
    const float a = srstd[g] * gamma[c];
    sc[j] = a;
    sh[j] = beta[c] - smean[g] * a;
  }
  const int rows_per_chunk = (HW + gridDim.x - 1) / gridDim.x;
  const int row_begin = blockIdx.x * rows_per_chunk;
  const int row_end = min(HW, row_begin + rows_per_chunk);
  const uint4* xb = reinterpret_cast<const uint4*>(x + static_cast<size_t>(f) * HW * C) + cv;
  uint4* yb = reinterpret_cast<uint4*>(y + static_cast<size_t>(f) * HW * C) + cv;
  int r = row_begin + r0;
  for (; r + 3 * rpi < row_end; r += 4 * rpi) {          // four independent 16-byte loads in flight per thread
    uint4 u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) u[k] = __ldg(xb + static_cast<size_t>(r + k * rpi) * nvec);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float v[8];
      unpack8(u[k], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = v[j] * sc[j] + sh[j];
        v[j] = silu ? silu_f(t) : t;
      }
      yb[static_cast<size_t>(r + k * rpi) * nvec] = pack8(v);
    }
  }
  for (; r < row_end; r += rpi) {
    float v[8];
    unpack8(__ldg(xb + static_cast<size_t>(r) * nvec), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = v[j] * sc[j] + sh[j];
      v[j] = silu ? silu_f(t) : t;
    }
    yb[static_cast<size_t>(r) * nvec] = pack8(v);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// spatial GroupNorm, single pass over HBM: statistics and apply in ONE kernel.  grid (nsplit, F) with every CTA resident
// at the same time (the host sizes nsplit from the occupancy); CTA (s, f) reduces its slice of frame f, publishes the
// partial sums, waits on a per-frame arrival counter until the nsplit slices of the frame are in, and then normalises
// the SAME slice - walking it backwards, so that the rows it read last (still in L2) are re-read first.  The two-kernel
// version above reads the tensor twice from HBM at the big levels (134 MB per activation > L2 once all frames are in
// flight); here the second read is an L2 hit for most of the slice.  Statistics stay deterministic (fixed-order sums).
// counters: [F][2] ints, zero before the first launch; the last CTA to leave a frame's barrier resets them.
// Co-residency is GUARANTEED, not assumed: the kernel is launched cooperatively (cudaLaunchAttributeCooperative), so
// the driver schedules the grid only when every CTA can be resident at once - whatever else runs on the device (other
// streams, NCCL kernels, MPS neighbours) - and refuses the launch (-> two-kernel version) if the grid cannot fit at all.
// The bounded spin remains as a last line of defence.  CCEDIT_GN_FUSED=0 selects the two-kernel version.
// ---------------------------------------------------------------------------------------------------------------
__global__ void gn_spatial_fused_kernel(const __half* __restrict__ x, __half* __restrict__ y,
                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                        float* __restrict__ partial, int* __restrict__ counters, int HW, int C, int nvec,
                                        int rpi, float eps, int silu) {
  extern __shared__ float gn_sm[];            // [2][rpi][C]
  __shared__ float smean[kGroups], srstd[kGroups];
  float* ssum = gn_sm;
  float* ssq = gn_sm + rpi * C;
  griddep_wait();                  // PDL: x comes from the previous kernel of the stream
  griddep_launch_dependents();
  const int f = blockIdx.y, split = blockIdx.x, nsplit = gridDim.x;
  const int cv = threadIdx.x % nvec, r0 = threadIdx.x / nvec;
  const int rows_per_split = (HW + nsplit - 1) / nsplit;
  const int row_begin = split * rows_per_split;
  const int row_end = min(HW, row_begin + rows_per_split);
  const uint4* xb = reinterpret_cast<const uint4*>(x + static_cast<size_t>(f) * HW * C) + cv;
  {
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
    int r = row_begin + r0;
    for (; r + 3 * rpi < row_end; r += 4 * rpi) {
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = __ldg(xb + static_cast<size_t>(r + k * rpi) * nvec);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float v[8];
        unpack8(u[k], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s[j] += v[j];
          q[j] += v[j] * v[j];
        }
      }
    }
    for (; r < row_end; r += rpi) {
      float v[8];
      unpack8(__ldg(xb + static_cast<size_t>(r) * nvec), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += v[j];
        q[j] += v[j] * v[j];
      }
    }
    park_channels(s, q, ssum, ssq, r0, C, cv * 8);
  }
  __syncthreads();
  if (threadIdx.x < kGroups) {
    const int cpg = C / kGroups, g = threadIdx.x;
    float ts = 0.f, tq = 0.f;
    for (int r = 0; r < rpi; ++r)
      for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
        ts += ssum[r * C + c];
        tq += ssq[r * C + c];
      }
    float* pp = partial + (static_cast<size_t>(f) * nsplit + split) * 2 * kGroups;
    pp[2 * g] = ts;
    pp[2 * g + 1] = tq;
    __threadfence();                                       // the partial sums are visible before the arrival below
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(&counters[2 * f], 1);
    unsigned spins = 0;
    while (atomicAdd(&counters[2 * f], 0) < nsplit) {      // all slices of this frame
      __nanosleep(64);
      if (++spins > (1u << 24)) __trap();                  // a scheduling assumption broke: fail instead of hanging
    }
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x < kGroups) {
    float s = 0.f, q = 0.f;
    for (int k = 0; k < nsplit; ++k) {
      const float* pp = partial + (static_cast<size_t>(f) * nsplit + k) * 2 * kGroups;
      s += __ldcg(pp + 2 * threadIdx.x);
      q += __ldcg(pp + 2 * threadIdx.x + 1);
    }
    const float n = static_cast<float>(HW) * static_cast<float>(C / kGroups);
    const float mean = s / n;
    const float var = fmaxf(q / n - mean * mean, 0.f);
    smean[threadIdx.x] = mean;
    srstd[threadIdx.x] = rsqrtf(var + eps);
  }
  __syncthreads();
  if (threadIdx.x == 0) {                                  // leave the barrier; the last one out re-arms it
    if (atomicAdd(&counters[2 * f + 1], 1) == nsplit - 1) {
      counters[2 * f] = 0;
      counters[2 * f + 1] = 0;
    }
  }
  const int cpg = C / kGroups;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cv * 8 + j;
    const int g = c / cpg;
    const float a = srstd[g] * gamma[c];
    sc[j] = a;
    sh[j] = beta[c] - smean[g] * a;
  }
  uint4* yb = reinterpret_cast<uint4*>(y + static_cast<size_t>(f) * HW * C) + cv;
  // backwards over the slice, four rows in flight per thread
  int r = row_end - 1 - r0;
  for (; r - 3 * rpi >= row_begin; r -= 4 * rpi) {
    uint4 u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) u[k] = __ldlu(xb + static_cast<size_t>(r - k * rpi) * nvec);   // last use: the line may leave L2
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float v[8];
      unpack8(u[k], v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = v[j] * sc[j] + sh[j];
        v[j] = silu ? silu_f(t) : t;
      }
      __stcs(yb + static_cast<size_t>(r - k * rpi) * nvec, pack8(v));   // streaming store: y must not push the slices still to be re-read out of L2
    }
  }
  for (; r >= row_begin; r -= rpi) {
    float v[8];
    unpack8(__ldlu(xb + static_cast<size_t>(r) * nvec), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = v[j] * sc[j] + sh[j];
      v[j] = silu ? silu_f(t) : t;
    }
    __stcs(yb + static_cast<size_t>(r) * nvec, pack8(v));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// temporal GroupNorm: one block handles `ppb` pixels; statistics over (C/32, T) per pixel.
// x: [B][T][HW][C]; second pass re-reads the block's rows (L1/L2 resident) instead of holding T vectors in registers.
// ---------------------------------------------------------------------------------------------------------------
__global__ void gn_temporal_kernel(const __half* __restrict__ x, __half* __restrict__ y,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, int B, int T, int HW,
                                   int C, int nvec, int ppb, float eps, int silu) {
  extern __shared__ float sg_dyn[];  // [2][ppb][C] per-channel sums, then [ppb][64] (mean, rstd) per group
  float* ssum = sg_dyn;
  float* ssq = sg_dyn + ppb * C;
  float* sstat = sg_dyn + 2 * ppb * C;
  griddep_wait();                  // PDL: x comes from the previous kernel of the stream
  griddep_launch_dependents();
  const int cv = threadIdx.x % nvec, pl = threadIdx.x / nvec;
  const long long pix = static_cast<long long>(blockIdx.x) * ppb + pl;  // over B*HW
  const bool active = pix < static_cast<long long>(B) * HW;
  const int b = active ? static_cast<int>(pix / HW) : 0;
  const int hw = active ? static_cast<int>(pix % HW) : 0;
  const size_t tstride = static_cast<size_t>(HW) * nvec;  // in uint4
  const uint4* xb = reinterpret_cast<const uint4*>(x) + (static_cast<size_t>(b) * T * HW + hw) * nvec + cv;
  uint4* yb = reinterpret_cast<uint4*>(y) + (static_cast<size_t>(b) * T * HW + hw) * nvec + cv;
  const int cpg = C / kGroups;
  {
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
    if (active) {
      int t = 0;
      for (; t + 3 < T; t += 4) {                          // four frames in flight per thread
        uint4 u[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) u[k] = __ldg(xb + (t + k) * tstride);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float v[8];
          unpack8(u[k], v);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            s[j] += v[j];
            q[j] += v[j] * v[j];
          }
        }
      }
      for (; t < T; ++t) {
        float v[8];
        unpack8(__ldg(xb + t * tstride), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s[j] += v[j];
          q[j] += v[j] * v[j];
        }
      }
    }
    park_channels(s, q, ssum, ssq, pl, C, cv * 8);
  }
  __syncthreads();
  const float n = static_cast<float>(T) * static_cast<float>(cpg);
  for (int idx = threadIdx.x; idx < ppb * kGroups; idx += blockDim.x) {
    const int p = idx / kGroups, g = idx % kGroups;
    float ts = 0.f, tq = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      ts += ssum[p * C + c];
      tq += ssq[p * C + c];
    }
    const float mean = ts / n;
    const float var = fmaxf(tq / n - mean * mean, 0.f);
    sstat[p * 2 * kGroups + 2 * g] = mean;
    sstat[p * 2 * kGroups + 2 * g + 1] = rsqrtf(var + eps);
  }
  __syncthreads();
  if (active) {
    const float* sg = sstat + pl * 2 * kGroups;
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cv * 8 + j;
      const int g = c / cpg;
      const float a = sg[2 * g + 1] * gamma[c];
      sc[j] = a;
      sh[j] = beta[c] - sg[2 * g] * a;
    }
    int t = 0;
    for (; t + 3 < T; t += 4) {
      uint4 u4[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u4[k] = __ldg(xb + (t + k) * tstride);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float v[8];
        unpack8(u4[k], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float u = v[j] * sc[j] + sh[j];
          v[j] = silu ? silu_f(u) : u;
        }
        __stcs(yb + (t + k) * tstride, pack8(v));        // streaming: y is not read again by this kernel
      }
    }
    for (; t < T; ++t) {
      float v[8];
      unpack8(__ldg(xb + t * tstride), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float u = v[j] * sc[j] + sh[j];
        v[j] = silu ? silu_f(u) : u;
      }
      __stcs(yb + t * tstride, pack8(v));
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm: one warp per token row, two-pass statistics on register-resident data.  Templated on the number of
// 16-byte vectors per lane (C <= 256 * NV) so that the narrow rows of the big levels (C = 320: NV = 2) do not pay the
// register footprint of the widest ones - occupancy is what hides the HBM latency of this kernel - and every warp
// works on two rows at a time to double the bytes in flight.
// ---------------------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256) layernorm_kernel(const __half* __restrict__ x, long long ldx, __half* __restrict__ y,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        long long M, int C, float eps, float2* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const long long row0 = (static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 2;
  if (row0 >= M) return;
  const int nvec = C >> 3;
  const bool two = row0 + 1 < M;
  float v[2][NV][8];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + (row0 + (r && two ? 1 : 0)) * ldx);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int iv = lane + 32 * i;
      if (iv < nvec) unpack8(__ldg(xr + iv), v[r][i]);
    }
  }
  float mean[2], rstd[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (lane + 32 * i < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[r][i][j];
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    mean[r] = s / static_cast<float>(C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (lane + 32 * i < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float dlt = v[r][i][j] - mean[r];
          q += dlt * dlt;
        }
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    rstd[r] = rsqrtf(q / static_cast<float>(C) + eps);
  }
  if (stats) {                                           // statistics only: the normalisation is folded into the next GEMM
    if (lane == 0) {
      stats[row0] = make_float2(mean[0], rstd[0]);
      if (two) stats[row0 + 1] = make_float2(mean[1], rstd[1]);
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int iv = lane + 32 * i;
    if (iv < nvec) {
      const float4* g4 = reinterpret_cast<const float4*>(gamma + iv * 8);
      const float4* b4 = reinterpret_cast<const float4*>(beta + iv * 8);
      const float4 ga = __ldg(g4), gb = __ldg(g4 + 1), ba = __ldg(b4), bb = __ldg(b4 + 1);
      const float gg[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
      const float be[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        if (r == 0 || two) {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = (v[r][i][j] - mean[r]) * rstd[r] * gg[j] + be[j];
          reinterpret_cast<uint4*>(y + (row0 + r) * static_cast<long long>(C))[iv] = pack8(o);
        }
      }
    }
  }
}

static int pick_rpi(int nvec) {
  int rpi = 256 / nvec;
  return rpi < 1 ? 1 : rpi;
}

}  // namespace ccedit

using namespace ccedit;

extern "C" int ccedit_groupnorm_spatial(const void* x, void* y, const float* gamma, const float* beta, float* partial,
                                        int32_t F, int32_t HW, int32_t C, float eps, int32_t silu, void* stream) {
  CCEDIT_CHECK_ARG(x && y && gamma && beta && partial, "ccedit_groupnorm_spatial: null pointer");
  CCEDIT_CHECK_ARG(F > 0 && HW > 0 && C > 0 && C % kGroups == 0 && C % 8 == 0 && C <= 4096,
                   "ccedit_groupnorm_spatial: bad shape F=%d HW=%d C=%d (C must be a multiple of 32)", F, HW, C);
  const int nvec = C / 8, rpi = pick_rpi(nvec);
  const int threads = nvec * rpi;
  long long bytes = static_cast<long long>(HW) * C * 2;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // ---- single-pass kernel when the whole grid can be resident at once (it synchronises the slices of a frame) ----
  {
    const size_t smem = 2 * static_cast<size_t>(rpi) * C * sizeof(float);
    int per_sm = 0;
    const int sms = device_sm_count();
    static const bool disabled = [] { const char* e = getenv("CCEDIT_GN_FUSED"); return e && atoi(e) == 0; }();
    if (!disabled && sms > 0 &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gn_spatial_fused_kernel, threads, smem) == cudaSuccess) {
      const long long capacity = static_cast<long long>(per_sm) * sms * 4 / 5;   // margin below the cooperative-launch limit
      int nsplit = static_cast<int>(capacity / F);
      const int want = static_cast<int>(bytes / 16384) > 0 ? static_cast<int>(bytes / 16384) : 1;   // >= 16 KB per slice
      if (nsplit > want) nsplit = want;
      if (nsplit > kMaxSplit) nsplit = kMaxSplit;
      if (nsplit >= 1 && 2 * F <= kGnCounterFloats) {
        int* counters = reinterpret_cast<int*>(partial);      // fixed place: [kGnCounterFloats] ints ahead of the partial sums
        const __half* xa = static_cast<const __half*>(x);
        __half* ya = static_cast<__half*>(y);
        float* pa = partial + kGnCounterFloats;
        void* args[] = {&xa, &ya, &gamma, &beta, &pa, &counters, &HW, &C, const_cast<int*>(&nvec), const_cast<int*>(&rpi),
                        &eps, &silu};
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(nsplit, F);
        cfg.blockDim = dim3(threads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = pdl_enabled(8) ? 2 : 1;
        cudaError_t le = cudaLaunchKernelExC(&cfg, reinterpret_cast<const void*>(gn_spatial_fused_kernel), args);
        if (le != cudaSuccess && cfg.numAttrs == 2) {          // the two attributes together refused: cooperative alone
          (void)cudaGetLastError();
          cfg.numAttrs = 1;
          le = cudaLaunchKernelExC(&cfg, reinterpret_cast<const void*>(gn_spatial_fused_kernel), args);
        }
        if (le == cudaSuccess) {
          g_launch_count.fetch_add(1, std::memory_order_relaxed);
          return CCEDIT_OK;
        }
        (void)cudaGetLastError();                              // grid refused (too large for co-residency): two-kernel version
      }
    }
  }
  int nsplit = static_cast<int>(bytes / 65536);
  if (nsplit < 1) nsplit = 1;
  if (nsplit > kMaxSplit) nsplit = kMaxSplit;
  partial += kGnCounterFloats;
  gn_spatial_stats_kernel<<<dim3(nsplit, F), threads, 2 * rpi * C * sizeof(float), st>>>(
      static_cast<const __half*>(x), partial, HW, C, nvec, rpi, nsplit);
  CCEDIT_CUDA_LAUNCH_CHECK("ccedit_groupnorm_spatial(stats)");
  int nchunk = static_cast<int>(bytes / 32768);
  if (nchunk < 1) nchunk = 1;
  if (nchunk > 64) nchunk = 64;
  gn_spatial_apply_kernel<<<dim3(nchunk, F), threads, 0, st>>>(static_cast<const __half*>(x), static_cast<__half*>(y),
                                                               gamma, beta, partial, HW, C, nvec, rpi, nsplit, eps, silu);
  g_launch_count.fetch_add(2, std::memory_order_relaxed);
  CCEDIT_CUDA_LAUNCH_CHECK("ccedit_groupnorm_spatial(apply)");
  return CCEDIT_OK;
}

extern "C" int ccedit_groupnorm_temporal(const void* x, void* y, const float* gamma, const float* beta, int32_t B,
                                         int32_t T, int32_t HW, int32_t C, float eps, int32_t silu, void* stream) {
  CCEDIT_CHECK_ARG(x && y && gamma && beta, "ccedit_groupnorm_temporal: null pointer");
  CCEDIT_CHECK_ARG(B > 0 && T > 0 && HW > 0 && C > 0 && C % kGroups == 0 && C % 8 == 0 && C <= 4096,
                   "ccedit_groupnorm_temporal: bad shape B=%d T=%d HW=%d C=%d", B, T, HW, C);
  const int nvec = C / 8, ppb = pick_rpi(nvec);
  const int threads = nvec * ppb;
  const long long npix = static_cast<long long>(B) * HW;
  const int grid = static_cast<int>((npix + ppb - 1) / ppb);
  (void)launch_pdl(8, gn_temporal_kernel, dim3(grid), dim3(threads), (2 * ppb * C + ppb * 2 * kGroups) * sizeof(float),
                   static_cast<cudaStream_t>(stream), static_cast<const __half*>(x), static_cast<__half*>(y), gamma, beta, B, T, HW,
                   C, nvec, ppb, eps, silu);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  CCEDIT_CUDA_LAUNCH_CHECK("ccedit_groupnorm_temporal");
  return CCEDIT_OK;
}

// Row statistics only (LayerNorm folded into the consuming GEMM): one warp per FOUR rows, the rows stay packed in
// registers (NV uint4 per lane and row) so that ~2.5 KB per warp are in flight; two-pass variance on the register copy.
template <int NV>
__global__ void __launch_bounds__(256) layernorm_stats_kernel(const __half* __restrict__ x, long long ldx,
                                                              float2* __restrict__ stats, long long M, int C, float eps) {
  constexpr int R = 4;
  const int lane = threadIdx.x & 31;
  const long long row0 = (static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5)) * R;
  if (row0 >= M) return;
  const int nvec = C >> 3;
  uint4 v[R][NV];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const long long row = row0 + r < M ? row0 + r : M - 1;
    const uint4* xr = reinterpret_cast<const uint4*>(x + row * ldx);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int iv = lane + 32 * i;
      v[r][i] = iv < nvec ? __ldg(xr + iv) : make_uint4(0, 0, 0, 0);
    }
  }
  const float invc = 1.f / static_cast<float>(C);
  float mean[R], q[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float f[8];
      unpack8(v[r][i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += f[j];
    }
    mean[r] = s;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int r = 0; r < R; ++r) mean[r] += __shfl_xor_sync(0xffffffffu, mean[r], o);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    mean[r] *= invc;
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (lane + 32 * i < nvec) {
        float f[8];
        unpack8(v[r][i], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float dlt = f[j] - mean[r];
          a = fmaf(dlt, dlt, a);
        }
      }
    }
    q[r] = a;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int r = 0; r < R; ++r) q[r] += __shfl_xor_sync(0xffffffffu, q[r], o);
  if (lane < R && row0 + lane < M) {
    float m = mean[0], qq = q[0];
#pragma unroll
    for (int r = 1; r < R; ++r)
      if (lane == r) { m = mean[r]; qq = q[r]; }
    stats[row0 + lane] = make_float2(m, rsqrtf(qq * invc + eps));
  }
}

static int layernorm_launch(const char* what, const void* x, int64_t ldx, void* y, const float* gamma, const float* beta,
                            int64_t M, int32_t C, float eps, float* stats, void* stream) {
  using namespace ccedit;
  CCEDIT_CHECK_ARG(M > 0 && C > 0 && C % 8 == 0 && C <= 2560 && ldx % 8 == 0,
                   "%s: bad shape M=%lld C=%d ldx=%lld (C %% 8 == 0, C <= 2560)", what, (long long)M, C, (long long)ldx);
  const int wpb = 8;
  const long long grid = (M + 2 * wpb - 1) / (2 * wpb);
  const dim3 g(static_cast<unsigned>(grid));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __half* xp = static_cast<const __half*>(x);
  __half* yp = static_cast<__half*>(y);
  float2* sp = reinterpret_cast<float2*>(stats);
  const int nv = (C / 8 + 31) / 32;
  switch (nv) {
    case 1: layernorm_kernel<1><<<g, wpb * 32, 0, st>>>(xp, ldx, yp, gamma, beta, M, C, eps, sp); break;
    case 2: layernorm_kernel<2><<<g, wpb * 32, 0, st>>>(xp, ldx, yp, gamma, beta, M, C, eps, sp); break;
    case 3: layernorm_kernel<3><<<g, wpb * 32, 0, st>>>(xp, ldx, yp, gamma, beta, M, C, eps, sp); break;
    case 4:
    case 5: layernorm_kernel<5><<<g, wpb * 32, 0, st>>>(xp, ldx, yp, gamma, beta, M, C, eps, sp); break;
    default: layernorm_kernel<10><<<g, wpb * 32, 0, st>>>(xp, ldx, yp, gamma, beta, M, C, eps, sp); break;
  }
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  CCEDIT_CUDA_LAUNCH_CHECK(what);
  return CCEDIT_OK;
}

extern "C" int ccedit_layernorm(const void* x, int64_t ldx, void* y, const float* gamma, const float* beta, int64_t M,
                                int32_t C, float eps, void* stream) {
  CCEDIT_CHECK_ARG(x && y && gamma && beta, "ccedit_layernorm: null pointer");
  return layernorm_launch("ccedit_layernorm", x, ldx, y, gamma, beta, M, C, eps, nullptr, stream);
}

extern "C" int ccedit_layernorm_stats(const void* x, int64_t ldx, float* stats, int64_t M, int32_t C, float eps,
                                      void* stream) {
  using namespace ccedit;
  CCEDIT_CHECK_ARG(x && stats, "ccedit_layernorm_stats: null pointer");
  CCEDIT_CHECK_ARG((reinterpret_cast<uintptr_t>(stats) & 7) == 0, "ccedit_layernorm_stats: stats must be 8-byte aligned");
  CCEDIT_CHECK_ARG(M > 0 && C > 0 && C % 8 == 0 && C <= 2560 && ldx % 8 == 0,
                   "ccedit_layernorm_stats: bad shape M=%lld C=%d ldx=%lld (C %% 8 == 0, C <= 2560)", (long long)M, C,
                   (long long)ldx);
  const int nv = (C / 8 + 31) / 32;
  if (nv > 5)                                            // very wide rows: the register-light generic kernel
    return layernorm_launch("ccedit_layernorm_stats", x, ldx, nullptr, nullptr, nullptr, M, C, eps, stats, stream);
  const int wpb = 8;
  const dim3 g(static_cast<unsigned>((M + 4 * wpb - 1) / (4 * wpb)));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __half* xp = static_cast<const __half*>(x);
  float2* sp = reinterpret_cast<float2*>(stats);
  switch (nv) {
    case 1: layernorm_stats_kernel<1><<<g, wpb * 32, 0, st>>>(xp, ldx, sp, M, C, eps); break;
    case 2: layernorm_stats_kernel<2><<<g, wpb * 32, 0, st>>>(xp, ldx, sp, M, C, eps); break;
    case 3: layernorm_stats_kernel<3><<<g, wpb * 32, 0, st>>>(xp, ldx, sp, M, C, eps); break;
    default: layernorm_stats_kernel<5><<<g, wpb * 32, 0, st>>>(xp, ldx, sp, M, C, eps); break;
  }
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  CCEDIT_CUDA_LAUNCH_CHECK("ccedit_layernorm_stats");
  return CCEDIT_OK;
}

// (sum, sum of squares) partials written by the producing GEMM's epilogue (ccedit_gemm_desc.stats_out) -> (mean, rstd).
namespace ccedit {
__global__ void layernorm_combine_kernel(const float2* __restrict__ partial, int P, long long M, float invc, float eps,
                                         float2* __restrict__ stats) {
  const long long m = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float s = 0.f, q = 0.f;
  for (int k = 0; k < P; ++k) {
    const float2 v = __ldg(partial + m * P + k);
    s += v.x;
    q += v.y;
  }
  const float mean = s * invc;
  const float var = fmaxf(fmaf(-mean, mean, q * invc), 0.f);
  stats[m] = make_float2(mean, rsqrtf(var + eps));
}
}  // namespace ccedit

extern "C" int ccedit_layernorm_stats_combine(const float* partial, int32_t P, float* stats, int64_t M, int32_t C,
                                              float eps, void* stream) {
  using namespace ccedit;
  CCEDIT_CHECK_ARG(partial && stats && P >= 1 && P <= 64 && M > 0 && C > 0, "ccedit_layernorm_stats_combine: bad arguments");
  CCEDIT_CHECK_ARG(((reinterpret_cast<uintptr_t>(partial) | reinterpret_cast<uintptr_t>(stats)) & 7) == 0,
                   "ccedit_layernorm_stats_combine: pointers must be 8-byte aligned");
  const int threads = 256;
  layernorm_combine_kernel<<<static_cast<unsigned>((M + threads - 1) / threads), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float2*>(partial), P, M, 1.f / static_cast<float>(C), eps, reinterpret_cast<float2*>(stats));
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  CCEDIT_CUDA_LAUNCH_CHECK("ccedit_layernorm_stats_combine");
  return CCEDIT_OK;
}

// First two layers of ControlNet2D.input_hint_block fused into one pass over the full-resolution hint video
// (controlmodel.py:215-219: conv3x3(hint_channels -> 16) + SiLU + conv3x3(16 -> 16) + SiLU, both stride 1, pad 1).
//
// At 34 frames x 512 x 768 these two layers are 13.4 M pixels x (27 + 144) x 16 MACs = 73 GFLOP - nothing - but 642 MB of
// unavoidable HBM traffic (read the 8-channel padded hint once, write 16 channels once).  Run through the tcgen05
// tap-GEMM they cost 4.8 ms per network call (K padded 3 -> 64 and 16 -> 64 per tap, 32-byte TMA rows, and a 428 MB
// intermediate written and re-read); here one CTA takes a 16 x 64 pixel tile of one frame:
//   1. cp.async the (16+4) x (64+4) x 8-channel input halo tile into shared memory (zero fill outside the image);
//   2. layer 0 on the (16+2) x (64+2) halo of layer 1 as an implicit GEMM on mma.sync.m16n8k16: M = 16 consecutive
//      halo pixels, K = 9 taps x 8 channels (two taps per k-step, padded to 80), N = 16; bias + SiLU; pixels outside
//      the image are forced to zero (they are layer 1's zero padding); fp16 result stays in shared memory;
//   3. layer 1 the same way (K = 9 taps x 16 channels, one tap per k-step); bias + SiLU; staged in shared memory and
//      written as whole 2 KB rows.
// The weights live in registers as mma B fragments (56 per thread).  Bound: HBM (algorithmic bytes = 48 B / pixel).
#include "common.cuh"
#include "../../include/ccedit_b200.h"

#include <atomic>
#include <mutex>

namespace ccedit {
extern std::atomic<long long> g_launch_count;

constexpr int kHsTH = 16, kHsTW = 64;                 // output tile
constexpr int kHsIH = kHsTH + 4, kHsIW = kHsTW + 4;   // input halo tile
constexpr int kHsMH = kHsTH + 2, kHsMW = kHsTW + 2;   // layer-0 output (layer-1 input) halo tile
constexpr int kHsThreads = 256;
constexpr int kHsInBytes = kHsIH * kHsIW * 16;        // 8 fp16 channels per pixel
constexpr int kHsMidBytes = kHsMH * kHsMW * 32;       // 16 fp16 channels per pixel
constexpr int kHsOutBytes = kHsTH * kHsTW * 32;
constexpr int kHsSmem = kHsInBytes + kHsMidBytes + kHsOutBytes;
constexpr int kHsK0 = 80, kHsK1 = 144;                // padded K of the two layers

__device__ __forceinline__ uint32_t hs_pack(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(kHsThreads, 2)
hint_stem01_kernel(const __half* __restrict__ x, __half* __restrict__ y, const __half* __restrict__ w0,
                   const float* __restrict__ b0, const __half* __restrict__ w1, const float* __restrict__ b1, int H,
                   int W) {
  extern __shared__ __align__(128) uint8_t hs_smem[];
  uint8_t* s_in = hs_smem;
  uint8_t* s_mid = hs_smem + kHsInBytes;
  uint8_t* s_out = s_mid + kHsMidBytes;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int x0 = blockIdx.x * kHsTW, y0 = blockIdx.y * kHsTH, f = blockIdx.z;
  const __half* xf = x + static_cast<long long>(f) * H * W * 8;

  // ---- 1. input halo tile ----
  for (int i = tid; i < kHsIH * kHsIW; i += kHsThreads) {
    const int iy = i / kHsIW, ix = i - iy * kHsIW;
    const int gy = y0 - 2 + iy, gx = x0 - 2 + ix;
    const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
    cp_async_16(smem_u32(s_in + i * 16), xf + (static_cast<long long>(ok ? gy : 0) * W + (ok ? gx : 0)) * 8, ok);
  }
  cp_async_commit();

  // ---- weights as B fragments: b0 = W[n = g (+8 per n-tile)][k = 16 ks + 2t, +1], b1 = ... [k + 8] ----
  uint32_t wf0[5][2][2], wf1[9][2][2];
#pragma unroll
  for (int ks = 0; ks < 5; ++ks)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const __half* p = w0 + (nt * 8 + g) * kHsK0 + ks * 16 + 2 * t;
      wf0[ks][nt][0] = *reinterpret_cast<const uint32_t*>(p);
      wf0[ks][nt][1] = *reinterpret_cast<const uint32_t*>(p + 8);
    }
#pragma unroll
  for (int ks = 0; ks < 9; ++ks)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const __half* p = w1 + (nt * 8 + g) * kHsK1 + ks * 16 + 2 * t;
      wf1[ks][nt][0] = *reinterpret_cast<const uint32_t*>(p);
      wf1[ks][nt][1] = *reinterpret_cast<const uint32_t*>(p + 8);
    }
  float bias0[2][2], bias1[2][2];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    bias0[nt][0] = b0[nt * 8 + 2 * t];
    bias0[nt][1] = b0[nt * 8 + 2 * t + 1];
    bias1[nt][0] = b1[nt * 8 + 2 * t];
    bias1[nt][1] = b1[nt * 8 + 2 * t + 1];
  }
  cp_async_wait<0>();
  __syncthreads();

  // ---- 2. layer 0 on the halo tile: m-tile = 16 consecutive pixels of the flattened kHsMH x kHsMW region ----
  {
    constexpr int npix = kHsMH * kHsMW;
    constexpr int ntile = (npix + 15) / 16;
    // ldmatrix.x4 row address of this lane: matrix id = lane / 8 -> (pixel half, tap parity), row = lane % 8
    const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8;
    const int ltap = lane >> 4;
    constexpr int kU = 1;                                  // m-tiles in flight per warp (measured: 1 = 594 us, 3-4 = 658 us)
    for (int mt0 = warp; mt0 < ntile; mt0 += kU * (kHsThreads / 32)) {
      int py[kU], px[kU];
      float acc[kU][2][4] = {};
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        int p = (mt0 + u * (kHsThreads / 32)) * 16 + lrow;
        p = p < npix ? p : npix - 1;
        py[u] = p / kHsMW;
        px[u] = p - py[u] * kHsMW;
      }
#pragma unroll
      for (int ks = 0; ks < 5; ++ks) {
        int tap = 2 * ks + ltap;
        tap = tap < 9 ? tap : 0;                          // k >= 72 carries zero weights: any readable address will do
        const int ty = tap / 3, tx = tap - ty * 3;
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          uint32_t a[4];
          ldmatrix_x4(a, smem_u32(s_in + ((py[u] + ty) * kHsIW + px[u] + tx) * 16));
          mma_m16n8k16(acc[u][0], a, wf0[ks][0]);
          mma_m16n8k16(acc[u][1], a, wf0[ks][1]);
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int mt = mt0 + u * (kHsThreads / 32);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {                   // rows g and g + 8 of the m-tile
          const int q = mt * 16 + g + hh * 8;
          if (q < npix) {
            const int qy = q / kHsMW, qx = q - qy * kHsMW;
            const int gy = y0 - 1 + qy, gx = x0 - 1 + qx;
            const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
              const float v0 = in ? silu_f(acc[u][nt][2 * hh] + bias0[nt][0]) : 0.f;
              const float v1 = in ? silu_f(acc[u][nt][2 * hh + 1] + bias0[nt][1]) : 0.f;
              *reinterpret_cast<uint32_t*>(s_mid + q * 32 + (nt * 8 + 2 * t) * 2) = hs_pack(v0, v1);
            }
          }
        }
      }
    }
  }
  __syncthreads();

  // ---- 3. layer 1: m-tile = 16 pixels of one output row; one tap (16 channels) per k-step ----
  {
    const int lpx = (lane & 7) + ((lane >> 3) & 1) * 8;    // pixel inside the m-tile
    const int lch = (lane >> 4) * 16;                      // channel half (bytes)
    constexpr int kU = 1;
    for (int mt0 = warp; mt0 < kHsTH * (kHsTW / 16); mt0 += kU * (kHsThreads / 32)) {
      int ry[kU], cx[kU];
      float acc[kU][2][4] = {};
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int mt = mt0 + u * (kHsThreads / 32);
        ry[u] = mt / (kHsTW / 16);
        cx[u] = (mt - ry[u] * (kHsTW / 16)) * 16;
      }
#pragma unroll
      for (int ks = 0; ks < 9; ++ks) {
        const int ty = ks / 3, tx = ks - ty * 3;
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          uint32_t a[4];
          ldmatrix_x4(a, smem_u32(s_mid + ((ry[u] + ty) * kHsMW + cx[u] + lpx + tx) * 32 + lch));
          mma_m16n8k16(acc[u][0], a, wf1[ks][0]);
          mma_m16n8k16(acc[u][1], a, wf1[ks][1]);
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint8_t* dst = s_out + (ry[u] * kHsTW + cx[u] + g + hh * 8) * 32;
#pragma unroll
          for (int nt = 0; nt < 2; ++nt)
            *reinterpret_cast<uint32_t*>(dst + (nt * 8 + 2 * t) * 2) =
                hs_pack(silu_f(acc[u][nt][2 * hh] + bias1[nt][0]), silu_f(acc[u][nt][2 * hh + 1] + bias1[nt][1]));
        }
    }
  }
  __syncthreads();

  // ---- 4. coalesced write-out: 16 bytes per thread, 2 KB per tile row ----
  __half* yf = y + static_cast<long long>(f) * H * W * 16;
  for (int i = tid; i < kHsTH * kHsTW * 2; i += kHsThreads) {
    const int pix = i >> 1, half = i & 1;
    const int ry = pix / kHsTW, rx = pix - ry * kHsTW;
    const int gy = y0 + ry, gx = x0 + rx;
    if (gy < H && gx < W)
      *reinterpret_cast<uint4*>(yf + (static_cast<long long>(gy) * W + gx) * 16 + half * 8) =
          *reinterpret_cast<const uint4*>(s_out + pix * 32 + half * 16);
  }
}

}  // namespace ccedit

extern "C" int ccedit_hint_stem01(const void* x, void* y, const void* w0, const float* b0, const void* w1,
                                  const float* b1, int32_t F, int32_t H, int32_t W, void* stream) {
  using namespace ccedit;
  CCEDIT_CHECK_ARG(x && y && w0 && b0 && w1 && b1, "ccedit_hint_stem01: null pointer");
  CCEDIT_CHECK_ARG(F >= 1 && H >= 1 && W >= 1 && F <= 65535, "ccedit_hint_stem01: bad shape F=%d H=%d W=%d", F, H, W);
  CCEDIT_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(w0) & 3) == 0 && (reinterpret_cast<uintptr_t>(w1) & 3) == 0,
                   "ccedit_hint_stem01: x/y must be 16-byte aligned, w0/w1 4-byte aligned");
  static std::atomic<bool> attr_set[kMaxDevices];         // function attributes belong to a device
  const int dev = current_device();
  CCEDIT_CHECK_ARG(dev >= 0, "ccedit_hint_stem01: no current CUDA device");
  if (!attr_set[dev].load(std::memory_order_acquire)) {
    const cudaError_t e = cudaFuncSetAttribute(hint_stem01_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHsSmem);
    if (e != cudaSuccess) {
      set_last_error("ccedit_hint_stem01: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return CCEDIT_ERR_CUDA;
    }
    attr_set[dev].store(true, std::memory_order_release);
  }
  dim3 grid((W + kHsTW - 1) / kHsTW, (H + kHsTH - 1) / kHsTH, F);
  hint_stem01_kernel<<<grid, kHsThreads, kHsSmem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(x), static_cast<__half*>(y), static_cast<const __half*>(w0), b0,
      static_cast<const __half*>(w1), b1, H, W);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  CCEDIT_CUDA_LAUNCH_CHECK("ccedit_hint_stem01");
  return CCEDIT_OK;
}

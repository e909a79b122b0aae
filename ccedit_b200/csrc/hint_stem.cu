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
              // 16-byte half of the pixel XOR-ed with bit 2 of the pixel index: layer 1's ldmatrix rows (consecutive
              // 32-byte pixels) then fall into eight different 16-byte slots of a 128-byte line (was a 2-way conflict)
              *reinterpret_cast<uint32_t*>(s_mid + q * 32 + ((nt ^ ((q >> 2) & 1)) << 4) + 4 * t) = hs_pack(v0, v1);
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
    const int lhalf = lane >> 4;                           // channel half
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
          const int q = (ry[u] + ty) * kHsMW + cx[u] + lpx + tx;
          ldmatrix_x4(a, smem_u32(s_mid + q * 32 + ((lhalf ^ ((q >> 2) & 1)) << 4)));
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


// ---------------------------------------------------------------------------------------------------------------
// Layers 2 and 3 of the hint stem fused the same way (controlmodel.py:220-223: conv3x3(16 -> 32, stride 2) + SiLU +
// conv3x3(32 -> 32) + SiLU).  Through the tap-GEMM they were a parity split of the 16-channel full-resolution tensor plus
// two GEMMs whose k-blocks are one tap of 16 / 32 channels each - all barrier hand-shakes, no work: 78 + 243 + 245 us for the
// 17 frames of a de-duplicated call, against 214 MB read + 107 MB written.  One CTA takes an 8 x 32 tile of the
// half-resolution output:
//   1. cp.async the (2 * 10 + 1) x (2 * 34 + 1) x 16-channel input window of the tile's 10 x 34 layer-2 halo (zero fill
//      outside the image = the convolution's padding);
//   2. layer 2 on the halo: M = 16 consecutive halo pixels, one tap (16 channels, input pixel (2 y + ky, 2 x + kx)) per
//      k-step, N = 32; bias + SiLU; halo pixels outside the image are layer 3's zero padding; fp16 result in shared memory;
//   3. layer 3: M = 16 pixels of an output row, 18 k-steps (9 taps x two 16-channel halves), N = 32, weights from shared
//      memory (padded rows: conflict-free fragment loads); bias + SiLU; staged over the dead input window and written as
//      whole 2 KB rows.
// Layer-2 weights live in registers as mma B fragments (72 per thread).  H and W must be even.
// Shared-memory layouts (ncu of the first version: 44 M bank conflicts, shared-memory pipe 73 % busy, mio_throttle 4.8
// per issue - an ldmatrix 8 x 8 fetch wants its eight 16-byte rows in eight different 16-byte slots of a 128-byte line):
//   * input window: even and odd columns in separate planes ([row][parity][column / 2][16 ch]), so that the stride-2
//     taps read consecutive 32-byte pixels, and the 16-byte half of a pixel XOR-ed with bit 2 of the plane column;
//   * layer-2 halo (64 bytes per pixel): the 16-byte chunk XOR-ed with bits 1-2 of the linear pixel index.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kH2TH = 8, kH2TW = 32;                        // output tile (half resolution)
constexpr int kH2MH = kH2TH + 2, kH2MW = kH2TW + 2;         // layer-2 output (layer-3 input) halo tile
constexpr int kH2IH = 2 * kH2MH + 1, kH2IW = 2 * kH2MW + 1; // input window (full resolution)
constexpr int kH2PW = (kH2IW + 1) / 2;                      // columns per parity plane of the input window
constexpr int kH2InBytes = kH2IH * 2 * kH2PW * 32;          // 16 fp16 channels per pixel
constexpr int kH2MidBytes = kH2MH * kH2MW * 64;             // 32 fp16 channels per pixel
constexpr int kH2K2 = 144, kH2K3 = 288, kH2W3Stride = kH2K3 + 8;   // halves; +8: fragment loads hit 32 different banks
constexpr int kH2W3Bytes = 32 * kH2W3Stride * 2;
constexpr int kH2Smem = kH2InBytes + kH2MidBytes + kH2W3Bytes;
static_assert(kH2TH * kH2TW * 64 <= kH2InBytes, "the output tile is staged over the input window");

__global__ void __launch_bounds__(kHsThreads, 2)
hint_stem23_kernel(const __half* __restrict__ x, __half* __restrict__ y, const __half* __restrict__ w2,
                   const float* __restrict__ b2, const __half* __restrict__ w3, const float* __restrict__ b3, int H, int W) {
  extern __shared__ __align__(128) uint8_t hs_smem[];
  uint8_t* s_in = hs_smem;
  uint8_t* s_mid = hs_smem + kH2InBytes;
  __half* s_w3 = reinterpret_cast<__half*>(s_mid + kH2MidBytes);
  uint8_t* s_out = s_in;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int H2 = H >> 1, W2 = W >> 1;
  const int x0 = blockIdx.x * kH2TW, y0 = blockIdx.y * kH2TH, f = blockIdx.z;      // tile origin at half resolution
  const __half* xf = x + static_cast<long long>(f) * H * W * 16;

  // ---- 1. input window: rows 2 y0 - 3 .., columns 2 x0 - 3 .. ----
  for (int i = tid; i < kH2IH * kH2IW * 2; i += kHsThreads) {
    const int pix = i >> 1, half = i & 1;
    const int iy = pix / kH2IW, ix = pix - iy * kH2IW;
    const int gy = 2 * y0 - 3 + iy, gx = 2 * x0 - 3 + ix;
    const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
    const int col = ix >> 1;
    cp_async_16(smem_u32(s_in + ((iy * 2 + (ix & 1)) * kH2PW + col) * 32 + ((half ^ ((col >> 2) & 1)) << 4)),
                xf + (static_cast<long long>(ok ? gy : 0) * W + (ok ? gx : 0)) * 16 + half * 8, ok);
  }
  cp_async_commit();
  // layer-3 weights -> shared memory (rows of 288 halves, padded to 296)
  for (int i = tid; i < 32 * (kH2K3 / 8); i += kHsThreads) {
    const int n = i / (kH2K3 / 8), c = i - n * (kH2K3 / 8);
    *reinterpret_cast<uint4*>(s_w3 + n * kH2W3Stride + c * 8) = *reinterpret_cast<const uint4*>(w3 + n * kH2K3 + c * 8);
  }
  // layer-2 weights as B fragments: b0 = W[n = nt * 8 + g][k = 16 ks + 2t, +1], b1 = ... [k + 8]
  uint32_t wf2[9][4][2];
#pragma unroll
  for (int ks = 0; ks < 9; ++ks)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const __half* p = w2 + (nt * 8 + g) * kH2K2 + ks * 16 + 2 * t;
      wf2[ks][nt][0] = *reinterpret_cast<const uint32_t*>(p);
      wf2[ks][nt][1] = *reinterpret_cast<const uint32_t*>(p + 8);
    }
  float bias2[4][2], bias3[4][2];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    bias2[nt][0] = b2[nt * 8 + 2 * t];
    bias2[nt][1] = b2[nt * 8 + 2 * t + 1];
    bias3[nt][0] = b3[nt * 8 + 2 * t];
    bias3[nt][1] = b3[nt * 8 + 2 * t + 1];
  }
  cp_async_wait<0>();
  __syncthreads();

  // ---- 2. layer 2 (stride 2) on the halo tile: m-tile = 16 consecutive pixels of the flattened kH2MH x kH2MW region ----
  {
    constexpr int npix = kH2MH * kH2MW;
    constexpr int ntile = (npix + 15) / 16;
    const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8;    // ldmatrix.x4: pixel of the m-tile this lane addresses
    const int lhalf = lane >> 4;                            // channel half
    for (int mt = warp; mt < ntile; mt += kHsThreads / 32) {
      int p = mt * 16 + lrow;
      p = p < npix ? p : npix - 1;
      const int py = p / kH2MW, px = p - py * kH2MW;
      // input pixel (2 py + ty, 2 px + tx): plane tx & 1, plane column px + (tx >> 1)
      const uint32_t row0 = smem_u32(s_in) + static_cast<uint32_t>((2 * py) * 2 * kH2PW) * 32u;
      uint32_t acol[2];                                     // byte offset of plane column px / px + 1 incl. the swizzled half
#pragma unroll
      for (int j = 0; j < 2; ++j) acol[j] = static_cast<uint32_t>((px + j) * 32 + ((lhalf ^ (((px + j) >> 2) & 1)) << 4));
      float acc[4][4] = {};
#pragma unroll
      for (int ks = 0; ks < 9; ++ks) {
        const int ty = ks / 3, tx = ks - ty * 3;
        uint32_t a[4];
        ldmatrix_x4(a, row0 + static_cast<uint32_t>((ty * 2 + (tx & 1)) * kH2PW * 32) + acol[tx >> 1]);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_m16n8k16(acc[nt], a, wf2[ks][nt]);
      }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {                       // rows g and g + 8 of the m-tile
        const int q = mt * 16 + g + hh * 8;
        if (q < npix) {
          const int qy = q / kH2MW, qx = q - qy * kH2MW;
          const int gy = y0 - 1 + qy, gx = x0 - 1 + qx;
          const bool in = gy >= 0 && gy < H2 && gx >= 0 && gx < W2;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            const float v0 = in ? silu_f(acc[nt][2 * hh] + bias2[nt][0]) : 0.f;
            const float v1 = in ? silu_f(acc[nt][2 * hh + 1] + bias2[nt][1]) : 0.f;
            *reinterpret_cast<uint32_t*>(s_mid + q * 64 + ((nt ^ ((q >> 1) & 3)) << 4) + 4 * t) = hs_pack(v0, v1);
          }
        }
      }
    }
  }
  __syncthreads();                                           // s_mid complete; the input window is dead from here on

  // ---- 3. layer 3: m-tile = 16 pixels of one output row; 18 k-steps = 9 taps x two 16-channel halves ----
  {
    const int lpx = (lane & 7) + ((lane >> 3) & 1) * 8;
    const int lhalf = lane >> 4;
    constexpr int ntile = kH2TH * (kH2TW / 16);              // 16: two per warp, processed together (B fragments shared)
    static_assert(ntile == 2 * (kHsThreads / 32), "two m-tiles per warp");
    int ry[2], cx[2], pbase[2];
    float acc[2][4][4] = {};
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int mt = warp + u * (kHsThreads / 32);
      ry[u] = mt / (kH2TW / 16);
      cx[u] = (mt - ry[u] * (kH2TW / 16)) * 16;
      pbase[u] = ry[u] * kH2MW + cx[u] + lpx;               // linear halo pixel of this lane's row for tap (0, 0)
    }
    const uint32_t smid = smem_u32(s_mid);
#pragma unroll
    for (int ks = 0; ks < 18; ++ks) {
      const int tap = ks >> 1, kc = ks & 1;
      const int ty = tap / 3, tx = tap - ty * 3;
      uint32_t bf[4][2];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const __half* p = s_w3 + (nt * 8 + g) * kH2W3Stride + ks * 16 + 2 * t;
        bf[nt][0] = *reinterpret_cast<const uint32_t*>(p);
        bf[nt][1] = *reinterpret_cast<const uint32_t*>(p + 8);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        uint32_t a[4];
        const int q = pbase[u] + ty * kH2MW + tx;                                    // halo pixel; chunk = 2 kc + half
        ldmatrix_x4(a, smid + static_cast<uint32_t>(q * 64 + (((2 * kc + lhalf) ^ ((q >> 1) & 3)) << 4)));
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_m16n8k16(acc[u][nt], a, bf[nt]);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint8_t* dst = s_out + (ry[u] * kH2TW + cx[u] + g + hh * 8) * 64;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
          *reinterpret_cast<uint32_t*>(dst + (nt * 8 + 2 * t) * 2) =
              hs_pack(silu_f(acc[u][nt][2 * hh] + bias3[nt][0]), silu_f(acc[u][nt][2 * hh + 1] + bias3[nt][1]));
      }
  }
  __syncthreads();

  // ---- 4. coalesced write-out: 16 bytes per thread, 2 KB per tile row ----
  __half* yf = y + static_cast<long long>(f) * H2 * W2 * 32;
  for (int i = tid; i < kH2TH * kH2TW * 4; i += kHsThreads) {
    const int pix = i >> 2, q4 = i & 3;
    const int ry = pix / kH2TW, rx = pix - ry * kH2TW;
    const int gy = y0 + ry, gx = x0 + rx;
    if (gy < H2 && gx < W2)
      *reinterpret_cast<uint4*>(yf + (static_cast<long long>(gy) * W2 + gx) * 32 + q4 * 8) =
          *reinterpret_cast<const uint4*>(s_out + pix * 64 + q4 * 16);
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

extern "C" int ccedit_hint_stem23(const void* x, void* y, const void* w2, const float* b2, const void* w3,
                                  const float* b3, int32_t F, int32_t H, int32_t W, void* stream) {
  using namespace ccedit;
  CCEDIT_CHECK_ARG(x && y && w2 && b2 && w3 && b3, "ccedit_hint_stem23: null pointer");
  CCEDIT_CHECK_ARG(F >= 1 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0 && F <= 65535,
                   "ccedit_hint_stem23: bad shape F=%d H=%d W=%d (H and W must be even)", F, H, W);
  CCEDIT_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(w2) & 3) == 0 && (reinterpret_cast<uintptr_t>(w3) & 15) == 0,
                   "ccedit_hint_stem23: x/y/w3 must be 16-byte aligned, w2 4-byte aligned");
  static std::atomic<bool> attr_set[kMaxDevices];         // function attributes belong to a device
  const int dev = current_device();
  CCEDIT_CHECK_ARG(dev >= 0, "ccedit_hint_stem23: no current CUDA device");
  if (!attr_set[dev].load(std::memory_order_acquire)) {
    const cudaError_t e = cudaFuncSetAttribute(hint_stem23_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kH2Smem);
    if (e != cudaSuccess) {
      set_last_error("ccedit_hint_stem23: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return CCEDIT_ERR_CUDA;
    }
    attr_set[dev].store(true, std::memory_order_release);
  }
  dim3 grid((W / 2 + kH2TW - 1) / kH2TW, (H / 2 + kH2TH - 1) / kH2TH, F);
  hint_stem23_kernel<<<grid, kHsThreads, kH2Smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(x), static_cast<__half*>(y), static_cast<const __half*>(w2), b2,
      static_cast<const __half*>(w3), b3, H, W);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  CCEDIT_CUDA_LAUNCH_CHECK("ccedit_hint_stem23");
  return CCEDIT_OK;
}

// Microbenchmark: MUFU.EX2 throughput per SM sub-partition with 1/2/4 warps, alone and mixed with the softmax's other
// instructions (FFMA, FADD, F2FP pack, 3-input max).  Prints clocks per warp-level MUFU.
#include <cstdio>
#include <cuda_fp16.h>
#include <cstdint>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void k(float* out, long long* clk, int iters, float c, float nm) {
  float r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = -0.01f * (threadIdx.x + i);
  float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float x0 = r[i], x1 = r[i + 1], x2 = r[i + 2], x3 = r[i + 3];
      if (MODE >= 1) { x0 = fmaf(x0, c, nm); x1 = fmaf(x1, c, nm); x2 = fmaf(x2, c, nm); x3 = fmaf(x3, c, nm); }
      float p0 = ex2(x0), p1 = ex2(x1), p2 = ex2(x2), p3 = ex2(x3);
      if (MODE >= 2) { s0 += p0; s1 += p1; s2 += p2; s3 += p3; }
      if (MODE >= 3) {
        __half2 h0 = __floats2half2_rn(p0, p1), h1 = __floats2half2_rn(p2, p3);
        acc ^= *reinterpret_cast<uint32_t*>(&h0) + *reinterpret_cast<uint32_t*>(&h1);
      }
      if (MODE < 3) { r[i] = p0 * 1e-3f - 1.f; r[i + 1] = p1 * 1e-3f - 1.f; r[i + 2] = p2 * 1e-3f - 1.f; r[i + 3] = p3 * 1e-3f - 1.f; }
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = s0 + s1 + s2 + s3 + r[5] + __uint_as_float(acc);
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
template <int MODE>
void run(int warps_per_smsp) {
  float* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
  int threads = warps_per_smsp * 4 * 32, iters = 2000;
  k<MODE><<<148, threads>>>(out, clk, 10, 1.1f, -0.5f);
  k<MODE><<<148, threads>>>(out, clk, iters, 1.1f, -0.5f);
  long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  printf("mode %d warps/SMSP %d: %.2f clk per warp-MUFU per SMSP (%.2f per warp)\n", MODE, warps_per_smsp,
         (double)h / (iters * 32.0 * warps_per_smsp), (double)h / (iters * 32.0));
  cudaFree(out); cudaFree(clk);
}
int main() {
  for (int w : {1, 2, 4}) { run<0>(w); run<1>(w); run<2>(w); run<3>(w); }
  return 0;
}

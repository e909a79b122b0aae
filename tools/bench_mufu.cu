// Microbenchmark: MUFU.EX2 throughput per SM sub-partition with 1/2/4 warps, alone and mixed with the softmax's other
// instructions (FFMA, FADD, F2FP pack, 3-input max).  Prints clocks per warp-level MUFU per sub-partition.
//   mode 0: ex2 only   1: + ffma (scale)   2: + fadd (row sum)   3: + f2fp pack   4: + max3   5: pack only (no ex2)
#include <cstdio>
#include <cuda_fp16.h>
#include <cstdint>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t packh(float a, float b) { uint32_t y; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(b), "f"(a)); return y; }
__device__ __forceinline__ float max3(float a, float b, float c) { float y; asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c)); return y; }
template <int MODE>
__global__ void k(float* out, long long* clk, int iters, float c, float nm) {
  float r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = -0.01f * (threadIdx.x + i);
  float s0 = 0, s1 = 0, s2 = 0, s3 = 0, mx = -1e30f;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float x0 = r[i], x1 = r[i + 1], x2 = r[i + 2], x3 = r[i + 3];
      if (MODE >= 4) { mx = max3(mx, x0, x1); mx = max3(mx, x2, x3); }
      if (MODE >= 1) { x0 = fmaf(x0, c, nm); x1 = fmaf(x1, c, nm); x2 = fmaf(x2, c, nm); x3 = fmaf(x3, c, nm); }
      float p0 = x0, p1 = x1, p2 = x2, p3 = x3;
      if (MODE != 5) { p0 = ex2(x0); p1 = ex2(x1); p2 = ex2(x2); p3 = ex2(x3); }
      if (MODE >= 2) { s0 += p0; s1 += p1; s2 += p2; s3 += p3; }
      if (MODE >= 3) { acc ^= packh(p0, p1); acc += packh(p2, p3); }
      r[i] = p0; r[i + 1] = p1; r[i + 2] = p2; r[i + 3] = p3;
    }
  }
  long long t1 = clock64();
  float tot = s0 + s1 + s2 + s3 + mx + __uint_as_float(acc);
#pragma unroll
  for (int i = 0; i < 32; ++i) tot += r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = tot;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
template <int MODE>
void run(int warps_per_smsp) {
  float* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
  int threads = warps_per_smsp * 4 * 32, iters = 2000;
  k<MODE><<<148, threads>>>(out, clk, 10, 0.999f, -0.001f);
  k<MODE><<<148, threads>>>(out, clk, iters, 0.999f, -0.001f);
  long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  printf("mode %d warps/SMSP %d: %.2f clk per element-instruction group per SMSP\n", MODE, warps_per_smsp,
         (double)h / (iters * 32.0 * warps_per_smsp));
  cudaFree(out); cudaFree(clk);
}
int main() {
  for (int w : {1, 2, 4}) { run<0>(w); run<1>(w); run<2>(w); run<3>(w); run<4>(w); run<5>(w); }
  return 0;
}

// Microbenchmark: issue rate of the legacy warp-level tensor-core path (mma.sync.m16n8k16 f16 x f16 -> f32, SASS
// HMMA.16816.F32; also m16n8k8) per SM sub-partition on sm_100a, with 1/2/4 warps per sub-partition and 8 independent
// accumulators per warp.  Prints clocks per HMMA per sub-partition and the chip-wide TFLOP/s that rate amounts to.
#include <cstdio>
#include <cstdint>
template <int K8>
__global__ void k(float* out, long long* clk, int iters) {
  float c[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  uint32_t a[4] = {0x3c003c00u + threadIdx.x, 0x3c003c00u, 0x38003800u, 0x3c003800u}, b[2] = {0x3c003c00u, 0x34003400u};
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (K8)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(b[0]));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
  }
  long long t1 = clock64();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = tot;
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
template <int K8>
void run(int warps_per_smsp) {
  float* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 8);
  int threads = warps_per_smsp * 4 * 32, iters = 4000;
  k<K8><<<148, threads>>>(out, clk, 10);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<K8><<<148, threads>>>(out, clk, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  const double n = (double)iters * 8.0 * warps_per_smsp;            // HMMAs per sub-partition
  const double flop = (K8 ? 2048.0 : 4096.0) * n * 4 * 148;
  printf("%s warps/SMSP %d: %.2f clk per HMMA per SMSP, %.1f TFLOP/s chip-wide\n", K8 ? "m16n8k8 " : "m16n8k16", warps_per_smsp,
         (double)h / n, flop / (ms * 1e-3) / 1e12);
  cudaFree(out); cudaFree(clk);
}
int main() {
  for (int w : {1, 2, 4, 8}) { run<0>(w); run<1>(w); }
  return 0;
}

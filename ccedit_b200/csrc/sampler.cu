// Fused elementwise math of one DPM++(2S) ancestral sampler step with classifier-free guidance and the eps-scaling
// denoiser - everything the reference does in ~15 PyTorch elementwise launches (plus a host sync) between and around
// the two network calls of a step:
//   DiscreteDenoiser.__call__     denoiser.py:22-40      net(x * c_in, idx, c) * c_out + x * c_skip   (EpsScaling: c_skip = 1)
//   VanillaCFG.__call__           guiders.py:25-29       x_u + scale * (x_c - x_u)
//   VanillaCFGTV2V.prepare_inputs guiders.py:56-67       cat([x] * 2)
//   DPMPP2SAncestralSampler       sampling.py:385-407    Euler step, midpoint x2 = m1 x - m2 d, x' = m3 x - m4 d2
//   ancestral_step                sampling.py:182-188    x' + noise * s_noise * sigma_up
// Three kernels per step: prepare (network input of call 1), mid (between the calls), final (after call 2).  All per-step
// scalars (quantised sigmas, c_in, c_out, timestep indices, the four multipliers ...) live in a device table written once
// per schedule; the kernels pick their row through a device-side step index, so ONE captured CUDA graph of a whole step
// (both network calls included) replays for every step.  Every operation is a separately rounded fp32 op in the
// reference's order (no FMA contraction): the fused step is bit-identical to the unfused PyTorch formulas.
// HBM-bound and tiny (the latent is 0.8 MB at the headline shape); what matters is that they are few and graph-capturable.
#include "common.cuh"
#include "../../include/ccedit_b200.h"

#include <atomic>

namespace ccedit {
extern std::atomic<long long> g_launch_count;

__device__ __forceinline__ float cfg_denoised(float eps_u, float eps_c, float x_in, float c_out, float scale) {
  // network(...) * c_out + input * c_skip for both halves, then x_u + scale * (x_c - x_u)
  const float du = __fadd_rn(__fmul_rn(eps_u, c_out), x_in);
  const float dc = __fadd_rn(__fmul_rn(eps_c, c_out), x_in);
  return __fadd_rn(du, __fmul_rn(scale, __fsub_rn(dc, du)));
}

// xin2[0:n] = xin2[n:2n] = x * c_in1 ; t2[0:2B] = idx1
__global__ void sampler_prepare_kernel(const float* __restrict__ x, float* __restrict__ xin2, long long* __restrict__ t2,
                                       const float* __restrict__ sc, const int* __restrict__ step, long long n, int B) {
  const float* s = sc + static_cast<long long>(*step) * CCEDIT_SAMPLER_ROW;
  const float c_in = s[CCEDIT_SC_CIN1];
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < 2 * B) t2[i] = static_cast<long long>(s[CCEDIT_SC_IDX1]);
  if (i >= n) return;
  const float v = __fmul_rn(x[i], c_in);
  xin2[i] = v;
  xin2[n + i] = v;
}

// after network call 1: d = CFG(denoiser), x_euler, and (unless the step is Euler-only) the midpoint x2 and the input of call 2
__global__ void sampler_mid_kernel(const float* __restrict__ x, const float* __restrict__ eps2, float* __restrict__ x_euler,
                                   float* __restrict__ x2, float* __restrict__ xin2, long long* __restrict__ t2,
                                   const float* __restrict__ sc, const int* __restrict__ step, float cfg_scale, long long n,
                                   int B) {
  const float* s = sc + static_cast<long long>(*step) * CCEDIT_SAMPLER_ROW;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < 2 * B) t2[i] = static_cast<long long>(s[CCEDIT_SC_IDX2]);
  if (i >= n) return;
  const float xv = x[i];
  const float d = cfg_denoised(eps2[i], eps2[n + i], xv, s[CCEDIT_SC_COUT1], cfg_scale);
  // x + (x - d) / sigma * (sigma_down - sigma)
  x_euler[i] = __fadd_rn(xv, __fmul_rn(__fdiv_rn(__fsub_rn(xv, d), s[CCEDIT_SC_SIGMA]), s[CCEDIT_SC_DSIGMA]));
  // x2 = m1 * x - m2 * d
  const float v2 = __fsub_rn(__fmul_rn(s[CCEDIT_SC_M1], xv), __fmul_rn(s[CCEDIT_SC_M2], d));
  x2[i] = v2;
  const float vin = __fmul_rn(v2, s[CCEDIT_SC_CIN2]);
  xin2[i] = vin;
  xin2[n + i] = vin;
}

// after network call 2 (or directly after mid when the step is Euler-only): the step's result (+ ancestral noise)
__global__ void sampler_final_kernel(const float* __restrict__ x, const float* __restrict__ x2, const float* __restrict__ x_euler,
                                     const float* __restrict__ eps2, const float* __restrict__ noise, float* __restrict__ x_out,
                                     const float* __restrict__ sc, const int* __restrict__ step, float cfg_scale, float s_noise,
                                     long long n) {
  const float* s = sc + static_cast<long long>(*step) * CCEDIT_SAMPLER_ROW;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = x_euler[i];
  if (s[CCEDIT_SC_EULER_ONLY] == 0.f) {
    const float d2 = cfg_denoised(eps2[i], eps2[n + i], x2[i], s[CCEDIT_SC_COUT2], cfg_scale);
    const float vd = __fsub_rn(__fmul_rn(s[CCEDIT_SC_M3], x[i]), __fmul_rn(s[CCEDIT_SC_M4], d2));
    if (s[CCEDIT_SC_SIGMA_DOWN] > 0.f) v = vd;             // torch.where(sigma_down > 0, x_dpmpp2s, x_euler)
  }
  if (s[CCEDIT_SC_NEXT_SIGMA] > 0.f)                       // torch.where(next_sigma > 0, x + noise * s_noise * sigma_up, x)
    v = __fadd_rn(v, __fmul_rn(__fmul_rn(noise[i], s_noise), s[CCEDIT_SC_SIGMA_UP]));
  x_out[i] = v;
}

static inline unsigned sampler_blocks(long long n, int B) {
  const long long m = n > 2 * B ? n : 2 * B;
  return static_cast<unsigned>((m + 255) / 256);
}

}  // namespace ccedit

using namespace ccedit;

extern "C" int ccedit_sampler_prepare(const float* x, float* xin2, int64_t* t2, const float* sc, const int32_t* step,
                                      int64_t n, int32_t B, void* stream) {
  CCEDIT_CHECK_ARG(x && xin2 && t2 && sc && step && n > 0 && B > 0, "ccedit_sampler_prepare: bad arguments");
  sampler_prepare_kernel<<<sampler_blocks(n, B), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, xin2, reinterpret_cast<long long*>(t2), sc, step, n, B);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  CCEDIT_CUDA_LAUNCH_CHECK("ccedit_sampler_prepare");
  return CCEDIT_OK;
}

extern "C" int ccedit_sampler_mid(const float* x, const float* eps2, float* x_euler, float* x2, float* xin2, int64_t* t2,
                                  const float* sc, const int32_t* step, float cfg_scale, int64_t n, int32_t B, void* stream) {
  CCEDIT_CHECK_ARG(x && eps2 && x_euler && x2 && xin2 && t2 && sc && step && n > 0 && B > 0, "ccedit_sampler_mid: bad arguments");
  sampler_mid_kernel<<<sampler_blocks(n, B), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, eps2, x_euler, x2, xin2, reinterpret_cast<long long*>(t2), sc, step, cfg_scale, n, B);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  CCEDIT_CUDA_LAUNCH_CHECK("ccedit_sampler_mid");
  return CCEDIT_OK;
}

extern "C" int ccedit_sampler_final(const float* x, const float* x2, const float* x_euler, const float* eps2, const float* noise,
                                    float* x_out, const float* sc, const int32_t* step, float cfg_scale, float s_noise,
                                    int64_t n, void* stream) {
  CCEDIT_CHECK_ARG(x && x2 && x_euler && eps2 && noise && x_out && sc && step && n > 0, "ccedit_sampler_final: bad arguments");
  sampler_final_kernel<<<sampler_blocks(n, 0), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, x2, x_euler, eps2, noise, x_out,
                                                                                            sc, step, cfg_scale, s_noise, n);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  CCEDIT_CUDA_LAUNCH_CHECK("ccedit_sampler_final");
  return CCEDIT_OK;
}

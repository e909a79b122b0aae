// Shared device helpers for the ccedit_b200 sm_100a kernels: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (UMMA + TMEM) PTX wrappers, error plumbing.  Written for sm_100a only.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace ccedit {

// ---------------------------------------------------------------------------------------------
// host-side error plumbing (C ABI returns int status; message kept per thread)
// ---------------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);
#define CCEDIT_OK 0
#define CCEDIT_ERR_INVALID 1
#define CCEDIT_ERR_CUDA 2

#define CCEDIT_CHECK_ARG(cond, ...)                  \
  do {                                               \
    if (!(cond)) {                                   \
      ::ccedit::set_last_error(__VA_ARGS__);         \
      return CCEDIT_ERR_INVALID;                     \
    }                                                \
  } while (0)

#define CCEDIT_CUDA_LAUNCH_CHECK(what)                                                   \
  do {                                                                                   \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) {                                                             \
      ::ccedit::set_last_error("%s: CUDA error: %s", what, cudaGetErrorString(_e));      \
      return CCEDIT_ERR_CUDA;                                                            \
    }                                                                                    \
  } while (0)

constexpr int kMaxDevices = 64;   // per-device caches (function attributes, SM counts) are arrays of this size
int current_device();             // cudaGetDevice() ordinal, -1 if unavailable / out of range
int device_sm_count();

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// small utilities
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// Programmatic dependent launch (PDL): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start
// while its predecessor in the stream is still running; griddepcontrol.wait blocks until that predecessor has completed
// and its writes are visible, griddepcontrol.launch_dependents lets the NEXT kernel be scheduled early.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
int pdl_mask();                // gemm_tc.cu: CCEDIT_PDL bit mask (read once): 1 tap-GEMM, 2 flash attention, 4 short-key / temporal attention, 8 GroupNorm
inline bool pdl_enabled(int bit = 1) { return (pdl_mask() & bit) != 0; }
// <<<grid, block, smem, stream>>> with the PDL attribute when enabled; the kernel must call griddep_wait() before it touches
// anything a predecessor may have written
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(int bit, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  const bool on = pdl_enabled(bit);
  cfg.attrs = on ? &attr : nullptr;
  cfg.numAttrs = on ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// SiLU as x * rcp(1 + 2^(-x log2 e)) on the two approximate MUFU ops (~2 ulp): 5 instructions.  The IEEE '/' drags a
// slow-path call into every use, and __fdividef adds a range fix-up (FSETP + 2 FMUL) for denominators above 2^126 that
// 1 + e^-x only reaches where the result is 0 either way (ncu of the GroupNorm kernels: FMUL 21 % of all instructions).
__device__ __forceinline__ float silu_f(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return x * r;
}
// exact-erf GELU (torch F.gelu default, reference attention.py:120-122): gelu(x) = x * Phi(x).
// Phi(-|x|) = exp2(q(-|x|)) with q a degree-6 minimax fit of log2(Phi) on [-5.5, 0] (Lawson iteration, tools/fit_gelu.py):
// relative error of Phi <= 2.7e-5 INCLUDING the left tail (the erf form 0.5*(1+erf) cancels there), max |gelu error|
// 4e-6; Phi(x > 0) = 1 - Phi(-x).  9 FMA/ALU-pipe instructions + one MUFU.EX2 per value - the erf form by Abramowitz &
// Stegun 7.1.26 used before cost 17 + two MUFU and made the GEGLU epilogue (not the MMAs) pace the FF-in GEMM at K = 320.
__device__ __forceinline__ float gelu_erf_f(float x) {
  const float a = fmaxf(-fabsf(x), -5.5f);
  float q = fmaf(2.615377861722895e-05f, a, 6.609817238438444e-04f);
  q = fmaf(q, a, 7.488313388526632e-03f);
  q = fmaf(q, a, 5.197041696700614e-02f);
  q = fmaf(q, a, -4.6032946090714766e-01f);
  q = fmaf(q, a, 1.1505840052884075f);
  q = fmaf(q, a, -1.00003606355367f);
  float phi;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(phi) : "f"(q));
  return fmaf(-fabsf(x), phi, fmaxf(x, 0.f));      // x > 0: x - x Phi(-x);  x < 0: x Phi(x) = -|x| Phi(-|x|)
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Non-blocking probe (try_wait may suspend the thread for a hardware time slice before answering "not yet")
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must not hang the GPU (a hung box is a lost lease); trap instead.
#ifndef CCEDIT_MBAR_SPIN_LIMIT
#define CCEDIT_MBAR_SPIN_LIMIT (1u << 21)   // >= 0.2 s of polling; no legitimate wait outlives its kernel (ms)
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > CCEDIT_MBAR_SPIN_LIMIT) __trap();   // surfaces as a launch failure instead of a hung GPU
  }
}

// ---------------------------------------------------------------------------------------------
// TMA
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, "
      "%7}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// TMA store (shared -> global, bulk async-group completion) and its group bookkeeping; issued by ONE thread, which is also
// the thread that later waits for the group (bulk groups are per-thread).
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3,
                                             int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
      ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() {   // at most N groups still READING their shared-memory source
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// Warp-collective forms of the TMA store / bulk-group / mbarrier operations of the staged GEMM epilogue: the whole
// converged warp calls them with warp-uniform operands and the elected lane issues (the same lane every time: bulk
// async-groups belong to the issuing thread).  From inside `if (lane == 0)` each of these costs an ELECT / R2UR loop.
__device__ __forceinline__ void tma_store_5d_warp(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3,
                                                  int c4) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];\n\t}"
      ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void bulk_commit_group_warp() {
  asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t@e cp.async.bulk.commit_group;\n\t}" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_group_read_warp() {
  asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t@e cp.async.bulk.wait_group.read %0;\n\t}" ::"n"(N)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_warp(uint64_t* bar) {
  asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t@e mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster_warp(uint32_t cluster_addr) {
  asm volatile(
      "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t@e mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n\t}"
      ::"r"(cluster_addr)
      : "memory");
}

// warp-collective TMA issue (see umma_*_warp below): whole converged warp calls, one elected lane issues
__device__ __forceinline__ void mbar_arrive_expect_tx_warp(uint64_t* bar, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_warp(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d_warp(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                 int c2, int c3, int c4) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, "
      "%7}], [%2];\n\t}" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05: TMEM alloc, UMMA, commit, ld
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T ; kind::f16 (fp16/bf16 in, fp32 accumulate). One thread issues.
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// Warp-collective variants: the WHOLE (converged) warp calls them with warp-uniform operands and one elected lane
// issues.  Issuing from inside an `if (lane == 0)` region instead makes the compiler wrap every tcgen05 instruction in
// an ELECT / R2UR.BROADCAST / branch loop (the operands must live in uniform registers), ~100 clocks per MMA - measured
// with ccedit_gemm_trace: 7 tcgen05 instructions took ~1000 clocks to issue.
__device__ __forceinline__ void umma_f16_ss_warp(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16_ts_warp(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_warp(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}
// ---- thread-block clusters: weight tiles are TMA-multicast to the CTAs of a cluster (tap_gemm_kernel<..., CL = 2>) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {       // every thread of every CTA of the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// ---- CTA pairs (tcgen05 cta_group::2): one MMA instruction of the leader CTA drives the tensor cores of both SMs ----
__device__ __forceinline__ uint32_t mapa_shared(uint32_t cta_addr, uint32_t rank) {   // same offset in CTA `rank` of the cluster
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {          // arrive on a (possibly remote) barrier
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {  // one warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA loads of a CTA pair: data lands in THIS CTA's shared memory, the transaction bytes are signalled on `bar_cluster`,
// a shared::cluster address that may belong to the peer (the leader's `full` barrier)
__device__ __forceinline__ void tma_load_2d_pair_warp(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d_pair_warp(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1,
                                                      int c2, int c3, int c4) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, "
      "%7}], [%2];\n\t}" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
// D[tmem, both CTAs] (+)= A[smem of both CTAs: 128 rows each] * B[smem of both CTAs: N/2 rows each]^T, issued by the leader
__device__ __forceinline__ void umma_f16_ss_pair_warp(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                      uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair_warp(uint64_t* bar, uint16_t mask) {   // arrives in every CTA of `mask`
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
// 2-D tile load delivered to the same shared-memory offset (and signalling the same mbarrier offset) in every CTA of `mask`
__device__ __forceinline__ void tma_load_2d_multicast_warp(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                           uint16_t mask) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], "
      "[%2], %5;\n\t}" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// tcgen05.commit arriving on the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_multicast_warp(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32b, 16 consecutive columns: thread i of the warp gets lane (base_lane+i), cols [c, c+16)
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// UMMA shared-memory descriptor for a K-major tile stored as rows of 128 bytes with the 128B swizzle
// (what TMA SWIZZLE_128B writes): 8-row groups are 1024 B apart (SBO), LBO unused (=1), version 1 (sm_100),
// layout type 2 (SWIZZLE_128B).  Field layout: cute/arch/mma_sm100_desc.hpp (SmemDescriptor).
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);  // start address, 16-byte units   [0,14)
  d |= static_cast<uint64_t>(1) << 16;                      // leading byte offset (ignored)  [16,30)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;              // stride byte offset = 1024 B    [32,46)
  d |= static_cast<uint64_t>(1) << 46;                      // descriptor version (sm_100)    [46,48)
  d |= static_cast<uint64_t>(2) << 61;                      // SWIZZLE_128B                   [61,64)
  return d;
}
// kind::f16 instruction descriptor: fp16 A/B (K-major both), fp32 D, M=128, N=n.
__host__ __device__ __forceinline__ uint32_t umma_idesc_f16(uint32_t m, uint32_t n) {
  uint32_t d = 0;
  d |= 1u << 4;          // D format: F32
  d |= 0u << 7;          // A format: F16
  d |= 0u << 10;         // B format: F16
  d |= 0u << 15;         // A K-major
  d |= 0u << 16;         // B K-major
  d |= (n >> 3) << 17;   // N / 8
  d |= (m >> 4) << 24;   // M / 16
  return d;
}

// ---------------------------------------------------------------------------------------------
// warp-level mma.sync m16n8k16 (fp16 in, fp32 acc) + ldmatrix, cp.async : used by the attention kernels
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_m16n8k16(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void cp_async_16(uint32_t smem_addr, const void* gptr, bool valid) {
  int sz = valid ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr), "l"(gptr), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
#endif  // __CUDACC__

}  // namespace ccedit

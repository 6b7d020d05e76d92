// common.cuh -- shared helpers for librlb200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "rlb200.h"

namespace rlb {

// ---- error plumbing (thread-local message, negative return codes; never throws) ----------------
void set_error(const char *fmt, ...);
int check_cuda(cudaError_t e, const char *what);
int check_launch(const char *kernel);
int sm_count();

#define RLB_REQUIRE(cond, code, ...)    \
  do {                                  \
    if (!(cond)) {                      \
      ::rlb::set_error(__VA_ARGS__);    \
      return (code);                    \
    }                                   \
  } while (0)

static inline cudaStream_t as_stream(rlb_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

bool pdl_enabled();  // RLB_PDL=0 in the environment turns programmatic dependent launch off (A/B measurements)

// <<<grid, block, smem, st>>> with the programmatic-stream-serialization attribute: the kernel may be scheduled while
// the previous kernel of the stream is still running and must call pdl_wait() before touching its results.  Under
// stream capture this becomes a programmatic dependency edge of the graph.
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- device helpers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// L2-coherent loads (bypass L1): used where another thread of the same launch may have written the line.
__device__ __forceinline__ float ld_cg(const float *p) { return __ldcg(p); }
__device__ __forceinline__ double ld_cg(const double *p) { return __ldcg(p); }

// Non-contracted arithmetic: the tree/sampler paths must round every multiply/add/subtract separately,
// exactly as the reference's scalar C++ (and separate torch kernels) do.  nvcc would otherwise be free
// to fuse  u*p_sum - left  into one FMA.
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }

template <typename T, bool IsMin>
__device__ __forceinline__ T tree_op(T a, T b) {
  // std::plus<T> / MinOp = std::min(lhs, rhs) -> (rhs < lhs) ? rhs : lhs   (csrc/segment_tree.h:266-298)
  if constexpr (IsMin) {
    return (b < a) ? b : a;
  } else {
    return add_rn(a, b);
  }
}

template <typename T>
struct Limits;
template <>
struct Limits<float> {
  __host__ __device__ static constexpr float max() { return 3.402823466e+38f; }
};
template <>
struct Limits<double> {
  __host__ __device__ static constexpr double max() { return 1.7976931348623158e+308; }
};

// ---- mbarrier / bulk-async-copy (TMA engine, non-tensor form) PTX ---------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint64_t *bar, uint32_t parity) {
  // try_wait with a suspend-time hint: the waiting thread is parked by the hardware (up to the hint, in ns) instead of
  // spinning through the issue slots of the warps that share its scheduler -- the row mover's elected lanes wait for
  // microseconds at a time next to the latency-critical tree kernels
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "RLB_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra RLB_DONE_%=;\n"
      "bra RLB_WAIT_%=;\n"
      "RLB_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(20000u)
      : "memory");
}
// global -> shared bulk copy; completion (bytes) is signalled on `bar`.  16-B aligned, size % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// same, with an L2 cache-policy hint (createpolicy): replay rows are read once -> evict_first keeps them from
// displacing the L2-resident segment trees.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar,
                                              uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
// shared -> global bulk copy, tracked by the thread's bulk async-group.
__device__ __forceinline__ void bulk_s2g(void *gdst, const void *smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// programmatic dependent launch (PDL): a kernel launched with the attribute may start while its predecessor in the
// stream is still draining; everything it does before pdl_wait() must not depend on the predecessor's results.
// pdl_trigger() in the predecessor lets the dependent's CTAs become resident early.  Both are no-ops for plain launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

}  // namespace rlb

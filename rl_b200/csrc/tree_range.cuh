// tree_range.cuh -- priority write for a CONTIGUOUS (modular) range of slots: the write path's mark_update.
//
// RoundRobinWriter.extend (writers.py:190-216) always writes slots  (cursor + arange(n)) % max_size  and then gives
// every one of them the SAME default priority (writers.py:232-235 -> samplers.py:1093-1096 -> update_priority).  The
// general update kernel (tree.cu) sorts and merges arbitrary indices; for a range none of that is needed:
//
//   * a node whose whole leaf span lies inside the range has a closed form: min = v and sum = v * 2^level.  That IS
//     what the reference computes (v + v, 2v + 2v, ... are exact in binary floating point), so it is bit-identical.
//     These ~2n nodes are plain coalesced stores, spread over as many CTAs as the caller gives the role;
//   * every other ancestor of a written leaf is an ancestor of one of the (at most four) boundary leaves of the (at
//     most two, when the range wraps) pieces: <= 4 nodes per level.  Their values form one serial chain from the
//     leaves to the root,  node = op(child0, child1),  where a child is a closed-form node, a node the chain produced
//     one level below, or an UNTOUCHED node whose old value was prefetched (all <= 8 * depth of them in one parallel
//     round trip, by the 4 x depth threads that each own one (boundary leaf, level) pair).
//
// Neither part waits for the other (the chain never reads a node the fill writes), so there is no grid-wide sync.
// The default priority itself ((max_priority + eps) ** alpha, samplers.py:886-893), its second pow (the reference's
// double-pow quirk, :1076) and the running max (:1054-1075) are computed in the kernel; since the new max depends on
// the old one, the CTA that takes the last ticket publishes it after every CTA has read the old value.
#pragma once

#include <math.h>
#include <string.h>

#include "common.cuh"

namespace rlb {

// torch.pow(x, scalar) semantics for fp32 tensors on CUDA (ATen/native/cuda/PowKernel.cu): dedicated
// kernels for 0.5 / -0.5 / -1, x*x, x*x*x, 1/(x*x) for 2 / 3 / -2, ::pow otherwise; exponent 0 -> 1,
// exponent 1 -> copy (ATen/native/Pow.cpp).  The exponent is cast to the tensor dtype first.
__device__ __forceinline__ float pow_like_torch(float x, float y) {
  if (y == 0.0f) return 1.0f;
  if (y == 1.0f) return x;
  if (y == 0.5f) return sqrtf(x);
  if (y == -0.5f) return rsqrtf(x);
  if (y == -1.0f) return 1.0f / x;
  if (y == 2.0f) return mul_rn(x, x);
  if (y == 3.0f) return mul_rn(mul_rn(x, x), x);
  if (y == -2.0f) return (float)(1.0 / (double)mul_rn(x, x));
  return powf(x, y);
}
__device__ __forceinline__ double pow_like_torch(double x, double y) {
  if (y == 0.0) return 1.0;
  if (y == 1.0) return x;
  if (y == 0.5) return sqrt(x);
  if (y == -0.5) return rsqrt(x);
  if (y == -1.0) return 1.0 / x;
  if (y == 2.0) return x * x;
  if (y == 3.0) return x * x * x;
  if (y == -2.0) return 1.0 / (x * x);
  return pow(x, y);
}

// max of floats through one fire-and-forget reduction: non-negative values order like their int bits, negative
// ones like their reversed unsigned bits (the buffer starts at -inf = 0xff800000, below / above all of them).
__device__ __forceinline__ void red_max_float(float *addr, float v) {
  if (v >= 0.0f)
    atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
  else
    atomicMin(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}

// ------------------------------------------------------------------------------------------------
enum RangeMode : int {
  kRangeValue = 0,     // *value (tree dtype) is the final leaf value
  kRangePriority = 1,  // *value (fp32) is a RAW priority: leaf = (p + eps) ** alpha, running max <- max(., p)
  kRangeDefault = 2,   // p = has_max ? (max + eps) ** alpha : first_default  (samplers.py:886-893), then as above
};

struct RangeParams {
  void *sum;  // either tree may be null
  void *mn;
  int64_t capacity;
  int depth;
  int mode;
  int64_t s[2], e[2];    // leaf ranges [s, e); e <= s means empty
  const void *value;     // device scalar (modes 0, 1)
  float *max_buf;        // running max of raw priorities (mode 2; optional in mode 1)
  unsigned int *ticket;  // zero-initialised word, left at zero (mode 2 with more than one CTA)
  float alpha, eps, first_default;
  int has_max;
};

constexpr int kRangeThreads = 256;
constexpr int kRangeMaxDepth = 32;  // one warp lane per level

__device__ __forceinline__ float scale_pow2(float v, int e) { return scalbnf(v, e); }
__device__ __forceinline__ double scale_pow2(double v, int e) { return scalbn(v, e); }

__device__ __forceinline__ bool range_covers(const RangeParams &R, int64_t node, int level) {
  const int64_t lo = (node << level) - R.capacity, hi = lo + (int64_t(1) << level);
  return (lo >= R.s[0] && hi <= R.e[0]) || (lo >= R.s[1] && hi <= R.e[1]);
}
// the k-th boundary leaf: s0, e0 - 1, s1, e1 - 1
__device__ __forceinline__ int64_t range_edge_leaf(const RangeParams &R, int k) {
  return (k & 1) ? R.e[k >> 1] - 1 : R.s[k >> 1];
}

// CTA `cta` of `nctas` (kRangeThreads threads each) cooperating on one range update.
template <typename T>
__device__ __forceinline__ void range_role(const RangeParams &R, int cta, int nctas) {
  T *sum = static_cast<T *>(R.sum), *mn = static_cast<T *>(R.mn);
  const int tid = threadIdx.x;
  // ---- the leaf value
  T v;
  if (R.mode == kRangeValue) {
    v = __ldg(static_cast<const T *>(R.value));
  } else {
    float p;
    if (R.mode == kRangePriority) {
      p = __ldg(static_cast<const float *>(R.value));
    } else {
      // "no priority seen yet" is decided on the device (the running max is still -inf), not by the host's flag, so
      // that a launch captured in a CUDA graph stays right when it is replayed after priorities have arrived.
      // ONE thread reads the running max and the CTA shares it: the ticket below (taken by thread 0 after the barrier)
      // then really means "every thread of this CTA has its copy of the OLD max" -- with per-thread reads the last
      // CTA could publish the new max while warps of an earlier CTA had not loaded yet, and those warps would fill
      // their nodes with a different leaf value.
      __shared__ float s_mx;
      if (tid == 0) s_mx = ld_cg(R.max_buf);
      __syncthreads();
      const float mx = s_mx;
      p = (R.has_max && mx > -INFINITY) ? pow_like_torch(add_rn(mx, R.eps), R.alpha) : R.first_default;
    }
    v = (T)pow_like_torch(add_rn(p, R.eps), R.alpha);
    if (R.max_buf && tid == 0) {  // publish the new running max once every CTA has read the old one
      bool last = cta == 0;
      if (nctas > 1 && R.mode == kRangeDefault) {
        __threadfence();
        last = atomicAdd(R.ticket, 1u) == (unsigned)nctas - 1u;
        if (last) *R.ticket = 0u;
      }
      if (last && p > -INFINITY) red_max_float(R.max_buf, p);
    }
  }
  // ---- closed-form nodes: level by level until a level has none
  const int64_t gthreads = (int64_t)nctas * kRangeThreads, gtid = (int64_t)cta * kRangeThreads + tid;
  for (int r = 0; r < 2; ++r) {
    if (R.e[r] <= R.s[r]) continue;
    T c = v;
    for (int l = 0; l < R.depth; ++l) {
      const int64_t base = R.capacity >> l;
      const int64_t first = base + ((R.s[r] + (int64_t(1) << l) - 1) >> l), last = base + (R.e[r] >> l);
      if (first >= last) break;
      for (int64_t i = first + gtid; i < last; i += gthreads) {
        if (sum) sum[i] = c;
        if (mn) mn[i] = v;
      }
      c = add_rn(c, c);
    }
  }
  if (cta != 0) return;
  // ---- boundary chain (CTA 0, 4 warps): warp k = boundary leaf k, lane = level - 1.  Whether a thread's node exists,
  // and which of its children are closed-form, depends only on the range, so every thread works that out (and loads
  // the untouched children) in parallel; only the values are serial: one shared-memory exchange + barrier per level.
  __shared__ long long x_node[2][4];  // nodes produced at the previous / current level (-1: none)
  __shared__ T x_sum[2][4], x_min[2][4];
  const int k = tid >> 5, l = (tid & 31) + 1;
  const bool mine = k < 4 && l <= R.depth && R.e[(k & 3) >> 1] > R.s[(k & 3) >> 1];
  int64_t node = -1;
  bool act = false;
  T imm_s[2] = {(T)0, (T)0}, imm_m[2] = {(T)0, (T)0};
  bool closed[2] = {false, false};
  if (mine) {
    const int64_t edge = R.capacity + range_edge_leaf(R, k);
    node = edge >> l;
    act = !range_covers(R, node, l);
    for (int j = 0; j < k; ++j)  // boundary leaves sharing this ancestor: the lower warp keeps it
      if (R.e[j >> 1] > R.s[j >> 1] && ((R.capacity + range_edge_leaf(R, j)) >> l) == node) act = false;
    if (act) {
      const T cst = scale_pow2(v, l - 1);  // closed-form sum one level below (v + v, 2v + 2v, ... are exact)
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const int64_t child = (node << 1) | ch;
        closed[ch] = range_covers(R, child, l - 1);
        imm_s[ch] = closed[ch] ? cst : (sum ? ld_cg(sum + child) : (T)0);  // (an untouched node unless the level
        imm_m[ch] = closed[ch] ? v : (mn ? ld_cg(mn + child) : (T)0);      //  below produced it: resolved in the loop)
      }
    }
  }
  if (tid < 8) x_node[tid >> 2][tid & 3] = -1;
  __syncthreads();
  T out_s = (T)0, out_m = (T)0;
  for (int step = 1; step <= R.depth; ++step) {
    const int cur = step & 1, prv = cur ^ 1;
    if (k < 4 && l == step) {
      if (act) {
        T cs[2], cm[2];
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          const int64_t child = (node << 1) | ch;
          cs[ch] = imm_s[ch];
          cm[ch] = imm_m[ch];
          if (!closed[ch]) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (x_node[prv][j] == child) {
                cs[ch] = x_sum[prv][j];
                cm[ch] = x_min[prv][j];
              }
            }
          }
        }
        out_s = tree_op<T, false>(cs[0], cs[1]);
        out_m = tree_op<T, true>(cm[0], cm[1]);
        x_sum[cur][k] = out_s;
        x_min[cur][k] = out_m;
      }
      x_node[cur][k] = act ? node : -1;
    }
    __syncthreads();
  }
  if (act) {
    if (sum) sum[node] = out_s;
    if (mn) mn[node] = out_m;
  }
}

// host side: leaf ranges of the slots (start + arange(n)) % modulo, n <= modulo
static inline void range_split(int64_t start, int64_t n, int64_t modulo, int64_t s[2], int64_t e[2]) {
  if (start + n <= modulo) {
    s[0] = start, e[0] = start + n, s[1] = e[1] = 0;
  } else {
    s[0] = start, e[0] = modulo, s[1] = 0, e[1] = start + n - modulo;
  }
}

static inline int range_params(RangeParams &R, const char *who, void *sum_tree, void *min_tree, int64_t capacity,
                               int dtype, int64_t start, int64_t n, int64_t modulo, int mode, const void *value,
                               double alpha, double eps, double first_default, int has_max, float *max_priority,
                               uint32_t *ticket) {
  RLB_REQUIRE((sum_tree || min_tree) && capacity > 0 && (capacity & (capacity - 1)) == 0, RLB_EINVAL,
              "%s: null trees or bad capacity", who);
  RLB_REQUIRE(dtype == RLB_F32 || dtype == RLB_F64, RLB_EINVAL, "%s: unsupported dtype %d", who, dtype);
  RLB_REQUIRE(modulo > 0 && modulo < capacity && start >= 0 && start < modulo && n >= 0 && n <= modulo, RLB_EINVAL,
              "%s: range start=%lld n=%lld outside modulo=%lld (capacity %lld)", who, (long long)start, (long long)n,
              (long long)modulo, (long long)capacity);
  RLB_REQUIRE(mode >= kRangeValue && mode <= kRangeDefault, RLB_EINVAL, "%s: unknown mode %d", who, mode);
  RLB_REQUIRE(mode == kRangeDefault || value, RLB_EINVAL, "%s: null value", who);
  RLB_REQUIRE(mode != kRangeDefault || (max_priority && ticket), RLB_EINVAL,
              "%s: the default-priority mode needs max_priority and ticket", who);
  memset(&R, 0, sizeof(R));
  R.sum = sum_tree;
  R.mn = min_tree;
  R.capacity = capacity;
  R.depth = 0;
  while ((int64_t(1) << R.depth) < capacity) ++R.depth;
  RLB_REQUIRE(R.depth <= kRangeMaxDepth, RLB_ELIMIT, "%s: tree too deep", who);
  R.mode = mode;
  range_split(start, n, modulo, R.s, R.e);
  R.value = value;
  R.max_buf = max_priority;
  R.ticket = ticket;
  R.alpha = (float)alpha;
  R.eps = (float)eps;
  R.first_default = (float)first_default;
  R.has_max = has_max;
  return RLB_OK;
}

static inline int range_ctas(int64_t n) {
  const int sms = sm_count();
  int64_t c = (2 * n + kRangeThreads * 8 - 1) / (kRangeThreads * 8);
  if (c > sms) c = sms;
  return c < 1 ? 1 : (int)c;
}

}  // namespace rlb

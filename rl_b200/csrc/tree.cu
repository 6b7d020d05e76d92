// tree.cu -- HBM-resident sum/min segment trees for prioritized replay (sm_100a).
//
// Replaces the reference's CPU tree (csrc/segment_tree.h:41-307) and its correctness-first CUDA port
// (csrc/cuda_segment_tree.cu:26-208).  The node layout and the value of every node are bit-identical
// to the reference (node k = op(node 2k, node 2k+1), one fp rounding per node, no FMA contraction), so
// the indices returned by the prefix-sum descent are exactly the reference's for the same draws.
// What changes is scheduling:
//   * sample  -- one launch fuses query(0,len) on both trees, mass = u*p_sum, the lower-bound descent,
//                clamp, leaf read and the importance weight.  The descent resolves FOUR tree levels per
//                memory round trip: the 2/4/8/16 descendants of a node sit in four contiguous, 16-B
//                aligned runs of the heap array, so they are fetched with independent 128-bit loads and
//                the four comparisons run from registers (5 dependent round trips for a 2^20 tree
//                instead of 20).
//   * update  -- last-writer-wins leaf scatter + recomputation of the TOUCHED ancestors only.  Batches
//                of <= 1024 run in ONE CTA: paths climb level-synchronously, exchanging child values
//                through a shared-memory hash keyed by parent id (one __syncthreads per level, no
//                global round trip between levels except the prefetch of untouched siblings); paths
//                that merge are carried on by a single thread.  Larger batches use a stamp-dedupe
//                scatter + one sweep launch per level + a single-CTA dense top.
// These are latency-bound pointer chases over an L2-resident array (16 MB for two 2^20 fp32 trees):
// no tensor cores, no smem tiling of the tree itself.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "tree_range.cuh"

namespace rlb {

long long *g_debug_ticks = nullptr;  // set by rlb_debug_set_tick_buffer (profiling only)

// ------------------------------------------------------------------------------------------------
// fill / rebuild / query / at
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void tree_fill_kernel(T *tree, int64_t n, T v) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) tree[i] = v;
}

template <typename T, bool IsMin>
__global__ void tree_level_dense_kernel(T *tree, int64_t level_start, int64_t level_count) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < level_count; i += stride) {
    const int64_t node = level_start + i;
    tree[node] = tree_op<T, IsMin>(tree[node << 1], tree[(node << 1) | 1]);
  }
}

// Dense recomputation of eleven levels at once: each CTA takes kRebuildTile consecutive leaves, reduces its subtree
// pairwise in shared memory (same child pairs, same operation order as the level-by-level form) and writes every node
// of it -- one launch instead of eleven for the bulk of a rebuild (load_state_dict, loads, the masked copy the
// prioritized slice sampler draws from).
constexpr int kRebuildTile = 2048;
template <typename T, bool IsMin>
__global__ void __launch_bounds__(256) tree_rebuild_tiles_kernel(T *tree, int64_t capacity) {
  __shared__ T buf[2][kRebuildTile];
  const int64_t leaf0 = capacity + (int64_t)blockIdx.x * kRebuildTile;
  for (int j = threadIdx.x; j < kRebuildTile; j += blockDim.x) buf[0][j] = tree[leaf0 + j];
  __syncthreads();
  int cur = 0;
  int64_t level_start = capacity >> 1;  // first node of the level being produced
  for (int w = kRebuildTile >> 1; w >= 1; w >>= 1, level_start >>= 1) {
    const int64_t node0 = level_start + (int64_t)blockIdx.x * w;
    for (int j = threadIdx.x; j < w; j += blockDim.x) {
      const T v = tree_op<T, IsMin>(buf[cur][2 * j], buf[cur][2 * j + 1]);
      buf[cur ^ 1][j] = v;
      tree[node0 + j] = v;
    }
    __syncthreads();
    cur ^= 1;
  }
}

// Dense recomputation of the top of the tree by ONE CTA: loads the W nodes [W, 2W) (W <= 1024, a
// power of two), reduces pairwise in shared memory and writes every node in [1, W).
template <typename T, bool IsMin>
__device__ __forceinline__ void tree_top_dense(T *tree, int W, T *sh) {
  const int tid = threadIdx.x;
  if (tid < W) sh[W + tid] = ld_cg(tree + W + tid);
  __syncthreads();
  for (int w = W >> 1; w >= 1; w >>= 1) {
    if (tid < w) {
      const int node = w + tid;
      const T v = tree_op<T, IsMin>(sh[node << 1], sh[(node << 1) | 1]);
      sh[node] = v;
      tree[node] = v;
    }
    __syncthreads();
  }
}

template <typename T>
__global__ void __launch_bounds__(1024) tree_top_kernel(T *sum, T *mn, int W) {
  __shared__ T sh[2048];
  if (sum) tree_top_dense<T, false>(sum, W, sh);
  __syncthreads();
  if (mn) tree_top_dense<T, true>(mn, W, sh);
}

// Reduce [l, r) exactly as SegmentTree::Query does (csrc/segment_tree.h:143-162): terms are folded
// into `ret` bottom-up, left term before right term at each level.
template <typename T, bool IsMin>
__device__ __forceinline__ T tree_query_walk(const T *__restrict__ tree, int64_t capacity, int64_t l, int64_t r,
                                             T identity) {
  T ret = identity;
  l |= capacity;
  r |= capacity;
  while (l < r) {
    if (l & 1) ret = tree_op<T, IsMin>(ret, tree[l++]);
    if (r & 1) ret = tree_op<T, IsMin>(ret, tree[--r]);
    l >>= 1;
    r >>= 1;
  }
  return ret;
}

template <typename T, bool IsMin>
__global__ void tree_query_kernel(const T *__restrict__ tree, int64_t size, int64_t capacity,
                                  const int64_t *__restrict__ l, const int64_t *__restrict__ r, T *out, int64_t n,
                                  int root_fast_path, T identity) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t li = l[i], ri = r[i];
  if (root_fast_path && li <= 0 && ri >= size) {
    out[i] = tree[1];
    return;
  }
  out[i] = tree_query_walk<T, IsMin>(tree, capacity, li, ri, identity);
}

template <typename T>
__global__ void tree_at_kernel(const T *__restrict__ tree, int64_t capacity, const int64_t *__restrict__ index,
                               T *out, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = tree[index[i] | capacity];
}

// ------------------------------------------------------------------------------------------------
// lower-bound descent (SumSegmentTree::ScanLowerBound, csrc/segment_tree.h:249-264)
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void descend_step(T &cur, int64_t &node, T lvalue) {
  // `if (current_value > lvalue) { current_value -= lvalue; index |= 1; }` with strict '>'
  if (cur > lvalue) {
    cur = sub_rn(cur, lvalue);
    node |= 1;
  }
}

// Generic one-level-per-load descent (any dtype, any remaining depth).
template <typename T>
__device__ __forceinline__ int64_t descend_plain(const T *__restrict__ tree, int64_t node, int levels, T &cur) {
  for (int k = 0; k < levels; ++k) {
    node <<= 1;
    descend_step(cur, node, __ldg(tree + node));
  }
  return node;
}

// fp32: four levels per round trip.  From `node`, the left children that the next four comparisons can
// possibly need are  tree[2n]; tree[4n+{0,2}]; tree[8n+{0,2,4,6}]; tree[16n+{0,2,..,14}]  -- each run is
// contiguous and 16-B aligned (for n >= 1), so they are fetched as 1 scalar + 1 + 2 + 4 independent
// LDG.128 and the data-dependent choice is made from registers.  The comparisons/subtractions are the
// reference's, in the reference's order.
__device__ __forceinline__ int64_t descend4_f32(const float *__restrict__ tree, int64_t node, float &cur,
                                                float &landed) {
  const float l1 = __ldg(tree + (node << 1));
  const float4 q2 = __ldg(reinterpret_cast<const float4 *>(tree + (node << 2)));
  const float4 q3a = __ldg(reinterpret_cast<const float4 *>(tree + (node << 3)));
  const float4 q3b = __ldg(reinterpret_cast<const float4 *>(tree + (node << 3) + 4));
  const float4 q4a = __ldg(reinterpret_cast<const float4 *>(tree + (node << 4)));
  const float4 q4b = __ldg(reinterpret_cast<const float4 *>(tree + (node << 4) + 4));
  const float4 q4c = __ldg(reinterpret_cast<const float4 *>(tree + (node << 4) + 8));
  const float4 q4d = __ldg(reinterpret_cast<const float4 *>(tree + (node << 4) + 12));
  // level +1
  int c = 0;
  if (cur > l1) {
    cur = sub_rn(cur, l1);
    c = 1;
  }
  // level +2: left child of (2n + c) is tree[4n + 2c]
  const float l2 = c ? q2.z : q2.x;
  c <<= 1;
  if (cur > l2) {
    cur = sub_rn(cur, l2);
    c |= 1;
  }
  // level +3: left child of (4n + c) is tree[8n + 2c], c in [0,4)
  float l3 = q3a.x;
  l3 = (c == 1) ? q3a.z : l3;
  l3 = (c == 2) ? q3b.x : l3;
  l3 = (c == 3) ? q3b.z : l3;
  c <<= 1;
  if (cur > l3) {
    cur = sub_rn(cur, l3);
    c |= 1;
  }
  // level +4: left child of (8n + c) is tree[16n + 2c], c in [0,8)
  float l4 = q4a.x;
  l4 = (c == 1) ? q4a.z : l4;
  l4 = (c == 2) ? q4b.x : l4;
  l4 = (c == 3) ? q4b.z : l4;
  l4 = (c == 4) ? q4c.x : l4;
  l4 = (c == 5) ? q4c.z : l4;
  l4 = (c == 6) ? q4d.x : l4;
  l4 = (c == 7) ? q4d.z : l4;
  c <<= 1;
  if (cur > l4) {
    cur = sub_rn(cur, l4);
    c |= 1;
  }
  // value of the node we end on (tree[16n + c]): it is one of the 16 floats already in registers, so when this
  // round reaches the leaves the caller needs no further load for the leaf priority
  const float4 qq = (c < 8) ? ((c < 4) ? q4a : q4b) : ((c < 12) ? q4c : q4d);
  const int e = c & 3;
  landed = (e == 0) ? qq.x : (e == 1) ? qq.y : (e == 2) ? qq.z : qq.w;
  return (node << 4) | c;
}

// fp32, a GROUP of G lanes per sample (small batches leave most lanes of the single-warp CTAs idle): up to
// log2(G) + 2 levels per round trip.  The left children the next K comparisons can possibly need are, level by level,
// the contiguous runs tree[(n << k) + 0 .. 2^k) -- 2^k / 4 aligned 16-byte units at level k >= 2, one 8-byte unit at
// level 1.  The deepest level's units go to lanes [0, 2^(K-2)) (register A), the levels above it are packed behind one
// another over the same lanes (register B: level k starts at lane 2^(K-2) - 2^(k-1)), so every lane issues at most two
// independent loads.  Then EVERY lane of the group walks the K comparisons -- the reference's, in the reference's order --
// fetching each candidate from its owner with a shuffle: `cur` and the path stay uniform across the group.
__device__ __forceinline__ int64_t descend_group_f32(const float *__restrict__ tree, int64_t node, int K, int G,
                                                  float &cur, float &landed) {
  // (deliberately compact: rolled loops, one copy -- a single-warp CTA runs this once, instruction fetch is part of its
  // latency)
  const int gl = threadIdx.x & (G - 1);
  const int nA = 1 << (K - 2);
  float4 qa = make_float4(0.f, 0.f, 0.f, 0.f), qb = qa;
  if (gl < nA) qa = __ldg(reinterpret_cast<const float4 *>(tree + (node << K)) + gl);
  {
    // my unit of the packed upper levels: lane gl holds level k with  nA - 2^(k-1) <= gl < nA - 2^(k-2)  (k >= 2),
    // lane nA - 1 holds level 1
    const int d = nA - gl;  // in [1, nA] for the lanes that hold something
    if (d >= 1 && K >= 2 && gl < nA) {
      if (d == 1) {
        const float2 t = __ldg(reinterpret_cast<const float2 *>(tree + (node << 1)));
        qb.x = t.x;
        qb.y = t.y;
      } else {
        const int k = 32 - __clz(d - 1) + 1;  // 2^(k-2) < d <= 2^(k-1)
        if (k < K) qb = __ldg(reinterpret_cast<const float4 *>(tree + (node << k)) + (gl - (nA - (1 << (k - 1)))));
      }
    }
  }
  int c = 0;
#pragma unroll 1
  for (int k = 1; k <= K; ++k) {
    const int e = c << 1;  // the left child of the node reached so far, as an element of level k's run
    const float4 q = (k == K) ? qa : qb;
    const int comp = e & 3;
    const float sel = comp == 0 ? q.x : (comp == 1 ? q.y : (comp == 2 ? q.z : q.w));
    const int src = (k == K ? 0 : nA - (1 << (k - 1))) + (e >> 2);
    const float l = __shfl_sync(0xffffffffu, sel, src, G);
    c = e;
    if (cur > l) {
      cur = sub_rn(cur, l);
      c |= 1;
    }
  }
  {  // value of the node we end on: element c of the deepest level's run
    const int comp = c & 3;
    const float sel = comp == 0 ? qa.x : (comp == 1 ? qa.y : (comp == 2 ? qa.z : qa.w));
    landed = __shfl_sync(0xffffffffu, sel, c >> 2, G);
  }
  return (node << K) | c;
}

// the remaining `rem` levels below `node`, cooperatively: rounds of at most log2(G) + 2 levels, evenly sized
__device__ __forceinline__ int64_t descend_group_all(const float *__restrict__ tree, int64_t node, int rem, int G,
                                                     float &cur, float &landed, bool &have_landed) {
  const int kMax = (G == 32 ? 5 : G == 16 ? 4 : 3) + 2;
  have_landed = false;
#pragma unroll 1
  while (rem >= 2) {
    const int rounds = (rem + kMax - 1) / kMax;
    const int K = (rem + rounds - 1) / rounds;
    if (K < 2) break;
    node = descend_group_f32(tree, node, K, G, cur, landed);
    rem -= K;
    have_landed = rem == 0;
  }
  if (rem > 0) {  // (a single left-over level)
    node = descend_plain<float>(tree, node, rem, cur);
    have_landed = false;
  }
  return node;
}

template <typename T>
__device__ __forceinline__ int64_t scan_lower_bound_one(const T *__restrict__ tree, int64_t size, int64_t capacity,
                                                        int depth, T value, T root, bool speculative = true) {
  if (value > root) return size;  // csrc/segment_tree.h:250-252
  int64_t node = 1;
  T cur = value;
  int lev = 0;
  if constexpr (sizeof(T) == 4) {
    if (speculative) {
      float landed;
      for (; lev + 4 <= depth; lev += 4) node = descend4_f32(tree, node, cur, landed);
    }
  }
  node = descend_plain<T>(tree, node, depth - lev, cur);
  return node ^ capacity;
}

template <typename T>
__global__ void tree_scan_kernel(const T *__restrict__ tree, int64_t size, int64_t capacity, int depth,
                                 const T *__restrict__ value, int64_t *out, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = scan_lower_bound_one<T>(tree, size, capacity, depth, value[i], __ldg(tree + 1));
}

// ------------------------------------------------------------------------------------------------
// fused PrioritizedSampler.sample arithmetic (samplers.py:895-956)
// ------------------------------------------------------------------------------------------------
// query(0, len) of one tree by one warp.  For l = 0 the walk only ever takes RIGHT terms:
// level j contributes tree[r_j - 1] iff r_j is odd, with r_j = (capacity + len) >> j, while l_j < r_j
// (l_j = capacity >> j).  The terms are independent loads (one lane per level), folded in level order
// by every lane redundantly so the association order is the reference's.
template <typename T, bool IsMin>
__device__ __forceinline__ T warp_query_prefix(const T *__restrict__ tree, int64_t capacity, int64_t len,
                                               T identity) {
  const int lane = threadIdx.x & 31;
  T ret = identity;
  // up to 64 levels in two passes of 32 lanes (capacity < 2^62)
  for (int base = 0; base < 64; base += 32) {
    const int j = base + lane;
    const int64_t lj = (j < 62) ? (capacity >> j) : 0;
    const int64_t rj = (j < 62) ? ((capacity + len) >> j) : 0;
    const bool take = (lj < rj) && (rj & 1);
    const T term = take ? __ldg(tree + (rj - 1)) : identity;
    const unsigned mask = __ballot_sync(0xffffffffu, take);
    for (int k = 0; k < 32; ++k) {
      const T t = __shfl_sync(0xffffffffu, term, k);
      if ((mask >> k) & 1u) ret = tree_op<T, IsMin>(ret, t);
    }
    if ((capacity >> (base + 31)) == 0) break;
  }
  return ret;
}

// Every lane chases its own root-to-leaf path, so each warp-level load touches 32 different lines and one SM's L1
// retires only about one such sector wavefront every ~2 cycles: with 128 samples per SM a 4-level round costs
// 2.6k cycles instead of one ~1k-cycle round trip (clock64-instrumented, profiles/README.md).  Small batches are
// therefore SPREAD: `spc` samples per 32-thread CTA, chosen so that the grid covers ~2 CTAs per SM (B=256 -> one
// sample per CTA on 256 CTAs), and use the 4-level speculative descent (latency-bound).  Large batches fill
// 128-thread CTAs and use the plain one-load-per-level descent: half the wavefronts, latency hidden by occupancy.
template <typename T>
__global__ void __launch_bounds__(128) per_sample_kernel(const T *__restrict__ sum, const T *__restrict__ mn,
                                                         int64_t size, int64_t capacity, int depth, int64_t len,
                                                         const T *__restrict__ u, int64_t B, T neg_beta,
                                                         int cpu_semantics, int speculative, int spc,
                                                         int64_t *__restrict__ index_out,
                                                         float *__restrict__ weight_out, T *leaf_out,
                                                         T *psum_pmin_out, int32_t *status, long long *dbg) {
#ifndef RLB_SAMPLE_TOP_LEVELS
#define RLB_SAMPLE_TOP_LEVELS 9
#endif
  constexpr int kTopLevels = RLB_SAMPLE_TOP_LEVELS;  // nodes 1..2^k-1 of the sum tree are staged in shared memory: k - 1 descent steps
  __shared__ T s_p[2];
  __shared__ T s_top[1 << kTopLevels];
  pdl_trigger();  // a dependent launch (the gather) may become resident now; it waits for this grid before the index
  pdl_wait();     // launched with PDL itself: the uniforms (torch.rand) and the trees are final from here on
  if (dbg && (blockIdx.x != 0 || threadIdx.x != 0)) dbg = nullptr;
  if (dbg) dbg[8] = (long long)clock64();
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  // Single-warp CTAs with at most 4 samples: the warp splits into groups of G = 32 / 16 / 8 lanes, one per sample, that
  // descend cooperatively (descend_group_f32); only the first lane of a group writes results.
  const bool coop = speculative && sizeof(T) == 4 && blockDim.x == 32 && spc <= 4 && depth >= kTopLevels;
  const int G = !coop ? 1 : (spc == 1 ? 32 : (spc == 2 ? 16 : 8));
  const int grp = coop ? (int)threadIdx.x / G : (int)threadIdx.x;
  const bool lane_on = grp < spc;  // spc == blockDim.x for full CTAs
  const int64_t i = lane_on ? blockIdx.x * (int64_t)spc + grp : B;
  // ---- round 0: every load below is independent -- the draw, both roots (or the prefix walks) and the top of the tree
  const T ui = (i < B) ? __ldg(u + i) : (T)0;
  const bool use_top = speculative && depth >= kTopLevels && sizeof(T) == 4;
  if (use_top) {
    for (int k = threadIdx.x * 4; k < (1 << kTopLevels); k += blockDim.x * 4) {
      const float4 q = __ldg(reinterpret_cast<const float4 *>(sum + k));  // coalesced; slot 0 is never read
      s_top[k] = (T)q.x;
      s_top[k + 1] = (T)q.y;
      s_top[k + 2] = (T)q.z;
      s_top[k + 3] = (T)q.w;
    }
  }
  // p_sum by warp 0, p_min by warp 1 (or both by warp 0 in single-warp CTAs)   (samplers.py:901-908)
  const bool root = cpu_semantics && (len >= size);  // csrc/segment_tree.h:145-147 fast path (l == 0)
  if (warp == 0) {
    const T v = root ? __ldg(sum + 1) : warp_query_prefix<T, false>(sum, capacity, len, (T)0);
    if ((threadIdx.x & 31) == 0) s_p[0] = v;
  }
  if (warp == (nwarps > 1 ? 1 : 0)) {
    const T v = root ? __ldg(mn + 1) : warp_query_prefix<T, true>(mn, capacity, len, Limits<T>::max());
    if ((threadIdx.x & 31) == 0) s_p[1] = v;
  }
  __syncthreads();
  const T p_sum = s_p[0], p_min = s_p[1];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (psum_pmin_out) {
      psum_pmin_out[0] = p_sum;
      psum_pmin_out[1] = p_min;
    }
    if (status) {
      int32_t st = 0;
      if (!(p_sum > (T)0)) st |= RLB_STATUS_NONPOS_PSUM;
      if (!(p_min > (T)0)) st |= RLB_STATUS_NONPOS_PMIN;
      if (st) atomicOr(status, st);
    }
  }
  if (i >= B && !coop) return;  // (cooperating lanes stay for the shuffles; a group without a sample walks a dummy path)
  if (dbg) dbg[9] = (long long)clock64();
  const T mass = mul_rn(ui, p_sum);  // samplers.py:919 / :923 -- one rounding, never fused downstream
  int64_t idx;
  T leaf;
  bool have_leaf = false;
  if constexpr (sizeof(T) == 4) {
    if (use_top) {
      // ScanLowerBound (csrc/segment_tree.h:249-264): 8 steps from shared memory, then 4 levels per round trip by one
      // lane, or up to 7 by a cooperating group of lanes
      const T rootv = s_top[1];
      const bool over = mass > rootv;  // -> size, no descent (segment_tree.h:250-252)
      int64_t node = 1;
      T cur = over ? (T)0 : mass;      // (cooperating lanes walk a harmless all-left path for such a sample)
#pragma unroll
      for (int k = 0; k < kTopLevels - 1; ++k) {
        node <<= 1;
        descend_step(cur, node, s_top[node]);
      }
      int lev = kTopLevels - 1;
      float landed = 0.f;
      if (coop) {
        bool hl = false;
        node = descend_group_all(sum, node, depth - lev, G, cur, landed, hl);
        have_leaf = hl && !over;
      } else if (!over) {
        for (; lev + 4 <= depth; lev += 4) node = descend4_f32(sum, node, cur, landed);
        have_leaf = (lev == depth) && (depth > kTopLevels - 1);
        node = descend_plain<T>(sum, node, depth - lev, cur);
      }
      idx = over ? size : (node ^ capacity);
      leaf = landed;
    } else {
      idx = scan_lower_bound_one<T>(sum, size, capacity, depth, mass, __ldg(sum + 1), speculative != 0);
    }
  } else {
    idx = scan_lower_bound_one<T>(sum, size, capacity, depth, mass, __ldg(sum + 1), speculative != 0);
  }
  if (coop && (i >= B || (threadIdx.x & (G - 1)) != 0)) return;  // one lane per sample goes on
  if (idx > len - 1) {  // samplers.py:933
    idx = len - 1;
    have_leaf = false;
  }
  if (dbg) dbg[10] = (long long)clock64() + (idx & 0);
  if (!have_leaf) leaf = __ldg(sum + (idx | capacity));
  if (cpu_semantics) {
    // samplers.py:935-943 (CPU trees only): walk left past zero-priority leaves
    while (leaf == (T)0) {
      idx -= 1;
      if (idx < 0) {
        if (status) atomicOr(status, RLB_STATUS_BACKOFF_FAIL);
        idx = 0;
        break;
      }
      leaf = __ldg(sum + (idx | capacity));
    }
  }
  index_out[i] = idx;
  if (leaf_out) leaf_out[i] = leaf;
  // samplers.py:953  torch.pow(weight / p_min, -beta).  On the reference's CPU branch that is a true division.  On its
  // CUDA branch p_min is a PYTHON FLOAT (pybind resolves query(0-d tensor, 0-d tensor) to the (int64, int64) overload
  // through __index__, csrc/cuda_segment_tree.h:268-270), and ATen's CUDA div by a CPU scalar multiplies by the
  // reciprocal instead (BinaryDivTrueKernel.cu) -- up to one ulp away.  `cpu_semantics` selects which one is reproduced.
  const T ratio = cpu_semantics ? leaf / p_min : mul_rn(leaf, (T)1 / p_min);
  const float w = (float)pow_like_torch(ratio, neg_beta);
  weight_out[i] = w;
  if (dbg) dbg[11] = (long long)clock64() + (w > 1e30f ? 1 : 0);
}

// ------------------------------------------------------------------------------------------------
// update: batches of <= 1024 in ONE launch
// ------------------------------------------------------------------------------------------------
// Semantics: input order, the LAST duplicate wins (csrc/segment_tree.h:222-226 / cuda_segment_tree.cu:32-37), then every
// touched ancestor = op(children).  fp addition commutes and min is exact, so once the winning leaf values are fixed
// the final heap does not depend on the order in which the reference's serial loop visits the items.
//
// The tree is cut at level `bot` above the leaves (bot = depth - kDenseLevels, W = 2^kDenseLevels nodes there):
//
//   * BELOW the cut -- sorted-merge climb.  The batch is sorted by (leaf, reversed position): the head of each run of
//     equal leaves is its last writer.  Head j's root path coincides with its left neighbour's from level
//     L_j = bitlength(leaf_j ^ leaf_{j-1}) upwards, so j carries its own nodes for levels < L_j and then hands its
//     value to the group on its left: the leader of that group (the leftmost item sharing the prefix leaf_j >> L_j, a
//     lower bound in the sorted keys) consumes, at iteration L_j - 1, the slot [L_j - 1][pos_leader] of the sibling
//     tile, which until then holds the OLD value of exactly the node j carries -- j simply overwrites it, and the
//     leader's loop body is the same whether its sibling was touched or not.  Every other sibling is untouched by the
//     batch: its OLD global value is what the reference's serial update would read.  A CTA barrier is only needed
//     after a level at which some item hands over (a bit mask of the merge levels says which; for 256 random items in
//     a 2^20 tree that is a handful of the levels).
//   * ABOVE the cut -- dense.  The W cut-level nodes of both trees are pulled into shared memory (coalesced, async),
//     the items that reach the cut overwrite their node, and the W - 1 nodes above are recomputed pairwise, exactly
//     node = op(node 2k, node 2k+1) as the reference computes them (untouched nodes come out bit-identical because the
//     invariant holds everywhere), and stored back coalesced.  Merges above the cut need no bookkeeping at all.
//
// The sibling reads below the cut are 2 x bot scattered 4-byte loads per item, and a single SM's L1 moves only about
// one sector wavefront every ~2 cycles, which made them the whole cost of earlier single-CTA versions.  So the launch
// is ONE THREAD-BLOCK CLUSTER of 8 CTAs: CTA 0 (the leader) sorts and climbs; the seven helpers read every (item,
// level) sibling of both trees and store it straight into the LEADER's shared-memory tile through distributed shared
// memory (st.shared::cluster), then release a cluster barrier the leader acquires just before it needs the tile --
// by then it has loaded its keys and sorted them, so the fetch costs it nothing.  No scratch in global memory, no
// atomic ticket, no second round trip (earlier versions: a grid of CTAs wrote a scratch tile, the last one to take a
// ticket re-read it).  Finished node values are streamed to the global trees by WRITER warps (the upper half of the
// leader) once the climb is done, while the compute warps recompute the dense top.
// FUSED (fp32): `value` holds RAW priorities; the leaf is (p + eps) ** alpha (samplers.py:1076, torch.pow semantics)
// and the maximum raw priority of the valid items is folded into *max_out.
#ifndef RLB_UPDATE_DENSE_LEVELS
#define RLB_UPDATE_DENSE_LEVELS 10
#endif
constexpr int kDenseLevels = RLB_UPDATE_DENSE_LEVELS;

__device__ __forceinline__ void atomic_max_float(float *addr, float v) {
  int *ia = reinterpret_cast<int *>(addr);
  int old = *ia;
  while (__int_as_float(old) < v) {
    const int assumed = old;
    old = atomicCAS(ia, assumed, __float_as_int(v));
    if (old == assumed) break;
  }
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

constexpr size_t kUpdWorkspaceHead = size_t(1) << 20;  // (reserved head of the workspace) stamps of the general path start here

__host__ __device__ inline int upd_bot_levels(int depth) { return depth > kDenseLevels ? depth - kDenseLevels : 0; }

constexpr int kUpdCluster = 8;  // CTAs per update launch: the leader + seven helpers (portable cluster size)
#ifndef RLB_UPDATE_HOIST
#define RLB_UPDATE_HOIST 2
#endif
#ifndef RLB_UPDATE_SPLIT_KEYS
#define RLB_UPDATE_SPLIT_KEYS 0
#endif
#ifndef RLB_UPDATE_HELPER_LOADS
#define RLB_UPDATE_HELPER_LOADS 4
#endif
constexpr int kUpdHoist = RLB_UPDATE_HOIST;   // sibling values of this many levels are held in registers during the climb
constexpr int kUpdMaxItemsPerRound = 1024;  // one item per thread
constexpr int kUpdMaxRounds = 8;            // batches up to 8192 items in one launch
constexpr int kUpdMaxItems = kUpdMaxItemsPerRound * kUpdMaxRounds;
// Batches above one round are SPLIT OVER SEVERAL CLUSTERS by leaf range (the items of cluster g are the leaves in
// [g, g + 1) * capacity / G): below the cut level the clusters touch disjoint nodes, so each runs the single-cluster
// algorithm on its own compacted share; each stores its slice of the cut level and the last one to finish (ticket)
// recomputes the dense top.  Scratch in the reserved head of the update workspace:
constexpr int kUpdMaxGroups = 16;            // 16 clusters x 8 CTAs = 128 of the 148 SMs
constexpr int kUpdGroupTarget = 320;         // split until a cluster's expected share is at most this
constexpr size_t kUpdTicketOff = 0;          // u32 ticket (returns to 0 at the end of every launch)
constexpr size_t kUpdCountOff = 256;         // i32[kUpdMaxGroups] items per cluster
constexpr size_t kUpdIlistOff = 4096;        // i32[kUpdMaxGroups][kUpdMaxItems] shard-local leaf indices, input order
constexpr size_t kUpdPlistOff = kUpdIlistOff + sizeof(int32_t) * kUpdMaxGroups * kUpdMaxItems;  // u16[...] input positions
static_assert(kUpdPlistOff + sizeof(uint16_t) * kUpdMaxGroups * kUpdMaxItems <= kUpdWorkspaceHead, "workspace head");


template <typename T>
__host__ __device__ inline size_t upd_smem_bytes(int np, int depth) {
  // xbuf u64[2*np] | sleaf u32[np] | spos u32[np] | Lpos i32[np] | sraw T[np] | sib T[2][bot][np] | cut heaps T[2][2W]
  const int bot = upd_bot_levels(depth);
  const size_t W = size_t(1) << (depth - bot);
  return (size_t)np * (16 + 4 + 4 + 4) + (size_t)np * sizeof(T) * (1 + 2 * (size_t)bot) + 4 * W * sizeof(T);
}

#define RLB_TICK(k)                                                   \
  do {                                                                \
    if (dbg && threadIdx.x == 0) dbg[(k)] = (long long)clock64();     \
  } while (0)

__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int lane_mask) {
  const unsigned lo = __shfl_xor_sync(0xffffffffu, (unsigned)v, lane_mask);
  const unsigned hi = __shfl_xor_sync(0xffffffffu, (unsigned)(v >> 32), lane_mask);
  return ((unsigned long long)hi << 32) | lo;
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive_release() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrival that publishes nothing: no release fence (a release has to wait for the thread's outstanding global
// stores to be acknowledged -- a microsecond the phases that carry no data should not pay)
__device__ __forceinline__ void cluster_arrive_relaxed() {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(const void *smem_ptr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(smem_ptr)), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_cluster(uint32_t addr, double v) {
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}

__device__ __forceinline__ uint32_t ld_cluster_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float ld_cluster(uint32_t addr, float) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ double ld_cluster(uint32_t addr, double) {
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}

// Sort NP unique keys, one per thread: bucket them by `bucket` (NP buckets over the leaf range + bucket NP for the
// entries that are not items: one key per bucket on average for the random leaves prioritized sampling produces), lay
// the buckets out with a prefix sum, and order each bucket by counting -- a key's rank is its bucket's base plus the
// number of smaller keys in the bucket.  Keys are unique, so the ranks are a permutation.  (A count-everything rank
// costs ~3k cycles of shared-memory bandwidth for 256 keys, a bitonic network ~11k for 1024; this is a handful of
// barriers.  Crowded buckets -- duplicates, clustered leaves -- only make the last loop longer.)
template <typename K>
__device__ __forceinline__ K bucket_rank_sort(K key, uint32_t bucket, K *bucketed, K *sorted, uint32_t *cnt,
                                              uint32_t *s_wsum, int NP, int tid, int lane) {
  cnt[tid] = 0u;
  if (tid == 0) cnt[NP] = 0u;
  __syncthreads();
  const uint32_t slot = atomicAdd(&cnt[bucket], 1u);
  __syncthreads();
  // exclusive prefix sum of the NP + 1 counts (the non-item bucket comes last: its base is the number of items)
  const uint32_t c = cnt[tid];
  uint32_t incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_wsum[tid >> 5] = incl;
  __syncthreads();
  uint32_t woff = 0;
  for (int w = 0; w < (tid >> 5); ++w) woff += s_wsum[w];
  const uint32_t excl = woff + incl - c;  // base of bucket `tid`
  __syncthreads();                        // (everybody has read cnt[] and s_wsum[])
  cnt[tid] = excl;
  if (tid == NP - 1) cnt[NP] = excl + c;  // base of the non-item bucket
  __syncthreads();
  const uint32_t base = cnt[bucket];
  bucketed[base + slot] = key;
  __syncthreads();
  const uint32_t bend = (bucket == (uint32_t)NP) ? (uint32_t)NP : cnt[bucket + 1];
  uint32_t rank = base;
  for (uint32_t j = base; j < bend; ++j) rank += (bucketed[j] < key);
  sorted[rank] = key;
  __syncthreads();
  return sorted[tid];
}

// This kernel runs once per launch on a cold instruction cache: its code size is part of its latency.  Register-hoisting
// the sibling values of 16 levels (kUpdHoist = 16, two fully unrolled loops) made it 0.9 us SLOWER than re-reading
// shared memory level by level (profiles/README.md: 9.1 us -> 8.2 us with kUpdHoist = 2).
// KEYS: 0 -- the key width of the sort is decided at run time (default), 1 / 2 -- one instantiation per width
// (RLB_UPDATE_SPLIT_KEYS=1: measured, no gain).
template <typename T, bool FUSED, int KEYS>
__global__ void __launch_bounds__(1024) tree_update_cta_kernel(T *sum, T *mn, int64_t capacity, int depth,
                                                               const int64_t *__restrict__ index_all,
                                                               const T *__restrict__ value_all, int n_all, int scalar,
                                                               float alpha, float eps, float *max_out,
                                                               long long *dbg, int64_t index_base,
                                                               int64_t index_limit, int groups,
                                                               unsigned char *__restrict__ mc_scratch) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int NP = (int)blockDim.x;          // a power of two: one item per thread and round
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int bot = upd_bot_levels(depth);   // levels climbed item by item; the rest is dense
  const int W = 1 << (depth - bot);        // nodes at the cut level
  const uint32_t crank = cluster_ctarank(), csize = cluster_nctarank();
  __shared__ unsigned s_lmask;             // bit L set: some item hands over at level L (1 <= L <= bot)
  __shared__ uint32_t s_wsum[32];
  unsigned long long *xbuf = reinterpret_cast<unsigned long long *>(smem_raw);  // [2][NP] sort buffers
  uint32_t *sleaf = reinterpret_cast<uint32_t *>(xbuf + 2 * NP);  // sorted leaf index (0xffffffff: not an item)
  uint32_t *spos = sleaf + NP;                                    // its input position
  int *Lpos = reinterpret_cast<int *>(spos + NP);                 // by INPUT POSITION: merge level of the item if it is
                                                                  // the last writer of its leaf, else 0
  T *sraw = reinterpret_cast<T *>(Lpos + NP);                     // leaf values by input position
  T *sib = sraw + NP;  // [2][bot][NP] by input position: sibling values in, finished ancestors out (staging tile)
  T *cut_s = sib + 2 * (size_t)bot * NP;  // heap layout over the top of the tree: node k at [k], 1 <= k < 2W
  T *cut_m = cut_s + 2 * (size_t)W;

  // Batches above NP items are applied in ROUNDS of NP consecutive items, in input order, inside this one launch (a
  // later round overwrites an earlier one: "the last duplicate wins" holds across rounds); the dense top is recomputed
  // once, after the last round.  Cluster barrier phases (every thread of every CTA arrives on each, in this order):
  //   1          "all CTAs are running" (remote shared-memory accesses are legal afterwards)
  //   per round: 2  helpers -> leader: the sibling tile is complete
  //              3  leader -> helpers: merge levels, leaf values and the staging tile are final -- scatter them
  //              4  helpers -> everybody: the round's nodes are in the trees, nobody reads the leader's tile any more
  //
  // groups > 1 (see kUpdMaxGroups): phase 1 additionally publishes the leader's compaction of this cluster's share --
  // local leaf indices + input positions, in input order -- and `n_all` becomes the share's size.
  __shared__ int s_flag;
  const uint32_t gcl = groups > 1 ? cluster_id_x() : 0u;
  const int32_t *ilist = nullptr;   // groups > 1: my cluster's items (already shard-local and in range)
  const uint16_t *plist = nullptr;  //             and their positions in the caller's arrays
  if (groups > 1) {
    int32_t *il = reinterpret_cast<int32_t *>(mc_scratch + kUpdIlistOff) + (size_t)gcl * kUpdMaxItems;
    uint16_t *pl = reinterpret_cast<uint16_t *>(mc_scratch + kUpdPlistOff) + (size_t)gcl * kUpdMaxItems;
    int *count = reinterpret_cast<int *>(mc_scratch + kUpdCountOff) + gcl;
    if (crank == 0) {
      const int64_t span = capacity / groups, lo = span * gcl;
      int base = 0;
      for (int c0 = 0; c0 < n_all; c0 += NP) {
        const int i = c0 + tid;
        int64_t ix = -1;
        if (i < n_all) ix = __ldg(index_all + i) - index_base;
        const bool in = ix >= 0 && ix < index_limit && ix >= lo && ix < lo + span;
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        if (lane == 0) s_wsum[tid >> 5] = (uint32_t)__popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
        for (int w = 0; w < (NP >> 5); ++w) {
          const int c = (int)s_wsum[w];
          if (w < (tid >> 5)) woff += c;
          tot += c;
        }
        if (in) {
          const int at = base + woff + __popc(bal & ((1u << lane) - 1u));
          il[at] = (int32_t)ix;
          pl[at] = (uint16_t)i;
        }
        base += tot;
        __syncthreads();
      }
      if (tid == 0) *count = base;
      n_all = base;
      cluster_arrive_release();  // phase 1: the lists are published to the helpers
    } else {
      cluster_arrive_relaxed();  // phase 1
      cluster_wait_acquire();
      n_all = __ldcg(count);
    }
    ilist = il;
    plist = pl;
    index_base = 0;
    index_limit = capacity;
  } else {
    cluster_arrive_relaxed();  // phase 1
  }
  const int rounds = (n_all + NP - 1) / NP;
  // shard-local leaf index of item i of round rnd, or -1: not an item of this shard
  auto item_ix = [&](int rnd, int i) -> int64_t {
    if (ilist) return (int64_t)__ldcg(ilist + (size_t)rnd * NP + i);
    const int64_t ix = __ldg(index_all + (size_t)rnd * NP + i) - index_base;
    return (ix >= 0 && ix < index_limit) ? ix : -1;
  };
  if (dbg && tid == 0 && crank <= 1) dbg[32 + 8 * crank] = (long long)globaltimer_ns();

  if (crank != 0) {
    for (int rnd = 0; rnd < rounds; ++rnd) {
      const int n = min(NP, n_all - rnd * NP);
      // ---- helpers, part 1: sib[t][l][i] (in the LEADER's shared memory) = tree_t[((capacity + index[i]) >> l) ^ 1]
      // for the levels below the cut.  The loads are issued before the cluster is known to be up, the stores after.
      constexpr int kMaxPer = RLB_UPDATE_HELPER_LOADS;
      const uint32_t per_tree = (uint32_t)bot * (uint32_t)NP, total = 2u * per_tree;
      const uint32_t nthreads = (csize - 1u) * (uint32_t)NP, g = (crank - 1u) * (uint32_t)NP + tid;
      const int64_t ix_self = (tid < n) ? item_ix(rnd, tid) : -1;  // (part 2: the item at position tid)
      T val[kMaxPer];
      uint32_t el[kMaxPer];
#pragma unroll
      for (int k = 0; k < kMaxPer; ++k) {
        const uint32_t e = g + (uint32_t)k * nthreads;
        el[k] = e;
        val[k] = (T)0;
        if (e < total) {
          const uint32_t t = e >= per_tree;
          const uint32_t rem = e - t * per_tree;
          const uint32_t l = rem / (uint32_t)NP;
          const uint32_t i = rem - l * (uint32_t)NP;
          const T *tree = t ? mn : sum;
          if (tree && i < (uint32_t)n) {
            const int64_t ix = item_ix(rnd, (int)i);
            if (ix >= 0) val[k] = ld_cg(tree + (((capacity + ix) >> l) ^ 1));
          }
        }
      }
      if (rnd == 0 && groups <= 1) cluster_wait_acquire();  // phase 1
#pragma unroll
      for (int k = 0; k < kMaxPer; ++k)
        if (el[k] < total) st_cluster(map_to_cta(sib + el[k], 0), val[k]);
      for (uint32_t e = g + kMaxPer * nthreads; e < total; e += nthreads) {  // (deep trees: more than 4 per thread)
        const uint32_t t = e >= per_tree;
        const uint32_t rem = e - t * per_tree;
        const uint32_t l = rem / (uint32_t)NP;
        const uint32_t i = rem - l * (uint32_t)NP;
        const T *tree = t ? mn : sum;
        T v = (T)0;
        if (tree && i < (uint32_t)n) {
          const int64_t ix = item_ix(rnd, (int)i);
          if (ix >= 0) v = ld_cg(tree + (((capacity + ix) >> l) ^ 1));
        }
        st_cluster(map_to_cta(sib + e, 0), v);
      }
      cluster_arrive_release();  // phase 2: my stores are performed before the leader goes on
      if (dbg && tid == 0 && crank == 1) dbg[32 + 8 + 1] = (long long)globaltimer_ns();
      cluster_wait_acquire();
      cluster_arrive_relaxed();  // phase 3: nothing to publish; wait for the leader's climb
      cluster_wait_acquire();
      if (dbg && tid == 0 && crank == 1) dbg[32 + 8 + 2] = (long long)globaltimer_ns();
      // ---- helpers, part 2: scatter my rows.  Row r (0, 1: the leaves of the sum / min tree; 2 + t * bot + l: level
      // l + 1 of tree t) holds one value per INPUT POSITION p; it is a finished node iff the item at p is the last
      // writer of its leaf and carried its path beyond level l (Lpos[p] > l + 1).  Rows are read from the leader's
      // shared memory in position order -- coalesced 128-byte DSMEM reads -- and the node id comes from the index this
      // helper read itself.
      {
        const int units = (int)csize - 1, unit = (int)crank - 1, nrows = 2 * bot + 2;
        const int Lp = (int)ld_cluster_u32(map_to_cta(Lpos + tid, 0));
        if (Lp > 0 && ix_self >= 0) {
          const uint32_t leafnode = (uint32_t)(capacity + ix_self);
          for (int r = unit; r < nrows; r += units) {
            if (r < 2) {
              T *tree = r ? mn : sum;
              if (tree) tree[leafnode] = ld_cluster(map_to_cta(sraw + tid, 0), T());
            } else {
              const int t = (r - 2) >= bot, l = (r - 2) - t * bot;
              T *tree = t ? mn : sum;
              if (tree && Lp > l + 1)
                tree[leafnode >> (l + 1)] = ld_cluster(map_to_cta(sib + ((size_t)(r - 2)) * NP + tid, 0), T());
            }
          }
        }
      }
      if (dbg && tid == 0 && crank == 1) dbg[32 + 8 + 3] = (long long)globaltimer_ns();
      // phase 4.  With another round to come the nodes just stored must be visible to the next round's sibling reads
      // (release); after the last round the kernel boundary does that
      if (rnd + 1 < rounds) {
        cluster_arrive_release();
        cluster_wait_acquire();
      } else {
        cluster_arrive_relaxed();
      }
    }
    return;
  }

  // ---- leader
  if (dbg && tid == 0) dbg[0] = (long long)clock64();
  {
    // coalesced 16-byte async copies of the cut level of both trees (land while the first round sorts)
    constexpr int kPer16 = 16 / (int)sizeof(T);
    if (W >= kPer16) {
      for (int k = tid * kPer16; k < W; k += NP * kPer16) {
        if (sum) cp_async16(cut_s + W + k, sum + W + k);
        if (mn) cp_async16(cut_m + W + k, mn + W + k);
      }
    } else {
      for (int k = tid; k < W; k += NP) {
        if (sum) cut_s[W + k] = ld_cg(sum + W + k);
        if (mn) cut_m[W + k] = ld_cg(mn + W + k);
      }
    }
  }
  const int pos_bits = 31 - __clz(NP);  // log2(NP)
  const bool key32 = KEYS == 1 ? true : (KEYS == 2 ? false : (depth + 1 + pos_bits) <= 32);
  const uint32_t none32 = 0xffffffffu >> pos_bits;  // all-ones leaf field of a 32-bit key
  const int bshift = depth > pos_bits ? depth - pos_bits : 0;   // leaf index -> one of NP buckets

  for (int rnd = 0; rnd < rounds; ++rnd) {
    const int n = min(NP, n_all - rnd * NP);
    const bool last_round = rnd + 1 == rounds;
    if (tid == 0) s_lmask = 0u;
    // my own item
    bool valid = false;
    int64_t my_ix = -1;
    T raw = (T)0;
    if (tid < n) {
      // index_base maps GLOBAL indices of a sharded buffer onto this shard; entries that fall outside
      // [0, index_limit) are skipped, like the negative "do not write" markers of samplers.py:1040-1052
      my_ix = item_ix(rnd, tid);
      valid = my_ix >= 0;
      const size_t at = plist ? (size_t)__ldcg(plist + (size_t)rnd * NP + tid) : (size_t)rnd * NP + tid;
      raw = scalar ? __ldg(value_all) : __ldg(value_all + at);
    }

    // ---- 1. keys: (leaf index, reversed input position) -- ascending order puts the last writer first; entries that
    // are not items get an all-ones leaf field (still unique through the position, so ranks are a permutation).  When
    // both fields fit in 32 bits (trees up to 2^21 slots with 1024-item batches) the sort runs on 32-bit keys.
    const uint32_t rpos = (uint32_t)(NP - 1 - tid);
    const uint32_t lf0 = valid ? (uint32_t)my_ix : (key32 ? none32 : 0xffffffffu);
    const uint32_t bucket = valid ? min(lf0 >> bshift, (uint32_t)NP - 1u) : (uint32_t)NP;
    if constexpr (FUSED) {
      if (max_out) {
        float p = valid ? (float)raw : -INFINITY;
        for (int o = 16; o > 0; o >>= 1) p = fmaxf(p, __shfl_xor_sync(0xffffffffu, p, o));
        if (lane == 0 && p > -INFINITY) red_max_float(max_out, p);
      }
      raw = (T)pow_like_torch(add_rn((float)raw, eps), alpha);  // the leaf value, while the loads above are in flight
    }
    sraw[tid] = raw;
    RLB_TICK(2);

    // ---- 2. sort (bucket_rank_sort above; the counters borrow the sleaf / spos arrays, written only afterwards)
    unsigned long long key;
    if (key32) {
      uint32_t *xb32 = reinterpret_cast<uint32_t *>(xbuf);
      key = bucket_rank_sort<uint32_t>((lf0 << pos_bits) | rpos, bucket, xb32, xb32 + NP, sleaf, s_wsum, NP, tid, lane);
    } else {
      key = bucket_rank_sort<unsigned long long>(((unsigned long long)lf0 << 32) | rpos, bucket, xbuf, xbuf + NP, sleaf,
                                                 s_wsum, NP, tid, lane);
    }
    RLB_TICK(3);

    // ---- 3. heads: the first entry of each run of equal leaves is its last writer.  No compaction: everything below
    // works on the sorted array as it is (entries that are not heads are simply dead).
    const uint32_t leaf_field = key32 ? ((uint32_t)key >> pos_bits) : (uint32_t)(key >> 32);
    const bool key_valid = key32 ? (leaf_field != none32) : (leaf_field != 0xffffffffu);
    const uint32_t myleaf = key_valid ? leaf_field : 0xffffffffu;
    const uint32_t pos = (uint32_t)(NP - 1) - (uint32_t)(key & (unsigned long long)(NP - 1));
    __syncthreads();  // (the sort's counters lived in sleaf / spos)
    sleaf[tid] = myleaf;
    spos[tid] = pos;
    if (rnd == 0) {
      cp_async_wait_all();
      cluster_wait_acquire();   // phase 1
    }
    // the sibling tile: every helper has stored its share into this CTA's shared memory once phase 2 completes
    cluster_arrive_relaxed();   // phase 2
    cluster_wait_acquire();
    __syncthreads();
    if (csize == 1 && bot > 0) {
      // (launched without helpers: fetch the siblings here -- correct, just slow)
      const uint32_t per_tree = (uint32_t)bot * (uint32_t)NP, total = 2u * per_tree;
      for (uint32_t e = tid; e < total; e += NP) {
        const uint32_t t = e >= per_tree;
        const uint32_t rem = e - t * per_tree;
        const uint32_t l = rem / (uint32_t)NP;
        const uint32_t i = rem - l * (uint32_t)NP;
        const T *tree = t ? mn : sum;
        T v = (T)0;
        if (tree && i < (uint32_t)n) {
          const int64_t ix = item_ix(rnd, (int)i);
          if (ix >= 0) v = ld_cg(tree + (((capacity + ix) >> l) ^ 1));
        }
        sib[e] = v;
      }
      __syncthreads();
    }
    RLB_TICK(4);

    // ---- 4. per head j: merge level L_j, and WHERE its hand-over goes (only merges below the cut hand over)
    const uint32_t left = tid > 0 ? sleaf[tid - 1] : 0u;
    bool alive = key_valid && (tid == 0 || left != myleaf);
    const uint32_t leafnode = (uint32_t)capacity + myleaf;
    int L = 0;
    T vs = (T)0, vm = (T)0;
    T *hand_s = nullptr, *hand_m = nullptr;
    if (alive) {
      const T v = sraw[pos];
      vs = v;
      vm = v;
      L = (tid == 0) ? depth + 1 : 32 - __clz(myleaf ^ left);
      if (L <= bot) {
        const uint32_t prefix = myleaf >> L;
        int lo = 0, hi = tid;  // first index in [0, tid) whose leaf has this prefix: the leader of the group on my left
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if ((sleaf[mid] >> L) < prefix) lo = mid + 1; else hi = mid;
        }
        const uint32_t lead_pos = spos[lo];
        hand_s = sib + (size_t)(L - 1) * NP + lead_pos;
        hand_m = sib + (size_t)bot * NP + (size_t)(L - 1) * NP + lead_pos;
        if (L == 1) {  // sibling leaves: hand the leaf value over before the first iteration
          *hand_s = vs;
          *hand_m = vm;
        }
      }
    }
    if (key_valid) Lpos[pos] = L;  // (0 for the losers of a duplicated leaf; positions that hold no item stay unread)
    {
      const unsigned m = __reduce_or_sync(0xffffffffu, (alive && L <= bot) ? (1u << L) : 0u);
      if (lane == 0 && m) atomicOr(&s_lmask, m);
    }
    __syncthreads();
    const unsigned lmask = s_lmask;
    RLB_TICK(5);

    // ---- 5. climb below the cut.  The sibling values of the first kUpdHoist levels are pulled into registers up
    // front (independent shared loads); a level into which somebody handed a value over is re-read after the barrier
    // that follows the hand-over.  The parent computed at level l overwrites the (consumed) sibling slot [l][pos]: the
    // tile doubles as the staging area the helpers scatter from.
    T *io_s = sib + pos;
    T *io_m = sib + (size_t)bot * NP + pos;
    T hs[kUpdHoist > 0 ? kUpdHoist : 1], hm[kUpdHoist > 0 ? kUpdHoist : 1];
#pragma unroll
    for (int l = 0; l < kUpdHoist; ++l) {
      hs[l] = (T)0;
      hm[l] = (T)0;
      if (l < bot) {
        hs[l] = io_s[(size_t)l * NP];
        hm[l] = io_m[(size_t)l * NP];
      }
    }
#pragma unroll
    for (int l = 0; l < kUpdHoist; ++l) {
      if (l < bot) {
        if (alive) {
          if (L == l + 1) {
            alive = false;  // my level-l value was handed over; the leader of the group on my left carries the parent
          } else {
            T os = hs[l], om = hm[l];
            if ((lmask >> (l + 1)) & 1u) {  // somebody may have handed its value into row l (possibly my slot)
              os = io_s[(size_t)l * NP];
              om = io_m[(size_t)l * NP];
            }
            vs = tree_op<T, false>(vs, os);          // IEEE addition commutes: operand order is immaterial
            vm = ((myleaf >> l) & 1u) ? tree_op<T, true>(om, vm) : tree_op<T, true>(vm, om);
            io_s[(size_t)l * NP] = vs;  // value of node (leaf >> (l + 1))
            io_m[(size_t)l * NP] = vm;
            if (L == l + 2 && L <= bot) {  // I merge at the next level: hand the value just computed to the leader
              *hand_s = vs;
              *hand_m = vm;
            }
          }
        }
        if ((lmask >> (l + 2)) & 1u) __syncthreads();  // somebody handed over: its leader reads it next iteration
      }
    }
    for (int l = kUpdHoist; l < bot; ++l) {  // (trees deeper than 2^(kDenseLevels + kUpdHoist))
      if (alive) {
        if (L == l + 1) {
          alive = false;
        } else {
          const T os = io_s[(size_t)l * NP], om = io_m[(size_t)l * NP];
          vs = tree_op<T, false>(vs, os);
          vm = ((myleaf >> l) & 1u) ? tree_op<T, true>(om, vm) : tree_op<T, true>(vm, om);
          io_s[(size_t)l * NP] = vs;
          io_m[(size_t)l * NP] = vm;
          if (L == l + 2 && L <= bot) {
            *hand_s = vs;
            *hand_m = vm;
          }
        }
      }
      if ((lmask >> (l + 2)) & 1u) __syncthreads();
    }
    RLB_TICK(6);

    // ---- 6. items that reached the cut overwrite their node of the dense top (kept in shared memory across rounds)
    if (alive) {
      const uint32_t node = leafnode >> bot;  // in [W, 2W)
      cut_s[node] = vs;
      cut_m[node] = vm;
    }
    if (dbg && tid == 0) dbg[12] = (long long)clock64();
    cluster_arrive_release();  // phase 3: merge levels + staging tile final -- the helpers scatter
    if (dbg && tid == 0) dbg[13] = (long long)clock64();
    if (csize == 1) {
      // without helpers (the whole tree is above the cut, or a plain launch): the leader scatters its own nodes
      __syncthreads();
      const int Lp = Lpos[tid];
      if (valid && Lp > 0) {
        const uint32_t ln = (uint32_t)(capacity + my_ix);
        if (sum) sum[ln] = sraw[tid];
        if (mn) mn[ln] = sraw[tid];
        for (int l = 0; l < bot && Lp > l + 1; ++l) {
          if (sum) sum[ln >> (l + 1)] = sib[(size_t)l * NP + tid];
          if (mn) mn[ln >> (l + 1)] = sib[((size_t)bot + l) * NP + tid];
        }
      }
    }
    if (!last_round) {
      cluster_wait_acquire();     // phase 3
      cluster_arrive_release();   // phase 4 (release: a no-helper round's stores above feed the next round's loads)
      cluster_wait_acquire();     // the round's nodes are in the trees; the tile may be refilled
      __syncthreads();
    }
  }

  if (rounds == 0) {  // (a cluster whose share of a split batch is empty never entered the loop)
    cp_async_wait_all();
    cluster_wait_acquire();  // phase 1
  }
  bool do_top = true;
  if (groups > 1) {
    // my slice of the cut level goes to the trees; the cluster that takes the last ticket reloads the whole level (the
    // other clusters' slices) and recomputes the top
    __syncthreads();
    const int slice = W / groups;
    for (int k = tid; k < slice; k += NP) {
      const int node = W + (int)gcl * slice + k;
      if (sum) sum[node] = cut_s[node];
      if (mn) mn[node] = cut_m[node];
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      unsigned *ticket = reinterpret_cast<unsigned *>(mc_scratch + kUpdTicketOff);
      const unsigned t = atomicAdd(ticket, 1u);
      s_flag = (t == (unsigned)groups - 1u);
      if (s_flag) *ticket = 0u;  // everybody has drawn: ready for the next launch / graph replay
      __threadfence();
    }
    __syncthreads();
    do_top = s_flag != 0;
    if (do_top) {
      for (int k = tid; k < W; k += NP) {
        if (sum) cut_s[W + k] = ld_cg(sum + W + k);
        if (mn) cut_m[W + k] = ld_cg(mn + W + k);
      }
    }
  }
  // ---- 7. dense top: the W - 1 nodes above the cut are recomputed, node = op(node 2k, node 2k+1).  Each thread
  // reduces the subtree over its own `per` consecutive cut nodes alone, warps continue with shuffles, warp 0 finishes.
  __syncthreads();
  if (do_top) {
    const int per = W >= NP ? W / NP : 1;      // cut nodes per thread
    const int A = W >= NP ? NP : W;            // threads that own a subtree; its root is node A + tid
    const bool act = tid < A;
    const int plev = 31 - __clz(per);          // levels inside the thread-private subtree
    T rs = (T)0, rm = (T)0;
    if (act) {
      const int r = A + tid;
      for (int lev = plev - 1; lev >= 0; --lev) {
        const int first = r << lev;
        for (int j = 0; j < (1 << lev); ++j) {
          const int node = first + j;
          if (sum) cut_s[node] = tree_op<T, false>(cut_s[node << 1], cut_s[(node << 1) | 1]);
          if (mn) cut_m[node] = tree_op<T, true>(cut_m[node << 1], cut_m[(node << 1) | 1]);
        }
      }
      if (sum) rs = cut_s[r];
      if (mn) rm = cut_m[r];
    }
    // warp level: lanes whose index is a multiple of 2^(s+1) combine with the lane 2^s to their right
    const int wlev = A >= 32 ? 5 : 31 - __clz(A);
    int node = A + tid;
    for (int sft = 0; sft < wlev; ++sft) {
      const T os = __shfl_down_sync(0xffffffffu, rs, 1 << sft);
      const T om = __shfl_down_sync(0xffffffffu, rm, 1 << sft);
      node >>= 1;
      if (act && (lane & ((2 << sft) - 1)) == 0) {
        rs = tree_op<T, false>(rs, os);
        rm = tree_op<T, true>(rm, om);
        if (sum) cut_s[node] = rs;
        if (mn) cut_m[node] = rm;
      }
    }
    if (A > 32) {  // A / 32 subtree roots left (nodes A/32 + warp), one per warp: warp 0 finishes
      __syncthreads();
      const int A2 = A >> 5;
      if (tid < 32) {
        const bool act2 = tid < A2;
        T qs = (T)0, qm = (T)0;
        if (act2) {
          if (sum) qs = cut_s[A2 + tid];
          if (mn) qm = cut_m[A2 + tid];
        }
        int node2 = A2 + tid;
        const int lev2 = 31 - __clz(A2);
        for (int sft = 0; sft < lev2; ++sft) {
          const T os = __shfl_down_sync(0xffffffffu, qs, 1 << sft);
          const T om = __shfl_down_sync(0xffffffffu, qm, 1 << sft);
          node2 >>= 1;
          if (act2 && (lane & ((2 << sft) - 1)) == 0) {
            qs = tree_op<T, false>(qs, os);
            qm = tree_op<T, true>(qm, om);
            if (sum) cut_s[node2] = qs;
            if (mn) cut_m[node2] = qm;
          }
        }
      }
    }
  }
  __syncthreads();
  if (dbg && tid == 0) dbg[14] = (long long)clock64();
  if (do_top) {
    for (int k = 1 + tid; k < W; k += NP) {
      if (sum) sum[k] = cut_s[k];
      if (mn) mn[k] = cut_m[k];
    }
  }
  RLB_TICK(7);
  if (dbg && tid == 0) dbg[32 + 1] = (long long)globaltimer_ns();
  if (rounds > 0) {
    cluster_wait_acquire();    // phase 3 of the last round (every helper arrived long ago)
    cluster_arrive_relaxed();  // phase 4
    cluster_wait_acquire();    // the helpers have finished reading this CTA's shared memory
  }
  if (dbg && tid == 0) dbg[32 + 2] = (long long)globaltimer_ns();
}

// ------------------------------------------------------------------------------------------------
// update: general path (any n)
// ------------------------------------------------------------------------------------------------
// stamp[leaf] = max over this call's items of (epoch << 32 | position): the item whose position
// survives is the last writer of that leaf.  The stamp array is persistent; `epoch` strictly increases
// from call to call so it never needs clearing.
__global__ void upd_stamp_kernel(unsigned long long *stamp, int64_t capacity, const int64_t *__restrict__ index,
                                 int64_t n, uint32_t epoch, int64_t index_base, int64_t index_limit) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t ix = index[i] - index_base;
  if (ix < 0 || ix >= index_limit) return;
  atomicMax(stamp + ix, ((unsigned long long)epoch << 32) | (unsigned long long)(uint32_t)i);
}

template <typename T>
__global__ void upd_leaf_kernel(T *sum, T *mn, const unsigned long long *__restrict__ stamp, int64_t capacity,
                                const int64_t *__restrict__ index, const T *__restrict__ value, int64_t n,
                                int scalar, int64_t index_base, int64_t index_limit) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t ix = index[i] - index_base;
  if (ix < 0 || ix >= index_limit) return;
  if ((uint32_t)(stamp[ix] & 0xffffffffull) != (uint32_t)i) return;  // a later duplicate wins
  const T v = scalar ? value[0] : value[i];
  if (sum) sum[capacity + ix] = v;
  if (mn) mn[capacity + ix] = v;
}

// Recompute the level-(k+1) ancestor of every item from its two (already final) children.  Items that
// share an ancestor write the same bits, so the race is benign.
template <typename T>
__global__ void upd_sweep_kernel(T *sum, T *mn, int64_t capacity, const int64_t *__restrict__ index, int64_t n,
                                 int shift, int64_t index_base, int64_t index_limit) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t ix = index[i] - index_base;
  if (ix < 0 || ix >= index_limit) return;
  const int64_t p = (capacity + ix) >> shift;
  if (sum) sum[p] = tree_op<T, false>(ld_cg(sum + (p << 1)), ld_cg(sum + ((p << 1) | 1)));
  if (mn) mn[p] = tree_op<T, true>(ld_cg(mn + (p << 1)), ld_cg(mn + ((p << 1) | 1)));
}

// ------------------------------------------------------------------------------------------------
// fused (priority + eps) ** alpha  (samplers.py:1076) + running max of the raw priorities
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) per_update_prep_kernel(const int64_t *__restrict__ index,
                                                              const float *__restrict__ priority, int64_t n,
                                                              int scalar, float alpha, float eps,
                                                              float *__restrict__ leaf, float *max_out,
                                                              int64_t index_base, int64_t index_limit) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  float p = -INFINITY;
  if (i < n) {
    const float raw = scalar ? priority[0] : priority[i];
    leaf[i] = pow_like_torch(add_rn(raw, eps), alpha);
    const int64_t ix = index[i] - index_base;
    if (ix >= 0 && ix < index_limit) p = raw;
  }
  if (max_out) {
    for (int o = 16; o > 0; o >>= 1) p = fmaxf(p, __shfl_xor_sync(0xffffffffu, p, o));
    __shared__ float sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = p;
    __syncthreads();
    if (threadIdx.x == 0) {
      float m = sh[0];
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, sh[w]);
      if (m > -INFINITY) atomic_max_float(max_out, m);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int ilog2_i64(int64_t v) {
  int d = 0;
  while ((int64_t(1) << d) < v) ++d;
  return d;
}

static bool is_pow2(int64_t v) { return v > 0 && (v & (v - 1)) == 0; }

template <typename T>
static int tree_fill_impl(void *tree, int64_t capacity, int is_min, cudaStream_t st) {
  const int64_t n = 2 * capacity;
  const int threads = 256;
  int64_t blocks = (n + threads - 1) / threads;
  const int64_t cap_blocks = (int64_t)sm_count() * 8;
  if (blocks > cap_blocks) blocks = cap_blocks;
  tree_fill_kernel<T><<<(unsigned)blocks, threads, 0, st>>>(static_cast<T *>(tree), n,
                                                            is_min ? Limits<T>::max() : (T)0);
  return check_launch("tree_fill_kernel");
}

template <typename T>
static int tree_rebuild_impl(void *tree_, int64_t capacity, int is_min, cudaStream_t st) {
  T *tree = static_cast<T *>(tree_);
  const int threads = 256;
  int64_t W = capacity >> 1;  // width of the deepest internal level
  if (capacity >= 2 * kRebuildTile) {  // eleven levels per launch: every tile's subtree, down to width capacity / 2048
    const unsigned tiles = (unsigned)(capacity / kRebuildTile);
    if (is_min)
      tree_rebuild_tiles_kernel<T, true><<<tiles, threads, 0, st>>>(tree, capacity);
    else
      tree_rebuild_tiles_kernel<T, false><<<tiles, threads, 0, st>>>(tree, capacity);
    int rc = check_launch("tree_rebuild_tiles_kernel");
    if (rc) return rc;
    W = capacity / kRebuildTile / 2;
  }
  for (; W >= 1024; W >>= 1) {
    int64_t blocks = (W + threads - 1) / threads;
    const int64_t cap_blocks = (int64_t)sm_count() * 16;
    if (blocks > cap_blocks) blocks = cap_blocks;
    if (is_min)
      tree_level_dense_kernel<T, true><<<(unsigned)blocks, threads, 0, st>>>(tree, W, W);
    else
      tree_level_dense_kernel<T, false><<<(unsigned)blocks, threads, 0, st>>>(tree, W, W);
    int rc = check_launch("tree_level_dense_kernel");
    if (rc) return rc;
  }
  // nodes [1024, 2048) are final here (for capacity <= 1024 the leaves are): dense top by one CTA
  const int Wtop = (int)((capacity <= 1024) ? capacity : 1024);
  if (Wtop >= 2) {
    tree_top_kernel<T><<<1, 1024, 0, st>>>(is_min ? nullptr : tree, is_min ? tree : nullptr, Wtop);
    return check_launch("tree_top_kernel");
  }
  return RLB_OK;
}

constexpr int kUpdateSmemLimit = 224 * 1024;  // dynamic shared memory the single-CTA update may use

static int getenv_int(const char *name, int fallback) {  // (A/B measurements: RLB_UPDATE_GROUPS=1 keeps one cluster)
  static int cached = -2;
  if (cached == -2) {
    const char *e = getenv(name);
    cached = e ? atoi(e) : fallback;
  }
  return cached;
}

// threads per CTA (= items per round) and clusters for a batch of n items
static void upd_shape(int64_t n, int depth, int *np_out, int *groups_out) {
  int np = 32, groups = 1;
  const int bot = upd_bot_levels(depth);
  const int64_t W = int64_t(1) << (depth - bot);
  if (n > kUpdMaxItemsPerRound && bot > 0 && getenv_int("RLB_UPDATE_GROUPS", -1) != 1) {
    groups = 2;
    while (groups < kUpdMaxGroups && groups < W && n / groups > kUpdGroupTarget) groups <<= 1;
    const int forced = getenv_int("RLB_UPDATE_GROUPS", -1);
    if (forced > 1 && forced <= kUpdMaxGroups && (forced & (forced - 1)) == 0 && forced <= W) groups = forced;
    // a cluster's share is binomial around n / groups: a quarter of head-room keeps it to one round almost always
    const int64_t share = n / groups + n / groups / 4;
    np = 256;
    while (np < share && np < kUpdMaxItemsPerRound) np <<= 1;
  } else {
    while (np < n && np < kUpdMaxItemsPerRound) np <<= 1;  // larger batches: rounds of np items inside the launch
  }
  *np_out = np;
  *groups_out = groups;
}

struct FusedPow {
  bool on = false;
  float alpha = 0.f, eps = 0.f;
  float *max_out = nullptr;
  int64_t index_base = 0;    // subtracted from every index (global -> shard-local)
  int64_t index_limit = -1;  // valid local indices are [0, index_limit); -1 = capacity
};

template <typename T, bool FUSED, int KEYS>
static int launch_update_cta_k(T *sum, T *mn, int64_t capacity, int depth, const int64_t *index, const T *value,
                               int64_t n, int scalar, const FusedPow &fp, void *workspace, cudaStream_t st);

template <typename T, bool FUSED>
static int launch_update_cta(T *sum, T *mn, int64_t capacity, int depth, const int64_t *index, const T *value,
                             int64_t n, int scalar, const FusedPow &fp, void *workspace, cudaStream_t st) {
#if RLB_UPDATE_SPLIT_KEYS
  int np = 32, groups = 1, pos_bits = 0;
  upd_shape(n, depth, &np, &groups);
  while ((1 << pos_bits) < np) ++pos_bits;
  if (depth + 1 + pos_bits <= 32)
    return launch_update_cta_k<T, FUSED, 1>(sum, mn, capacity, depth, index, value, n, scalar, fp, workspace, st);
  return launch_update_cta_k<T, FUSED, 2>(sum, mn, capacity, depth, index, value, n, scalar, fp, workspace, st);
#else
  return launch_update_cta_k<T, FUSED, 0>(sum, mn, capacity, depth, index, value, n, scalar, fp, workspace, st);
#endif
}

template <typename T, bool FUSED, int KEYS>
static int launch_update_cta_k(T *sum, T *mn, int64_t capacity, int depth, const int64_t *index, const T *value,
                               int64_t n, int scalar, const FusedPow &fp, void *workspace, cudaStream_t st) {
  int np = 32, groups = 1;
  upd_shape(n, depth, &np, &groups);
  const size_t smem = upd_smem_bytes<T>(np, depth);
  static bool attr_set_dev[64] = {};  // function attributes are per device
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool &attr_set = attr_set_dev[cur_dev & 63];
  if (!attr_set) {
    int rc = check_cuda(cudaFuncSetAttribute(tree_update_cta_kernel<T, FUSED, KEYS>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, kUpdateSmemLimit),
                        "cudaFuncSetAttribute(tree_update_cta_kernel)");
    if (rc) return rc;
    rc = check_cuda(cudaFuncSetAttribute(tree_update_cta_kernel<T, FUSED, KEYS>,
                                         cudaFuncAttributePreferredSharedMemoryCarveout,
                                         cudaSharedmemCarveoutMaxShared),
                    "cudaFuncSetAttribute(tree_update_cta_kernel, carveout)");
    if (rc) return rc;
    attr_set = true;
  }
  // one cluster: the leader + the helpers (none needed when the whole tree is above the cut)
  const int bot = upd_bot_levels(depth);
  const unsigned cluster = bot > 0 ? kUpdCluster : 1;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(cluster * (unsigned)groups);
  cfg.blockDim = dim3(np);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int rc = check_cuda(cudaLaunchKernelEx(&cfg, tree_update_cta_kernel<T, FUSED, KEYS>, sum, mn, capacity, depth, index, value,
                                         (int)n, scalar, fp.alpha, fp.eps, fp.max_out, g_debug_ticks, fp.index_base,
                                         fp.index_limit < 0 ? capacity : fp.index_limit, groups,
                                         static_cast<unsigned char *>(workspace)),
                      "tree_update_cta_kernel");
  if (rc) return rc;
  return check_launch("tree_update_cta_kernel");
}

template <typename T>
static bool update_fits_cta(int64_t n, int64_t capacity, int depth) {
  // the cluster kernel takes up to kUpdMaxItems items in ONE launch (capturable, no epoch stamp); beyond that the
  // stamp + sweep path below is the better algorithm (it is bandwidth-, not latency-bound)
  if (n > (int64_t)kUpdMaxItems || capacity > (int64_t(1) << 30)) return false;
  int np = 32, groups = 1;
  upd_shape(n, depth, &np, &groups);
  return upd_smem_bytes<T>(np, depth) <= (size_t)kUpdateSmemLimit;
}


template <typename T>
static int tree_update_impl(void *sum_, void *mn_, int64_t capacity, const int64_t *index, const void *value_,
                            int64_t n, int scalar, void *workspace, size_t workspace_bytes, uint32_t epoch,
                            cudaStream_t st, const FusedPow &fp = FusedPow()) {
  T *sum = static_cast<T *>(sum_);
  T *mn = static_cast<T *>(mn_);
  const T *value = static_cast<const T *>(value_);
  const int depth = ilog2_i64(capacity);
  RLB_REQUIRE(workspace != nullptr && workspace_bytes >= kUpdWorkspaceHead + (size_t)capacity * sizeof(unsigned long long),
              RLB_EINVAL, "rlb_tree_update: needs a zero-initialised workspace of rlb_tree_update_workspace_bytes(size)");
  if (update_fits_cta<T>(n, capacity, depth)) {
    if constexpr (sizeof(T) == 4) {
      if (fp.on)
        return launch_update_cta<T, true>(sum, mn, capacity, depth, index, value, n, scalar, fp, workspace, st);
    }
    return launch_update_cta<T, false>(sum, mn, capacity, depth, index, value, n, scalar, fp, workspace, st);
  }
  // general path
  unsigned long long *stamp =
      reinterpret_cast<unsigned long long *>(static_cast<unsigned char *>(workspace) + kUpdWorkspaceHead);
  const int threads = 256;
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  const int64_t ibase = fp.index_base, ilimit = fp.index_limit < 0 ? capacity : fp.index_limit;
  upd_stamp_kernel<<<blocks, threads, 0, st>>>(stamp, capacity, index, n, epoch, ibase, ilimit);
  int rc = check_launch("upd_stamp_kernel");
  if (rc) return rc;
  upd_leaf_kernel<T><<<blocks, threads, 0, st>>>(sum, mn, stamp, capacity, index, value, n, scalar, ibase, ilimit);
  rc = check_launch("upd_leaf_kernel");
  if (rc) return rc;
  // sweep the levels of width >= 1024 (touched ancestors only), then one CTA recomputes the dense top
  int shift = 1;
  for (int64_t W = capacity >> 1; W >= 1024; W >>= 1, ++shift) {
    upd_sweep_kernel<T><<<blocks, threads, 0, st>>>(sum, mn, capacity, index, n, shift, ibase, ilimit);
    rc = check_launch("upd_sweep_kernel");
    if (rc) return rc;
  }
  const int Wtop = (int)((capacity <= 1024) ? capacity : 1024);
  if (Wtop >= 2) {
    tree_top_kernel<T><<<1, 1024, 0, st>>>(sum, mn, Wtop);
    rc = check_launch("tree_top_kernel");
  }
  return rc;
}


// ------------------------------------------------------------------------------------------------
// contiguous-range priority write (tree_range.cuh)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kRangeThreads) tree_range_kernel(const __grid_constant__ RangeParams R) {
  range_role<T>(R, (int)blockIdx.x, (int)gridDim.x);
}

}  // namespace rlb

using namespace rlb;

extern "C" {

/* profiling aid (not part of the public header): device buffer of >= 8 int64 that the single-launch update
 * kernel fills with clock64() stamps at its phase boundaries; NULL disables. */
void rlb_debug_set_tick_buffer(void *p) { rlb::g_debug_ticks = static_cast<long long *>(p); }

int64_t rlb_tree_capacity(int64_t size) {
  int64_t c = 1;
  for (; c <= size; c <<= 1) {
  }
  return c;
}

size_t rlb_tree_update_workspace_bytes(int64_t size) {
  // 1 MB head (launch ticket + the sibling scratch tile of the <=1024 path) followed by one 64-bit
  // (epoch, position) stamp per addressable leaf slot for larger batches
  return rlb::kUpdWorkspaceHead + (size_t)rlb_tree_capacity(size) * sizeof(unsigned long long);
}

int rlb_tree_fill(void *tree, int64_t capacity, int is_min, int dtype, rlb_stream_t stream) {
  RLB_REQUIRE(tree && is_pow2(capacity), RLB_EINVAL, "rlb_tree_fill: null tree or capacity not a power of two");
  if (dtype == RLB_F32) return tree_fill_impl<float>(tree, capacity, is_min, as_stream(stream));
  if (dtype == RLB_F64) return tree_fill_impl<double>(tree, capacity, is_min, as_stream(stream));
  RLB_REQUIRE(false, RLB_EINVAL, "rlb_tree_fill: unsupported dtype %d", dtype);
}

int rlb_tree_rebuild(void *tree, int64_t capacity, int is_min, int dtype, rlb_stream_t stream) {
  RLB_REQUIRE(tree && is_pow2(capacity), RLB_EINVAL, "rlb_tree_rebuild: null tree or capacity not a power of two");
  if (dtype == RLB_F32) return tree_rebuild_impl<float>(tree, capacity, is_min, as_stream(stream));
  if (dtype == RLB_F64) return tree_rebuild_impl<double>(tree, capacity, is_min, as_stream(stream));
  RLB_REQUIRE(false, RLB_EINVAL, "rlb_tree_rebuild: unsupported dtype %d", dtype);
}

int rlb_tree_update(void *sum_tree, void *min_tree, int64_t capacity, const int64_t *index, const void *value,
                    int64_t n, int scalar, int dtype, void *workspace, size_t workspace_bytes, uint32_t epoch,
                    rlb_stream_t stream) {
  RLB_REQUIRE((sum_tree || min_tree) && is_pow2(capacity), RLB_EINVAL,
              "rlb_tree_update: no tree given or capacity not a power of two");
  RLB_REQUIRE(n >= 0, RLB_EINVAL, "rlb_tree_update: negative n");
  if (n == 0) return RLB_OK;
  RLB_REQUIRE(index && value, RLB_EINVAL, "rlb_tree_update: null index/value");
  if (dtype == RLB_F32)
    return tree_update_impl<float>(sum_tree, min_tree, capacity, index, value, n, scalar, workspace,
                                   workspace_bytes, epoch, as_stream(stream));
  if (dtype == RLB_F64)
    return tree_update_impl<double>(sum_tree, min_tree, capacity, index, value, n, scalar, workspace,
                                    workspace_bytes, epoch, as_stream(stream));
  RLB_REQUIRE(false, RLB_EINVAL, "rlb_tree_update: unsupported dtype %d", dtype);
}

int rlb_tree_query(const void *tree, int64_t size, int64_t capacity, int is_min, int dtype, const int64_t *l,
                   const int64_t *r, void *out, int64_t n, int root_fast_path, rlb_stream_t stream) {
  RLB_REQUIRE(tree && is_pow2(capacity) && n >= 0, RLB_EINVAL, "rlb_tree_query: bad arguments");
  if (n == 0) return RLB_OK;
  RLB_REQUIRE(l && r && out, RLB_EINVAL, "rlb_tree_query: null l/r/out");
  const int threads = 128;
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  cudaStream_t st = as_stream(stream);
  if (dtype == RLB_F32) {
    if (is_min)
      tree_query_kernel<float, true><<<blocks, threads, 0, st>>>((const float *)tree, size, capacity, l, r,
                                                                 (float *)out, n, root_fast_path,
                                                                 Limits<float>::max());
    else
      tree_query_kernel<float, false><<<blocks, threads, 0, st>>>((const float *)tree, size, capacity, l, r,
                                                                  (float *)out, n, root_fast_path, 0.0f);
  } else if (dtype == RLB_F64) {
    if (is_min)
      tree_query_kernel<double, true><<<blocks, threads, 0, st>>>((const double *)tree, size, capacity, l, r,
                                                                  (double *)out, n, root_fast_path,
                                                                  Limits<double>::max());
    else
      tree_query_kernel<double, false><<<blocks, threads, 0, st>>>((const double *)tree, size, capacity, l, r,
                                                                   (double *)out, n, root_fast_path, 0.0);
  } else {
    RLB_REQUIRE(false, RLB_EINVAL, "rlb_tree_query: unsupported dtype %d", dtype);
  }
  return check_launch("tree_query_kernel");
}

int rlb_tree_at(const void *tree, int64_t capacity, int dtype, const int64_t *index, void *out, int64_t n,
                rlb_stream_t stream) {
  RLB_REQUIRE(tree && is_pow2(capacity) && n >= 0, RLB_EINVAL, "rlb_tree_at: bad arguments");
  if (n == 0) return RLB_OK;
  RLB_REQUIRE(index && out, RLB_EINVAL, "rlb_tree_at: null index/out");
  const int threads = 128;
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  if (dtype == RLB_F32)
    tree_at_kernel<float><<<blocks, threads, 0, as_stream(stream)>>>((const float *)tree, capacity, index,
                                                                     (float *)out, n);
  else if (dtype == RLB_F64)
    tree_at_kernel<double><<<blocks, threads, 0, as_stream(stream)>>>((const double *)tree, capacity, index,
                                                                      (double *)out, n);
  else
    RLB_REQUIRE(false, RLB_EINVAL, "rlb_tree_at: unsupported dtype %d", dtype);
  return check_launch("tree_at_kernel");
}

int rlb_tree_scan_lower_bound(const void *sum_tree, int64_t size, int64_t capacity, int dtype, const void *value,
                              int64_t *out, int64_t n, rlb_stream_t stream) {
  RLB_REQUIRE(sum_tree && is_pow2(capacity) && n >= 0, RLB_EINVAL, "rlb_tree_scan_lower_bound: bad arguments");
  if (n == 0) return RLB_OK;
  RLB_REQUIRE(value && out, RLB_EINVAL, "rlb_tree_scan_lower_bound: null value/out");
  const int depth = ilog2_i64(capacity);
  const int threads = 128;
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  if (dtype == RLB_F32)
    tree_scan_kernel<float><<<blocks, threads, 0, as_stream(stream)>>>((const float *)sum_tree, size, capacity,
                                                                       depth, (const float *)value, out, n);
  else if (dtype == RLB_F64)
    tree_scan_kernel<double><<<blocks, threads, 0, as_stream(stream)>>>((const double *)sum_tree, size, capacity,
                                                                        depth, (const double *)value, out, n);
  else
    RLB_REQUIRE(false, RLB_EINVAL, "rlb_tree_scan_lower_bound: unsupported dtype %d", dtype);
  return check_launch("tree_scan_kernel");
}

int rlb_per_sample(const void *sum_tree, const void *min_tree, int64_t size, int64_t capacity, int dtype,
                   int64_t len, const void *u, int64_t B, double beta, int cpu_semantics, int64_t *index_out,
                   float *weight_out, void *leaf_out, void *psum_pmin_out, int32_t *status, rlb_stream_t stream) {
  RLB_REQUIRE(sum_tree && min_tree && is_pow2(capacity), RLB_EINVAL, "rlb_per_sample: null tree or bad capacity");
  RLB_REQUIRE(size > 0 && size < capacity, RLB_EINVAL, "rlb_per_sample: size must be in (0, capacity)");
  RLB_REQUIRE(len > 0 && len <= size, RLB_EINVAL,
              "rlb_per_sample: len=%lld outside (0, size=%lld] (Cannot sample from an empty storage.)",
              (long long)len, (long long)size);
  RLB_REQUIRE(B >= 0, RLB_EINVAL, "rlb_per_sample: negative batch");
  if (B == 0) return RLB_OK;
  RLB_REQUIRE(u && index_out && weight_out, RLB_EINVAL, "rlb_per_sample: null u/index_out/weight_out");
  const int depth = ilog2_i64(capacity);
  const int sms = sm_count();
  const int speculative = B <= (int64_t)sms * 64;  // latency-bound regime: 4 levels per round trip
  int threads = 128, spc = 128;
  if (speculative) {  // spread: ~2 single-warp CTAs per SM, each with as few samples as that allows
    threads = 32;
    spc = (int)((B + 2 * (int64_t)sms - 1) / (2 * (int64_t)sms));
    if (spc < 1) spc = 1;
    if (spc > 32) spc = 32;
  }
  const unsigned blocks = (unsigned)((B + spc - 1) / spc);
  int rc;
  if (dtype == RLB_F32)
    rc = check_cuda(launch_pdl(per_sample_kernel<float>, dim3(blocks), dim3(threads), 0, as_stream(stream),
                               (const float *)sum_tree, (const float *)min_tree, size, capacity, depth, len,
                               (const float *)u, B, (float)(-beta), cpu_semantics, speculative, spc, index_out,
                               weight_out, (float *)leaf_out, (float *)psum_pmin_out, status, g_debug_ticks),
                    "per_sample_kernel");
  else if (dtype == RLB_F64)
    rc = check_cuda(launch_pdl(per_sample_kernel<double>, dim3(blocks), dim3(threads), 0, as_stream(stream),
                               (const double *)sum_tree, (const double *)min_tree, size, capacity, depth, len,
                               (const double *)u, B, (double)(-beta), cpu_semantics, speculative, spc, index_out,
                               weight_out, (double *)leaf_out, (double *)psum_pmin_out, status, g_debug_ticks),
                    "per_sample_kernel");
  else
    RLB_REQUIRE(false, RLB_EINVAL, "rlb_per_sample: unsupported dtype %d", dtype);
  if (rc) return rc;
  return check_launch("per_sample_kernel");
}

int rlb_per_update(void *sum_tree, void *min_tree, int64_t capacity, const int64_t *index, const float *priority,
                   int64_t n, int scalar, double alpha, double eps, float *leaf_scratch, float *max_priority_out,
                   void *workspace, size_t workspace_bytes, uint32_t epoch, int64_t index_base, int64_t index_limit,
                   rlb_stream_t stream) {
  RLB_REQUIRE((sum_tree || min_tree) && is_pow2(capacity), RLB_EINVAL, "rlb_per_update: bad tree/capacity");
  RLB_REQUIRE(n >= 0, RLB_EINVAL, "rlb_per_update: negative n");
  if (n == 0) return RLB_OK;
  RLB_REQUIRE(index && priority, RLB_EINVAL, "rlb_per_update: null index/priority");
  RLB_REQUIRE(leaf_scratch || update_fits_cta<float>(n, capacity, ilog2_i64(capacity)), RLB_EINVAL,
              "rlb_per_update: this batch size needs leaf_scratch[n]");
  cudaStream_t st = as_stream(stream);
  if (update_fits_cta<float>(n, capacity, ilog2_i64(capacity))) {
    // one launch: pow, running max, election, leaf write and the climb all happen inside the CTA kernel
    FusedPow fp;
    fp.on = true;
    fp.alpha = (float)alpha;
    fp.eps = (float)eps;
    fp.max_out = max_priority_out;
    fp.index_base = index_base;
    fp.index_limit = index_limit;
    return tree_update_impl<float>(sum_tree, min_tree, capacity, index, priority, n, scalar, workspace,
                                   workspace_bytes, epoch, st, fp);
  }
  const int threads = 256;
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  per_update_prep_kernel<<<blocks, threads, 0, st>>>(index, priority, n, scalar, (float)alpha, (float)eps,
                                                     leaf_scratch, max_priority_out, index_base,
                                                     index_limit < 0 ? capacity : index_limit);
  int rc = check_launch("per_update_prep_kernel");
  if (rc) return rc;
  FusedPow plain;  // pow already applied; only the index mapping is forwarded
  plain.index_base = index_base;
  plain.index_limit = index_limit;
  return tree_update_impl<float>(sum_tree, min_tree, capacity, index, leaf_scratch, n, /*scalar=*/0, workspace,
                                 workspace_bytes, epoch, st, plain);
}

int rlb_tree_update_range(void *sum_tree, void *min_tree, int64_t capacity, int dtype, int64_t start, int64_t n,
                          int64_t modulo, int mode, const void *value, double alpha, double eps, double first_default,
                          int has_max, float *max_priority, uint32_t *ticket, rlb_stream_t stream) {
  if (n == 0) return RLB_OK;
  RangeParams R;
  int rc = range_params(R, "rlb_tree_update_range", sum_tree, min_tree, capacity, dtype, start, n, modulo, mode, value,
                        alpha, eps, first_default, has_max, max_priority, ticket);
  if (rc) return rc;
  const int ctas = range_ctas(n);
  if (dtype == RLB_F32)
    tree_range_kernel<float><<<ctas, kRangeThreads, 0, as_stream(stream)>>>(R);
  else
    tree_range_kernel<double><<<ctas, kRangeThreads, 0, as_stream(stream)>>>(R);
  return check_launch("tree_range_kernel");
}

}  // extern "C"

// slice.cu -- trajectory-slice sampling for replay storages laid out as a ring of time steps (sm_100a).
//
// Replaces the index arithmetic of SliceSampler (data/replay_buffers/samplers.py:1207-2300) for 1-d storages:
//
//   rlb_traj_table   _find_start_stop_traj :1652-1706 + _end_to_start_stop :1708-1743.  The reference derives the
//                    (start, stop, length) table of all trajectories with nonzero / roll-by-mask / index arithmetic
//                    (~15 launches and a host sync) and, with strict_length, filters it by boolean indexing
//                    (:1993-2010, two more syncs) on EVERY sample unless cache_values is set.  Here: one pass that
//                    marks trajectory ends per 16384-slot tile and scans the tile counts in the last CTA to arrive, one
//                    pass that emits the table (each end knows the previous end from an in-tile max-scan plus the tile
//                    prefix, so start / length need no second table pass) and counts the trajectories that are long
//                    enough, and -- only when short ones must be dropped -- one pass that emits the compacted table.
//   rlb_slice_index  _get_index :2058-2215 (span = False): relative start = floor(u * (len - seq + 1)) in the
//                    reference's fp32 arithmetic, ring wrap, truncated markers, padded form with mask; the reference
//                    spends ~25 small launches here.
//
// Integer / byte work: results are bit-exact against the reference (tests/golden/slice_golden.npz).  Bound by the one
// byte (or eight, for trajectory ids) per slot the table pass must read: HBM streaming, no reuse.
#include "common.cuh"

namespace rlb {

constexpr int kTrajThreads = 256;
constexpr int kTrajItems = 64;  // consecutive slots per thread = four 128-bit loads of flag bytes in flight
constexpr int kTrajTile = kTrajThreads * kTrajItems;

struct TrajWorkspace {  // header of the caller's workspace; per-tile arrays and the flag bytes follow
  unsigned int ticket[2];
  int force_last;  // no end found in a full ring: the last slot closes the only trajectory (:1700-1703)
  int pad_;
  long long overall_last;
};

struct TrajParams {
  const void *signal;
  int kind;
  int at_capacity;
  int filter;
  int pad_;
  int64_t L, cursor, min_len;
  int64_t *start, *stop, *length, *counts;
  TrajWorkspace *ws;
  int *tile_cnt;         // ends per tile, then exclusive prefix
  long long *tile_last;  // last end position in the tile (-1), then last end position BEFORE the tile
  int *tile_kcnt;        // kept trajectories per tile, then exclusive prefix
  uint8_t *flags;        // RLB_TRAJ_ID: the id comparison of pass 1, one byte per slot, for the later passes
  int64_t n_tiles;
};

__device__ __forceinline__ uint32_t nonzero_bytes(uint32_t w) {  // bit k = (byte k != 0)
  return ((w & 0xffu) ? 1u : 0u) | ((w & 0xff00u) ? 2u : 0u) | ((w & 0xff0000u) ? 4u : 0u) | ((w >> 24) ? 8u : 0u);
}

__device__ __forceinline__ uint64_t nonzero_bytes16(const uint4 q) {
  return (uint64_t)(nonzero_bytes(q.x) | (nonzero_bytes(q.y) << 4) | (nonzero_bytes(q.z) << 8) |
                    (nonzero_bytes(q.w) << 12));
}

// the end-of-trajectory flags of slots [base, base + kTrajItems) from a byte array, as a bit mask
__device__ __forceinline__ uint64_t flags_from_bytes(const uint8_t *f, int64_t base, int64_t L) {
  if (base >= L) return 0ull;
  if (base + kTrajItems <= L && (reinterpret_cast<uintptr_t>(f + base) & 15u) == 0) {
    uint4 q[kTrajItems / 16];
#pragma unroll
    for (int v = 0; v < kTrajItems / 16; ++v) q[v] = __ldg(reinterpret_cast<const uint4 *>(f + base) + v);
    uint64_t m = 0;
#pragma unroll
    for (int v = 0; v < kTrajItems / 16; ++v) m |= nonzero_bytes16(q[v]) << (16 * v);
    return m;
  }
  uint64_t m = 0;
  for (int j = 0; j < kTrajItems && base + j < L; ++j) m |= (uint64_t)(f[base + j] != 0 ? 1 : 0) << j;
  return m;
}

// The tile's raw flags.  RLB_TRAJ_END, or a later pass of RLB_TRAJ_ID: straight from the byte array.  Pass 1 of
// RLB_TRAJ_ID: ids are compared with coalesced (striped) loads, staged in shared memory, re-read blocked, and saved
// as bytes so that the other passes read one byte per slot instead of eight.
template <bool FIRST_PASS>
__device__ __forceinline__ uint64_t traj_tile_flags(const TrajParams &P, uint8_t *sh_flags) {
  const int64_t tile0 = (int64_t)blockIdx.x * kTrajTile;
  const int64_t base = tile0 + (int64_t)threadIdx.x * kTrajItems;
  uint64_t m;
  if (P.kind == RLB_TRAJ_END) {
    m = flags_from_bytes(static_cast<const uint8_t *>(P.signal), base, P.L);
  } else if (!FIRST_PASS) {
    m = flags_from_bytes(P.flags, base, P.L);
  } else {
    const int64_t *id = static_cast<const int64_t *>(P.signal);
#pragma unroll 4
    for (int r = 0; r < kTrajItems; ++r) {
      const int64_t i = tile0 + (int64_t)r * kTrajThreads + threadIdx.x;
      bool e = false;
      if (i < P.L) e = (i + 1 < P.L) ? (__ldg(id + i) != __ldg(id + i + 1)) : (P.at_capacity ? id[i] != id[0] : true);
      sh_flags[r * kTrajThreads + threadIdx.x] = e ? 1 : 0;
    }
    __syncthreads();
    m = 0;
#pragma unroll
    for (int v = 0; v < kTrajItems / 16; ++v) {
      const uint4 q = *(reinterpret_cast<const uint4 *>(sh_flags + threadIdx.x * kTrajItems) + v);
      m |= nonzero_bytes16(q) << (16 * v);
      if (base < P.L) *(reinterpret_cast<uint4 *>(P.flags + base) + v) = q;  // (padded to whole tiles)
    }
  }
  return m;
}

// the reference's boundary rules on top of the raw flags (:1675-1677, :1683-1703)
__device__ __forceinline__ uint64_t traj_apply_rules(const TrajParams &P, uint64_t m, int64_t base, int force_last) {
  const int64_t last = P.L - 1;
  if ((!P.at_capacity || force_last) && last >= base && last < base + kTrajItems) m |= 1ull << (int)(last - base);
  if (P.at_capacity && P.cursor >= base && P.cursor < base + kTrajItems) m |= 1ull << (int)(P.cursor - base);
  return m;
}

// block-wide exclusive scans over one value per thread (kTrajThreads threads); `total` = reduction over the block
__device__ __forceinline__ int block_excl_sum(int v, int *total, int *sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) sh[warp] = inc;
  __syncthreads();
  int base = 0, tot = 0;
  for (int w = 0; w < kTrajThreads / 32; ++w) {
    if (w < warp) base += sh[w];
    tot += sh[w];
  }
  __syncthreads();
  *total = tot;
  return base + inc - v;
}
__device__ __forceinline__ long long block_excl_max(long long v, long long *total, long long *sh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const long long t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d && t > inc) inc = t;
  }
  long long excl = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane == 0) excl = -1;
  if (lane == 31) sh[warp] = inc;
  __syncthreads();
  long long base = -1, tot = -1;
  for (int w = 0; w < kTrajThreads / 32; ++w) {
    if (w < warp && sh[w] > base) base = sh[w];
    if (sh[w] > tot) tot = sh[w];
  }
  __syncthreads();
  *total = tot;
  return excl > base ? excl : base;
}

// the last CTA to take ticket `which` returns true (and resets the ticket for the next call)
__device__ __forceinline__ bool traj_last_cta(TrajWorkspace *ws, int which, int *sh_flag) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(&ws->ticket[which], 1u);
    *sh_flag = (t == gridDim.x - 1);
    if (*sh_flag) ws->ticket[which] = 0u;
  }
  __syncthreads();
  if (*sh_flag) __threadfence();
  return *sh_flag != 0;
}

// pass 1: ends per tile; the last CTA turns the per-tile figures into prefixes
__global__ void __launch_bounds__(kTrajThreads) traj_mark_kernel(const TrajParams P) {
  __shared__ __align__(16) uint8_t sh_flags[kTrajTile];
  __shared__ int sh_i[kTrajThreads / 32];
  __shared__ long long sh_l[kTrajThreads / 32];
  __shared__ int sh_flag;
  const int64_t base = (int64_t)blockIdx.x * kTrajTile + (int64_t)threadIdx.x * kTrajItems;
  const uint64_t m = traj_apply_rules(P, traj_tile_flags<true>(P, sh_flags), base, 0);
  const int cnt = __popcll(m);
  const long long last = m ? base + (63 - __clzll(m)) : -1;
  int tot;
  long long tl;
  block_excl_sum(cnt, &tot, sh_i);
  block_excl_max(last, &tl, sh_l);
  if (threadIdx.x == 0) {
    P.tile_cnt[blockIdx.x] = tot;
    P.tile_last[blockIdx.x] = tl;
  }
  if (!traj_last_cta(P.ws, 0, &sh_flag)) return;
  // exclusive prefix of counts and running "last end before this tile", kTrajThreads tiles at a time
  int carry = 0;
  long long carry_last = -1;
  for (int64_t t0 = 0; t0 < P.n_tiles; t0 += kTrajThreads) {
    const int64_t t = t0 + threadIdx.x;
    const int c = t < P.n_tiles ? __ldcg(P.tile_cnt + t) : 0;
    const long long l = t < P.n_tiles ? __ldcg(P.tile_last + t) : -1;
    int ctot;
    long long ltot;
    const int ce = block_excl_sum(c, &ctot, sh_i);
    const long long le = block_excl_max(l, &ltot, sh_l);
    if (t < P.n_tiles) {
      P.tile_cnt[t] = carry + ce;
      P.tile_last[t] = le > carry_last ? le : carry_last;
    }
    carry += ctot;
    if (ltot > carry_last) carry_last = ltot;
  }
  if (threadIdx.x == 0) {
    const int force = carry == 0;  // (only possible at capacity: otherwise the last slot is always an end)
    P.ws->force_last = force;
    P.ws->overall_last = force ? P.L - 1 : carry_last;
    P.counts[0] = force ? 1 : carry;
  }
}

// pass 2 (EMIT_KEPT = false): the full table (unless filtered) and the count of long-enough trajectories per tile;
// pass 3 (EMIT_KEPT = true): the table of long-enough trajectories only.
template <bool EMIT_KEPT>
__global__ void __launch_bounds__(kTrajThreads) traj_emit_kernel(const TrajParams P) {
  __shared__ int sh_i[kTrajThreads / 32];
  __shared__ long long sh_l[kTrajThreads / 32];
  __shared__ int sh_flag;
  const int force = P.ws->force_last;
  const long long overall_last = P.ws->overall_last;
  const int64_t base = (int64_t)blockIdx.x * kTrajTile + (int64_t)threadIdx.x * kTrajItems;
  const uint64_t m = traj_apply_rules(P, traj_tile_flags<false>(P, nullptr), base, force);
  const int cnt = __popcll(m);
  const long long last = m ? base + (63 - __clzll(m)) : -1;
  int tot;
  long long tl;
  const int rank = block_excl_sum(cnt, &tot, sh_i);
  long long prev = block_excl_max(last, &tl, sh_l);
  const long long before_tile = P.tile_last[blockIdx.x];
  if (before_tile > prev) prev = before_tile;
  // first sweep over this thread's ends: how many are long enough
  int kept = 0;
  {
    long long p = prev;
    for (uint64_t mm = m; mm; mm &= mm - 1) {
      const int64_t i = base + (__ffsll((long long)mm) - 1);
      const long long q = p >= 0 ? p : overall_last - P.L;  // the ring: the first trajectory starts after the last end
      int64_t s0 = q + 1;
      if (s0 < 0) s0 += P.L;
      int64_t n = i - s0 + 1;
      if (n <= 0) n += P.L;
      if (n >= P.min_len) ++kept;
      p = i;
    }
  }
  int64_t k;
  bool emit;
  if (!EMIT_KEPT) {
    k = (int64_t)P.tile_cnt[blockIdx.x] + rank;
    emit = !P.filter;
  } else {
    int ktot;
    const int krank = block_excl_sum(kept, &ktot, sh_i);
    k = (int64_t)P.tile_kcnt[blockIdx.x] + krank;
    emit = true;
  }
  if (emit) {
    long long p = prev;
    for (uint64_t mm = m; mm; mm &= mm - 1) {
      const int64_t i = base + (__ffsll((long long)mm) - 1);
      const long long q = p >= 0 ? p : overall_last - P.L;
      int64_t s0 = q + 1;
      if (s0 < 0) s0 += P.L;
      int64_t n = i - s0 + 1;
      if (n <= 0) n += P.L;
      p = i;
      if (EMIT_KEPT && n < P.min_len) continue;
      P.start[k] = s0;
      P.stop[k] = i;
      P.length[k] = n;
      ++k;
    }
  }
  if (EMIT_KEPT) return;
  int ktot;
  block_excl_sum(kept, &ktot, sh_i);
  if (threadIdx.x == 0) P.tile_kcnt[blockIdx.x] = ktot;
  if (!traj_last_cta(P.ws, 1, &sh_flag)) return;
  int carry = 0;
  for (int64_t t0 = 0; t0 < P.n_tiles; t0 += kTrajThreads) {
    const int64_t t = t0 + threadIdx.x;
    const int c = t < P.n_tiles ? __ldcg(P.tile_kcnt + t) : 0;
    int ctot;
    const int ce = block_excl_sum(c, &ctot, sh_i);
    if (t < P.n_tiles) P.tile_kcnt[t] = carry + ce;
    carry += ctot;
  }
  if (threadIdx.x == 0) P.counts[1] = carry;
}

// ---- slice expansion -----------------------------------------------------------------------------------
struct SliceParams {
  const int64_t *start, *length, *traj_draw, *out_offset;
  const float *u;
  int64_t n_traj, num_slices, seq_length, storage_length;
  int64_t span0, span1;  // SliceSampler(span=(left, right)): 0 = off, -1 = True, k > 0 = at most k steps outside
  int variable, pad_output;
  int64_t *index_out, *seq_out;
  uint8_t *truncated_out, *mask_out;
  // optional: the stored one-byte done / terminated flags of the sampled slots, so that the info the sampler returns
  // (done | truncated, terminated; samplers.py:2190-2205) needs no second gather
  const uint8_t *done_src, *term_src;
  uint8_t *done_out, *term_out;
};

// one warp per slice
__global__ void __launch_bounds__(128) slice_index_kernel(const SliceParams P) {
  const int lane = threadIdx.x & 31;
  const int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= P.num_slices) return;
  int64_t t = __ldg(P.traj_draw + s);
  if (t < 0) t = 0;
  if (t >= P.n_traj) t = P.n_traj - 1;
  const int64_t len = __ldg(P.length + t);
  int64_t seq = P.seq_length;
  if (P.variable && len < seq) seq = len;                          // lengths[traj_idx].clamp_max(seq_length), :2037
  // :2071-2097: the range the relative start is drawn from; a span lets the slice hang out of the trajectory on the
  // left (start_point < 0) and / or on the right (end_point beyond the last indexable start)
  const int64_t end_point = P.span1 == 0 ? len - seq + 1 : (P.span1 < 0 ? len + 1 : len - P.span1);
  const int64_t start_point = P.span0 == 0 ? 0 : (P.span0 < 0 ? 1 - seq : -P.span0);
  // torch.rand(fp32) * int64 tensor: the integer is converted to fp32 and the product rounded once (:2099-2102)
  const float prod = mul_rn(__ldg(P.u + s), (float)(end_point - start_point));
  int64_t rel = (int64_t)floorf(prod) + start_point;
  if (P.span0 && rel < 0) {                                        // :2104-2111: fewer elements, from the first step
    seq += rel;
    rel = 0;
  }
  if (P.span1 && rel + seq > len) seq = len - rel;                 // :2112-2118
  if (seq < 0) seq = 0;  // (a trajectory shorter than the span: nothing of the slice is inside it)
  const int64_t first = __ldg(P.start + t) + rel;
  if (P.seq_out && lane == 0) P.seq_out[s] = seq;
  if (!P.index_out) return;
  const bool padded = P.variable && P.pad_output;
  const int64_t width = (P.variable && !P.pad_output) ? seq : P.seq_length;
  const int64_t off = (P.variable && !P.pad_output) ? __ldg(P.out_offset + s) : s * P.seq_length;
  const int64_t last_real = seq > 0 ? seq - 1 : 0;
  for (int64_t j = lane; j < width; j += 32) {
    const bool real = j < seq;
    int64_t ix = first + (real ? j : last_real);                   // padded steps repeat the last real index, :2151-2158
    ix %= P.storage_length;
    if (ix < 0) ix += P.storage_length;
    P.index_out[off + j] = ix;
    const uint8_t trunc = (j == last_real) ? 1 : 0;                              // :2178-2187
    if (P.truncated_out) P.truncated_out[off + j] = trunc;
    if (P.done_out) P.done_out[off + j] = (P.done_src && __ldg(P.done_src + ix)) ? 1 : trunc;
    if (P.term_out) P.term_out[off + j] = (P.term_src && __ldg(P.term_src + ix)) ? 1 : 0;
    if (padded && P.mask_out) P.mask_out[off + j] = real ? 1 : 0;
  }
}

// ---- prioritized slices: starts that would run past the end of their trajectory get zero mass ------------------
// PrioritizedSliceSampler zeroes them in the sum tree before every draw and restores them afterwards
// (samplers.py:2854-2888, 2910-2918: two tree updates of n_traj * (seq - 1) items per sample).  Here the sampler draws
// from a masked COPY of the leaves: this kernel zeroes the last min(len, seq - 1) leaves of every trajectory in the copy
// (one warp per trajectory), rlb_tree_rebuild turns it into a tree, and the draw is the ordinary rlb_per_sample.
template <typename T>
__global__ void __launch_bounds__(128) slice_mask_starts_kernel(T *leaves, const int64_t *stop, const int64_t *length,
                                                                 int64_t n_traj, int64_t seq_length,
                                                                 int64_t ring_length) {
  const int lane = threadIdx.x & 31;
  const int64_t k = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (k >= n_traj) return;
  const int64_t sp = __ldg(stop + k), len = __ldg(length + k);
  const int64_t m = len < seq_length - 1 ? len : seq_length - 1;
  for (int64_t j = lane; j < m; j += 32) {
    int64_t i = sp - j;
    if (i < 0) i += ring_length;
    leaves[i] = (T)0;
  }
}

}  // namespace rlb

using namespace rlb;

extern "C" {

size_t rlb_traj_table_workspace_bytes(int64_t L) {
  const int64_t tiles = (L + kTrajTile - 1) / kTrajTile + 1;
  // header, per-tile (last end, count, kept count), one flag byte per slot padded to whole tiles
  return 64 + (size_t)tiles * (sizeof(int) * 2 + sizeof(long long)) + 64 + (size_t)tiles * kTrajTile;
}

int rlb_traj_table(const void *signal, int kind, int64_t L, int at_capacity, int64_t cursor, int64_t min_len,
                   int filter, int64_t *start, int64_t *stop, int64_t *length, int64_t *counts, void *workspace,
                   size_t workspace_bytes, rlb_stream_t stream) {
  RLB_REQUIRE(kind == RLB_TRAJ_END || kind == RLB_TRAJ_ID, RLB_EINVAL, "rlb_traj_table: unknown signal kind %d", kind);
  RLB_REQUIRE(L > 0, RLB_EINVAL, "rlb_traj_table: empty storage (L=%lld)", (long long)L);
  RLB_REQUIRE(signal && start && stop && length && counts && workspace, RLB_EINVAL, "rlb_traj_table: null argument");
  RLB_REQUIRE(workspace_bytes >= rlb_traj_table_workspace_bytes(L), RLB_EINVAL, "rlb_traj_table: workspace too small");
  RLB_REQUIRE(cursor < L, RLB_EINVAL, "rlb_traj_table: cursor %lld outside the storage", (long long)cursor);
  TrajParams P;
  memset(&P, 0, sizeof(P));
  P.signal = signal;
  P.kind = kind;
  P.at_capacity = at_capacity ? 1 : 0;
  P.filter = filter ? 1 : 0;
  P.L = L;
  P.cursor = cursor;
  P.min_len = min_len;
  P.start = start;
  P.stop = stop;
  P.length = length;
  P.counts = counts;
  P.n_tiles = (L + kTrajTile - 1) / kTrajTile;
  static_assert(sizeof(TrajWorkspace) <= 64, "workspace header");
  RLB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15u) == 0, RLB_EINVAL, "rlb_traj_table: workspace must be 16-B aligned");
  uint8_t *w = static_cast<uint8_t *>(workspace);
  P.ws = reinterpret_cast<TrajWorkspace *>(w);
  w += 64;
  P.tile_last = reinterpret_cast<long long *>(w);
  w += sizeof(long long) * (size_t)(P.n_tiles + 1);
  P.tile_cnt = reinterpret_cast<int *>(w);
  w += sizeof(int) * (size_t)(P.n_tiles + 1);
  P.tile_kcnt = reinterpret_cast<int *>(w);
  w += sizeof(int) * (size_t)(P.n_tiles + 1);
  w += (16 - (reinterpret_cast<uintptr_t>(w) & 15u)) & 15u;
  P.flags = w;
  RLB_REQUIRE(P.n_tiles < (int64_t(1) << 31), RLB_ELIMIT, "rlb_traj_table: storage too long");
  cudaStream_t st = as_stream(stream);
  traj_mark_kernel<<<(unsigned)P.n_tiles, kTrajThreads, 0, st>>>(P);
  int rc = check_launch("traj_mark_kernel");
  if (rc) return rc;
  traj_emit_kernel<false><<<(unsigned)P.n_tiles, kTrajThreads, 0, st>>>(P);
  rc = check_launch("traj_emit_kernel<count>");
  if (rc || !filter) return rc;
  traj_emit_kernel<true><<<(unsigned)P.n_tiles, kTrajThreads, 0, st>>>(P);
  return check_launch("traj_emit_kernel<kept>");
}

int rlb_slice_index(const int64_t *start, const int64_t *length, int64_t n_traj, const int64_t *traj_draw,
                    const float *u, int64_t num_slices, int64_t seq_length, int64_t storage_length, int variable,
                    int pad_output, int64_t span_left, int64_t span_right, const int64_t *out_offset, int64_t *index_out, uint8_t *truncated_out,
                    uint8_t *mask_out, int64_t *seq_out, const uint8_t *done_src, const uint8_t *term_src,
                    uint8_t *done_out, uint8_t *term_out, rlb_stream_t stream) {
  RLB_REQUIRE(num_slices >= 0 && seq_length > 0 && storage_length > 0 && n_traj > 0, RLB_EINVAL,
              "rlb_slice_index: bad sizes (num_slices=%lld seq_length=%lld storage_length=%lld n_traj=%lld)",
              (long long)num_slices, (long long)seq_length, (long long)storage_length, (long long)n_traj);
  if (num_slices == 0) return RLB_OK;
  RLB_REQUIRE(start && length && traj_draw && u && (index_out || seq_out), RLB_EINVAL, "rlb_slice_index: null argument");
  RLB_REQUIRE(!(variable && !pad_output && index_out) || out_offset, RLB_EINVAL,
              "rlb_slice_index: variable-length output needs out_offset");
  SliceParams P;
  memset(&P, 0, sizeof(P));
  P.start = start;
  P.length = length;
  P.traj_draw = traj_draw;
  P.out_offset = out_offset;
  P.u = u;
  P.n_traj = n_traj;
  P.num_slices = num_slices;
  P.seq_length = seq_length;
  P.storage_length = storage_length;
  RLB_REQUIRE(span_left < seq_length && span_right < seq_length, RLB_EINVAL,
              "rlb_slice_index: The right and left span must be strictly lower than the sequence length");
  RLB_REQUIRE(!(span_left || span_right) || variable, RLB_EINVAL,
              "rlb_slice_index: a span makes slice lengths variable: pass variable=1");
  P.span0 = span_left;
  P.span1 = span_right;
  P.variable = variable ? 1 : 0;
  P.pad_output = pad_output ? 1 : 0;
  P.index_out = index_out;
  P.seq_out = seq_out;
  P.truncated_out = truncated_out;
  P.mask_out = mask_out;
  P.done_src = done_src;
  P.term_src = term_src;
  P.done_out = done_out;
  P.term_out = term_out;
  const int wpb = 4;
  slice_index_kernel<<<(unsigned)((num_slices + wpb - 1) / wpb), 32 * wpb, 0, as_stream(stream)>>>(P);
  return check_launch("slice_index_kernel");
}

int rlb_slice_mask_starts(void *leaves, int dtype, const int64_t *stop, const int64_t *length, int64_t n_traj,
                          int64_t seq_length, int64_t ring_length, rlb_stream_t stream) {
  RLB_REQUIRE(n_traj >= 0 && seq_length > 0 && ring_length > 0, RLB_EINVAL, "rlb_slice_mask_starts: bad sizes");
  if (n_traj == 0 || seq_length == 1) return RLB_OK;
  RLB_REQUIRE(leaves && stop && length, RLB_EINVAL, "rlb_slice_mask_starts: null argument");
  const int wpb = 4;
  const unsigned blocks = (unsigned)((n_traj + wpb - 1) / wpb);
  if (dtype == RLB_F32)
    slice_mask_starts_kernel<float><<<blocks, 32 * wpb, 0, as_stream(stream)>>>((float *)leaves, stop, length, n_traj,
                                                                              seq_length, ring_length);
  else if (dtype == RLB_F64)
    slice_mask_starts_kernel<double><<<blocks, 32 * wpb, 0, as_stream(stream)>>>((double *)leaves, stop, length, n_traj,
                                                                               seq_length, ring_length);
  else
    RLB_REQUIRE(false, RLB_EINVAL, "rlb_slice_mask_starts: unsupported dtype %d", dtype);
  return check_launch("slice_mask_starts_kernel");
}

}  // extern "C"

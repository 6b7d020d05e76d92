// gather.cu -- minibatch row gather / scatter for HBM-resident replay storage (sm_100a).
//
// Replaces TensorStorage.get / .set for tensor indices (data/replay_buffers/storages.py:1242-1263,
// 1028-1096), which in the reference is one aten::index (vectorized_gather_kernel) / index_put_ launch
// PER LEAF plus TensorDict bookkeeping.  Here every leaf of the sampled batch moves in ONE launch:
//
//     dst[k][b, :] = src[k][index[b], :]          (gather)      dst[k][index[b], :] = src[k][b, :]   (scatter)
//
// A pure byte mover: HBM-bandwidth bound, zero reuse, no tensor cores.  Two roles share the grid:
//
//   * bulk-DMA role (CTAs [0, bulk_ctas)) -- wide, 16-B aligned rows (Atari frame stacks: 28 224 B).  The
//     concatenated output byte space of all bulk leaves is cut into equal contiguous ranges, one per
//     warp-pipeline (perfect balance even at B = 256, where there are only ~3 rows per SM).  Each
//     pipeline is driven by ONE elected lane: `cp.async.bulk` global->shared into a ring of 8 KB stages
//     (completion on an mbarrier), then `cp.async.bulk` shared->global out of the stage; the ring keeps
//     up to kAhead loads and kStages-kAhead stores in flight per pipeline with no register staging and
//     no per-byte instructions.  The lanes of the warp prefetch the row indices 32 at a time.
//   * vector role (remaining CTAs) -- narrow or unaligned leaves (actions, rewards, flags, 1.5 KB
//     observations): flat (row, vector) units, widest vector the leaf's alignment allows (16/8/4/2/1 B),
//     4 independent units per thread, streaming cache hints.
//
// Index semantics follow torch indexing: negative indices wrap by `len`; out-of-range indices are reported through
// the status word (torch would raise IndexError) -- a gather reads the clamped row, a scatter drops the write.
#include <stdlib.h>

#include "common.cuh"
#include "tree_range.cuh"

namespace rlb {

// tunables (overridable at build time for the sweeps recorded in profiles/README.md)
#ifndef RLB_GATHER_PIPES
#define RLB_GATHER_PIPES 4
#endif
#ifndef RLB_GATHER_STAGES
#define RLB_GATHER_STAGES 5
#endif
#ifndef RLB_GATHER_AHEAD
#define RLB_GATHER_AHEAD 3
#endif
#ifndef RLB_GATHER_CHUNK
#define RLB_GATHER_CHUNK 8192
#endif
constexpr int kPipes = RLB_GATHER_PIPES;         // DMA pipelines (warps) per bulk CTA
constexpr int kGatherThreads = 32 * kPipes;      // both roles
constexpr int kStages = RLB_GATHER_STAGES;       // ring depth per pipeline (5 x 8 KB x 4 = 161 KB/CTA: leaves room
                                                 // for a co-resident priority-update CTA)
constexpr int kAhead = RLB_GATHER_AHEAD;         // loads kept in flight ahead of the store front
constexpr uint32_t kChunk = RLB_GATHER_CHUNK;    // bytes per stage
constexpr int kVecUnroll = 4;                    // units per thread per tile (vector role)
constexpr int kTileUnits = kGatherThreads * kVecUnroll;
constexpr int64_t kBulkMinRowBytes = 4096;       // AUTO mode: rows at least this wide use the DMA role
#ifndef RLB_GATHER_PEER_RESERVED_SMS
#define RLB_GATHER_PEER_RESERVED_SMS 20
#endif
constexpr int kPeerReservedSms = RLB_GATHER_PEER_RESERVED_SMS;  // SMs a peer-broadcasting launch leaves free

struct GatherLeaf {
  const uint8_t *src;
  uint8_t *dst;
  int64_t row_bytes;
  int64_t stride;   // gather: source row stride; scatter: destination row stride
  int64_t ostride;  // row stride of the OTHER (batch-ordered) side: gather dst / scatter src
  int64_t first;    // first 16-B unit (bulk role) / first tile (vector role) of this leaf
  int64_t units;    // vector role: B * (row_bytes >> vec_log2)
  uint32_t upr;     // vector role: vectors per row
  int32_t vec_log2; // vector role: log2(vector bytes)
  int32_t bulk;     // 1 -> bulk-DMA role
  int32_t foff;     // frame leaf: which frame of the transition's window, relative to its newest one (j - k)
  // frame leaf (de-duplicated frame-stack storage, framestack.cu): the slot index is translated through the slot's
  // frame word  fpos[slot] = env << 40 | position  into the row  env * ring + (position + foff) mod ring  of the
  // frame pool (= src); null for ordinary leaves
  const int64_t *fpos;
  const int64_t *fhead;  // per-env count of frames pushed so far (eviction check), may be null
  int64_t ring;          // frames per env ring
};

struct GatherParams {
  GatherLeaf leaf[RLB_MAX_LEAVES];
  // every destination byte is written to dst + peer_delta[p] for p < n_peers: delta 0 is the local buffer, the
  // others are the same location of the symmetric receive buffer on peer GPUs (NVLink peer memory), so the
  // gathered rows are broadcast by the gather kernel itself -- no separate all-gather
  int64_t peer_delta[RLB_MAX_PEERS];
  // != 0: dst + mc_delta is the same location in a multicast mapping of all ranks' buffers: bulk-role stages are stored
  // once with multimem.st (replicated by the NVSwitch into every member, this GPU included) instead of n_peers copies
  int64_t mc_delta;
  int n_peers;
  int rotate;  // rotate the peer order per piece (see bulk_role)
  const int64_t *index;  // null: the implicit modular range  (ibase + b) % len  (the writer's cursor; B <= len)
  int64_t ibase;
  int64_t B;
  int64_t len;
  int32_t *status;
  int64_t bulk_units;  // total 16-B units over bulk leaves
  int64_t vec_tiles;   // total tiles over vector leaves
  int n_leaves;
  int bulk_ctas;
};

__device__ __forceinline__ int64_t fix_index(int64_t ix, int64_t len, int32_t *status) {
  if (ix < 0) ix += len;
  if (ix < 0 || ix >= len) {
    if (status) atomicOr(status, RLB_STATUS_INDEX_OOB);
    ix = ix < 0 ? 0 : len - 1;
  }
  return ix;
}

// gather side: range-check the slot, then (frame leaves) translate it into the frame-pool row
__device__ __forceinline__ int64_t resolve_row(const GatherParams &P, const GatherLeaf &L, int64_t ix) {
  ix = fix_index(ix, P.len, P.status);
  if (L.fpos) {
    const int64_t w = __ldg(L.fpos + ix);
    const int64_t env = w >> RLB_FRAME_ENV_SHIFT, q = (w & RLB_FRAME_POS_MASK) + L.foff;
    // positions [head - ring, head) of an env's log are still in its ring
    if (L.fhead && P.status && (q < 0 || __ldg(L.fhead + env) - q > L.ring)) atomicOr(P.status, RLB_STATUS_FRAME_EVICTED);
    ix = env * L.ring + (q < 0 ? 0 : q % L.ring);
  }
  return ix;
}

// IMPLICIT (write path only): no index array, slot = (ibase + b) mod len
template <bool IMPLICIT>
__device__ __forceinline__ int64_t load_index(const GatherParams &P, const GatherLeaf &L, int64_t b) {
  if constexpr (!IMPLICIT) {
    return __ldg(P.index + b);
  } else {
    const int64_t ix = P.ibase + b;
    return ix >= P.len ? ix - P.len : ix;
  }
}

// ---- vector role -------------------------------------------------------------------------------------
template <typename V>
__device__ __forceinline__ V ld_stream(const V *p) {
  return __ldcs(p);
}
template <>
__device__ __forceinline__ uint8_t ld_stream<uint8_t>(const uint8_t *p) {
  return *p;
}
template <>
__device__ __forceinline__ uint16_t ld_stream<uint16_t>(const uint16_t *p) {
  return *p;
}
template <typename V>
__device__ __forceinline__ void st_stream(V *p, V v) {
  __stcs(p, v);
}
template <>
__device__ __forceinline__ void st_stream<uint8_t>(uint8_t *p, uint8_t v) {
  *p = v;
}
template <>
__device__ __forceinline__ void st_stream<uint16_t>(uint16_t *p, uint16_t v) {
  *p = v;
}

template <typename V, bool SCATTER, bool IMPLICIT>
__device__ __forceinline__ void vec_tile(const GatherLeaf &L, const GatherParams &P, int64_t tile_in_leaf) {
  const int64_t u0 = tile_in_leaf * kTileUnits + threadIdx.x;
  int64_t b[kVecUnroll], ix[kVecUnroll];
  uint32_t j[kVecUnroll];
  bool ok[kVecUnroll];
#pragma unroll
  for (int k = 0; k < kVecUnroll; ++k) {
    const int64_t u = u0 + (int64_t)k * kGatherThreads;
    ok[k] = u < L.units;
    if (L.units <= 0xffffffffll) {  // 32-bit divide whenever the leaf's unit space allows it
      const uint32_t u32 = ok[k] ? (uint32_t)u : 0u;
      const uint32_t q = u32 / L.upr;
      b[k] = q;
      j[k] = u32 - q * L.upr;
    } else {
      b[k] = ok[k] ? u / L.upr : 0;
      j[k] = ok[k] ? (uint32_t)(u - b[k] * L.upr) : 0;
    }
    ix[k] = ok[k] ? load_index<IMPLICIT>(P, L, b[k]) : 0;
  }
  V val[kVecUnroll];
#pragma unroll
  for (int k = 0; k < kVecUnroll; ++k) {
    if (ok[k]) {
      if constexpr (SCATTER && !IMPLICIT) {
        // a WRITE through an out-of-range index is dropped (and reported), never redirected to slot 0 / len - 1:
        // clamping would silently overwrite somebody else's transition
        int64_t w = ix[k] < 0 ? ix[k] + P.len : ix[k];
        if (w < 0 || w >= P.len) {
          if (P.status) atomicOr(P.status, RLB_STATUS_INDEX_OOB);
          ok[k] = false;
          continue;
        }
        ix[k] = w;
      } else {
        ix[k] = resolve_row(P, L, ix[k]);
      }
      const int64_t srow = SCATTER ? b[k] * L.ostride : ix[k] * L.stride;
      val[k] = ld_stream(reinterpret_cast<const V *>(L.src + srow) + j[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < kVecUnroll; ++k) {
    if (ok[k]) {
      const int64_t drow = SCATTER ? ix[k] * L.stride : b[k] * L.ostride;
      for (int p = 0; p < P.n_peers; ++p)
        st_stream(reinterpret_cast<V *>(L.dst + P.peer_delta[p] + drow) + j[k], val[k]);
    }
  }
}

template <bool SCATTER, bool IMPLICIT = false>
__device__ __forceinline__ void vector_role(const GatherParams &P, int64_t first_tile, int64_t tile_stride) {
  for (int64_t tile = first_tile; tile < P.vec_tiles; tile += tile_stride) {
    int l = -1;
    for (int k = 0; k < P.n_leaves; ++k) {
      if (!P.leaf[k].bulk && P.leaf[k].first <= tile) l = k;  // leaves are laid out in increasing `first`
    }
    if (l < 0) return;
    const GatherLeaf &L = P.leaf[l];
    const int64_t t = tile - L.first;
    switch (L.vec_log2) {
      case 4: vec_tile<uint4, SCATTER, IMPLICIT>(L, P, t); break;
      case 3: vec_tile<uint2, SCATTER, IMPLICIT>(L, P, t); break;
      case 2: vec_tile<uint32_t, SCATTER, IMPLICIT>(L, P, t); break;
      case 1: vec_tile<uint16_t, SCATTER, IMPLICIT>(L, P, t); break;
      default: vec_tile<uint8_t, SCATTER, IMPLICIT>(L, P, t); break;
    }
  }
}

// one 16-byte store into a multicast mapping: the NVSwitch writes it into every GPU bound to the multicast object
__device__ __forceinline__ void multimem_st16(void *mc_addr, const uint4 v) {
  asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "f"(__uint_as_float(v.x)),
               "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
               : "memory");
}

// ---- bulk-DMA role -----------------------------------------------------------------------------------
struct PipeSmem {
  uint64_t full[kStages];   // mbarriers: stage filled by the g->s bulk copy
  uint8_t *dst[kStages];    // destination of the piece staged in each slot
  uint32_t bytes[kStages];
  uint32_t pad_;
};

constexpr size_t kPipeHeaderBytes = (sizeof(PipeSmem) * kPipes + 1023) / 1024 * 1024;

template <bool SCATTER, bool IMPLICIT = false>
__device__ __forceinline__ void bulk_role(const GatherParams &P, uint8_t *ring, PipeSmem *ps, int cta) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int64_t npipes = (int64_t)P.bulk_ctas * kPipes;
  const int64_t pipe = (int64_t)cta * kPipes + warp;
  // equal split of the 16-B unit space
  const int64_t q = P.bulk_units / npipes, rem = P.bulk_units % npipes;
  int64_t pos = pipe * q + (pipe < rem ? pipe : rem);
  const int64_t end = pos + q + (pipe < rem ? 1 : 0);
  if (pos >= end) return;
  uint8_t *my_ring = ring + (size_t)warp * kStages * kChunk;
  PipeSmem *my = ps + warp;
  const uint64_t policy = l2_policy_evict_first();
  if (lane == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&my->full[s], 1);
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  __syncwarp();
  pdl_wait();  // everything above overlapped the tail of the kernel that produced the index (PDL launches)

  // locate the first piece
  int l = 0;
  for (int k = 0; k < P.n_leaves; ++k) {
    if (P.leaf[k].bulk && P.leaf[k].first <= pos) l = k;
  }
  int64_t leaf_end = P.leaf[l].first + P.B * (P.leaf[l].row_bytes >> 4);  // in units
  int64_t byte_in_leaf = (pos - P.leaf[l].first) << 4;
  int64_t b = byte_in_leaf / P.leaf[l].row_bytes;
  int64_t off = byte_in_leaf - b * P.leaf[l].row_bytes;
  // index window: lane j holds index[win0 + j]
  // (gather: already range-checked and, for frame leaves, translated -- one dependent load per 32 rows, not per piece)
  auto window = [&](const GatherLeaf &L, int64_t first) -> int64_t {
    if (first + lane >= P.B) return 0;
    const int64_t raw = load_index<IMPLICIT>(P, L, first + lane);
    if constexpr (SCATTER) return raw; else return resolve_row(P, L, raw);
  };
  int64_t win0 = b;
  int64_t my_ix = window(P.leaf[l], win0);

  int64_t n_loaded = 0, n_stored = 0;
  bool more = true;
  while (more || n_stored < n_loaded) {
    // ---- keep up to kAhead loads in flight ahead of the store front
    while (more && n_loaded < n_stored + kAhead) {
      const GatherLeaf &L = P.leaf[l];
      if (b >= win0 + 32) {  // warp-uniform: refill the index window
        win0 = b;
        my_ix = window(L, win0);
      }
      int64_t ix = __shfl_sync(0xffffffffu, my_ix, (int)(b - win0));
      if constexpr (SCATTER) ix = fix_index(ix, P.len, lane == 0 ? P.status : nullptr);
      int64_t nbytes = L.row_bytes - off;
      if (nbytes > (int64_t)kChunk) nbytes = kChunk;
      const int64_t left = (end - pos) << 4;
      if (nbytes > left) nbytes = left;
      const int stage = (int)(n_loaded % kStages);
      if (lane == 0) {
        // the slot was last used by piece n_loaded - kStages: its s->g copy must have finished READING
        // shared memory.  At most kStages - kAhead younger stores may still be pending.
        if (n_loaded >= kStages) bulk_wait_read<kStages - kAhead>();
        my->dst[stage] = L.dst + (SCATTER ? ix * L.stride : b * L.ostride) + off;
        my->bytes[stage] = (uint32_t)nbytes;
        mbar_arrive_expect_tx(&my->full[stage], (uint32_t)nbytes);
        bulk_g2s_hint(my_ring + (size_t)stage * kChunk, L.src + (SCATTER ? b * L.ostride : ix * L.stride) + off,
                      (uint32_t)nbytes, &my->full[stage], policy);
      }
      ++n_loaded;
      // advance the cursor
      pos += nbytes >> 4;
      off += nbytes;
      if (off == L.row_bytes) {
        off = 0;
        ++b;
      }
      if (pos >= end) {
        more = false;
      } else if (pos >= leaf_end) {  // next bulk leaf
        int nl = l;
        for (int k = 0; k < P.n_leaves; ++k) {
          if (P.leaf[k].bulk && P.leaf[k].first == leaf_end) nl = k;
        }
        l = nl;
        leaf_end = P.leaf[l].first + P.B * (P.leaf[l].row_bytes >> 4);
        b = 0;
        off = 0;
        win0 = 0;
        my_ix = window(P.leaf[l], 0);
      }
    }
    // ---- retire the oldest staged piece: wait for its bytes, then DMA it out
    if (n_stored < n_loaded) {
      const int stage = (int)(n_stored % kStages);
      if (P.mc_delta != 0) {
        // multicast: the whole warp copies the stage out of shared memory, 16 bytes per lane and store
        mbar_wait_parity(&my->full[stage], (uint32_t)((n_stored / kStages) & 1));
        __syncwarp();  // (lane 0 wrote the piece's destination and size)
        const uint8_t *sp = my_ring + (size_t)stage * kChunk;
        uint8_t *dp = my->dst[stage] + P.mc_delta;
        const uint32_t units = my->bytes[stage] >> 4;
        for (uint32_t u0 = 0; u0 < units; u0 += 128) {  // four independent 16-byte loads, then four stores, per lane
          uint4 v[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t u = u0 + (uint32_t)q * 32u + lane;
            if (u < units) v[q] = *reinterpret_cast<const uint4 *>(sp + ((size_t)u << 4));
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t u = u0 + (uint32_t)q * 32u + lane;
            if (u < units) multimem_st16(dp + ((size_t)u << 4), v[q]);
          }
        }
        __syncwarp();  // the stage may be refilled once every lane has read it
      } else if (lane == 0) {
        mbar_wait_parity(&my->full[stage], (uint32_t)((n_stored / kStages) & 1));
        fence_proxy_async_smem();
        // local buffer and, when sharded, every peer's receive buffer.  The destinations are visited in an order that
        // ROTATES from piece to piece and from pipeline to pipeline: every rank runs this same loop, and a fixed order
        // would aim all of them at the same receiver at the same time
        const int np = P.n_peers;
        int p = P.rotate ? (int)((n_stored + (int64_t)blockIdx.x * kPipes + (threadIdx.x >> 5)) % np) : 0;
        for (int q = 0; q < np; ++q) {
          bulk_s2g(my->dst[stage] + P.peer_delta[p], my_ring + (size_t)stage * kChunk, my->bytes[stage]);
          p = p + 1 == np ? 0 : p + 1;
        }
        bulk_commit();
      }
      ++n_stored;
    }
  }
  // the CTA may retire once the DMA engine has READ the last stages (the writes complete on their own and are
  // visible at the kernel boundary); with NVLink peers the publish kernel that follows relies on completed writes,
  // so wait for them in full
  if (lane == 0) {
    if (P.n_peers > 1) bulk_wait_all(); else bulk_wait_read<0>();
  }
  __syncwarp();
}

template <bool SCATTER>
__global__ void __launch_bounds__(kGatherThreads) gather_kernel(const __grid_constant__ GatherParams P) {
  extern __shared__ __align__(128) uint8_t gsmem[];
  pdl_trigger();  // whatever follows in the stream (the trailer kernel, the next gather) may start launching
  if ((int)blockIdx.x < P.bulk_ctas) {
    PipeSmem *ps = reinterpret_cast<PipeSmem *>(gsmem);
    uint8_t *ring = gsmem + kPipeHeaderBytes;  // PipeSmem[kPipes] header; stages stay 128-B aligned
    bulk_role<SCATTER>(P, ring, ps, (int)blockIdx.x);
    return;
  }
  const int vec_ctas = (int)gridDim.x - P.bulk_ctas;
  pdl_wait();
  vector_role<SCATTER>(P, (int64_t)blockIdx.x - P.bulk_ctas, vec_ctas);
}

// The fused write path: TensorStorage.set of a writer batch (storages.py:1028-1096) and the
// default-priority update of the same slots (writers.py:232-235 -> samplers.py:1093-1096) in ONE launch.  CTAs
// [0, row_ctas) move the rows (scatter with the implicit modular index), the last range_ctas CTAs write the trees
// (tree_range.cuh).  The two halves touch disjoint memory and never wait for each other.
template <typename T>
__global__ void __launch_bounds__(kRangeThreads) extend_kernel(const __grid_constant__ GatherParams P,
                                                                const __grid_constant__ RangeParams R,
                                                                int tree_ctas) {
  extern __shared__ __align__(128) uint8_t gsmem[];
  if ((int)blockIdx.x < tree_ctas) {
    range_role<T>(R, (int)blockIdx.x, tree_ctas);
    return;
  }
  if (threadIdx.x >= kGatherThreads) return;  // the row roles are written for kGatherThreads threads
  const int cta = (int)blockIdx.x - tree_ctas, row_ctas = (int)gridDim.x - tree_ctas;
  if (cta < P.bulk_ctas) {
    bulk_role<true, true>(P, gsmem + kPipeHeaderBytes, reinterpret_cast<PipeSmem *>(gsmem), cta);
    return;
  }
  vector_role<true, true>(P, (int64_t)cta - P.bulk_ctas, row_ctas - P.bulk_ctas);
}

constexpr size_t kBulkSmemBytes = kPipeHeaderBytes + (size_t)kPipes * kStages * kChunk;
static_assert(kBulkSmemBytes <= 227 * 1024, "bulk ring exceeds the shared memory of an SM");
static_assert(kAhead >= 1 && kAhead < kStages, "need 1 <= kAhead < kStages");

static int pick_vec_log2(const void *src, const void *dst, int64_t row_bytes, int64_t stride, int64_t ostride) {
  for (int lg = 4; lg > 0; --lg) {
    const uintptr_t a = uintptr_t(1) << lg;
    if (reinterpret_cast<uintptr_t>(src) % a == 0 && reinterpret_cast<uintptr_t>(dst) % a == 0 &&
        row_bytes % (int64_t)a == 0 && stride % (int64_t)a == 0 && ostride % (int64_t)a == 0)
      return lg;
  }
  return 0;
}

// Fills the kernel parameters and the CTA split (bulk_ctas in P, vec_ctas returned).  `index` may be null for the
// implicit modular range starting at `ibase` (write path).
template <bool SCATTER>
static int plan_rows(GatherParams &P, int &vec_ctas_out, const void *const *src, void *const *dst,
                     const int64_t *row_bytes, const int64_t *stride, const int64_t *ostride,
                     const int64_t *peer_delta, int n_peers, int n_leaves, const int64_t *index, int64_t ibase,
                     int64_t B, int64_t len, int mode, int32_t *status, const char *who, int reserved_sms = 0,
                     const rlb_frame_leaf *frames = nullptr, int64_t mc_delta = 0) {
  RLB_REQUIRE(src && dst && row_bytes && stride, RLB_EINVAL, "%s: null argument", who);
  const int sms = sm_count();
  if (sms <= 0) return RLB_ENODEV;

  RLB_REQUIRE(n_peers >= 0 && n_peers <= RLB_MAX_PEERS && (n_peers == 0 || peer_delta), RLB_ELIMIT,
              "%s: n_peers=%d outside [0, %d] or null peer_delta", who, n_peers, RLB_MAX_PEERS);
  memset(&P, 0, sizeof(P));
  RLB_REQUIRE(mc_delta % 16 == 0 && (mc_delta == 0 || !SCATTER), RLB_EINVAL, "%s: bad multicast_delta", who);
  P.mc_delta = mc_delta;
  static const bool no_rotate = [] {  // (A/B measurements: RLB_GATHER_PEER_ROTATE=0 keeps the fixed peer order)
    const char *e = getenv("RLB_GATHER_PEER_ROTATE");
    return e && e[0] == '0';
  }();
  P.rotate = (n_peers > 1 && !no_rotate) ? 1 : 0;
  P.n_peers = n_peers > 0 ? n_peers : 1;  // no peer list = the local buffer only
  for (int p = 0; p < n_peers; ++p) {
    RLB_REQUIRE(peer_delta[p] % 16 == 0, RLB_EINVAL, "%s: peer_delta[%d] is not 16-byte aligned", who, p);
    P.peer_delta[p] = peer_delta[p];
  }
  P.index = index;
  P.ibase = ibase;
  P.B = B;
  P.len = len;
  P.status = status;
  P.n_leaves = n_leaves;
  int64_t bulk_units = 0, vec_tiles = 0;
  for (int k = 0; k < n_leaves; ++k) {
    RLB_REQUIRE(src[k] && dst[k] && row_bytes[k] >= 0 && stride[k] >= row_bytes[k], RLB_EINVAL,
                "%s: leaf %d has a null pointer, negative row_bytes or stride < row_bytes", who, k);
    GatherLeaf &L = P.leaf[k];
    L.src = static_cast<const uint8_t *>(src[k]);
    L.dst = static_cast<uint8_t *>(dst[k]);
    L.row_bytes = row_bytes[k];
    L.stride = stride[k];
    L.ostride = ostride ? ostride[k] : row_bytes[k];
    L.fpos = nullptr;
    L.fhead = nullptr;
    L.ring = 0;
    L.foff = 0;
    if (frames && frames[k].fpos) {
      RLB_REQUIRE(frames[k].ring > 0 && frames[k].offset <= 0, RLB_EINVAL, "%s: leaf %d: bad frame window", who, k);
      L.fpos = frames[k].fpos;
      L.fhead = frames[k].head;
      L.ring = frames[k].ring;
      L.foff = frames[k].offset;
    }
    RLB_REQUIRE(L.ostride >= row_bytes[k], RLB_EINVAL, "%s: leaf %d batch-side stride < row_bytes", who, k);
    const int lg = pick_vec_log2(src[k], dst[k], row_bytes[k], stride[k], L.ostride);
    const bool eligible = lg == 4 && row_bytes[k] >= 16;  // (rlb_scatter asks for the vector role)
    const bool want = (mode == RLB_GATHER_BULK) || (mode == RLB_GATHER_AUTO && row_bytes[k] >= kBulkMinRowBytes);
    if (row_bytes[k] == 0) {
      L.bulk = 0;
      L.first = vec_tiles;
      L.units = 0;
      L.upr = 1;
      L.vec_log2 = 0;
    } else if (eligible && want) {
      L.bulk = 1;
      L.first = bulk_units;
      bulk_units += B * (row_bytes[k] >> 4);
    } else {
      L.bulk = 0;
      L.vec_log2 = lg;
      const int64_t upr = row_bytes[k] >> lg;
      RLB_REQUIRE(upr < (int64_t(1) << 32), RLB_ELIMIT, "%s: leaf %d row too wide for its alignment", who, k);
      L.upr = (uint32_t)upr;
      L.units = B * upr;
      L.first = vec_tiles;
      vec_tiles += (L.units + kTileUnits - 1) / kTileUnits;
    }
  }
  P.bulk_units = bulk_units;
  P.vec_tiles = vec_tiles;
  // grid: one bulk CTA per SM (192 KB of ring each); vector CTAs take SMs of their own when there are few
  // of them, otherwise they oversubscribe (they need no shared memory when no bulk leaf exists).
  int bulk_ctas = 0, vec_ctas = 0;
  if (vec_tiles > 0) {
    const int64_t cap = (int64_t)sms * (bulk_units > 0 ? 1 : 16);
    vec_ctas = (int)(vec_tiles < cap ? vec_tiles : cap);
  }
  if (bulk_units > 0) {
    const int64_t bytes = bulk_units << 4;
    int64_t want_pipes = (bytes + 4095) / 4096;  // at least ~4 KB per pipeline
    int64_t want_ctas = (want_pipes + kPipes - 1) / kPipes;
    int avail = sms - reserved_sms - (vec_ctas < sms / 4 ? vec_ctas : sms / 4);
    if (avail < 1) avail = 1;
    bulk_ctas = (int)(want_ctas < avail ? want_ctas : avail);
    if (bulk_ctas < 1) bulk_ctas = 1;
  }
  P.bulk_ctas = bulk_ctas;
  vec_ctas_out = vec_ctas;
  return RLB_OK;
}

// dynamic shared memory opt-in, once per (kernel, device)
template <typename K>
static int allow_bulk_smem(K kernel, bool *attr_set_dev, const char *name) {
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool &attr_set = attr_set_dev[cur_dev & 63];
  if (attr_set) return RLB_OK;
  int rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBulkSmemBytes),
                      name);
  if (rc) return rc;
  // ask for the largest shared-memory carveout so that a CTA of another kernel (the priority update runs
  // concurrently on a side stream) can still become resident next to a 161 KB gather CTA
  rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                       cudaSharedmemCarveoutMaxShared),
                  name);
  if (rc) return rc;
  attr_set = true;
  return RLB_OK;
}

template <bool SCATTER>
static int launch_rows(const void *const *src, void *const *dst, const int64_t *row_bytes, const int64_t *stride,
                       const int64_t *ostride, const int64_t *peer_delta, int n_peers, int n_leaves,
                       const int64_t *index, int64_t B, int64_t len, int mode, int32_t *status, cudaStream_t st,
                       const char *who, const rlb_frame_leaf *frames = nullptr, int64_t mc_delta = 0) {
  RLB_REQUIRE(n_leaves >= 0 && n_leaves <= RLB_MAX_LEAVES, RLB_ELIMIT, "%s: n_leaves=%d exceeds RLB_MAX_LEAVES=%d",
              who, n_leaves, RLB_MAX_LEAVES);
  RLB_REQUIRE(B >= 0 && len >= 0, RLB_EINVAL, "%s: negative B or len", who);
  if (n_leaves == 0 || B == 0) return RLB_OK;
  RLB_REQUIRE(len > 0, RLB_EINVAL, "%s: cannot index an empty storage (len == 0)", who);
  RLB_REQUIRE(index, RLB_EINVAL, "%s: null argument", who);
  GatherParams P;
  int vec_ctas = 0;
  // with NVLink peers the launch is bound by the wire, not by HBM: leave some SMs to whatever runs beside the
  // exchange (the priority write-back's cluster, the next draw) instead of parking a 161 KB CTA on every one
  const int reserved = n_peers > 1 ? kPeerReservedSms : 0;
  int rc = plan_rows<SCATTER>(P, vec_ctas, src, dst, row_bytes, stride, ostride, peer_delta, n_peers, n_leaves, index,
                              0, B, len, mode, status, who, reserved, frames, mc_delta);
  if (rc) return rc;
  const size_t smem = P.bulk_ctas > 0 ? kBulkSmemBytes : 0;
  static bool attr_set_dev[64] = {};  // function attributes are per device
  if (smem && (rc = allow_bulk_smem(gather_kernel<SCATTER>, attr_set_dev, "cudaFuncSetAttribute(gather_kernel)")))
    return rc;
  rc = check_cuda(launch_pdl(gather_kernel<SCATTER>, dim3(P.bulk_ctas + vec_ctas), dim3(kGatherThreads), smem, st, P),
                  SCATTER ? "gather_kernel<scatter>" : "gather_kernel<gather>");
  if (rc) return rc;
  return check_launch(SCATTER ? "gather_kernel<scatter>" : "gather_kernel<gather>");
}

template <typename T>
static int launch_extend(const GatherParams &P, int row_ctas, const RangeParams &R, int tree_ctas,
                         cudaStream_t st) {
  const size_t smem = P.bulk_ctas > 0 ? kBulkSmemBytes : 0;
  static bool attr_set_dev[64] = {};
  int rc;
  if (smem && (rc = allow_bulk_smem(extend_kernel<T>, attr_set_dev, "cudaFuncSetAttribute(extend_kernel)"))) return rc;
  extend_kernel<T><<<tree_ctas + row_ctas, kRangeThreads, smem, st>>>(P, R, tree_ctas);
  return check_launch("extend_kernel");
}

}  // namespace rlb

using namespace rlb;

extern "C" {

int rlb_gather(const void *const *src, void *const *dst, const int64_t *row_bytes, const int64_t *src_stride_bytes,
               const int64_t *dst_stride_bytes, const int64_t *peer_delta, int n_peers, int n_leaves,
               const int64_t *index, int64_t B, int64_t len, int mode, int32_t *status, rlb_stream_t stream) {
  RLB_REQUIRE(mode == RLB_GATHER_AUTO || mode == RLB_GATHER_VECTOR || mode == RLB_GATHER_BULK, RLB_EINVAL,
              "rlb_gather: unknown mode %d", mode);
  return launch_rows<false>(src, dst, row_bytes, src_stride_bytes, dst_stride_bytes, peer_delta, n_peers, n_leaves,
                            index, B, len, mode, status, as_stream(stream), "rlb_gather");
}

int rlb_gather_ex(const void *const *src, void *const *dst, const int64_t *row_bytes, const int64_t *src_stride_bytes,
                  const int64_t *dst_stride_bytes, int n_leaves, const rlb_gather_opts *opts, const int64_t *index,
                  int64_t B, int64_t len, int mode, int32_t *status, rlb_stream_t stream) {
  RLB_REQUIRE(mode == RLB_GATHER_AUTO || mode == RLB_GATHER_VECTOR || mode == RLB_GATHER_BULK, RLB_EINVAL,
              "rlb_gather_ex: unknown mode %d", mode);
  RLB_REQUIRE(opts, RLB_EINVAL, "rlb_gather_ex: null opts");
  return launch_rows<false>(src, dst, row_bytes, src_stride_bytes, dst_stride_bytes, opts->peer_delta, opts->n_peers,
                            n_leaves, index, B, len, mode, status, as_stream(stream), "rlb_gather_ex", opts->frames,
                            opts->multicast_delta);
}

int rlb_scatter(const void *const *src, void *const *dst, const int64_t *row_bytes, const int64_t *dst_stride_bytes,
                int n_leaves, const int64_t *index, int64_t B, int64_t len, int32_t *status, rlb_stream_t stream) {
  return launch_rows<true>(src, dst, row_bytes, dst_stride_bytes, nullptr, nullptr, 0, n_leaves, index, B, len,
                           RLB_GATHER_VECTOR, status, as_stream(stream), "rlb_scatter");
}

int rlb_extend(const void *const *src, void *const *dst, const int64_t *row_bytes, const int64_t *dst_stride_bytes,
               const int64_t *src_stride_bytes, int n_leaves, int64_t cursor, int64_t n, int64_t max_size,
               void *sum_tree, void *min_tree, int64_t capacity, int dtype, int mode, const void *value, double alpha,
               double eps, double first_default, int has_max, float *max_priority, uint32_t *ticket,
               rlb_stream_t stream) {
  RLB_REQUIRE(n_leaves >= 0 && n_leaves <= RLB_MAX_LEAVES, RLB_ELIMIT, "rlb_extend: n_leaves=%d exceeds %d",
              n_leaves, RLB_MAX_LEAVES);
  RLB_REQUIRE(max_size > 0 && cursor >= 0 && cursor < max_size && n >= 0 && n <= max_size, RLB_EINVAL,
              "rlb_extend: cursor=%lld n=%lld outside max_size=%lld", (long long)cursor, (long long)n,
              (long long)max_size);
  if (n == 0) return RLB_OK;
  const bool trees = sum_tree || min_tree;
  RLB_REQUIRE(trees || n_leaves > 0, RLB_EINVAL, "rlb_extend: nothing to write");
  RangeParams R;
  memset(&R, 0, sizeof(R));
  int tree_ctas = 0, rc;
  if (trees) {
    rc = range_params(R, "rlb_extend", sum_tree, min_tree, capacity, dtype, cursor, n, max_size, mode, value, alpha,
                      eps, first_default, has_max, max_priority, ticket);
    if (rc) return rc;
    tree_ctas = range_ctas(n);
    if (tree_ctas > 8) tree_ctas = 8;  // the movers need the SMs; 2048 threads write the closed-form nodes fast enough
  }
  GatherParams P;
  memset(&P, 0, sizeof(P));
  int vec_ctas = 0;
  if (n_leaves > 0 &&
      (rc = plan_rows<true>(P, vec_ctas, src, dst, row_bytes, dst_stride_bytes, src_stride_bytes, nullptr, 0, n_leaves,
                            nullptr, cursor, n, max_size, RLB_GATHER_AUTO, nullptr, "rlb_extend", tree_ctas)))
    return rc;
  const int row_ctas = P.bulk_ctas + vec_ctas;
  if (trees && dtype == RLB_F64) return launch_extend<double>(P, row_ctas, R, tree_ctas, as_stream(stream));
  return launch_extend<float>(P, row_ctas, R, tree_ctas, as_stream(stream));
}

}  // extern "C"

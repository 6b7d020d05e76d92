// shard.cu -- small fused kernels around the sharded buffer's single all-gather.
//
// The packed minibatch row carries, after the storage leaves, a 20-byte trailer
//     int64 global_index | f32 p_i (leaf priority) | f32 S_r (shard sum) | f32 m_r (shard min)
// rlb_shard_pack fills the trailer of the local rows (one launch instead of five elementwise ones) and
// rlb_shard_weights turns the gathered trailers into importance weights normalised over the WHOLE sharded buffer,
//     w_i = ((p_i / S_r) / min_b (m_b / S_b)) ** -beta ,
// which is the reference formula (samplers.py:945-953) when there is a single shard.  Both are latency-sized
// (B <= a few thousand rows); neither touches the row payload.
//
// The same two kernels carry the split-phase exchange of the NVLink transport (no library barrier on the chain):
// every rank owns `flags[W]` (u64, in its symmetric receive allocation).  rlb_shard_pack -- which runs after the
// gather kernel that pushed the rows into every peer -- ends with a system-scope RELEASE store of this draw's
// sequence number into slot `rank` of every peer's flags; rlb_shard_weights starts with an ACQUIRE spin until
// all W slots of its own flags have reached the sequence number of the draw it finalises.  Sequence numbers are
// kept in device counters the kernels increment themselves, so a captured step replays correctly.
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace rlb {

__device__ __forceinline__ float pow_like_torch_f(float x, float y) {
  if (y == 0.0f) return 1.0f;
  if (y == 1.0f) return x;
  if (y == 0.5f) return sqrtf(x);
  if (y == -0.5f) return rsqrtf(x);
  if (y == -1.0f) return 1.0f / x;
  if (y == 2.0f) return __fmul_rn(x, x);
  if (y == 3.0f) return __fmul_rn(__fmul_rn(x, x), x);
  if (y == -2.0f) return (float)(1.0 / (double)__fmul_rn(x, x));
  return powf(x, y);
}

struct PeerList {
  int64_t delta[RLB_MAX_PEERS];
  int n;
};

__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ONE CTA.  Trailers of the B local rows go to the local buffer and to the same rows of every peer's receive
// buffer; then (flags != null) the draw is published: seq = ++*seq_counter, release-stored into slot `rank` of
// every destination's flag array.  The rows themselves were stored by the gather kernel launched before this one
// on the same stream, so they are complete (kernel boundary) before the release below is issued.
__global__ void __launch_bounds__(1024) shard_pack_kernel(uint8_t *rows, int64_t row_bytes, int64_t meta_off,
                                                          const int64_t *__restrict__ index,
                                                          const float *__restrict__ leaf,
                                                          const float *__restrict__ psum_pmin, int64_t index_base,
                                                          int64_t B, const PeerList peers,
                                                          unsigned long long *flags, unsigned long long *seq_counter,
                                                          int rank) {
  __shared__ unsigned long long s_seq;
  pdl_wait();  // PDL launch: the gather kernel before this one has completed and its stores are performed
  if (flags && threadIdx.x == 0) {
    s_seq = *seq_counter + 1ull;
    *seq_counter = s_seq;
  }
  const float S = psum_pmin[0], mn = psum_pmin[1];
  for (int64_t b = threadIdx.x; b < B; b += blockDim.x) {
    const int64_t gi = index[b] + index_base;
    const float p = leaf[b];
    for (int q = 0; q < peers.n; ++q) {  // the local rows and the same rows of every peer's receive buffer
      uint8_t *m = rows + peers.delta[q] + b * row_bytes + meta_off;  // 8-byte aligned by construction of the layout
      *reinterpret_cast<int64_t *>(m) = gi;
      float *f = reinterpret_cast<float *>(m + 8);
      f[0] = p;
      f[1] = S;
      f[2] = mn;
    }
  }
  if (!flags) return;
  // every thread's trailer stores happen-before the barrier; the releasing store below (system scope, issued by a
  // thread that passed the barrier) is cumulative over them -- the grid-sync pattern: bar.sync, then ONE fence
  __syncthreads();
  if ((int)threadIdx.x < peers.n) {
    unsigned long long *dst =
        reinterpret_cast<unsigned long long *>(reinterpret_cast<uint8_t *>(flags) + peers.delta[threadIdx.x]) + rank;
    st_release_sys_u64(dst, s_seq);
  }
}

// one CTA: pass 1 reduces min_b(m_b / S_b) over all gathered rows, pass 2 writes the weights (and a contiguous
// copy of the global indices for the priority write-back that follows)
__global__ void __launch_bounds__(1024) shard_weights_kernel(const uint8_t *rows, int64_t row_bytes,
                                                             int64_t meta_off, int64_t B, float neg_beta,
                                                             float *__restrict__ weight_out,
                                                             int64_t *__restrict__ index_out,
                                                             const unsigned long long *flags,
                                                             unsigned long long *wait_counter, int n_ranks,
                                                             long long timeout_ns, int32_t *status) {
  if (flags) {
    // close the exchange: every rank's rows of draw number `target` have landed here.  Bounded spin: a peer that
    // died must not hang this GPU -- the status word reports it and the batch is returned as is.
    __shared__ unsigned long long s_target;
    if (threadIdx.x == 0) {
      s_target = *wait_counter + 1ull;
      *wait_counter = s_target;
    }
    __syncthreads();
    if ((int)threadIdx.x < n_ranks) {
      const unsigned long long target = s_target;
      const unsigned long long t0 = global_timer_ns();
      while (ld_acquire_sys_u64(flags + threadIdx.x) < target) {
        if ((long long)(global_timer_ns() - t0) > timeout_ns) {
          if (status) atomicOr(status, RLB_STATUS_EXCHANGE_TIMEOUT);
          break;
        }
        __nanosleep(64);
      }
    }
    __syncthreads();
  }
  __shared__ float sh[32];
  float mn = INFINITY;
  for (int64_t b = threadIdx.x; b < B; b += blockDim.x) {
    // L2-coherent loads: the trailers may have been written by a peer GPU while this kernel was already running
    const float *f = reinterpret_cast<const float *>(rows + b * row_bytes + meta_off + 8);
    mn = fminf(mn, __fdiv_rn(__ldcg(f + 2), __ldcg(f + 1)));
  }
  for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mn;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : INFINITY;
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (threadIdx.x == 0) sh[0] = v;
  }
  __syncthreads();
  const float gmin = sh[0];
  for (int64_t b = threadIdx.x; b < B; b += blockDim.x) {
    const uint8_t *m = rows + b * row_bytes + meta_off;
    const float *f = reinterpret_cast<const float *>(m + 8);
    const float ratio = __fdiv_rn(__ldcg(f), __ldcg(f + 1));
    weight_out[b] = pow_like_torch_f(__fdiv_rn(ratio, gmin), neg_beta);
    if (index_out) index_out[b] = __ldcg(reinterpret_cast<const long long *>(m));
  }
}

}  // namespace rlb

using namespace rlb;

extern "C" {

int rlb_shard_pack(void *rows, int64_t row_bytes, int64_t meta_offset, const int64_t *index, const float *leaf,
                   const float *psum_pmin, int64_t index_base, int64_t B, const int64_t *peer_delta, int n_peers,
                   uint64_t *flags, uint64_t *seq_counter, int rank, rlb_stream_t stream) {
  RLB_REQUIRE(B >= 0 && row_bytes > 0 && meta_offset >= 0 && meta_offset + 20 <= row_bytes && meta_offset % 8 == 0,
              RLB_EINVAL, "rlb_shard_pack: bad layout");
  RLB_REQUIRE(n_peers >= 0 && n_peers <= RLB_MAX_PEERS && (n_peers == 0 || peer_delta), RLB_ELIMIT,
              "rlb_shard_pack: bad peer list");
  RLB_REQUIRE(!flags || (seq_counter && rank >= 0 && rank < RLB_MAX_PEERS && reinterpret_cast<uintptr_t>(flags) % 8 == 0),
              RLB_EINVAL, "rlb_shard_pack: flags need a sequence counter, an 8-byte aligned base and a valid rank");
  if (B == 0 && !flags) return RLB_OK;
  RLB_REQUIRE(B == 0 || (rows && index && leaf && psum_pmin), RLB_EINVAL, "rlb_shard_pack: null pointer");
  RLB_REQUIRE(psum_pmin, RLB_EINVAL, "rlb_shard_pack: null psum_pmin");
  PeerList peers;
  memset(&peers, 0, sizeof(peers));
  peers.n = n_peers > 0 ? n_peers : 1;
  for (int p = 0; p < n_peers; ++p) {
    RLB_REQUIRE(peer_delta[p] % 8 == 0, RLB_EINVAL, "rlb_shard_pack: peer_delta[%d] is not 8-byte aligned", p);
    peers.delta[p] = peer_delta[p];
  }
  int threads = 32;
  while (threads < B && threads < 1024) threads <<= 1;
  int rc = check_cuda(launch_pdl(shard_pack_kernel, dim3(1), dim3(threads), 0, as_stream(stream),
                                 static_cast<uint8_t *>(rows), row_bytes, meta_offset, index, leaf, psum_pmin, index_base,
                                 B, peers, reinterpret_cast<unsigned long long *>(flags),
                                 reinterpret_cast<unsigned long long *>(seq_counter), rank),
                      "shard_pack_kernel");
  if (rc) return rc;
  return check_launch("shard_pack_kernel");
}

int rlb_shard_weights(const void *rows, int64_t row_bytes, int64_t meta_offset, int64_t B, double beta,
                      float *weight_out, int64_t *index_out, const uint64_t *flags, uint64_t *wait_counter,
                      int n_ranks, double timeout_s, int32_t *status, rlb_stream_t stream) {
  RLB_REQUIRE(B >= 0 && row_bytes > 0 && meta_offset >= 0 && meta_offset + 20 <= row_bytes && meta_offset % 8 == 0,
              RLB_EINVAL, "rlb_shard_weights: bad layout");
  RLB_REQUIRE(!flags || (wait_counter && n_ranks > 0 && n_ranks <= RLB_MAX_PEERS), RLB_EINVAL,
              "rlb_shard_weights: flags need a wait counter and 0 < n_ranks <= RLB_MAX_PEERS");
  if (B == 0 && !flags) return RLB_OK;
  RLB_REQUIRE(B == 0 || (rows && weight_out), RLB_EINVAL, "rlb_shard_weights: null pointer");
  int threads = 32;
  while (threads < B && threads < 1024) threads <<= 1;
  const long long timeout_ns = (long long)((timeout_s > 0 ? timeout_s : 10.0) * 1e9);
  shard_weights_kernel<<<1, threads, 0, as_stream(stream)>>>(
      static_cast<const uint8_t *>(rows), row_bytes, meta_offset, B, (float)(-beta), weight_out, index_out,
      reinterpret_cast<const unsigned long long *>(flags), reinterpret_cast<unsigned long long *>(wait_counter),
      n_ranks, timeout_ns, status);
  return check_launch("shard_weights_kernel");
}

}  // extern "C"

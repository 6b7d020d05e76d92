// shard.cu -- small fused kernels around the sharded buffer's single all-gather.
//
// The packed minibatch row carries, after the storage leaves, a 20-byte trailer
//     int64 global_index | f32 p_i (leaf priority) | f32 S_r (shard sum) | f32 m_r (shard min)
// rlb_shard_pack fills the trailer of the local rows (one launch instead of five elementwise ones) and
// rlb_shard_weights turns the gathered trailers into importance weights normalised over the WHOLE sharded buffer,
//     w_i = ((p_i / S_r) / min_b (m_b / S_b)) ** -beta ,
// which is the reference formula (samplers.py:945-953) when there is a single shard.  Both are latency-sized
// (B <= a few thousand rows); neither touches the row payload.
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

__global__ void shard_pack_kernel(uint8_t *rows, int64_t row_bytes, int64_t meta_off, const int64_t *__restrict__ index,
                                  const float *__restrict__ leaf, const float *__restrict__ psum_pmin,
                                  int64_t index_base, int64_t B, const PeerList peers) {
  const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int64_t gi = index[b] + index_base;
  const float p = leaf[b], S = psum_pmin[0], mn = psum_pmin[1];
  for (int q = 0; q < peers.n; ++q) {  // the local rows and the same rows of every peer's receive buffer
    uint8_t *m = rows + peers.delta[q] + b * row_bytes + meta_off;  // 8-byte aligned by construction of the layout
    *reinterpret_cast<int64_t *>(m) = gi;
    float *f = reinterpret_cast<float *>(m + 8);
    f[0] = p;
    f[1] = S;
    f[2] = mn;
  }
}

// one CTA: pass 1 reduces min_b(m_b / S_b) over all gathered rows, pass 2 writes the weights (and a contiguous
// copy of the global indices for the priority write-back that follows)
__global__ void __launch_bounds__(1024) shard_weights_kernel(const uint8_t *__restrict__ rows, int64_t row_bytes,
                                                             int64_t meta_off, int64_t B, float neg_beta,
                                                             float *__restrict__ weight_out,
                                                             int64_t *__restrict__ index_out) {
  __shared__ float sh[32];
  float mn = INFINITY;
  for (int64_t b = threadIdx.x; b < B; b += blockDim.x) {
    const float *f = reinterpret_cast<const float *>(rows + b * row_bytes + meta_off + 8);
    mn = fminf(mn, __fdiv_rn(f[2], f[1]));
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
    const float ratio = __fdiv_rn(f[0], f[1]);
    weight_out[b] = pow_like_torch_f(__fdiv_rn(ratio, gmin), neg_beta);
    if (index_out) index_out[b] = *reinterpret_cast<const int64_t *>(m);
  }
}

}  // namespace rlb

using namespace rlb;

extern "C" {

int rlb_shard_pack(void *rows, int64_t row_bytes, int64_t meta_offset, const int64_t *index, const float *leaf,
                   const float *psum_pmin, int64_t index_base, int64_t B, const int64_t *peer_delta, int n_peers,
                   rlb_stream_t stream) {
  RLB_REQUIRE(B >= 0 && row_bytes > 0 && meta_offset >= 0 && meta_offset + 20 <= row_bytes && meta_offset % 8 == 0,
              RLB_EINVAL, "rlb_shard_pack: bad layout");
  if (B == 0) return RLB_OK;
  RLB_REQUIRE(rows && index && leaf && psum_pmin, RLB_EINVAL, "rlb_shard_pack: null pointer");
  RLB_REQUIRE(n_peers >= 0 && n_peers <= RLB_MAX_PEERS && (n_peers == 0 || peer_delta), RLB_ELIMIT,
              "rlb_shard_pack: bad peer list");
  PeerList peers;
  memset(&peers, 0, sizeof(peers));
  peers.n = n_peers > 0 ? n_peers : 1;
  for (int p = 0; p < n_peers; ++p) peers.delta[p] = peer_delta[p];
  const int threads = 128;
  shard_pack_kernel<<<(unsigned)((B + threads - 1) / threads), threads, 0, as_stream(stream)>>>(
      static_cast<uint8_t *>(rows), row_bytes, meta_offset, index, leaf, psum_pmin, index_base, B, peers);
  return check_launch("shard_pack_kernel");
}

int rlb_shard_weights(const void *rows, int64_t row_bytes, int64_t meta_offset, int64_t B, double beta,
                      float *weight_out, int64_t *index_out, rlb_stream_t stream) {
  RLB_REQUIRE(B >= 0 && row_bytes > 0 && meta_offset >= 0 && meta_offset + 20 <= row_bytes && meta_offset % 8 == 0,
              RLB_EINVAL, "rlb_shard_weights: bad layout");
  if (B == 0) return RLB_OK;
  RLB_REQUIRE(rows && weight_out, RLB_EINVAL, "rlb_shard_weights: null pointer");
  int threads = 32;
  while (threads < B && threads < 1024) threads <<= 1;
  shard_weights_kernel<<<1, threads, 0, as_stream(stream)>>>(static_cast<const uint8_t *>(rows), row_bytes, meta_offset,
                                                            B, (float)(-beta), weight_out, index_out);
  return check_launch("shard_weights_kernel");
}

}  // extern "C"

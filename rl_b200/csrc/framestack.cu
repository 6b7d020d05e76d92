// framestack.cu -- write side of the de-duplicated frame-stack storage (sm_100a); the read side is the frame-leaf
// translation inside the gather kernel (gather.cu: resolve_row).
//
// SURVEY.md section 8(f)-1.  The reference keeps, for every transition, the k-frame observation stack and the k-frame
// next-observation stack it was handed (TensorStorage.set, data/replay_buffers/storages.py:1028-1096): 2k frames per
// transition, of which 2k - 1 repeat frames of the neighbouring transitions of the same episode.  Here every frame of an
// environment's stream is logged once, in that environment's ring of the frame pool:
//
//   episode start   log  obs[i, 0] ... obs[i, k - 1]  (the reset stack, whatever padding produced it), then next[i, k - 1]
//   any other step  log  next[i, k - 1]               (obs[i] is the predecessor's next-stack, already logged)
//
// so that transition i, whose newest frame sits at log position p, always has
//   obs[i, j] = log[p - k + j],  next[i, j] = log[p - k + 1 + j]          (j in [0, k)).
// Its frame word  env << 40 | p  is the only per-transition pixel state; rlb_gather_ex turns (slot -> frame word ->
// pool rows) inside the gather launch.  Byte copies only: a batch read back is bit-identical to what was written as long as
// the stream really is a frame stack (FrameStackStorage(validate=True) checks exactly that).
//
// Two launches per extend: (1) frame_plan_kernel, one CTA per environment: block-wide exclusive scan of the frames
// each step logs (1, or k + 1 at an episode start) -> log positions, frame words, new ring heads; (2) frame_copy_kernel,
// one CTA per (transition, frame): straight 16-byte copies, HBM-bound, ~1.1 frames written per transition.
#include "common.cuh"

namespace rlb {

constexpr int kPlanThreads = 256;
constexpr int kCopyThreads = 128;

struct FramePush {
  const uint8_t *obs, *next;
  int64_t obs_stride, next_stride;  // bytes between consecutive transitions
  const uint8_t *is_init, *done;
  uint8_t *last_done;
  int64_t *head;
  uint8_t *pool;
  int64_t *fpos;
  uint8_t *init_scratch;
  int64_t n, steps, frame_bytes, ring;
  int n_envs, layout, k, pad_;
};

__device__ __forceinline__ int64_t push_row(const FramePush &P, int64_t env, int64_t step) {
  return P.layout == 0 ? env * P.steps + step : step * P.n_envs + env;
}

__global__ void __launch_bounds__(kPlanThreads) frame_plan_kernel(const FramePush P) {
  __shared__ int64_t warp_tot[kPlanThreads / 32];
  const int64_t env = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t per = (P.steps + kPlanThreads - 1) / kPlanThreads;
  const int64_t t0 = (int64_t)tid * per, t1 = min(t0 + per, P.steps);
  auto starts = [&](int64_t t) -> bool {
    if (P.is_init) return P.is_init[push_row(P, env, t)] != 0;
    return t == 0 ? P.last_done[env] != 0 : P.done[push_row(P, env, t - 1)] != 0;
  };
  int64_t mine = 0;
  for (int64_t t = t0; t < t1; ++t) mine += starts(t) ? P.k + 1 : 1;
  // exclusive scan over the CTA
  int64_t incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int64_t up = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += up;
  }
  if (lane == 31) warp_tot[wid] = incl;
  __syncthreads();
  int64_t before = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kPlanThreads / 32; ++w) {
    if (w < wid) before += warp_tot[w];
    total += warp_tot[w];
  }
  const int64_t head0 = P.head[env];
  int64_t pos = head0 + before + incl - mine;
  for (int64_t t = t0; t < t1; ++t) {
    const bool s = starts(t);
    const int64_t row = push_row(P, env, t);
    if (s) pos += P.k;
    P.fpos[row] = (env << RLB_FRAME_ENV_SHIFT) | pos;
    P.init_scratch[row] = s ? 1 : 0;
    ++pos;
  }
  __syncthreads();  // every thread has read head / last_done
  if (tid == 0) {
    P.head[env] = head0 + total;
    if (P.done) P.last_done[env] = P.done[push_row(P, env, P.steps - 1)];
    else if (P.is_init) P.last_done[env] = 0;
  }
}

// grid (n, k + 1): CTA (i, f) copies frame f of transition i's logged frames -- f < k: reset frame obs[i, f] (episode
// starts only), f == k: the newest frame next[i, k - 1]
template <typename V>
__global__ void __launch_bounds__(kCopyThreads) frame_copy_kernel(const FramePush P) {
  const int64_t i = blockIdx.x;
  const int f = blockIdx.y;
  if (f < P.k && !P.init_scratch[i]) return;
  const int64_t w = P.fpos[i];
  const int64_t env = w >> RLB_FRAME_ENV_SHIFT, pos = (w & RLB_FRAME_POS_MASK) - (P.k - f);
  const uint8_t *src = f < P.k ? P.obs + i * P.obs_stride + (int64_t)f * P.frame_bytes
                               : P.next + i * P.next_stride + (int64_t)(P.k - 1) * P.frame_bytes;
  uint8_t *dst = P.pool + (env * P.ring + pos % P.ring) * P.frame_bytes;
  const int64_t units = P.frame_bytes / (int64_t)sizeof(V);
  const V *s = reinterpret_cast<const V *>(src);
  V *d = reinterpret_cast<V *>(dst);
  for (int64_t u = threadIdx.x; u < units; u += kCopyThreads) d[u] = s[u];
}

}  // namespace rlb

using namespace rlb;

extern "C" {

int rlb_framestack_push(const void *obs, const void *next_obs, int64_t obs_row_stride, int64_t next_row_stride,
                        const uint8_t *is_init, const uint8_t *done, uint8_t *last_done, int64_t *head, void *pool,
                        int64_t *fpos_out, uint8_t *init_scratch, int64_t n, int n_envs, int layout, int k,
                        int64_t frame_bytes, int64_t ring, rlb_stream_t stream) {
  RLB_REQUIRE(n >= 0 && n_envs > 0 && k > 0 && frame_bytes > 0 && ring > 0, RLB_EINVAL, "rlb_framestack_push: bad sizes");
  if (n == 0) return RLB_OK;
  RLB_REQUIRE(obs && next_obs && last_done && head && pool && fpos_out && init_scratch, RLB_EINVAL,
              "rlb_framestack_push: null argument");
  RLB_REQUIRE(is_init || done, RLB_EINVAL, "rlb_framestack_push: need is_init or done to find the episode starts");
  RLB_REQUIRE(layout == 0 || layout == 1, RLB_EINVAL, "rlb_framestack_push: layout must be 0 (env-major) or 1");
  RLB_REQUIRE(n % n_envs == 0, RLB_EINVAL, "rlb_framestack_push: %lld transitions do not divide into %d environments",
              (long long)n, n_envs);
  RLB_REQUIRE(n_envs < (1 << 22), RLB_EINVAL, "rlb_framestack_push: too many environments");
  const int64_t steps = n / n_envs;
  // a call may log up to steps * (k + 1) frames per env: they must not lap the ring within the call
  RLB_REQUIRE(steps * (k + 1) <= ring, RLB_EINVAL,
              "rlb_framestack_push: %lld steps x (k + 1) frames exceed the env ring of %lld frames", (long long)steps,
              (long long)ring);
  FramePush P;
  P.obs = (const uint8_t *)obs;
  P.next = (const uint8_t *)next_obs;
  P.obs_stride = obs_row_stride;
  P.next_stride = next_row_stride;
  P.is_init = is_init;
  P.done = done;
  P.last_done = last_done;
  P.head = head;
  P.pool = (uint8_t *)pool;
  P.fpos = fpos_out;
  P.init_scratch = init_scratch;
  P.n = n;
  P.steps = steps;
  P.frame_bytes = frame_bytes;
  P.ring = ring;
  P.n_envs = n_envs;
  P.layout = layout;
  P.k = k;
  P.pad_ = 0;
  cudaStream_t st = as_stream(stream);
  frame_plan_kernel<<<(unsigned)n_envs, kPlanThreads, 0, st>>>(P);
  int rc = check_launch("frame_plan_kernel");
  if (rc != RLB_OK) return rc;
  const dim3 grid((unsigned)n, (unsigned)(k + 1));
  RLB_REQUIRE(n <= 0x7fffffffll && k + 1 <= 65535, RLB_EINVAL, "rlb_framestack_push: batch too large for one launch");
  const uintptr_t align = (uintptr_t)obs | (uintptr_t)next_obs | (uintptr_t)pool | (uintptr_t)obs_row_stride |
                          (uintptr_t)next_row_stride | (uintptr_t)frame_bytes;
  if (align % 16 == 0)
    frame_copy_kernel<uint4><<<grid, kCopyThreads, 0, st>>>(P);
  else if (align % 4 == 0)
    frame_copy_kernel<uint32_t><<<grid, kCopyThreads, 0, st>>>(P);
  else
    frame_copy_kernel<uint8_t><<<grid, kCopyThreads, 0, st>>>(P);
  return check_launch("frame_copy_kernel");
}

}  // extern "C"

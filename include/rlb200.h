/*
 * rlb200.h -- C ABI of the B200-native replay-and-advantage engine (librlb200.so).
 *
 * This is the drop-in boundary for the ONE hot path of pytorch/rl (TorchRL 0.12) that this
 * repository accelerates:  ReplayBuffer.sample (PrioritizedSampler segment-tree sample + storage
 * gather), update_priority (segment-tree write-back) and vec_generalized_advantage_estimate.
 * Each entry point names the reference interface it replaces (paths relative to
 * /root/reference/torchrl/).  INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *   - plain C, no torch / C++ types: raw device pointers, sizes, a CUDA stream handle;
 *   - every pointer marked [dev] is a device pointer owned by the caller (PyTorch);
 *     [host] pointers are ordinary host arrays read during the call;
 *   - the library allocates nothing, never synchronises, never throws: kernels are enqueued on
 *     `stream` (pass torch.cuda.current_stream().cuda_stream) and the call returns
 *     RLB_OK or a negative RLB_E* code; rlb_last_error() gives the message for this thread;
 *   - `dtype`: RLB_F32 or RLB_F64 selects the tree value type, as the reference's
 *     {Sum,Min}SegmentTreeFp{32,64} do (csrc/pybind.cpp:21-34);
 *   - a segment tree is the reference's implicit binary heap: `2*capacity` values, leaf i at
 *     `capacity + i`, node k = op(node 2k, node 2k+1), capacity = smallest power of two
 *     STRICTLY greater than size (csrc/segment_tree.h:44-48).  The layout is bit-compatible
 *     with the reference's CUDA tree tensor `values_` (csrc/cuda_segment_tree.h:29-38).
 *   - there is NO CPU fallback: without a CUDA device every entry point except
 *     rlb_version / rlb_last_error / rlb_tree_capacity / rlb_*_workspace_bytes fails.
 */
#ifndef RLB200_H_
#define RLB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RLB_VERSION 100 /* 0.1.0 */

#define RLB_OK 0
#define RLB_EINVAL (-1)   /* bad argument (null pointer, negative size, unsupported dtype ...) */
#define RLB_ECUDA (-2)    /* a CUDA runtime call or kernel launch failed */
#define RLB_ENODEV (-3)   /* no usable CUDA device */
#define RLB_ELIMIT (-4)   /* argument exceeds a compiled-in limit (e.g. leaves per call) */

#define RLB_F32 0
#define RLB_F64 1

#define RLB_MAX_LEAVES 24 /* leaves gathered by one rlb_gather launch; callers chunk above this */
#define RLB_MAX_PEERS 16  /* destinations (local + NVLink peers) one launch can broadcast to */

/* gather modes (rlb_gather `mode`) */
#define RLB_GATHER_AUTO 0   /* bulk-DMA (TMA) staging for wide 16-B aligned rows, vector path otherwise */
#define RLB_GATHER_VECTOR 1 /* force the 128-bit vectorised LDG/STG path for every leaf */
#define RLB_GATHER_BULK 2   /* force bulk-DMA staging wherever a leaf is eligible */

/* bits of the [dev] int32 status word the kernels OR into (optional, may be NULL) */
#define RLB_STATUS_INDEX_OOB 1    /* gather: an index outside [-len, len) was clamped; scatter: that write was dropped */
#define RLB_STATUS_NONPOS_PSUM 2  /* sample: p_sum <= 0  (samplers.py:911-912, CPU-only check there) */
#define RLB_STATUS_NONPOS_PMIN 4  /* sample: p_min <= 0  (samplers.py:913-914) */
#define RLB_STATUS_BACKOFF_FAIL 8 /* sample: zero-weight back-off ran below index 0 (samplers.py:940-941) */
#define RLB_STATUS_EXCHANGE_TIMEOUT 16 /* sharded exchange: a peer's rows did not arrive within the time limit */

typedef void *rlb_stream_t; /* cudaStream_t */

int rlb_version(void);
const char *rlb_last_error(void);
/* Number of SMs of the current device (grid sizing is a multiple of this); <0 on error. */
int rlb_device_sm_count(void);

/* Keep [ptr, ptr+bytes) (the sampler's sum+min trees: 16 MB for 1M slots) resident in the 126 MB L2 for
 * kernels subsequently launched on (or captured from) `stream`: sets the persisting-L2 carve-out and the
 * stream's access-policy window (hit = persisting, miss = streaming).  The 29 MB that every sampled batch
 * streams through L2 otherwise evicts the lower tree levels between steps.  ptr == NULL clears the window.
 * Returns > 0 (MiB set aside + 1) on success, 0 when the device has no persisting L2, < 0 on error. */
int rlb_l2_persist(const void *ptr, size_t bytes, rlb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Segment tree  -- replaces the pybind classes {Sum,Min}SegmentTreeFp{32,64} /
 * Cuda{Sum,Min}SegmentTreeFp{32,64} of torchrl._torchrl
 * (csrc/segment_tree.h:41-307, csrc/cuda_segment_tree.cu:26-208, csrc/pybind.cpp:21-34).
 * ------------------------------------------------------------------------------------------- */

/* capacity rule of SegmentTree::SegmentTree (csrc/segment_tree.h:44-48). Pure host arithmetic. */
int64_t rlb_tree_capacity(int64_t size);

/* values_.assign(2*capacity, identity) (csrc/segment_tree.h:47; cuda_segment_tree.h:34-37).
 * identity = 0 (sum) / numeric_limits<T>::max() (min). */
int rlb_tree_fill(void *tree /*[dev] 2*capacity values*/, int64_t capacity, int is_min, int dtype,
                  rlb_stream_t stream);

/* LoadValues (csrc/segment_tree.h:200-207): leaves [capacity, capacity+size) are already in place;
 * recompute every internal node bottom-up, node = op(left, right). */
int rlb_tree_rebuild(void *tree /*[dev]*/, int64_t capacity, int is_min, int dtype, rlb_stream_t stream);

/* Bytes of persistent scratch rlb_tree_update / rlb_per_update need for a tree of `size` leaves (one
 * buffer per sampler, zero-initialised ONCE by the caller, shared by the sum and the min tree; calls
 * that share a workspace must be stream-ordered). */
size_t rlb_tree_update_workspace_bytes(int64_t size);

/* SegmentTree::Update, batch form (csrc/segment_tree.h:83-139,216-226; CUDA reference
 * SetLeavesKernel + RecomputeTree, csrc/cuda_segment_tree.cu:26-49,100-132), applied to the sum
 * and the min tree of one sampler in the same call (samplers.py:1077-1078).  Semantics: updates
 * are applied in input order, the LAST duplicate wins; afterwards every ancestor of a touched leaf
 * equals op(left child, right child) -- bit-identical to the serial reference.  Only touched
 * ancestors are recomputed.  `value` holds n elements, or one when scalar != 0.
 * Either tree may be NULL.  `epoch` must be a value never used before with this workspace
 * (callers pass a running counter starting at 1; on wrap-around clear the workspace). */
int rlb_tree_update(void *sum_tree /*[dev]*/, void *min_tree /*[dev]*/, int64_t capacity,
                    const int64_t *index /*[dev] n*/, const void *value /*[dev]*/, int64_t n, int scalar,
                    int dtype, void *workspace /*[dev]*/, size_t workspace_bytes, uint32_t epoch,
                    rlb_stream_t stream);

/* SegmentTree::Query, batch form (csrc/segment_tree.h:143-162; QueryKernel
 * csrc/cuda_segment_tree.cu:51-73): out[i] = op-reduce of leaves [l[i], r[i]).
 * root_fast_path != 0 reproduces the CPU class (returns the root when l<=0 && r>=size);
 * 0 reproduces the CUDA reference, which always walks. */
int rlb_tree_query(const void *tree /*[dev]*/, int64_t size, int64_t capacity, int is_min, int dtype,
                   const int64_t *l /*[dev] n*/, const int64_t *r /*[dev] n*/, void *out /*[dev] n*/,
                   int64_t n, int root_fast_path, rlb_stream_t stream);

/* SegmentTree::At, batch form (csrc/segment_tree.h:56-79,210-214; cuda_segment_tree.h:50-60). */
int rlb_tree_at(const void *tree /*[dev]*/, int64_t capacity, int dtype, const int64_t *index /*[dev] n*/,
                void *out /*[dev] n*/, int64_t n, rlb_stream_t stream);

/* SumSegmentTree::ScanLowerBound, batch form (csrc/segment_tree.h:249-264,289-294;
 * ScanLowerBoundKernel csrc/cuda_segment_tree.cu:75-98): first index whose inclusive prefix sum
 * is >= value; `size` when value > root. */
int rlb_tree_scan_lower_bound(const void *sum_tree /*[dev]*/, int64_t size, int64_t capacity, int dtype,
                              const void *value /*[dev] n*/, int64_t *out /*[dev] n*/, int64_t n,
                              rlb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * PrioritizedSampler.sample arithmetic in ONE launch (data/replay_buffers/samplers.py:895-956):
 *   p_sum = sum.query(0,len); p_min = min.query(0,len); mass = u*p_sum; index = scan_lower_bound(mass)
 *   index.clamp_max_(len-1); leaf = sum[index]; [CPU semantics: zero-weight back-off :935-943]
 *   weight = (leaf / p_min) ** -beta
 * `u` are the uniform draws in [0,1) (the caller keeps torch.rand(B, generator) so the RNG stream
 * is the reference's, samplers.py:918/923).  cpu_semantics != 0 selects the CPU class behaviour
 * (root fast path in query, back-off loop); 0 selects the CUDA reference behaviour.
 * Outputs: index[B] int64, weight[B] fp32 (priority_weight), optional leaf[B] (tree dtype) and
 * psum_pmin[2] (tree dtype).  fp32 arithmetic is never FMA-contracted. */
int rlb_per_sample(const void *sum_tree /*[dev]*/, const void *min_tree /*[dev]*/, int64_t size,
                   int64_t capacity, int dtype, int64_t len, const void *u /*[dev] B, tree dtype*/,
                   int64_t B, double beta, int cpu_semantics, int64_t *index_out /*[dev] B*/,
                   float *weight_out /*[dev] B*/, void *leaf_out /*[dev] B or NULL*/,
                   void *psum_pmin_out /*[dev] 2 or NULL*/, int32_t *status /*[dev] or NULL*/,
                   rlb_stream_t stream);

/* PrioritizedSampler.update_priority arithmetic (samplers.py:1076-1078) fused with the tree
 * write: leaf = (priority + eps) ** alpha in fp32 (torch.pow semantics), then rlb_tree_update on
 * both trees.  index < 0 entries are skipped (MaxValueWriter convention, samplers.py:1040-1052).
 * If max_priority_out != NULL the maximum RAW priority over the valid entries is atomically folded
 * into *max_priority_out (max, no reset): with a buffer the caller initialises once to -inf this IS
 * the reference's running `_max_priority` (samplers.py:1054-1075) without a host round trip.
 * fp32 trees only.  `index_base` is subtracted from every index and entries outside
 * [0, index_limit) are skipped (index_limit < 0 means capacity): a rank of the capacity-sharded buffer
 * passes GLOBAL indices with base = rank * shard_capacity and rewrites only what it owns. */
int rlb_per_update(void *sum_tree /*[dev]*/, void *min_tree /*[dev]*/, int64_t capacity,
                   const int64_t *index /*[dev] n*/, const float *priority /*[dev] n or 1*/, int64_t n,
                   int scalar, double alpha, double eps, float *leaf_scratch /*[dev] n; may be NULL when n <= 1024*/,
                   float *max_priority_out /*[dev] 1 or NULL*/, void *workspace /*[dev]*/,
                   size_t workspace_bytes, uint32_t epoch, int64_t index_base, int64_t index_limit,
                   rlb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Storage gather -- replaces TensorStorage.get for a tensor index
 * (data/replay_buffers/storages.py:1242-1263: storage[:len][index] per leaf, which bottoms out
 * in aten::index / vectorized_gather_kernel).  For every leaf k in ONE launch:
 *     dst[k][b, :] = src[k][index[b], :]        b in [0, B), rows are `row_bytes[k]` bytes,
 * source rows `src_stride_bytes[k]` apart, destination rows `dst_stride_bytes[k]` apart (NULL =
 * contiguous; a stride lets every leaf land in its column of one packed [B, row] buffer, e.g. the
 * send buffer of the sharded buffer's all-gather, with no staging copy).  With n_peers > 0 every destination
 * byte is written to dst + peer_delta[p] for each p (16-B aligned byte offsets; include 0 for the local
 * copy): passing the offsets of the peers' symmetric receive buffers (NVLink peer memory) makes the gather
 * kernel broadcast the rows itself -- gather + all-gather fused in one launch.  Negative indices wrap
 * (index + len) as in torch indexing; out-of-range indices set RLB_STATUS_INDEX_OOB in *status
 * and are clamped (torch raises IndexError; callers that want the exception read the status). */
int rlb_gather(const void *const *src /*[host] n_leaves [dev] pointers*/,
               void *const *dst /*[host] n_leaves [dev] pointers*/, const int64_t *row_bytes /*[host]*/,
               const int64_t *src_stride_bytes /*[host]*/, const int64_t *dst_stride_bytes /*[host] or NULL*/,
               const int64_t *peer_delta /*[host] n_peers or NULL*/, int n_peers, int n_leaves,
               const int64_t *index /*[dev] B*/, int64_t B, int64_t len, int mode,
               int32_t *status /*[dev] or NULL*/, rlb_stream_t stream);

/* ---- de-duplicated frame-stack storage (no reference counterpart: the reference stores the
 * k-frame observation stack AND the k-frame next-observation stack of every transition, storages.py:1028-1096) --------
 * Every frame of an environment's stream is kept ONCE, in that environment's ring of the frame pool
 * pool[n_envs * ring, frame_bytes]; a transition owns the frame word
 *     fpos = env << RLB_FRAME_ENV_SHIFT | position            (position = index of its NEWEST frame in the env's log)
 * and its observation / next-observation stacks are the log positions [position - k, position) / (position - k, position].
 * An episode's first transition logs its k reset frames before its newest one, so the rule has no special cases. */
#define RLB_FRAME_ENV_SHIFT 40
#define RLB_FRAME_POS_MASK ((1ll << RLB_FRAME_ENV_SHIFT) - 1)
#define RLB_STATUS_FRAME_EVICTED 32 /* a gathered transition's frames were already overwritten in the env's ring */

typedef struct rlb_frame_leaf {
  const int64_t *fpos; /* [dev] per-slot frame words, or NULL: an ordinary leaf */
  const int64_t *head; /* [dev] per-env count of frames logged so far (eviction check), or NULL */
  int64_t ring;        /* frames per env ring */
  int32_t offset;      /* which frame of the window: j - k for j in [0, k] (0 = the newest) */
  int32_t reserved;
} rlb_frame_leaf;

/* rlb_gather with options (all optional; a zeroed struct makes it rlb_gather without peers):
 *   frames            leaf k, when frames[k].fpos != NULL, reads row  env * ring + (position + offset) mod ring  of src[k]
 *                     (the frame pool) for the slot index[b]: the stacks are rebuilt by the gather kernel itself, written
 *                     into consecutive frame slots of the batch through dst / dst_stride_bytes;
 *   peer_delta        as in rlb_gather;
 *   multicast_delta   != 0: byte offset from dst[k] to its alias in an NVLink-SHARP multicast mapping of the symmetric
 *                     receive buffers (cuMulticast* / torch symmetric memory `multicast_ptr`).  Wide rows are then stored
 *                     ONCE with multimem.st -- the switch replicates them into every member GPU, this one included --
 *                     instead of n_peers unicast copies; narrow leaves still go through peer_delta. */
typedef struct rlb_gather_opts {
  const rlb_frame_leaf *frames; /* [host] n_leaves entries, or NULL */
  const int64_t *peer_delta;    /* [host] n_peers entries, or NULL */
  int64_t multicast_delta;
  int32_t n_peers;
  int32_t reserved;
} rlb_gather_opts;

int rlb_gather_ex(const void *const *src /*[host]*/, void *const *dst /*[host]*/, const int64_t *row_bytes /*[host]*/,
                  const int64_t *src_stride_bytes /*[host]*/, const int64_t *dst_stride_bytes /*[host] or NULL*/,
                  int n_leaves, const rlb_gather_opts *opts /*[host]*/, const int64_t *index /*[dev] B*/, int64_t B,
                  int64_t len, int mode, int32_t *status /*[dev] or NULL*/, rlb_stream_t stream);

/* Logs the frames of n incoming transitions (n_envs environments x n / n_envs consecutive steps each; layout 0: row
 * i = env * steps + step, the flattened [E, T] collector batch; layout 1: row i = step * n_envs + env) and returns their
 * frame words.  A transition is an episode start when is_init[i] (or, is_init == NULL, when the previous transition
 * of its environment had done = 1 -- carried across calls in last_done, which starts at 1): its k observation frames
 * are logged, then its newest frame  next_obs[i, k - 1]; every other transition logs the newest frame only and
 * shares the rest with its predecessors.  Two launches: the per-env position scan, then the frame copies. */
int rlb_framestack_push(const void *obs /*[dev] n x k x frame_bytes*/, const void *next_obs /*[dev] same*/,
                        int64_t obs_row_stride, int64_t next_row_stride, const uint8_t *is_init /*[dev] n or NULL*/,
                        const uint8_t *done /*[dev] n or NULL*/, uint8_t *last_done /*[dev] n_envs*/,
                        int64_t *head /*[dev] n_envs*/, void *pool /*[dev]*/, int64_t *fpos_out /*[dev] n*/,
                        uint8_t *init_scratch /*[dev] n*/, int64_t n, int n_envs, int layout, int k,
                        int64_t frame_bytes, int64_t ring, rlb_stream_t stream);

/* TensorStorage.set for a tensor cursor (storages.py:1028-1096: storage[cursor] = data per leaf,
 * aten::index_put_):  dst[k][index[b], :] = src[k][b, :].  Duplicate indices: the last wins only
 * if the caller passes unique indices (round-robin writers do, writers.py:190-216). */
int rlb_scatter(const void *const *src /*[host]*/, void *const *dst /*[host]*/,
                const int64_t *row_bytes /*[host]*/, const int64_t *dst_stride_bytes /*[host]*/, int n_leaves,
                const int64_t *index /*[dev] B*/, int64_t B, int64_t len, int32_t *status /*[dev] or NULL*/,
                rlb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Sharded minibatch trailer (new; the reference has no sharded buffer -- SURVEY.md section 8e).  A packed
 * minibatch row ends with  int64 global_index | f32 p_i | f32 S_r | f32 m_r  at byte `meta_offset`
 * (8-byte aligned).  rlb_shard_pack fills the trailer of the B local rows from rlb_per_sample's outputs
 * (index + index_base, leaf, psum_pmin); rlb_shard_weights reads the trailers of the B gathered rows and
 * writes  w_i = ((p_i/S_r) / min_b(m_b/S_b)) ** -beta  (= samplers.py:945-953 for one shard) plus,
 * optionally, a contiguous copy of the global indices.
 *
 * Split-phase exchange (NVLink transport, `flags` != NULL): `flags` is this rank's array of one uint64 per rank
 * inside its symmetric receive allocation (zero-initialised; peers address it with the same `peer_delta` as the
 * rows).  rlb_shard_pack -- enqueued after the rlb_gather that pushed the rows into every peer -- increments
 * *seq_counter and RELEASE-stores the new value (system scope) into slot `rank` of every destination's flags.
 * rlb_shard_weights increments *wait_counter and ACQUIRE-spins until all `n_ranks` slots of its own flags have
 * reached that value (at most `timeout_s` seconds, then RLB_STATUS_EXCHANGE_TIMEOUT is ORed into *status and the
 * kernel carries on), then computes the weights.  Both counters live in device memory and are advanced by the
 * kernels, so a captured step replays correctly; the i-th wait pairs with the i-th publish of every rank. */
int rlb_shard_pack(void *rows /*[dev] B x row_bytes*/, int64_t row_bytes, int64_t meta_offset,
                   const int64_t *index /*[dev] B*/, const float *leaf /*[dev] B*/,
                   const float *psum_pmin /*[dev] 2*/, int64_t index_base, int64_t B,
                   const int64_t *peer_delta /*[host] or NULL*/, int n_peers, uint64_t *flags /*[dev] or NULL*/,
                   uint64_t *seq_counter /*[dev] or NULL*/, int rank, rlb_stream_t stream);
int rlb_shard_weights(const void *rows /*[dev] B x row_bytes*/, int64_t row_bytes, int64_t meta_offset, int64_t B,
                      double beta, float *weight_out /*[dev] B*/, int64_t *index_out /*[dev] B or NULL*/,
                      const uint64_t *flags /*[dev] or NULL*/, uint64_t *wait_counter /*[dev] or NULL*/, int n_ranks,
                      double timeout_s, int32_t *status /*[dev] or NULL*/, rlb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Generalized advantage estimation -- replaces vec_generalized_advantage_estimate /
 * generalized_advantage_estimate for scalar gamma, lmbda
 * (objectives/value/functional.py:270-370 -> _fast_vec_gae :211-267; loop :119-180).
 * Tensors are contiguous [rows, T, F] (time at dim -2), fp32 (RLB_F32) or fp64 (RLB_F64),
 * done/terminated are one byte per element (torch.bool).
 *     delta_t = r_t + gamma*(1-term_t)*v'_t - v_t ;  A_t = delta_t + gammalmbda*(1-done_t)*A_{t+1}
 *     advantage = A ; value_target = A + v
 * `gammalmbda` is the caller's fp32 (or fp64) product gamma*lmbda (the reference multiplies two
 * 0-d tensors, functional.py:249).  One launch, no host sync, forward only. */
int rlb_gae(const void *state_value /*[dev]*/, const void *next_state_value /*[dev]*/,
            const void *reward /*[dev]*/, const uint8_t *done /*[dev]*/, const uint8_t *terminated /*[dev]*/,
            double gamma, double gammalmbda, int64_t rows, int64_t T, int64_t F, int dtype,
            void *advantage /*[dev]*/, void *value_target /*[dev]*/, rlb_stream_t stream);

/* TD(lambda) / TD(1) return -- replaces vec_td_lambda_return_estimate / td_lambda_return_estimate and (lmbda = 1)
 * vec_td1_return_estimate / td1_return_estimate for scalar gamma, lmbda
 * (objectives/value/functional.py:790-899, 993-1210, 464-570, 648-707).  Same reverse scan as rlb_gae:
 *     nv_t = (1-term_t)*v'_t ;  G_t = r_t + gamma*((1-lmbda)*nv_t + lmbda*G'_{t+1}),
 *     G'_{t+1} = nv_t where done_t or t = T-1 (bootstrap at the end of the window), else G_{t+1}.
 * The caller passes gammalmbda = gamma*lmbda and one_minus_lmbda rounded as the reference's tensor ops round
 * them (functional.py:1034-1046). */
int rlb_td_lambda_return(const void *next_state_value /*[dev]*/, const void *reward /*[dev]*/,
                         const uint8_t *done /*[dev]*/, const uint8_t *terminated /*[dev]*/, double gamma,
                         double gammalmbda, double one_minus_lmbda, int64_t rows, int64_t T, int64_t F, int dtype,
                         void *returns /*[dev]*/, rlb_stream_t stream);

/* ---- write path ------------------------------------------------------------------------
 * RoundRobinWriter.extend (writers.py:190-216) writes the slots (cursor + arange(n)) % max_size and gives all of them
 * the sampler's default priority (writers.py:232-235 -> samplers.py:1093-1096).  For such a modular RANGE the tree
 * update needs no sort / merge: see csrc/tree_range.cuh.
 *
 * mode RLB_RANGE_VALUE    *value (tree dtype) is the leaf value as is.
 *      RLB_RANGE_PRIORITY *value (fp32) is a raw priority: leaf = (p + eps) ** alpha; *max_priority <- max(., p).
 *      RLB_RANGE_DEFAULT  p = has_max ? (*max_priority + eps) ** alpha : first_default   (samplers.py:886-893), then
 *                         as RLB_RANGE_PRIORITY -- the whole of PrioritizedSampler.mark_update for a writer batch.
 * ticket: one zero-initialised 32-bit word the kernel leaves at zero (RLB_RANGE_DEFAULT only).  n <= modulo. */
#define RLB_RANGE_VALUE 0
#define RLB_RANGE_PRIORITY 1
#define RLB_RANGE_DEFAULT 2
int rlb_tree_update_range(void *sum_tree /*[dev]*/, void *min_tree /*[dev]*/, int64_t capacity, int dtype,
                          int64_t start, int64_t n, int64_t modulo, int mode, const void *value /*[dev] scalar*/,
                          double alpha, double eps, double first_default, int has_max,
                          float *max_priority /*[dev]*/, uint32_t *ticket /*[dev]*/, rlb_stream_t stream);

/* TensorStorage.set for a writer batch (storages.py:1028-1096) fused with the range update above, ONE launch:
 * dst[k][(cursor + b) % max_size, :] = src[k][b, :] for every leaf k (src_stride_bytes NULL = packed rows) and the trees
 * as rlb_tree_update_range(start = cursor, modulo = max_size).  sum_tree == min_tree == NULL: rows only;
 * n_leaves == 0: trees only. */
int rlb_extend(const void *const *src /*[host] of [dev]*/, void *const *dst /*[host] of [dev]*/,
               const int64_t *row_bytes /*[host]*/, const int64_t *dst_stride_bytes /*[host]*/,
               const int64_t *src_stride_bytes /*[host] or NULL*/, int n_leaves, int64_t cursor, int64_t n,
               int64_t max_size, void *sum_tree /*[dev]*/, void *min_tree /*[dev]*/, int64_t capacity, int dtype,
               int mode, const void *value /*[dev] scalar*/, double alpha, double eps, double first_default,
               int has_max, float *max_priority /*[dev]*/, uint32_t *ticket /*[dev]*/, rlb_stream_t stream);

/* ---- trajectory slices -------------------------------------------------------------------
 * SliceSampler (samplers.py:1207-2300); N-d storages run one table per ring (column) on the host side.
 *
 * rlb_traj_table: the (start, stop, length) table of the trajectories stored in a ring of L slots
 * (_find_start_stop_traj :1652-1706, _end_to_start_stop :1708-1743).  signal: RLB_TRAJ_END = L end-of-trajectory bytes,
 * RLB_TRAJ_ID = L int64 trajectory ids (an end is where the id changes; at capacity the ring closes on slot 0).  Not at
 * capacity the last slot always ends a trajectory; at capacity slot `cursor` (the last one written, -1 = unknown) does,
 * and slot L-1 when there is no end at all.  Entries are ordered by stop.  counts[0] = number of trajectories,
 * counts[1] = how many are at least min_len long; filter != 0 keeps only those in the table (strict_length, :1993-2010).
 * start / stop / length need room for L entries.  workspace: rlb_traj_table_workspace_bytes(L) bytes, zeroed once.
 *
 * rlb_slice_index: _get_index :2058-2215.  Slice s takes trajectory traj_draw[s] (the output of
 * torch.randint(n_traj, ...)) and starts floor(u[s] * (len - seq + 1)) steps into it (fp32 product, as torch.rand() *
 * int64 tensor); index = (start + step) % storage_length; truncated marks the last real step of every slice.
 * variable != 0: slices of trajectories shorter than seq_length are shortened (strict_length = False, :2033-2037); then
 * pad_output pads them to seq_length by repeating the last real index and writes mask, otherwise slice s is written at
 * out_offset[s] (exclusive cumsum of seq_out, obtained by a first call with index_out == NULL).
 * span_left / span_right (SliceSampler(span=...), :2071-2118): 0 = off, -1 = True, k > 0: the slice may start up to k steps
 * (True: seq_length - 1) before its trajectory / run up to k steps past its end; the part outside is cut off, so lengths
 * vary (requires variable != 0).
 * done_src / term_src: the storage's one-byte-per-slot done / terminated flags; done_out = done_src[index] | truncated and
 * term_out = term_src[index] (zero / truncated alone when a source is NULL) -- the info of samplers.py:2190-2205. */
#define RLB_TRAJ_END 0
#define RLB_TRAJ_ID 1
size_t rlb_traj_table_workspace_bytes(int64_t L);
int rlb_traj_table(const void *signal /*[dev]*/, int kind, int64_t L, int at_capacity, int64_t cursor, int64_t min_len,
                   int filter, int64_t *start /*[dev]*/, int64_t *stop /*[dev]*/, int64_t *length /*[dev]*/,
                   int64_t *counts /*[dev] 2*/, void *workspace /*[dev]*/, size_t workspace_bytes, rlb_stream_t stream);
int rlb_slice_index(const int64_t *start /*[dev]*/, const int64_t *length /*[dev]*/, int64_t n_traj,
                    const int64_t *traj_draw /*[dev]*/, const float *u /*[dev]*/, int64_t num_slices, int64_t seq_length,
                    int64_t storage_length, int variable, int pad_output, int64_t span_left, int64_t span_right,
                    const int64_t *out_offset /*[dev] or NULL*/,
                    int64_t *index_out /*[dev] or NULL*/, uint8_t *truncated_out /*[dev] or NULL*/,
                    uint8_t *mask_out /*[dev] or NULL*/, int64_t *seq_out /*[dev] or NULL*/,
                    const uint8_t *done_src /*[dev] or NULL*/, const uint8_t *term_src /*[dev] or NULL*/,
                    uint8_t *done_out /*[dev] or NULL*/, uint8_t *term_out /*[dev] or NULL*/, rlb_stream_t stream);

/* PrioritizedSliceSampler (samplers.py:2575-3028): a slice must not start within the last seq_length - 1 steps of its
 * trajectory.  The reference zeroes those leaves in the sum tree before each draw and restores them (:2854-2918); here the
 * draw uses a masked copy of the leaves: this call zeroes leaves[(stop[k] - j) mod ring_length], j < min(length[k],
 * seq_length - 1), for every trajectory k of an rlb_traj_table (unfiltered); rlb_tree_rebuild + rlb_per_sample follow. */
int rlb_slice_mask_starts(void *leaves /*[dev] the copy's leaf level: tree + capacity*/, int dtype,
                          const int64_t *stop /*[dev]*/, const int64_t *length /*[dev]*/, int64_t n_traj,
                          int64_t seq_length, int64_t ring_length, rlb_stream_t stream);

/* The bare reverse scan  out_t = d_t + c_t * out_{t+1}  (out_T = 0) over contiguous [rows, T, F] coefficient
 * tensors.  V-trace (vtrace_advantage_estimate, functional.py:1297-1382: vs_minus_v) and GAE with per-step
 * gamma / lmbda tensors (functional.py:317-370, rolling) are this scan after an elementwise prologue. */
int rlb_affine_scan(const void *d /*[dev]*/, const void *c /*[dev]*/, int64_t rows, int64_t T, int64_t F, int dtype,
                    void *out /*[dev]*/, rlb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RLB200_H_ */

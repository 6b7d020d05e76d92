/*
 * TEST INFRASTRUCTURE -- CPU oracle for the replay-and-advantage hot path.
 *
 * A plain-C restatement of the reference algorithms, used ONLY by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs as the
 * checker.  The product (rl_b200 + librlb200.so) never links, loads or calls this file.
 *
 * Parity pinning (see DESIGN.md "Oracle"): every function here is checked in
 * tests/test_oracle.py against (a) the known-answer vectors of the reference's own tests
 * (test/rb/test_prioritized.py:113-140, test/rb/test_rb_core.py:598-600) and (b) the
 * UNMODIFIED reference compiled from /root/reference/torchrl/csrc (oracle/_ref, built by
 * oracle/build_ref.py), plus golden GAE vectors produced by importing the reference's
 * Python functionals (tests/golden/make_golden.py).
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC  (see oracle/Makefile).  -ffp-contract=off
 * matters: the reference evaluates  a*b  and  (a*b)+c  as separately rounded fp32 tensor ops.
 *
 * Each function cites the reference file:line (relative to /root/reference/torchrl/) it follows.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * Segment tree  (csrc/segment_tree.h:41-307)
 *   values[2*capacity], leaf i lives at  i | capacity, node = op(left child, right child).
 * ---------------------------------------------------------------------------------------- */

typedef struct {
  int64_t size;
  int64_t capacity;
  int is_min;
  float identity;
  float *values;
} orc_tree;

/* csrc/segment_tree.h:44-48 -- capacity is the smallest power of two STRICTLY greater than size
 * (the loop runs while capacity <= size), every slot starts at the identity element. */
orc_tree *orc_tree_new(int64_t size, int is_min) {
  orc_tree *t = (orc_tree *)malloc(sizeof(orc_tree));
  t->size = size;
  t->is_min = is_min;
  t->identity = is_min ? FLT_MAX : 0.0f; /* segment_tree.h:270,303: T(0) / numeric_limits<T>::max() */
  for (t->capacity = 1; t->capacity <= size; t->capacity <<= 1) {
  }
  t->values = (float *)malloc(sizeof(float) * 2 * (size_t)t->capacity);
  for (int64_t i = 0; i < 2 * t->capacity; ++i) t->values[i] = t->identity;
  return t;
}

void orc_tree_free(orc_tree *t) {
  if (t) {
    free(t->values);
    free(t);
  }
}

int64_t orc_tree_capacity(const orc_tree *t) { return t->capacity; }
int64_t orc_tree_size(const orc_tree *t) { return t->size; }
const float *orc_tree_values(const orc_tree *t) { return t->values; }

static inline float orc_op(const orc_tree *t, float a, float b) {
  /* std::plus / MinOp = std::min(lhs, rhs)  (segment_tree.h:266-298) */
  if (t->is_min) return (b < a) ? b : a;
  return a + b;
}

/* csrc/segment_tree.h:83-88 -- point update, ancestors recomputed bottom-up to the root. */
static inline void orc_tree_update1(orc_tree *t, int64_t index, float value) {
  index |= t->capacity;
  for (t->values[index] = value; index > 1; index >>= 1) {
    t->values[index >> 1] = orc_op(t, t->values[index], t->values[index ^ 1]);
  }
}

/* csrc/segment_tree.h:216-226 -- batch update is a serial loop in input order (last duplicate
 * wins); `scalar` mirrors the value.numel()==1 overloads (:127-139). */
void orc_tree_update(orc_tree *t, const int64_t *index, const float *value, int64_t n, int scalar) {
  for (int64_t i = 0; i < n; ++i) orc_tree_update1(t, index[i], scalar ? value[0] : value[i]);
}

/* csrc/segment_tree.h:56,210-214 */
void orc_tree_at(const orc_tree *t, const int64_t *index, float *out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) out[i] = t->values[index[i] | t->capacity];
}

/* csrc/segment_tree.h:143-162 -- reduce [l, r); whole-range fast path returns the root. */
float orc_tree_query(const orc_tree *t, int64_t l, int64_t r) {
  if (l <= 0 && r >= t->size) return t->values[1];
  float ret = t->identity;
  l |= t->capacity;
  r |= t->capacity;
  while (l < r) {
    if (l & 1) ret = orc_op(t, ret, t->values[l++]);
    if (r & 1) ret = orc_op(t, ret, t->values[--r]);
    l >>= 1;
    r >>= 1;
  }
  return ret;
}

/* csrc/cuda_segment_tree.cu:51-73 -- the CUDA reference's query has NO root fast path; it always
 * walks.  Kept separately because the device sampler path (samplers.py:901-905) uses this one. */
float orc_tree_query_walk(const orc_tree *t, int64_t l, int64_t r) {
  float ret = t->identity;
  l |= t->capacity;
  r |= t->capacity;
  while (l < r) {
    if (l & 1) ret = orc_op(t, ret, t->values[l++]);
    if (r & 1) ret = orc_op(t, ret, t->values[--r]);
    l >>= 1;
    r >>= 1;
  }
  return ret;
}

/* csrc/segment_tree.h:249-264 -- first index whose inclusive prefix sum is >= value.
 * Strict `>` comparisons; returns size_ when value exceeds the root. */
static inline int64_t orc_scan1(const orc_tree *t, float value) {
  if (value > t->values[1]) return t->size;
  int64_t index = 1;
  float cur = value;
  while (index < t->capacity) {
    index <<= 1;
    const float lvalue = t->values[index];
    if (cur > lvalue) {
      cur -= lvalue;
      index |= 1;
    }
  }
  return index ^ t->capacity;
}

/* csrc/segment_tree.h:289-294 */
void orc_tree_scan_lower_bound(const orc_tree *t, const float *value, int64_t *out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) out[i] = orc_scan1(t, value[i]);
}

/* csrc/segment_tree.h:194-207 -- DumpValues / LoadValues (leaves only, full rebuild). */
void orc_tree_dump_leaves(const orc_tree *t, float *out) {
  memcpy(out, t->values + t->capacity, sizeof(float) * (size_t)t->size);
}
void orc_tree_load_leaves(orc_tree *t, const float *leaves) {
  memcpy(t->values + t->capacity, leaves, sizeof(float) * (size_t)t->size);
  for (int64_t i = t->capacity - 1; i > 0; --i)
    t->values[i] = orc_op(t, t->values[i << 1], t->values[(i << 1) | 1]);
}

/* ------------------------------------------------------------------------------------------
 * PrioritizedSampler.sample arithmetic  (data/replay_buffers/samplers.py:895-956)
 *   Given the uniform draws u[B] (the reference takes them from torch.rand(B, generator),
 *   :918/:923) this restates everything that follows, in the reference's operation order:
 *     p_sum = sum_tree.query(0, len); p_min = min_tree.query(0, len)           :904-908
 *     mass  = u * p_sum                      (one fp32 rounding)              :919/:923
 *     index = sum_tree.scan_lower_bound(mass); index.clamp_max_(len - 1)       :927,933
 *     weight = sum_tree[index]                                                  :934
 *     (CPU only) while weight == 0: index -= 1 ...                              :935-943
 *   The final  (weight / p_min) ** -beta  (:953) is torch.pow and is applied by the Python
 *   side of the oracle (oracle/per_oracle.py) so that it is literally the reference's call.
 *   walk_query != 0 selects the CUDA reference's walking query (no root fast path).
 *   Returns 0, or -1 "non-positive p_sum", -2 "non-positive p_min", -3 back-off underflow.
 * ---------------------------------------------------------------------------------------- */
int orc_per_sample(const orc_tree *sum, const orc_tree *mn, int64_t len, const float *u, int64_t B,
                   int cpu_checks, int walk_query, int64_t *index, float *leaf, float *p_sum_out,
                   float *p_min_out) {
  const float p_sum = walk_query ? orc_tree_query_walk(sum, 0, len) : orc_tree_query(sum, 0, len);
  const float p_min = walk_query ? orc_tree_query_walk(mn, 0, len) : orc_tree_query(mn, 0, len);
  *p_sum_out = p_sum;
  *p_min_out = p_min;
  if (cpu_checks) {
    if (p_sum <= 0) return -1; /* samplers.py:911-912 */
    if (p_min <= 0) return -2; /* samplers.py:913-914 */
  }
  for (int64_t i = 0; i < B; ++i) {
    const float mass = u[i] * p_sum;
    int64_t idx = orc_scan1(sum, mass);
    if (idx > len - 1) idx = len - 1;
    float w = sum->values[idx | sum->capacity];
    if (cpu_checks) {
      while (w == 0.0f) { /* samplers.py:937-943 */
        idx -= 1;
        if (idx < 0) return -3;
        w = sum->values[idx | sum->capacity];
      }
    }
    index[i] = idx;
    leaf[i] = w;
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Generalized advantage estimation  (objectives/value/functional.py:119-180, the time loop,
 * which is the semantic ground truth that `vec_generalized_advantage_estimate` :270-370 must
 * agree with -- test/objectives/test_values.py:643-681).
 *   tensors are [rows, T, F] contiguous (time at dim -2, functional.py:147-148)
 *   delta_t = r_t + (gamma * not_terminated_t) * v'_t - v_t                      :170-171
 *   disc_t  = (lmbda * gamma) * not_done_t                                       :172
 *   A_t     = delta_t + A_{t+1} * disc_t, A_T = 0 (prev_advantage = 0)           :169,173-176
 *   target  = A + v                                                              :178
 * gamma / lmbda arrive as fp32 0-d tensors from the GAE module (advantages.py:1456-1467), so
 * lmbda*gamma is itself rounded to fp32 before use; callers pass that product as `gammalmbda`.
 * ---------------------------------------------------------------------------------------- */
void orc_gae_f32(const float *v, const float *nv, const float *r, const uint8_t *done,
                 const uint8_t *term, float gamma, float gammalmbda, int64_t rows, int64_t T,
                 int64_t F, float *adv, float *tgt) {
  for (int64_t b = 0; b < rows; ++b) {
    for (int64_t f = 0; f < F; ++f) {
      float prev = 0.0f;
      for (int64_t t = T - 1; t >= 0; --t) {
        const int64_t i = (b * T + t) * F + f;
        const float g_nt = gamma * (float)(term[i] ? 0 : 1);
        const float gnv = g_nt * nv[i];
        const float s = r[i] + gnv;
        const float delta = s - v[i];
        const float disc = gammalmbda * (float)(done[i] ? 0 : 1);
        const float pd = prev * disc;
        prev = delta + pd;
        adv[i] = prev;
        tgt[i] = prev + v[i];
      }
    }
  }
}

/* float64 evaluation of the same recurrence on the fp32 inputs: the accuracy ground truth that
 * both the reference's fp32 paths and our kernel are measured against. */
void orc_gae_f64(const float *v, const float *nv, const float *r, const uint8_t *done,
                 const uint8_t *term, double gamma, double gammalmbda, int64_t rows, int64_t T,
                 int64_t F, double *adv, double *tgt) {
  for (int64_t b = 0; b < rows; ++b) {
    for (int64_t f = 0; f < F; ++f) {
      double prev = 0.0;
      for (int64_t t = T - 1; t >= 0; --t) {
        const int64_t i = (b * T + t) * F + f;
        const double delta = (double)r[i] + gamma * (term[i] ? 0.0 : 1.0) * (double)nv[i] - (double)v[i];
        const double disc = gammalmbda * (done[i] ? 0.0 : 1.0);
        prev = delta + prev * disc;
        adv[i] = prev;
        tgt[i] = prev + (double)v[i];
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * TD(lambda) return  (objectives/value/functional.py:790-899, rolling branch, scalar gamma / lmbda;
 * lmbda = 1 is TD(1), functional.py:464-570 / 648-707):
 *   nv_t = v'_t * not_terminated_t                                                  :850-853
 *   g = nv_{T-1}; for t = T-1 .. 0:  g = g*(1-done_t) + nv_t*done_t                 :888-894
 *                                    g = ret_t = r_t + gamma*((1-lmbda)*nv_t + lmbda*g)   :895-897
 * ---------------------------------------------------------------------------------------- */
void orc_td_lambda_f32(const float *nv, const float *r, const uint8_t *done, const uint8_t *term, float gamma,
                       float lmbda, int64_t rows, int64_t T, int64_t F, float *ret) {
  const float oml = 1.0f - lmbda;
  for (int64_t b = 0; b < rows; ++b) {
    for (int64_t f = 0; f < F; ++f) {
      const int64_t last = (b * T + (T - 1)) * F + f;
      float g = nv[last] * (float)(term[last] ? 0 : 1);
      for (int64_t t = T - 1; t >= 0; --t) {
        const int64_t i = (b * T + t) * F + f;
        const float nvt = nv[i] * (float)(term[i] ? 0 : 1);
        const float dn = (float)(done[i] ? 1 : 0);
        const float g1 = g * (1.0f - dn);
        const float g2 = nvt * dn;
        g = g1 + g2;
        const float a = oml * nvt;
        const float bb = lmbda * g;
        const float s = a + bb;
        const float gs = gamma * s;
        g = r[i] + gs;
        ret[i] = g;
      }
    }
  }
}

void orc_td_lambda_f64(const float *nv, const float *r, const uint8_t *done, const uint8_t *term, double gamma,
                       double lmbda, int64_t rows, int64_t T, int64_t F, double *ret) {
  for (int64_t b = 0; b < rows; ++b) {
    for (int64_t f = 0; f < F; ++f) {
      const int64_t last = (b * T + (T - 1)) * F + f;
      double g = term[last] ? 0.0 : (double)nv[last];
      for (int64_t t = T - 1; t >= 0; --t) {
        const int64_t i = (b * T + t) * F + f;
        const double nvt = term[i] ? 0.0 : (double)nv[i];
        if (done[i]) g = nvt;
        g = (double)r[i] + gamma * ((1.0 - lmbda) * nvt + lmbda * g);
        ret[i] = g;
      }
    }
  }
}

/* ----------------------------------------------------------------------------------------
 * The bare reverse scan  out_t = d_t + c_t * out_{t+1},  out_T = 0.  This is the python loop of
 * vtrace_advantage_estimate (functional.py:1360-1368: `delta_t + discount_t * c_t * vs_minus_v[-1]`, with
 * c = discount_t * c_t formed first, as python evaluates it) and the recurrence that the rolled gamma tensor of
 * vec_generalized_advantage_estimate (functional.py:317-370, value/utils.py:130-181) unrolls to.
 * Product and sum are rounded separately, as torch's elementwise ops do.
 * ---------------------------------------------------------------------------------------- */
void orc_affine_scan_f32(const float *d, const float *c, int64_t rows, int64_t T, int64_t F, float *out) {
  for (int64_t b = 0; b < rows; ++b)
    for (int64_t f = 0; f < F; ++f) {
      float a = 0.0f;
      for (int64_t t = T - 1; t >= 0; --t) {
        const int64_t i = (b * T + t) * F + f;
        const float m = c[i] * a;
        a = d[i] + m;
        out[i] = a;
      }
    }
}

void orc_affine_scan_f64(const double *d, const double *c, int64_t rows, int64_t T, int64_t F, double *out) {
  for (int64_t b = 0; b < rows; ++b)
    for (int64_t f = 0; f < F; ++f) {
      double a = 0.0;
      for (int64_t t = T - 1; t >= 0; --t) {
        const int64_t i = (b * T + t) * F + f;
        const double m = c[i] * a;
        a = d[i] + m;
        out[i] = a;
      }
    }
}

/* ------------------------------------------------------------------------------------------
 * Storage gather  (data/replay_buffers/storages.py:1242-1263): storage[:len][index] per leaf,
 * i.e. a row copy out[b,:] = src[index[b],:].  The arithmetic lives in torch (aten::index); this
 * byte-level restatement exists so the C-ABI gather can be checked without torch semantics in
 * the loop.  Negative indices wrap like Python/torch indexing (index + len).
 * Returns -1 on an out-of-range index (torch raises IndexError).
 * ---------------------------------------------------------------------------------------- */
int orc_gather_rows(const uint8_t *src, int64_t row_bytes, int64_t src_stride_bytes, int64_t len,
                    const int64_t *index, int64_t B, uint8_t *dst) {
  for (int64_t b = 0; b < B; ++b) {
    int64_t i = index[b];
    if (i < 0) i += len;
    if (i < 0 || i >= len) return -1;
    memcpy(dst + b * row_bytes, src + i * src_stride_bytes, (size_t)row_bytes);
  }
  return 0;
}

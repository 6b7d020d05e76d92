// gae.cu -- generalized advantage estimation as ONE discounted reverse-scan kernel (sm_100a).
//
// Replaces vec_generalized_advantage_estimate -> _fast_vec_gae (objectives/value/functional.py:211-370):
// ~15 torch kernels (masked_scatter padding, cumprod filter, conv1d, boolean un-padding) and >= 2 host
// syncs (value/utils.py:208,283-284) become one launch that reads each input element once and writes
// each output element once (22 B / element in fp32): an HBM-bandwidth-bound streaming kernel, no
// contraction, no tensor cores.
//
// Maths (the reference loop, functional.py:164-178, is the semantic ground truth):
//     delta_t = r_t + gamma*(1-term_t)*v'_t - v_t
//     A_t     = delta_t + c_t * A_{t+1},  c_t = gammalmbda*(1-done_t),  A_T = 0
// i.e. a suffix scan of the affine maps x -> delta_t + c_t*x under composition
//     (c1,d1) o (c2,d2) = (c1*c2, d1 + c1*d2).
// F == 1 layout [rows, T] (time contiguous, functional.py:147-148): one warp per row, each lane owns four
// consecutive time steps (one 128-bit load per input array), composes them serially, the 32 lane maps
// are combined with a 5-step shuffle scan (work-efficient variants buy nothing at warp width), and the
// four local values are then re-derived serially from the lane's incoming value so the result has the
// accuracy of the serial recurrence.  Rows longer than 128 steps are walked tile by tile from the end
// with the carry A_{tile end} held in a register.
// F > 1 layout [rows, T, F]: one thread per (row, feature) column, serial in time, coalesced across
// features.
//
// The same scan serves the sibling estimators: MODE 1 = TD(lambda) / TD(1) returns
// (functional.py:843-899, scalar gamma / lmbda):  G_t = r_t + gamma*((1-lmbda)*nv_t + lmbda*G'_{t+1}),
// nv_t = (1-term_t)*v'_t, G' = nv_t where done_t (or at the last step) -- i.e. the affine map
//     d_t = r_t + gamma*nv_t*(done_t ? 1 : 1-lmbda),   c_t = done_t ? 0 : gamma*lmbda.
// MODE 2 = the bare scan  A_t = d_t + c_t*A_{t+1}  over caller-supplied coefficient tensors: V-trace
// (functional.py:1297-1382) and GAE with per-step gamma / lmbda tensors (functional.py:317-370) reduce to it after an
// elementwise prologue, without the [B, T, T] gamma tensor of value/utils.py:130-181.
#include <stdlib.h>

#include "common.cuh"

namespace rlb {

template <typename T>
struct Vec4 {
  T v[4];
};

template <typename T>
__device__ __forceinline__ Vec4<T> load4(const T *p);
template <>
__device__ __forceinline__ Vec4<float> load4<float>(const float *p) {
  const float4 q = __ldg(reinterpret_cast<const float4 *>(p));
  return {{q.x, q.y, q.z, q.w}};
}
template <>
__device__ __forceinline__ Vec4<double> load4<double>(const double *p) {
  const double2 a = __ldg(reinterpret_cast<const double2 *>(p));
  const double2 b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
  return {{a.x, a.y, b.x, b.y}};
}
__device__ __forceinline__ void store4(float *p, const Vec4<float> &x) {
  __stcs(reinterpret_cast<float4 *>(p), make_float4(x.v[0], x.v[1], x.v[2], x.v[3]));
}
__device__ __forceinline__ void store4(double *p, const Vec4<double> &x) {
  __stcs(reinterpret_cast<double2 *>(p), make_double2(x.v[0], x.v[1]));
  __stcs(reinterpret_cast<double2 *>(p) + 1, make_double2(x.v[2], x.v[3]));
}

constexpr int kGaeWarpsPerCta = 8;
constexpr int kGaeTile = 128;  // time steps per warp iteration (32 lanes x 4)

// VEC: rows are 16-B aligned for T (and 4-B aligned for the flag bytes) and T % 4 == 0.
// MODE 0: GAE (outputs advantage and value_target; needs v).  MODE 1: TD(lambda) return (output `adv` only; `v`
// and `tgt` are unused, `oml` = 1 - lmbda).
template <typename T, int MODE>
__device__ __forceinline__ void scan_coeffs(T vv, T nvv, T rr, bool dn, bool tm, bool last, T gamma, T gl, T oml,
                                            T &d, T &c) {
  if constexpr (MODE == 0) {
    d = (rr + (tm ? (T)0 : gamma) * nvv) - vv;
    c = dn ? (T)0 : gl;
  } else if constexpr (MODE == 2) {
    d = rr;   // `r` carries d_t
    c = nvv;  // `nv` carries c_t
  } else {
    const bool cut = dn || last;
    const T nvt = tm ? (T)0 : nvv;
    d = rr + (nvt * gamma) * (cut ? (T)1 : oml);
    c = cut ? (T)0 : gl;
  }
}

// One 128-step tile of one row, as seen by one warp: each lane's four (d, c) pairs + state values.
template <typename T, bool VEC, int MODE>
struct GaeTile {
  T d[4], c[4], sv[4];

  __device__ __forceinline__ void load(const T *__restrict__ v, const T *__restrict__ nv, const T *__restrict__ r,
                                       const uint8_t *__restrict__ done, const uint8_t *__restrict__ term, int64_t base,
                                       int64_t t0, int64_t Tlen, T gamma, T gl, T oml, int64_t es = 1) {
    // es: element stride between consecutive steps (F for one feature column of a [T, F] row; scalar path only)
    if (VEC && t0 + 4 <= Tlen) {
      Vec4<T> qv = {{(T)0, (T)0, (T)0, (T)0}};
      if constexpr (MODE == 0) qv = load4<T>(v + base + t0);
      const Vec4<T> qn = load4<T>(nv + base + t0);
      const Vec4<T> qr = load4<T>(r + base + t0);
      uchar4 qd = make_uchar4(0, 0, 0, 0), qt = make_uchar4(0, 0, 0, 0);
      if constexpr (MODE != 2) {
        qd = __ldg(reinterpret_cast<const uchar4 *>(done + base + t0));
        qt = __ldg(reinterpret_cast<const uchar4 *>(term + base + t0));
      }
      const uint8_t dd[4] = {qd.x, qd.y, qd.z, qd.w};
      const uint8_t tt[4] = {qt.x, qt.y, qt.z, qt.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        sv[j] = qv.v[j];
        scan_coeffs<T, MODE>(qv.v[j], qn.v[j], qr.v[j], dd[j] != 0, tt[j] != 0, t0 + j == Tlen - 1, gamma, gl, oml,
                             d[j], c[j]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t t = t0 + j;
        if (t < Tlen) {
          const int64_t i = base + t * es;
          const T vv = (MODE == 0) ? __ldg(v + i) : (T)0;
          sv[j] = vv;
          const bool dn = (MODE != 2) ? (__ldg(done + i) != 0) : false;
          const bool tm = (MODE != 2) ? (__ldg(term + i) != 0) : false;
          scan_coeffs<T, MODE>(vv, __ldg(nv + i), __ldg(r + i), dn, tm, t == Tlen - 1, gamma, gl, oml, d[j], c[j]);
        } else {  // beyond the row: A = 0 there, contributes nothing
          sv[j] = (T)0;
          d[j] = (T)0;
          c[j] = (T)0;
        }
      }
    }
  }

  // inclusive suffix scan of the lane maps: afterwards A_{t0(lane)} = Bs + Cs * (A entering the tile from the right)
  __device__ __forceinline__ void scan(int lane, T &Bs, T &Cs) const {
    T Bq = d[3], Cq = c[3];  // lane-local composition: A_{t0} = Bq + Cq * A_{t0+4}
#pragma unroll
    for (int j = 2; j >= 0; --j) {
      Bq = d[j] + c[j] * Bq;
      Cq = c[j] * Cq;
    }
    Bs = Bq;
    Cs = Cq;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const T Bo = __shfl_down_sync(0xffffffffu, Bs, o);
      const T Co = __shfl_down_sync(0xffffffffu, Cs, o);
      if (lane + o < 32) {
        Bs = Bs + Cs * Bo;
        Cs = Cs * Co;
      }
    }
  }

  // the four local values re-derived serially from the value entering the lane; returns A at the tile's first step
  __device__ __forceinline__ T finish(int lane, T Bs, T Cs, T carry, int64_t base, int64_t t0, int64_t Tlen,
                                      T *__restrict__ adv, T *__restrict__ tgt, int64_t es = 1) const {
    const T a_first = Bs + Cs * carry;
    // value entering this lane from the right = A at the first step of lane+1 (carry for lane 31)
    T a_next = __shfl_down_sync(0xffffffffu, a_first, 1);
    if (lane == 31) a_next = carry;
    Vec4<T> oa, ot;
#pragma unroll
    for (int j = 3; j >= 0; --j) {
      a_next = d[j] + c[j] * a_next;
      oa.v[j] = a_next;
      ot.v[j] = a_next + sv[j];  // value_target = advantage + state_value (functional.py:178)
    }
    if (VEC && t0 + 4 <= Tlen) {
      store4(adv + base + t0, oa);
      if constexpr (MODE == 0) store4(tgt + base + t0, ot);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (t0 + j < Tlen) {
          adv[base + (t0 + j) * es] = oa.v[j];
          if constexpr (MODE == 0) tgt[base + (t0 + j) * es] = ot.v[j];
        }
      }
    }
    return __shfl_sync(0xffffffffu, oa.v[0], 0);
  }
};

// Many rows: one WARP per row (blockDim.x / 32 rows per CTA); rows longer than a tile are walked from the end.
// F > 1 (scalar path): `rows` counts (row, feature) COLUMNS of [rows / F, T, F] tensors; column q = row * F + f starts
// at row * T * F + f and steps by F.  The F warps of a row sit next to each other in a CTA, so the row's cache lines are
// fetched once and shared through L1 (a thread-per-column walk touches 4 bytes of every 32-byte sector per step).
template <typename T, bool VEC, int MODE>
__global__ void __launch_bounds__(kGaeWarpsPerCta * 32) gae_rows_kernel(
    const T *__restrict__ v, const T *__restrict__ nv, const T *__restrict__ r, const uint8_t *__restrict__ done,
    const uint8_t *__restrict__ term, T gamma, T gl, T oml, int64_t rows, int64_t Tlen, T *__restrict__ adv,
    T *__restrict__ tgt, int64_t F) {
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;  // whole warps leave together
  pdl_wait();
  const int64_t base = VEC ? row * Tlen : (row / F) * Tlen * F + (row % F);
  const int64_t ntiles = (Tlen + kGaeTile - 1) / kGaeTile;
  T carry = (T)0;  // A at the first step of the tile processed before (later in time); prev_advantage = 0
  for (int64_t tile = ntiles - 1; tile >= 0; --tile) {
    const int64_t t0 = tile * kGaeTile + 4 * lane;  // first of this lane's four steps
    GaeTile<T, VEC, MODE> g;
    g.load(v, nv, r, done, term, base, t0, Tlen, gamma, gl, oml, F);
    T Bs, Cs;
    g.scan(lane, Bs, Cs);
    carry = g.finish(lane, Bs, Cs, carry, base, t0, Tlen, adv, tgt, F);
  }
}

// [rows, T, F] with F in {2, 3, 4} and T % 4 == 0 (16-byte aligned arrays): a lane's 4 steps x F features are 4F CONTIGUOUS
// values of the row, so every array moves with the same coalesced 128-bit accesses as the F == 1 tile; the F scans of a lane
// run side by side in registers.
template <typename T, int F, int MODE>
__global__ void __launch_bounds__(kGaeWarpsPerCta * 32) gae_rows_f_kernel(
    const T *__restrict__ v, const T *__restrict__ nv, const T *__restrict__ r, const uint8_t *__restrict__ done,
    const uint8_t *__restrict__ term, T gamma, T gl, T oml, int64_t rows, int64_t Tlen, T *__restrict__ adv,
    T *__restrict__ tgt) {
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  pdl_wait();
  const int64_t base = row * Tlen * F;
  const int64_t ntiles = (Tlen + kGaeTile - 1) / kGaeTile;
  T carry[F];
#pragma unroll
  for (int f = 0; f < F; ++f) carry[f] = (T)0;
  for (int64_t tile = ntiles - 1; tile >= 0; --tile) {
    const int64_t t0 = tile * kGaeTile + 4 * lane;
    T d[4 * F], c[4 * F], sv[4 * F];  // element (step j, feature f) at [j * F + f]
    const bool in = t0 < Tlen;        // (T % 4 == 0: a lane's four steps are inside the row together)
    if (in) {
      const int64_t e0 = base + t0 * F;
#pragma unroll
      for (int q = 0; q < F; ++q) {
        Vec4<T> qv = {{(T)0, (T)0, (T)0, (T)0}};
        if constexpr (MODE == 0) qv = load4<T>(v + e0 + 4 * q);
        const Vec4<T> qn = load4<T>(nv + e0 + 4 * q);
        const Vec4<T> qr = load4<T>(r + e0 + 4 * q);
        uchar4 qd = make_uchar4(0, 0, 0, 0), qt = make_uchar4(0, 0, 0, 0);
        if constexpr (MODE != 2) {
          qd = __ldg(reinterpret_cast<const uchar4 *>(done + e0 + 4 * q));
          qt = __ldg(reinterpret_cast<const uchar4 *>(term + e0 + 4 * q));
        }
        const uint8_t dd[4] = {qd.x, qd.y, qd.z, qd.w};
        const uint8_t tt[4] = {qt.x, qt.y, qt.z, qt.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int e = 4 * q + k;  // flat position among the lane's 4F values: step e / F
          sv[e] = qv.v[k];
          scan_coeffs<T, MODE>(qv.v[k], qn.v[k], qr.v[k], dd[k] != 0, tt[k] != 0, t0 + e / F == Tlen - 1, gamma, gl, oml,
                               d[e], c[e]);
        }
      }
    } else {
#pragma unroll
      for (int e = 0; e < 4 * F; ++e) {
        d[e] = (T)0;
        c[e] = (T)0;
        sv[e] = (T)0;
      }
    }
    T oa[4 * F], ot[4 * F];
#pragma unroll
    for (int f = 0; f < F; ++f) {
      T Bs = d[3 * F + f], Cs = c[3 * F + f];  // lane-local composition, then the warp scan (as GaeTile::scan)
#pragma unroll
      for (int j = 2; j >= 0; --j) {
        Bs = d[j * F + f] + c[j * F + f] * Bs;
        Cs = c[j * F + f] * Cs;
      }
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const T Bo = __shfl_down_sync(0xffffffffu, Bs, o);
        const T Co = __shfl_down_sync(0xffffffffu, Cs, o);
        if (lane + o < 32) {
          Bs = Bs + Cs * Bo;
          Cs = Cs * Co;
        }
      }
      const T a_first = Bs + Cs * carry[f];
      T a_next = __shfl_down_sync(0xffffffffu, a_first, 1);
      if (lane == 31) a_next = carry[f];
#pragma unroll
      for (int j = 3; j >= 0; --j) {  // the four local values re-derived serially (accuracy of the serial recurrence)
        a_next = d[j * F + f] + c[j * F + f] * a_next;
        oa[j * F + f] = a_next;
        ot[j * F + f] = a_next + sv[j * F + f];
      }
      carry[f] = __shfl_sync(0xffffffffu, oa[f], 0);
    }
    if (in) {
      const int64_t e0 = base + t0 * F;
#pragma unroll
      for (int q = 0; q < F; ++q) {
        Vec4<T> xa, xt;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          xa.v[k] = oa[4 * q + k];
          xt.v[k] = ot[4 * q + k];
        }
        store4(adv + e0 + 4 * q, xa);
        if constexpr (MODE == 0) store4(tgt + e0 + 4 * q, xt);
      }
    }
  }
}

// Few rows (the reference benchmark's [300,500], [32,512], [1,512]): one CTA per row, one WARP per TILE.  All tiles of a
// chunk are loaded at once (the serial walk above pays one DRAM round trip per tile); lane 0 of each warp publishes
// its tile's composite map, and every warp folds the maps of the tiles to its right -- the serial recurrence at tile
// granularity -- to get the value entering its tile.  Rows longer than blockDim.x * 4 steps go chunk by chunk.
constexpr int kGaeMaxRowWarps = 16;
template <typename T, bool VEC, int MODE>
__global__ void __launch_bounds__(kGaeMaxRowWarps * 32) gae_row_cta_kernel(
    const T *__restrict__ v, const T *__restrict__ nv, const T *__restrict__ r, const uint8_t *__restrict__ done,
    const uint8_t *__restrict__ term, T gamma, T gl, T oml, int64_t rows, int64_t Tlen, T *__restrict__ adv,
    T *__restrict__ tgt) {
  __shared__ T sB[kGaeMaxRowWarps], sC[kGaeMaxRowWarps];
  pdl_trigger();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int64_t row = blockIdx.x;
  const int64_t base = row * Tlen;
  const int64_t ntiles = (Tlen + kGaeTile - 1) / kGaeTile;
  pdl_wait();
  T chunk_carry = (T)0;
  for (int64_t hi = ntiles; hi > 0; hi -= nw) {      // tiles [hi - nw, hi) of this chunk, later tiles first
    const int64_t tile = hi - nw + warp;             // may be negative in the last chunk: an empty tile
    const int64_t t0 = tile * kGaeTile + 4 * lane;
    GaeTile<T, VEC, MODE> g;
    T Bs = (T)0, Cs = (T)1;                          // identity map for an empty tile
    if (tile >= 0) {
      g.load(v, nv, r, done, term, base, t0, Tlen, gamma, gl, oml);
      g.scan(lane, Bs, Cs);
    }
    if (lane == 0) {
      sB[warp] = Bs;
      sC[warp] = Cs;
    }
    __syncthreads();
    T cin = chunk_carry, mine = chunk_carry;
    for (int w = nw - 1; w >= 0; --w) {              // value entering tile w = A at the first step of tile w + 1
      if (w == warp) mine = cin;
      cin = sB[w] + sC[w] * cin;
    }
    chunk_carry = cin;                               // A at the first step of the chunk
    if (tile >= 0) g.finish(lane, Bs, Cs, mine, base, t0, Tlen, adv, tgt);
    __syncthreads();
  }
}

// [rows, T, F] with F > 1: thread per (row, f) column; consecutive threads -> consecutive f (coalesced).
template <typename T, int MODE>
__global__ void __launch_bounds__(256) gae_cols_kernel(const T *__restrict__ v, const T *__restrict__ nv,
                                                       const T *__restrict__ r, const uint8_t *__restrict__ done,
                                                       const uint8_t *__restrict__ term, T gamma, T gl, T oml,
                                                       int64_t rows, int64_t Tlen, int64_t F, T *__restrict__ adv,
                                                       T *__restrict__ tgt) {
  const int64_t col = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (col >= rows * F) return;
  const int64_t row = col / F, f = col - row * F;
  const int64_t base = row * Tlen * F + f;
  T a = (T)0;
#pragma unroll 4
  for (int64_t t = Tlen - 1; t >= 0; --t) {
    const int64_t i = base + t * F;
    const T vv = (MODE == 0) ? __ldg(v + i) : (T)0;
    T dlt, cc;
    const bool dn = (MODE != 2) ? (__ldg(done + i) != 0) : false;
    const bool tm = (MODE != 2) ? (__ldg(term + i) != 0) : false;
    scan_coeffs<T, MODE>(vv, __ldg(nv + i), __ldg(r + i), dn, tm, t == Tlen - 1, gamma, gl, oml, dlt, cc);
    a = dlt + cc * a;
    adv[i] = a;
    if constexpr (MODE == 0) tgt[i] = a + vv;
  }
}

template <typename T, int MODE>
static int gae_impl(const void *v, const void *nv, const void *r, const uint8_t *done, const uint8_t *term,
                    double gamma, double gl, double oml, int64_t rows, int64_t Tlen, int64_t F, void *adv, void *tgt,
                    cudaStream_t st) {
  const T *pv = static_cast<const T *>(v), *pn = static_cast<const T *>(nv), *pr = static_cast<const T *>(r);
  T *pa = static_cast<T *>(adv), *pt = static_cast<T *>(tgt);
  if (F == 1) {
    auto al = [](const void *p, uintptr_t a) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) % a) == 0; };
    const bool vec = (Tlen % 4 == 0) && al(v, 16) && al(nv, 16) && al(r, 16) && al(adv, 16) && al(tgt, 16) &&
                     al(done, 4) && al(term, 4);
    const int sms = sm_count();
    const int64_t ntiles = (Tlen + kGaeTile - 1) / kGaeTile;
    if (ntiles >= 2 && rows <= 4 * (int64_t)sms) {  // few long rows: CTA per row, warp per tile
      const int nw = (int)(ntiles < kGaeMaxRowWarps ? ntiles : kGaeMaxRowWarps);
      if (vec)
        return check_cuda(launch_pdl(gae_row_cta_kernel<T, true, MODE>, dim3((unsigned)rows), dim3(nw * 32), 0, st, pv,
                                     pn, pr, done, term, (T)gamma, (T)gl, (T)oml, rows, Tlen, pa, pt),
                          "gae_row_cta_kernel");
      return check_cuda(launch_pdl(gae_row_cta_kernel<T, false, MODE>, dim3((unsigned)rows), dim3(nw * 32), 0, st, pv,
                                   pn, pr, done, term, (T)gamma, (T)gl, (T)oml, rows, Tlen, pa, pt),
                        "gae_row_cta_kernel");
    }
    // many rows: warp per row; with few (short) rows one warp per CTA so that they spread over the SMs
    const int wpc = rows >= 8 * (int64_t)sms ? kGaeWarpsPerCta : (rows >= 2 * (int64_t)sms ? 2 : 1);
    const int64_t blocks = (rows + wpc - 1) / wpc;
    RLB_REQUIRE(blocks < (int64_t(1) << 31), RLB_ELIMIT, "rlb_gae: too many rows for one launch");
    if (vec)
      return check_cuda(launch_pdl(gae_rows_kernel<T, true, MODE>, dim3((unsigned)blocks), dim3(wpc * 32), 0, st, pv,
                                   pn, pr, done, term, (T)gamma, (T)gl, (T)oml, rows, Tlen, pa, pt, (int64_t)1),
                        "gae_rows_kernel");
    return check_cuda(launch_pdl(gae_rows_kernel<T, false, MODE>, dim3((unsigned)blocks), dim3(wpc * 32), 0, st, pv,
                                 pn, pr, done, term, (T)gamma, (T)gl, (T)oml, rows, Tlen, pa, pt, (int64_t)1),
                      "gae_rows_kernel");
  }
  {
    // F in {2, 3, 4}: contiguous 4F-value lanes (gae_rows_f_kernel) when the vector accesses are aligned
    auto al = [](const void *p, uintptr_t a) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) % a) == 0; };
    const bool vecf = F >= 2 && F <= 4 && (Tlen % 4 == 0) && al(v, 16) && al(nv, 16) && al(r, 16) && al(adv, 16) &&
                      al(tgt, 16) && al(done, 4) && al(term, 4);
    static const bool no_f = [] {
      const char *e = getenv("RLB_GAE_NO_F_KERNEL");
      return e && e[0] == '1';
    }();
    if (vecf && !no_f) {
      const int sms = sm_count();
      const int wpc = rows >= 8 * (int64_t)sms ? kGaeWarpsPerCta : (rows >= 2 * (int64_t)sms ? 2 : 1);
      const int64_t blocks = (rows + wpc - 1) / wpc;
      RLB_REQUIRE(blocks < (int64_t(1) << 31), RLB_ELIMIT, "rlb_gae: too many rows for one launch");
      const dim3 grid((unsigned)blocks), block(wpc * 32);
      cudaError_t e;
      if (F == 2)
        e = launch_pdl(gae_rows_f_kernel<T, 2, MODE>, grid, block, 0, st, pv, pn, pr, done, term, (T)gamma, (T)gl, (T)oml,
                       rows, Tlen, pa, pt);
      else if (F == 3)
        e = launch_pdl(gae_rows_f_kernel<T, 3, MODE>, grid, block, 0, st, pv, pn, pr, done, term, (T)gamma, (T)gl, (T)oml,
                       rows, Tlen, pa, pt);
      else
        e = launch_pdl(gae_rows_f_kernel<T, 4, MODE>, grid, block, 0, st, pv, pn, pr, done, term, (T)gamma, (T)gl, (T)oml,
                       rows, Tlen, pa, pt);
      return check_cuda(e, "gae_rows_f_kernel");
    }
  }
  static const bool force_cols = [] {  // (A/B measurements: RLB_GAE_COLUMN_KERNEL=1 keeps the thread-per-column walk)
    const char *e = getenv("RLB_GAE_COLUMN_KERNEL");
    return e && e[0] == '1';
  }();
  if (F <= 16 && Tlen >= 16 && !force_cols) {
    // narrow feature dim: a warp per (row, feature) column, time-parallel like F == 1 (see gae_rows_kernel)
    const int64_t cols = rows * F;
    const int64_t blocks = (cols + kGaeWarpsPerCta - 1) / kGaeWarpsPerCta;
    RLB_REQUIRE(blocks < (int64_t(1) << 31), RLB_ELIMIT, "rlb_gae: too many columns for one launch");
    return check_cuda(launch_pdl(gae_rows_kernel<T, false, MODE>, dim3((unsigned)blocks), dim3(kGaeWarpsPerCta * 32), 0,
                                 st, pv, pn, pr, done, term, (T)gamma, (T)gl, (T)oml, cols, Tlen, pa, pt, F),
                      "gae_rows_kernel");
  }
  const int64_t cols = rows * F;
  const int64_t blocks = (cols + 255) / 256;
  RLB_REQUIRE(blocks < (int64_t(1) << 31), RLB_ELIMIT, "rlb_gae: too many columns for one launch");
  gae_cols_kernel<T, MODE><<<(unsigned)blocks, 256, 0, st>>>(pv, pn, pr, done, term, (T)gamma, (T)gl, (T)oml, rows,
                                                             Tlen, F, pa, pt);
  return check_launch("gae_cols_kernel");
}

}  // namespace rlb

using namespace rlb;

extern "C" int rlb_gae(const void *state_value, const void *next_state_value, const void *reward,
                       const uint8_t *done, const uint8_t *terminated, double gamma, double gammalmbda,
                       int64_t rows, int64_t T, int64_t F, int dtype, void *advantage, void *value_target,
                       rlb_stream_t stream) {
  RLB_REQUIRE(rows >= 0 && T >= 0 && F >= 1, RLB_EINVAL, "rlb_gae: bad shape rows=%lld T=%lld F=%lld",
              (long long)rows, (long long)T, (long long)F);
  if (rows == 0 || T == 0) return RLB_OK;
  RLB_REQUIRE(state_value && next_state_value && reward && done && terminated && advantage && value_target,
              RLB_EINVAL, "rlb_gae: null pointer");
  if (dtype == RLB_F32)
    return gae_impl<float, 0>(state_value, next_state_value, reward, done, terminated, gamma, gammalmbda, 0.0, rows, T,
                              F, advantage, value_target, as_stream(stream));
  if (dtype == RLB_F64)
    return gae_impl<double, 0>(state_value, next_state_value, reward, done, terminated, gamma, gammalmbda, 0.0, rows,
                               T, F, advantage, value_target, as_stream(stream));
  RLB_REQUIRE(false, RLB_EINVAL, "rlb_gae: unsupported dtype %d", dtype);
}

extern "C" int rlb_td_lambda_return(const void *next_state_value, const void *reward, const uint8_t *done,
                                    const uint8_t *terminated, double gamma, double gammalmbda,
                                    double one_minus_lmbda, int64_t rows, int64_t T, int64_t F, int dtype,
                                    void *returns, rlb_stream_t stream) {
  RLB_REQUIRE(rows >= 0 && T >= 0 && F >= 1, RLB_EINVAL, "rlb_td_lambda_return: bad shape rows=%lld T=%lld F=%lld",
              (long long)rows, (long long)T, (long long)F);
  if (rows == 0 || T == 0) return RLB_OK;
  RLB_REQUIRE(next_state_value && reward && done && terminated && returns, RLB_EINVAL,
              "rlb_td_lambda_return: null pointer");
  if (dtype == RLB_F32)
    return gae_impl<float, 1>(nullptr, next_state_value, reward, done, terminated, gamma, gammalmbda, one_minus_lmbda,
                              rows, T, F, returns, nullptr, as_stream(stream));
  if (dtype == RLB_F64)
    return gae_impl<double, 1>(nullptr, next_state_value, reward, done, terminated, gamma, gammalmbda,
                               one_minus_lmbda, rows, T, F, returns, nullptr, as_stream(stream));
  RLB_REQUIRE(false, RLB_EINVAL, "rlb_td_lambda_return: unsupported dtype %d", dtype);
}

extern "C" int rlb_affine_scan(const void *d, const void *c, int64_t rows, int64_t T, int64_t F, int dtype, void *out,
                               rlb_stream_t stream) {
  RLB_REQUIRE(rows >= 0 && T >= 0 && F >= 1, RLB_EINVAL, "rlb_affine_scan: bad shape rows=%lld T=%lld F=%lld",
              (long long)rows, (long long)T, (long long)F);
  if (rows == 0 || T == 0) return RLB_OK;
  RLB_REQUIRE(d && c && out, RLB_EINVAL, "rlb_affine_scan: null pointer");
  if (dtype == RLB_F32)
    return gae_impl<float, 2>(nullptr, c, d, nullptr, nullptr, 0.0, 0.0, 0.0, rows, T, F, out, nullptr,
                              as_stream(stream));
  if (dtype == RLB_F64)
    return gae_impl<double, 2>(nullptr, c, d, nullptr, nullptr, 0.0, 0.0, 0.0, rows, T, F, out, nullptr,
                               as_stream(stream));
  RLB_REQUIRE(false, RLB_EINVAL, "rlb_affine_scan: unsupported dtype %d", dtype);
}

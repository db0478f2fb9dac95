// Plain-state HMM word discoverers of hmm/ (states = concept tokens of the caption).
//
//   prob domain: hmm/hmm_word_discoverer.py       forward :110-125, backward :127-138,
//                updateInitialCounts :140-151, updateTransitionCounts :153-186,
//                updateObservationCounts :188-203, M-step :275-296, align :301-329
//   log  domain: hmm/audio_hmm_word_discoverer.py forward :148-168, backward :170-185,
//                counts :187-254, M-step :354-389, align :396-427
//
// Mapping: one warp per caption pair, lane j = state j (n <= 16); alpha history of the pair in
// shared memory; per-warp register accumulators for the init / transition counts (persistent
// warps, fixed pair->warp assignment, so the two-level reduction is deterministic); the
// observation counts go through a static postings index (slots sorted by (concept, phone) table
// entry) and a segmented reduction -- no atomics anywhere.  The path is latency/HBM-bound
// (~210 B and ~3.5 kFLOP per pair), so the design goal is many independent warps, not tensor cores.
#include <stdlib.h>

#include <mutex>

#include "mwd_common.cuh"

namespace mwd {

constexpr int kHmmWarpsPerSm = 24;

int hmm_warps_total() { return sm_count() * kHmmWarpsPerSm; }

struct HmmArgs {
  const int32_t* tgt_off;
  const int32_t* tgt;
  const int32_t* src_off;
  const int32_t* src;
  const int64_t* slot_off;
  const double* init;    // row of this n
  const double* trans;   // table of this n
  const double* obs;
  double* pair_ll;
  double* post;
  double* part_init;     // full partial tables
  double* part_trans;
  double* alpha_out;
  double* beta_out;
  const double* emis;    // dense log emissions (n_slots) or null
  int64_t lo, hi;
  int n, Vf, Tmax, warps_per_cta, total_warps;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, s));
  return v;
}
// scipy.special.logsumexp over the lanes (inactive lanes pass -inf)
__device__ __forceinline__ double warp_lse(double v) {
  double m = warp_max(v);
  if (!(fabs(m) < INFINITY)) m = 0.0;
  double s = warp_sum(exp(v - m));
  return log(s) + m;
}
__device__ __forceinline__ double lse2(double a, double b) {
  if (a == -INFINITY) return b;
  if (b == -INFINITY) return a;
  double m = fmax(a, b);
  return m + log(exp(a - m) + exp(b - m));
}

// NN > 0: the bucket's state count is a compile-time constant (state loops unroll, the per-lane
// accumulator arrays shrink from kNMax to NN registers); NN == 0: generic (n > 8).
template <bool LOG, int NN>
__global__ void __launch_bounds__(256) hmm_estep_kernel(const HmmArgs a) {
  constexpr int NU = NN > 0 ? NN : kNMax;
  const int n = NN > 0 ? NN : a.n, Vf = a.Vf;
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  const int j = lane;
  const bool on = j < n;
  extern __shared__ double smem[];
  double* s_A = smem;                       // [n][n]
  double* s_pi = s_A + kNMax * kNMax;       // [n]
  const int per_warp = a.Tmax * n + kNMax + kNMax * kNMax;
  double* s_al = s_pi + kNMax + (size_t)wic * per_warp;   // [Tmax][n]
  double* s_x = s_al + (size_t)a.Tmax * n;                // [n]
  double* s_E = s_x + kNMax;                              // [n][n] (log trans counts)
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) s_A[e] = a.trans[e];
  for (int e = threadIdx.x; e < n; e += blockDim.x) s_pi[e] = a.init[e];
  __syncthreads();

  const double ident = LOG ? -INFINITY : 0.0;
  double init_acc = ident;
  double tacc[NU];
#pragma unroll
  for (int i = 0; i < NU; ++i) tacc[i] = ident;

  const int gw = blockIdx.x * a.warps_per_cta + wic;
  if (wic < a.warps_per_cta) {
    for (int64_t pair = a.lo + gw; pair < a.hi; pair += a.total_warps) {
      const int f0 = a.src_off[pair];
      const int T = a.src_off[pair + 1] - f0;
      const int32_t* f = a.src + f0;
      const int ej = on ? a.tgt[a.tgt_off[pair] + j] : 0;
      const double* orow = a.obs + (size_t)ej * Vf;
      const int64_t slot0 = a.slot_off[pair];
      auto emis = [&](int t) -> double {
        if (LOG && a.emis) return on ? a.emis[slot0 + (int64_t)t * n + j] : 0.0;   // continuous model
        double b = on ? orow[f[t]] : 0.0;
        return (b != b) ? 0.0 : b;           // absent pair: 0 in both classes (:122 / :161)
      };
      // ------------------------------------------------------------ forward
      double al;
      {
        double b0 = on ? ((LOG && a.emis) ? a.emis[slot0 + j] : orow[f[0]]) : 0.0;
        if (LOG) al = on ? s_pi[j] + b0 : -INFINITY;          // :158 (absent -> KeyError upstream)
        else al = on ? s_pi[j] * ((b0 != b0) ? 0.0 : b0) : 0.0;   // :114-118
      }
      if (on) s_al[j] = al;
      for (int t = 0; t + 1 < T; ++t) {
        __syncwarp();
        const double b = emis(t + 1);
        const double* at = s_al + (size_t)t * n;
        if (LOG) {
          double m = -INFINITY;
#pragma unroll
          for (int i = 0; i < NU; ++i)
            if (NN > 0 || i < n) m = fmax(m, s_A[i * n + (on ? j : 0)] + at[i]);
          if (!(fabs(m) < INFINITY)) m = 0.0;
          double s = 0.0;
#pragma unroll
          for (int i = 0; i < NU; ++i)
            if (NN > 0 || i < n) s += exp(s_A[i * n + (on ? j : 0)] + at[i] - m);
          al = log(s) + m + b;                                 // :165
        } else {
          double acc = 0.0;
#pragma unroll
          for (int i = 0; i < NU; ++i)
            if (NN > 0 || i < n) acc = fma(s_A[i * n + (on ? j : 0)], at[i], acc);
          al = acc * b;                                        // :123
        }
        if (on) s_al[(size_t)(t + 1) * n + j] = al;
      }
      __syncwarp();
      double ll;
      {
        double last = on ? s_al[(size_t)(T - 1) * n + j] : (LOG ? -INFINITY : 0.0);
        ll = LOG ? warp_lse(last) : log(warp_sum(last));       // :312 / :244-245
      }
      if (lane == 0) a.pair_ll[pair] = ll;
      // ------------------------------------------------------------ backward + counts
      if (LOG) {
        if (T >= 2) {   // transition counts from the LAST t only (:204-229)
          const double bl = emis(T - 1);
          for (int i = 0; i < n; ++i)
            if (on) s_E[i * n + j] = s_al[(size_t)(T - 2) * n + i] + s_A[i * n + j] + bl;   // beta_{T-1} = 0
          __syncwarp();
          if (on) {
#pragma unroll
            for (int i = 0; i < NU; ++i) {
              if (NN == 0 && i >= n) break;
              const int dlt = j - i;
              double m = -INFINITY;
              for (int r = 0; r < n; ++r) {
                int c = r + dlt;
                if (c >= 0 && c < n) m = fmax(m, s_E[r * n + c]);
              }
              if (!(fabs(m) < INFINITY)) m = 0.0;
              double s = 0.0;
              for (int r = 0; r < n; ++r) {
                int c = r + dlt;
                if (c >= 0 && c < n) s += exp(s_E[r * n + c] - m);
              }
              tacc[i] = lse2(tacc[i], log(s) + m);
            }
          }
          __syncwarp();
        }
        double beta = on ? 0.0 : -INFINITY;
        double ic = -INFINITY, nrm = -INFINITY;
        for (int t = T - 1; t >= 0; --t) {
          const double alv = on ? s_al[(size_t)t * n + j] : -INFINITY;
          const double v = on ? alv + beta : -INFINITY;
          if (a.alpha_out && on) { a.alpha_out[slot0 + (int64_t)t * n + j] = alv; a.beta_out[slot0 + (int64_t)t * n + j] = beta; }
          ic = lse2(ic, v);
          nrm = lse2(nrm, warp_lse(v));
          if (on) s_al[(size_t)t * n + j] = v;
          if (t > 0) {
            const double bb = beta + emis(t);
            if (on) s_x[j] = bb;
            __syncwarp();
            double m = -INFINITY;
#pragma unroll
            for (int c = 0; c < NU; ++c)
              if (NN > 0 || c < n) m = fmax(m, s_A[(on ? j : 0) * n + c] + s_x[c]);
            if (!(fabs(m) < INFINITY)) m = 0.0;
            double s = 0.0;
#pragma unroll
            for (int c = 0; c < NU; ++c)
              if (NN > 0 || c < n) s += exp(s_A[(on ? j : 0) * n + c] + s_x[c] - m);
            beta = on ? log(s) + m : -INFINITY;                // :182
            __syncwarp();
          }
        }
        init_acc = lse2(init_acc, ic);                         // :192-194
        __syncwarp();
        for (int t = 0; t < T; ++t)
          if (on) a.post[slot0 + (int64_t)t * n + j] = s_al[(size_t)t * n + j] - nrm;   // :244-246
        __syncwarp();
      } else {
        double beta = on ? 1.0 : 0.0;
        for (int t = T - 1; t >= 0; --t) {
          const double alv = on ? s_al[(size_t)t * n + j] : 0.0;
          const double g = alv * beta;
          const double G = warp_sum(g);
          const double gam = g / G;                            // :147-148, :193
          if (on) {
            init_acc += gam;
            a.post[slot0 + (int64_t)t * n + j] = gam;
            if (a.alpha_out) { a.alpha_out[slot0 + (int64_t)t * n + j] = alv; a.beta_out[slot0 + (int64_t)t * n + j] = beta; }
          }
          if (t > 0) {
            const double bb = beta * emis(t);
            // xi_{t-1}[i][j] = (alpha_{t-1}[i] * bb[j]) * A[i][j], normalised over (i, j)  (:161-162)
            const double* ap = s_al + (size_t)(t - 1) * n;
            double xv[NU];
            double col = 0.0;
#pragma unroll
            for (int i = 0; i < NU; ++i) {
              xv[i] = (i < n && on) ? (ap[i] * bb) * s_A[i * n + j] : 0.0;
              col += xv[i];
            }
            const double Z = warp_sum(col);
#pragma unroll
            for (int i = 0; i < NU; ++i) tacc[i] += xv[i] / Z;
            if (on) s_x[j] = bb;
            __syncwarp();
            double acc = 0.0;
#pragma unroll
            for (int c = 0; c < NU; ++c)
              if (NN > 0 || c < n) acc = fma(s_A[(on ? j : 0) * n + c], s_x[c], acc);
            beta = on ? acc : 0.0;                             // :136
            __syncwarp();
          }
        }
      }
    }
    // per-warp partial rows (one row per persistent warp and per n)
    // (accumulated, not assigned: a state count may be split into several launches by caption length)
    if (on) {
      double* pi_out = a.part_init + ((size_t)gw * (kNMax + 1) + n) * kNMax + j;
      *pi_out = LOG ? lse2(*pi_out, init_acc) : *pi_out + init_acc;
      double* pt = a.part_trans + ((size_t)gw * (kNMax + 1) + n) * (kNMax * kNMax);
#pragma unroll
      for (int i = 0; i < NU; ++i)
        if (i < n) pt[i * n + j] = LOG ? lse2(pt[i * n + j], tacc[i]) : pt[i * n + j] + tacc[i];
    }
  }
}

// exact logsumexp_i(w[i * stride] + x[i]): the out-of-line fallback of the packed kernel's weighted form
__device__ __noinline__ double lse_exact(const double* w, int stride, const double* x, int nn) {
  double m = -INFINITY;
  for (int i = 0; i < nn; ++i) m = fmax(m, w[i * stride] + x[i]);
  if (!(fabs(m) < INFINITY)) m = 0.0;
  double sm = 0.0;
  for (int i = 0; i < nn; ++i) sm += exp(w[i * stride] + x[i] - m);
  return log(sm) + m;
}

// ------------------------------------------------------------------------------------------
// Packed form of the discrete-observation E-step for n <= 8: a warp owns G = 32 / NN caption
// pairs at once (lane = (sub-pair, state)), so all lanes carry a state instead of n of 32.
// The alpha lattice of the warp's G pairs lives in a per-warp GLOBAL scratch slab laid out
// [t][lane] (one coalesced 256-byte row per step, L2-resident); shared memory only holds the
// 32-double exchange rows through which the states of a pair see each other, so residency is
// bounded by registers, not by the longest caption.  The backward sweep re-reads each lane's own
// alpha two steps ahead of its use.  Sums over the states of a pair are segment sums through an
// exchange row.  Sub-pairs of a warp run in lock-step up to the longest caption of the group
// (pairs are sorted by T, so the group is nearly uniform).  Same arithmetic per pair as
// hmm_estep_kernel.
// ------------------------------------------------------------------------------------------
template <bool LOG, int NN>
constexpr int hmm_packed_warp_doubles() {
  return 4 * 32 + (LOG ? (32 / NN) * NN * NN : 0) + (32 / NN) * (NN * NN + NN);
}

template <bool LOG, int NN>
__global__ void __launch_bounds__(128) hmm_estep_packed_kernel(const HmmArgs a, double* scratch) {
  constexpr int G = 32 / NN;
  const int n = NN, Vf = a.Vf;
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  const int sub = lane / NN, j = lane - sub * NN;
  const bool lane_on = sub < G;
  extern __shared__ double smem[];
  constexpr int per_warp = hmm_packed_warp_doubles<LOG, NN>();
  double* s_w = smem + (size_t)wic * per_warp;
  double* s_cur = s_w;                                        // [2][32] alpha exchange (double-buffered)
  double* s_red = s_w + 64;                                   // [32] segment sums
  double* s_xr = s_w + 96;                                    // [32] beta * b exchange
  double* s_E = s_w + 128 + (lane_on ? sub : 0) * NN * NN;    // [NN][NN] per sub-pair (LOG)
  double* s_fin = s_w + 128 + (LOG ? G * NN * NN : 0);        // [G][NN*NN + NN]
  const int seg0 = (lane_on ? sub : 0) * NN;
  auto seg_sum = [&](double v) -> double {
    __syncwarp();
    s_red[lane] = v;
    __syncwarp();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < NN; ++i) t += s_red[seg0 + i];
    return t;
  };
  // scipy.special.logsumexp over the states of the pair (max, sum of exp(v - max) in state order, log);
  // every lane exponentiates its OWN term once and the terms are summed through an exchange row
  auto seg_lse = [&](double v) -> double {
    __syncwarp();
    s_red[lane] = v;
    __syncwarp();
    double m = -INFINITY;
#pragma unroll
    for (int i = 0; i < NN; ++i) m = fmax(m, s_red[seg0 + i]);
    if (!(fabs(m) < INFINITY)) m = 0.0;
    s_cur[lane] = exp(v - m);
    __syncwarp();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < NN; ++i) t += s_cur[seg0 + i];
    return log(t) + m;
  };
  // logsumexp_i(w_log[i] + x[i]) for a fixed weight vector (a column / row of log A) and the pair's state
  // vector x (own entry x_own, all entries in row[seg0..]): with mg = max_i x[i] the sum is
  // sum_i exp(w_log[i]) * exp(x[i] - mg) -- ONE exp per lane (its own entry, shared through `xch`) instead
  // of one per (i, j).  If the weighted sum underflows (the dominant x[i] carries a zero / tiny weight)
  // the exact per-term form is evaluated instead.  `w_log(i)` reads log A from global memory.
  auto lse_weighted = [&](const double (&w_lin)[NN], const double* w_log, int w_stride, const double* row,
                          double x_own, double* xch, bool act) -> double {
    double mg = -INFINITY;
#pragma unroll
    for (int i = 0; i < NN; ++i) mg = fmax(mg, row[seg0 + i]);
    if (!(fabs(mg) < INFINITY)) mg = 0.0;
    xch[lane] = act ? exp(x_own - mg) : 0.0;
    __syncwarp();
    double S = 0.0;
#pragma unroll
    for (int i = 0; i < NN; ++i) S = fma(w_lin[i], xch[seg0 + i], S);
    if (S > 1e-280 || !act) return log(S) + mg;
    return lse_exact(w_log, w_stride, row + seg0, NN);
  };
  double acol[NN], arow[NN];                       // A[i][j], A[j][c]  (LOG: exp of the log table)
#pragma unroll
  for (int i = 0; i < NN; ++i) {
    acol[i] = LOG ? exp(a.trans[i * n + j]) : a.trans[i * n + j];
    arow[i] = LOG ? exp(a.trans[j * n + i]) : a.trans[j * n + i];
  }
  const double* lcol = a.trans + j;        // log A[i][j] = lcol[i * n]   (LOG, rare paths)
  const double* lrow = a.trans + j * n;    // log A[j][c] = lrow[c]
  const double pi_j = a.init[j];
  const double ident = LOG ? -INFINITY : 0.0;
  double init_acc = ident;
  double tacc[NN];
#pragma unroll
  for (int i = 0; i < NN; ++i) tacc[i] = ident;

  const int gw = blockIdx.x * a.warps_per_cta + wic;
  double* g_row = scratch + (size_t)gw * a.Tmax * 32;        // [Tmax][32]
  double* g_al = g_row + lane;
  for (int64_t base = a.lo + (int64_t)gw * G; base < a.hi; base += (int64_t)a.total_warps * G) {
    const int64_t pair = base + sub;
    const bool on = lane_on && pair < a.hi;
    int f0 = 0, T = 0;
    if (on) {
      f0 = a.src_off[pair];
      T = a.src_off[pair + 1] - f0;
    }
    const int Tw = __reduce_max_sync(0xffffffffu, T);
    const int32_t* f = a.src + f0;
    const int ej = on ? a.tgt[a.tgt_off[pair] + j] : 0;
    const double* orow = a.obs + (size_t)ej * Vf;
    const int64_t slot0 = on ? a.slot_off[pair] : 0;
    auto emis = [&](int t) -> double {
      if (LOG && a.emis) return a.emis[slot0 + (int64_t)t * NN + j];      // continuous (segment) model
      double b = orow[f[t]];
      return (b != b) ? 0.0 : b;           // absent pair: 0 in both classes (:122 / :161)
    };
    auto own_alpha = [&](int t) -> double {    // this lane's alpha_t, 0 outside the pair's lattice
      return (on && t >= 0 && t < T) ? g_al[(size_t)t * 32] : 0.0;
    };
    // ------------------------------------------------------------ forward
    double al = ident;
    if (on) {
      const double b0 = (LOG && a.emis) ? a.emis[slot0 + j] : orow[f[0]];
      al = LOG ? pi_j + b0 : pi_j * ((b0 != b0) ? 0.0 : b0);   // :158 / :114-118
      g_al[0] = al;
    }
    __syncwarp();
    s_cur[lane] = al;
    // emissions are fetched one step ahead of their use: the gather never sits on the recursion's chain
    double b_next = (on && 1 < T) ? emis(1) : 0.0;
    for (int t = 0; t + 1 < Tw; ++t) {
      __syncwarp();
      const double b = b_next;
      b_next = (on && t + 2 < T) ? emis(t + 2) : 0.0;
      const bool step = on && t + 1 < T;
      if (LOG) {
        const double nv = lse_weighted(acol, lcol, n, s_cur + (t & 1) * 32, al, s_xr, step) + b;   // :165
        if (step) al = nv;
      } else if (step) {
        const double* at = s_cur + (t & 1) * 32 + seg0;
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < NN; ++i) acc = fma(acol[i], at[i], acc);
        al = acc * b;                                          // :123
      }
      if (step) g_al[(size_t)(t + 1) * 32] = al;
      s_cur[((t + 1) & 1) * 32 + lane] = al;      // past the pair's end: alpha_{T-1} stays
    }
    {
      const double last = on ? al : ident;                          // alpha_{T-1}
      const double ll = LOG ? seg_lse(last) : log(seg_sum(last));   // :312 / :244-245
      if (on && j == 0) a.pair_ll[pair] = ll;
    }
    __syncwarp();        // the warp's alpha rows are visible to all its lanes from here
    // ------------------------------------------------------------ backward + counts
    if (LOG) {
      // transition counts from the LAST t only (:204-229)
      if (on && T >= 2) {
        const double bl = emis(T - 1);
        const double* a2 = g_row + (size_t)(T - 2) * 32 + seg0;
#pragma unroll
        for (int i = 0; i < NN; ++i) s_E[i * NN + j] = a2[i] + lcol[i * n] + bl;   // beta_{T-1} = 0
      }
      __syncwarp();
      if (on && T >= 2) {
#pragma unroll
        for (int i = 0; i < NN; ++i) {
          const int dlt = j - i;
          double m = -INFINITY;
#pragma unroll
          for (int r = 0; r < NN; ++r) {
            const int c = r + dlt;
            if (c >= 0 && c < NN) m = fmax(m, s_E[r * NN + c]);
          }
          if (!(fabs(m) < INFINITY)) m = 0.0;
          double sm = 0.0;
#pragma unroll
          for (int r = 0; r < NN; ++r) {
            const int c = r + dlt;
            if (c >= 0 && c < NN) sm += exp(s_E[r * NN + c] - m);
          }
          tacc[i] = lse2(tacc[i], log(sm) + m);
        }
      }
      __syncwarp();
      double beta = on ? 0.0 : -INFINITY;
      // running logsumexp over t of v (ic) and of tot (nrm) as (max, scaled sum): one exp per step each
      double ic_m = -INFINITY, ic_s = 0.0, nrm_m = -INFINITY, nrm_s = 0.0;
      auto lse_push = [](double& m, double& sacc, double v) {
        if (v == -INFINITY) return;
        const double d = v - m;
        const double e = exp(-fabs(d));
        if (d > 0.0) {
          sacc = sacc * e + 1.0;
          m = v;
        } else {
          sacc += e;
        }
      };
      double cur = own_alpha(Tw - 1), nx1 = own_alpha(Tw - 2), nx2 = own_alpha(Tw - 3);
      double e_next = (on && Tw - 1 < T && Tw - 1 > 0) ? emis(Tw - 1) : 0.0;
      for (int t = Tw - 1; t >= 0; --t) {
        const bool act = on && t < T;
        const double v = act ? cur + beta : -INFINITY;
        cur = nx1;
        nx1 = nx2;
        nx2 = own_alpha(t - 3);
        const double e_t = e_next;                       // emission of step t, fetched one step earlier
        e_next = (on && t - 1 < T && t - 1 > 0) ? emis(t - 1) : 0.0;
        const double tot = seg_lse(v);
        if (act) {
          lse_push(ic_m, ic_s, v);
          lse_push(nrm_m, nrm_s, tot);
          g_al[(size_t)t * 32] = v;
        }
        if (t > 0) {
          const double x = act ? beta + e_t : -INFINITY;
          s_xr[lane] = x;
          __syncwarp();
          const double nb = lse_weighted(arow, lrow, 1, s_xr, x, s_red, act);   // :182
          if (act) beta = nb;
        }
      }
      const double ic = ic_s > 0.0 ? log(ic_s) + ic_m : -INFINITY;
      const double nrm = nrm_s > 0.0 ? log(nrm_s) + nrm_m : -INFINITY;
      if (on) {
        init_acc = lse2(init_acc, ic);                         // :192-194
        for (int t = 0; t < T; ++t) a.post[slot0 + (int64_t)t * NN + j] = g_al[(size_t)t * 32] - nrm;   // :244-246
      }
      __syncwarp();
    } else {
      double beta = on ? 1.0 : 0.0;
      double cur = own_alpha(Tw - 1), nx1 = own_alpha(Tw - 2), nx2 = own_alpha(Tw - 3);
      double e_next = (on && Tw - 1 < T && Tw - 1 > 0) ? emis(Tw - 1) : 0.0;
      for (int t = Tw - 1; t >= 0; --t) {
        const bool act = on && t < T;
        const double alv = cur, prev = nx1;      // alpha_t[j], alpha_{t-1}[j] (0 outside the lattice)
        cur = nx1;
        nx1 = nx2;
        nx2 = own_alpha(t - 3);
        const double e_t = e_next;                       // emission of step t, fetched one step earlier
        e_next = (on && t - 1 < T && t - 1 > 0) ? emis(t - 1) : 0.0;
        const double g = alv * beta;
        const double Gs = seg_sum(g);
        if (act) {
          const double gam = g / Gs;                           // :147-148, :193
          init_acc += gam;
          a.post[slot0 + (int64_t)t * NN + j] = gam;
        }
        if (t > 0) {
          const double bb = act ? beta * e_t : 0.0;
          s_cur[lane] = prev;
          s_xr[lane] = bb;
          __syncwarp();
          // xi_{t-1}[i][j] = (alpha_{t-1}[i] * bb[j]) * A[i][j], normalised over (i, j)  (:161-162)
          const double* ap = s_cur + seg0;
          double xv[NN];
          double col = 0.0;
#pragma unroll
          for (int i = 0; i < NN; ++i) {
            xv[i] = act ? (ap[i] * bb) * acol[i] : 0.0;
            col += xv[i];
          }
          double acc = 0.0;
#pragma unroll
          for (int c = 0; c < NN; ++c) acc = fma(arow[c], s_xr[seg0 + c], acc);
          const double Z = seg_sum(col);
          if (act) {
#pragma unroll
            for (int i = 0; i < NN; ++i) tacc[i] += xv[i] / Z;
            beta = acc;                                        // :136
          }
        }
      }
      __syncwarp();
    }
  }
  // combine the G sub-pair lanes of every state in fixed order, then the per-warp partial rows
  __syncwarp();
  if (lane_on) {
    double* fin = s_fin + sub * (NN * NN + NN);
    fin[NN * NN + j] = init_acc;
#pragma unroll
    for (int i = 0; i < NN; ++i) fin[i * NN + j] = tacc[i];
  }
  __syncwarp();
  if (lane < NN) {
    double ia = ident;
    double ta[NN];
#pragma unroll
    for (int i = 0; i < NN; ++i) ta[i] = ident;
    for (int g = 0; g < G; ++g) {
      const double* fin = s_fin + g * (NN * NN + NN);
      ia = LOG ? lse2(ia, fin[NN * NN + lane]) : ia + fin[NN * NN + lane];
#pragma unroll
      for (int i = 0; i < NN; ++i) ta[i] = LOG ? lse2(ta[i], fin[i * NN + lane]) : ta[i] + fin[i * NN + lane];
    }
    double* pi_out = a.part_init + ((size_t)gw * (kNMax + 1) + n) * kNMax + lane;
    *pi_out = LOG ? lse2(*pi_out, ia) : *pi_out + ia;
    double* pt = a.part_trans + ((size_t)gw * (kNMax + 1) + n) * (kNMax * kNMax);
#pragma unroll
    for (int i = 0; i < NN; ++i) pt[i * n + lane] = LOG ? lse2(pt[i * n + lane], ta[i]) : pt[i * n + lane] + ta[i];
  }
}

// ---------------------------------------------------------------- reductions
// one warp per (concept, phone) table entry, over its postings
// Entries with more than `big` postings (a frequent concept x a frequent phone on Zipf-distributed
// corpora) would serialise on one warp: they are only listed here (big_list / big_count) and reduced
// by the whole grid in hmm_postings_big_kernel.
template <bool LOG>
__global__ void hmm_postings_kernel(const double* __restrict__ post, const int64_t* __restrict__ idx,
                                    const int64_t* __restrict__ off, int64_t entries,
                                    double* __restrict__ out, int64_t big, int64_t* __restrict__ big_list,
                                    unsigned long long* __restrict__ big_count) {
  const int64_t e = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (e >= entries) return;
  const int64_t lo = off[e], hi = off[e + 1];
  if (hi - lo > big) {
    if (lane == 0) big_list[atomicAdd(big_count, 1ull)] = e;
    return;
  }
  if (LOG) {
    double m = -INFINITY;
    for (int64_t q = lo + lane; q < hi; q += 32) m = fmax(m, post[idx[q]]);
    m = warp_max(m);
    if (m == -INFINITY) { if (lane == 0) out[e] = -INFINITY; return; }
    double s = 0.0;
    for (int64_t q = lo + lane; q < hi; q += 32) s += exp(post[idx[q]] - m);
    s = warp_sum(s);
    if (lane == 0) out[e] = log(s) + m;
  } else {
    double s = 0.0;
    for (int64_t q = lo + lane; q < hi; q += 32) s += post[idx[q]];
    s = warp_sum(s);
    if (lane == 0) out[e] = s;
  }
}

// ---------------------------------------------------------------- M-step
struct HmmLens { int lens[kNMax + 1]; int n; };

// Reduction of the per-warp partial rows of the initial / transition counts, restricted to the state
// counts the corpus actually has: one CTA per (state count m, table), thread = (row lane, element of
// the m or m*m block); every row lane walks its rows in order, lanes combine in fixed order.
// LOG: max over all rows first, then sum of exp(v - max) -- the two-pass logsumexp of the row loop
// it replaces.  Everything outside the used blocks is the identity (written by hmm_fill_kernel).
__global__ void hmm_fill_kernel(double* __restrict__ out, int64_t elems, double v) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < elems) out[e] = v;
}

constexpr int kReduceThreads = 1024;

template <bool LOG>
__global__ void __launch_bounds__(kReduceThreads) hmm_reduce_blocks_kernel(
    const HmmLens la, const double* __restrict__ part_init, const double* __restrict__ part_trans, int rows,
    double* __restrict__ out_init, double* __restrict__ out_trans) {
  const int m = la.lens[blockIdx.x >> 1];
  const bool is_trans = blockIdx.x & 1;
  const int E = is_trans ? m * m : m;
  const size_t stride = is_trans ? (size_t)(kNMax + 1) * kNMax * kNMax : (size_t)(kNMax + 1) * kNMax;
  const size_t off = is_trans ? (size_t)m * kNMax * kNMax : (size_t)m * kNMax;
  const double* src = (is_trans ? part_trans : part_init) + off;
  double* dst = (is_trans ? out_trans : out_init) + off;
  const int lanes = kReduceThreads / E;                  // row lanes per element (>= 4 for m <= 16)
  const int e = threadIdx.x % E, rl = threadIdx.x / E;
  const bool live = rl < lanes;
  __shared__ double s_part[kReduceThreads];
  __shared__ double s_max[kNMax * kNMax];
  double mx = 0.0;
  if (LOG) {
    double v = -INFINITY;
    if (live)
      for (int r = rl; r < rows; r += lanes) v = fmax(v, src[(size_t)r * stride + e]);
    s_part[threadIdx.x] = v;
    __syncthreads();
    if (threadIdx.x < E) {
      double t = -INFINITY;
      for (int l = 0; l < lanes; ++l) t = fmax(t, s_part[l * E + threadIdx.x]);
      s_max[threadIdx.x] = t;
    }
    __syncthreads();
    mx = s_max[e];
  }
  double acc = 0.0;
  if (live && !(LOG && mx == -INFINITY))
    for (int r = rl; r < rows; r += lanes) {
      const double v = src[(size_t)r * stride + e];
      acc += LOG ? exp(v - mx) : v;
    }
  s_part[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < E) {
    double t = 0.0;
    for (int l = 0; l < lanes; ++l) t += s_part[l * E + threadIdx.x];
    if (LOG) dst[threadIdx.x] = (s_max[threadIdx.x] == -INFINITY) ? -INFINITY : log(t) + s_max[threadIdx.x];
    else dst[threadIdx.x] = t;
  }
}

template <bool LOG>
__global__ void hmm_mstep_init_trans_kernel(HmmLens la, const double* __restrict__ initC,
                                            const double* __restrict__ transC, double* __restrict__ accI,
                                            double* __restrict__ accT, double* __restrict__ init,
                                            double* __restrict__ trans) {
  const int m = la.lens[blockIdx.x];
  __shared__ double sC[kNMax * kNMax], sI[kNMax], sJ[2 * kNMax];
  const int tid = threadIdx.x;
  const double* ic = initC + (size_t)m * kNMax;
  const double* tc = transC + (size_t)m * kNMax * kNMax;
  double* io = init + (size_t)m * kNMax;
  double* to = trans + (size_t)m * kNMax * kNMax;
  if (LOG) {
    double* ai = accI + (size_t)m * kNMax;
    double* at = accT + (size_t)m * kNMax * kNMax;
    for (int e = tid; e < m * m; e += blockDim.x) { at[e] = lse2(at[e], tc[e]); sC[e] = at[e]; }
    for (int e = tid; e < m; e += blockDim.x) { ai[e] = lse2(ai[e], ic[e]); sI[e] = ai[e]; }
    __syncthreads();
    if (tid < m) {   // row r = tid: trans[r] -= LSE_s(acc[r][s])   (:364-369)
      double mx = -INFINITY;
      for (int c = 0; c < m; ++c) mx = fmax(mx, sC[tid * m + c]);
      if (!(fabs(mx) < INFINITY)) mx = 0.0;
      double s = 0.0;
      for (int c = 0; c < m; ++c) s += exp(sC[tid * m + c] - mx);
      const double nf = log(s) + mx;
      for (int c = 0; c < m; ++c) to[tid * m + c] = sC[tid * m + c] - nf;
    }
    if (tid == 0) {  // init -= LSE(acc)   (:355-360)
      double mx = -INFINITY;
      for (int c = 0; c < m; ++c) mx = fmax(mx, sI[c]);
      if (!(fabs(mx) < INFINITY)) mx = 0.0;
      double s = 0.0;
      for (int c = 0; c < m; ++c) s += exp(sI[c] - mx);
      const double nf = log(s) + mx;
      for (int c = 0; c < m; ++c) io[c] = sI[c] - nf;
    }
  } else {
    for (int e = tid; e < m * m; e += blockDim.x) sC[e] = tc[e];
    __syncthreads();
    // Toeplitz pooling (:165-177), linear so applied to the summed counts
    for (int dlt = tid; dlt < 2 * m - 1; dlt += blockDim.x) {
      int offd = dlt - (m - 1);
      double s = 0.0;
      for (int r = 0; r < m; ++r) {
        int c = r + offd;
        if (c >= 0 && c < m) s += sC[r * m + c];
      }
      sJ[dlt] = s;
    }
    __syncthreads();
    if (tid < m) {
      double tot = 0.0;
      for (int c = 0; c < m; ++c) tot += sJ[c - tid + m - 1];
      if (tot != 0.0)
        for (int c = 0; c < m; ++c) to[tid * m + c] = sJ[c - tid + m - 1] / tot;   // :279-286
    }
    if (tid == 0) {
      double tot = 0.0;
      for (int c = 0; c < m; ++c) tot += ic[c];
      for (int c = 0; c < m; ++c) io[c] = ic[c] / tot;                             // :276-277
    }
  }
}

// one warp per concept row: obs[tw][fw] = c / sum_fw c  (present entries only)
template <bool LOG>
__global__ void hmm_mstep_obs_kernel(const double* __restrict__ obsC, double* __restrict__ accO, int Vt,
                                     int Vf, double* __restrict__ obs) {
  const int tw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (tw >= Vt) return;
  const size_t base = (size_t)tw * Vf;
  if (LOG) {
    double m = -INFINITY;
    for (int fw = lane; fw < Vf; fw += 32) {
      double o = obs[base + fw];
      if (o == o) {
        double v = lse2(accO[base + fw], obsC[base + fw]);
        accO[base + fw] = v;
        m = fmax(m, v);
      }
    }
    m = warp_max(m);
    if (!(fabs(m) < INFINITY)) m = 0.0;
    double s = 0.0;
    for (int fw = lane; fw < Vf; fw += 32)
      if (obs[base + fw] == obs[base + fw]) s += exp(accO[base + fw] - m);
    s = warp_sum(s);
    const double nf = log(s) + m;
    for (int fw = lane; fw < Vf; fw += 32)
      if (obs[base + fw] == obs[base + fw]) obs[base + fw] = accO[base + fw] - nf;   // :388-389
  } else {
    double s = 0.0;
    for (int fw = lane; fw < Vf; fw += 32)
      if (obs[base + fw] == obs[base + fw]) s += obsC[base + fw];
    s = warp_sum(s);
    for (int fw = lane; fw < Vf; fw += 32)
      if (obs[base + fw] == obs[base + fw]) obs[base + fw] = obsC[base + fw] / s;    // :288-296
  }
}

// ---------------------------------------------------------------- postings: large entries
constexpr int kPostSplit = 128;     // slices (= CTAs) per large entry
constexpr int kPostThreads = 256;

// CTA s reduces slice s of every listed entry: slices are cut by position, threads stride the slice,
// lanes / warps combine in fixed order -> the value does not depend on the order of the list.
// part[b][s] = {max, sum exp(v - max)} (LOG) or {0, sum}.
template <bool LOG>
__global__ void __launch_bounds__(kPostThreads) hmm_postings_big_kernel(
    const double* __restrict__ post, const int64_t* __restrict__ idx, const int64_t* __restrict__ off,
    const int64_t* __restrict__ big_list, const unsigned long long* __restrict__ big_count,
    double* __restrict__ part) {
  __shared__ double s_w[kPostThreads / 32];
  __shared__ double s_b;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t nbig = (int64_t)*big_count;
  auto cta_reduce = [&](double v, bool is_max) -> double {
    v = is_max ? warp_max(v) : warp_sum(v);
    __syncthreads();
    if (lane == 0) s_w[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = s_w[0];
      for (int w = 1; w < kPostThreads / 32; ++w) t = is_max ? fmax(t, s_w[w]) : t + s_w[w];
      s_b = t;
    }
    __syncthreads();
    return s_b;
  };
  for (int64_t b = 0; b < nbig; ++b) {
    const int64_t e = big_list[b];
    const int64_t lo = off[e], len = off[e + 1] - lo;
    const int64_t q0 = lo + len * blockIdx.x / kPostSplit, q1 = lo + len * (blockIdx.x + 1) / kPostSplit;
    double m = 0.0;
    if (LOG) {
      double v = -INFINITY;
      for (int64_t q = q0 + threadIdx.x; q < q1; q += kPostThreads) v = fmax(v, post[idx[q]]);
      m = cta_reduce(v, true);
    }
    double sacc = 0.0;
    if (!(LOG && m == -INFINITY))
      for (int64_t q = q0 + threadIdx.x; q < q1; q += kPostThreads) {
        const double v = post[idx[q]];
        sacc += LOG ? exp(v - m) : v;
      }
    sacc = cta_reduce(sacc, false);
    if (threadIdx.x == 0) {
      part[(b * kPostSplit + blockIdx.x) * 2] = m;
      part[(b * kPostSplit + blockIdx.x) * 2 + 1] = sacc;
    }
  }
}

// one thread per listed entry: slices combined in slice order
template <bool LOG>
__global__ void hmm_postings_big_combine_kernel(const int64_t* __restrict__ big_list,
                                                const unsigned long long* __restrict__ big_count,
                                                const double* __restrict__ part, double* __restrict__ out) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= (int64_t)*big_count) return;
  const double* pb = part + b * kPostSplit * 2;
  if (LOG) {
    double M = -INFINITY;
    for (int s2 = 0; s2 < kPostSplit; ++s2) M = fmax(M, pb[2 * s2]);
    if (M == -INFINITY) { out[big_list[b]] = -INFINITY; return; }
    double S = 0.0;
    for (int s2 = 0; s2 < kPostSplit; ++s2)
      if (pb[2 * s2] != -INFINITY) S += pb[2 * s2 + 1] * exp(pb[2 * s2] - M);
    out[big_list[b]] = log(S) + M;
  } else {
    double S = 0.0;
    for (int s2 = 0; s2 < kPostSplit; ++s2) S += pb[2 * s2 + 1];
    out[big_list[b]] = S;
  }
}

// ---------------------------------------------------------------- align
struct HmmAlignArgs {
  const int32_t* tgt_off;
  const int32_t* tgt;
  const int32_t* src_off;
  const int32_t* src;
  const double* init;
  const double* trans;
  const double* obs;
  int32_t* alignment;
  double* align_probs;
  const int64_t* ap_off;
  const double* emis;
  const int64_t* slot_off;
  int64_t n_pairs;
  int64_t lo, hi;          // pair range of this launch
  int Vf, Tmax, warps_per_cta;
  double unk;
};

__device__ __forceinline__ bool np_greater_h(double cand, double best) {
  return (cand > best) || (cand != cand && best == best);
}

template <bool LOG>
__global__ void __launch_bounds__(256) hmm_align_kernel(const HmmAlignArgs a) {
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  const int64_t pair = a.lo + (int64_t)blockIdx.x * a.warps_per_cta + wic;
  if (wic >= a.warps_per_cta || pair >= a.hi) return;
  extern __shared__ double smem[];
  const size_t per_warp = 2 * kNMax * sizeof(double) + (size_t)a.Tmax * kNMax;
  unsigned char* base = reinterpret_cast<unsigned char*>(smem) + (size_t)wic * ((per_warp + 7) / 8 * 8);
  double* s_sc = reinterpret_cast<double*>(base);           // [2][NMAX]
  unsigned char* s_bp = base + 2 * kNMax * sizeof(double);  // [Tmax][NMAX]
  const int e0 = a.tgt_off[pair];
  const int n = a.tgt_off[pair + 1] - e0;
  const int f0 = a.src_off[pair];
  const int T = a.src_off[pair + 1] - f0;
  const int32_t* f = a.src + f0;
  const int j = lane;
  const bool on = j < n;
  const double* A = a.trans + (size_t)n * MWD_TRANS_STRIDE;
  const double* pi = a.init + (size_t)n * MWD_INIT_STRIDE;
  const double* orow = a.obs + (size_t)(on ? a.tgt[e0 + j] : 0) * a.Vf;
  double* ap = a.align_probs ? a.align_probs + a.ap_off[pair] : nullptr;
  const int64_t slot0 = a.emis ? a.slot_off[pair] : 0;
  double sc = 0.0;
  if (on) {
    double b0 = a.emis ? a.emis[slot0 + j] : orow[f[0]];
    sc = LOG ? pi[j] + b0 : pi[j] * b0;                       // :307 / :402
    s_sc[j] = sc;
  }
  __syncwarp();
  for (int t = 1; t < T; ++t) {
    const double* prev = s_sc + ((t - 1) & 1) * kNMax;
    double* next = s_sc + (t & 1) * kNMax;
    if (on) {
      double b = a.emis ? a.emis[slot0 + (int64_t)t * n + j] : orow[f[t]];
      if (b != b) b = a.unk;                                  // :311 / :407
      double best = LOG ? __dadd_rn(__dadd_rn(prev[0], A[j]), b) : __dmul_rn(__dmul_rn(prev[0], A[j]), b);
      int arg = 0;
      for (int i = 1; i < n; ++i) {
        double cand = LOG ? __dadd_rn(__dadd_rn(prev[i], A[i * n + j]), b)
                          : __dmul_rn(__dmul_rn(prev[i], A[i * n + j]), b);
        if (np_greater_h(cand, best)) { best = cand; arg = i; }
      }
      s_bp[t * kNMax + j] = (unsigned char)arg;
      sc = best;
      next[j] = sc;
    }
    __syncwarp();
    if (ap && on) {
      if (LOG) {
        ap[(size_t)(t - 1) * n + j] = sc;                     // :414
      } else {
        double tot = 0.0;
        for (int i = 0; i < n; ++i) tot += next[i];
        ap[(size_t)(t - 1) * n + j] = sc / tot;               // :316
      }
    }
  }
  if (lane == 0) {
    const double* fin = s_sc + ((T - 1) & 1) * kNMax;
    double best = fin[0];
    int cur = 0;
    for (int i = 1; i < n; ++i)
      if (np_greater_h(fin[i], best)) { best = fin[i]; cur = i; }
    a.alignment[f0 + T - 1] = cur;
    for (int t = T - 1; t > 0; --t) {
      cur = s_bp[t * kNMax + cur];
      a.alignment[f0 + t - 1] = cur;
    }
  }
}

// Packed Viterbi for n <= 8: a warp decodes G = 32 / NN pairs at once (lane = (sub-pair, state)), scores
// exchanged through a double-buffered 32-double row, back-pointers one byte per (t, lane); the G
// back-traces run in parallel, one lane per pair.  Same products / comparisons per pair, in the same
// order, as hmm_align_kernel (first index wins ties, NaN handling of np.argmax).
template <bool LOG, int NN>
__global__ void __launch_bounds__(256) hmm_align_packed_kernel(const HmmAlignArgs a) {
  constexpr int G = 32 / NN;
  const int n = NN;
  const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
  const int sub = lane / NN, j = lane - sub * NN;
  const bool lane_on = sub < G;
  const int seg0 = (lane_on ? sub : 0) * NN;
  extern __shared__ double smem[];
  const size_t per_warp = 2 * 32 * sizeof(double) + (size_t)a.Tmax * 32;
  unsigned char* wbase = reinterpret_cast<unsigned char*>(smem) + (size_t)wic * per_warp;
  double* s_sc = reinterpret_cast<double*>(wbase);           // [2][32]
  unsigned char* s_bp = wbase + 2 * 32 * sizeof(double);     // [Tmax][32]
  const int64_t pair = a.lo + ((int64_t)blockIdx.x * a.warps_per_cta + wic) * G + sub;
  const bool on = lane_on && pair < a.hi;
  int e0 = 0, f0 = 0, T = 0;
  if (on) {
    e0 = a.tgt_off[pair];
    f0 = a.src_off[pair];
    T = a.src_off[pair + 1] - f0;
  }
  const int Tw = __reduce_max_sync(0xffffffffu, T);
  if (Tw == 0) return;
  const int32_t* f = a.src + f0;
  const double* A = a.trans + (size_t)n * MWD_TRANS_STRIDE;
  double acol[NN];
#pragma unroll
  for (int i = 0; i < NN; ++i) acol[i] = A[i * n + j];
  const double* orow = a.obs + (size_t)(on ? a.tgt[e0 + j] : 0) * a.Vf;
  double* ap = (a.align_probs && on) ? a.align_probs + a.ap_off[pair] : nullptr;
  const int64_t slot0 = (a.emis && on) ? a.slot_off[pair] : 0;
  auto emis = [&](int t) -> double {
    double b = a.emis ? a.emis[slot0 + (int64_t)t * n + j] : orow[f[t]];
    return (b != b) ? a.unk : b;                              // :311 / :407
  };
  double sc = 0.0;
  if (on) {
    const double b0 = a.emis ? a.emis[slot0 + j] : orow[f[0]];
    const double pj = a.init[(size_t)n * MWD_INIT_STRIDE + j];
    sc = LOG ? pj + b0 : pj * b0;                             // :307 / :402
  }
  s_sc[lane] = sc;
  double b_next = (on && 1 < T) ? emis(1) : 0.0;
  for (int t = 1; t < Tw; ++t) {
    __syncwarp();
    const double* prev = s_sc + ((t - 1) & 1) * 32 + seg0;
    const bool act = on && t < T;
    const double b = b_next;
    b_next = (on && t + 1 < T) ? emis(t + 1) : 0.0;
    if (act) {
      double best = LOG ? __dadd_rn(__dadd_rn(prev[0], acol[0]), b) : __dmul_rn(__dmul_rn(prev[0], acol[0]), b);
      int arg = 0;
#pragma unroll
      for (int i = 1; i < NN; ++i) {
        const double cand = LOG ? __dadd_rn(__dadd_rn(prev[i], acol[i]), b)
                                : __dmul_rn(__dmul_rn(prev[i], acol[i]), b);
        if (np_greater_h(cand, best)) { best = cand; arg = i; }
      }
      s_bp[t * 32 + lane] = (unsigned char)arg;
      sc = best;
    }
    double* next = s_sc + (t & 1) * 32;
    next[lane] = sc;                  // finished pairs keep re-posting their final scores
    if (a.align_probs) {              // warp-uniform
      if (LOG) {
        if (act) ap[(size_t)(t - 1) * n + j] = sc;            // :414
      } else {
        __syncwarp();
        if (act) {
          double tot = 0.0;
#pragma unroll
          for (int i = 0; i < NN; ++i) tot += next[seg0 + i];
          ap[(size_t)(t - 1) * n + j] = sc / tot;             // :316
        }
      }
    }
  }
  __syncwarp();
  if (on && j == 0) {
    const double* fin = s_sc + ((Tw - 1) & 1) * 32 + seg0;
    double best = fin[0];
    int cur = 0;
#pragma unroll
    for (int i = 1; i < NN; ++i)
      if (np_greater_h(fin[i], best)) { best = fin[i]; cur = i; }
    a.alignment[f0 + T - 1] = cur;
    for (int t = T - 1; t > 0; --t) {
      cur = s_bp[t * 32 + seg0 + cur];
      a.alignment[f0 + t - 1] = cur;
    }
  }
}

// ---------------------------------------------------------------- Gaussian segment emissions
// lnorm[w][m] = -(D/2 log 2pi + 1/2 sum_d log var)   (smt/audio_gmm_word_discoverer.py:58)
__global__ void gauss_lnorm_kernel(const double* __restrict__ var, int D, double* __restrict__ lnorm) {
  const int wm = blockIdx.x;
  double s = 0.0;
  for (int d = threadIdx.x; d < D; d += 32) s += log(var[(size_t)wm * D + d]);
  s = warp_sum(s);
  if (threadIdx.x == 0) lnorm[wm] = -(0.5 * D * log(2.0 * 3.14159265358979323846) + 0.5 * s);
}

// one warp per (pair, t): lanes over the embedding dimensions
template <typename FT>
__global__ void __launch_bounds__(256) gauss_emission_kernel(
    const int32_t* __restrict__ tgt_off, const int32_t* __restrict__ tgt, const int32_t* __restrict__ src_off,
    const int64_t* __restrict__ slot_off, int64_t n_pairs, const FT* __restrict__ emb, int D, int M,
    const double* __restrict__ lprior, const double* __restrict__ means, const double* __restrict__ var,
    const double* __restrict__ lnorm, double* __restrict__ emis, double* __restrict__ resp, int64_t n_rows,
    const int32_t* __restrict__ row_pair) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // embedding row
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  const int pair = row_pair[row];
  const int t = (int)(row - src_off[pair]);
  const int e0 = tgt_off[pair];
  const int n = tgt_off[pair + 1] - e0;
  const FT* x = emb + row * D;
  const int64_t slot = slot_off[pair] + (int64_t)t * n;
  for (int j = 0; j < n; ++j) {
    const int w = tgt[e0 + j];
    double comp[8];
    double mx = -INFINITY;
    for (int m = 0; m < M; ++m) {
      const double* mu = means + ((size_t)w * M + m) * D;
      const double* vr = var + ((size_t)w * M + m) * D;
      double q = 0.0;
      for (int d = lane; d < D; d += 32) {
        const double z = (double)x[d] - mu[d];
        q += z * z / (2.0 * vr[d]);
      }
      q = warp_sum(q);
      comp[m] = lprior[w * M + m] + lnorm[w * M + m] - q;                 // :59, :395-401
      mx = fmax(mx, comp[m]);
    }
    if (!(fabs(mx) < INFINITY)) mx = 0.0;
    double s = 0.0;
    for (int m = 0; m < M; ++m) s += exp(comp[m] - mx);
    const double lb = log(s) + mx;
    if (lane == 0) {
      emis[slot + j] = lb;
      for (int m = 0; m < M; ++m) resp[(slot + j) * M + m] = comp[m] - lb;
    }
  }
}

// one CTA per word; thread d owns dimension d; slots of the word in postings order
template <typename FT>
__global__ void __launch_bounds__(128) gauss_stats_kernel(
    const int64_t* __restrict__ word_idx, const int64_t* __restrict__ word_off,
    const int32_t* __restrict__ slot_row, const FT* __restrict__ emb, int D, int M,
    const double* __restrict__ post, const double* __restrict__ resp, double* __restrict__ stats) {
  const int w = blockIdx.x;
  const int64_t lo = word_off[w], hi = word_off[w + 1];
  const int SW = 1 + 2 * D;
  for (int m = 0; m < M; ++m) {
    double* out = stats + ((size_t)w * M + m) * SW;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
      double sx = 0.0, sxx = 0.0, sw = 0.0;
      for (int64_t q = lo; q < hi; ++q) {
        const int64_t slot = word_idx[q];
        const double wg = exp(post[slot] + resp[slot * M + m]);
        const double xv = (double)emb[(size_t)slot_row[slot] * D + d];
        sw += wg;
        sx = fma(wg, xv, sx);
        sxx = fma(wg, xv * xv, sxx);
      }
      out[1 + d] = sx;
      out[1 + D + d] = sxx;
      if (d == 0) out[0] = sw;
    }
  }
}

// Sliced form (the default): CTA (w, s) accumulates slice s of word w's postings -- slices are cut by
// position, so the hot words of a corpus (NULL is a state of EVERY utterance) spread over the grid
// instead of serialising one CTA.  Per 128-posting tile the weights exp(post + resp) are computed once,
// one per thread, and shared; thread d then walks the tile for its dimension (coalesced embedding rows).
// part[w][s][m] = {sum wgt, sum wgt*x (D), sum wgt*x^2 (D)}; gauss_stats_combine_kernel adds the slices
// in slice order.
template <typename FT>
__global__ void __launch_bounds__(128) gauss_stats_slice_kernel(
    const int64_t* __restrict__ word_idx, const int64_t* __restrict__ word_off,
    const int32_t* __restrict__ slot_row, const FT* __restrict__ emb, int D, int M,
    const double* __restrict__ post, const double* __restrict__ resp, double* __restrict__ part) {
  const int w = blockIdx.x, sl = blockIdx.y, S = gridDim.y;
  const int64_t lo0 = word_off[w], len = word_off[w + 1] - lo0;
  const int64_t lo = lo0 + len * sl / S, hi = lo0 + len * (sl + 1) / S;
  const int SW = 1 + 2 * D;
  __shared__ double s_wg[128];
  __shared__ int32_t s_row[128];
  for (int m = 0; m < M; ++m) {
    double* out = part + (((size_t)w * S + sl) * M + m) * SW;
    for (int d0 = 0; d0 < D; d0 += 128) {
      const int d = d0 + threadIdx.x;
      double sx = 0.0, sxx = 0.0, sw = 0.0;
      for (int64_t q0 = lo; q0 < hi; q0 += 128) {
        const int cnt = (int)((hi - q0 < 128) ? hi - q0 : 128);
        __syncthreads();
        if ((int)threadIdx.x < cnt) {
          const int64_t slot = word_idx[q0 + threadIdx.x];
          s_wg[threadIdx.x] = exp(post[slot] + resp[slot * M + m]);
          s_row[threadIdx.x] = slot_row[slot];
        }
        __syncthreads();
        if (d < D)
          for (int i = 0; i < cnt; ++i) {
            const double wg = s_wg[i];
            const double xv = (double)emb[(size_t)s_row[i] * D + d];
            sw += wg;
            sx = fma(wg, xv, sx);
            sxx = fma(wg, xv * xv, sxx);
          }
      }
      if (d < D) {
        out[1 + d] = sx;
        out[1 + D + d] = sxx;
        if (d == 0) out[0] = sw;
      }
    }
  }
}

__global__ void gauss_stats_combine_kernel(const double* __restrict__ part, int S, int row_len,
                                           double* __restrict__ stats) {
  // blockIdx.x = w; row_len = M * (1 + 2D)
  for (int e = threadIdx.x; e < row_len; e += blockDim.x) {
    double t = 0.0;
    for (int sl = 0; sl < S; ++sl) t += part[((size_t)blockIdx.x * S + sl) * row_len + e];
    stats[(size_t)blockIdx.x * row_len + e] = t;
  }
}

__global__ void gauss_update_kernel(int M, int D, const double* __restrict__ stats, int update_var,
                                    double* __restrict__ lprior, double* __restrict__ means,
                                    double* __restrict__ var) {
  const int w = blockIdx.x;
  const int SW = 1 + 2 * D;
  double tot = 0.0;
  for (int m = 0; m < M; ++m) tot += stats[((size_t)w * M + m) * SW];
  for (int m = 0; m < M; ++m) {
    const double* st = stats + ((size_t)w * M + m) * SW;
    const double sw = st[0];
    if (sw > 0.0) {
      for (int d = threadIdx.x; d < D; d += blockDim.x) {
        const double mu = st[1 + d] / sw;
        means[((size_t)w * M + m) * D + d] = mu;
        if (update_var) {
          double v = st[1 + D + d] / sw - mu * mu;
          var[((size_t)w * M + m) * D + d] = v < 1e-6 ? 1e-6 : v;
        }
      }
    }
    if (M > 1 && tot > 0.0 && threadIdx.x == 0) lprior[w * M + m] = log(sw / tot);
  }
}

int sum_doubles(const double* x, int64_t n, double* blk_scratch, double* out, cudaStream_t st);

}  // namespace mwd

using namespace mwd;

extern "C" int mwd_hmm_warps(void) { return hmm_warps_total(); }

namespace {
// The library's own stream-ordered pool (one per device) with an unbounded release threshold: a freed
// scratch block stays in the pool, so the next call's allocation is a pointer bump, not a cudaMalloc.
int scratch_pool(cudaMemPool_t* out) {
  static std::mutex mu;
  static cudaMemPool_t pools[64] = {};
  int dev = 0;
  MWD_CHECK_CUDA(cudaGetDevice(&dev));
  MWD_REQUIRE(dev >= 0 && dev < 64, "device ordinal %d out of range", dev);
  std::lock_guard<std::mutex> lock(mu);
  if (!pools[dev]) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    MWD_CHECK_CUDA(cudaMemPoolCreate(&pools[dev], &props));
    uint64_t keep = ~0ull;
    MWD_CHECK_CUDA(cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep));
  }
  *out = pools[dev];
  return 0;
}

struct ScratchGuard {   // stream-ordered free on every exit path
  cudaStream_t st;
  void* ptr = nullptr;
  explicit ScratchGuard(cudaStream_t s) : st(s) {}
  ~ScratchGuard() { if (ptr) cudaFreeAsync(ptr, st); }
};
}  // namespace

extern "C" int64_t mwd_hmm_counts_len(int Vt, int Vf) {
  return (int64_t)Vt * Vf + (int64_t)(kNMax + 1) * kNMax + (int64_t)(kNMax + 1) * kNMax * kNMax + 1;
}

extern "C" int mwd_hmm_estep(const mwd_hmm_problem* p, void* stream) {
  cudaStream_t st = as_stream(stream);
  const int total = hmm_warps_total();
  const char* pk_env = getenv("MWD_HMM_PACKED");
  // alpha scratch of the packed kernel: [warp][Tmax][32] doubles for the longest caption of any packed
  // launch group, stream-ordered so the pool hands the same block back on every call
  ScratchGuard guard(st);
  double* scratch = nullptr;
  {
    int t_packed = 0;
    for (int b = 0; b < p->n_buckets; ++b)
      if (p->bucket_n[b] <= 8 && p->bucket_lo[b + 1] > p->bucket_lo[b] && p->bucket_tmax[b] > t_packed)
        t_packed = p->bucket_tmax[b];
    if (t_packed > 0 && !p->alpha_out && !(pk_env && atoi(pk_env) == 0)) {
      cudaMemPool_t pool;
      if (int rc = scratch_pool(&pool)) return rc;
      MWD_CHECK_CUDA(cudaMallocFromPoolAsync(&guard.ptr, (size_t)total * t_packed * 32 * sizeof(double), pool, st));
      scratch = static_cast<double*>(guard.ptr);
    }
  }
  for (int b = 0; b < p->n_buckets; ++b) {
    const int n = p->bucket_n[b];
    const int64_t lo = p->bucket_lo[b], hi = p->bucket_lo[b + 1];
    if (hi <= lo) continue;
    MWD_REQUIRE(n >= 1 && n <= kNMax, "bucket %d: %d states outside [1,%d]", b, n, kNMax);
    const int Tmax = p->bucket_tmax[b];
    const size_t fixed = (size_t)(kNMax * kNMax + kNMax) * sizeof(double);
    const size_t per_warp = ((size_t)Tmax * n + kNMax + kNMax * kNMax) * sizeof(double);
    int wpc = (int)((220 * 1024 - fixed) / per_warp);
    if (wpc > 8) wpc = 8;
    MWD_REQUIRE(wpc >= 1, "caption of %d tokens x %d states does not fit in shared memory", Tmax, n);
    // keep the warp count (rows of the partial tables) fixed: grid * wpc == total
    while (total % wpc) --wpc;
    const int grid = total / wpc;
    HmmArgs a;
    a.tgt_off = p->tgt_off; a.tgt = p->tgt; a.src_off = p->src_off; a.src = p->src;
    a.slot_off = p->slot_off;
    a.init = p->init + (size_t)n * MWD_INIT_STRIDE;
    a.trans = p->trans + (size_t)n * MWD_TRANS_STRIDE;
    a.obs = p->obs;
    a.pair_ll = p->pair_ll; a.post = p->post;
    a.part_init = p->part_init; a.part_trans = p->part_trans;
    a.alpha_out = p->alpha_out; a.beta_out = p->beta_out;
    a.emis = p->emis;
    a.lo = lo; a.hi = hi; a.n = n; a.Vf = p->n_src_types; a.Tmax = Tmax;
    a.warps_per_cta = wpc; a.total_warps = total;
    const size_t smem = fixed + (size_t)wpc * per_warp;
    auto launch = [&](auto kern) -> int {
      MWD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<grid, wpc * 32, smem, st>>>(a);
      return 0;
    };
    int rc;
    // packed kernel (several pairs per warp): every mode but the dense alpha / beta dump of forward()
    if (n <= 8 && !p->alpha_out && !(pk_env && atoi(pk_env) == 0)) {
      const int pwpc = 4;
      MWD_REQUIRE(total % pwpc == 0, "warp count %d not a multiple of %d", total, pwpc);
      HmmArgs b2 = a;
      b2.warps_per_cta = pwpc;
      auto launch_p = [&](auto kern, int warp_doubles) -> int {
        const size_t psmem = (size_t)pwpc * warp_doubles * sizeof(double);
        MWD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
        // persistent warps: never launch more CTAs than are resident at once (rows of the partial tables
        // past the launched warps keep their identity fill)
        int occ = 0;
        MWD_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, pwpc * 32, psmem));
        MWD_REQUIRE(occ >= 1, "packed recursion kernel does not fit on an SM");
        int ctas = occ * sm_count();
        if (ctas > total / pwpc) ctas = total / pwpc;
        HmmArgs b3 = b2;
        b3.total_warps = ctas * pwpc;
        kern<<<ctas, pwpc * 32, psmem, st>>>(b3, scratch);
        return 0;
      };
      switch (n) {
#define MWD_HP(V)                                                                                              \
  case V:                                                                                                      \
    rc = p->log_domain ? launch_p(hmm_estep_packed_kernel<true, V>, hmm_packed_warp_doubles<true, V>())        \
                       : launch_p(hmm_estep_packed_kernel<false, V>, hmm_packed_warp_doubles<false, V>());     \
    break;
        MWD_HP(1) MWD_HP(2) MWD_HP(3) MWD_HP(4) MWD_HP(5) MWD_HP(6) MWD_HP(7) MWD_HP(8)
#undef MWD_HP
        default: rc = 2; break;
      }
      if (rc) return rc;
      MWD_CHECK_LAUNCH();
      continue;
    }
    switch (n <= 8 ? n : 0) {
#define MWD_HN(V)                                                                                     \
  case V:                                                                                             \
    rc = p->log_domain ? launch(hmm_estep_kernel<true, V>) : launch(hmm_estep_kernel<false, V>);      \
    break;
      MWD_HN(1) MWD_HN(2) MWD_HN(3) MWD_HN(4) MWD_HN(5) MWD_HN(6) MWD_HN(7) MWD_HN(8)
#undef MWD_HN
      default: rc = p->log_domain ? launch(hmm_estep_kernel<true, 0>) : launch(hmm_estep_kernel<false, 0>); break;
    }
    if (rc) return rc;
    MWD_CHECK_LAUNCH();
  }
  return 0;
}

extern "C" int mwd_hmm_reduce(const mwd_hmm_problem* p, const int64_t* post_idx, const int64_t* post_off,
                              double* counts, void* stream) {
  cudaStream_t st = as_stream(stream);
  const int rows = hmm_warps_total();
  const int64_t oe = (int64_t)p->n_tgt_types * p->n_src_types;
  const int64_t ie = (int64_t)(kNMax + 1) * kNMax;
  const int64_t te = (int64_t)(kNMax + 1) * kNMax * kNMax;
  const unsigned og = (unsigned)((oe + 7) / 8);
  // state counts present in this shard (launch groups are sorted by state count)
  HmmLens la;
  la.n = 0;
  for (int b = 0; b < p->n_buckets; ++b) {
    const int m = p->bucket_n[b];
    MWD_REQUIRE(m >= 1 && m <= kNMax, "bucket %d: %d states outside [1,%d]", b, m, kNMax);
    bool seen = false;
    for (int i = 0; i < la.n; ++i) seen = seen || la.lens[i] == m;
    if (!seen) la.lens[la.n++] = m;
  }
  const unsigned fg = (unsigned)((ie + te + 255) / 256);
  // observation counts over the postings index; entries past `big` postings go through the split path
  const bool want_obs = !(p->log_domain && p->emis) && oe > 0;
  ScratchGuard guard(st);
  if (want_obs) {
    int64_t big = p->n_slots / 1024;
    if (big < 8192) big = 8192;
    if (const char* be = getenv("MWD_HMM_POST_BIG")) big = atoll(be);
    if (big < 1) big = 1;
    while ((p->n_slots / big + 1) * kPostSplit * 16 > (int64_t(1) << 30)) big *= 2;   // bound the scratch
    const int64_t cap = p->n_slots / big + 1;                 // most entries that can exceed `big`
    const size_t list_bytes = ((size_t)cap * sizeof(int64_t) + 15) / 16 * 16;
    cudaMemPool_t pool;
    if (int rc = scratch_pool(&pool)) return rc;
    MWD_CHECK_CUDA(cudaMallocFromPoolAsync(&guard.ptr, 16 + list_bytes + (size_t)cap * kPostSplit * 2 * sizeof(double),
                                           pool, st));
    auto* big_count = static_cast<unsigned long long*>(guard.ptr);
    auto* big_list = reinterpret_cast<int64_t*>(static_cast<char*>(guard.ptr) + 16);
    auto* part = reinterpret_cast<double*>(static_cast<char*>(guard.ptr) + 16 + list_bytes);
    MWD_CHECK_CUDA(cudaMemsetAsync(big_count, 0, 16, st));
    if (p->log_domain) {
      hmm_postings_kernel<true><<<og, 256, 0, st>>>(p->post, post_idx, post_off, oe, counts, big, big_list, big_count);
      hmm_postings_big_kernel<true><<<kPostSplit, kPostThreads, 0, st>>>(p->post, post_idx, post_off, big_list,
                                                                        big_count, part);
      hmm_postings_big_combine_kernel<true><<<(unsigned)((cap + 127) / 128), 128, 0, st>>>(big_list, big_count,
                                                                                         part, counts);
    } else {
      hmm_postings_kernel<false><<<og, 256, 0, st>>>(p->post, post_idx, post_off, oe, counts, big, big_list, big_count);
      hmm_postings_big_kernel<false><<<kPostSplit, kPostThreads, 0, st>>>(p->post, post_idx, post_off, big_list,
                                                                         big_count, part);
      hmm_postings_big_combine_kernel<false><<<(unsigned)((cap + 127) / 128), 128, 0, st>>>(big_list, big_count,
                                                                                          part, counts);
    }
    MWD_CHECK_LAUNCH();
  }
  if (p->log_domain) {
    hmm_fill_kernel<<<fg, 256, 0, st>>>(counts + oe, ie + te, -INFINITY);
    if (la.n)
      hmm_reduce_blocks_kernel<true><<<2 * la.n, kReduceThreads, 0, st>>>(la, p->part_init, p->part_trans, rows,
                                                                        counts + oe, counts + oe + ie);
  } else {
    hmm_fill_kernel<<<fg, 256, 0, st>>>(counts + oe, ie + te, 0.0);
    if (la.n)
      hmm_reduce_blocks_kernel<false><<<2 * la.n, kReduceThreads, 0, st>>>(la, p->part_init, p->part_trans, rows,
                                                                         counts + oe, counts + oe + ie);
  }
  MWD_CHECK_LAUNCH();
  // log-likelihood sum; stage-1 partials parked in the (already consumed) head of part_trans
  return sum_doubles(p->pair_ll, p->n_pairs, p->part_trans, counts + oe + ie + te, st);
}

extern "C" int mwd_hmm_mstep(const mwd_hmm_mstep_args* a, void* stream) {
  cudaStream_t st = as_stream(stream);
  MWD_REQUIRE(a->n_lens >= 1 && a->n_lens <= kNMax, "n_lens %d outside [1,%d]", a->n_lens, kNMax);
  HmmLens la;
  la.n = a->n_lens;
  for (int i = 0; i < a->n_lens; ++i) {
    MWD_REQUIRE(a->lens[i] >= 1 && a->lens[i] <= kNMax, "length %d outside [1,%d]", a->lens[i], kNMax);
    la.lens[i] = a->lens[i];
  }
  const int Vt = a->n_tgt_types, Vf = a->n_src_types;
  const int64_t oe = (int64_t)Vt * Vf;
  const int64_t ie = (int64_t)(kNMax + 1) * kNMax;
  const double* obsC = a->counts;
  const double* initC = a->counts + oe;
  const double* transC = a->counts + oe + ie;
  if (a->log_domain) {
    MWD_REQUIRE(a->acc != nullptr, "log-domain M-step needs the running accumulators");
    hmm_mstep_init_trans_kernel<true><<<a->n_lens, 64, 0, st>>>(la, initC, transC, a->acc + oe, a->acc + oe + ie,
                                                             a->init, a->trans);
    if (Vf > 0) hmm_mstep_obs_kernel<true><<<(Vt + 7) / 8, 256, 0, st>>>(obsC, a->acc, Vt, Vf, a->obs);
  } else {
    hmm_mstep_init_trans_kernel<false><<<a->n_lens, 64, 0, st>>>(la, initC, transC, nullptr, nullptr, a->init,
                                                              a->trans);
    hmm_mstep_obs_kernel<false><<<(Vt + 7) / 8, 256, 0, st>>>(obsC, nullptr, Vt, Vf, a->obs);
  }
  MWD_CHECK_LAUNCH();
  return 0;
}

extern "C" int mwd_hmm_align(const mwd_hmm_problem* p, double unk_prob, int32_t* alignment,
                             double* align_probs, const int64_t* ap_off, void* stream) {
  cudaStream_t st = as_stream(stream);
  if (p->n_pairs <= 0) return 0;
  MWD_REQUIRE(align_probs == nullptr || ap_off != nullptr, "align_probs needs ap_off");
  MWD_REQUIRE(p->n_buckets > 0 && p->bucket_lo && p->bucket_n && p->bucket_tmax, "align needs the launch groups");
  HmmAlignArgs a;
  a.tgt_off = p->tgt_off; a.tgt = p->tgt; a.src_off = p->src_off; a.src = p->src;
  a.init = p->init; a.trans = p->trans; a.obs = p->obs;
  a.alignment = alignment; a.align_probs = align_probs; a.ap_off = ap_off;
  a.emis = p->emis; a.slot_off = p->slot_off;
  a.n_pairs = p->n_pairs; a.Vf = p->n_src_types; a.Tmax = p->t_max; a.unk = unk_prob;
  const char* pk_env = getenv("MWD_HMM_PACKED");
  const bool packed_ok = !(pk_env && atoi(pk_env) == 0);
  for (int b = 0; b < p->n_buckets; ++b) {
    const int n = p->bucket_n[b];
    a.lo = p->bucket_lo[b];
    a.hi = p->bucket_lo[b + 1];
    if (a.hi <= a.lo) continue;
    MWD_REQUIRE(n >= 1 && n <= kNMax, "bucket %d: %d states outside [1,%d]", b, n, kNMax);
    a.Tmax = p->bucket_tmax[b];
    const int64_t npairs = a.hi - a.lo;
    if (n <= 8 && packed_ok) {
      // several pairs per warp; back-pointers are one byte per (t, lane)
      const int G = 32 / n;
      const size_t per_warp = 2 * 32 * sizeof(double) + (size_t)a.Tmax * 32;
      int wpc = (int)((220 * 1024) / per_warp);
      if (wpc > 8) wpc = 8;
      MWD_REQUIRE(wpc >= 1, "caption of %d tokens does not fit in shared memory", a.Tmax);
      a.warps_per_cta = wpc;
      const size_t smem = per_warp * wpc;
      const int64_t groups = (npairs + G - 1) / G;
      const int64_t grid = (groups + wpc - 1) / wpc;
      MWD_REQUIRE(grid <= 0x7fffffff, "too many pairs for one launch");
      auto launch_p = [&](auto kern) -> int {
        MWD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)grid, wpc * 32, smem, st>>>(a);
        return 0;
      };
      int rc;
      switch (n) {
#define MWD_AP(V)                                                                                                 \
  case V:                                                                                                         \
    rc = p->log_domain ? launch_p(hmm_align_packed_kernel<true, V>) : launch_p(hmm_align_packed_kernel<false, V>); \
    break;
        MWD_AP(1) MWD_AP(2) MWD_AP(3) MWD_AP(4) MWD_AP(5) MWD_AP(6) MWD_AP(7) MWD_AP(8)
#undef MWD_AP
        default: rc = 2; break;
      }
      if (rc) return rc;
      MWD_CHECK_LAUNCH();
      continue;
    }
    size_t per_warp = 2 * kNMax * sizeof(double) + (size_t)a.Tmax * kNMax;
    per_warp = (per_warp + 7) / 8 * 8;
    int wpc = (int)((220 * 1024) / per_warp);
    if (wpc > 8) wpc = 8;
    MWD_REQUIRE(wpc >= 1, "caption of %d tokens does not fit in shared memory", a.Tmax);
    a.warps_per_cta = wpc;
    const size_t smem = per_warp * wpc;
    const int64_t grid = (npairs + wpc - 1) / wpc;
    MWD_REQUIRE(grid <= 0x7fffffff, "too many pairs for one launch");
    if (p->log_domain) {
      MWD_CHECK_CUDA(cudaFuncSetAttribute(hmm_align_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      hmm_align_kernel<true><<<(unsigned)grid, wpc * 32, smem, st>>>(a);
    } else {
      MWD_CHECK_CUDA(cudaFuncSetAttribute(hmm_align_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      hmm_align_kernel<false><<<(unsigned)grid, wpc * 32, smem, st>>>(a);
    }
    MWD_CHECK_LAUNCH();
  }
  return 0;
}

extern "C" int mwd_hmm_gauss_emission(const mwd_hmm_problem* p, const void* emb, int emb_is_f64, int D,
                                      int M, const double* lprior, const double* means, const double* var,
                                      double* lnorm, double* emis, double* resp, void* stream) {
  cudaStream_t st = as_stream(stream);
  MWD_REQUIRE(M >= 1 && M <= 8, "n_mix %d outside [1,8]", M);
  MWD_REQUIRE(p->row_pair != nullptr, "segment model needs row_pair");
  const int64_t rows = p->n_src_rows;
  gauss_lnorm_kernel<<<p->n_tgt_types * M, 32, 0, st>>>(var, D, lnorm);
  if (rows > 0) {
    const unsigned grid = (unsigned)((rows + 7) / 8);
    if (emb_is_f64)
      gauss_emission_kernel<double><<<grid, 256, 0, st>>>(p->tgt_off, p->tgt, p->src_off, p->slot_off, p->n_pairs,
          (const double*)emb, D, M, lprior, means, var, lnorm, emis, resp, rows, p->row_pair);
    else
      gauss_emission_kernel<float><<<grid, 256, 0, st>>>(p->tgt_off, p->tgt, p->src_off, p->slot_off, p->n_pairs,
          (const float*)emb, D, M, lprior, means, var, lnorm, emis, resp, rows, p->row_pair);
  }
  MWD_CHECK_LAUNCH();
  return 0;
}

extern "C" int mwd_hmm_gauss_stats(const mwd_hmm_problem* p, const void* emb, int emb_is_f64, int D, int M,
                                   const double* resp, const int64_t* word_idx, const int64_t* word_off,
                                   double* stats, void* stream) {
  cudaStream_t st = as_stream(stream);
  MWD_REQUIRE(p->slot_row != nullptr, "segment model needs slot_row");
  const char* sp_env = getenv("MWD_GAUSS_STATS_SPLIT");
  if (!(sp_env && atoi(sp_env) == 0) && p->n_tgt_types > 0) {
    const int row_len = M * (1 + 2 * D);
    int S = 32;
    while (S > 1 && (size_t)p->n_tgt_types * S * row_len * sizeof(double) > (size_t(1) << 30)) S /= 2;
    ScratchGuard guard(st);
    cudaMemPool_t pool;
    if (int rc = scratch_pool(&pool)) return rc;
    MWD_CHECK_CUDA(cudaMallocFromPoolAsync(&guard.ptr, (size_t)p->n_tgt_types * S * row_len * sizeof(double), pool, st));
    double* part = static_cast<double*>(guard.ptr);
    const dim3 grid((unsigned)p->n_tgt_types, (unsigned)S);
    if (emb_is_f64)
      gauss_stats_slice_kernel<double><<<grid, 128, 0, st>>>(word_idx, word_off, p->slot_row, (const double*)emb, D, M,
                                                            p->post, resp, part);
    else
      gauss_stats_slice_kernel<float><<<grid, 128, 0, st>>>(word_idx, word_off, p->slot_row, (const float*)emb, D, M,
                                                           p->post, resp, part);
    gauss_stats_combine_kernel<<<p->n_tgt_types, 128, 0, st>>>(part, S, row_len, stats);
    MWD_CHECK_LAUNCH();
    return 0;
  }
  if (emb_is_f64)
    gauss_stats_kernel<double><<<p->n_tgt_types, 128, 0, st>>>(word_idx, word_off, p->slot_row, (const double*)emb,
                                                              D, M, p->post, resp, stats);
  else
    gauss_stats_kernel<float><<<p->n_tgt_types, 128, 0, st>>>(word_idx, word_off, p->slot_row, (const float*)emb,
                                                             D, M, p->post, resp, stats);
  MWD_CHECK_LAUNCH();
  return 0;
}

extern "C" int mwd_hmm_gauss_update(int Vt, int M, int D, const double* stats, int update_var,
                                    double* lprior, double* means, double* var, void* stream) {
  gauss_update_kernel<<<Vt, 128, 0, as_stream(stream)>>>(M, D, stats, update_var, lprior, means, var);
  MWD_CHECK_LAUNCH();
  return 0;
}

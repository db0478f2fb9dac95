// K1 -- fused forward / backward / expected-count kernel of the (region i, concept k)-state HMM.
//
// Replaces the per-caption Python loop body of
//   hmm_dnn/image_phone_hmm_word_discoverer.py: forward :276-304, backward :314-335,
//   updateInitialCounts :347-362, updateTransitionCounts :374-416, updateStateCounts :426-433,
//   computeAvgLogLikelihood :523-531 and the phoneCounts / conceptCountsA lines :233,:235.
//
// Work decomposition (one launch per bucket of equal n):
//   * a persistent CTA processes kPairsPerCta caption-image pairs in lock-step;
//   * 8 lanes share one (pair, region i) row of the (i,k) lattice, lane l owns concepts
//     k = l + 8*j (j < KG), so the per-row sums over k are KG local adds + a 3-step shuffle;
//   * the cross-region coupling (s_t -> c_t, r_t -> w_t) goes through a double-buffered shared
//     exchange array with ONE __syncthreads per time step;
//   * alpha is checkpointed every B steps into an L2-resident per-CTA scratch; the backward sweep
//     recomputes the B-1 slices in between into shared memory (barrier-free: the coupling terms
//     c_t were stored on the way forward), so no (T,n,K) tensor ever reaches HBM;
//   * everything stays in the raw float64 probability domain with the reference's EPS floors.
//   * the recursion kernel only publishes per-step row statistics (s_t, floor-sum, xi diagonal,
//     r_t: 4n doubles per step); the init / transition count bookkeeping (EPS floors, per-step
//     normalisers) runs in a separate, fully parallel post-pass kernel (ik_counts_kernel), so it
//     never sits on the recursion's critical path;
//   * counts are accumulated without atomics: init/trans in registers of fixed warps, phone
//     counts by read-modify-write of a per-CTA table that only this CTA touches, in t order.
#include <stdlib.h>

#include "ik_estep.cuh"

namespace mwd {


constexpr int BMAX = 8; // max checkpoint interval
constexpr int kEPP = 4; // pairs per CTA

// phone-count table update for one column k and the PP rows deferred by the previous step.
// All loads are issued first; equal phone ids are forwarded so the result equals the sequential
// slot-ordered read-modify-write.
template <int PP>
__device__ __forceinline__ void drain_column(double* tab, int K, int k, const int* xs, const double* cA) {
  double v[PP];
#pragma unroll
  for (int s = 0; s < PP; ++s) v[s] = (xs[s] >= 0) ? tab[xs[s] * K + k] : 0.0;
#pragma unroll
  for (int s = 0; s < PP; ++s) {
    if (xs[s] < 0) continue;
#pragma unroll
    for (int q = 0; q < s; ++q)
      if (xs[q] == xs[s]) v[s] = v[q];          // latest earlier slot with the same phone wins
    v[s] += cA[s * K + k];
  }
#pragma unroll
  for (int s = 0; s < PP; ++s)
    if (xs[s] >= 0) tab[xs[s] * K + k] = v[s];
}

// concept_alignment of the PP rows deferred by the previous step: warp w takes rows w, w + nwarps, ...
// (argmax_k conceptCountsA[t][k], first index on ties, printAlignment :628)
template <int PP>
__device__ __forceinline__ void argmax_rows_deferred(int32_t* ca_out, int K, const long long* pos, const double* cA) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int s = warp; s < PP; s += nwarps) {
    const long long at = pos[s];
    if (at < 0) continue;
    double bv = 0.0;
    int bk = 0x7fffffff;
    for (int k = lane; k < K; k += 32) {
      const double v = cA[s * K + k];
      if (bk == 0x7fffffff || argmax_better(v, bv)) { bv = v; bk = k; }
    }
    const int kbest = warp_argmax_nonneg(bv, bk);
    if (lane == 0) ca_out[at] = kbest;
  }
}

constexpr int kEstepThreadsPerSm = 480;   // register budget: 65536 / 480 = 136 per thread (3 x 160-thread CTAs)
constexpr int estep_min_blocks(int nn) { return nn <= 0 ? 1 : (kEstepThreadsPerSm / (32 * nn) < 1 ? 1 : kEstepThreadsPerSm / (32 * nn)); }

// KG = ceil(K/8) exactly: only the last of a lane's KG concepts can fall outside [0,K).
// NN > 0: n is a compile-time constant (index math folds, the n-loops unroll); NN == 0: generic.
template <int KG, int NN, bool TAB>
__global__ void __launch_bounds__(NN > 0 ? 32 * NN : 512, estep_min_blocks(NN))
ik_estep_kernel(const EstepArgs a) {
  constexpr int PP = kEPP;
  constexpr int KS = KG * kLanesPerRow;
  const int n = (NN > 0) ? NN : a.n;
  const int K = a.K, B = a.B;
  const double eps = a.eps;
  const int tid = threadIdx.x;
  const int l8 = tid & 7;
  const int grp = tid >> 3;
  const int slot = grp / n;
  const int i = grp - slot * n;
  const int TX = a.Tmax;
  const bool kv_last = (l8 + 8 * (KG - 1)) < K;   // validity of this lane's last concept

  extern __shared__ double smem[];
  // shared-memory map (offsets in doubles)
  const int o_exch = PP * B * n * KS;                     // [2][PP][NMAX]
  const int o_aoff = o_exch + 2 * PP * kNMax;             // [n][n]
  const int o_d = o_aoff + kNMax * kNMax;                 // [n]
  const int o_pi = o_d + kNMax;                           // [n]
  const int o_in = o_pi + kNMax;                          // [PP]  1 / max(L, EPS)
  const int o_cA = o_in + PP;                             // [2][PP][K]
  const int o_tab = o_cA + 2 * PP * K;                    // [P][K]   (TAB)
  const int o_x = o_tab + (TAB ? a.P * K : 0);            // ints: [PP][TX]
  double* s_buf = smem;
  double* s_exch = smem + o_exch;
  double* s_aoff = smem + o_aoff;
  double* s_inorm = smem + o_in;
  double* s_cA = smem + o_cA;
  int* s_x = reinterpret_cast<int*>(smem + o_x);
  __shared__ int s_T[PP];
  __shared__ int s_xs[2][PP];                             // phone id of the deferred row, -1 = none
  __shared__ long long s_pos[2][PP];                      // its position p0 + t in the packed phone array, -1 = none

  // transition / initial tables of this n
  for (int e = tid; e < n * n; e += blockDim.x) {
    int r = e / n, c = e - r * n;
    double v = a.trans[e];
    s_aoff[e] = (r == c) ? 0.0 : v;
    if (r == c) smem[o_d + r] = v;
  }
  for (int e = tid; e < n; e += blockDim.x) smem[o_pi + e] = a.init[e];
  if (TAB)
    for (int e = tid; e < a.P * K; e += blockDim.x) smem[o_tab + e] = 0.0;

  // checkpoint slices are padded to whole 128-byte lines (SL doubles) so that a consumed slice can
  // be dropped from L2 with discard.global.L2 instead of being written back to HBM
  const int SL = (n * KS + 15) & ~15;
  double* cta_scr = a.scratch + (size_t)blockIdx.x * a.cta_scratch;
  double* my_ckpt = cta_scr + (size_t)slot * a.NC * SL + i * KS + l8;               // + c*SL + 8j
  double* my_hist = cta_scr + (size_t)PP * a.NC * SL + (size_t)slot * a.Tmax * n + i;  // + t*n : c_t[i]
  const int buf_row = (slot * B * n + i) * KS + l8;                                  // + tt*n*KS + 8j
  const int nKS = n * KS;
  // part_phone == NULL (dense-emission classes): no phone table, the caller consumes cA_out instead
  const bool tab_on = a.part_phone != nullptr;
  double* g_phone = a.part_phone + (size_t)blockIdx.x * a.P * K;
  double* tab = TAB ? (smem + o_tab) : g_phone;
  const int* my_x = s_x + slot * TX;
  const double* obsT = a.obsT + l8;

  const int64_t npairs = a.hi - a.lo;
  const int64_t nquads = (npairs + PP - 1) / PP;
  __syncthreads();
  const double d_i = smem[o_d + i];
  const double pi_i = smem[o_pi + i];
  double aoff_row[NN > 0 ? NN : 1];     // Aoff[i][:] in registers when n is static
  if (NN > 0) {
#pragma unroll
    for (int j = 0; j < NN; ++j) aoff_row[j] = s_aoff[i * n + j];
  }

  for (int64_t quad = blockIdx.x; quad < nquads; quad += gridDim.x) {
    const int64_t pair = a.lo + quad * PP + slot;
    const bool valid = pair < a.hi;
    int T = 0;
    int64_t r0 = 0, p0 = 0;
    double* my_stats = a.stats;
    if (valid) {
      p0 = a.phone_off[pair];
      T = a.phone_off[pair + 1] - (int32_t)p0;
      r0 = a.region_off[pair];
      my_stats = a.stats + 4 * a.slot_off[pair] + i;      // + (t*4 + q)*n
    }
    __syncthreads();  // previous quad fully done (exchange + buffers reusable)
    if (i == 0 && l8 == 0) {
      s_T[slot] = T;
      s_xs[0][slot] = -1;
      s_xs[1][slot] = -1;
      s_pos[0][slot] = -1;
      s_pos[1][slot] = -1;
    }
    for (int t = i * kLanesPerRow + l8; t < T; t += n * kLanesPerRow)
      s_x[slot * TX + t] = a.phones[p0 + t];
    __syncthreads();
    int Tmax = 0;
#pragma unroll
    for (int s = 0; s < PP; ++s) Tmax = max(Tmax, s_T[s]);

    double pz[KG];
    {
      const double* prow = a.pz + (r0 + i) * K + l8;
#pragma unroll
      for (int j = 0; j < KG; ++j)
        pz[j] = (valid && (j < KG - 1 || kv_last)) ? prow[8 * j] : 0.0;
    }

    // ------------------------------------------------------------------ forward sweep
    double al[KG];
    {
      const double* orow = obsT + (valid ? my_x[0] : 0) * K;
#pragma unroll
      for (int j = 0; j < KG; ++j) {
        double o = (j < KG - 1 || kv_last) ? __ldg(orow + 8 * j) : 0.0;
        al[j] = (pi_i * pz[j]) * o;
      }
    }
    for (int t = 0; t < Tmax; ++t) {
      const bool act = t < T;
      const int par = t & 1;
      double s = 0.0;
      double onext[KG];
      if (act) {
        if (t + 1 < T) {   // next step's emissions: issued before the reduction / barrier
          const double* orow = obsT + my_x[t + 1] * K;
#pragma unroll
          for (int j = 0; j < KG; ++j) onext[j] = (j < KG - 1 || kv_last) ? __ldg(orow + 8 * j) : 0.0;
        }
        if (t % B == 0 && !a.ll_only) {
          double* dst = my_ckpt + (size_t)(t / B) * SL;
#pragma unroll
          for (int j = 0; j < KG; ++j) __stcg(dst + 8 * j, al[j]);
        }
#pragma unroll
        for (int j = 0; j < KG; ++j) s += al[j];
      }
      s = row8_sum(s);
      const int exo = (par * PP + slot) * kNMax;
      if (l8 == 0) s_exch[exo + i] = s;
      __syncthreads();
      if (act) {
        const double* ex = s_exch + exo;
        if (t == T - 1) {
          // negative sentinel in the (unused) s slot of the last row: see ik_counts_small_kernel
          if (l8 == 0 && !a.ll_only) __stcg(my_stats + (t * 4 + 0) * n, -1.0);
          if (i == 0 && l8 == 0) {
            double L = 0.0;
            for (int j = 0; j < n; ++j) L += ex[j];
            L = floor_at(L, a.eps);
            a.pair_ll[pair] = log(L);                              // :529
            // sum_{i,k} alpha_t beta_t equals the sentence likelihood at every t, so the floored
            // normaliser of updateStateCounts (:430) is one constant per pair
            s_inorm[slot] = 1.0 / L;
          }
        } else {
          double c = 0.0;
#pragma unroll
          for (int j = 0; j < (NN > 0 ? NN : kNMax); ++j)
            if (NN > 0 || j < n) c = fma(s_aoff[j * n + i], ex[j], c);
          if (l8 == 0 && !a.ll_only) {
            __stcg(my_hist + t * n, c);
            __stcg(my_stats + (t * 4 + 0) * n, s);
          }
#pragma unroll
          for (int j = 0; j < KG; ++j) al[j] = onext[j] * fma(d_i, al[j], c * pz[j]);
        }
      }
    }

    if (a.ll_only) continue;

    // ------------------------------------------------------------------ backward sweep
    double bo[KG];          // beta_{t+1} * o_{t+1}
#pragma unroll
    for (int j = 0; j < KG; ++j) bo[j] = 0.0;
    double w = 0.0;         // (Aoff r_{t+1})[i]
    int step = 0;           // parity counter of the exchange buffers

    const int nblk = (Tmax + B - 1) / B;
    for (int c = nblk - 1; c >= 0; --c) {
      const int t0 = c * B;
      const int len = min(B, Tmax - t0);
      __syncthreads();  // column reads of the previous block are complete
      if (t0 < T) {
        const double* src = my_ckpt + (size_t)c * SL;
        double cb[BMAX];
#pragma unroll
        for (int tt = 1; tt < BMAX; ++tt)
          cb[tt] = (tt < len && t0 + tt < T) ? __ldcg(my_hist + (t0 + tt - 1) * n) : 0.0;
#pragma unroll
        for (int j = 0; j < KG; ++j) {
          al[j] = __ldcg(src + 8 * j);
          s_buf[buf_row + 8 * j] = al[j];
        }
#pragma unroll
        for (int tt = 1; tt < BMAX; ++tt) {
          if (tt < len && t0 + tt < T) {
            const double* orow = obsT + my_x[t0 + tt] * K;
            double* dst = s_buf + buf_row + tt * nKS;
#pragma unroll
            for (int j = 0; j < KG; ++j) {
              double o = (j < KG - 1 || kv_last) ? __ldg(orow + 8 * j) : 0.0;
              al[j] = o * fma(d_i, al[j], cb[tt] * pz[j]);
              dst[8 * j] = al[j];
            }
          }
        }
      }
      for (int tt = len - 1; tt >= 0; --tt, ++step) {
        const int t = t0 + tt;
        const bool act = t < T;
        const int par = step & 1;
        double sumF = 0.0, dg = 0.0, rr = 0.0;
        int x = 0;
        if (act) {
          x = my_x[t];
          const bool last = (t == T - 1);
          const double* orow = obsT + x * K;
          double* row = s_buf + buf_row + tt * nKS;
#pragma unroll
          for (int j = 0; j < KG; ++j) {
            const bool kv = (j < KG - 1) || kv_last;
            double av = row[8 * j];
            double beta = last ? 1.0 : fma(d_i, bo[j], w);
            dg = fma(av, bo[j], dg);
            double g = av * beta;
            sumF += kv ? floor_at(g, eps) : 0.0;
            double o = kv ? __ldg(orow + 8 * j) : 0.0;
            bo[j] = beta * o;
            rr = fma(bo[j], pz[j], rr);
            row[8 * j] = g;
          }
          dg *= d_i;
        }
        sumF = row8_sum(sumF);
        dg = row8_sum(dg);
        rr = row8_sum(rr);
        const int exo = (par * PP + slot) * kNMax;
        if (l8 == 0) {
          s_exch[exo + i] = rr;
          if (act) {   // row statistics of this step for the count post-pass
            __stcg(my_stats + (t * 4 + 1) * n, sumF);
            __stcg(my_stats + (t * 4 + 2) * n, dg);
            __stcg(my_stats + (t * 4 + 3) * n, rr);
          }
        }
        __syncthreads();
        if (tt == len - 1) {
          // every thread has re-read its part of checkpoint slice c: the slice is dead, drop it
          const int lines = SL >> 4;
          for (int ln = tid; ln < PP * lines; ln += blockDim.x) {
            const int sl = ln / lines, l = ln - sl * lines;
            if (t0 < s_T[sl]) {
              const double* dead = cta_scr + ((size_t)sl * a.NC + c) * SL + l * 16;
              asm volatile("discard.global.L2 [%0], 128;" ::"l"(dead) : "memory");
            }
          }
        }
        // drain the phone-count rows deferred by the previous step: thread k owns column k of
        // the per-CTA table for every slot, so the read-modify-writes never race and their
        // order (t descending, slot ascending) is fixed.
        if (tab_on)
          for (int k = tid; k < K; k += blockDim.x)
            drain_column<PP>(tab, K, k, s_xs[par ^ 1], s_cA + (par ^ 1) * PP * K);
        if (a.ca_out) argmax_rows_deferred<PP>(a.ca_out, K, s_pos[par ^ 1], s_cA + (par ^ 1) * PP * K);
        if (i == 0 && l8 == 0) {
          s_xs[par][slot] = act ? x : -1;
          s_pos[par][slot] = act ? (long long)(p0 + t) : -1;
        }
        if (act) {
          const double* ex = s_exch + exo;
          double wn = 0.0;
          if (NN > 0) {
#pragma unroll
            for (int j = 0; j < (NN > 0 ? NN : 1); ++j) wn = fma(aoff_row[j], ex[j], wn);
          } else {
            for (int j = 0; j < n; ++j) wn = fma(s_aoff[i * n + j], ex[j], wn);
          }
          w = wn;
          // conceptCountsA[t][k] = sum_i gamma_t[i][k]; phoneCounts[k][x_t] += ...  (:430,:233,:235)
          const double inorm = s_inorm[slot];
          const double* col = s_buf + (slot * B + tt) * nKS;
          double* crow = s_cA + (par * PP + slot) * K;
          for (int k = i * kLanesPerRow + l8; k < K; k += n * kLanesPerRow) {
            double v = 0.0;
#pragma unroll
            for (int ii = 0; ii < (NN > 0 ? NN : kNMax); ++ii)
              if (NN > 0 || ii < n) v += col[ii * KS + k];
            v *= inorm;
            crow[k] = v;
            if (a.cA_out) a.cA_out[(p0 + t) * K + k] = v;
          }
        }
      }
    }
    // flush the last deferred phone-count rows
    {
      const int par = step & 1;
      __syncthreads();
      if (tab_on)
        for (int k = tid; k < K; k += blockDim.x)
          drain_column<PP>(tab, K, k, s_xs[par ^ 1], s_cA + (par ^ 1) * PP * K);
      if (a.ca_out) argmax_rows_deferred<PP>(a.ca_out, K, s_pos[par ^ 1], s_cA + (par ^ 1) * PP * K);
    }
  }

  if (a.ll_only) return;
  // ---------------------------------------------------------------- per-CTA phone-count table
  __syncthreads();
  if (TAB && tab_on)
    for (int e = tid; e < a.P * K; e += blockDim.x) g_phone[e] += smem[o_tab + e];
}

// ------------------------------------------------------------------------------------------
// Count post-pass: updateInitialCounts (:347-362) and updateTransitionCounts (:374-416) from the
// per-step row statistics.  One warp per pair (persistent, fixed pair->warp map), lanes own the
// n x n transition entries; every time step is independent, so nothing here is latency-critical.
// ------------------------------------------------------------------------------------------
struct CountsArgs {
  const int32_t* phone_off;
  const double* trans;     // trans[n]
  const double* stats;
  const int64_t* slot_off;
  double* part_init;
  double* part_trans;
  int64_t lo, hi;
  int n, total_warps;
  double eps;            // xi floor (MWD_EPS or 0)
  int compact;           // rows written by the float32 kernel: see StatRow
  int K;                 // concept count (compact rows of an all-floored step stand for K * eps)
};

// One (pair, t) row of the statistics the E-step kernels hand to the count post-pass.  Float64 kernels: 4 n doubles
// [s | F | dg | r].  The float32 kernel (ik_estep_warp32.cu) writes the SAME slot compactly -- 4 n float32 mantissas,
// then the binary exponents of s, of (F, dg) and of r, then a flag "every entry of the step was floored" (F = K eps)
// -- 16 n + 16 bytes instead of 32 n, rows packed densely from the start of the bucket's slot range: the un-scaling
// (double)m * 2^e  it used to do before the store happens here, with identical results.
struct StatRow {
  const double* row;
  double es, ef, er;     // 2^e factors (1 for float64 rows)
  bool compact, floored;
  __device__ __forceinline__ StatRow(const double* r, bool c) : row(r), es(1.0), ef(1.0), er(1.0), compact(c), floored(false) {}
  // row g (counted from the bucket's first row) of a bucket whose statistics start at `base`
  __device__ __forceinline__ static const double* at(const double* base, int64_t g, int n, bool c) {
    return c ? reinterpret_cast<const double*>(reinterpret_cast<const float*>(base) + g * (4 * n + 4)) : base + g * 4 * n;
  }
  __device__ __forceinline__ const double* next(int n) const { return at(row, 1, n, compact); }
  __device__ __forceinline__ static double p2(int e) { return __longlong_as_double((long long)(max(e, -1022) + 1023) << 52); }
  __device__ __forceinline__ void load_exps(int n) {
    if (compact) {
      const int4 e = __ldcs(reinterpret_cast<const int4*>(reinterpret_cast<const float*>(row) + 4 * n));
      es = p2(e.x); ef = p2(e.y); er = p2(e.z); floored = e.w != 0;
    }
  }
  __device__ __forceinline__ double raw(int idx) const {
    return compact ? (double)__ldcs(reinterpret_cast<const float*>(row) + idx) : __ldcs(row + idx);
  }
};

__device__ __forceinline__ double warp_sum_all(double v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

__global__ void __launch_bounds__(256) ik_counts_kernel(const CountsArgs a) {
  constexpr int EPL = (kNMax * kNMax + 31) / 32;   // transition entries per lane (8)
  const int n = a.n;
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nn = n * n;
  const int epl = (nn + 31) >> 5;                  // entries per lane actually used
  // per-lane constants of its entries e = lane + 32 q -> (r, c): offsets into the stats rows
  int off_a[EPL], off_b[EPL];                      // xi = stats[off_a] * coef * next[off_b] (off-diag)
  double coef[EPL];                                //      stats[off_a]                      (diag: coef < 0)
#pragma unroll
  for (int q = 0; q < EPL; ++q) {
    const int e = lane + 32 * q;
    off_a[q] = 0; off_b[q] = 0; coef[q] = 0.0;
    if (e < nn) {
      const int r = e / n, c = e - r * n;
      if (r == c) { off_a[q] = 2 * n + r; coef[q] = -1.0; }
      else { off_a[q] = r; off_b[q] = 3 * n + c; coef[q] = a.trans[e]; }
    }
  }
  double acc_t[EPL];
#pragma unroll
  for (int q = 0; q < EPL; ++q) acc_t[q] = 0.0;
  double acc_i = 0.0;
  for (int64_t pair = a.lo + gw; pair < a.hi; pair += a.total_warps) {
    const int T = a.phone_off[pair + 1] - a.phone_off[pair];
    const int64_t s_lo = a.slot_off[a.lo];
    const double* base = a.stats + 4 * s_lo;
    const int64_t g0 = (a.slot_off[pair] - s_lo) / n;      // the pair's first row, counted from the bucket's first
    for (int t = 0; t < T; ++t) {
      StatRow row(StatRow::at(base, g0 + t, n, a.compact != 0), a.compact != 0);
      row.load_exps(n);
      // updateInitialCounts: sum_k max(g,EPS) per region over its total (:355)
      const double f = (lane < n) ? (row.floored ? (double)a.K * a.eps : row.raw(n + lane) * row.ef) : 0.0;
      double xv[EPL];
      double z = 0.0;
      if (t < T - 1) {
        // xi_t = diag(d o beta alpha) + s_t Aoff r_{t+1}, EPS-floored, normalised (:388-396)
        StatRow nxt(row.next(n), a.compact != 0);
        nxt.load_exps(n);
#pragma unroll
        for (int q = 0; q < EPL; ++q) {
          xv[q] = 0.0;
          if (q < epl && lane + 32 * q < nn) {
            // off_a is a slot of s (< n) for an off-diagonal entry, of dg for a diagonal one; off_b a slot of r
            const double u = row.raw(off_a[q]) * ((coef[q] < 0.0) ? row.ef : row.es);
            const double xi = (coef[q] < 0.0) ? u : (u * coef[q]) * (nxt.raw(off_b[q]) * nxt.er);
            xv[q] = floor_at(xi, a.eps);
            z += xv[q];
          }
        }
      }
      const double Ft = warp_sum_all(f);
      acc_i += f / Ft;
      if (t < T - 1) {
        const double iz = 1.0 / warp_sum_all(z);
#pragma unroll
        for (int q = 0; q < EPL; ++q)
          if (q < epl) acc_t[q] = fma(xv[q], iz, acc_t[q]);
      }
    }
  }
  // combine the CTA's 8 warps in fixed order, then add into this CTA's partial row
  __shared__ double s_red[8][kNMax * kNMax + kNMax];
  const int wic = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < EPL; ++q) {
    const int e = lane + 32 * q;
    if (e < nn) s_red[wic][e] = acc_t[q];
  }
  if (lane < n) s_red[wic][kNMax * kNMax + lane] = acc_i;
  __syncthreads();
  double* pt = a.part_trans + ((size_t)blockIdx.x * (kNMax + 1) + n) * (kNMax * kNMax);
  double* pi_out = a.part_init + ((size_t)blockIdx.x * (kNMax + 1) + n) * kNMax;
  for (int e = threadIdx.x; e < nn; e += blockDim.x) {
    double v = 0.0;
    for (int w2 = 0; w2 < 8; ++w2) v += s_red[w2][e];
    pt[e] += v;
  }
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    double v = 0.0;
    for (int w2 = 0; w2 < 8; ++w2) v += s_red[w2][kNMax * kNMax + e];
    pi_out[e] += v;
  }
}

// Same post-pass for small n (<= 6): lanes run over TIME, each lane keeps all n*n + n accumulators
// in registers, so a (pair, t) step costs ~10 warp-instructions instead of ~100.
template <int N>
__global__ void __launch_bounds__(256) ik_counts_small_kernel(const CountsArgs a) {
  const double eps = a.eps;
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  double aoff[N * N];
#pragma unroll
  for (int e = 0; e < N * N; ++e) aoff[e] = (e / N == e % N) ? 0.0 : a.trans[e];
  double acc_t[N * N], acc_i[N];
#pragma unroll
  for (int e = 0; e < N * N; ++e) acc_t[e] = 0.0;
#pragma unroll
  for (int e = 0; e < N; ++e) acc_i[e] = 0.0;
  // Row-parallel: the bucket's (pair, t) rows are contiguous in `stats` (4N doubles each); every warp
  // takes one contiguous chunk of rows, a lane one row at a time -- no per-pair metadata, full lanes.
  // A pair's last row carries a negative sentinel in its s slot (no xi there).
  const int64_t s_lo = a.slot_off[a.lo], rows = (a.slot_off[a.hi] - s_lo) / N;   // slots = T*n per pair
  const double* base = a.stats + 4 * s_lo;
  const int64_t chunk = (rows + a.total_warps - 1) / a.total_warps;
  const int64_t g_end = min(rows, (int64_t)(gw + 1) * chunk);
  for (int64_t g = (int64_t)gw * chunk + lane; g < g_end; g += 32) {
    {
      double f[N], sv[N], dgv[N], rn[N], Ft = 0.0;
      bool has_next;
      if (a.compact) {
        // the lane's whole row (4 N mantissas, 3 exponents, flag) in N + 1 128-bit loads; of the next row only r and its exponent
        const float* rp = reinterpret_cast<const float*>(base) + g * (4 * N + 4);
        float rw[4 * N + 4];
#pragma unroll
        for (int q = 0; q < N + 1; ++q) {
          const float4 v = __ldcs(reinterpret_cast<const float4*>(rp) + q);
          rw[4 * q] = v.x; rw[4 * q + 1] = v.y; rw[4 * q + 2] = v.z; rw[4 * q + 3] = v.w;
        }
        const double es = StatRow::p2(__float_as_int(rw[4 * N])), ef = StatRow::p2(__float_as_int(rw[4 * N + 1]));
        const bool floored = __float_as_int(rw[4 * N + 3]) != 0;
#pragma unroll
        for (int r = 0; r < N; ++r) {
          sv[r] = (double)rw[r] * es;
          f[r] = floored ? (double)a.K * eps : (double)rw[N + r] * ef;
          dgv[r] = (double)rw[2 * N + r] * ef;
        }
        has_next = !(sv[0] < 0.0) && g + 1 < rows;
        if (has_next) {
          const float* nx = rp + (4 * N + 4);
          const double er = StatRow::p2(__float_as_int(__ldcs(nx + 4 * N + 2)));
#pragma unroll
          for (int r = 0; r < N; ++r) rn[r] = (double)__ldcs(nx + 3 * N + r) * er;
        }
      } else {
        const double* row = base + (size_t)g * 4 * N;
#pragma unroll
        for (int r = 0; r < N; ++r) { f[r] = __ldcs(row + N + r); sv[r] = __ldcs(row + r); }
        has_next = !(sv[0] < 0.0) && g + 1 < rows;
        if (has_next) {
#pragma unroll
          for (int r = 0; r < N; ++r) {
            dgv[r] = __ldcs(row + 2 * N + r);
            rn[r] = __ldcs(row + 4 * N + 3 * N + r);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < N; ++r) Ft += f[r];
      const double iF = 1.0 / Ft;
#pragma unroll
      for (int r = 0; r < N; ++r) acc_i[r] = fma(f[r], iF, acc_i[r]);                 // :355
      if (has_next) {
        double xv[N * N], z = 0.0;
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
          for (int c = 0; c < N; ++c) {
            const double xi = (r == c) ? dgv[r] : (sv[r] * aoff[r * N + c]) * rn[c];    // :388-389
            xv[r * N + c] = floor_at(xi, eps);                                            // :396
            z += xv[r * N + c];
          }
        const double iz = 1.0 / z;
#pragma unroll
        for (int e = 0; e < N * N; ++e) acc_t[e] = fma(xv[e], iz, acc_t[e]);
      }
    }
  }
  // lanes -> warp totals (fixed butterfly order), then the CTA's 8 warps in fixed order
  __shared__ double s_red[8][N * N + N];
  const int wic = threadIdx.x >> 5;
#pragma unroll
  for (int e = 0; e < N * N; ++e) {
    double v = warp_sum_all(acc_t[e]);
    if (lane == 0) s_red[wic][e] = v;
  }
#pragma unroll
  for (int e = 0; e < N; ++e) {
    double v = warp_sum_all(acc_i[e]);
    if (lane == 0) s_red[wic][N * N + e] = v;
  }
  __syncthreads();
  double* pt = a.part_trans + ((size_t)blockIdx.x * (kNMax + 1) + N) * (kNMax * kNMax);
  double* pi_out = a.part_init + ((size_t)blockIdx.x * (kNMax + 1) + N) * kNMax;
  for (int e = threadIdx.x; e < N * N + N; e += blockDim.x) {
    double v = 0.0;
    for (int w2 = 0; w2 < 8; ++w2) v += s_red[w2][e];
    if (e < N * N) pt[e] += v;
    else pi_out[e - N * N] += v;
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int kg_for(int K) { return (K + kLanesPerRow - 1) / kLanesPerRow; }

// (KG, n) combinations compiled with a static n: the MSCOCO (K=65..72) and Flickr (K=97..104)
// concept counts for every n <= 10; everything else runs the generic kernel.
static bool estep_static_n(int n, int KG) { return (KG == 9 || KG == 13) && n >= 1 && n <= 10; }

struct EstepPlan {
  int KG, KS, PP, B, NC, threads, grid, tab;
  size_t smem;
  int64_t cta_scratch;  // doubles
};

static size_t estep_fixed_smem(int PP, int K, int P, int Tmax, int tab) {
  return ((size_t)2 * PP * kNMax + kNMax * kNMax + 2 * kNMax + PP + (size_t)2 * PP * K +
          (tab ? (size_t)P * K : 0)) * sizeof(double) + (size_t)PP * Tmax * sizeof(int) + 64;
}

static EstepPlan plan_bucket(int n, int K, int P, int Tmax, int64_t npairs) {
  EstepPlan pl;
  pl.KG = kg_for(K);
  pl.KS = pl.KG * kLanesPerRow;
  pl.PP = kPairsPerCta;
  // The per-CTA phone-count table lives in global memory by default (L2-resident, ~25 KB per CTA):
  // measured equal in speed to a shared-memory copy (profiles/r01_bench_history.md) and it leaves
  // the shared memory to a longer checkpoint interval B, i.e. less checkpoint traffic.
  // Tuning knobs (experiments only): MWD_ESTEP_TAB=1 (shared-memory table), MWD_ESTEP_CTAS=<target
  // CTAs/SM>, MWD_ESTEP_B=<minimum checkpoint interval>.
  pl.tab = 0;
  if (const char* e = getenv("MWD_ESTEP_TAB")) pl.tab = (atoi(e) && (size_t)P * K * sizeof(double) <= 64 * 1024) ? 1 : 0;
  const size_t fixed = estep_fixed_smem(pl.PP, K, P, Tmax, pl.tab);
  const size_t slice = (size_t)pl.PP * n * pl.KS * sizeof(double);
  pl.threads = pl.PP * n * kLanesPerRow;
  // aim for 3 co-resident CTAs per SM, at least 2 alpha slices per block, at most BMAX
  const size_t smem_sm = 224 * 1024;
  int target = 3;
  if (const char* e = getenv("MWD_ESTEP_CTAS")) target = atoi(e) > 0 ? atoi(e) : target;
  const int min_b = getenv("MWD_ESTEP_B") ? atoi(getenv("MWD_ESTEP_B")) : 2;
  int B = 0;
  for (; target >= 1; --target) {
    size_t budget = smem_sm / target;
    if (budget <= fixed + 1024) continue;
    B = (int)((budget - fixed - 1024) / slice);
    if (B >= min_b || target == 1) break;
  }
  if (B < 1) B = 1;
  if (B > BMAX) B = BMAX;
  if (B > Tmax) B = Tmax > 0 ? Tmax : 1;
  pl.B = B;
  pl.NC = (Tmax + B - 1) / B;
  if (pl.NC < 1) pl.NC = 1;
  pl.smem = fixed + (size_t)B * slice;
  int64_t nquads = (npairs + pl.PP - 1) / pl.PP;
  int per_sm = (int)(smem_sm / (pl.smem + 1024));
  int by_threads = 2048 / pl.threads;
  int by_regs = estep_static_n(n, pl.KG) ? estep_min_blocks(n) : 1;
  if (per_sm > by_threads) per_sm = by_threads;
  if (per_sm > by_regs) per_sm = by_regs;
  if (per_sm > kEstepCtasPerSm) per_sm = kEstepCtasPerSm;
  if (per_sm < 1) per_sm = 1;
  int64_t grid = (int64_t)sm_count() * per_sm;
  if (grid > nquads) grid = nquads;
  if (grid < 1) grid = 1;
  pl.grid = (int)grid;
  const int64_t SL = ((int64_t)n * pl.KS + 15) & ~(int64_t)15;   // 128-byte aligned checkpoint slices
  pl.cta_scratch = (((int64_t)pl.PP * ((int64_t)pl.NC * SL + (int64_t)Tmax * n)) + 15) & ~(int64_t)15;
  return pl;
}

template <int KG, int NN, bool TAB>
static int launch_estep(const EstepArgs& args, const EstepPlan& pl, cudaStream_t st) {
  auto kern = ik_estep_kernel<KG, NN, TAB>;
  MWD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)pl.smem));
  kern<<<pl.grid, pl.threads, pl.smem, st>>>(args);
  MWD_CHECK_LAUNCH();
  return 0;
}

template <int KG, int NN>
static int launch_estep_tab(const EstepArgs& args, const EstepPlan& pl, cudaStream_t st) {
  if (pl.tab) return launch_estep<KG, NN, true>(args, pl, st);
  return launch_estep<KG, NN, false>(args, pl, st);
}

template <int KG>
static int launch_estep_static(const EstepArgs& args, const EstepPlan& pl, cudaStream_t st) {
  switch (args.n) {
#define MWD_N(V) case V: return launch_estep_tab<KG, V>(args, pl, st);
    MWD_N(1) MWD_N(2) MWD_N(3) MWD_N(4) MWD_N(5) MWD_N(6) MWD_N(7) MWD_N(8) MWD_N(9) MWD_N(10)
#undef MWD_N
  }
  return launch_estep_tab<KG, 0>(args, pl, st);
}

template <int KG>
static int launch_estep_kg(const EstepArgs& args, const EstepPlan& pl, cudaStream_t st) {
  return launch_estep_tab<KG, 0>(args, pl, st);
}
template <>
int launch_estep_kg<9>(const EstepArgs& args, const EstepPlan& pl, cudaStream_t st) {
  return launch_estep_static<9>(args, pl, st);
}
template <>
int launch_estep_kg<13>(const EstepArgs& args, const EstepPlan& pl, cudaStream_t st) {
  return launch_estep_static<13>(args, pl, st);
}

}  // namespace mwd

using namespace mwd;

extern "C" int64_t mwd_ik_scratch_bytes(const mwd_ik_problem* p) {
  int64_t need = 0;
  for (int b = 0; b < p->n_buckets; ++b) {
    int64_t np_ = p->bucket_lo[b + 1] - p->bucket_lo[b];
    if (np_ <= 0) continue;
    EstepPlan pl = plan_bucket(p->bucket_n[b], p->n_concepts, p->n_phone_types, p->bucket_tmax[b], np_);
    int64_t bytes = pl.cta_scratch * pl.grid * (int64_t)sizeof(double);
    if (bytes > need) need = bytes;
    bytes = estep_warp_scratch(p->bucket_n[b], p->n_concepts, p->n_phone_types, p->bucket_tmax[b], np_) * (int64_t)sizeof(double);
    if (bytes > need) need = bytes;
    bytes = estep_warp32_scratch(p->bucket_n[b], p->n_concepts, p->n_phone_types, p->bucket_tmax[b], np_) * (int64_t)sizeof(double);
    if (bytes > need) need = bytes;
  }
  return need;
}

static int estep_impl(const mwd_ik_problem* p, void* stream, int ll_only) {
  MWD_REQUIRE(p->n_concepts >= 1 && p->n_concepts <= MWD_KMAX, "n_concepts %d outside [1,%d]",
              p->n_concepts, MWD_KMAX);
  MWD_REQUIRE(ll_only || (p->stats != nullptr && p->slot_off != nullptr), "mwd_ik_estep needs stats and slot_off");
  cudaStream_t st = as_stream(stream);
  for (int b = 0; b < p->n_buckets; ++b) {
    const int n = p->bucket_n[b];
    const int64_t lo = p->bucket_lo[b], hi = p->bucket_lo[b + 1];
    if (hi <= lo) continue;
    MWD_REQUIRE(n >= 1 && n <= MWD_NMAX, "bucket %d: n=%d outside [1,%d]", b, n, MWD_NMAX);
    // scaled-float32 lattice (MWD_MIXED_RECURSION) where an instantiation exists and the caller consumes phone counts
    const int64_t warp32_scr = ((p->mixed_precision & MWD_MIXED_RECURSION) && p->part_phone != nullptr)
                                   ? estep_warp32_scratch(n, p->n_concepts, p->n_phone_types, p->bucket_tmax[b], hi - lo) : 0;
    const bool use_warp32 = warp32_scr > 0;
    const int64_t warp_scr = use_warp32 ? warp32_scr
                                        : estep_warp_scratch(n, p->n_concepts, p->n_phone_types, p->bucket_tmax[b], hi - lo);
    const bool use_warp = warp_scr > 0;
    EstepPlan pl = plan_bucket(n, p->n_concepts, p->n_phone_types, p->bucket_tmax[b], hi - lo);
    if (use_warp) {   // the warp-per-pair kernel plans its own launch; only the scratch check applies
      pl.smem = 0;
      pl.grid = 1;
      pl.cta_scratch = warp_scr;
    }
    MWD_REQUIRE(pl.smem <= 227 * 1024, "estep shared memory %zu exceeds 227 KB (n=%d K=%d T=%d)",
                pl.smem, n, p->n_concepts, p->bucket_tmax[b]);
    MWD_REQUIRE(ll_only || pl.cta_scratch * pl.grid * (int64_t)sizeof(double) <= p->scratch_bytes,
                "estep scratch too small: need %lld bytes, have %lld",
                (long long)(pl.cta_scratch * pl.grid * (int64_t)sizeof(double)),
                (long long)p->scratch_bytes);
    MWD_REQUIRE(pl.grid <= estep_grid_rows(), "estep grid %d exceeds partial rows %d", pl.grid,
                estep_grid_rows());
    EstepArgs a;
    a.region_off = p->region_off;
    a.phone_off = p->phone_off;
    a.phones = p->phones;
    a.pz = p->pz;
    a.init = p->init + (size_t)n * MWD_INIT_STRIDE;
    a.trans = p->trans + (size_t)n * MWD_TRANS_STRIDE;
    a.obsT = p->obsT;
    a.pair_ll = p->pair_ll;
    a.cA_out = p->concept_counts_a;
    a.ca_out = p->concept_alignment;
    a.part_phone = p->part_phone;
    a.part_init = p->part_init;
    a.part_trans = p->part_trans;
    a.scratch = p->scratch;
    a.stats = p->stats;
    a.slot_off = p->slot_off;
    a.lo = lo;
    a.hi = hi;
    a.cta_scratch = pl.cta_scratch;
    a.n = n;
    a.K = p->n_concepts;
    a.P = p->n_phone_types;
    a.B = pl.B;
    a.NC = pl.NC;
    a.Tmax = p->bucket_tmax[b];
    a.ll_only = ll_only;
    a.eps = p->no_floor ? 0.0 : MWD_EPS;
    int rc = 0;
    if (use_warp32) rc = estep_warp32_launch(a, st);
    else if (use_warp) rc = estep_warp_launch(a, st);
    else switch (pl.KG) {
#define MWD_KG(G) case G: rc = launch_estep_kg<G>(a, pl, st); break;
      MWD_KG(1) MWD_KG(2) MWD_KG(3) MWD_KG(4) MWD_KG(5) MWD_KG(6) MWD_KG(7) MWD_KG(8)
      MWD_KG(9) MWD_KG(10) MWD_KG(11) MWD_KG(12) MWD_KG(13) MWD_KG(14) MWD_KG(15) MWD_KG(16)
#undef MWD_KG
      default:
        set_error("n_concepts %d needs KG=%d > 16", p->n_concepts, pl.KG);
        return 2;
    }
    if (rc) return rc;
    if (!ll_only) {
      CountsArgs c;
      c.phone_off = p->phone_off;
      c.trans = a.trans;
      c.stats = p->stats;
      c.slot_off = p->slot_off;
      c.part_init = p->part_init;
      c.part_trans = p->part_trans;
      c.lo = lo;
      c.hi = hi;
      c.n = n;
      c.eps = a.eps;
      c.compact = use_warp32 ? 1 : 0;
      c.K = p->n_concepts;
      c.total_warps = estep_grid_rows() * 8;     // one partial row per CTA, 8 warps per CTA
      switch (n) {
        case 1: ik_counts_small_kernel<1><<<estep_grid_rows(), 256, 0, st>>>(c); break;
        case 2: ik_counts_small_kernel<2><<<estep_grid_rows(), 256, 0, st>>>(c); break;
        case 3: ik_counts_small_kernel<3><<<estep_grid_rows(), 256, 0, st>>>(c); break;
        case 4: ik_counts_small_kernel<4><<<estep_grid_rows(), 256, 0, st>>>(c); break;
        case 5: ik_counts_small_kernel<5><<<estep_grid_rows(), 256, 0, st>>>(c); break;
        case 6: ik_counts_small_kernel<6><<<estep_grid_rows(), 256, 0, st>>>(c); break;
        default: ik_counts_kernel<<<estep_grid_rows(), 256, 0, st>>>(c); break;
      }
      MWD_CHECK_LAUNCH();
    }
  }
  return 0;
}

extern "C" int mwd_ik_estep(const mwd_ik_problem* p, void* stream) { return estep_impl(p, stream, 0); }

extern "C" int mwd_ik_loglik(const mwd_ik_problem* p, void* stream) { return estep_impl(p, stream, 1); }

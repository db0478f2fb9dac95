// K1w -- warp-per-pair form of the fused forward / backward / expected-count kernel of the
// (region i, concept k)-state HMM (same reference lines as ik_estep.cu:
//   hmm_dnn/image_phone_hmm_word_discoverer.py forward :276-304, backward :314-335,
//   updateStateCounts :426-433, computeAvgLogLikelihood :523-531, phoneCounts / conceptCountsA
//   :233,:235; the init / transition counts are finished by ik_counts_*_kernel from the row
//   statistics published here).
//
// Why a second kernel: the CTA-per-4-pairs kernel needs one __syncthreads per time step and moves
// every cross-region quantity through shared memory; ncu shows it bound by the LSU data pipe
// (60 % busy) and barrier latency, FP64 pipe 19 %.  Here ONE WARP owns one caption-image pair:
//   * lane = (region i, sub-lane j), LPR = 32 / n lanes per region row, lane owns concepts
//     k = j + LPR * q (q < KG); the whole (i,k) lattice of a step is 32 x KG registers;
//   * row sums (s_t, r_t, floor-sum, xi diagonal) are shuffle trees inside the row, the
//     cross-region coupling c_t = Aoff^T s_t / w_t = Aoff r_t is n broadcast shuffles -- no block
//     barrier, no shared-memory exchange, warps never wait for each other;
//   * alpha is checkpointed every B steps to an L2-resident per-warp scratch ([q][lane] layout,
//     one coalesced 256-byte store per register) and the B-1 slices in between are recomputed
//     into the warp's private shared-memory block on the way back;
//   * the only shared-memory transpose left is the column sum  sum_i gamma_t[i][k]  (phone counts,
//     conceptCountsA): rows are written with a stride == LPR (mod 16) doubles, so both the
//     row-owner stores and the column-owner loads are bank-conflict free;
//   * phone counts go to a per-WARP table (L2-resident, fixed warp -> pair map, column k is always
//     touched by the same lane in t order), so the result is reproducible without atomics.
#include <stdlib.h>

#include "ik_estep.cuh"

namespace mwd {

constexpr int kWpc = 4;            // warps per CTA
constexpr int kWarpCtasPerSm = 3;  // 12 warps per SM (register budget 65536 / 384 = 170)
// concepts per lane above which the register-resident lattice needs the 255-register budget of
// 2 CTAs (8 warps) per SM: n = 7..10 at K = 65, n = 5..6 at K = 100
constexpr int kWarpBigKG = 14;
constexpr int warp_ctas_per_sm(int KG) { return KG >= kWarpBigKG ? 2 : kWarpCtasPerSm; }
constexpr int kWBmax = 8;          // max checkpoint interval

// L2 residency hints: the alpha checkpoints are the only global data with reuse (written in the
// forward sweep, read back once in the backward sweep) -> evict_last; the per-step row statistics
// are written once for the count post-pass and pz / cA stream through -> evict_first (.cs).
__device__ __forceinline__ uint64_t l2_policy_keep() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void st_keep(double* p, double v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}

constexpr bool is_pow2(int v) { return (v & (v - 1)) == 0; }
constexpr int pow2_below(int v) { int p = 1; while (p * 2 < v) p *= 2; return p; }   // largest 2^e < v (v >= 2)

// sum over the LPR lanes of a region row; the total is valid in the row's first lane (j == 0)
template <int LPR>
__device__ __forceinline__ double row_sum_head(double v, int j) {
  if constexpr (LPR == 1) {
    return v;
  } else if constexpr (is_pow2(LPR)) {
#pragma unroll
    for (int off = LPR / 2; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    return v;
  } else {
    constexpr int P2 = pow2_below(LPR);
    double u = __shfl_down_sync(0xffffffffu, v, P2);
    if (j + P2 < LPR) v += u;
#pragma unroll
    for (int off = P2 / 2; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    return v;
  }
}

// OBS_S: the emission table obsT (P x K doubles) is staged in shared memory once per CTA (the L1
// left beside a large shared-memory carve-out is too small to keep it resident).
// RB ("register block"): the checkpoint interval is fixed to 2 and the one recomputed alpha slice
// stays in registers, so alpha never touches shared memory (only the gamma slice does, for the
// column sums) and o_{t+1} is loaded once for the recompute and the backward step.
// GEN ("generic width"): the instantiation is wider than the concept count needs (KG > ceil(K / LPR)), so ANY concept
// group of a lane can lie beyond K, not just the last one; validity comes from per-lane bit masks.  Lets every
// K <= LPR * KG run on the warp kernel instead of dropping to the CTA-per-4-pairs kernel (ik_estep.cu).
template <int N, int KG, bool OBS_S, bool RB, bool GEN = false>
__global__ void __launch_bounds__(kWpc * 32, warp_ctas_per_sm(KG)) ik_estep_warp_kernel(const EstepArgs a) {
  constexpr int LPR = 32 / N;
  constexpr int ROWL = N * LPR;                                  // lanes that own lattice rows
  constexpr int KS0 = LPR * KG;                                  // concepts per row incl. padding
  constexpr int KS = KS0 + (((LPR - KS0) % 16) + 16) % 16;       // smem row stride == LPR (mod 16)
  constexpr int SL = KG * 32;                                    // doubles per checkpoint slice
  constexpr int KC = (KS0 + 31) / 32;                            // columns per lane in the column pass
  const int K = a.K;
  const int B = RB ? 2 : a.B;
  const double eps = a.eps;
  const int lane = threadIdx.x & 31;
  const int wic = threadIdx.x >> 5;
  const int gw = blockIdx.x * kWpc + wic;
  const int total_warps = gridDim.x * kWpc;
  const bool on = lane < ROWL;
  const int i = on ? lane / LPR : 0;
  const int j = on ? lane - i * LPR : 0;
  const bool head = on && j == 0;
  const bool kv_last = on && (j + LPR * (KG - 1) < K);           // validity of the lane's last concept
  unsigned ldmask = 0;                                           // GEN: bit q = column j + LPR q lies inside the table row
  if constexpr (GEN) {
#pragma unroll
    for (int q = 0; q < KG; ++q) ldmask |= (j + LPR * q < K) ? (1u << q) : 0u;
  }
  const unsigned kvmask = on ? ldmask : 0u;                      // GEN: bit q = concept j + LPR q of this lane exists
  auto kvalid = [&](int q) -> bool {
    if constexpr (GEN) return (kvmask >> q) & 1u;
    else return (q < KG - 1) ? on : kv_last;
  };
  auto ldvalid = [&](int q) -> bool {
    if constexpr (GEN) return (ldmask >> q) & 1u;
    else return q < KG - 1 || kv_last;
  };

  extern __shared__ double smem[];
  const int obs_elems = OBS_S ? ((a.P * K + 1) & ~1) : 0;
  // per warp: [B][N][KS] alpha / gamma block (RB: two gamma slices used alternately)
  double* buf = smem + obs_elems + (size_t)wic * B * (N * KS);
  if (OBS_S) {
    for (int e = threadIdx.x; e < a.P * K; e += blockDim.x) smem[e] = a.obsT[e];
    __syncthreads();
  }
  double* my_buf = buf + i * KS + j;                             // + tt*N*KS + LPR*q

  const double d_i = a.trans[i * N + i];
  const double pi_i = a.init[i];

  double* scr = a.scratch + (size_t)gw * a.cta_scratch;          // [NC][SL] checkpoints, [Tmax][N] c_t
  double* my_ckpt = scr + lane;                                  // + c*SL + 32*q
  double* hist = scr + (size_t)a.NC * SL + i;                    // + t*N
  // part_phone == NULL (dense-emission classes): no phone table, the caller consumes cA_out instead
  const bool tab_on = a.part_phone != nullptr;
  double* tab = a.part_phone + (size_t)gw * a.P * K;
  const double* obs_j = (OBS_S ? smem : a.obsT) + j;             // + x*K + LPR*q

  const uint64_t keep = l2_policy_keep();

  auto load_obs = [&](double (&o)[KG], int x) {
    const double* orow = obs_j + x * K;
#pragma unroll
    for (int q = 0; q < KG; ++q)
      o[q] = ldvalid(q) ? (OBS_S ? orow[LPR * q] : __ldg(orow + LPR * q)) : 0.0;
  };

  for (int64_t pair = a.lo + gw; pair < a.hi; pair += total_warps) {
    const int64_t p0 = a.phone_off[pair];
    const int T = a.phone_off[pair + 1] - (int32_t)p0;
    const int64_t r0 = a.region_off[pair];
    const int32_t* ph = a.phones + p0;
    double* st = a.stats + 4 * a.slot_off[pair] + i;             // + (t*4 + q)*N
    if (T <= 0) continue;

    double pz[KG];
    {
      const double* prow = a.pz + (r0 + i) * K + j;
#pragma unroll
      for (int q = 0; q < KG; ++q) pz[q] = kvalid(q) ? __ldcs(prow + LPR * q) : 0.0;
    }

    // ------------------------------------------------------------------ forward sweep
    double al[KG];
    double inorm = 0.0;
    {
      double acol[N];                                            // Aoff[:, i]
#pragma unroll
      for (int jp = 0; jp < N; ++jp) acol[jp] = (jp == i) ? 0.0 : a.trans[jp * N + i];
      int xn = 0;
      {
        double o[KG];
        load_obs(o, ph[0]);
#pragma unroll
        for (int q = 0; q < KG; ++q) al[q] = (pi_i * pz[q]) * o[q];
        if (T > 1) xn = ph[1];
      }
      int to_ckpt = 0;     // steps until the next checkpoint
      int cidx = 0;
      for (int t = 0; t < T; ++t) {
        double onext[KG];
        if (t + 1 < T) {   // next step's emissions, issued before the reductions
          load_obs(onext, xn);
          if (t + 2 < T) xn = ph[t + 2];
        }
        if (!a.ll_only) {
          if (to_ckpt == 0) {
            double* dst = my_ckpt + (size_t)cidx * SL;
#pragma unroll
            for (int q = 0; q < KG; ++q) st_keep(dst + 32 * q, al[q], keep);
            to_ckpt = B;
            ++cidx;
          }
          --to_ckpt;
        }
        double s = 0.0, s_b = 0.0;
#pragma unroll
        for (int q = 0; q < KG; ++q) {
          if (q & 1) s_b += al[q];
          else s += al[q];
        }
        s = row_sum_head<LPR>(s + s_b, j);
        double sv[N];
#pragma unroll
        for (int jp = 0; jp < N; ++jp) sv[jp] = __shfl_sync(0xffffffffu, s, jp * LPR);
        if (t == T - 1) {
          double L = 0.0;
#pragma unroll
          for (int jp = 0; jp < N; ++jp) L += sv[jp];
          L = floor_at(L, eps);
          if (lane == 0) a.pair_ll[pair] = log(L);                       // :529
        // s_{T-1} is never used by the counts; a negative sentinel marks the pair's last row for the
        // row-parallel count post-pass (ik_counts_small_kernel)
        if (head && !a.ll_only) __stcs(st + (t * 4 + 0) * N, -1.0);
          // sum_{i,k} alpha_t beta_t equals the sentence likelihood at every t, so the floored
          // normaliser of updateStateCounts (:430) is one constant per pair
          inorm = 1.0 / L;
        } else {
          double c = 0.0;
#pragma unroll
          for (int jp = 0; jp < N; ++jp) c = fma(acol[jp], sv[jp], c);
          if (head && !a.ll_only) {
            __stcg(hist + t * N, c);
            __stcs(st + (t * 4 + 0) * N, s);
          }
#pragma unroll
          for (int q = 0; q < KG; ++q) al[q] = onext[q] * fma(d_i, al[q], c * pz[q]);
        }
      }
    }
    if (a.ll_only) continue;

    // ------------------------------------------------------------------ backward sweep
    double arow[N];                                              // Aoff[i, :]
#pragma unroll
    for (int jp = 0; jp < N; ++jp) arow[jp] = (jp == i) ? 0.0 : a.trans[i * N + jp];
    double bo[KG];          // beta_{t+1} * o_{t+1}
#pragma unroll
    for (int q = 0; q < KG; ++q) bo[q] = 0.0;
    double w = 1.0;         // (Aoff r_{t+1})[i]; with bo = 0 the first step gets beta_{T-1} = fma(d, 0, 1) = 1

    // one backward step: alpha_t in av, emissions o_t in o; gamma slice goes to gslice (shared)
    auto bwd_step = [&](int t, int x, const double (&av)[KG], const double (&o)[KG], double* gslice) {
      // phone-count cells of this step (column owner = lane, fixed), loaded early
      double tabv[KC];
      double* trow = tab + x * K + lane;
#pragma unroll
      for (int m = 0; m < KC; ++m) tabv[m] = (tab_on && lane + 32 * m < K) ? __ldcg(trow + 32 * m) : 0.0;
      double* grow = gslice + i * KS + j;
      double sumF = 0.0, dg = 0.0, rr = 0.0, sumF_b = 0.0, dg_b = 0.0, rr_b = 0.0;   // two chains each
#pragma unroll
      for (int q = 0; q < KG; ++q) {
        const bool kv = kvalid(q);
        const double beta = fma(d_i, bo[q], w);
        const double g = av[q] * beta;
        const double f = kv ? floor_at(g, eps) : 0.0;
        if (q & 1) {
          dg_b = fma(av[q], bo[q], dg_b);
          sumF_b += f;
        } else {
          dg = fma(av[q], bo[q], dg);
          sumF += f;
        }
        bo[q] = beta * o[q];
        if (q & 1) rr_b = fma(bo[q], pz[q], rr_b);
        else rr = fma(bo[q], pz[q], rr);
        if (on) grow[LPR * q] = g;
      }
      sumF += sumF_b;
      rr += rr_b;
      dg = (dg + dg_b) * d_i;
      __syncwarp();       // gamma slice visible to the column owners
      // column loads first, the row reductions overlap their latency
      const double* col = gslice + lane;
      double cs[KC];
#pragma unroll
      for (int m = 0; m < KC; ++m) {
        cs[m] = 0.0;
        if (lane + 32 * m < K) {
#pragma unroll
          for (int ii = 0; ii < N; ++ii) cs[m] += col[ii * KS + 32 * m];
        }
      }
      sumF = row_sum_head<LPR>(sumF, j);
      dg = row_sum_head<LPR>(dg, j);
      rr = row_sum_head<LPR>(rr, j);
      if (head) {         // row statistics of this step for the count post-pass
        __stcs(st + (t * 4 + 1) * N, sumF);
        __stcs(st + (t * 4 + 2) * N, dg);
        __stcs(st + (t * 4 + 3) * N, rr);
      }
      double wn = 0.0;
#pragma unroll
      for (int jp = 0; jp < N; ++jp) wn = fma(arow[jp], __shfl_sync(0xffffffffu, rr, jp * LPR), wn);
      w = wn;
      // conceptCountsA[t][k] = sum_i gamma_t[i][k]; phoneCounts[k][x_t] += ...  (:430,:233,:235)
#pragma unroll
      for (int m = 0; m < KC; ++m) {
        cs[m] *= inorm;
        if (tab_on && lane + 32 * m < K) __stcg(trow + 32 * m, tabv[m] + cs[m]);
      }
      if (a.cA_out) {
        double* crow = a.cA_out + (p0 + t) * K + lane;
#pragma unroll
        for (int m = 0; m < KC; ++m)
          if (lane + 32 * m < K) crow[32 * m] = cs[m];
      }
      if (a.ca_out) {     // concept_alignment[t] = argmax_k cA[t][k] (first index on ties, :628)
        double bv = 0.0;
        int bk = 0x7fffffff;
#pragma unroll
        for (int m = 0; m < KC; ++m)
          if (lane + 32 * m < K && (bk == 0x7fffffff || argmax_better(cs[m], bv))) { bv = cs[m]; bk = lane + 32 * m; }
        const int kbest = warp_argmax_nonneg(bv, bk);
        if (lane == 0) a.ca_out[p0 + t] = kbest;
      }
    };
    auto drop_slice = [&](int c) {
      // checkpoint slice c is dead: drop it from L2 instead of letting it be written back
      for (int ln = lane; ln < SL / 16; ln += 32) {
        const double* dead = scr + (size_t)c * SL + ln * 16;
        asm volatile("discard.global.L2 [%0], 128;" ::"l"(dead) : "memory");
      }
    };

    const int nblk = (T + B - 1) / B;
    if constexpr (RB) {
      int par = 0;
      for (int c = nblk - 1; c >= 0; --c) {
        const int t0 = 2 * c;
        double a0[KG];
        {
          const double* src = my_ckpt + (size_t)c * SL;
#pragma unroll
          for (int q = 0; q < KG; ++q) a0[q] = __ldcg(src + 32 * q);
        }
        const int x0 = ph[t0];
        if (t0 + 1 < T) {
          const int x1 = ph[t0 + 1];
          const double cb = __ldcg(hist + t0 * N);
          double o1[KG], a1[KG];
          load_obs(o1, x1);
#pragma unroll
          for (int q = 0; q < KG; ++q) a1[q] = o1[q] * fma(d_i, a0[q], cb * pz[q]);
          bwd_step(t0 + 1, x1, a1, o1, buf + par * (N * KS));
          par ^= 1;
        }
        {
          double o0[KG];
          load_obs(o0, x0);
          bwd_step(t0, x0, a0, o0, buf + par * (N * KS));
          par ^= 1;
        }
        drop_slice(c);
      }
    } else {
      int xb = ph[T - 1];     // phone of the step about to be processed
      for (int c = nblk - 1; c >= 0; --c) {
        const int t0 = c * B;
        const int len = min(B, T - t0);
        __syncwarp();         // column reads of the previous block are complete
        {
          const double* src = my_ckpt + (size_t)c * SL;
#pragma unroll
          for (int q = 0; q < KG; ++q) al[q] = __ldcg(src + 32 * q);
          if (on) {
#pragma unroll
            for (int q = 0; q < KG; ++q) my_buf[LPR * q] = al[q];
          }
          for (int tt = 1; tt < len; ++tt) {
            const double cb = __ldcg(hist + (t0 + tt - 1) * N);
            double o[KG];
            load_obs(o, ph[t0 + tt]);
            double* dst = my_buf + tt * (N * KS);
#pragma unroll
            for (int q = 0; q < KG; ++q) {
              al[q] = o[q] * fma(d_i, al[q], cb * pz[q]);
              if (on) dst[LPR * q] = al[q];
            }
          }
        }
        for (int tt = len - 1; tt >= 0; --tt) {
          const int t = t0 + tt;
          const int x = xb;
          if (t > 0) xb = ph[t - 1];
          double av[KG], o[KG];
          const double* row = my_buf + tt * (N * KS);
#pragma unroll
          for (int q = 0; q < KG; ++q) av[q] = row[LPR * q];
          load_obs(o, x);
          bwd_step(t, x, av, o, buf + tt * (N * KS));
        }
        drop_slice(c);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct WarpPlan {
  int KG, B, NC, grid, obs_s, rb;
  bool gen;               // generic-width instantiation (KG wider than the concept count needs)
  size_t smem;
  int64_t warp_scratch;   // doubles per warp
};

// (n, KG) instantiations: K = 65 (MSCOCO, run_image2phone.py:43) and K = 50 / 100 (Flickr30k,
// run_image2phone.py:73 / image_phone_hmm_word_discoverer.py:735) for the n they are fast for, plus K = 40 / 80.
// Other concept counts run a generic-width instantiation (MWD_WARP_GEN_COMBOS) where one is wide enough, else the
// CTA-per-4-pairs kernel (ik_estep.cu).
#define MWD_WARP_COMBOS(X)                                                              \
  X(1, 2) X(1, 3) X(1, 4) X(2, 4) X(2, 5) X(2, 7) X(3, 5) X(3, 7) X(3, 10) X(4, 7) X(4, 9) \
  X(4, 13) X(5, 9) X(5, 11) X(6, 10) X(6, 13)                                             \
  X(5, 17) X(7, 13) X(7, 17) X(8, 13) X(8, 17) X(9, 17) X(9, 22) X(10, 17) X(10, 22)                 \
  X(6, 20) X(7, 25) X(8, 25)                                                                          \
  /* K = 40 and K = 80 (any n <= 8, K = 40 also n = 9, 10): no drop to the CTA-per-4-pairs kernel */ \
  X(2, 3) X(3, 4) X(4, 5) X(5, 7) X(6, 8) X(7, 10) X(8, 10) X(9, 14) X(10, 14)                       \
  X(3, 8) X(4, 10) X(5, 14) X(6, 16) X(7, 20) X(8, 20)
// register-block variant: compiled (and the default) for the MSCOCO concept count K = 65, n <= 8 (at
// n = 5 it measured 55.1 vs 60.3 ms at 1M pairs); MWD_ESTEPW_RB=0 selects the shared-memory block
// variant instead
#define MWD_WARP_RB_COMBOS(X) X(1, 3) X(2, 5) X(3, 7) X(4, 9) X(5, 11) X(6, 13) X(7, 17) X(8, 17)

// generic-width instantiations (shared-memory-block variant): every exact width again with the per-group validity
// mask, plus the widths that take n <= 4 up to K = 128; wider float64 lattices than these fall back to ik_estep.cu
#define MWD_WARP_GEN_COMBOS(X) MWD_WARP_COMBOS(X) X(2, 8) X(3, 13) X(4, 16)

static int warp_gen_width(int n, int KG) {
  int best = 0;
#define X(NN, GG) if (n == NN && GG >= KG && (best == 0 || GG < best)) best = GG;
  MWD_WARP_GEN_COMBOS(X)
#undef X
  return best;
}

static int warp_kg(int n, int K) {
  const int lpr = 32 / n;
  return (K + lpr - 1) / lpr;
}

static bool warp_combo(int n, int KG) {
#define X(NN, GG) if (n == NN && KG == GG) return true;
  MWD_WARP_COMBOS(X)
#undef X
  return false;
}

static bool warp_enabled() {
  const char* e = getenv("MWD_ESTEP_WARP");
  return !(e && atoi(e) == 0);
}

static bool plan_warp(int n, int K, int P, int Tmax, int64_t npairs, WarpPlan* pl) {
  if (n < 1 || n > 10 || !warp_enabled()) return false;
  const int lpr = 32 / n;
  pl->KG = warp_kg(n, K);
  pl->gen = false;
  if (!warp_combo(n, pl->KG)) {
    pl->KG = warp_gen_width(n, pl->KG);
    if (pl->KG == 0) return false;
    pl->gen = true;
  }
  const int ks0 = lpr * pl->KG;
  const int ks = ks0 + (((lpr - ks0) % 16) + 16) % 16;
  const size_t slice = (size_t)kWpc * n * ks * sizeof(double);     // one alpha slice of every warp of a CTA
  const int cps = warp_ctas_per_sm(pl->KG);
  size_t budget = (size_t)224 * 1024 / cps - 1024;
  // emission table in shared memory when it leaves room for a checkpoint interval of >= 3
  const size_t obs_bytes = (((size_t)P * K + 1) & ~(size_t)1) * sizeof(double);
  pl->obs_s = (obs_bytes + 3 * slice <= budget) ? 1 : 0;
  if (const char* e = getenv("MWD_ESTEPW_OBS")) pl->obs_s = (atoi(e) != 0 && obs_bytes + 2 * slice <= budget) ? 1 : 0;
  if (pl->obs_s) budget -= obs_bytes;
  pl->rb = 0;
#define X(NN, GG) if (n == NN && pl->KG == GG && !pl->gen) pl->rb = 1;
  MWD_WARP_RB_COMBOS(X)
#undef X
  if (const char* e = getenv("MWD_ESTEPW_RB")) { if (atoi(e) == 0) pl->rb = 0; }
  if (pl->rb) {   // register block: B = 2, two gamma slices per warp, table in shared memory if it fits
    pl->obs_s = (obs_bytes + 2 * slice <= (size_t)224 * 1024 / cps - 1024) ? 1 : 0;
    budget = 2 * slice;
  }
  int B = (int)(budget / slice);
  if (const char* e = getenv("MWD_ESTEPW_B")) { int v = atoi(e); if (v >= 1 && v < B) B = v; }
  if (B > kWBmax) B = kWBmax;
  if (B > Tmax) B = Tmax > 0 ? Tmax : 1;
  if (B < 2 && Tmax > 1) return false;
  pl->B = B;
  pl->NC = (Tmax + B - 1) / B;
  if (pl->NC < 1) pl->NC = 1;
  pl->smem = (size_t)B * slice + (pl->obs_s ? obs_bytes : 0);
  int ctas_per_sm = cps;
  if (const char* e = getenv("MWD_ESTEPW_CTAS")) { int v = atoi(e); if (v >= 1 && v < ctas_per_sm) ctas_per_sm = v; }
  int64_t grid = (int64_t)sm_count() * ctas_per_sm;
  const int64_t need = (npairs + kWpc - 1) / kWpc;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  if (grid * kWpc > estep_grid_rows()) return false;
  pl->grid = (int)grid;
  pl->warp_scratch = (((int64_t)pl->NC * pl->KG * 32 + (int64_t)Tmax * n) + 15) & ~(int64_t)15;
  return true;
}

bool estep_warp_supported(int n, int K, int P) {
  WarpPlan pl;
  return plan_warp(n, K, P, 64, 1 << 20, &pl);
}

int64_t estep_warp_scratch(int n, int K, int P, int Tmax, int64_t npairs) {
  WarpPlan pl;
  if (!plan_warp(n, K, P, Tmax, npairs, &pl)) return 0;
  return pl.warp_scratch * pl.grid * kWpc;
}

template <int N, int KG, bool OBS_S, bool RB, bool GEN = false>
static int launch_warp(const EstepArgs& a, const WarpPlan& pl, cudaStream_t st) {
  auto kern = ik_estep_warp_kernel<N, KG, OBS_S, RB, GEN>;
  MWD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  kern<<<pl.grid, kWpc * 32, pl.smem, st>>>(a);
  MWD_CHECK_LAUNCH();
  return 0;
}

int estep_warp_launch(EstepArgs a, cudaStream_t st) {
  WarpPlan pl;
  MWD_REQUIRE(plan_warp(a.n, a.K, a.P, a.Tmax, a.hi - a.lo, &pl), "warp E-step: unsupported (n=%d, K=%d)", a.n, a.K);
  a.B = pl.B;
  a.NC = pl.NC;
  a.cta_scratch = pl.warp_scratch;
  if (pl.gen) {
#define X(NN, GG)                                                                     \
  if (a.n == NN && pl.KG == GG)                                                       \
    return pl.obs_s ? launch_warp<NN, GG, true, false, true>(a, pl, st) : launch_warp<NN, GG, false, false, true>(a, pl, st);
    MWD_WARP_GEN_COMBOS(X)
#undef X
    set_error("warp E-step: no generic-width instantiation for (n=%d, KG=%d)", a.n, pl.KG);
    return 2;
  }
  if (pl.rb) {
#define X(NN, GG)                                                                     \
  if (a.n == NN && pl.KG == GG)                                                       \
    return pl.obs_s ? launch_warp<NN, GG, true, true>(a, pl, st) : launch_warp<NN, GG, false, true>(a, pl, st);
    MWD_WARP_RB_COMBOS(X)
#undef X
  }
#define X(NN, GG)                                                                     \
  if (a.n == NN && pl.KG == GG)                                                       \
    return pl.obs_s ? launch_warp<NN, GG, true, false>(a, pl, st) : launch_warp<NN, GG, false, false>(a, pl, st);
  MWD_WARP_COMBOS(X)
#undef X
  set_error("warp E-step: no instantiation for (n=%d, KG=%d)", a.n, pl.KG);
  return 2;
}

}  // namespace mwd

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
//   * counts are accumulated without atomics: init/trans in registers of fixed threads, phone
//     counts by read-modify-write of a per-CTA table that only this CTA touches, in t order.
#include "mwd_common.cuh"

namespace mwd {

struct EstepArgs {
  const int32_t* region_off;
  const int32_t* phone_off;
  const int32_t* phones;
  const double* pz;
  const double* init;    // init[n] row
  const double* trans;   // trans[n] table, [i*n+j]
  const double* obsT;
  double* pair_ll;
  double* cA_out;        // may be null
  double* part_phone;    // [grid][P*K]
  double* part_init;     // [grid][(NMAX+1)*NMAX]
  double* part_trans;    // [grid][(NMAX+1)*NMAX*NMAX]
  double* scratch;
  int64_t lo, hi;
  int64_t cta_scratch;   // doubles per CTA
  int n, K, P, B, NC, Tmax;
  int ll_only;           // 1: forward sweep + log-likelihood only
};

constexpr int NQ = 5;  // exchanged per-row quantities: floor-sum, g-sum, diag, r, xi-row-sum

template <int KG>
__global__ void __launch_bounds__(kPairsPerCta * kNMax * kLanesPerRow)
ik_estep_kernel(const EstepArgs a) {
  constexpr int KS = KG * kLanesPerRow;
  constexpr int PP = kPairsPerCta;
  const int n = a.n, K = a.K, B = a.B;
  const int tid = threadIdx.x;
  const int l8 = tid & 7;
  const int grp = tid >> 3;
  const int slot = grp / n;
  const int i = grp - slot * n;
  constexpr int NJ = kNMax / kLanesPerRow;  // xi columns per lane

  extern __shared__ double smem[];
  double* s_buf = smem;                                   // [PP][B][n][KS]
  double* s_exch = s_buf + (size_t)PP * B * n * KS;       // [2][PP][NQ][NMAX]
  double* s_aoff = s_exch + 2 * PP * NQ * kNMax;          // [n][n]
  double* s_d = s_aoff + kNMax * kNMax;                   // [n]
  double* s_pi = s_d + kNMax;                             // [n]
  double* s_cA = s_pi + kNMax;                            // [2][PP][K] deferred phone-count rows
  __shared__ int s_T[PP];
  __shared__ int s_xs[2][PP];                             // phone id of the deferred row, -1 = none

  // transition / initial tables of this n
  for (int e = tid; e < n * n; e += blockDim.x) {
    int r = e / n, c = e - r * n;
    double v = a.trans[e];
    s_aoff[e] = (r == c) ? 0.0 : v;
    if (r == c) s_d[r] = v;
  }
  for (int e = tid; e < n; e += blockDim.x) s_pi[e] = a.init[e];

  double* cta_scr = a.scratch + (size_t)blockIdx.x * a.cta_scratch;
  double* ckpt = cta_scr;                                           // [PP][NC][n][KS]
  double* hist = cta_scr + (size_t)PP * a.NC * n * KS;              // [PP][Tmax][2][n]
  double* my_ckpt = ckpt + ((size_t)slot * a.NC * n + i) * KS;      // + c*n*KS + k
  double* my_hist = hist + (size_t)slot * a.Tmax * 2 * n;           // + (t*2+which)*n + i
  double* my_buf = s_buf + ((size_t)slot * B * n + i) * KS;         // + tt*n*KS + k
  double* my_phone = a.part_phone + (size_t)blockIdx.x * a.P * K;

  // persistent accumulators
  double init_acc = 0.0;          // lane 0 of each row: sum_t floor-sum_i / total
  double trans_acc[NJ];           // lane l of row i: xi[i][l + 8*jj]
#pragma unroll
  for (int jj = 0; jj < NJ; ++jj) trans_acc[jj] = 0.0;

  const int64_t npairs = a.hi - a.lo;
  const int64_t nquads = (npairs + PP - 1) / PP;
  __syncthreads();
  const double d_i = s_d[i];
  const double pi_i = s_pi[i];

  for (int64_t quad = blockIdx.x; quad < nquads; quad += gridDim.x) {
    const int64_t pair = a.lo + quad * PP + slot;
    const bool valid = pair < a.hi;
    int T = 0;
    const int32_t* ph = a.phones;
    int64_t r0 = 0, p0 = 0;
    if (valid) {
      p0 = a.phone_off[pair];
      T = a.phone_off[pair + 1] - (int32_t)p0;
      ph = a.phones + p0;
      r0 = a.region_off[pair];
    }
    __syncthreads();  // previous quad fully done (exchange + buffers reusable)
    if (i == 0 && l8 == 0) {
      s_T[slot] = T;
      s_xs[0][slot] = -1;
      s_xs[1][slot] = -1;
    }
    __syncthreads();
    int Tmax = 0;
#pragma unroll
    for (int s = 0; s < PP; ++s) Tmax = max(Tmax, s_T[s]);

    double pz[KG];
#pragma unroll
    for (int j = 0; j < KG; ++j) {
      int k = l8 + 8 * j;
      pz[j] = (valid && k < K) ? a.pz[(r0 + i) * K + k] : 0.0;
    }

    // ------------------------------------------------------------------ forward sweep
    double al[KG];
    {
      int x = valid ? ph[0] : 0;
      const double* orow = a.obsT + (size_t)x * K;
#pragma unroll
      for (int j = 0; j < KG; ++j) {
        int k = l8 + 8 * j;
        double o = (k < K) ? __ldg(orow + k) : 0.0;
        al[j] = (pi_i * pz[j]) * o;
      }
    }
    for (int t = 0; t < Tmax; ++t) {
      const bool act = t < T;
      const int par = t & 1;
      double s = 0.0;
      if (act) {
        if (t % B == 0 && !a.ll_only) {
          double* dst = my_ckpt + (size_t)(t / B) * n * KS;
#pragma unroll
          for (int j = 0; j < KG; ++j) __stcg(dst + l8 + 8 * j, al[j]);
        }
#pragma unroll
        for (int j = 0; j < KG; ++j) s += al[j];
      }
      s = row8_sum(s);
      if (l8 == 0) s_exch[((par * PP + slot) * NQ + 0) * kNMax + i] = s;
      __syncthreads();
      if (act) {
        const double* ex = s_exch + ((par * PP + slot) * NQ + 0) * kNMax;
        if (t == T - 1) {
          if (i == 0 && l8 == 0) {
            double L = 0.0;
            for (int j = 0; j < n; ++j) L += ex[j];
            a.pair_ll[pair] = log(floor_eps(L));
          }
        } else {
          double c = 0.0;
          for (int j = 0; j < n; ++j) c = fma(s_aoff[j * n + i], ex[j], c);
          if (l8 == 0 && !a.ll_only) {
            __stcg(my_hist + (size_t)(t * 2 + 0) * n + i, c);
            __stcg(my_hist + (size_t)(t * 2 + 1) * n + i, s);
          }
          int x = ph[t + 1];
          const double* orow = a.obsT + (size_t)x * K;
#pragma unroll
          for (int j = 0; j < KG; ++j) {
            int k = l8 + 8 * j;
            double o = (k < K) ? __ldg(orow + k) : 0.0;
            al[j] = o * fma(d_i, al[j], c * pz[j]);
          }
        }
      }
    }

    if (a.ll_only) continue;

    // ------------------------------------------------------------------ backward sweep
    double bo[KG];          // beta_{t+1} * o_{t+1}
#pragma unroll
    for (int j = 0; j < KG; ++j) bo[j] = 0.0;
    double w = 0.0;         // (Aoff r_{t+1})[i]
    double r_next[NJ];      // r_{t+1}[j] for this lane's xi columns
    double x_hold[NJ];      // floored xi_t[i][j] waiting for its normaliser
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) { r_next[jj] = 0.0; x_hold[jj] = 0.0; }
    bool pend = false;
    double zrow = 0.0;
    int step = 0;           // parity counter of exchange buffers (continues from forward)

    const int nblk = (Tmax + B - 1) / B;
    for (int c = nblk - 1; c >= 0; --c) {
      const int t0 = c * B;
      const int len = min(B, Tmax - t0);
      __syncthreads();  // column reads of the previous block are complete
      if (t0 < T) {
        const double* src = my_ckpt + (size_t)c * n * KS;
#pragma unroll
        for (int j = 0; j < KG; ++j) {
          al[j] = __ldcg(src + l8 + 8 * j);
          my_buf[l8 + 8 * j] = al[j];
        }
        for (int tt = 1; tt < len && t0 + tt < T; ++tt) {
          const int t = t0 + tt;
          double cprev = __ldcg(my_hist + (size_t)((t - 1) * 2 + 0) * n + i);
          int x = ph[t];
          const double* orow = a.obsT + (size_t)x * K;
          double* dst = my_buf + (size_t)tt * n * KS;
#pragma unroll
          for (int j = 0; j < KG; ++j) {
            int k = l8 + 8 * j;
            double o = (k < K) ? __ldg(orow + k) : 0.0;
            al[j] = o * fma(d_i, al[j], cprev * pz[j]);
            dst[l8 + 8 * j] = al[j];
          }
        }
      }
      for (int tt = len - 1; tt >= 0; --tt, ++step) {
        const int t = t0 + tt;
        const bool act = t < T;
        const int par = step & 1;
        double sumF = 0.0, sumG = 0.0, dg = 0.0, rr = 0.0, s_cur = 0.0;
        int x = 0;
        if (act) {
          x = ph[t];
          const bool last = (t == T - 1);
          if (!last) s_cur = __ldcg(my_hist + (size_t)(t * 2 + 1) * n + i);
          const double* orow = a.obsT + (size_t)x * K;
          double* row = my_buf + (size_t)tt * n * KS;
#pragma unroll
          for (int j = 0; j < KG; ++j) {
            int k = l8 + 8 * j;
            const bool kv = k < K;
            double av = row[l8 + 8 * j];
            double beta = last ? 1.0 : fma(d_i, bo[j], w);
            dg = fma(av, bo[j], dg);
            double g = av * beta;
            sumG += g;
            sumF += kv ? floor_eps(g) : 0.0;
            double o = kv ? __ldg(orow + k) : 0.0;
            bo[j] = beta * o;
            rr = fma(bo[j], pz[j], rr);
            row[l8 + 8 * j] = g;
          }
          dg *= d_i;
        }
        sumF = row8_sum(sumF);
        sumG = row8_sum(sumG);
        dg = row8_sum(dg);
        rr = row8_sum(rr);
        double zr = row8_sum(zrow);
        zrow = 0.0;
        if (l8 == 0) {
          double* ex = s_exch + ((par * PP + slot) * NQ) * kNMax + i;
          ex[0 * kNMax] = sumF;
          ex[1 * kNMax] = sumG;
          ex[2 * kNMax] = dg;
          ex[3 * kNMax] = rr;
          ex[4 * kNMax] = zr;
        }
        __syncthreads();
        // drain the phone-count rows deferred by the previous step: thread k owns column k of
        // the per-CTA table for every slot, so the read-modify-writes never race and their
        // order (t descending, slot ascending) is fixed.
        for (int k = tid; k < K; k += blockDim.x) {
#pragma unroll
          for (int s = 0; s < PP; ++s) {
            int xs = s_xs[par ^ 1][s];
            if (xs >= 0) my_phone[(size_t)xs * K + k] += s_cA[((par ^ 1) * PP + s) * K + k];
          }
        }
        if (i == 0 && l8 == 0) s_xs[par][slot] = act ? x : -1;
        if (act) {
          const double* ex = s_exch + ((par * PP + slot) * NQ) * kNMax;
          double Ft = 0.0, Gt = 0.0, Zt = 0.0, wn = 0.0;
          for (int j = 0; j < n; ++j) {
            Ft += ex[0 * kNMax + j];
            Gt += ex[1 * kNMax + j];
            Zt += ex[4 * kNMax + j];
            wn = fma(s_aoff[i * n + j], ex[3 * kNMax + j], wn);
          }
          if (l8 == 0) init_acc += sumF / Ft;                    // :355
          if (pend) {                                            // normalise xi_{t+1}  (:396)
#pragma unroll
            for (int jj = 0; jj < NJ; ++jj) trans_acc[jj] += x_hold[jj] / Zt;
            pend = false;
          }
          if (t < T - 1) {                                       // xi_t  (:388-389)
            zrow = 0.0;
#pragma unroll
            for (int jj = 0; jj < NJ; ++jj) {
              int j = l8 + 8 * jj;
              double xv = 0.0;
              if (j < n) {
                double xi = (j == i) ? ex[2 * kNMax + i] : (s_cur * s_aoff[i * n + j]) * r_next[jj];
                xv = floor_eps(xi);
              }
              x_hold[jj] = xv;
              zrow += xv;
            }
            pend = true;
          }
#pragma unroll
          for (int jj = 0; jj < NJ; ++jj) {
            int j = l8 + 8 * jj;
            r_next[jj] = (j < n) ? ex[3 * kNMax + j] : 0.0;
          }
          w = wn;
          // conceptCountsA[t][k] = sum_i gamma_t[i][k]; phoneCounts[k][x_t] += ...  (:430,:233,:235)
          const double norm = floor_eps(Gt);
          const double* col = s_buf + ((size_t)slot * B + tt) * n * KS;
          double* crow = s_cA + (par * PP + slot) * K;
          for (int k = i * kLanesPerRow + l8; k < K; k += n * kLanesPerRow) {
            double v = 0.0;
            for (int ii = 0; ii < n; ++ii) v += col[ii * KS + k];
            v /= norm;
            crow[k] = v;
            if (a.cA_out) a.cA_out[(p0 + t) * K + k] = v;
          }
        }
      }
    }
    // flush the last pending xi normaliser (xi_0)
    {
      const int par = step & 1;
      double zr = row8_sum(zrow);
      zrow = 0.0;
      if (l8 == 0) s_exch[((par * PP + slot) * NQ + 4) * kNMax + i] = zr;
      __syncthreads();
      for (int k = tid; k < K; k += blockDim.x) {
#pragma unroll
        for (int s = 0; s < PP; ++s) {
          int xs = s_xs[par ^ 1][s];
          if (xs >= 0) my_phone[(size_t)xs * K + k] += s_cA[((par ^ 1) * PP + s) * K + k];
        }
      }
      if (pend) {
        const double* ex = s_exch + ((par * PP + slot) * NQ + 4) * kNMax;
        double Zt = 0.0;
        for (int j = 0; j < n; ++j) Zt += ex[j];
#pragma unroll
        for (int jj = 0; jj < NJ; ++jj) trans_acc[jj] += x_hold[jj] / Zt;
        pend = false;
      }
    }
  }

  if (a.ll_only) return;
  // ---------------------------------------------------------------- per-CTA partial tables
  __syncthreads();
  double* red = s_buf;  // reuse: [PP][n][n] then [PP][n]
#pragma unroll
  for (int jj = 0; jj < NJ; ++jj) {
    int j = l8 + 8 * jj;
    if (j < n) red[(slot * n + i) * n + j] = trans_acc[jj];
  }
  double* red_i = red + PP * n * n;
  if (l8 == 0) red_i[slot * n + i] = init_acc;
  __syncthreads();
  double* pt = a.part_trans + ((size_t)blockIdx.x * (kNMax + 1) + n) * (kNMax * kNMax);
  double* pi_out = a.part_init + ((size_t)blockIdx.x * (kNMax + 1) + n) * kNMax;
  for (int e = tid; e < n * n; e += blockDim.x) {
    double v = 0.0;
    for (int s = 0; s < PP; ++s) v += red[s * n * n + e];
    pt[e] += v;
  }
  for (int e = tid; e < n; e += blockDim.x) {
    double v = 0.0;
    for (int s = 0; s < PP; ++s) v += red_i[s * n + e];
    pi_out[e] += v;
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int kg_for(int K) {
  if (K <= 16) return 2;
  if (K <= 32) return 4;
  if (K <= 64) return 8;
  if (K <= 72) return 9;
  if (K <= 104) return 13;
  return 16;
}

struct EstepPlan {
  int KG, KS, B, NC, threads, grid;
  size_t smem;
  int64_t cta_scratch;  // doubles
};

static size_t estep_smem(int KS, int B, int n, int K) {
  return ((size_t)kPairsPerCta * B * n * KS + 2 * kPairsPerCta * NQ * kNMax + kNMax * kNMax +
          2 * kNMax + 2 * kPairsPerCta * K) * sizeof(double);
}

static EstepPlan plan_bucket(int n, int K, int Tmax, int64_t npairs) {
  EstepPlan pl;
  pl.KG = kg_for(K);
  pl.KS = pl.KG * kLanesPerRow;
  size_t slice = (size_t)kPairsPerCta * n * pl.KS * sizeof(double);
  int B = (int)((44 * 1024) / slice);
  if (B < 1) B = 1;
  if (B > 8) B = 8;
  if (B > Tmax) B = Tmax > 0 ? Tmax : 1;
  pl.B = B;
  pl.NC = (Tmax + B - 1) / B;
  if (pl.NC < 1) pl.NC = 1;
  pl.threads = kPairsPerCta * n * kLanesPerRow;
  pl.smem = estep_smem(pl.KS, B, n, K);
  int64_t nquads = (npairs + kPairsPerCta - 1) / kPairsPerCta;
  int per_sm = (int)((220 * 1024) / (pl.smem + 1024));
  int by_threads = 2048 / pl.threads;
  if (per_sm > by_threads) per_sm = by_threads;
  if (per_sm > kEstepCtasPerSm) per_sm = kEstepCtasPerSm;
  if (per_sm < 1) per_sm = 1;
  int64_t grid = (int64_t)sm_count() * per_sm;
  if (grid > nquads) grid = nquads;
  if (grid < 1) grid = 1;
  pl.grid = (int)grid;
  pl.cta_scratch = (int64_t)kPairsPerCta * ((int64_t)pl.NC * n * pl.KS + (int64_t)Tmax * 2 * n);
  return pl;
}

template <int KG>
static int launch_estep(const EstepArgs& args, const EstepPlan& pl, cudaStream_t st) {
  auto kern = ik_estep_kernel<KG>;
  MWD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)pl.smem));
  kern<<<pl.grid, pl.threads, pl.smem, st>>>(args);
  MWD_CHECK_LAUNCH();
  return 0;
}

}  // namespace mwd

using namespace mwd;

extern "C" int64_t mwd_ik_scratch_bytes(const mwd_ik_problem* p) {
  int64_t need = 0;
  for (int b = 0; b < p->n_buckets; ++b) {
    int64_t np_ = p->bucket_lo[b + 1] - p->bucket_lo[b];
    if (np_ <= 0) continue;
    EstepPlan pl = plan_bucket(p->bucket_n[b], p->n_concepts, p->bucket_tmax[b], np_);
    int64_t bytes = pl.cta_scratch * pl.grid * (int64_t)sizeof(double);
    if (bytes > need) need = bytes;
  }
  return need;
}

static int estep_impl(const mwd_ik_problem* p, void* stream, int ll_only) {
  MWD_REQUIRE(p->n_concepts >= 1 && p->n_concepts <= MWD_KMAX, "n_concepts %d outside [1,%d]",
              p->n_concepts, MWD_KMAX);
  cudaStream_t st = as_stream(stream);
  for (int b = 0; b < p->n_buckets; ++b) {
    const int n = p->bucket_n[b];
    const int64_t lo = p->bucket_lo[b], hi = p->bucket_lo[b + 1];
    if (hi <= lo) continue;
    MWD_REQUIRE(n >= 1 && n <= MWD_NMAX, "bucket %d: n=%d outside [1,%d]", b, n, MWD_NMAX);
    EstepPlan pl = plan_bucket(n, p->n_concepts, p->bucket_tmax[b], hi - lo);
    MWD_REQUIRE(pl.smem <= 227 * 1024, "estep shared memory %zu exceeds 227 KB (n=%d K=%d)",
                pl.smem, n, p->n_concepts);
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
    a.part_phone = p->part_phone;
    a.part_init = p->part_init;
    a.part_trans = p->part_trans;
    a.scratch = p->scratch;
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
    int rc = 0;
    switch (pl.KG) {
      case 2: rc = launch_estep<2>(a, pl, st); break;
      case 4: rc = launch_estep<4>(a, pl, st); break;
      case 8: rc = launch_estep<8>(a, pl, st); break;
      case 9: rc = launch_estep<9>(a, pl, st); break;
      case 13: rc = launch_estep<13>(a, pl, st); break;
      default: rc = launch_estep<16>(a, pl, st); break;
    }
    if (rc) return rc;
  }
  return 0;
}

extern "C" int mwd_ik_estep(const mwd_ik_problem* p, void* stream) { return estep_impl(p, stream, 0); }

extern "C" int mwd_ik_loglik(const mwd_ik_problem* p, void* stream) { return estep_impl(p, stream, 1); }

// K1w32 -- the warp-per-pair forward / backward / expected-count kernel with the (i,k) lattice in SCALED float32
// (opt-in: mwd_ik_problem.mixed_precision & MWD_MIXED_RECURSION).  Same reference lines as ik_estep_warp.cu
// (hmm_dnn/image_phone_hmm_word_discoverer.py forward :276-304, backward :314-335, updateStateCounts :426-433,
// computeAvgLogLikelihood :523-531, phoneCounts / conceptCountsA :233,:235).
//
// Why it is legitimate: every EPS floor of the reference acts on a raw probability, but each
// one can be restated on a scaled value once the common scale of the lattice at that step is known:
//     alpha_t = a^_t * 2^ea_t,   beta_t = b^_t * 2^eb_t           (ONE integer exponent per (pair, t))
//     max(alpha beta, EPS) = 2^(ea+eb) * max(a^ b^, EPS * 2^-(ea+eb))
// and if EPS * 2^-(ea+eb) is beyond float32 every entry is floored, the floored row sum is exactly K * EPS.
// The lattice values (32 x KG per step) are float32 in [2^-126, ~1] after power-of-two renormalisation (exact);
// everything that leaves the warp is float64 again: the per-step row statistics handed to the count post-pass travel
// as float32 mantissas + the step's exponents and are un-scaled to float64 by the reader (StatRow in ik_estep.cu; the
// xi / init floors are applied there on raw float64 values as before -- 92 instead of 160 bytes per step), the phone table
// accumulates  float64(sum_i gamma^) * 2^(ea+eb) / max(L, EPS)  in float64, log-likelihood in float64.
// The emission table is staged in shared memory as float32 with one power-of-two shift per phone type.
//
// What it buys on B200: the FP64 pipe issues one warp instruction per 2 cycles per scheduler, the FP32 pipe one per
// cycle; the lattice takes half the registers (5 x KG instead of 10 x KG for pz / alpha / recompute / o / beta o),
// so 4 CTAs of 4 warps fit per SM instead of 3; shuffles and shared-memory words are 32-bit; the alpha checkpoints
// are half the bytes.  Accuracy against the float64 kernel: tests/test_gpu_mixed_precision.py (1e-5 gate).
#include <stdlib.h>

#include "ik_estep.cuh"

namespace mwd {

namespace {

#ifndef MWD_W32_KEEP
#define MWD_W32_KEEP 1    // L2 evict_last policy on: bit 0 the checkpoint stores, bit 1 the phone-table updates.  Measured at
                          // 1 M MSCOCO pairs (time / DRAM bytes per launch): 0: 35.58 ms / 50.3 GB, 1: 35.15 / 42.6,
                          // 2: 35.23 / 43.2, 3: 35.19 / 42.2 -- the streamed row statistics and posteriors leave L2 first
#endif
#ifndef MWD_W32_PF
#define MWD_W32_PF 0      // checkpoint prefetch of the backward sweep: 0 off, 1 into L1, 2 into registers (measured at
                          // 1 M MSCOCO pairs: 35.49 / 35.49 / 37.42 ms -- the L2 latency of the slice is already hidden)
#endif
constexpr int kWpc32 = 4;            // warps per CTA
// CTAs per SM by lattice width: 16 warps at 128 registers up to 12 concepts per lane, 12 warps at 168 up to 17, else 8
constexpr int ctas32(int KG) { return KG <= 12 ? 4 : (KG <= 17 ? 3 : 2); }

constexpr bool is_pow2_(int v) { return (v & (v - 1)) == 0; }
constexpr int pow2_below_(int v) { int p = 1; while (p * 2 < v) p *= 2; return p; }

template <int LPR>
__device__ __forceinline__ float row_sum_head32(float v, int j) {
  if constexpr (LPR == 1) {
    return v;
  } else if constexpr (is_pow2_(LPR)) {
#pragma unroll
    for (int off = LPR / 2; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    return v;
  } else {
    constexpr int P2 = pow2_below_(LPR);
    float u = __shfl_down_sync(0xffffffffu, v, P2);
    if (j + P2 < LPR) v += u;
#pragma unroll
    for (int off = P2 / 2; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    return v;
  }
}

__device__ __forceinline__ uint64_t l2_keep_policy32() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void st_keep32(float* p, float v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_keep64(double* p, double v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}

// 2^e as a double / float.  pow2d clamps below at 2^-1022: raw values under that are 0 or denormal in the reference too
__device__ __forceinline__ double pow2d(int e) { return __longlong_as_double((long long)(max(e, -1022) + 1023) << 52); }
__device__ __forceinline__ float pow2f(int e) { return __int_as_float((e + 127) << 23); }
// binary exponent of a positive normal float
__device__ __forceinline__ int expof(float m) { return ((__float_as_int(m) >> 23) & 0xff) - 127; }
// un-scale: x * 2^e in float64
__device__ __forceinline__ double unscale(float x, int e) { return (double)x * pow2d(e); }
// argmax over a warp of non-negative floats, first index on ties (np.argmax); lanes without a candidate pass (0, INT_MAX)
__device__ __forceinline__ int warp_argmax_nonneg32(float v, int k) {
  const unsigned b = __float_as_uint(v);
  const unsigned mb = __reduce_max_sync(0xffffffffu, b);
  return __reduce_min_sync(0xffffffffu, b == mb ? k : 0x7fffffff);
}

// GEN ("generic width"): the instantiation is wider than the concept count needs (KG > ceil(K / LPR)), so ANY concept
// group of a lane can lie beyond K, not just the last one; validity comes from a per-lane bit mask.  Lets every
// K <= LPR * KG run on the warp kernel instead of dropping to the CTA-per-4-pairs kernel (ik_estep.cu).
template <int N, int KG, bool GEN = false>
__global__ void __launch_bounds__(kWpc32 * 32, ctas32(KG)) ik_estep_warp32_kernel(const EstepArgs a) {
  constexpr int LPR = 32 / N;
  constexpr int ROWL = N * LPR;
  constexpr int KS0 = LPR * KG;
  constexpr int KS = KS0 + (((LPR - KS0) % 32) + 32) % 32;       // smem row stride == LPR (mod 32) words
  constexpr int SL = KG * 32;                                    // floats per checkpoint slice
  constexpr int KC = (KS0 + 31) / 32;
  const int K = a.K, P = a.P;
  const double eps = a.eps;
  const int lane = threadIdx.x & 31;
  const int wic = threadIdx.x >> 5;
  const int gw = blockIdx.x * kWpc32 + wic;
  const int total_warps = gridDim.x * kWpc32;
  const bool on = lane < ROWL;
  const int i = on ? lane / LPR : 0;
  const int j = on ? lane - i * LPR : 0;
  // (ptxas re-derives lane / LPR and the scratch / table base addresses inside the time loops, ~30 integer
  // instructions per (pair, t) on the ncu source page; pinning them in registers with empty asm statements
  // changed nothing, 35.4 -> 35.5 ms: the kernel is bound by its dependency chains, not by the issue count)
  const bool head = on && j == 0;
  const bool kv_last = on && (j + LPR * (KG - 1) < K);
  unsigned kvmask = 0;                                            // GEN: bit q = concept j + LPR q of this lane exists
  if constexpr (GEN) {
#pragma unroll
    for (int q = 0; q < KG; ++q) kvmask |= (on && j + LPR * q < K) ? (1u << q) : 0u;
  }
  auto kvalid = [&](int q) -> bool {
    if constexpr (GEN) return (kvmask >> q) & 1u;
    else return (q < KG - 1) ? on : kv_last;
  };

  extern __shared__ float smem32[];
  float* obsS = smem32;                                          // [P][KS0] scaled emission table, columns >= K are 0
  int* shS = reinterpret_cast<int*>(obsS + ((P * KS0 + 3) & ~3));   // [P] shift of each phone type
  float* buf = reinterpret_cast<float*>(shS + ((P + 3) & ~3)) + (size_t)wic * 2 * (N * KS);   // two gamma slices per warp

  // stage the emission table: row x scaled by 2^sh[x], sh = -(exponent of the row maximum) - 1 (max lands in [0.5,1))
  for (int x = wic; x < P; x += kWpc32) {
    double m = 0.0;
    for (int k = lane; k < K; k += 32) m = fmax(m, a.obsT[(size_t)x * K + k]);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, s));
    int sh = 0;
    if (m > 0.0 && m < 1e300) sh = 1022 - (int)((__double_as_longlong(m) >> 52) & 0x7ff);   // -(e) - 1, e = unbiased exponent
    if (sh > 1000) sh = 1000;
    if (lane == 0) shS[x] = sh;
    const double sc = pow2d(sh);
    for (int k = lane; k < KS0; k += 32) obsS[x * KS0 + k] = (k < K) ? (float)(a.obsT[(size_t)x * K + k] * sc) : 0.0f;
  }
  __syncthreads();

  // EPS = epsm * 2^epse (epsm in [0.5, 1)); 0 for the un-floored classes
  int epse = 0;
  const float epsm = (float)frexp(eps, &epse);

  const float d_i = (float)a.trans[i * N + i];
  const float pi_i = (float)a.init[i];
  float acol[N], arow[N];
#pragma unroll
  for (int jp = 0; jp < N; ++jp) {
    acol[jp] = (jp == i) ? 0.0f : (float)a.trans[jp * N + i];
    arow[jp] = (jp == i) ? 0.0f : (float)a.trans[i * N + jp];
  }

  // per-warp scratch (float words): [NC][SL] checkpoints | [Tmax][N] c_t | [NC] exponent of each checkpoint
  float* scr = reinterpret_cast<float*>(a.scratch + (size_t)gw * a.cta_scratch);
  float* my_ckpt = scr + lane;
  float* hist = scr + (size_t)a.NC * SL + i;
  int* ck_exp = reinterpret_cast<int*>(scr + (size_t)a.NC * SL + (size_t)a.Tmax * N);
  const bool tab_on = a.part_phone != nullptr;
  double* tab = a.part_phone + (size_t)gw * P * K;
  const float* obs_j = obsS + j;

  const uint64_t keep = (MWD_W32_KEEP != 0) ? l2_keep_policy32() : 0;
  int ccol[KC];
  bool cok[KC];
#pragma unroll
  for (int m = 0; m < KC; ++m) {
    cok[m] = lane + 32 * m < K;
    ccol[m] = min(lane + 32 * m, KS0 - 1);
  }

  auto load_obs = [&](float (&o)[KG], int x) {
    const float* orow = obs_j + x * KS0;
#pragma unroll
    for (int q = 0; q < KG; ++q) o[q] = orow[LPR * q];      // padded rows: no guard on the last column group
  };
  auto warp_max = [&](const float (&v)[KG], float extra) {
    float m = extra;
#pragma unroll
    for (int q = 0; q < KG; ++q) m = fmaxf(m, v[q]);
    if (!on) m = 0.0f;                     // lanes beyond the lattice mirror lane 0 with pz = 0: keep them out
    return __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(m)));   // non-negative: bit order == value order
  };

#ifndef MWD_W32_LONGEST_FIRST
#define MWD_W32_LONGEST_FIRST 0
#endif
  // a bucket is sorted by caption length.  Taking the longest captions first (so that the launch does not end on a few
  // warps still working through a 125-phone caption) measured SLOWER at 1 M pairs, 34.56 -> 34.79 ms: neighbouring warps
  // on neighbouring pairs in ascending order keep the streamed inputs and the statistics rows adjacent
#if MWD_W32_LONGEST_FIRST
  for (int64_t pair = a.hi - 1 - gw; pair >= a.lo; pair -= total_warps) {
#else
  for (int64_t pair = a.lo + gw; pair < a.hi; pair += total_warps) {
#endif
    const int64_t p0 = a.phone_off[pair];
    const int T = a.phone_off[pair + 1] - (int32_t)p0;
    const int64_t r0 = a.region_off[pair];
    const int32_t* ph = a.phones + p0;
    // row t of the pair's statistics: 4 n float32 mantissas [s | F | dg | r], the exponents of s, (F, dg), r and the
    // all-floored flag (StatRow in ik_estep.cu un-scales them).  Rows of 4 n + 4 words are packed densely from the
    // start of the bucket's slot range (they need 16 n + 16 of the 32 n bytes a float64 row has), in global row order
    constexpr int RS = 4 * N + 4;             // row stride in 32-bit words
    const int64_t s_lo = a.slot_off[a.lo];
    float* st = reinterpret_cast<float*>(a.stats + 4 * s_lo) + ((a.slot_off[pair] - s_lo) / N) * RS + i;
    int* st_e = reinterpret_cast<int*>(st - i) + 4 * N;
    if (T <= 0) continue;

    float pz[KG];
    {
      const double* prow = a.pz + (r0 + i) * K + j;
#pragma unroll
      for (int q = 0; q < KG; ++q) pz[q] = kvalid(q) ? (float)__ldcs(prow + LPR * q) : 0.0f;
    }

    // ------------------------------------------------------------------ forward sweep
    float al[KG];
    double inorm = 0.0;
    int ea;
    {
      int xn = 0;
      {
        float o[KG];
        const int x0 = ph[0];
        load_obs(o, x0);
#pragma unroll
        for (int q = 0; q < KG; ++q) al[q] = (pi_i * pz[q]) * o[q];
        ea = -shS[x0];
        if (T > 1) xn = ph[1];
      }
      int cidx = 0;
      for (int t = 0; t < T; ++t) {
        float onext[KG];
        int shn = 0;
        if (t + 1 < T) {
          load_obs(onext, xn);
          shn = shS[xn];
          if (t + 2 < T) xn = ph[t + 2];
        }
        if ((t & 3) == 0) {               // power-of-two renormalisation (exact); even t, so checkpoints carry it
          const float m = warp_max(al, 0.0f);
          if (m > 0.0f) {
            const int e = expof(m);
            const float sc = pow2f(-e);
#pragma unroll
            for (int q = 0; q < KG; ++q) al[q] *= sc;
            ea += e;
          }
        }
        if (!a.ll_only && (t & 1) == 0) {
          float* dst = my_ckpt + (size_t)cidx * SL;
#pragma unroll
          for (int q = 0; q < KG; ++q) {
            if (MWD_W32_KEEP & 1) st_keep32(dst + 32 * q, al[q], keep);
            else __stcg(dst + 32 * q, al[q]);
          }
          if (lane == 0) __stcg(ck_exp + cidx, ea);
          ++cidx;
        }
        float s = 0.0f, s_b = 0.0f;
#pragma unroll
        for (int q = 0; q < KG; ++q) {
          if (q & 1) s_b += al[q];
          else s += al[q];
        }
        s = row_sum_head32<LPR>(s + s_b, j);
        float sv[N];
#pragma unroll
        for (int jp = 0; jp < N; ++jp) sv[jp] = __shfl_sync(0xffffffffu, s, jp * LPR);
        if (t == T - 1) {
          float Lh = 0.0f;
#pragma unroll
          for (int jp = 0; jp < N; ++jp) Lh += sv[jp];
          double L = unscale(Lh, ea);
          L = floor_at(L, eps);
          if (lane == 0) a.pair_ll[pair] = log(L);                       // :529
          if (head && !a.ll_only) __stcs(st + t * RS, -1.0f);            // sentinel: last row of the pair
          if (lane == 0 && !a.ll_only) __stcs(st_e + t * RS, 0);
          inorm = 1.0 / L;
        } else {
          float c = 0.0f;
#pragma unroll
          for (int jp = 0; jp < N; ++jp) c = fmaf(acol[jp], sv[jp], c);
          if (head && !a.ll_only) {
            __stcg(hist + t * N, c);
            __stcs(st + t * RS, s);
          }
          if (lane == 0 && !a.ll_only) __stcs(st_e + t * RS, ea);
#pragma unroll
          for (int q = 0; q < KG; ++q) al[q] = onext[q] * fmaf(d_i, al[q], c * pz[q]);
          ea -= shn;
        }
      }
    }
    if (a.ll_only) continue;

    // ------------------------------------------------------------------ backward sweep
    float bo[KG];           // beta_{t+1} * o_{t+1}, scaled by 2^-eb
#pragma unroll
    for (int q = 0; q < KG; ++q) bo[q] = 0.0f;
    float w = 1.0f;         // (Aoff r_{t+1})[i], same scale; with bo = 0 the first step gets beta_{T-1} = 1
    int eb = 0;

    // one backward step: alpha_t (scaled by 2^-eat) in av, scaled emissions o_t in o (shift sh)
    auto bwd_step = [&](int t, int x, int sh, int eat, const float (&av)[KG], const float (&o)[KG], float* gslice) {
      if ((t & 3) == 3) {   // renormalise the carried (bo, w)
        const float m = warp_max(bo, on ? w : 0.0f);
        if (m > 0.0f) {
          const int e = expof(m);
          const float sc = pow2f(-e);
#pragma unroll
          for (int q = 0; q < KG; ++q) bo[q] *= sc;
          w *= sc;
          eb += e;
        }
      }
      const int eg = eat + eb;                    // exponent of gamma^, dg^
      // the EPS floor in scaled units, EPS * 2^-eg = epsm * 2^E: beyond 2^100 every entry of the step is floored
      // (gamma^ <= ~8), below 2^-125 it cannot matter; no float64 arithmetic on this path
      const int E = epse - eg;
      const bool all_floored = (epsm > 0.0f) && (E > 100);
      const float epsf = (all_floored || E < -125) ? 0.0f : epsm * pow2f(E);
      const double sg = pow2d(eg);
      double tabv[KC];
      double* trow = tab + x * K + lane;
      // (unconditional loads from a clamped column instead of this guarded form: +6 ms, the loads then sit on the
      // critical path of the step)
#pragma unroll
      for (int m = 0; m < KC; ++m) tabv[m] = (tab_on && cok[m]) ? __ldcg(trow + 32 * m) : 0.0;
      float* grow = gslice + i * KS + j;
      float sumF = 0.0f, dg = 0.0f, rr = 0.0f, sumF_b = 0.0f, dg_b = 0.0f, rr_b = 0.0f;
#pragma unroll
      for (int q = 0; q < KG; ++q) {
        const bool kv = kvalid(q);
        const float beta = fmaf(d_i, bo[q], w);
        const float g = av[q] * beta;
        const float f = kv ? fmaxf(g, epsf) : 0.0f;
        if (q & 1) {
          dg_b = fmaf(av[q], bo[q], dg_b);
          sumF_b += f;
        } else {
          dg = fmaf(av[q], bo[q], dg);
          sumF += f;
        }
        bo[q] = beta * o[q];
        if (q & 1) rr_b = fmaf(bo[q], pz[q], rr_b);
        else rr = fmaf(bo[q], pz[q], rr);
        if (on) grow[LPR * q] = g;
      }
      sumF += sumF_b;
      rr += rr_b;
      dg = (dg + dg_b) * d_i;
      const int ebo = eb - sh;                    // exponent of the new bo / rr / w
      __syncwarp();
      // column sums: unconditional loads from a clamped column (lanes past K re-read column KS0-1: a broadcast, no
      // bank conflict), masked afterwards -- the guarded form compiles to one branch region per column chunk
      float cs[KC];
#pragma unroll
      for (int m = 0; m < KC; ++m) {
        const float* col = gslice + ccol[m];
        float acc = 0.0f;
#pragma unroll
        for (int ii = 0; ii < N; ++ii) acc += col[ii * KS];
        cs[m] = cok[m] ? acc : 0.0f;
      }
      sumF = row_sum_head32<LPR>(sumF, j);
      dg = row_sum_head32<LPR>(dg, j);
      rr = row_sum_head32<LPR>(rr, j);
      {                   // row statistics of this step for the count post-pass (head lanes store)
        float* sp = st + t * RS;
        if (head) {
          __stcs(sp + N, sumF);
          __stcs(sp + 2 * N, dg);
          __stcs(sp + 3 * N, rr);
        }
        if (lane == 0) {    // exponents of (F, dg) and of r; flag: every entry floored, the K concepts of a row give EPS each
          __stcs(st_e + t * RS + 1, eg);
          __stcs(reinterpret_cast<int2*>(st_e + t * RS + 2), make_int2(ebo, all_floored ? 1 : 0));
        }
      }
      float wn = 0.0f;
#pragma unroll
      for (int jp = 0; jp < N; ++jp) wn = fmaf(arow[jp], __shfl_sync(0xffffffffu, rr, jp * LPR), wn);
      w = wn;
      eb = ebo;
      // conceptCountsA[t][k] = sum_i gamma_t[i][k] / max(L, EPS);  phoneCounts[k][x_t] += ...  (:430,:233,:235)
      const double gsc = sg * inorm;
      double cs64[KC];
#pragma unroll
      for (int m = 0; m < KC; ++m) {
        const double v = (double)cs[m] * gsc;
        if (tab_on && cok[m]) {
          if (MWD_W32_KEEP & 2) st_keep64(trow + 32 * m, tabv[m] + v, keep);
          else __stcg(trow + 32 * m, tabv[m] + v);
        }
        cs64[m] = v;
      }
      if (a.cA_out != nullptr) {      // materialised conceptCountsA (off by default): one uniform branch per step
        double* crow = a.cA_out + (p0 + t) * K + lane;
#pragma unroll
        for (int m = 0; m < KC; ++m)
          if (cok[m]) crow[32 * m] = cs64[m];
      }
      if (a.ca_out) {     // concept_alignment[t] = argmax_k cA[t][k] (first index on ties, :628); scale-invariant
        float bv = 0.0f;
        int bk = 0x7fffffff;
#pragma unroll
        for (int m = 0; m < KC; ++m)
          if (cok[m] && (bk == 0x7fffffff || __float_as_uint(cs[m]) > __float_as_uint(bv))) { bv = cs[m]; bk = lane + 32 * m; }
        const int kbest = warp_argmax_nonneg32(bv, bk);
        if (lane == 0) a.ca_out[p0 + t] = kbest;
      }
    };

    auto drop_slice = [&](int c) {
      // checkpoint slice c is dead: drop it from L2 instead of letting it be written back to HBM
      for (int ln = lane; ln < SL / 32; ln += 32) {
        const float* dead = scr + (size_t)c * SL + ln * 32;
        asm volatile("discard.global.L2 [%0], 128;" ::"l"(dead) : "memory");
      }
    };
    const int nblk = (T + 1) / 2;
    int par = 0;
#if MWD_W32_PF == 2
    float a0n[KG];
    {
      const float* src = my_ckpt + (size_t)(nblk - 1) * SL;
#pragma unroll
      for (int q = 0; q < KG; ++q) a0n[q] = __ldcg(src + 32 * q);
    }
#endif
    for (int c = nblk - 1; c >= 0; --c) {
      const int t0 = 2 * c;
      float a0[KG];
#if MWD_W32_PF == 2
#pragma unroll
      for (int q = 0; q < KG; ++q) a0[q] = a0n[q];
      if (c > 0) {        // the next block's checkpoint travels while this block computes
        const float* src = my_ckpt + (size_t)(c - 1) * SL;
#pragma unroll
        for (int q = 0; q < KG; ++q) a0n[q] = __ldcg(src + 32 * q);
      }
#else
      {
        const float* src = my_ckpt + (size_t)c * SL;
#pragma unroll
        for (int q = 0; q < KG; ++q) a0[q] = MWD_W32_PF ? __ldca(src + 32 * q) : __ldcg(src + 32 * q);
      }
#if MWD_W32_PF == 1
      if (c > 0) {        // pull the next block's checkpoint (one 128-byte line per register) from L2 into L1 meanwhile;
                          // the lines were written by this very warp, so its own L1 cannot hold them stale
        const float* nxt = my_ckpt + (size_t)(c - 1) * SL;
#pragma unroll
        for (int q = 0; q < KG; ++q) asm volatile("prefetch.global.L1 [%0];" ::"l"(nxt + 32 * q));
      }
#endif
#endif
      const int e0 = __ldcg(ck_exp + c);
      const int x0 = ph[t0];
      if (t0 + 1 < T) {
        const int x1 = ph[t0 + 1];
        const float cb = __ldcg(hist + t0 * N);
        float o1[KG], a1[KG];
        load_obs(o1, x1);
        const int sh1 = shS[x1];
#pragma unroll
        for (int q = 0; q < KG; ++q) a1[q] = o1[q] * fmaf(d_i, a0[q], cb * pz[q]);
        bwd_step(t0 + 1, x1, sh1, e0 - sh1, a1, o1, buf + par * (N * KS));
        par ^= 1;
      }
      {
        float o0[KG];
        load_obs(o0, x0);
        bwd_step(t0, x0, shS[x0], e0, a0, o0, buf + par * (N * KS));
        par ^= 1;
      }
      drop_slice(c);
    }
  }
}

struct Warp32Plan {
  int KG, NC, grid;
  bool gen;               // generic-width instantiation (KG wider than the concept count needs)
  size_t smem;
  int64_t warp_scratch;   // doubles per warp
};

// (n, KG) instantiations: every n <= 10 at the concept counts of the reference's configurations (K = 65: MSCOCO,
// K = 50 / 100: Flickr30k) plus K = 40 / 80 for the generic tests
#define MWD_WARP32_COMBOS(X)                                                                            \
  X(1, 2) X(1, 3) X(1, 4) X(2, 3) X(2, 4) X(2, 5) X(2, 7) X(3, 4) X(3, 5) X(3, 7) X(3, 8) X(3, 10) X(4, 5) X(4, 7) \
  X(4, 9) X(4, 10) X(4, 13) X(5, 7) X(5, 9) X(5, 11) X(5, 14) X(5, 17) X(6, 8) X(6, 10) X(6, 13) X(6, 16) X(6, 20) \
  X(7, 10) X(7, 13) X(7, 17) X(7, 20) X(7, 25) X(8, 10) X(8, 13) X(8, 17) X(8, 20) X(8, 25) X(9, 14) X(9, 17)     \
  X(9, 22) X(9, 27) X(10, 14) X(10, 17) X(10, 22) X(10, 27)

// generic-width instantiations: every exact width again with the per-group validity mask, plus the widths that take
// n <= 8 up to K = 128 (n = 9, 10: K <= 81 -- the state of a wider lattice does not fit the register file)
#define MWD_WARP32_GEN_COMBOS(X) \
  MWD_WARP32_COMBOS(X) X(2, 8) X(3, 13) X(4, 16) X(5, 22) X(6, 26) X(7, 32) X(8, 32)

static bool warp32_combo(int n, int KG) {
#define X(NN, GG) if (n == NN && KG == GG) return true;
  MWD_WARP32_COMBOS(X)
#undef X
  return false;
}
// smallest generic-width instantiation of n that holds KG concept groups per lane; 0 if none
static int warp32_gen_width(int n, int KG) {
  int best = 0;
#define X(NN, GG) if (n == NN && GG >= KG && (best == 0 || GG < best)) best = GG;
  MWD_WARP32_GEN_COMBOS(X)
#undef X
  return best;
}

static bool plan_warp32(int n, int K, int P, int Tmax, int64_t npairs, Warp32Plan* pl) {
  if (n < 1 || n > 10) return false;
  if (const char* e = getenv("MWD_ESTEP_WARP32")) { if (atoi(e) == 0) return false; }
  const int lpr = 32 / n;
  pl->KG = (K + lpr - 1) / lpr;
  pl->gen = false;
  if (!warp32_combo(n, pl->KG)) {
    pl->KG = warp32_gen_width(n, pl->KG);
    if (pl->KG == 0) return false;
    pl->gen = true;
  }
  const int ks0 = lpr * pl->KG;
  const int ks = ks0 + (((lpr - ks0) % 32) + 32) % 32;
  pl->smem = ((size_t)((P * ks0 + 3) & ~3) + ((P + 3) & ~3) + (size_t)kWpc32 * 2 * n * ks) * sizeof(float);
  const int cps = ctas32(pl->KG);
  if (pl->smem > (size_t)224 * 1024 / cps - 1024) return false;
  pl->NC = (Tmax + 1) / 2;
  if (pl->NC < 1) pl->NC = 1;
  int64_t grid = (int64_t)sm_count() * cps;
  const int64_t need = (npairs + kWpc32 - 1) / kWpc32;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  if (grid * kWpc32 > estep_grid_rows()) return false;
  pl->grid = (int)grid;
  const int64_t words = (int64_t)pl->NC * pl->KG * 32 + (int64_t)Tmax * n + pl->NC;
  pl->warp_scratch = (((words + 1) / 2) + 15) & ~(int64_t)15;
  return true;
}

template <int N, int KG, bool GEN>
static int launch_warp32(const EstepArgs& a, const Warp32Plan& pl, cudaStream_t st) {
  auto kern = ik_estep_warp32_kernel<N, KG, GEN>;
  MWD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  kern<<<pl.grid, kWpc32 * 32, pl.smem, st>>>(a);
  MWD_CHECK_LAUNCH();
  return 0;
}

}  // namespace

int64_t estep_warp32_scratch(int n, int K, int P, int Tmax, int64_t npairs) {
  Warp32Plan pl;
  if (!plan_warp32(n, K, P, Tmax, npairs, &pl)) return 0;
  return pl.warp_scratch * pl.grid * kWpc32;
}

int estep_warp32_launch(EstepArgs a, cudaStream_t st) {
  Warp32Plan pl;
  MWD_REQUIRE(plan_warp32(a.n, a.K, a.P, a.Tmax, a.hi - a.lo, &pl), "float32 warp E-step: unsupported (n=%d, K=%d)", a.n, a.K);
  a.B = 2;
  a.NC = pl.NC;
  a.cta_scratch = pl.warp_scratch;
  if (!pl.gen) {
#define X(NN, GG) if (a.n == NN && pl.KG == GG) return launch_warp32<NN, GG, false>(a, pl, st);
    MWD_WARP32_COMBOS(X)
#undef X
  } else {
#define X(NN, GG) if (a.n == NN && pl.KG == GG) return launch_warp32<NN, GG, true>(a, pl, st);
    MWD_WARP32_GEN_COMBOS(X)
#undef X
  }
  set_error("float32 warp E-step: no instantiation for (n=%d, KG=%d)", a.n, pl.KG);
  return 2;
}

}  // namespace mwd

// K2 -- restricted-chain concept posterior (updateConceptCounts,
// hmm_dnn/image_phone_hmm_word_discoverer.py:443-465).
//
// For every (region i, concept k) the reference clamps region i to concept k and runs a plain
// n-state HMM forward pass with marginal emissions e_t[j] = sum_k' pz[j][k'] obs[k'][x_t]
// (row i replaced by obs[k][x_t]); the chain likelihoods L(i,k) give
//   conceptCounts[i][k] = pz[i][k] L(i,k) / sum_k pz[i][k] L(i,k).
// This is 2*T*K*n^3 flop per pair -- the dominant term of the whole EM iteration.
//
// Mapping: one CTA per pair, one thread per chain (i,k) with the n-vector state in registers
// (N is a template parameter so the n x n recurrence is fully unrolled); the transition matrix
// lives in __constant__ memory so every DFMA takes it as an immediate constant-bank operand, the
// marginal emissions are broadcast from shared memory, obs[.][x_t] comes from the transposed
// (P x K) table through L1.  Raw float64 probability domain, no floors -- as the reference.
//
// Mixed-precision variant (mwd_ik_problem.mixed_precision & MWD_MIXED_CONCEPT): the same chains in
// float32 on the FFMA pipe (twice the FP64 rate on B200).  updateConceptCounts has no EPS floor and
// its result is a ratio over k, so the chains may be rescaled freely: the emission table is pre-scaled
// per phone type by a power of two (common to all chains of a pair, cancels in the ratio) and every
// chain carries its own integer exponent, renormalised every 8 steps; only the last step --
// pz * L / sum_k pz * L -- runs in float64.  Accuracy ~1e-6 relative on conceptCounts (tests:
// tests/test_gpu_mixed_precision.py); entries of obs below 2^-126 of their phone's maximum flush to 0.
//
// The clamped emission o_k(x_t) is carried as a float32 PAIR (hi, lo).  A float32 rounding of the (P x K)
// emission table is the SAME relative error in every pair of the corpus, so it survives the sum over pairs
// in the posterior gradient (CPU emulation, profiles/r02_mixed_precision.md section 5: gradient error 1.4e-7 of
// its scale with a rounded table, 9e-9 with the pair; splitting the marginal emissions or the transition
// matrix as well changes nothing -- their roundings differ from pair to pair and average out).
#include <stdlib.h>

#include <mutex>

#ifndef MWD_C32_PROLOGUE_F32
#define MWD_C32_PROLOGUE_F32 1      // float32 phone-table prologue of the float32 chains (0: float64)
#endif
#include <type_traits>

#include "mwd_common.cuh"

namespace mwd {

__constant__ double c_trans[(kNMax + 1) * kNMax * kNMax];
__constant__ double c_init[(kNMax + 1) * kNMax];
__constant__ float c_trans32[(kNMax + 1) * kNMax * kNMax];
__constant__ float c_init32[(kNMax + 1) * kNMax];

// The constant tables are process-wide: two engines driving this entry point from different streams or
// host threads would overwrite them under a running kernel.  Every call therefore (a) holds this mutex
// while it enqueues its copies and kernels and (b) first makes its stream wait for the event recorded
// after the previous call's last kernel -- calls are serialised on the device, whatever stream they use.
static std::mutex g_const_mutex;
static cudaEvent_t g_const_event[16] = {};

struct ConceptArgs {
  const int32_t* region_off;
  const int32_t* phone_off;
  const int32_t* phones;
  const double* pz;
  const double* obsT;
  double* cC;
  int64_t lo, hi;
  int K, Tmax;
};

template <int N>
__global__ void __launch_bounds__(1024) ik_concept_kernel(const ConceptArgs a) {
  const int K = a.K;
  const int64_t pair = a.lo + blockIdx.x;
  const int tid = threadIdx.x;
  const int p0 = a.phone_off[pair];
  const int T = a.phone_off[pair + 1] - p0;
  const int64_t r0 = a.region_off[pair];
  const int32_t* ph = a.phones + p0;

  extern __shared__ double smem[];
  double* s_pz = smem;                    // [N][K]
  double* s_num = s_pz + N * K;           // [N][K]
  double* s_e = s_num + N * K;            // [Tmax][N]
  double* s_row = s_e + (size_t)a.Tmax * N;  // [N]
  int* s_x = reinterpret_cast<int*>(s_row + kNMax);  // [Tmax]

  for (int e = tid; e < N * K; e += blockDim.x) s_pz[e] = a.pz[r0 * K + e];
  for (int t = tid; t < T; t += blockDim.x) s_x[t] = ph[t];
  __syncthreads();
  // marginal emissions e[t][j] = sum_k pz[j][k] * obs[k][x_t]  (sequential k, like a GEMM k-loop)
  for (int e = tid; e < T * N; e += blockDim.x) {
    int t = e / N, j = e - t * N;
    const double* orow = a.obsT + (size_t)s_x[t] * K;
    const double* prow = s_pz + j * K;
    double acc = 0.0;
    for (int k = 0; k < K; ++k) acc = fma(prow[k], __ldg(orow + k), acc);
    s_e[e] = acc;
  }
  __syncthreads();

  const double* A = c_trans + N * (kNMax * kNMax);
  const double* pi = c_init + N * kNMax;
  // one step of the restricted chain: G = (F A) * e', e'[j] = e_t[j] except e'[i] = o.
  // SI >= 0: the clamped region is a compile-time constant (warp-uniform chains), so the emission
  // pick costs nothing; SI < 0: generic per-thread select.
  auto chain_step = [&](auto si_tag, const double (&F)[N], double (&G)[N], const double* et, double o, int i) {
    constexpr int SI = decltype(si_tag)::value;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double acc = 0.0;
#pragma unroll
      for (int l = 0; l < N; ++l) acc = fma(F[l], A[l * N + j], acc);
      if (SI >= 0) G[j] = acc * ((j == SI) ? o : et[j]);
      else G[j] = acc * ((j == i) ? o : et[j]);
    }
  };
  // whole restricted chain of (region i, concept k); returns pz[i][k] * L(i,k)
  auto run_chain = [&](auto si_tag, int i, int k) {
    constexpr int SI = decltype(si_tag)::value;
    const double* ocol = a.obsT + k;
    double F[N], G[N];
    {
      double o = __ldg(ocol + (size_t)s_x[0] * K);
#pragma unroll
      for (int j = 0; j < N; ++j) F[j] = pi[j] * ((j == (SI >= 0 ? SI : i)) ? o : s_e[j]);
    }
    int t = 1;
    // two steps per trip (F -> G -> F): no register copies between steps
    for (; t + 1 < T; t += 2) {
      const double o0 = __ldg(ocol + (size_t)s_x[t] * K);
      const double o1 = __ldg(ocol + (size_t)s_x[t + 1] * K);
      chain_step(si_tag, F, G, s_e + t * N, o0, i);
      chain_step(si_tag, G, F, s_e + (t + 1) * N, o1, i);
    }
    if (t < T) {
      const double o0 = __ldg(ocol + (size_t)s_x[t] * K);
      chain_step(si_tag, F, G, s_e + t * N, o0, i);
#pragma unroll
      for (int j = 0; j < N; ++j) F[j] = G[j];
    }
    double lik = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) lik += F[j];
    return lik;
  };
  // chain -> thread map: the first N * KF chains (KF = K rounded down to whole warps) are laid out
  // region-major with KF concepts per region, so every warp there works on ONE region; the
  // (K - KF) * N left-over chains follow and take the generic path.
  // (only for N <= 6: the 64-register budget of the 1024-thread launch bound has no room for the
  // N-vector state of the larger specialisations)
  const int KF = (N <= 6) ? (K & ~31) : 0;
  for (int c = tid; c < N * K; c += blockDim.x) {
    int i, k;
    double lik;
    if (c < N * KF) {
      i = c / KF;
      k = c - i * KF;
      switch (i) {   // warp-uniform
#define MWD_SI(V) \
  case V:         \
    if constexpr (V < N && N <= 6) lik = run_chain(std::integral_constant<int, V>{}, i, k); else lik = 0.0; \
    break;
        MWD_SI(0) MWD_SI(1) MWD_SI(2) MWD_SI(3) MWD_SI(4) MWD_SI(5) MWD_SI(6) MWD_SI(7)
        MWD_SI(8) MWD_SI(9) MWD_SI(10) MWD_SI(11) MWD_SI(12) MWD_SI(13) MWD_SI(14) MWD_SI(15)
#undef MWD_SI
        default: lik = 0.0;
      }
    } else {
      const int cc = c - N * KF, rem = K - KF;
      i = cc / rem;
      k = KF + (cc - i * rem);
      lik = run_chain(std::integral_constant<int, -1>{}, i, k);
    }
    s_num[i * K + k] = s_pz[i * K + k] * lik;
  }
  __syncthreads();
  // row sums over k, one warp per region
  const int warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
  for (int i = warp; i < N; i += nwarp) {
    double s = 0.0;
    for (int k = lane; k < K; k += 32) s += s_num[i * K + k];
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    if (lane == 0) s_row[i] = s;
  }
  __syncthreads();
  for (int c = tid; c < N * K; c += blockDim.x) a.cC[r0 * K + c] = s_num[c] / s_row[c / K];
}

// ------------------------------------------------------------------------------------------ float32 chains
// packed float32 pairs (sm_100a FFMA2 / FMUL2); ptxas folds dup2() of a register, uniform-register or constant
// operand into the instruction's broadcast form, so it costs nothing
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(d)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
        "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;"
      : "=l"(d)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 dup2(float v) { return make_float2(v, v); }
struct ConceptArgs32 {
  const int32_t* region_off;
  const int32_t* phone_off;
  const int32_t* phones;
  const double* pz;
  const double* obsT;       // (P x K) float64: marginal emissions are formed from it in float64
  const float4* obsS;       // (P x ceil(K/2)) {hi[k], hi[k+1], lo[k], lo[k+1]}, k even: float32 pair (hi, lo) of
                            // obsT[x][k] * 2^shift[x]; one 128-bit load serves the two chains of a paired lane
  const int32_t* shift;     // (P)
  const double* obsKP;      // (K x P) float64, column x scaled by 2^shift[x] (phone-table form of the marginal emissions)
  const float* obsKPf;      // the same table rounded to float32 (float32 prologue)
  int P, kseg_n;            // phone inventory; k-segments of the phone-table prologue (0: direct per-(t, j) dot products)
  const float* trans32;     // float32 copies of the full tables in GLOBAL memory (register-resident path)
  const float* init32;
  double* cC;
  int64_t lo, hi;
  int K, Tmax;
};

// obsS[x][k] = obsT[x][k] * 2^shift[x], shift[x] = -ceil(log2(max_k obsT[x][k])) (0 for an all-zero row);
// one warp per phone type.  Also converts the transition / initial tables to float32.
__global__ void concept_prepare32_kernel(const double* __restrict__ obsT, int P, int K, float4* __restrict__ obsS,
                                         double* __restrict__ obsKP, float* __restrict__ obsKPf,
                                         int32_t* __restrict__ shift, const double* __restrict__ trans,
                                         const double* __restrict__ init, float* __restrict__ trans32,
                                         float* __restrict__ init32) {
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nth = gridDim.x * blockDim.x;
  for (int e = gtid; e < (kNMax + 1) * kNMax * kNMax; e += nth) trans32[e] = (float)trans[e];
  for (int e = gtid; e < (kNMax + 1) * kNMax; e += nth) init32[e] = (float)init[e];
  const int lane = threadIdx.x & 31;
  const int warp = gtid >> 5, nwarp = nth >> 5;
  for (int x = warp; x < P; x += nwarp) {
    double m = 0.0;
    for (int k = lane; k < K; k += 32) m = fmax(m, obsT[(size_t)x * K + k]);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, s));
    int sh = 0;
    if (m > 0.0 && m < INFINITY) {
      int e;
      frexp(m, &e);            // m = f * 2^e, f in [0.5, 1)
      sh = -e;
    }
    if (lane == 0) shift[x] = sh;
    const int Kq = (K + 1) >> 1;
    float* row = reinterpret_cast<float*>(obsS + (size_t)x * Kq);
    for (int k = lane; k < 2 * Kq; k += 32) {
      const double v = (k < K) ? ldexp(obsT[(size_t)x * K + k], sh) : 0.0;
      const float hi = (float)v;
      row[(k >> 1) * 4 + (k & 1)] = hi;
      row[(k >> 1) * 4 + (k & 1) + 2] = (float)(v - (double)hi);
      if (k < K) {
        obsKP[(size_t)k * P + x] = v;
        obsKPf[(size_t)k * P + x] = hi;
      }
    }
  }
}

// chain -> warp map of the float32 kernel (n <= 8): per region, nd "paired" warps own 64 concepts each (a lane runs
// k and k + 1 in the two halves of packed float32 pairs), then at most one warp of 32 single chains; both tiers
// work on ONE region per warp, so the clamped region is a compile-time constant of the code they run.  The
// K % 32 left-over concepts of all regions -- of all PPC pairs of the CTA -- are packed into "tail" warps whose
// lanes each carry their own clamped region.  n > 8: tail warps only.
struct ConceptMap { int nd, ns, rem, wpp, tpp; };
template <int N>
__host__ __device__ inline ConceptMap concept_map(int K) {
  ConceptMap m;
  m.nd = (N <= 8) ? K / 64 : 0;
  m.ns = (N <= 8) ? (K - 64 * m.nd) / 32 : 0;
  m.rem = K - 64 * m.nd - 32 * m.ns;
  m.wpp = N * (m.nd + m.ns);      // region-bound warps per pair
  m.tpp = N * m.rem;              // tail chains per pair
  return m;
}
__host__ __device__ inline int concept_block_warps(const ConceptMap& m, int ppc) {
  return ppc * m.wpp + (ppc * m.tpp + 31) / 32;
}

// One CTA = PPC consecutive pairs of a bucket (same n, neighbouring T: the corpus is sorted by (n, T)).  PPC > 1
// exists to fill the tail warps: at K = 65 a pair has 5 left-over chains, six pairs fill 30 of a warp's 32 lanes.
template <int N, int PPC>
__global__ void __launch_bounds__(1024) ik_concept32_kernel(const ConceptArgs32 a) {
  constexpr int NP = (N <= 4) ? 4 : ((N <= 8) ? 8 : 16);      // row stride of the emission slab (floats)
  const int K = a.K, NK = N * K, Tmax = a.Tmax;
  const int64_t pair0 = a.lo + (int64_t)blockIdx.x * PPC;
  const int npair = (a.hi - pair0 < PPC) ? (int)(a.hi - pair0) : PPC;
  const int tid = threadIdx.x, warp_id = tid >> 5, lane_id = tid & 31, nwarps = blockDim.x >> 5;
  const int64_t r0 = a.region_off[pair0];      // the CTA's pairs own the contiguous region rows r0 .. r0 + npair * N

  extern __shared__ double smem[];
  double* s_pz = smem;                                              // [PPC][N][K]
  double* s_num = s_pz + PPC * NK;                                  // [PPC][N][K] chain mantissa, then pz * L
  float* s_e = reinterpret_cast<float*>(s_num + PPC * NK);       // [PPC][Tmax][NP] scaled marginal emissions
  unsigned* s_off = reinterpret_cast<unsigned*>(s_e + (size_t)PPC * Tmax * NP);   // [PPC][Tmax] byte offset of row x_t in obsS
  int* s_exp = reinterpret_cast<int*>(s_off + PPC * Tmax);          // [PPC][N][K] chain exponents
  double* s_part = reinterpret_cast<double*>(s_exp + PPC * NK + ((PPC * (Tmax + NK)) & 1));   // [PPC][S][P][N] (8-byte aligned)

  // float32 copy of the posteriors for the float32 prologue; it borrows s_num, which is first written when the chains
  // publish their results
  float* s_pzf = reinterpret_cast<float*>(s_num);
  for (int e = tid; e < npair * NK; e += blockDim.x) {
    const double v = a.pz[r0 * K + e];
    s_pz[e] = v;
    if (MWD_C32_PROLOGUE_F32) s_pzf[e] = (float)v;
  }
  for (int q = 0; q < npair; ++q) {
    const int p0 = a.phone_off[pair0 + q], T = a.phone_off[pair0 + q + 1] - p0;
    for (int t = tid; t < T; t += blockDim.x)
      s_off[q * Tmax + t] = (unsigned)a.phones[p0 + t] * (unsigned)((K + 1) >> 1) * (unsigned)sizeof(float4);
  }
  __syncthreads();
  // marginal emissions e[t][j] = sum_k pz[j][k] * obs[k][x_t] in float64, scaled like the table row and rounded
  // to float32.  Phone-table form (small inventories): M[x][j] = sum_k pz[j][k] * obs[k][x] for every phone type
  // x -- one thread per (k-segment, x) with the n sums in registers, the (K x P) table read coalesced along x and
  // the posterior broadcast from shared memory -- then e[t] = M[x_t].  The direct form (thread per (t, j),
  // K-term dot product) gathers its table rows by phone: 3x the L1 wavefronts and 5x the dependent-FMA depth
  // at P = 49, T = 50 (ncu: 47 % of the warps' resident time was spent before the chains started).
  if (a.kseg_n > 0) {
    const int P = a.P, S = a.kseg_n, kseg = (K + S - 1) / S;
    float* s_partf = reinterpret_cast<float*>(s_part);
    for (int item = tid; item < npair * S * P; item += blockDim.x) {
      const int q = item / (S * P), r = item - q * (S * P), seg = r / P, x = r - seg * P;
      const int k1 = min(K, (seg + 1) * kseg);
      if (MWD_C32_PROLOGUE_F32) {
        // float32 phone table: the marginal emissions are rounded to float32 below anyway, and their rounding differs
        // from pair to pair (it averages out over the corpus, profiles/r02_mixed_precision.md section 5)
        const float* pzq = s_pzf + q * NK;
        float acc[N];
#pragma unroll
        for (int j = 0; j < N; ++j) acc[j] = 0.0f;
        for (int k = seg * kseg; k < k1; ++k) {
          const float o = __ldg(a.obsKPf + (size_t)k * P + x);
#pragma unroll
          for (int j = 0; j < N; ++j) acc[j] = fmaf(pzq[j * K + k], o, acc[j]);
        }
        float* dst = s_partf + ((size_t)(q * S + seg) * P + x) * N;
#pragma unroll
        for (int j = 0; j < N; ++j) dst[j] = acc[j];
      } else {
        const double* pzq = s_pz + q * NK;
        double acc[N];
#pragma unroll
        for (int j = 0; j < N; ++j) acc[j] = 0.0;
        for (int k = seg * kseg; k < k1; ++k) {
          const double o = __ldg(a.obsKP + (size_t)k * P + x);
#pragma unroll
          for (int j = 0; j < N; ++j) acc[j] = fma(pzq[j * K + k], o, acc[j]);
        }
        double* dst = s_part + ((size_t)(q * S + seg) * P + x) * N;
#pragma unroll
        for (int j = 0; j < N; ++j) dst[j] = acc[j];
      }
    }
    __syncthreads();
    for (int q = 0; q < npair; ++q) {
      const int p0 = a.phone_off[pair0 + q], T = a.phone_off[pair0 + q + 1] - p0;
      for (int e = tid; e < T * N; e += blockDim.x) {
        const int t = e / N, j = e - t * N;
        const int x = a.phones[p0 + t];
        if (MWD_C32_PROLOGUE_F32) {
          float v = 0.0f;
          for (int seg = 0; seg < S; ++seg) v += s_partf[((size_t)(q * S + seg) * P + x) * N + j];
          s_e[((size_t)q * Tmax + t) * NP + j] = v;
        } else {
          double v = 0.0;
          for (int seg = 0; seg < S; ++seg) v += s_part[((size_t)(q * S + seg) * P + x) * N + j];
          s_e[((size_t)q * Tmax + t) * NP + j] = (float)v;
        }
      }
    }
  } else {
    for (int q = 0; q < npair; ++q) {
      const int p0 = a.phone_off[pair0 + q], T = a.phone_off[pair0 + q + 1] - p0;
      for (int e = tid; e < T * N; e += blockDim.x) {
        const int t = e / N, j = e - t * N;
        const int x = a.phones[p0 + t];
        const double* orow = a.obsT + (size_t)x * K;
        const double* prow = s_pz + q * NK + j * K;
        double acc = 0.0;
        for (int k = 0; k < K; ++k) acc = fma(prow[k], __ldg(orow + k), acc);
        s_e[((size_t)q * Tmax + t) * NP + j] = (float)ldexp(acc, __ldg(a.shift + x));
      }
    }
  }
  __syncthreads();

  const float* A = c_trans32 + N * (kNMax * kNMax);
  const float* pi = c_init32 + N * kNMax;
  // per-pair view handed to the chain runners
  struct PairCtx { int T; const float* se; const unsigned* so; double* num; int* ex; };
  auto pair_ctx = [&](int q) {
    PairCtx c;
    c.T = a.phone_off[pair0 + q + 1] - a.phone_off[pair0 + q];
    c.se = s_e + (size_t)q * Tmax * NP;
    c.so = s_off + q * Tmax;
    c.num = s_num + q * NK;
    c.ex = s_exp + q * NK;
    return c;
  };
  auto publish = [&](const PairCtx& c, int i, int k, float lik, int ex) {
    c.num[i * K + k] = (double)lik;
    c.ex[i * K + k] = (lik > 0.0f) ? ex : -0x40000000;
  };
  // one step of a single chain: G = (F A) * e', e'[j] = e_t[j] except e'[clamped] = o.hi + o.lo
  auto chain_step = [&](auto si_tag, const float (&F)[N], float (&G)[N], const float* et, float2 o, int i) {
    constexpr int SI = decltype(si_tag)::value;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      float acc = 0.0f;
#pragma unroll
      for (int l = 0; l < N; ++l) acc = fmaf(F[l], A[l * N + j], acc);
      if (SI >= 0) G[j] = (j == SI) ? fmaf(acc, o.x, acc * o.y) : acc * et[j];
      else G[j] = (j == i) ? fmaf(acc, o.x, acc * o.y) : acc * et[j];
    }
  };
  // power-of-two renormalisation of a chain: exact, the exponent moves to `ex`
  auto renorm = [&](float (&F)[N], int& ex) {
    float m = F[0];
#pragma unroll
    for (int j = 1; j < N; ++j) m = fmaxf(m, F[j]);
    if (m > 0.0f) {
      const int e = ((__float_as_int(m) >> 23) & 0xff) - 127;
      const float sc = __int_as_float((127 - e) << 23);
#pragma unroll
      for (int j = 0; j < N; ++j) F[j] *= sc;
      ex += e;
    }
  };
  // one chain (i, k); SI >= 0: the clamped region is a compile-time constant, SI < 0: per-lane select
  auto run_single = [&](auto si_tag, const PairCtx& c, int i, int k, int& ex) {
    constexpr int SI = decltype(si_tag)::value;
    const int T = c.T;
    const char* po = reinterpret_cast<const char*>(a.obsS) + (k >> 1) * 16 + (k & 1) * 4;
    auto ld_o = [&](int t) {
      const float* q = reinterpret_cast<const float*>(po + c.so[t]);
      return make_float2(__ldg(q), __ldg(q + 2));
    };
    float F[N], G[N];
    {
      const float2 o = ld_o(0);
#pragma unroll
      for (int j = 0; j < N; ++j)
        F[j] = (j == (SI >= 0 ? SI : i)) ? fmaf(pi[j], o.x, pi[j] * o.y) : pi[j] * c.se[j];
    }
    ex = 0;
    int t = 1, trips = 0;
    // two steps per trip (F -> G -> F); a power-of-two renormalisation every 4th trip (scaled emissions are
    // <= 1, so 8 steps cannot overflow and lose at most a few hundred binades of headroom).  The loop is NOT
    // unrolled further: with several code variants per CTA (one per clamped region, single / paired, the
    // tail) a larger body thrashes the instruction cache (ncu: 14 no-instruction stalls per issue with an 8-step body)
#pragma unroll 1
    for (; t + 1 < T; t += 2) {
      const float2 o0 = ld_o(t), o1 = ld_o(t + 1);
      chain_step(si_tag, F, G, c.se + t * NP, o0, i);
      chain_step(si_tag, G, F, c.se + (t + 1) * NP, o1, i);
      if ((++trips & 3) == 0) renorm(F, ex);
    }
    if (t < T) {
      chain_step(si_tag, F, G, c.se + t * NP, ld_o(t), i);
#pragma unroll
      for (int j = 0; j < N; ++j) F[j] = G[j];
    }
    renorm(F, ex);
    float lik = 0.0f;
#pragma unroll
    for (int j = 0; j < N; ++j) lik += F[j];
    return lik;
  };
  // the chains (SI, k) and (SI, k + 1), k even, of one lane in the two halves of packed float32 pairs: every
  // FFMA2 / FMUL2 advances both chains (sm_100a issues a packed FMA at the FMA rate of two scalar ones,
  // profiles/r02_fma_forms_microbench.txt), the transition matrix and the marginal emissions enter as
  // broadcast operands, and one 128-bit load brings both emission pairs
  auto run_pair = [&](auto si_tag, const PairCtx& c, int k, int (&ex)[2], float (&lik)[2]) {
    constexpr int SI = decltype(si_tag)::value;
    const int T = c.T;
    const char* po = reinterpret_cast<const char*>(a.obsS) + (k >> 1) * 16;
    auto ld_o = [&](int t) { return __ldg(reinterpret_cast<const float4*>(po + c.so[t])); };
    auto step2 = [&](const float2 (&Fv)[N], float2 (&Gv)[N], const float* et, const float4 o) {
#pragma unroll
      for (int j = 0; j < N; ++j) {
        float2 acc = mul2(Fv[0], dup2(A[j]));
#pragma unroll
        for (int l = 1; l < N; ++l) acc = fma2(Fv[l], dup2(A[l * N + j]), acc);
        Gv[j] = (j == SI) ? fma2(acc, make_float2(o.x, o.y), mul2(acc, make_float2(o.z, o.w))) : mul2(acc, dup2(et[j]));
      }
    };
    // exact power-of-two renormalisation of both halves, each with its own exponent
    auto renorm2 = [&](float2 (&Fv)[N]) {
      float mx = Fv[0].x, my = Fv[0].y;
#pragma unroll
      for (int j = 1; j < N; ++j) { mx = fmaxf(mx, Fv[j].x); my = fmaxf(my, Fv[j].y); }
      const int e0 = (mx > 0.0f) ? ((__float_as_int(mx) >> 23) & 0xff) - 127 : 0;
      const int e1 = (my > 0.0f) ? ((__float_as_int(my) >> 23) & 0xff) - 127 : 0;
      const float2 sc = make_float2(__int_as_float((127 - e0) << 23), __int_as_float((127 - e1) << 23));
#pragma unroll
      for (int j = 0; j < N; ++j) Fv[j] = mul2(Fv[j], sc);
      ex[0] += e0;
      ex[1] += e1;
    };
    float2 F[N], G[N];
    {
      const float4 o = ld_o(0);
#pragma unroll
      for (int j = 0; j < N; ++j)
        F[j] = (j == SI) ? make_float2(fmaf(pi[j], o.x, pi[j] * o.z), fmaf(pi[j], o.y, pi[j] * o.w)) : dup2(pi[j] * c.se[j]);
    }
    ex[0] = ex[1] = 0;
    int t = 1, trips = 0;
#pragma unroll 1
    for (; t + 1 < T; t += 2) {
      const float4 o0 = ld_o(t), o1 = ld_o(t + 1);
      step2(F, G, c.se + t * NP, o0);
      step2(G, F, c.se + (t + 1) * NP, o1);
      if ((++trips & 3) == 0) renorm2(F);
    }
    if (t < T) {
      step2(F, G, c.se + t * NP, ld_o(t));
#pragma unroll
      for (int j = 0; j < N; ++j) F[j] = G[j];
    }
    renorm2(F);
    lik[0] = lik[1] = 0.0f;
#pragma unroll
    for (int j = 0; j < N; ++j) { lik[0] += F[j].x; lik[1] += F[j].y; }
  };

  const ConceptMap cm = concept_map<N>(K);
  const int si_warps = PPC * cm.wpp, all_warps = concept_block_warps(cm, PPC);
  for (int w = warp_id; w < all_warps; w += nwarps) {
    if (w < si_warps) {
      const int q = w / cm.wpp, ww = w - q * cm.wpp;
      if (q >= npair) continue;
      const PairCtx c = pair_ctx(q);
      if (ww < N * cm.nd) {
        const int i = ww / cm.nd, k = (ww - i * cm.nd) * 64 + 2 * lane_id;
        int ex[2] = {0, 0};
        float lik[2] = {0.0f, 0.0f};
        switch (i) {   // warp-uniform
#define MWD_SI(V) \
  case V:         \
    if constexpr (V < N && N <= 8) run_pair(std::integral_constant<int, V>{}, c, k, ex, lik); \
    break;
          MWD_SI(0) MWD_SI(1) MWD_SI(2) MWD_SI(3) MWD_SI(4) MWD_SI(5) MWD_SI(6) MWD_SI(7)
#undef MWD_SI
          default: break;
        }
        publish(c, i, k, lik[0], ex[0]);
        publish(c, i, k + 1, lik[1], ex[1]);
      } else {
        const int i = ww - N * cm.nd, k = cm.nd * 64 + lane_id;
        int ex = 0;
        float lik = 0.0f;
        switch (i) {   // warp-uniform
#define MWD_SI(V) \
  case V:         \
    if constexpr (V < N && N <= 8) lik = run_single(std::integral_constant<int, V>{}, c, i, k, ex); \
    break;
          MWD_SI(0) MWD_SI(1) MWD_SI(2) MWD_SI(3) MWD_SI(4) MWD_SI(5) MWD_SI(6) MWD_SI(7)
#undef MWD_SI
          default: break;
        }
        publish(c, i, k, lik, ex);
      }
    } else {
      const int cidx = (w - si_warps) * 32 + lane_id;
      const int q = (cm.tpp > 0) ? cidx / cm.tpp : PPC;
      if (q < npair) {
        const int r = cidx - q * cm.tpp;
        const int i = r / cm.rem, k = cm.nd * 64 + cm.ns * 32 + (r - i * cm.rem);
        const PairCtx c = pair_ctx(q);
        int ex = 0;
        const float lik = run_single(std::integral_constant<int, -1>{}, c, i, k, ex);
        publish(c, i, k, lik, ex);
      }
    }
  }
  __syncthreads();
  // per (pair, region): largest chain exponent, then pz * L relative to it in float64, row sum, reciprocal
  for (int row = warp_id; row < npair * N; row += nwarps) {
    const int q = row / N, i = row - q * N;
    const int* ex = s_exp + q * NK + i * K;
    double* num = s_num + q * NK + i * K;
    const double* pzr = s_pz + q * NK + i * K;
    int em = -0x40000000;
    for (int k = lane_id; k < K; k += 32) em = max(em, ex[k]);
    em = __reduce_max_sync(0xffffffffu, em);
    double s = 0.0;
    for (int k = lane_id; k < K; k += 32) {
      const int de = ex[k] - em;      // <= 0; below 2^-1000 of the row's largest chain a term is exactly negligible
      const double v = (de < -1000) ? 0.0 : pzr[k] * (num[k] * __hiloint2double((1023 + de) << 20, 0));
      num[k] = v;
      s += v;
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    const double inv = 1.0 / s;
    double* out = a.cC + (r0 + row) * K;
    for (int k = lane_id; k < K; k += 32) out[k] = num[k] * inv;
  }
}

// shared memory of one CTA of the float32 kernel (the carve-up at the top of the kernel)
template <int N>
static size_t concept32_smem(int K, int Tmax, int ppc, int kseg_n, int P) {
  constexpr int NP = (N <= 4) ? 4 : ((N <= 8) ? 8 : 16);
  return (size_t)ppc * ((size_t)2 * N * K * sizeof(double) + (size_t)Tmax * NP * sizeof(float) +
                        ((size_t)Tmax + (size_t)N * K) * sizeof(int) + (size_t)kseg_n * P * N * sizeof(double)) + 8;
}
// k-segments of the phone-table prologue for a CTA of `threads` threads: as many as keep every thread busy once
// (at most 4); 0 = direct form, for inventories much larger than a caption (the table costs n K P flop per pair,
// the direct form n K T) or too large for shared memory
static int concept32_ksegs(int threads, int P, int Tmax, int N, int ppc) {
  if (P > 2 * Tmax + 32) return 0;
  int S = threads / (ppc * P);
  S = S < 1 ? 1 : (S > 4 ? 4 : S);
  while (S > 1 && (size_t)ppc * S * P * N * sizeof(double) > 24 * 1024) --S;
  if ((size_t)ppc * S * P * N * sizeof(double) > 24 * 1024) return 0;
  return S;
}

template <int N, int PPC>
static int launch_concept32(ConceptArgs32 a, cudaStream_t st) {
  const int64_t npairs = a.hi - a.lo, nblocks = (npairs + PPC - 1) / PPC;
  int threads = 32 * concept_block_warps(concept_map<N>(a.K), PPC);
  if (threads > 1024) threads = 1024;
  a.kseg_n = concept32_ksegs(threads, a.P, a.Tmax, N, PPC);
  const size_t smem = concept32_smem<N>(a.K, a.Tmax, PPC, a.kseg_n, a.P);
  auto kern = ik_concept32_kernel<N, PPC>;
  if (smem > 48 * 1024)
    MWD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MWD_REQUIRE(smem <= 227 * 1024, "concept chains: shared memory %zu exceeds 227 KB (n=%d, T=%d)", smem, N, a.Tmax);
  MWD_REQUIRE(nblocks <= 0x7fffffff, "bucket too large for one launch");
  kern<<<(unsigned)nblocks, threads, smem, st>>>(a);
  MWD_CHECK_LAUNCH();
  return 0;
}

// pairs per CTA: as many as fill the tail warp (K % 32 left-over chains per region), at most 6 and
// only while the CTA stays within 1024 threads and a quarter of the SM's shared memory (MWD_CONCEPT32_PPC overrides)
template <int N>
static int launch_concept32_auto(const ConceptArgs32& a, cudaStream_t st) {
  const ConceptMap cm = concept_map<N>(a.K);
  int ppc = 1;
  static const int forced = getenv("MWD_CONCEPT32_PPC") ? atoi(getenv("MWD_CONCEPT32_PPC")) : 0;
  if (N <= 8 && cm.tpp > 0) {
    auto fits = [&](int q) {
      return 32 * concept_block_warps(cm, q) <= 1024 && concept32_smem<N>(a.K, a.Tmax, q, 4, a.P) <= 100 * 1024;
    };
    auto fill = [&](int q) { return (double)(q * cm.tpp) / (32.0 * ((q * cm.tpp + 31) / 32)); };
    const int cand[4] = {1, 2, 3, 6};
    if (forced >= 1) {
      for (int c = 0; c < 4; ++c)
        if (cand[c] <= forced && fits(cand[c])) ppc = cand[c];
    } else {
      // measured at K = 65, n = 5 (5 tail chains per pair), 1 M pairs: 30.7 / 29.9 / 27.1 / 27.1 ms at 1 / 2 / 3 / 6
      double best = fill(1);
      for (int c = 1; c < 4 && best < 0.9; ++c)
        if (fits(cand[c]) && fill(cand[c]) > best + 0.05) { ppc = cand[c]; best = fill(cand[c]); }
    }
  }
  if constexpr (N <= 8) {
    switch (ppc) {
      case 6: return launch_concept32<N, 6>(a, st);
      case 3: return launch_concept32<N, 3>(a, st);
      case 2: return launch_concept32<N, 2>(a, st);
      default: break;
    }
  }
  return launch_concept32<N, 1>(a, st);
}

template <int N>
static int launch_concept(const ConceptArgs& a, cudaStream_t st) {
  int64_t npairs = a.hi - a.lo;
  int threads = ((N * a.K + 31) / 32) * 32;
  if (threads > 1024) threads = 1024;
  size_t smem = ((size_t)2 * N * a.K + (size_t)a.Tmax * N + kNMax) * sizeof(double) +
                (size_t)a.Tmax * sizeof(int);
  auto kern = ik_concept_kernel<N>;
  if (smem > 48 * 1024)
    MWD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MWD_REQUIRE(npairs <= 0x7fffffff, "bucket too large for one launch");
  kern<<<(unsigned)npairs, threads, smem, st>>>(a);
  MWD_CHECK_LAUNCH();
  return 0;
}

}  // namespace mwd

using namespace mwd;

// library-owned float32 staging of the mixed-precision path (scaled emission table, shifts, tables),
// grown on demand, one set per device
struct Stage32 {
  float4* obsS = nullptr; double* obsKP = nullptr; int32_t* shift = nullptr; float* trans32 = nullptr; float* init32 = nullptr;
  size_t cap = 0;
};
static Stage32 g_stage[16];

// Which buckets take the float32 chains (MWD_MIXED_CONCEPT): measured per bucket at 1 M pairs (ncu launch lists of the
// Flickr and coco10 shapes, profiles/r02_concept_per_bucket.txt) the float32 kernel wins from n = 4 on -- below that the
// n x n products are too small to pay for its prologue / exponent bookkeeping (n = 3: 1.87 vs 1.44 ms) -- and loses
// again for n >= 6 when the concept count leaves a 32-wide single-chain warp per region beside the paired ones
// (K = 100: 2 n + 1 code variants per CTA; n = 8: 6.9 vs 2.8 ms).  MWD_CONCEPT32_MIN_N overrides the lower bound.
static bool concept_bucket_f32(int n, int K) {
  static const int min_n = getenv("MWD_CONCEPT32_MIN_N") ? atoi(getenv("MWD_CONCEPT32_MIN_N")) : 4;
  if (n < min_n) return false;
  const bool single_warps = n <= 8 && (K % 64) >= 32;
  return !(single_warps && n >= 6);
}

static int concept_counts_f32(const mwd_ik_problem* p, cudaStream_t st, int dev) {
  const int K = p->n_concepts, P = p->n_phone_types;
  Stage32& sg = g_stage[dev];
  const size_t need = (size_t)P * K;
  if (sg.cap < need) {
    // growing is rare (first call / a larger phone inventory): settle outstanding work, then reallocate
    MWD_CHECK_CUDA(cudaDeviceSynchronize());
    if (sg.obsS) { cudaFree(sg.obsS); cudaFree(sg.shift); cudaFree(sg.obsKP); }
    MWD_CHECK_CUDA(cudaMalloc(&sg.obsS, (size_t)P * ((K + 1) / 2) * sizeof(float4)));
    MWD_CHECK_CUDA(cudaMalloc(&sg.shift, (size_t)P * sizeof(int32_t)));
    MWD_CHECK_CUDA(cudaMalloc(&sg.obsKP, need * (sizeof(double) + sizeof(float))));   // float64 table, then its float32 copy
    sg.cap = need;
  }
  if (!sg.trans32) {
    MWD_CHECK_CUDA(cudaMalloc(&sg.trans32, sizeof(float) * (kNMax + 1) * kNMax * kNMax));
    MWD_CHECK_CUDA(cudaMalloc(&sg.init32, sizeof(float) * (kNMax + 1) * kNMax));
  }
  concept_prepare32_kernel<<<(P * 32 + 255) / 256 < 8 ? 8 : (P * 32 + 255) / 256, 256, 0, st>>>(
      p->obsT, P, K, sg.obsS, sg.obsKP, reinterpret_cast<float*>(sg.obsKP + sg.cap), sg.shift, p->trans, p->init,
      sg.trans32, sg.init32);
  MWD_CHECK_LAUNCH();
  MWD_CHECK_CUDA(cudaMemcpyToSymbolAsync(c_trans32, sg.trans32, sizeof(float) * (kNMax + 1) * kNMax * kNMax, 0,
                                         cudaMemcpyDeviceToDevice, st));
  MWD_CHECK_CUDA(cudaMemcpyToSymbolAsync(c_init32, sg.init32, sizeof(float) * (kNMax + 1) * kNMax, 0,
                                         cudaMemcpyDeviceToDevice, st));
  for (int b = 0; b < p->n_buckets; ++b) {
    const int n = p->bucket_n[b];
    ConceptArgs32 a;
    a.region_off = p->region_off;
    a.phone_off = p->phone_off;
    a.phones = p->phones;
    a.pz = p->pz;
    a.obsT = p->obsT;
    a.obsS = sg.obsS;
    a.shift = sg.shift;
    a.obsKP = sg.obsKP;
    a.obsKPf = reinterpret_cast<const float*>(sg.obsKP + sg.cap);
    a.P = P;
    a.kseg_n = 0;
    a.trans32 = sg.trans32;
    a.init32 = sg.init32;
    a.cC = p->concept_counts;
    a.lo = p->bucket_lo[b];
    a.hi = p->bucket_lo[b + 1];
    a.K = K;
    a.Tmax = p->bucket_tmax[b];
    if (a.hi <= a.lo || !concept_bucket_f32(n, K)) continue;
    int rc;
    switch (n) {
#define MWD_CASE(NN) case NN: rc = launch_concept32_auto<NN>(a, st); break;
      MWD_CASE(1) MWD_CASE(2) MWD_CASE(3) MWD_CASE(4) MWD_CASE(5) MWD_CASE(6) MWD_CASE(7) MWD_CASE(8)
      MWD_CASE(9) MWD_CASE(10) MWD_CASE(11) MWD_CASE(12) MWD_CASE(13) MWD_CASE(14) MWD_CASE(15)
      MWD_CASE(16)
#undef MWD_CASE
      default:
        set_error("bucket %d: n=%d outside [1,%d]", b, n, MWD_NMAX);
        return 2;
    }
    if (rc) return rc;
  }
  return 0;
}

// only_non_f32: run just the buckets the float32 path leaves to the float64 kernel
static int concept_counts_f64(const mwd_ik_problem* p, cudaStream_t st, bool only_non_f32) {
  // the whole (tiny) parameter tables go to constant memory once per call, device-to-device
  MWD_CHECK_CUDA(cudaMemcpyToSymbolAsync(c_trans, p->trans, sizeof(double) * (kNMax + 1) * kNMax * kNMax,
                                         0, cudaMemcpyDeviceToDevice, st));
  MWD_CHECK_CUDA(cudaMemcpyToSymbolAsync(c_init, p->init, sizeof(double) * (kNMax + 1) * kNMax, 0,
                                         cudaMemcpyDeviceToDevice, st));
  for (int b = 0; b < p->n_buckets; ++b) {
    const int n = p->bucket_n[b];
    ConceptArgs a;
    a.region_off = p->region_off;
    a.phone_off = p->phone_off;
    a.phones = p->phones;
    a.pz = p->pz;
    a.obsT = p->obsT;
    a.cC = p->concept_counts;
    a.lo = p->bucket_lo[b];
    a.hi = p->bucket_lo[b + 1];
    a.K = p->n_concepts;
    a.Tmax = p->bucket_tmax[b];
    if (a.hi <= a.lo || (only_non_f32 && concept_bucket_f32(n, p->n_concepts))) continue;
    int rc;
    switch (n) {
#define MWD_CASE(NN) case NN: rc = launch_concept<NN>(a, st); break;
      MWD_CASE(1) MWD_CASE(2) MWD_CASE(3) MWD_CASE(4) MWD_CASE(5) MWD_CASE(6) MWD_CASE(7) MWD_CASE(8)
      MWD_CASE(9) MWD_CASE(10) MWD_CASE(11) MWD_CASE(12) MWD_CASE(13) MWD_CASE(14) MWD_CASE(15)
      MWD_CASE(16)
#undef MWD_CASE
      default:
        set_error("bucket %d: n=%d outside [1,%d]", b, n, MWD_NMAX);
        return 2;
    }
    if (rc) return rc;
  }
  return 0;
}

extern "C" int mwd_ik_concept_counts(const mwd_ik_problem* p, void* stream) {
  cudaStream_t st = as_stream(stream);
  int dev = 0;
  MWD_CHECK_CUDA(cudaGetDevice(&dev));
  MWD_REQUIRE(dev >= 0 && dev < 16, "device ordinal %d outside [0,16)", dev);
  std::lock_guard<std::mutex> guard(g_const_mutex);
  // under CUDA-graph capture (IKEngine.em_iteration_graph) the cross-stream event chain cannot be recorded: a captured
  // iteration is single-stream by construction, and replays are ordered by the stream they are launched on
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  MWD_CHECK_CUDA(cudaStreamIsCapturing(st, &cap));
  const bool capturing = cap != cudaStreamCaptureStatusNone;
  if (!g_const_event[dev]) MWD_CHECK_CUDA(cudaEventCreateWithFlags(&g_const_event[dev], cudaEventDisableTiming));
  else if (!capturing) MWD_CHECK_CUDA(cudaStreamWaitEvent(st, g_const_event[dev], 0));   // previous user of the constant tables
  // float32 chains need the scaled (P x K) float table: only for a real phone inventory, not for the dense
  // per-frame emission tables of the image-audio classes (P = number of frames)
  const bool f32 = (p->mixed_precision & MWD_MIXED_CONCEPT) && p->part_phone != nullptr &&
                   (int64_t)p->n_phone_types * p->n_concepts <= (1 << 22);
  bool any32 = false, any64 = !f32;
  for (int b = 0; f32 && b < p->n_buckets; ++b) {
    if (p->bucket_lo[b + 1] <= p->bucket_lo[b]) continue;
    if (concept_bucket_f32(p->bucket_n[b], p->n_concepts)) any32 = true;
    else any64 = true;
  }
  int rc = any32 ? concept_counts_f32(p, st, dev) : 0;
  if (rc) return rc;
  rc = any64 ? concept_counts_f64(p, st, f32) : 0;
  if (rc) return rc;
  if (!capturing) MWD_CHECK_CUDA(cudaEventRecord(g_const_event[dev], st));
  return 0;
}

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
#include <mutex>
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
struct ConceptArgs32 {
  const int32_t* region_off;
  const int32_t* phone_off;
  const int32_t* phones;
  const double* pz;
  const double* obsT;       // (P x K) float64: marginal emissions are formed from it in float64
  const float* obsS;        // (P x K) float32, row x pre-scaled by 2^shift[x]
  const int32_t* shift;     // (P)
  const float* trans32;     // float32 copies of the full tables in GLOBAL memory (register-resident path)
  const float* init32;
  double* cC;
  int64_t lo, hi;
  int K, Tmax;
};

// obsS[x][k] = obsT[x][k] * 2^shift[x], shift[x] = -ceil(log2(max_k obsT[x][k])) (0 for an all-zero row);
// one warp per phone type.  Also converts the transition / initial tables to float32.
__global__ void concept_prepare32_kernel(const double* __restrict__ obsT, int P, int K, float* __restrict__ obsS,
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
    for (int k = lane; k < K; k += 32) obsS[(size_t)x * K + k] = (float)ldexp(obsT[(size_t)x * K + k], sh);
  }
}

template <int N>
__global__ void __launch_bounds__(1024) ik_concept32_kernel(const ConceptArgs32 a) {
  constexpr int NP = (N <= 4) ? 4 : ((N <= 8) ? 8 : 16);      // row stride of the emission slab (floats)
  // (measured alternatives, all slower on B200 at 1 M MSCOCO pairs: A in 25 registers via global loads -> 60
  // registers, 3 CTAs/SM, 44.3 ms; the caption's emission rows staged in shared memory -> 35.3 ms; this form,
  // A through the constant bank and o_t[k] from the L1-resident global table: 33.6 ms; float64 kernel: 40.9 ms)
  constexpr bool kRegA = false;
  const int K = a.K;
  const int64_t pair = a.lo + blockIdx.x;
  const int tid = threadIdx.x;
  const int p0 = a.phone_off[pair];
  const int T = a.phone_off[pair + 1] - p0;
  const int64_t r0 = a.region_off[pair];
  const int32_t* ph = a.phones + p0;

  extern __shared__ double smem[];
  double* s_pz = smem;                                    // [N][K]
  double* s_num = s_pz + N * K;                           // [N][K] chain mantissa, then pz * L
  double* s_row = s_num + N * K;                          // [kNMax]
  float* s_e = reinterpret_cast<float*>(s_row + kNMax);   // [Tmax][NP] scaled marginal emissions
  int* s_x = reinterpret_cast<int*>(s_e + (size_t)a.Tmax * NP);   // [Tmax]
  int* s_exp = s_x + a.Tmax;                              // [N][K] chain exponents
  int* s_emax = s_exp + N * K;                            // [kNMax]

  for (int e = tid; e < N * K; e += blockDim.x) s_pz[e] = a.pz[r0 * K + e];
  for (int t = tid; t < T; t += blockDim.x) s_x[t] = ph[t];
  __syncthreads();
  // marginal emissions e[t][j] = sum_k pz[j][k] * obs[k][x_t] in float64 (sequential k, exactly the
  // float64 kernel's values), then scaled like the table row and rounded to float32
  for (int e = tid; e < T * N; e += blockDim.x) {
    int t = e / N, j = e - t * N;
    const int x = s_x[t];
    const double* orow = a.obsT + (size_t)x * K;
    const double* prow = s_pz + j * K;
    double acc = 0.0;
    for (int k = 0; k < K; ++k) acc = fma(prow[k], __ldg(orow + k), acc);
    s_e[t * NP + j] = (float)ldexp(acc, __ldg(a.shift + x));
  }
  __syncthreads();

  const float* A = c_trans32 + N * (kNMax * kNMax);
  const float* pi = c_init32 + N * kNMax;
  // n <= 6: the n x n matrix lives in registers for the whole kernel (three-register FFMA issues at full rate
  // on sm_100a, profiles/microbench/fma_forms.cu).  It is read from GLOBAL memory on purpose: values loaded from
  // the constant bank are re-materialised by the compiler inside the loop (LDCU: 8 % of the issue slots, ncu).
  float Ar[kRegA ? N * N : 1];
  if (kRegA) {
    const float* Ag = a.trans32 + N * (kNMax * kNMax);
#pragma unroll
    for (int q = 0; q < N * N; ++q) Ar[q] = __ldg(Ag + q);
  }
  auto chain_step = [&](auto si_tag, const float (&F)[N], float (&G)[N], const float* et, float o, int i) {
    constexpr int SI = decltype(si_tag)::value;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      float acc = 0.0f;
#pragma unroll
      for (int l = 0; l < N; ++l) acc = fmaf(F[l], kRegA ? Ar[l * N + j] : A[l * N + j], acc);
      if (SI >= 0) G[j] = acc * ((j == SI) ? o : et[j]);
      else G[j] = acc * ((j == i) ? o : et[j]);
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
  auto run_chain = [&](auto si_tag, int i, int k, int& ex) {
    constexpr int SI = decltype(si_tag)::value;
    const float* po = a.obsS + k;
    float F[N], G[N];
    {
      const float o = __ldg(po + (size_t)s_x[0] * K);
#pragma unroll
      for (int j = 0; j < N; ++j) F[j] = pi[j] * ((j == (SI >= 0 ? SI : i)) ? o : s_e[j]);
    }
    ex = 0;
    int t = 1, trips = 0;
    // two steps per trip (F -> G -> F); a power-of-two renormalisation every 4th trip (scaled emissions are
    // <= 1, so 8 steps cannot overflow and lose at most a few hundred binades of headroom).  The loop is NOT
    // unrolled further: with six code variants per CTA (five clamped regions + the generic tail) a larger body
    // thrashes the instruction cache (ncu: 14 no-instruction stalls per issue with an 8-step body)
#pragma unroll 1
    for (; t + 1 < T; t += 2) {
      const float o0 = __ldg(po + (size_t)s_x[t] * K);
      const float o1 = __ldg(po + (size_t)s_x[t + 1] * K);
      chain_step(si_tag, F, G, s_e + t * NP, o0, i);
      chain_step(si_tag, G, F, s_e + (t + 1) * NP, o1, i);
      if ((++trips & 3) == 0) renorm(F, ex);
    }
    if (t < T) {
      const float o0 = __ldg(po + (size_t)s_x[t] * K);
      chain_step(si_tag, F, G, s_e + t * NP, o0, i);
#pragma unroll
      for (int j = 0; j < N; ++j) F[j] = G[j];
    }
    renorm(F, ex);
    float lik = 0.0f;
#pragma unroll
    for (int j = 0; j < N; ++j) lik += F[j];
    return lik;
  };
  const int KF = (N <= 8) ? (K & ~31) : 0;
  for (int c = tid; c < N * K; c += blockDim.x) {
    int i, k, ex = 0;
    float lik;
    if (c < N * KF) {
      i = c / KF;
      k = c - i * KF;
      switch (i) {   // warp-uniform
#define MWD_SI(V) \
  case V:         \
    if constexpr (V < N && N <= 8) lik = run_chain(std::integral_constant<int, V>{}, i, k, ex); else lik = 0.0f; \
    break;
        MWD_SI(0) MWD_SI(1) MWD_SI(2) MWD_SI(3) MWD_SI(4) MWD_SI(5) MWD_SI(6) MWD_SI(7)
#undef MWD_SI
        default: lik = 0.0f;
      }
    } else {
      const int cc = c - N * KF, rem = K - KF;
      i = cc / rem;
      k = KF + (cc - i * rem);
      lik = run_chain(std::integral_constant<int, -1>{}, i, k, ex);
    }
    s_num[i * K + k] = (double)lik;
    s_exp[i * K + k] = (lik > 0.0f) ? ex : -0x40000000;
  }
  __syncthreads();
  // per region: largest chain exponent, then pz * L relative to it in float64, row sum, normalise
  const int warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
  for (int i = warp; i < N; i += nwarp) {
    int em = -0x40000000;
    for (int k = lane; k < K; k += 32) em = max(em, s_exp[i * K + k]);
    em = __reduce_max_sync(0xffffffffu, em);
    double s = 0.0;
    for (int k = lane; k < K; k += 32) {
      const int de = s_exp[i * K + k] - em;
      const double v = (de < -2000) ? 0.0 : s_pz[i * K + k] * ldexp(s_num[i * K + k], de);
      s_num[i * K + k] = v;
      s += v;
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    if (lane == 0) s_row[i] = s;
  }
  __syncthreads();
  for (int c = tid; c < N * K; c += blockDim.x) a.cC[r0 * K + c] = s_num[c] / s_row[c / K];
}

template <int N>
static int launch_concept32(const ConceptArgs32& a, cudaStream_t st) {
  constexpr int NP = (N <= 4) ? 4 : ((N <= 8) ? 8 : 16);
  int64_t npairs = a.hi - a.lo;
  int threads = ((N * a.K + 31) / 32) * 32;
  if (threads > 1024) threads = 1024;
  size_t smem = ((size_t)2 * N * a.K + kNMax) * sizeof(double) + (size_t)a.Tmax * NP * sizeof(float) +
                ((size_t)a.Tmax + (size_t)N * a.K + kNMax) * sizeof(int);
  auto kern = ik_concept32_kernel<N>;
  if (smem > 48 * 1024)
    MWD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MWD_REQUIRE(smem <= 227 * 1024, "concept chains: shared memory %zu exceeds 227 KB (n=%d, T=%d)", smem, N, a.Tmax);
  MWD_REQUIRE(npairs <= 0x7fffffff, "bucket too large for one launch");
  kern<<<(unsigned)npairs, threads, smem, st>>>(a);
  MWD_CHECK_LAUNCH();
  return 0;
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
  float* obsS = nullptr; int32_t* shift = nullptr; float* trans32 = nullptr; float* init32 = nullptr;
  size_t cap = 0;
};
static Stage32 g_stage[16];

static int concept_counts_f32(const mwd_ik_problem* p, cudaStream_t st, int dev) {
  const int K = p->n_concepts, P = p->n_phone_types;
  Stage32& sg = g_stage[dev];
  const size_t need = (size_t)P * K;
  if (sg.cap < need) {
    // growing is rare (first call / a larger phone inventory): settle outstanding work, then reallocate
    MWD_CHECK_CUDA(cudaDeviceSynchronize());
    if (sg.obsS) { cudaFree(sg.obsS); cudaFree(sg.shift); }
    MWD_CHECK_CUDA(cudaMalloc(&sg.obsS, need * sizeof(float)));
    MWD_CHECK_CUDA(cudaMalloc(&sg.shift, (size_t)P * sizeof(int32_t)));
    sg.cap = need;
  }
  if (!sg.trans32) {
    MWD_CHECK_CUDA(cudaMalloc(&sg.trans32, sizeof(float) * (kNMax + 1) * kNMax * kNMax));
    MWD_CHECK_CUDA(cudaMalloc(&sg.init32, sizeof(float) * (kNMax + 1) * kNMax));
  }
  concept_prepare32_kernel<<<(P * 32 + 255) / 256 < 8 ? 8 : (P * 32 + 255) / 256, 256, 0, st>>>(
      p->obsT, P, K, sg.obsS, sg.shift, p->trans, p->init, sg.trans32, sg.init32);
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
    a.trans32 = sg.trans32;
    a.init32 = sg.init32;
    a.cC = p->concept_counts;
    a.lo = p->bucket_lo[b];
    a.hi = p->bucket_lo[b + 1];
    a.K = K;
    a.Tmax = p->bucket_tmax[b];
    if (a.hi <= a.lo) continue;
    int rc;
    switch (n) {
#define MWD_CASE(NN) case NN: rc = launch_concept32<NN>(a, st); break;
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

static int concept_counts_f64(const mwd_ik_problem* p, cudaStream_t st) {
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
    if (a.hi <= a.lo) continue;
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
  int rc = f32 ? concept_counts_f32(p, st, dev) : concept_counts_f64(p, st);
  if (rc) return rc;
  if (!capturing) MWD_CHECK_CUDA(cudaEventRecord(g_const_event[dev], st));
  return 0;
}

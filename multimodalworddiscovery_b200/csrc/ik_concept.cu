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
#include <type_traits>

#include "mwd_common.cuh"

namespace mwd {

__constant__ double c_trans[(kNMax + 1) * kNMax * kNMax];
__constant__ double c_init[(kNMax + 1) * kNMax];

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

extern "C" int mwd_ik_concept_counts(const mwd_ik_problem* p, void* stream) {
  cudaStream_t st = as_stream(stream);
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

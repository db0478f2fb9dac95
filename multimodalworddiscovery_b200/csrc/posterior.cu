// K3 / K4 -- image-posterior emission GEMM + row softmax, and the posterior-parameter gradient.
//
//   K3: softmaxLayer, hmm_dnn/image_phone_hmm_word_discoverer.py:533-541 (linear) and
//       hmm_dnn/image_phone_gaussian_hmm_word_discoverer.py:501-510 (RBF; expanded to
//       (2 v.mu_k - |mu_k|^2)/width, the row constant -|v|^2/width cancels in the softmax).
//       regions x concepts GEMM  pz = softmax_rows([V,1] W^T), float64 accumulate.
//   K4: updateSoftmaxWeight, :475-488 / gaussian :488-499:
//       grad = (conceptCounts - pz)^T [V,1]  (K x (D+1)), split over rows of V into
//       kGradSplits deterministic partials that are then summed in fixed order.
//
// Both are float64 SIMT GEMMs (DFMA pipe): at K=65, D=512 the two GEMMs are ~1/3 of the
// iteration's float64 work; 1e-5 parity on W over tens of EM iterations rules out TF32/BF16
// tensor-core inputs, and B200's FP64 tensor rate equals its DFMA rate.
#include "mwd_common.cuh"

namespace mwd {

constexpr int BM = 64;   // rows (regions) per CTA tile
constexpr int BK = 16;   // reduction chunk
constexpr int TM = 4;    // rows per thread
// thread grid: 16 (columns, tx) x 16 (rows, ty); thread owns rows ty*4..+3, cols tx + 16*jn

template <typename FT>
__device__ __forceinline__ double load_feat(const FT* p) { return (double)(*p); }

// ------------------------------------------------------------------------------ K3
template <int TN, typename FT>
__global__ void __launch_bounds__(256)
posterior_kernel(const FT* __restrict__ feats, int64_t R, int D, const double* __restrict__ W,
                 int K, double* __restrict__ pz) {
  constexpr int KP = 16 * TN;  // padded concept count
  __shared__ double sV[BK][BM + 4];
  __shared__ double sW[BK][KP];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t row0 = (int64_t)blockIdx.x * BM;
  const int ldw = D + 1;

  double acc[TM][TN];
#pragma unroll
  for (int m = 0; m < TM; ++m)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[m][j] = 0.0;

  for (int d0 = 0; d0 < D; d0 += BK) {
    // V tile: 64 rows x 16 d; thread loads 4 elements (row = e/16, dd = e%16 -> coalesced over d)
#pragma unroll
    for (int e = threadIdx.x; e < BM * BK; e += 256) {
      int r = e >> 4, dd = e & 15;
      int64_t gr = row0 + r;
      int d = d0 + dd;
      sV[dd][r] = (gr < R && d < D) ? load_feat(feats + gr * D + d) : 0.0;
    }
    for (int e = threadIdx.x; e < KP * BK; e += 256) {
      int k = e >> 4, dd = e & 15;
      int d = d0 + dd;
      sW[dd][k] = (k < K && d < D) ? W[(size_t)k * ldw + d] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int dd = 0; dd < BK; ++dd) {
      double v[TM], w[TN];
#pragma unroll
      for (int m = 0; m < TM; ++m) v[m] = sV[dd][ty * TM + m];
#pragma unroll
      for (int j = 0; j < TN; ++j) w[j] = sW[dd][tx + 16 * j];
#pragma unroll
      for (int m = 0; m < TM; ++m)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[m][j] = fma(v[m], w[j], acc[m][j]);
    }
    __syncthreads();
  }
  // bias, then exp(x - logsumexp(x)) per row (scipy.special.logsumexp: max-shifted)
#pragma unroll
  for (int m = 0; m < TM; ++m) {
    double mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int k = tx + 16 * j;
      if (k < K) {
        acc[m][j] += W[(size_t)k * ldw + D];
        mx = fmax(mx, acc[m][j]);
      }
    }
#pragma unroll
    for (int s = 8; s > 0; s >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, s));
    double sum = 0.0;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int k = tx + 16 * j;
      if (k < K) sum += exp(acc[m][j] - mx);
    }
#pragma unroll
    for (int s = 8; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
    const double lse = log(sum) + mx;
    const int64_t gr = row0 + ty * TM + m;
    if (gr < R) {
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        int k = tx + 16 * j;
        if (k < K) pz[gr * K + k] = exp(acc[m][j] - lse);
      }
    }
  }
}

// expanded RBF weights: Wexp[k][d] = 2 mu[k][d]/width, Wexp[k][D] = -|mu_k|^2/width
__global__ void gaussian_expand_kernel(const double* __restrict__ mus, int K, int D, double width,
                                       double* __restrict__ Wexp) {
  const int k = blockIdx.x;
  __shared__ double s_part[32];
  double ss = 0.0;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    double m = mus[(size_t)k * D + d];
    Wexp[(size_t)k * (D + 1) + d] = 2.0 * m / width;
    ss = fma(m, m, ss);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, s);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_part[w];
    Wexp[(size_t)k * (D + 1) + D] = -t / width;
  }
}

template <int TN>
static int launch_posterior(const void* feats, int is64, int64_t R, int D, const double* W, int K,
                            double* pz, cudaStream_t st) {
  if (R <= 0) return 0;
  int64_t grid = (R + BM - 1) / BM;
  MWD_REQUIRE(grid <= 0x7fffffff, "too many regions for one launch");
  if (is64)
    posterior_kernel<TN, double><<<(unsigned)grid, 256, 0, st>>>((const double*)feats, R, D, W, K, pz);
  else
    posterior_kernel<TN, float><<<(unsigned)grid, 256, 0, st>>>((const float*)feats, R, D, W, K, pz);
  MWD_CHECK_LAUNCH();
  return 0;
}

static int posterior_dispatch(const void* feats, int is64, int64_t R, int D, const double* W, int K,
                              double* pz, cudaStream_t st) {
  MWD_REQUIRE(K >= 1 && K <= MWD_KMAX, "n_concepts %d outside [1,%d]", K, MWD_KMAX);
  switch ((K + 15) / 16) {
    case 1: return launch_posterior<1>(feats, is64, R, D, W, K, pz, st);
    case 2: return launch_posterior<2>(feats, is64, R, D, W, K, pz, st);
    case 3: return launch_posterior<3>(feats, is64, R, D, W, K, pz, st);
    case 4: return launch_posterior<4>(feats, is64, R, D, W, K, pz, st);
    case 5: return launch_posterior<5>(feats, is64, R, D, W, K, pz, st);
    case 6: return launch_posterior<6>(feats, is64, R, D, W, K, pz, st);
    case 7: return launch_posterior<7>(feats, is64, R, D, W, K, pz, st);
    default: return launch_posterior<8>(feats, is64, R, D, W, K, pz, st);
  }
}

// ------------------------------------------------------------------------------ K4
// partial[s][k][d] = sum_{r in split s} (cC - pz)[r][k] * [V,1][r][d]
constexpr int BD = 64;   // feature columns per CTA tile (thread: 4 consecutive d)
constexpr int BR = 16;   // rows per smem chunk

template <int TN, typename FT>
__global__ void __launch_bounds__(256)
posterior_grad_kernel(const FT* __restrict__ feats, int64_t R, int D, const double* __restrict__ cC,
                      const double* __restrict__ pz, int K, int64_t rows_per_split,
                      double* __restrict__ partial) {
  constexpr int KP = 16 * TN;
  __shared__ double sDl[BR][KP];       // Delta tile
  __shared__ double sV[BR][BD + 4];
  const int tx = threadIdx.x & 15;     // concept direction: k = tx + 16*j
  const int ty = threadIdx.x >> 4;     // feature direction: d = d0 + ty*4 + m
  const int d0 = blockIdx.x * BD;
  const int split = blockIdx.y;
  const int64_t rbeg = (int64_t)split * rows_per_split;
  const int64_t rend = min(R, rbeg + rows_per_split);
  const int ld = D + 1;

  double acc[TN][TM];
#pragma unroll
  for (int j = 0; j < TN; ++j)
#pragma unroll
    for (int m = 0; m < TM; ++m) acc[j][m] = 0.0;

  for (int64_t r0 = rbeg; r0 < rend; r0 += BR) {
    for (int e = threadIdx.x; e < BR * KP; e += 256) {
      int rr = e / KP, k = e - rr * KP;
      int64_t r = r0 + rr;
      sDl[rr][k] = (r < rend && k < K) ? (cC[r * K + k] - pz[r * K + k]) : 0.0;
    }
    for (int e = threadIdx.x; e < BR * BD; e += 256) {
      int rr = e >> 6, dd = e & 63;
      int64_t r = r0 + rr;
      int d = d0 + dd;
      double v = 0.0;
      if (r < rend) {
        if (d < D) v = load_feat(feats + r * D + d);
        else if (d == D) v = 1.0;
      }
      sV[rr][dd] = v;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < BR; ++rr) {
      double dl[TN], v[TM];
#pragma unroll
      for (int j = 0; j < TN; ++j) dl[j] = sDl[rr][tx + 16 * j];
#pragma unroll
      for (int m = 0; m < TM; ++m) v[m] = sV[rr][ty * TM + m];
#pragma unroll
      for (int j = 0; j < TN; ++j)
#pragma unroll
        for (int m = 0; m < TM; ++m) acc[j][m] = fma(dl[j], v[m], acc[j][m]);
    }
    __syncthreads();
  }
  double* out = partial + (size_t)split * K * ld;
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    int k = tx + 16 * j;
    if (k >= K) continue;
#pragma unroll
    for (int m = 0; m < TM; ++m) {
      int d = d0 + ty * TM + m;
      if (d < ld) out[(size_t)k * ld + d] = acc[j][m];
    }
  }
}

__global__ void grad_reduce_kernel(const double* __restrict__ partial, int splits, int64_t elems,
                                   double* __restrict__ grad) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= elems) return;
  double s = 0.0;
  for (int sp = 0; sp < splits; ++sp) s += partial[(size_t)sp * elems + e];
  grad[e] = s;
}

template <int TN>
static int launch_grad(const mwd_ik_problem* p, double* partial, cudaStream_t st) {
  const int D = p->feat_dim, K = p->n_concepts;
  const int64_t R = p->n_regions;
  int64_t rps = (R + kGradSplits - 1) / kGradSplits;
  rps = ((rps + BR - 1) / BR) * BR;
  if (rps < BR) rps = BR;
  dim3 grid((D + 1 + BD - 1) / BD, kGradSplits);
  if (p->feat_is_f64)
    posterior_grad_kernel<TN, double><<<grid, 256, 0, st>>>((const double*)p->feats, R, D,
                                                            p->concept_counts, p->pz, K, rps, partial);
  else
    posterior_grad_kernel<TN, float><<<grid, 256, 0, st>>>((const float*)p->feats, R, D,
                                                           p->concept_counts, p->pz, K, rps, partial);
  MWD_CHECK_LAUNCH();
  return 0;
}

}  // namespace mwd

using namespace mwd;

extern "C" int mwd_posterior_linear(const void* feats, int feat_is_f64, int64_t n_regions,
                                    int feat_dim, const double* W, int n_concepts, double* pz,
                                    void* stream) {
  return posterior_dispatch(feats, feat_is_f64, n_regions, feat_dim, W, n_concepts, pz,
                            as_stream(stream));
}

extern "C" int mwd_posterior_gaussian(const void* feats, int feat_is_f64, int64_t n_regions,
                                      int feat_dim, const double* mus, double width, int n_concepts,
                                      double* w_scratch, double* pz, void* stream) {
  cudaStream_t st = as_stream(stream);
  MWD_REQUIRE(n_concepts >= 1 && n_concepts <= MWD_KMAX, "n_concepts %d outside [1,%d]", n_concepts,
              MWD_KMAX);
  gaussian_expand_kernel<<<n_concepts, 128, 0, st>>>(mus, n_concepts, feat_dim, width, w_scratch);
  MWD_CHECK_LAUNCH();
  return posterior_dispatch(feats, feat_is_f64, n_regions, feat_dim, w_scratch, n_concepts, pz, st);
}

extern "C" int mwd_ik_posterior_grad(const mwd_ik_problem* p, double* grad_partials, double* grad,
                                     void* stream) {
  cudaStream_t st = as_stream(stream);
  const int K = p->n_concepts;
  MWD_REQUIRE(K >= 1 && K <= MWD_KMAX, "n_concepts %d outside [1,%d]", K, MWD_KMAX);
  int rc;
  switch ((K + 15) / 16) {
    case 1: rc = launch_grad<1>(p, grad_partials, st); break;
    case 2: rc = launch_grad<2>(p, grad_partials, st); break;
    case 3: rc = launch_grad<3>(p, grad_partials, st); break;
    case 4: rc = launch_grad<4>(p, grad_partials, st); break;
    case 5: rc = launch_grad<5>(p, grad_partials, st); break;
    case 6: rc = launch_grad<6>(p, grad_partials, st); break;
    case 7: rc = launch_grad<7>(p, grad_partials, st); break;
    default: rc = launch_grad<8>(p, grad_partials, st); break;
  }
  if (rc) return rc;
  const int64_t elems = (int64_t)K * (p->feat_dim + 1);
  grad_reduce_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, st>>>(grad_partials, kGradSplits,
                                                                      elems, grad);
  MWD_CHECK_LAUNCH();
  return 0;
}

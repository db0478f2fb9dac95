// K3 / K4 -- image-posterior emission GEMM + row softmax, and the posterior-parameter gradient.
//
//   K3: softmaxLayer, hmm_dnn/image_phone_hmm_word_discoverer.py:533-541 (linear) and
//       hmm_dnn/image_phone_gaussian_hmm_word_discoverer.py:501-510 (RBF; expanded to
//       (2 v.mu_k - |mu_k|^2)/width, the row constant -|v|^2/width cancels in the softmax).
//       regions x concepts GEMM  pz = softmax_rows([V,1] W^T).
//   K4: updateSoftmaxWeight, :475-488 / gaussian :488-499:
//       grad = (conceptCounts - pz)^T [V,1]  (K x (D+1)), split over rows of V into
//       kGradSplits deterministic partials that are then summed in fixed order.
//
// Both GEMMs run on the FP64 tensor path (mma.sync.m8n8k4.f64, "DMMA"): float64 inputs and
// float64 accumulation, i.e. bit-for-bit the precision class of the reference's NumPy matmul.
// 1e-5 parity of W over tens of EM iterations rules out TF32/BF16 operands; tcgen05 has no
// float64 kind, so mma.sync is the tensor instruction that applies here.  One DMMA retires
// 256 MACs per warp instruction versus 32 for a DFMA, which is what lifts these kernels off the
// issue-slot limit of the SIMT version.
#include "mwd_common.cuh"

namespace mwd {

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

constexpr int BK = 16;        // reduction chunk (feature dims for K3, rows for K4 sub-steps)
constexpr int LDS_PAD = 20;   // row stride (doubles) of a 16-wide fp64 smem tile: 160 B == 32 mod 128

template <typename FT>
__device__ __forceinline__ double load_feat(const FT* p) { return (double)__ldg(p); }

// ------------------------------------------------------------------------------ K3
// CTA = 8 warps; warp w owns rows [w*8*MT, (w+1)*8*MT) of the CTA's 64*MT-row tile and all
// NT*8 concept columns: MT*NT accumulator fragments (2 doubles each).  The next 16-wide chunk of
// V (raw element type) and W is prefetched into registers while the DMMAs of the current chunk
// run; the fp32->fp64 conversion happens only when the registers are staged into shared memory.
// EPI 0: + bias, row softmax (image posterior);  EPI 1: + bias, ReLU (hidden layer of the two-layer
// class, hmm_dnn/image_phone_hmm_dnn_word_discoverer.py:573-579).  ldo = row stride of the output.
// TAIL: K == 8*NT + 1 (the MSCOCO concept count 65 = 8*8 + 1): the last concept column is not padded
// to a ninth 8-wide DMMA tile (11 % of the tensor work for one column) but accumulated with plain
// DFMAs from the A fragments the lanes already hold -- lane (g4,l4) has V[row m*8+g4][kk*4+l4], i.e.
// the 4 lanes of a group cover the chunk of a row; their partial dot products are summed in the epilogue.
template <int NT, int MT, typename FT, int EPI, bool TAIL = false>
__global__ void __launch_bounds__(256, (NT * MT <= 18) ? 2 : 1)
posterior_kernel(const FT* __restrict__ feats, int64_t R, int D, const double* __restrict__ W,
                 int K, double* __restrict__ pz, int ldo) {
  constexpr int BM = 64 * MT;
  constexpr int KP = 8 * NT + (TAIL ? 1 : 0);
  __shared__ double sV[BM * LDS_PAD];
  __shared__ double sW[KP * LDS_PAD];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g4 = lane >> 2, l4 = lane & 3;
  const int64_t row0 = (int64_t)blockIdx.x * BM;
  const int ldw = D + 1;

  double acc[MT][NT][2];
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[m][j][0] = acc[m][j][1] = 0.0;
  double tail[MT];                           // partial logits of concept 8*NT (TAIL)
#pragma unroll
  for (int m = 0; m < MT; ++m) tail[m] = 0.0;

  constexpr int VPT = BM / 16;               // V elements per thread per chunk (row = tid/16 + 16 q)
  constexpr int WPT = (KP + 15) / 16;        // W elements per thread per chunk (k   = tid/16 + 16 q)
  FT vreg[VPT];
  double wreg[WPT];
  const int dd = tid & 15, rr = tid >> 4;
  // per-thread base pointers; rows/concepts past the end are clamped (loaded, then zeroed)
  const FT* vptr[VPT];
  bool vok[VPT];
#pragma unroll
  for (int q = 0; q < VPT; ++q) {
    int64_t gr = row0 + rr + 16 * q;
    vok[q] = gr < R;
    vptr[q] = feats + (vok[q] ? gr : 0) * D + dd;
  }
  const double* wptr[WPT];
  bool wok[WPT];
#pragma unroll
  for (int q = 0; q < WPT; ++q) {
    int k = rr + 16 * q;
    wok[q] = k < K;
    wptr[q] = W + (size_t)(wok[q] ? k : 0) * ldw + dd;
  }

  auto fetch = [&](int d0) {
    const bool dok = d0 + dd < D;
#pragma unroll
    for (int q = 0; q < VPT; ++q) vreg[q] = (vok[q] && dok) ? __ldg(vptr[q] + d0) : FT(0);
#pragma unroll
    for (int q = 0; q < WPT; ++q) wreg[q] = (wok[q] && dok) ? __ldg(wptr[q] + d0) : 0.0;
  };
  auto stage = [&]() {
#pragma unroll
    for (int q = 0; q < VPT; ++q) sV[(rr + 16 * q) * LDS_PAD + dd] = (double)vreg[q];
#pragma unroll
    for (int q = 0; q < WPT; ++q)
      if (rr + 16 * q < KP) sW[(rr + 16 * q) * LDS_PAD + dd] = wreg[q];
  };

  fetch(0);
  for (int d0 = 0; d0 < D; d0 += BK) {
    stage();
    __syncthreads();
    if (d0 + BK < D) fetch(d0 + BK);   // next chunk's global loads fly during the DMMAs
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) {
      double af[MT], bf[NT];
#pragma unroll
      for (int m = 0; m < MT; ++m) af[m] = sV[((warp * MT + m) * 8 + g4) * LDS_PAD + kk * 4 + l4];
#pragma unroll
      for (int j = 0; j < NT; ++j) bf[j] = sW[(j * 8 + g4) * LDS_PAD + kk * 4 + l4];
#pragma unroll
      for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int j = 0; j < NT; ++j) dmma(acc[m][j][0], acc[m][j][1], af[m], bf[j]);
      if (TAIL) {
        const double wt = sW[(8 * NT) * LDS_PAD + kk * 4 + l4];
#pragma unroll
        for (int m = 0; m < MT; ++m) tail[m] = fma(af[m], wt, tail[m]);
      }
    }
    __syncthreads();
  }
  if (TAIL) {   // the 4 lanes of a group hold the partial sums of one row: every lane gets the total
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      tail[m] += __shfl_xor_sync(0xffffffffu, tail[m], 1);
      tail[m] += __shfl_xor_sync(0xffffffffu, tail[m], 2);
    }
  }
  // epilogue: + bias, exp(x - logsumexp(x)) per row (scipy.special.logsumexp is max-shifted).
  // Fragment layout: lane holds row g4, columns j*8 + l4*2 + {0,1}; the 4 lanes l4=0..3 of a
  // group share the row.
  if (EPI == 1) {
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      const int64_t gr = row0 + (warp * MT + m) * 8 + g4;
      if (gr >= R) continue;
#pragma unroll
      for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          int k = j * 8 + l4 * 2 + h;
          if (k < K) {
            const double x = acc[m][j][h] + __ldg(W + (size_t)k * ldw + D);
            pz[gr * ldo + k] = (x > 0.0) ? x : 0.0;
          }
        }
      if (TAIL && l4 == 0) {
        const double x = tail[m] + __ldg(W + (size_t)(8 * NT) * ldw + D);
        pz[gr * ldo + 8 * NT] = (x > 0.0) ? x : 0.0;
      }
    }
    return;
  }
#pragma unroll
  for (int m = 0; m < MT; ++m) {
    double mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        int k = j * 8 + l4 * 2 + h;
        if (k < K) {
          acc[m][j][h] += __ldg(W + (size_t)k * ldw + D);
          mx = fmax(mx, acc[m][j][h]);
        }
      }
    if (TAIL) {
      tail[m] += __ldg(W + (size_t)(8 * NT) * ldw + D);
      mx = fmax(mx, tail[m]);
    }
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    double sum = 0.0;
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        int k = j * 8 + l4 * 2 + h;
        if (k < K) sum += exp(acc[m][j][h] - mx);
      }
    if (TAIL && l4 == 0) sum += exp(tail[m] - mx);    // counted once per row
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    const double lse = log(sum) + mx;
    const int64_t gr = row0 + (warp * MT + m) * 8 + g4;
    if (gr < R) {
#pragma unroll
      for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          int k = j * 8 + l4 * 2 + h;
          if (k < K) pz[gr * ldo + k] = exp(acc[m][j][h] - lse);
        }
      if (TAIL && l4 == 0) pz[gr * ldo + 8 * NT] = exp(tail[m] - lse);
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

template <int NT, int EPI, bool TAIL = false>
static int launch_posterior(const void* feats, int is64, int64_t R, int D, const double* W, int K,
                            double* pz, int ldo, cudaStream_t st) {
  if (R <= 0) return 0;
  constexpr int MT = (NT <= 9) ? 2 : 1;
  int64_t grid = (R + 64 * MT - 1) / (64 * MT);
  MWD_REQUIRE(grid <= 0x7fffffff, "too many regions for one launch");
  if (is64)
    posterior_kernel<NT, MT, double, EPI, TAIL><<<(unsigned)grid, 256, 0, st>>>((const double*)feats, R, D, W, K, pz, ldo);
  else
    posterior_kernel<NT, MT, float, EPI, TAIL><<<(unsigned)grid, 256, 0, st>>>((const float*)feats, R, D, W, K, pz, ldo);
  MWD_CHECK_LAUNCH();
  return 0;
}

template <int EPI>
static int posterior_dispatch_epi(const void* feats, int is64, int64_t R, int D, const double* W, int K,
                                  double* pz, int ldo, cudaStream_t st) {
  MWD_REQUIRE(K >= 1 && K <= MWD_KMAX, "%d output columns outside [1,%d]", K, MWD_KMAX);
  // K = 8*NT + 1 for the concept counts in use (65 = MSCOCO, 33 = tests): tail-column variant
  if (K == 65) return launch_posterior<8, EPI, true>(feats, is64, R, D, W, K, pz, ldo, st);
  if (K == 33) return launch_posterior<4, EPI, true>(feats, is64, R, D, W, K, pz, ldo, st);
  switch ((K + 7) / 8) {
#define MWD_NT(T) case T: return launch_posterior<T, EPI>(feats, is64, R, D, W, K, pz, ldo, st);
    MWD_NT(1) MWD_NT(2) MWD_NT(3) MWD_NT(4) MWD_NT(5) MWD_NT(6) MWD_NT(7) MWD_NT(8)
    MWD_NT(9) MWD_NT(10) MWD_NT(11) MWD_NT(12) MWD_NT(13) MWD_NT(14) MWD_NT(15) MWD_NT(16)
#undef MWD_NT
  }
  set_error("%d output columns not supported", K);
  return 2;
}

static int posterior_dispatch(const void* feats, int is64, int64_t R, int D, const double* W, int K,
                              double* pz, cudaStream_t st) {
  return posterior_dispatch_epi<0>(feats, is64, R, D, W, K, pz, K, st);
}

// ------------------------------------------------------------------------------ K4
// partial[s][k][d] = sum_{r in split s} (cC - pz)[r][k] * [V,1][r][d]
// CTA tile: all K concepts (MT8 = ceil(K/8) m-tiles) x 128 feature columns; warp w owns the 16
// feature columns [w*16, w*16+16) (2 n-tiles) and all m-tiles: 2*MT8 accumulator fragments, 2+MT8
// fragment loads per 2*MT8 DMMAs.  The reduction runs over rows in chunks of 16 with register
// prefetch of the next chunk.  The bias column d == D ([V,1]'s ones) is the column sum of Delta;
// the blockIdx.x == 0 CTAs accumulate it from the staged Delta tile.
constexpr int BD = 128;
constexpr int BR = 16;

// TAIL: K == 8*MT8 + 1: the last Delta row is accumulated with DFMAs against the B fragments instead of
// padding a ninth m-tile (see posterior_kernel).
template <int MT8, typename FT, bool TAIL = false>
__global__ void __launch_bounds__(256, (MT8 <= 9) ? 2 : 1)
posterior_grad_kernel(const FT* __restrict__ feats, int64_t R, int D, const double* __restrict__ cC,
                      const double* __restrict__ pz, int ldc, int K, int64_t rows_per_split,
                      double* __restrict__ partial, int accumulate) {
  constexpr int KP = 8 * MT8 + (TAIL ? 1 : 0);
  constexpr int LDD = 8 * (MT8 + (TAIL ? 1 : 0)) + 4;   // Delta tile [BR][KP], padded: 608 B == 96 mod 128 (KP=72)
  constexpr int LDV = BD + 4;            // V tile [BR][BD], padded: 1056 B == 32 mod 128
  __shared__ double sDl[BR * LDD];
  __shared__ double sV[BR * LDV];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g4 = lane >> 2, l4 = lane & 3;
  const int d0 = blockIdx.x * BD;
  const int split = blockIdx.y;
  const int64_t rbeg = (int64_t)split * rows_per_split;
  const int64_t rend = min(R, rbeg + rows_per_split);
  const int ld = D + 1;
  const bool do_bias = (blockIdx.x == 0);

  double acc[MT8][2][2];
#pragma unroll
  for (int m = 0; m < MT8; ++m)
#pragma unroll
    for (int j = 0; j < 2; ++j) acc[m][j][0] = acc[m][j][1] = 0.0;
  double bias_acc = 0.0;
  double tail[2] = {0.0, 0.0};           // TAIL: partial grad[8*MT8][warp*16 + j*8 + g4] over r == l4 (mod 4)

  // Delta tile: BR*KP elements, thread handles element e = tid + 256 q -> (row e / KP, k e % KP)
  constexpr int DPT = (BR * KP + 255) / 256;
  // V tile: BR*BD elements, thread handles column tid & 127, rows (tid >> 7) + 2 q
  constexpr int VPT = BR / 2;
  double creg[DPT], preg[DPT];   // raw cC / pz; subtracted only at staging time
  FT vreg[VPT];
  const int vcol = tid & 127, vrow = tid >> 7;
  const bool vcol_ok = d0 + vcol < D;

  auto fetch = [&](int64_t r0) {
#pragma unroll
    for (int q = 0; q < DPT; ++q) {
      int e = tid + 256 * q;
      int rr = e / KP, k = e - rr * KP;
      int64_t r = r0 + rr;
      const bool ok = (e < BR * KP && r < rend && k < K);
      creg[q] = ok ? __ldg(cC + r * ldc + k) : 0.0;
      preg[q] = (ok && pz) ? __ldg(pz + r * ldc + k) : 0.0;
    }
#pragma unroll
    for (int q = 0; q < VPT; ++q) {
      int64_t r = r0 + vrow + 2 * q;
      vreg[q] = (vcol_ok && r < rend) ? __ldg(feats + r * D + d0 + vcol) : FT(0);
    }
  };
  auto stage = [&]() {
#pragma unroll
    for (int q = 0; q < DPT; ++q) {
      int e = tid + 256 * q;
      int rr = e / KP, k = e - rr * KP;
      if (e < BR * KP) sDl[rr * LDD + k] = creg[q] - preg[q];
    }
#pragma unroll
    for (int q = 0; q < VPT; ++q) sV[(vrow + 2 * q) * LDV + vcol] = (double)vreg[q];
  };

  if (rbeg < rend) fetch(rbeg);
  for (int64_t r0 = rbeg; r0 < rend; r0 += BR) {
    stage();
    __syncthreads();
    if (r0 + BR < rend) fetch(r0 + BR);
#pragma unroll
    for (int kk = 0; kk < BR / 4; ++kk) {
      // A = Delta^T fragment: A[k = m*8+g4][r = kk*4+l4];  B = V fragment: B[r = kk*4+l4][d = warp*16 + j*8 + g4]
      double bf[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) bf[j] = sV[(kk * 4 + l4) * LDV + warp * 16 + j * 8 + g4];
#pragma unroll
      for (int m = 0; m < MT8; ++m) {
        const double af = sDl[(kk * 4 + l4) * LDD + m * 8 + g4];
#pragma unroll
        for (int j = 0; j < 2; ++j) dmma(acc[m][j][0], acc[m][j][1], af, bf[j]);
      }
      if (TAIL) {
        const double dt = sDl[(kk * 4 + l4) * LDD + 8 * MT8];
#pragma unroll
        for (int j = 0; j < 2; ++j) tail[j] = fma(dt, bf[j], tail[j]);
      }
    }
    if (do_bias && tid < KP) {
#pragma unroll
      for (int rr = 0; rr < BR; ++rr) bias_acc += sDl[rr * LDD + tid];
    }
    __syncthreads();
  }
  // C fragment: row k = m*8 + g4, cols d = d0 + warp*16 + j*8 + l4*2 + {0,1}
  double* out = partial + (size_t)split * K * ld;
#pragma unroll
  for (int m = 0; m < MT8; ++m) {
    int k = m * 8 + g4;
    if (k >= K) continue;
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        int d = d0 + warp * 16 + j * 8 + l4 * 2 + h;
        if (d < D) {
          double* o = out + (size_t)k * ld + d;
          *o = accumulate ? (*o + acc[m][j][h]) : acc[m][j][h];
        }
      }
  }
  if (TAIL) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      tail[j] += __shfl_xor_sync(0xffffffffu, tail[j], 1);
      tail[j] += __shfl_xor_sync(0xffffffffu, tail[j], 2);
      const int d = d0 + warp * 16 + j * 8 + g4;
      if (l4 == 0 && d < D) {
        double* o = out + (size_t)(8 * MT8) * ld + d;
        *o = accumulate ? (*o + tail[j]) : tail[j];
      }
    }
  }
  if (do_bias && tid < K) {
    double* o = out + (size_t)tid * ld + D;
    *o = accumulate ? (*o + bias_acc) : bias_acc;
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

struct GradIn {
  const void* feats; int feat_is_f64; int64_t R; int D;
  const double* delta; const double* minus; int ldc; int K;
};

// Row splits of the gradient GEMM: kGradSplits x 4 d-tiles fill the machine at D = 512; narrower
// feature matrices (fewer d-tiles) get proportionally more row splits so the grid stays ~296 CTAs.
static int grad_splits_for(int D) {
  const int dt = (D + BD - 1) / BD;
  int m = 4 / (dt < 1 ? 1 : dt);
  if (m < 1) m = 1;
  return kGradSplits * m;
}

template <int MT8, bool TAIL = false>
static int launch_grad(const GradIn& g, double* partial, int accumulate, cudaStream_t st) {
  const int D = g.D, K = g.K;
  const int64_t R = g.R;
  const int splits = grad_splits_for(D);
  int64_t rps = (R + splits - 1) / splits;
  rps = ((rps + BR - 1) / BR) * BR;
  if (rps < BR) rps = BR;
  dim3 grid((D + BD - 1) / BD, splits);
  if (g.feat_is_f64)
    posterior_grad_kernel<MT8, double, TAIL><<<grid, 256, 0, st>>>((const double*)g.feats, R, D, g.delta, g.minus,
                                                             g.ldc, K, rps, partial, accumulate);
  else
    posterior_grad_kernel<MT8, float, TAIL><<<grid, 256, 0, st>>>((const float*)g.feats, R, D, g.delta, g.minus,
                                                            g.ldc, K, rps, partial, accumulate);
  MWD_CHECK_LAUNCH();
  return 0;
}

// eps[r][h] = (hidden[r][h] > 0) * sum_k (cC - pz)[r][k] * W[k][h]   (ReLU back-propagation,
// hmm_dnn/image_phone_hmm_dnn_word_discoverer.py:510-511,526)
__global__ void backprop_hidden_kernel(const double* __restrict__ cC, const double* __restrict__ pz,
                                       const double* __restrict__ W, const double* __restrict__ hidden,
                                       int64_t R, int K, int H, double* __restrict__ eps) {
  extern __shared__ double s_d[];          // Delta rows of this CTA: [rows][K]
  const int rows = blockDim.y;
  const int64_t r0 = (int64_t)blockIdx.x * rows;
  for (int e = threadIdx.y * blockDim.x + threadIdx.x; e < rows * K; e += rows * blockDim.x) {
    int rr = e / K, k = e - rr * K;
    int64_t r = r0 + rr;
    s_d[e] = (r < R) ? (cC[r * K + k] - pz[r * K + k]) : 0.0;
  }
  __syncthreads();
  const int64_t r = r0 + threadIdx.y;
  if (r >= R) return;
  const double* d = s_d + threadIdx.y * K;
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    double acc = 0.0;
    for (int k = 0; k < K; ++k) acc = fma(d[k], __ldg(W + (size_t)k * (H + 1) + h), acc);
    eps[r * H + h] = (hidden[r * H + h] > 0.0) ? acc : 0.0;
  }
}

__global__ void sgd_update_kernel(const double* __restrict__ grad, int64_t elems, double scale, double lr,
                                  double momentum, double* __restrict__ param) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= elems) return;
  param[e] = (1.0 - momentum) * param[e] + lr * (scale * grad[e]);
}

}  // namespace mwd

using namespace mwd;

static int grad_generic(const GradIn& g, double* grad_partials, int accumulate, cudaStream_t st) {
  const int K = g.K;
  MWD_REQUIRE(K >= 1 && K <= MWD_KMAX, "%d gradient rows outside [1,%d]", K, MWD_KMAX);
  int rc = 2;
  if (K == 65) return launch_grad<8, true>(g, grad_partials, accumulate, st);
  if (K == 33) return launch_grad<4, true>(g, grad_partials, accumulate, st);
  switch ((K + 7) / 8) {
#define MWD_MT(T) case T: rc = launch_grad<T>(g, grad_partials, accumulate, st); break;
    MWD_MT(1) MWD_MT(2) MWD_MT(3) MWD_MT(4) MWD_MT(5) MWD_MT(6) MWD_MT(7) MWD_MT(8)
    MWD_MT(9) MWD_MT(10) MWD_MT(11) MWD_MT(12) MWD_MT(13) MWD_MT(14) MWD_MT(15) MWD_MT(16)
#undef MWD_MT
  }
  return rc;
}

static int grad_partials_impl(const mwd_ik_problem* p, double* grad_partials, int accumulate,
                              cudaStream_t st) {
  GradIn g{p->feats, p->feat_is_f64, p->n_regions, p->feat_dim, p->concept_counts, p->pz, p->n_concepts,
           p->n_concepts};
  return grad_generic(g, grad_partials, accumulate, st);
}


extern "C" int mwd_hidden_relu(const void* feats, int feat_is_f64, int64_t n_regions, int feat_dim,
                               const double* V, int hidden_dim, double* hidden, void* stream) {
  cudaStream_t st = as_stream(stream);
  for (int c0 = 0; c0 < hidden_dim; c0 += MWD_KMAX) {     // column chunks of <= 128 hidden units
    const int cols = hidden_dim - c0 < MWD_KMAX ? hidden_dim - c0 : MWD_KMAX;
    int rc = posterior_dispatch_epi<1>(feats, feat_is_f64, n_regions, feat_dim, V + (size_t)c0 * (feat_dim + 1),
                                       cols, hidden + c0, hidden_dim, st);
    if (rc) return rc;
  }
  return 0;
}

extern "C" int mwd_backprop_hidden(const double* concept_counts, const double* pz, const double* W,
                                   const double* hidden, int64_t n_regions, int n_concepts, int hidden_dim,
                                   double* eps, void* stream) {
  if (n_regions <= 0) return 0;
  dim3 block(64, 4);
  const size_t smem = (size_t)block.y * n_concepts * sizeof(double);
  backprop_hidden_kernel<<<(unsigned)((n_regions + block.y - 1) / block.y), block, smem, as_stream(stream)>>>(
      concept_counts, pz, W, hidden, n_regions, n_concepts, hidden_dim, eps);
  MWD_CHECK_LAUNCH();
  return 0;
}

extern "C" int64_t mwd_outer_grad_partials_len(int n_rows_out, int feat_dim) {
  const int rows = n_rows_out < MWD_KMAX ? n_rows_out : MWD_KMAX;
  return (int64_t)grad_splits_for(feat_dim) * rows * (feat_dim + 1);
}

extern "C" int mwd_outer_grad(const void* feats, int feat_is_f64, int64_t n_regions, int feat_dim,
                              const double* delta, const double* minus, int n_rows_out, double* grad_partials,
                              double* grad, void* stream) {
  cudaStream_t st = as_stream(stream);
  const int ld = feat_dim + 1;
  for (int c0 = 0; c0 < n_rows_out; c0 += MWD_KMAX) {      // row chunks of <= 128 output rows
    const int rows = n_rows_out - c0 < MWD_KMAX ? n_rows_out - c0 : MWD_KMAX;
    GradIn g{feats, feat_is_f64, n_regions, feat_dim, delta + c0, minus ? minus + c0 : nullptr, n_rows_out, rows};
    int rc = grad_generic(g, grad_partials, 0, st);
    if (rc) return rc;
    const int64_t elems = (int64_t)rows * ld;
    grad_reduce_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, st>>>(grad_partials, grad_splits_for(feat_dim),
                                                                        elems, grad + (size_t)c0 * ld);
    MWD_CHECK_LAUNCH();
  }
  return 0;
}

extern "C" int mwd_sgd_update(double* param, const double* grad, int64_t elems, double scale, double lr,
                              double momentum, void* stream) {
  sgd_update_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, as_stream(stream)>>>(grad, elems, scale, lr,
                                                                                    momentum, param);
  MWD_CHECK_LAUNCH();
  return 0;
}


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

extern "C" int mwd_ik_posterior_grad_partial(const mwd_ik_problem* p, double* grad_partials,
                                             int accumulate, void* stream) {
  return grad_partials_impl(p, grad_partials, accumulate, as_stream(stream));
}

extern "C" int mwd_ik_posterior_grad_finish(int n_concepts, int feat_dim, const double* grad_partials,
                                            double* grad, void* stream) {
  const int64_t elems = (int64_t)n_concepts * (feat_dim + 1);
  grad_reduce_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, as_stream(stream)>>>(
      grad_partials, grad_splits_for(feat_dim), elems, grad);
  MWD_CHECK_LAUNCH();
  return 0;
}

extern "C" int mwd_ik_posterior_grad(const mwd_ik_problem* p, double* grad_partials, double* grad,
                                     void* stream) {
  int rc = grad_partials_impl(p, grad_partials, 0, as_stream(stream));
  if (rc) return rc;
  return mwd_ik_posterior_grad_finish(p->n_concepts, p->feat_dim, grad_partials, grad, stream);
}

// K6 -- Viterbi align + cluster (+ row argmax), and the dense single-pair forward / backward
// used by the class methods forward() / backward().
//
//   align   : hmm_dnn/image_phone_hmm_word_discoverer.py:543-584
//   cluster : :586-597
//   argmax of conceptCountsA rows: printAlignment :628
//   forward : :276-304     backward : :314-335
//
// The Viterbi recursion is done in the raw float64 probability domain with the reference's
// exact multiplication order ((scores[i]*A[i][j])*p[t][j]), strict '>' scans so the first index
// wins ties exactly like np.argmax, and the EPS score floor -- the integer outputs are bit-exact
// whenever the marginal emissions p[t][j] are (they are computed as a sequential-k FMA chain,
// the order of a GEMM micro-kernel).
#include "mwd_common.cuh"

namespace mwd {

struct DecodeArgs {
  const int32_t* region_off;
  const int32_t* phone_off;
  const int32_t* phones;
  const double* pz;
  const double* init;    // full tables
  const double* trans;
  const double* obsT;
  int32_t* alignment;
  double* align_probs;   // may be null
  const int64_t* ap_off;
  int32_t* image_concepts;
  double* cluster_scores;   // may be null: R x K
  int64_t n_pairs;
  int K, Tmax, floor_norm, given_alignment;
};

__device__ __forceinline__ bool np_greater(double cand, double best) {
  // np.argmax / np.max semantics: the first NaN wins and sticks
  return (cand > best) || (cand != cand && best == best);
}

__global__ void __launch_bounds__(128) ik_decode_kernel(const DecodeArgs a) {
  const int64_t pair = blockIdx.x;
  const int tid = threadIdx.x;
  const int K = a.K;
  const int p0 = a.phone_off[pair];
  const int T = a.phone_off[pair + 1] - p0;
  const int64_t r0 = a.region_off[pair];
  const int n = (int)(a.region_off[pair + 1] - r0);
  const int32_t* ph = a.phones + p0;

  extern __shared__ double smem[];
  double* s_pz = smem;                              // [n][K]
  double* s_p = s_pz + kNMax * K;                   // [Tmax][n]
  double* s_sc = s_p + (size_t)a.Tmax * kNMax;      // [2][NMAX]
  int* s_x = reinterpret_cast<int*>(s_sc + 2 * kNMax);     // [Tmax]
  unsigned char* s_bp = reinterpret_cast<unsigned char*>(s_x) + (((size_t)a.Tmax * sizeof(int) + 7) & ~(size_t)7);  // [Tmax][NMAX]
  // cluster scores [n][K] live in their own region after the back-pointers (8-byte aligned)
  double* s_cl = reinterpret_cast<double*>(s_bp + (((size_t)a.Tmax * kNMax + 7) & ~(size_t)7));

  for (int e = tid; e < n * K; e += blockDim.x) s_pz[e] = a.pz[r0 * K + e];
  for (int t = tid; t < T; t += blockDim.x) s_x[t] = ph[t];
  __syncthreads();
  // p[t][i] = sum_k pz[i][k] obs[k][x_t]   (:550)
  for (int e = tid; e < T * n; e += blockDim.x) {
    int t = e / n, i = e - t * n;
    const double* orow = a.obsT + (size_t)s_x[t] * K;
    const double* prow = s_pz + i * K;
    double acc = 0.0;
    for (int k = 0; k < K; ++k) acc = fma(prow[k], __ldg(orow + k), acc);
    s_p[t * n + i] = acc;
  }
  __syncthreads();

  const double* A = a.trans + (size_t)n * MWD_TRANS_STRIDE;
  const double* pi = a.init + (size_t)n * MWD_INIT_STRIDE;
  double* ap = a.align_probs ? a.align_probs + a.ap_off[pair] : nullptr;

  if (tid < 32 && !a.given_alignment) {
    const int j = tid;
    double sc = 0.0;
    if (j < n) {
      sc = pi[j] * s_p[j];                                       // :555
      s_sc[j] = sc;
      if (ap) ap[j] = sc;                                        // :559 (raw, un-normalised)
    }
    __syncwarp();
    int cur = 0;
    for (int t = 1; t < T; ++t) {
      const double* prev = s_sc + ((t - 1) & 1) * kNMax;
      double* next = s_sc + (t & 1) * kNMax;
      if (j < n) {
        const double pt = s_p[t * n + j];
        double best = __dmul_rn(__dmul_rn(prev[0], A[j]), pt);   // (scores[i]*A[i][j])*p[t][j]
        int arg = 0;
        for (int i = 1; i < n; ++i) {
          double cand = __dmul_rn(__dmul_rn(prev[i], A[i * n + j]), pt);
          if (np_greater(cand, best)) { best = cand; arg = i; }
        }
        s_bp[t * kNMax + j] = (unsigned char)arg;                // :562
        sc = (a.floor_norm & 2) ? best : floor_eps(best);        // :564 (two-layer :612 does not floor)
        next[j] = sc;
      }
      __syncwarp();
      if (ap && j < n) {
        double tot = 0.0;
        for (int i = 0; i < n; ++i) tot += (a.floor_norm & 1) ? floor_eps(next[i]) : next[i];
        ap[(size_t)t * n + j] = sc / tot;                        // :571 / gaussian :583
      }
    }
    if (j == 0) {
      const double* fin = s_sc + ((T - 1) & 1) * kNMax;
      double best = fin[0];
      for (int i = 1; i < n; ++i)
        if (np_greater(fin[i], best)) { best = fin[i]; cur = i; }
      a.alignment[p0 + T - 1] = cur;                             // :576-582
      for (int t = T - 1; t > 0; --t) {
        cur = s_bp[t * kNMax + cur];
        a.alignment[p0 + t - 1] = cur;
      }
    }
  }
  __syncthreads();
  // cluster (:586-597): scores[i][k] = pz[i][k] * prod_{t: align[t]==i} obs[k][x_t], in t order
  for (int e = tid; e < n * K; e += blockDim.x) {
    int i = e / K, k = e - i * K;
    double sc = s_pz[e];
    for (int t = 0; t < T; ++t)
      if (a.alignment[p0 + t] == i) sc = __dmul_rn(sc, __ldg(a.obsT + (size_t)s_x[t] * K + k));
    s_cl[e] = sc;
    if (a.cluster_scores) a.cluster_scores[r0 * K + e] = sc;
  }
  __syncthreads();
  if (tid < n) {
    const double* row = s_cl + tid * K;
    double best = row[0];
    int arg = 0;
    for (int k = 1; k < K; ++k)
      if (np_greater(row[k], best)) { best = row[k]; arg = k; }
    a.image_concepts[r0 + tid] = arg;
  }
}

__global__ void argmax_rows_kernel(const double* __restrict__ rows, int64_t n_rows, int n_cols,
                                   int32_t* __restrict__ out) {
  // one warp per row, lanes strided over columns; ties -> lowest index
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  const double* r = rows + row * n_cols;
  double best = 0.0;
  int arg = -1;
  for (int k = lane; k < n_cols; k += 32) {
    double v = r[k];
    if (arg < 0 || np_greater(v, best)) { best = v; arg = k; }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    double ob = __shfl_xor_sync(0xffffffffu, best, s);
    int oa = __shfl_xor_sync(0xffffffffu, arg, s);
    if (oa >= 0) {
      bool take;
      if (arg < 0) take = true;
      else if (best != best) take = (ob != ob) && oa < arg;          // both NaN: lower index
      else if (ob != ob) take = true;                                // NaN beats number
      else take = (ob > best) || (ob == best && oa < arg);
      if (take) { best = ob; arg = oa; }
    }
  }
  if (lane == 0) out[row] = arg;
}

// ---------------------------------------------------------------- dense single-pair sweeps
__global__ void ik_forward_dense_kernel(const double* __restrict__ pz, const int32_t* __restrict__ ph,
                                        int T, int n, int K, const double* __restrict__ init,
                                        const double* __restrict__ trans,
                                        const double* __restrict__ obsT, double* __restrict__ out) {
  __shared__ double s_s[kNMax], s_c[kNMax];
  const int tid = threadIdx.x;
  const double* A = trans + (size_t)n * MWD_TRANS_STRIDE;
  const double* pi = init + (size_t)n * MWD_INIT_STRIDE;
  for (int e = tid; e < n * K; e += blockDim.x) {
    int i = e / K, k = e - i * K;
    out[e] = (pi[i] * pz[e]) * obsT[(size_t)ph[0] * K + k];
  }
  __syncthreads();
  for (int t = 0; t + 1 < T; ++t) {
    const double* cur = out + (size_t)t * n * K;
    double* nxt = out + (size_t)(t + 1) * n * K;
    if (tid < n) {
      double s = 0.0;
      for (int k = 0; k < K; ++k) s += cur[tid * K + k];
      s_s[tid] = s;
    }
    __syncthreads();
    if (tid < n) {
      double c = 0.0;
      for (int j = 0; j < n; ++j)
        if (j != tid) c = fma(A[j * n + tid], s_s[j], c);
      s_c[tid] = c;
    }
    __syncthreads();
    for (int e = tid; e < n * K; e += blockDim.x) {
      int i = e / K, k = e - i * K;
      double o = obsT[(size_t)ph[t + 1] * K + k];
      nxt[e] = (A[i * n + i] * cur[e]) * o + s_c[i] * (pz[e] * o);
    }
    __syncthreads();
  }
}

__global__ void ik_backward_dense_kernel(const double* __restrict__ pz, const int32_t* __restrict__ ph,
                                         int T, int n, int K, const double* __restrict__ trans,
                                         const double* __restrict__ obsT, double* __restrict__ out) {
  __shared__ double s_r[kNMax], s_w[kNMax];
  const int tid = threadIdx.x;
  const double* A = trans + (size_t)n * MWD_TRANS_STRIDE;
  for (int e = tid; e < n * K; e += blockDim.x) out[(size_t)(T - 1) * n * K + e] = 1.0;
  __syncthreads();
  for (int t = T - 1; t > 0; --t) {
    const double* cur = out + (size_t)t * n * K;
    double* prv = out + (size_t)(t - 1) * n * K;
    const double* orow = obsT + (size_t)ph[t] * K;
    if (tid < n) {
      double r = 0.0;
      for (int k = 0; k < K; ++k) r += cur[tid * K + k] * (pz[tid * K + k] * orow[k]);
      s_r[tid] = r;
    }
    __syncthreads();
    if (tid < n) {
      double w = 0.0;
      for (int j = 0; j < n; ++j)
        if (j != tid) w = fma(A[tid * n + j], s_r[j], w);
      s_w[tid] = w;
    }
    __syncthreads();
    for (int e = tid; e < n * K; e += blockDim.x) {
      int i = e / K, k = e - i * K;
      prv[e] = A[i * n + i] * (cur[e] * orow[k]) + s_w[i];
    }
    __syncthreads();
  }
}

}  // namespace mwd

using namespace mwd;

extern "C" int mwd_ik_decode(const mwd_ik_problem* p, int floor_norm, int given_alignment,
                             int32_t* alignment, double* align_probs, const int64_t* ap_off,
                             int32_t* image_concepts, double* cluster_scores, void* stream) {
  cudaStream_t st = as_stream(stream);
  if (p->n_pairs <= 0) return 0;
  MWD_REQUIRE(p->n_pairs <= 0x7fffffff, "too many pairs for one launch");
  MWD_REQUIRE(align_probs == nullptr || ap_off != nullptr, "align_probs needs ap_off");
  DecodeArgs a;
  a.region_off = p->region_off;
  a.phone_off = p->phone_off;
  a.phones = p->phones;
  a.pz = p->pz;
  a.init = p->init;
  a.trans = p->trans;
  a.obsT = p->obsT;
  a.alignment = alignment;
  a.align_probs = align_probs;
  a.ap_off = ap_off;
  a.image_concepts = image_concepts;
  a.n_pairs = p->n_pairs;
  a.K = p->n_concepts;
  a.Tmax = p->t_max;
  a.floor_norm = floor_norm;
  a.given_alignment = given_alignment;
  a.cluster_scores = cluster_scores;
  // s_pz [NMAX][K] | s_p [Tmax][NMAX] | s_sc [2][NMAX] | s_x [Tmax] ints | s_bp [Tmax][NMAX] bytes | s_cl [NMAX][K]
  size_t smem = ((size_t)kNMax * a.K + (size_t)a.Tmax * kNMax + 2 * kNMax) * sizeof(double) +
                (((size_t)a.Tmax * sizeof(int) + 7) & ~(size_t)7) + (((size_t)a.Tmax * kNMax + 7) & ~(size_t)7) +
                (size_t)kNMax * a.K * sizeof(double);
  MWD_REQUIRE(smem <= 227 * 1024, "decode shared memory %zu exceeds 227 KB (t_max=%d)", smem, a.Tmax);
  if (smem > 48 * 1024)
    MWD_CHECK_CUDA(cudaFuncSetAttribute(ik_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
  ik_decode_kernel<<<(unsigned)p->n_pairs, 128, smem, st>>>(a);
  MWD_CHECK_LAUNCH();
  return 0;
}

extern "C" int mwd_argmax_rows(const double* rows, int64_t n_rows, int n_cols, int32_t* out,
                               void* stream) {
  if (n_rows <= 0) return 0;
  int64_t grid = (n_rows + 7) / 8;
  MWD_REQUIRE(grid <= 0x7fffffff, "too many rows");
  argmax_rows_kernel<<<(unsigned)grid, 256, 0, as_stream(stream)>>>(rows, n_rows, n_cols, out);
  MWD_CHECK_LAUNCH();
  return 0;
}

extern "C" int mwd_ik_forward_dense(const double* pz_pair, const int32_t* phones_pair, int T, int n,
                                    int K, const double* init, const double* trans,
                                    const double* obsT, double* out, void* stream) {
  MWD_REQUIRE(T >= 1 && n >= 1 && n <= kNMax, "bad shape T=%d n=%d", T, n);
  ik_forward_dense_kernel<<<1, 256, 0, as_stream(stream)>>>(pz_pair, phones_pair, T, n, K, init, trans,
                                                            obsT, out);
  MWD_CHECK_LAUNCH();
  return 0;
}

extern "C" int mwd_ik_backward_dense(const double* pz_pair, const int32_t* phones_pair, int T, int n,
                                     int K, const double* trans, const double* obsT, double* out,
                                     void* stream) {
  MWD_REQUIRE(T >= 1 && n >= 1 && n <= kNMax, "bad shape T=%d n=%d", T, n);
  ik_backward_dense_kernel<<<1, 256, 0, as_stream(stream)>>>(pz_pair, phones_pair, T, n, K, trans, obsT,
                                                             out);
  MWD_CHECK_LAUNCH();
  return 0;
}

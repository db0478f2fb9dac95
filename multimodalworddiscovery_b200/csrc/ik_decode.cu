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

// One WARP owns one pair (persistent warps, fixed pair -> warp stride): no block barriers, the
// Viterbi scores live in lanes j < n and travel by shuffles, back-pointers / marginal emissions /
// the alignment sit in the warp's private shared-memory slab.  Arithmetic order is unchanged from
// the reference: sequential-k FMA chain for p[t][i], (scores[i]*A[i][j])*p[t][j] scanned in i with
// strict '>', products of cluster() in t order.
constexpr int kDecWarps = 4;   // warps per CTA

// NN > 0: every pair of the launch has n == NN regions (one launch per bucket; the state loops unroll
// without predicates); NN == 0: n is read per pair (single-pair calls, n > 10).
template <int NN>
__global__ void __launch_bounds__(kDecWarps * 32) ik_decode_kernel(const DecodeArgs a, int64_t lo, int64_t hi, int nmax,
                                                                   int slab_bytes) {
  constexpr int NU = NN > 0 ? NN : kNMax;      // unroll bound of the state loops
  const int lane = threadIdx.x & 31;
  const int wic = threadIdx.x >> 5;
  const int K = a.K;
  extern __shared__ double smem[];
  unsigned char* slab = reinterpret_cast<unsigned char*>(smem) + (size_t)wic * slab_bytes;
  double* s_pz = reinterpret_cast<double*>(slab);                    // [nmax][K]  pz, later cluster scores
  double* s_p = s_pz + (size_t)nmax * K;                             // [Tmax][nmax] marginal emissions
  int* s_x = reinterpret_cast<int*>(s_p + (size_t)a.Tmax * nmax);    // [Tmax] phone ids
  int* s_ali = s_x + a.Tmax;                                         // [Tmax] alignment
  unsigned char* s_bp = reinterpret_cast<unsigned char*>(s_ali + a.Tmax);   // [Tmax][nmax] back-pointers

  const int64_t gw = (int64_t)blockIdx.x * kDecWarps + wic;
  const int64_t nwarps = (int64_t)gridDim.x * kDecWarps;
  for (int64_t pair = lo + gw; pair < hi; pair += nwarps) {
    const int p0 = a.phone_off[pair];
    const int T = a.phone_off[pair + 1] - p0;
    const int64_t r0 = a.region_off[pair];
    const int n = NN > 0 ? NN : (int)(a.region_off[pair + 1] - r0);
    const int32_t* ph = a.phones + p0;
    __syncwarp();   // the previous pair's slab is no longer read
    for (int e = lane; e < n * K; e += 32) s_pz[e] = a.pz[r0 * K + e];
    for (int t = lane; t < T; t += 32) s_x[t] = ph[t];
    if (a.given_alignment)
      for (int t = lane; t < T; t += 32) s_ali[t] = a.alignment[p0 + t];
    __syncwarp();
    const double* A = a.trans + (size_t)n * MWD_TRANS_STRIDE;
    const double* pi = a.init + (size_t)n * MWD_INIT_STRIDE;
    double* ap = a.align_probs ? a.align_probs + a.ap_off[pair] : nullptr;

    if (!a.given_alignment) {
      // p[t][i] = sum_k pz[i][k] obs[k][x_t]   (:550), one (t,i) item per lane
      // (four items per lane in flight: each item is a K-long dependent FMA chain)
      for (int e0 = lane; e0 < T * n; e0 += 128) {
        const double* orow[4];
        const double* prow[4];
        double acc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = min(e0 + 32 * u, T * n - 1);
          const int t = e / n, i = e - t * n;
          orow[u] = a.obsT + (size_t)s_x[t] * K;
          prow[u] = s_pz + i * K;
          acc[u] = 0.0;
        }
        for (int k = 0; k < K; ++k) {
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[u] = fma(prow[u][k], __ldg(orow[u] + k), acc[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (e0 + 32 * u < T * n) s_p[e0 + 32 * u] = acc[u];
      }
      __syncwarp();
      const int j = lane;
      const bool on = j < n;
      double acol[NU];                       // A[i][j], i < n
#pragma unroll
      for (int i = 0; i < NU; ++i) acol[i] = (on && i < n) ? A[i * n + j] : 0.0;
      double sc = on ? pi[j] * s_p[j] : 0.0;                         // :555
      if (ap && on) ap[j] = sc;                                      // :559 (raw, un-normalised)
      for (int t = 1; t < T; ++t) {
        const double pt = on ? s_p[t * n + j] : 0.0;
        double best = 0.0;
        int arg = 0;
#pragma unroll
        for (int i = 0; i < NU; ++i) {
          if (NN > 0 || i < n) {                                     // warp-uniform
            const double prev = __shfl_sync(0xffffffffu, sc, i);
            const double cand = __dmul_rn(__dmul_rn(prev, acol[i]), pt);   // (scores[i]*A[i][j])*p[t][j]
            if (i == 0 || np_greater(cand, best)) { best = cand; arg = i; }
          }
        }
        if (on) s_bp[t * n + j] = (unsigned char)arg;                // :562
        sc = (a.floor_norm & 2) ? best : floor_eps(best);            // :564 (two-layer :612 does not floor)
        if (ap) {                                                    // warp-uniform
          double tot = 0.0;
#pragma unroll
          for (int i = 0; i < NU; ++i) {
            if (NN > 0 || i < n) {
              const double v = __shfl_sync(0xffffffffu, sc, i);
              tot += (a.floor_norm & 1) ? floor_eps(v) : v;
            }
          }
          if (on) ap[(size_t)t * n + j] = sc / tot;                  // :571 / gaussian :583
        }
      }
      // final argmax over states (first index on ties), then the back-trace by lane 0
      int cur = 0;
      {
        double best = __shfl_sync(0xffffffffu, sc, 0);
#pragma unroll
        for (int i = 1; i < NU; ++i) {
          if (NN > 0 || i < n) {
            const double v = __shfl_sync(0xffffffffu, sc, i);
            if (np_greater(v, best)) { best = v; cur = i; }
          }
        }
      }
      __syncwarp();                                                  // back-pointers visible to lane 0
      if (lane == 0) {
        s_ali[T - 1] = cur;                                          // :576-582
        for (int t = T - 1; t > 0; --t) {
          cur = s_bp[t * n + cur];
          s_ali[t - 1] = cur;
        }
      }
      __syncwarp();
      for (int t = lane; t < T; t += 32) a.alignment[p0 + t] = s_ali[t];
    }
    // cluster (:586-597): scores[i][k] = pz[i][k] * prod_{t: align[t]==i} obs[k][x_t], in t order
    // time-major: step t multiplies the K scores of the aligned region by obs[:, x_t] (lanes over k);
    // every (i,k) product still runs in t order, each lane only ever touches its own columns
    if constexpr (NN > 0 && NN <= 6) {
      // scores of all NN regions in registers (4 column slots per lane cover K <= 128); the aligned
      // region of a step is warp-uniform, so picking its register row is a uniform branch
      double sc[NN][4];
#pragma unroll
      for (int i = 0; i < NN; ++i)
#pragma unroll
        for (int m = 0; m < 4; ++m) sc[i][m] = (lane + 32 * m < K) ? s_pz[i * K + lane + 32 * m] : 0.0;
      for (int t = 0; t < T; ++t) {
        const int ai = s_ali[t];
        const double* orow = a.obsT + (size_t)s_x[t] * K + lane;
        double o[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) o[m] = (lane + 32 * m < K) ? __ldg(orow + 32 * m) : 0.0;
#pragma unroll
        for (int i = 0; i < NN; ++i)
          if (ai == i) {
#pragma unroll
            for (int m = 0; m < 4; ++m) sc[i][m] = __dmul_rn(sc[i][m], o[m]);
          }
      }
#pragma unroll
      for (int i = 0; i < NN; ++i)
#pragma unroll
        for (int m = 0; m < 4; ++m)
          if (lane + 32 * m < K) s_pz[i * K + lane + 32 * m] = sc[i][m];
    } else {
      for (int t = 0; t < T; ++t) {
        double* row = s_pz + s_ali[t] * K;
        const double* orow = a.obsT + (size_t)s_x[t] * K;
        for (int k = lane; k < K; k += 32) row[k] = __dmul_rn(row[k], __ldg(orow + k));
      }
    }
    __syncwarp();
    if (a.cluster_scores)
      for (int e = lane; e < n * K; e += 32) a.cluster_scores[r0 * K + e] = s_pz[e];
    if (lane < n) {
      const double* row = s_pz + lane * K;
      double best = row[0];
      int arg = 0;
      for (int k = 1; k < K; ++k)
        if (np_greater(row[k], best)) { best = row[k]; arg = k; }
      a.image_concepts[r0 + lane] = arg;
    }
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
  // per-warp slab: s_pz [nmax][K] | s_p [Tmax][nmax] doubles | s_x, s_ali [Tmax] ints | s_bp [Tmax][nmax] bytes
  int nmax = 0;
  for (int b = 0; b < p->n_buckets; ++b)
    if (p->bucket_lo[b + 1] > p->bucket_lo[b] && p->bucket_n[b] > nmax) nmax = p->bucket_n[b];
  if (nmax <= 0 || nmax > kNMax) nmax = kNMax;          // no bucket descriptors (single-pair calls)
  const int Tm = a.Tmax > 0 ? a.Tmax : 1;
  a.Tmax = Tm;
  size_t slab = ((size_t)nmax * a.K + (size_t)Tm * nmax) * sizeof(double) + (size_t)2 * Tm * sizeof(int) +
                (size_t)Tm * nmax;
  slab = (slab + 15) & ~(size_t)15;
  const size_t smem = slab * kDecWarps;
  MWD_REQUIRE(smem <= 227 * 1024, "decode shared memory %zu exceeds 227 KB (t_max=%d)", smem, a.Tmax);
  int per_sm = (int)((size_t)224 * 1024 / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 16 / kDecWarps * 4) per_sm = 16 / kDecWarps * 4;    // at most 64 warps per SM
  auto launch = [&](auto kern, int64_t lo, int64_t hi) -> int {
    if (smem > 48 * 1024)
      MWD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t grid = (int64_t)sm_count() * per_sm;
    const int64_t need = (hi - lo + kDecWarps - 1) / kDecWarps;
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, kDecWarps * 32, smem, st>>>(a, lo, hi, nmax, (int)slab);
    return 0;
  };
  if (p->n_buckets <= 0) {
    int rc = launch(ik_decode_kernel<0>, 0, p->n_pairs);
    if (rc) return rc;
  }
  for (int b = 0; b < p->n_buckets; ++b) {
    const int64_t lo = p->bucket_lo[b], hi = p->bucket_lo[b + 1];
    if (hi <= lo) continue;
    int rc;
    switch (p->bucket_n[b]) {
#define MWD_DN(V) case V: rc = launch(ik_decode_kernel<V>, lo, hi); break;
      MWD_DN(1) MWD_DN(2) MWD_DN(3) MWD_DN(4) MWD_DN(5) MWD_DN(6) MWD_DN(7) MWD_DN(8) MWD_DN(9) MWD_DN(10)
#undef MWD_DN
      default: rc = launch(ik_decode_kernel<0>, lo, hi); break;
    }
    if (rc) return rc;
  }
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

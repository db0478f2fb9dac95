// K5 -- deterministic second-level count reduction and the M-step of trainUsingEM
// (hmm_dnn/image_phone_hmm_word_discoverer.py:238-258, gaussian variant :238-264, parameter
// update of updateSoftmaxWeight :487-488 / gaussian :494-499).
//
// counts buffer layout (float64), the unit that is all-reduced across GPUs:
//   [0, P*K)                      phoneC, transposed: [p][k]
//   [.., +(NMAX+1)*NMAX)          initC[m][i]
//   [.., +(NMAX+1)*NMAX*NMAX)     transC[m][i*m+j]   (un-pooled; Toeplitz pooling happens here)
//   [.., +1)                      sum over pairs of log-likelihood
#include "mwd_common.cuh"

namespace mwd {

// Two-level fixed-order column sum of a [rows][elems] partial table (rows = 1776 on a B200: a single
// serial pass per column is latency bound, 0.6 ms).  Stage 1 (grid.y = kRowGroups): group g sums its
// rows in order and parks the result IN PLACE in its first row; stage 2 sums the group heads in order.
constexpr int kRowGroups = 32;
__global__ void reduce_rows_stage1_kernel(double* __restrict__ part, int rows, int64_t elems) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= elems) return;
  const int per = (rows + kRowGroups - 1) / kRowGroups;
  const int r0 = blockIdx.y * per;
  const int r1 = min(rows, r0 + per);
  if (r0 >= r1) return;
  double s = 0.0;
  for (int r = r0; r < r1; ++r) s += part[(size_t)r * elems + e];
  part[(size_t)r0 * elems + e] = s;
}
__global__ void reduce_rows_stage2_kernel(const double* __restrict__ part, int rows, int64_t elems,
                                          double* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= elems) return;
  const int per = (rows + kRowGroups - 1) / kRowGroups;
  double s = 0.0;
  for (int r0 = 0; r0 < rows; r0 += per) s += part[(size_t)r0 * elems + e];
  out[e] = s;
}
static void reduce_rows(double* part, int rows, int64_t elems, double* out, cudaStream_t st) {
  const unsigned gx = (unsigned)((elems + 255) / 256);
  reduce_rows_stage1_kernel<<<dim3(gx, kRowGroups), 256, 0, st>>>(part, rows, elems);
  reduce_rows_stage2_kernel<<<gx, 256, 0, st>>>(part, rows, elems, out);
}

// fixed-shape two-stage sum of pair_ll: stage 1 = 256 CTAs x 256 threads, strided; stage 2 = 1 CTA
constexpr int kLLBlocks = 256;
__global__ void ll_stage1_kernel(const double* __restrict__ ll, int64_t n, double* __restrict__ blk) {
  __shared__ double s[256];
  double acc = 0.0;
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < n; e += (int64_t)kLLBlocks * 256)
    acc += ll[e];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) s[threadIdx.x] += s[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) blk[blockIdx.x] = s[0];
}
__global__ void ll_stage2_kernel(const double* __restrict__ blk, double* __restrict__ out) {
  __shared__ double s[kLLBlocks];
  s[threadIdx.x] = blk[threadIdx.x];
  __syncthreads();
  for (int w = kLLBlocks / 2; w > 0; w >>= 1) {
    if ((int)threadIdx.x < w) s[threadIdx.x] += s[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = s[0];
}

// fixed-shape deterministic sum of n doubles (blk_scratch: >= kLLBlocks doubles)
int sum_doubles(const double* x, int64_t n, double* blk_scratch, double* out, cudaStream_t st) {
  ll_stage1_kernel<<<kLLBlocks, 256, 0, st>>>(x, n, blk_scratch);
  ll_stage2_kernel<<<1, kLLBlocks, 0, st>>>(blk_scratch, out);
  MWD_CHECK_LAUNCH();
  return 0;
}

// init / trans M-step for one m (one CTA per distinct length)
struct LensArg { int lens[kNMax + 1]; int n; };

__global__ void mstep_init_trans_kernel(LensArg la, int gaussian, int toeplitz, int freeze_trans,
                                        const double* __restrict__ initC,
                                        const double* __restrict__ transC,
                                        double* __restrict__ init, double* __restrict__ trans) {
  const int m = la.lens[blockIdx.x];
  __shared__ double sC[kNMax * kNMax];
  __shared__ double sJ[2 * kNMax];
  __shared__ double sTot[kNMax];
  __shared__ double sI;
  const double* ic = initC + (size_t)m * kNMax;
  const double* tc = transC + (size_t)m * kNMax * kNMax;
  double* io = init + (size_t)m * kNMax;
  double* to = trans + (size_t)m * kNMax * kNMax;
  const int tid = threadIdx.x;
  for (int e = tid; e < m * m; e += blockDim.x) sC[e] = tc[e];
  __syncthreads();
  if (toeplitz) {
    // every [s][s'] receives the sum of its diagonal s'-s (:399-413; linear, so applied to sums)
    for (int dlt = tid; dlt < 2 * m - 1; dlt += blockDim.x) {
      int off = dlt - (m - 1);
      double s = 0.0;
      for (int r = 0; r < m; ++r) {
        int c = r + off;
        if (c >= 0 && c < m) s += sC[r * m + c];
      }
      sJ[dlt] = s;
    }
    __syncthreads();
    for (int e = tid; e < m * m; e += blockDim.x) {
      int r = e / m, c = e - r * m;
      sC[e] = sJ[c - r + m - 1];
    }
    __syncthreads();
  }
  if (tid < m) {
    double t = 0.0;
    for (int c = 0; c < m; ++c) t += gaussian ? floor_eps(sC[tid * m + c]) : sC[tid * m + c];
    sTot[tid] = t;
  }
  if (tid == 0) {
    double t = 0.0;
    for (int c = 0; c < m; ++c) t += gaussian ? floor_eps(ic[c]) : ic[c];
    sI = t;
  }
  __syncthreads();
  if (!freeze_trans)
    for (int e = tid; e < m * m; e += blockDim.x) {
      int r = e / m;
      if (sTot[r] != 0.0) to[e] = (gaussian ? floor_eps(sC[e]) : sC[e]) / sTot[r];
    }
  if (tid < m) io[tid] = (gaussian ? floor_eps(ic[tid]) : ic[tid]) / sI;
}

// obsT[p][k] = phoneC[p][k] / sum_p max(phoneC[p][k], EPS)    (:255-256)
__global__ void mstep_obs_kernel(const double* __restrict__ phoneC, int P, int K, double eps,
                                 double* __restrict__ obsT) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  double norm = 0.0;
  for (int p = 0; p < P; ++p) norm += floor_at(phoneC[(size_t)p * K + k], eps);
  for (int p = 0; p < P; ++p) obsT[(size_t)p * K + k] = phoneC[(size_t)p * K + k] / norm;
}

// W = (1-momentum) W + lr * grad / N          (linear, :483,:487-488)
__global__ void mstep_w_kernel(const double* __restrict__ grad, int64_t elems, double invN, double lr,
                               double momentum, double* __restrict__ W) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= elems) return;
  W[e] = (1.0 - momentum) * W[e] + lr * (invN * grad[e]);
}

// mus = (1-momentum) mus + lr * (grad[:, :D] - grad[:, D] * mus) / (N * width)  (gaussian :494-499)
__global__ void mstep_mus_kernel(const double* __restrict__ grad, int K, int D, double scale,
                                 double lr, double momentum, double* __restrict__ mus) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)K * D) return;
  int k = (int)(e / D), d = (int)(e - (int64_t)k * D);
  double g = grad[(size_t)k * (D + 1) + d] - grad[(size_t)k * (D + 1) + D] * mus[e];
  mus[e] = (1.0 - momentum) * mus[e] + lr * (scale * g);
}

}  // namespace mwd

using namespace mwd;

extern "C" int64_t mwd_ik_counts_len(int K, int P) {
  return (int64_t)P * K + (int64_t)(kNMax + 1) * kNMax + (int64_t)(kNMax + 1) * kNMax * kNMax + 1;
}

extern "C" int mwd_ik_partial_sizes(int K, int P, mwd_partial_sizes* out) {
  int64_t g = estep_grid_rows();
  out->phone_elems = g * P * K;
  out->init_elems = g * (kNMax + 1) * kNMax;
  out->trans_elems = g * (kNMax + 1) * kNMax * kNMax;
  return 0;
}

extern "C" int mwd_ik_reduce_counts(const mwd_ik_problem* p, double* counts, void* stream) {
  cudaStream_t st = as_stream(stream);
  const int rows = estep_grid_rows();
  const int64_t pe = (int64_t)p->n_phone_types * p->n_concepts;
  const int64_t ie = (int64_t)(kNMax + 1) * kNMax;
  const int64_t te = (int64_t)(kNMax + 1) * kNMax * kNMax;
  // (the partial tables are consumed: stage 1 overwrites the group-head rows in place)
  if (p->part_phone)
    reduce_rows(p->part_phone, rows, pe, counts, st);
  else   // dense-emission classes fill counts[0:pe] themselves (mwd_concept_phone_counts)
    MWD_CHECK_CUDA(cudaMemsetAsync(counts, 0, (size_t)pe * sizeof(double), st));
  reduce_rows(p->part_init, rows, ie, counts + pe, st);
  reduce_rows(p->part_trans, rows, te, counts + pe + ie, st);
  MWD_CHECK_LAUNCH();
  // log-likelihood: stage-1 partials are parked in the (already consumed) head of part_init
  return sum_doubles(p->pair_ll, p->n_pairs, p->part_init, counts + pe + ie + te, st);
}

extern "C" int mwd_ik_mstep(const mwd_ik_mstep_args* a, void* stream) {
  cudaStream_t st = as_stream(stream);
  const int K = a->n_concepts, P = a->n_phone_types, D = a->feat_dim;
  MWD_REQUIRE(a->n_lens >= 1 && a->n_lens <= kNMax, "n_lens %d outside [1,%d]", a->n_lens, kNMax);
  LensArg la;
  la.n = a->n_lens;
  for (int i = 0; i < a->n_lens; ++i) {
    MWD_REQUIRE(a->lens[i] >= 1 && a->lens[i] <= kNMax, "length %d outside [1,%d]", a->lens[i], kNMax);
    la.lens[i] = a->lens[i];
  }
  const int64_t pe = (int64_t)P * K;
  const int64_t ie = (int64_t)(kNMax + 1) * kNMax;
  const double* phoneC = a->counts;
  const double* initC = a->counts + pe;
  const double* transC = a->counts + pe + ie;
  const int no_floors = (a->flags & MWD_MSTEP_NO_FLOORS) ? 1 : 0;
  const int floor_tables = (!no_floors && (a->gaussian || (a->flags & MWD_MSTEP_FLOOR_TABLES))) ? 1 : 0;
  mstep_init_trans_kernel<<<a->n_lens, 256, 0, st>>>(la, floor_tables, a->toeplitz,
                                                     (a->flags & MWD_MSTEP_FREEZE_TRANS) ? 1 : 0, initC, transC,
                                                     a->init, a->trans);
  mstep_obs_kernel<<<(K + 127) / 128, 128, 0, st>>>(phoneC, P, K, no_floors ? -1.0 : kEps, a->obsT);
  const double invN = 1.0 / (double)a->n_pairs_global;
  if (a->flags & MWD_MSTEP_NO_POSTERIOR) {
    MWD_CHECK_LAUNCH();
    return 0;
  }
  if (a->gaussian) {
    const int64_t elems = (int64_t)K * D;
    const double scale = 1.0 / ((double)a->n_pairs_global * a->width);
    mstep_mus_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, st>>>(a->grad, K, D, scale, a->lr,
                                                                      a->momentum, a->posterior_param);
  } else {
    const int64_t elems = (int64_t)K * (D + 1);
    mstep_w_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, st>>>(a->grad, elems, invN, a->lr,
                                                                    a->momentum, a->posterior_param);
  }
  MWD_CHECK_LAUNCH();
  return 0;
}

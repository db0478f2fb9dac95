// Dense-emission pieces of ImageAudioHMMWordDiscoverer (SURVEY 8 f2,
// hmm_dnn/image_audio_hmm_word_discoverer.py).  The class is the image-phone HMM with the discrete
// emission obs[:, x_t] replaced by
//     E[t][k] = sum_ph phoneProbs[k][ph] * p(ph | a_t)            (:286-288, :320-322, ...)
// so the recursion / concept / decode kernels run unchanged on `obsT = E` (Ttot x K) with the
// identity "phone id" sequence x_t = global frame index; what is new is
//   * mwd_dense_emission        -- E from the frame posteriors and the phone table,
//   * mwd_concept_phone_counts  -- updateConceptPhoneCounts (:486-493) summed over t into the
//                                  (concept x phone) count table of trainUsingEM :230-231.
#include "mwd_common.cuh"

// Both operations are GEMMs over the frames and go through the FP64 tensor-path kernels of
// posterior.cu: E = [PH, 1] . [phoneProbs | 0]^T is the hidden-layer GEMM (its ReLU epilogue is the
// identity on non-negative sums), the concept-phone counts are the outer-product gradient GEMM
// cAn^T . [PH, 1] with deterministic row-split partials.

namespace mwd {

// V[k][p] = ppT[p][k] (p < nP), V[k][nP] = 0: the phone table as a (K x (nP+1)) weight matrix
__global__ void phone_table_as_weights_kernel(const double* __restrict__ ppT, int nP, int K,
                                              double* __restrict__ V) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= K * (nP + 1)) return;
  const int k = e / (nP + 1), p = e - k * (nP + 1);
  V[e] = (p < nP) ? ppT[p * K + k] : 0.0;
}

// cA[t][:] /= (sum_k cA[t] * sum_p ph[t]) -- the normaliser of the per-frame outer product (:491);
// one warp per frame, lanes stride the row, fixed butterfly order.
__global__ void __launch_bounds__(256) cpc_normalise_kernel(double* __restrict__ cA, const double* __restrict__ ph,
                                                            int64_t T, int K, int nP) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarp = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t t = warp; t < T; t += nwarp) {
    double* row = cA + t * K;
    const double* prow = ph + t * nP;
    double a = 0.0, b = 0.0;
    for (int e = lane; e < K; e += 32) a += row[e];
    for (int e = lane; e < nP; e += 32) b += prow[e];
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, m);
      b += __shfl_xor_sync(0xffffffffu, b, m);
    }
    const double sc = 1.0 / (a * b);
    for (int e = lane; e < K; e += 32) row[e] *= sc;
  }
}

// counts_t[p][k] = grad[k][p]  (grad: K x (nP+1), last column = bias sums, dropped)
__global__ void cpc_transpose_kernel(const double* __restrict__ grad, int K, int nP, double* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= K * nP) return;
  const int p = e / K, k = e - p * K;
  out[e] = grad[k * (nP + 1) + p];
}

}  // namespace mwd

using namespace mwd;

extern "C" int mwd_hidden_relu(const void* feats, int feat_is_f64, int64_t n_regions, int feat_dim,
                               const double* V, int hidden_dim, double* hidden, void* stream);
extern "C" int mwd_outer_grad(const void* feats, int feat_is_f64, int64_t n_regions, int feat_dim,
                              const double* delta, const double* minus, int n_rows_out, double* grad_partials,
                              double* grad, void* stream);
extern "C" int64_t mwd_outer_grad_partials_len(int n_rows_out, int feat_dim);

extern "C" int mwd_dense_emission(const double* frame_post, const double* phone_probs_t, int64_t n_frames,
                                  int n_phones, int n_concepts, double* v_scratch, double* emis, void* stream) {
  MWD_REQUIRE(n_concepts >= 1 && n_concepts <= MWD_KMAX, "n_concepts %d outside [1,%d]", n_concepts, MWD_KMAX);
  MWD_REQUIRE(n_phones >= 1 && n_phones <= 128, "n_phones %d outside [1,128]", n_phones);
  if (n_frames <= 0) return 0;
  const int elems = n_concepts * (n_phones + 1);
  phone_table_as_weights_kernel<<<(elems + 255) / 256, 256, 0, as_stream(stream)>>>(phone_probs_t, n_phones,
                                                                                   n_concepts, v_scratch);
  MWD_CHECK_LAUNCH();
  // relu(x) == x here: every term of the sum is a product of probabilities
  return mwd_hidden_relu(frame_post, 1, n_frames, n_phones, v_scratch, n_concepts, emis, stream);
}

extern "C" int64_t mwd_concept_phone_partials_len(int n_concepts, int n_phones) {
  return mwd_outer_grad_partials_len(n_concepts, n_phones) + (int64_t)n_concepts * (n_phones + 1);
}

extern "C" int mwd_concept_phone_counts(double* concept_counts_a, const double* frame_post, int64_t n_frames,
                                        int n_concepts, int n_phones, double* partials, double* counts_t,
                                        void* stream) {
  const int K = n_concepts, nP = n_phones;
  MWD_REQUIRE(K >= 1 && K <= MWD_KMAX, "n_concepts %d outside [1,%d]", K, MWD_KMAX);
  MWD_REQUIRE(nP >= 1 && nP <= 128, "n_phones %d outside [1,128]", nP);
  cudaStream_t st = as_stream(stream);
  if (n_frames <= 0) {
    MWD_CHECK_CUDA(cudaMemsetAsync(counts_t, 0, sizeof(double) * K * nP, st));
    return 0;
  }
  int64_t blocks = (n_frames + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  cpc_normalise_kernel<<<(unsigned)blocks, 256, 0, st>>>(concept_counts_a, frame_post, n_frames, K, nP);
  MWD_CHECK_LAUNCH();
  double* grad = partials + mwd_outer_grad_partials_len(K, nP);
  int rc = mwd_outer_grad(frame_post, 1, n_frames, nP, concept_counts_a, nullptr, K, partials, grad, stream);
  if (rc) return rc;
  cpc_transpose_kernel<<<(K * nP + 255) / 256, 256, 0, st>>>(grad, K, nP, counts_t);
  MWD_CHECK_LAUNCH();
  return 0;
}

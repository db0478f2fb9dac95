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

namespace mwd {

// E[t][k] = sum_p ph[t][p] * ppT[p][k]; ppT (nP x K, the obsT layout) staged in shared memory.
__global__ void __launch_bounds__(256) dense_emission_kernel(const double* __restrict__ ph,
                                                             const double* __restrict__ ppT, int64_t T,
                                                             int nP, int K, double* __restrict__ E) {
  extern __shared__ double s_pp[];
  for (int e = threadIdx.x; e < nP * K; e += blockDim.x) s_pp[e] = ppT[e];
  __syncthreads();
  const int64_t total = T * K;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = e / K;
    const int k = (int)(e - t * K);
    const double* row = ph + t * nP;
    double acc = 0.0;
    for (int p = 0; p < nP; ++p) acc = fma(__ldg(row + p), s_pp[p * K + k], acc);
    E[e] = acc;
  }
}

constexpr int kCpcChunks = 296;   // time chunks (one CTA each) of the concept-phone count reduction
constexpr int kCpcTile = 16;      // frames staged per step
constexpr int kCpcR = 16;         // max phones per thread

// partial[chunk][p][k] = sum over the chunk's frames of cA[t][k] * ph[t][p] / (sum_k cA[t] * sum_p ph[t]):
// the per-frame normalised outer product of :490-491, accumulated in frame order.
// Thread (g, k) owns concept k and phones p = g + PG * r.
__global__ void __launch_bounds__(1024) concept_phone_partial_kernel(const double* __restrict__ cA,
                                                                     const double* __restrict__ ph, int64_t T,
                                                                     int K, int nP, double* __restrict__ part) {
  extern __shared__ double sm[];
  double* s_c = sm;                         // [tile][K]
  double* s_p = s_c + kCpcTile * K;         // [tile][nP]
  double* s_sc = s_p + kCpcTile * nP;       // [tile]
  const int PG = blockDim.x / K;            // phone groups
  const int g = threadIdx.x / K, k = threadIdx.x - g * K;
  const bool owner = g < PG;
  const int64_t per = (T + gridDim.x - 1) / gridDim.x;
  const int64_t t_lo = (int64_t)blockIdx.x * per;
  const int64_t t_hi = (t_lo + per < T) ? t_lo + per : T;
  double acc[kCpcR];
#pragma unroll
  for (int r = 0; r < kCpcR; ++r) acc[r] = 0.0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int64_t t0 = t_lo; t0 < t_hi; t0 += kCpcTile) {
    const int len = (int)((t_hi - t0 < kCpcTile) ? t_hi - t0 : kCpcTile);
    __syncthreads();
    for (int e = threadIdx.x; e < len * K; e += blockDim.x) s_c[e] = cA[t0 * K + e];
    for (int e = threadIdx.x; e < len * nP; e += blockDim.x) s_p[e] = ph[t0 * nP + e];
    __syncthreads();
    for (int tt = warp; tt < len; tt += nwarp) {       // 1 / (sum_k cA[t] * sum_p ph[t]), one warp per frame
      double a = 0.0, b = 0.0;
      for (int e = lane; e < K; e += 32) a += s_c[tt * K + e];
      for (int e = lane; e < nP; e += 32) b += s_p[tt * nP + e];
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, m);
        b += __shfl_xor_sync(0xffffffffu, b, m);
      }
      if (lane == 0) s_sc[tt] = 1.0 / (a * b);
    }
    __syncthreads();
    if (owner) {
      for (int tt = 0; tt < len; ++tt) {
        const double c = s_c[tt * K + k] * s_sc[tt];
#pragma unroll
        for (int r = 0; r < kCpcR; ++r) {
          const int p = g + PG * r;
          if (p < nP) acc[r] = fma(c, s_p[tt * nP + p], acc[r]);
        }
      }
    }
  }
  if (owner) {
    double* out = part + (size_t)blockIdx.x * nP * K;
#pragma unroll
    for (int r = 0; r < kCpcR; ++r) {
      const int p = g + PG * r;
      if (p < nP) out[p * K + k] = acc[r];
    }
  }
}

__global__ void cpc_reduce_kernel(const double* __restrict__ part, int rows, int64_t elems,
                                  double* __restrict__ out) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= elems) return;
  double s = 0.0;
  for (int r = 0; r < rows; ++r) s += part[(size_t)r * elems + e];
  out[e] = s;
}

}  // namespace mwd

using namespace mwd;

extern "C" int mwd_dense_emission(const double* frame_post, const double* phone_probs_t, int64_t n_frames,
                                  int n_phones, int n_concepts, double* emis, void* stream) {
  MWD_REQUIRE(n_concepts >= 1 && n_concepts <= MWD_KMAX, "n_concepts %d outside [1,%d]", n_concepts, MWD_KMAX);
  MWD_REQUIRE(n_phones >= 1 && n_phones <= 128, "n_phones %d outside [1,128]", n_phones);
  if (n_frames <= 0) return 0;
  const size_t smem = (size_t)n_phones * n_concepts * sizeof(double);
  if (smem > 48 * 1024)
    MWD_CHECK_CUDA(cudaFuncSetAttribute(dense_emission_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t blocks = (n_frames * n_concepts + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  dense_emission_kernel<<<(unsigned)blocks, 256, smem, as_stream(stream)>>>(frame_post, phone_probs_t, n_frames,
                                                                           n_phones, n_concepts, emis);
  MWD_CHECK_LAUNCH();
  return 0;
}

extern "C" int64_t mwd_concept_phone_partials_len(int n_concepts, int n_phones) {
  return (int64_t)kCpcChunks * n_concepts * n_phones;
}

extern "C" int mwd_concept_phone_counts(const double* concept_counts_a, const double* frame_post, int64_t n_frames,
                                        int n_concepts, int n_phones, double* partials, double* counts_t,
                                        void* stream) {
  const int K = n_concepts, nP = n_phones;
  MWD_REQUIRE(K >= 1 && K <= MWD_KMAX, "n_concepts %d outside [1,%d]", K, MWD_KMAX);
  MWD_REQUIRE(nP >= 1 && nP <= 128, "n_phones %d outside [1,128]", nP);
  cudaStream_t st = as_stream(stream);
  int threads = 1024;
  MWD_REQUIRE((threads / K) * kCpcR >= nP, "concept-phone counts: %d phones do not fit %d groups x %d", nP,
              threads / K, kCpcR);
  const size_t smem = ((size_t)kCpcTile * (K + nP) + kCpcTile) * sizeof(double);
  concept_phone_partial_kernel<<<kCpcChunks, threads, smem, st>>>(concept_counts_a, frame_post, n_frames, K, nP,
                                                                  partials);
  const int64_t elems = (int64_t)K * nP;
  cpc_reduce_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, st>>>(partials, kCpcChunks, elems, counts_t);
  MWD_CHECK_LAUNCH();
  return 0;
}

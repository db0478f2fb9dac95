// Shared declarations of the mwd_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "mwd_b200.h"

namespace mwd {

void set_error(const char* fmt, ...);

#define MWD_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::mwd::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,                 \
                       cudaGetErrorString(_e));                                           \
      return 1;                                                                           \
    }                                                                                     \
  } while (0)

#define MWD_REQUIRE(cond, ...)                                                            \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      ::mwd::set_error(__VA_ARGS__);                                                      \
      return 2;                                                                           \
    }                                                                                     \
  } while (0)

#define MWD_CHECK_LAUNCH() MWD_CHECK_CUDA(cudaGetLastError())

constexpr double kEps = MWD_EPS;
constexpr int kNMax = MWD_NMAX;
constexpr int kPairsPerCta = 4;   // pairs processed in lock-step by one estep CTA
constexpr int kLanesPerRow = 8;   // threads that share one (pair, region) row of the (i,k) lattice
constexpr int kGradSplits = 74;   // row splits of the posterior-gradient GEMM (x4 d-tiles = 2 CTAs/SM)
constexpr int kEstepCtasPerSm = 16; // rows of the per-CTA (per-warp for K1w: 12, K1w32: 16) partial tables per SM

int sm_count();
int estep_grid_rows();

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// np.maximum(x, EPS): NaN propagates (fmax would swallow it)
__device__ __forceinline__ double floor_eps(double x) { return (x < kEps) ? kEps : x; }
// same with a run-time floor: EPS, or 0 for the classes whose counts are NOT floored
// (image_audio_gaussian_hmm_word_discoverer.py:369-371,414-415,449-451) -- on probabilities a
// floor of 0 is the identity, so the un-floored mode costs no extra instruction
__device__ __forceinline__ double floor_at(double x, double eps) { return (x < eps) ? eps : x; }

__device__ __forceinline__ double shfl_xor_f64(double v, int mask) {
  return __shfl_xor_sync(0xffffffffu, v, mask);
}

// sum over the 8 lanes that share a (pair, region) row; every lane receives the total
__device__ __forceinline__ double row8_sum(double v) {
  v += shfl_xor_f64(v, 4);
  v += shfl_xor_f64(v, 2);
  v += shfl_xor_f64(v, 1);
  return v;
}

}  // namespace mwd

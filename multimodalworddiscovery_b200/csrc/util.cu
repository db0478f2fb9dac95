// Small stream-ordered utilities of the product path, so that one EM iteration launches only
// kernels of this library (+ NCCL): zero / constant fill of the partial-count tables, the
// deterministic sum of the per-pair log-likelihoods, and the fixed-rank-order combination of
// the packed count buffers gathered from all ranks (dist.py).
#include <math.h>

#include "mwd_common.cuh"

namespace mwd {

int sum_doubles(const double* x, int64_t n, double* blk_scratch, double* out, cudaStream_t st);

__global__ void fill_kernel(double* __restrict__ p, int64_t n, double v) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
    p[e] = v;
}

// out[e] = sum_r g[r][e] in rank order (log_domain: logsumexp over ranks, NaN-propagating, the
// plain-sum entry ll_index excepted)
__global__ void rank_reduce_kernel(const double* __restrict__ g, int world, int64_t n, int log_domain,
                                   int64_t ll_index, double* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  if (!log_domain || e == ll_index) {
    double s = 0.0;
    for (int r = 0; r < world; ++r) s += g[(size_t)r * n + e];
    out[e] = s;
    return;
  }
  double m = -INFINITY;
  bool nan = false;
  for (int r = 0; r < world; ++r) {
    const double v = g[(size_t)r * n + e];
    nan |= (v != v);
    m = (v > m) ? v : m;
  }
  if (nan) { out[e] = NAN; return; }
  if (isinf(m)) { out[e] = m; return; }
  double s = 0.0;
  for (int r = 0; r < world; ++r) s += exp(g[(size_t)r * n + e] - m);
  out[e] = m + log(s);
}

}  // namespace mwd

using namespace mwd;

extern "C" int mwd_fill_f64(double* p, int64_t n, double value, void* stream) {
  if (n <= 0) return 0;
  cudaStream_t st = as_stream(stream);
  if (value == 0.0) {
    MWD_CHECK_CUDA(cudaMemsetAsync(p, 0, (size_t)n * sizeof(double), st));
    return 0;
  }
  int64_t grid = (n + 255) / 256;
  if (grid > 4 * sm_count()) grid = 4 * sm_count();
  fill_kernel<<<(unsigned)grid, 256, 0, st>>>(p, n, value);
  MWD_CHECK_LAUNCH();
  return 0;
}

extern "C" int mwd_sum_f64(const double* x, int64_t n, double* scratch256, double* out, void* stream) {
  return sum_doubles(x, n, scratch256, out, as_stream(stream));
}

extern "C" int mwd_rank_reduce(const double* gathered, int world, int64_t n, int log_domain, int64_t ll_index,
                               double* out, void* stream) {
  if (n <= 0) return 0;
  MWD_REQUIRE(world >= 1, "world %d < 1", world);
  rank_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(gathered, world, n, log_domain,
                                                                                 ll_index, out);
  MWD_CHECK_LAUNCH();
  return 0;
}

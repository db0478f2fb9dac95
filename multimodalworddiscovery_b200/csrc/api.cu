// Library-level entry points: error reporting, device geometry.
#include <stdarg.h>

#include "mwd_common.cuh"

namespace mwd {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)
      v = 148;
    cached = v;
    cached_dev = dev;
  }
  return cached;
}

int estep_grid_rows() { return sm_count() * kEstepCtasPerSm; }

}  // namespace mwd

extern "C" const char* mwd_last_error(void) { return mwd::g_err; }

extern "C" int mwd_version(void) { return 100; }

extern "C" int mwd_get_geometry(mwd_geometry* out) {
  int dev = 0;
  MWD_CHECK_CUDA(cudaGetDevice(&dev));
  out->sm_count = mwd::sm_count();
  out->estep_grid = mwd::estep_grid_rows();
  out->grad_splits = mwd::kGradSplits;
  return 0;
}

extern "C" int mwd_abi_sizeof(int which) {
  switch (which) {
    case 0: return (int)sizeof(mwd_geometry);
    case 1: return (int)sizeof(mwd_ik_problem);
    case 2: return (int)sizeof(mwd_partial_sizes);
    case 3: return (int)sizeof(mwd_ik_mstep_args);
    case 4: return (int)sizeof(mwd_hmm_problem);
    case 5: return (int)sizeof(mwd_hmm_mstep_args);
    default: return -1;
  }
}

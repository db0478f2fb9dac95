// Shared between the two E-step kernels (ik_estep.cu: CTA-per-4-pairs generic kernel,
// ik_estep_warp.cu: warp-per-pair kernel for small n).
#pragma once

#include "mwd_common.cuh"

namespace mwd {

struct EstepArgs {
  const int32_t* region_off;
  const int32_t* phone_off;
  const int32_t* phones;
  const double* pz;
  const double* init;    // init[n] row
  const double* trans;   // trans[n] table, [i*n+j]
  const double* obsT;
  double* pair_ll;
  double* cA_out;        // may be null
  int32_t* ca_out;       // may be null: argmax_k of the cA row of every phone (printAlignment :628)
  double* part_phone;    // [grid][P*K]
  double* part_init;     // [grid][(NMAX+1)*NMAX]
  double* part_trans;    // [grid][(NMAX+1)*NMAX*NMAX]
  double* scratch;
  double* stats;         // [slot_off[N]][4]: per (pair, t): s_t[n], floor-sum[n], xi-diag[n], r_t[n]
  const int64_t* slot_off;
  int64_t lo, hi;
  int64_t cta_scratch;   // doubles per CTA
  int n, K, P, B, NC, Tmax;
  int ll_only;           // 1: forward sweep + log-likelihood only
  double eps;            // floor of the likelihood / gamma / xi normalisers: MWD_EPS, or 0 (un-floored classes)
};

// argmax over a warp of NON-NEGATIVE doubles (or NaN) with np.argmax semantics: the largest value wins, the
// first index among equals, and a NaN beats every number.  For such values the IEEE bit pattern is
// monotone as an unsigned integer (NaN patterns sort above +inf), so the reduction is three REDUX
// instructions instead of five rounds of 64-bit shuffles.  Each lane passes its own best candidate
// (v, k); lanes without a candidate pass (0.0, INT_MAX).
__device__ __forceinline__ int warp_argmax_nonneg(double v, int k) {
  const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
  const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
  const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
  return __reduce_min_sync(0xffffffffu, (hi == mh && lo == ml) ? k : 0x7fffffff);
}
// strict "candidate beats incumbent" on the same ordering
__device__ __forceinline__ bool argmax_better(double v, double best) {
  return (unsigned long long)__double_as_longlong(v) > (unsigned long long)__double_as_longlong(best);
}

// Warp-per-pair kernel (ik_estep_warp.cu).  estep_warp_supported: true when an instantiation exists
// for (n, K); estep_warp_scratch: doubles of checkpoint scratch the launch needs for this bucket.
bool estep_warp_supported(int n, int K, int P);
int64_t estep_warp_scratch(int n, int K, int P, int Tmax, int64_t npairs);
int estep_warp_launch(EstepArgs a, cudaStream_t st);
// Scaled-float32 form of the warp-per-pair kernel (ik_estep_warp32.cu, MWD_MIXED_RECURSION); scratch == 0: no
// instantiation / shared memory for this shape, the float64 kernels run instead.
int64_t estep_warp32_scratch(int n, int K, int P, int Tmax, int64_t npairs);
int estep_warp32_launch(EstepArgs a, cudaStream_t st);

}  // namespace mwd

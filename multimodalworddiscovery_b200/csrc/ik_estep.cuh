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

// Warp-per-pair kernel (ik_estep_warp.cu).  estep_warp_supported: true when an instantiation exists
// for (n, K); estep_warp_scratch: doubles of checkpoint scratch the launch needs for this bucket.
bool estep_warp_supported(int n, int K, int P);
int64_t estep_warp_scratch(int n, int K, int P, int Tmax, int64_t npairs);
int estep_warp_launch(EstepArgs a, cudaStream_t st);

}  // namespace mwd

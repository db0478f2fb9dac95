// K4-tc -- posterior-parameter gradient GEMM on the Blackwell tensor cores (tcgen05 / TMEM / TMA).
//
//   updateSoftmaxWeight, hmm_dnn/image_phone_hmm_word_discoverer.py:475-488 (and the Gaussian class' mus update,
//   image_phone_gaussian_hmm_word_discoverer.py:488-499, which consumes the same product):
//       grad[k][d] = sum_r (conceptCounts - pz)[r][k] * [V,1][r][d]          (K x (D+1), reduction over ALL regions)
//
// Opt-in (MWD_MIXED_GRAD): the gradient has no EPS floor (SURVEY 8a census).  Arithmetic: split-TF32 on both sides,
//     v = v_hi + v_lo,  delta = d_hi + d_lo,   v*delta ~= v_hi d_hi + v_hi d_lo + v_lo d_hi     (error O(2^-22))
// accumulated in fp32 TMEM for `flush_every` row-blocks (512 rows), then added to a float64 partial table of the CTA
// (L2-resident); the CTAs' partials are summed in fixed order afterwards -> deterministic, no float atomics.
//
// The product is formed TRANSPOSED, grad^T[d][k], so that the big operand streams through TMA untouched:
//   A = V^T : M = 128 feature dims per UMMA, "MN-major" -- a [16 rows][32 dims] fp32 box loaded with the 32-byte-atom
//             128-byte swizzle (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) IS four canonical MN-major SWIZZLE_128B_BASE32B
//             atoms (4 rows x 128 B each), no transpose needed.  (The 16-byte-atom swizzle is rejected for MN-major
//             32-bit operands: the MMA silently produces zeros -- measured with tools/umma_layout_probe.py.)
//   B = delta : N = NPAD concepts, K-major without swizzle (core matrices of 8 concepts x 4 rows), built in shared
//             memory by the delta-builder warps from the float64 conceptCounts / pz rows (also the column sums =
//             the bias column of the gradient, accumulated in float64 per thread, fixed order)
//   D = up to 4 accumulators (D <= 512) x NPAD columns of TMEM
// Warp roles: 0 TMA producer | 1 UMMA issuer | 2-5 feature splitter (v -> v_hi, v_lo) | 6-9 TMEM flush |
//             10-13 delta builder.  One persistent CTA per SM owns a contiguous range of row-blocks.
#include <stdlib.h>

#include "mwd_common.cuh"
#include "tc_common.cuh"

namespace mwd {
namespace {

using namespace tc;

constexpr int GT_BR = 16;                 // rows per row-block = 2 UMMA k-steps of 8
constexpr int GT_THREADS = 448;
constexpr int GT_BOX_BYTES = GT_BR * 128; // one [16 rows][32 dims] fp32 box
constexpr int GT_MAX_STAGES = 8;
constexpr int GT_SMEM_BUDGET = 227 * 1024;

struct GradTcArgs {
  int64_t n_rows;
  int64_t n_rb;          // row-blocks in total
  int64_t rb_per_cta;
  int32_t D, K, NPAD;
  int32_t n_dblk;        // ceil(D / 32) boxes per row-block
  int32_t n_mtiles;      // ceil(D / 128) accumulators
  int32_t stages;
  int32_t split_mode;
  int32_t flush_every;   // row-blocks per fp32 accumulation group
  int32_t prefetch;      // row-blocks of features prefetched into L2 ahead of the TMA loads
  const double* cC;
  const double* pz;
  double* partials;      // [gridDim.x][K][D+1], pre-zeroed or carrying earlier chunks; always accumulated into
};

__host__ __device__ inline int gt_a_bytes(int n_dblk) { return n_dblk * GT_BOX_BYTES; }
__host__ __device__ inline int gt_b_bytes(int NPAD) { return 2 * NPAD * 32; }
__host__ __device__ inline int gt_stage_bytes(int n_dblk, int NPAD) {
  return 2 * gt_a_bytes(n_dblk) + 2 * gt_b_bytes(NPAD);
}

template <int NKK>   // ceil(NPAD / 32): concept columns handled per delta-builder lane
__global__ void __launch_bounds__(GT_THREADS, 1)
posterior_grad_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmPF,
                         const GradTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = a.stages, NPAD = a.NPAD, K = a.K, D = a.D;
  const int a_bytes = gt_a_bytes(a.n_dblk), b_bytes = gt_b_bytes(NPAD);
  const int stage_bytes = 2 * a_bytes + 2 * b_bytes;

  uint8_t* stage_base = smem;
  double* colsum = reinterpret_cast<double*>(smem + (size_t)S * stage_bytes);    // [4 builder warps][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(colsum + 4 * 128);
  uint64_t* full_raw = bars;
  uint64_t* xform = bars + GT_MAX_STAGES;
  uint64_t* empty = bars + 2 * GT_MAX_STAGES;
  uint64_t* acc_full = bars + 3 * GT_MAX_STAGES;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_raw[s], 1);
      mbar_init(&xform[s], 256);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 128);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmA);
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t rb_lo = (int64_t)blockIdx.x * a.rb_per_cta;
  const int64_t rb_hi = (rb_lo + a.rb_per_cta < a.n_rb) ? rb_lo + a.rb_per_cta : a.n_rb;
  const int64_t my_rb = rb_hi > rb_lo ? rb_hi - rb_lo : 0;
  double* part = a.partials + (size_t)blockIdx.x * K * (D + 1);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const uint64_t pol = policy_evict_first();
      // L2 prefetch a.prefetch row-blocks ahead (16 rows x 256 dims per instruction, un-swizzled map)
      int64_t pf = 0;
      auto prefetch_next = [&]() {
        if (pf < my_rb) {
          const int row = (int)((rb_lo + pf) * GT_BR);
          for (int c = 0; c < D; c += 256) tma_prefetch_l2_2d(&tmPF, c, row);
          ++pf;
        }
      };
      for (int i = 0; i < a.prefetch; ++i) prefetch_next();
      for (int64_t i = 0; i < my_rb; ++i) {
        const int s = (int)(i % S);
        const uint32_t ph = (uint32_t)(i / S) & 1u;
        prefetch_next();
        mbar_wait(&empty[s], ph ^ 1u);
        uint8_t* st = stage_base + (size_t)s * stage_bytes;
        mbar_arrive_expect_tx(&full_raw[s], (uint32_t)a_bytes);
        const int row = (int)((rb_lo + i) * GT_BR);
        for (int blk = 0; blk < a.n_dblk; ++blk)
          tma_load_2d_hint(st + blk * GT_BOX_BYTES, &tmA, &full_raw[s], blk * 32, row, pol);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ UMMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(128, NPAD, /*A MN-major*/ 1, /*B K-major*/ 0);
      uint32_t group = 0;
      for (int64_t i = 0; i < my_rb; ++i) {
        const int s = (int)(i % S);
        const uint32_t ph = (uint32_t)(i / S) & 1u;
        const bool first = (i % a.flush_every) == 0;
        if (first) {
          mbar_wait(acc_empty, (group & 1u) ^ 1u);     // previous group flushed
          tc_fence_after();
        }
        mbar_wait(&xform[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(stage_base + (size_t)s * stage_bytes);
        const uint32_t sb = sa + 2 * a_bytes;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint64_t b_hi = umma_desc(sb + ks * (b_bytes / 2), 128, 256, kSwizzleNone);
          const uint64_t b_lo = umma_desc(sb + b_bytes + ks * (b_bytes / 2), 128, 256, kSwizzleNone);
          for (int mt = 0; mt < a.n_mtiles; ++mt) {
            const uint32_t ao = sa + mt * 4 * GT_BOX_BYTES + ks * 1024;
            const uint64_t a_hi = umma_desc(ao, GT_BOX_BYTES, 512, kSwizzle128Base32);
            const uint64_t a_lo = umma_desc(ao + a_bytes, GT_BOX_BYTES, 512, kSwizzle128Base32);
            const uint32_t d = tmem_base + (uint32_t)(mt * NPAD);
            umma_tf32(d, a_hi, b_hi, idesc, (first && ks == 0) ? 0u : 1u);
            umma_tf32(d, a_hi, b_lo, idesc, 1u);
            umma_tf32(d, a_lo, b_hi, idesc, 1u);
          }
        }
        umma_commit(&empty[s]);
        if (((i + 1) % a.flush_every) == 0 || i + 1 == my_rb) {
          umma_commit(acc_full);
          ++group;
        }
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------------ feature splitter (128 threads)
    const int t = threadIdx.x - 64;
    const int n_chunks = a_bytes / 16;
    for (int64_t i = 0; i < my_rb; ++i) {
      const int s = (int)(i % S);
      const uint32_t ph = (uint32_t)(i / S) & 1u;
      mbar_wait(&full_raw[s], ph);
      float4* hi = reinterpret_cast<float4*>(stage_base + (size_t)s * stage_bytes);
      float4* lo = reinterpret_cast<float4*>(stage_base + (size_t)s * stage_bytes + a_bytes);
      for (int c0 = 0; c0 < n_chunks; c0 += 128 * 8) {
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = c0 + t + 128 * j;
          v[j] = (c < n_chunks) ? hi[c] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = c0 + t + 128 * j;
          if (c < n_chunks) {
            float4 h, l;
            if (a.split_mode == 0) {
              h.x = tf32_round(v[j].x); h.y = tf32_round(v[j].y); h.z = tf32_round(v[j].z); h.w = tf32_round(v[j].w);
              hi[c] = h;
            } else {
              h.x = __uint_as_float(__float_as_uint(v[j].x) & 0xffffe000u);
              h.y = __uint_as_float(__float_as_uint(v[j].y) & 0xffffe000u);
              h.z = __uint_as_float(__float_as_uint(v[j].z) & 0xffffe000u);
              h.w = __uint_as_float(__float_as_uint(v[j].w) & 0xffffe000u);
            }
            l.x = tf32_round(v[j].x - h.x); l.y = tf32_round(v[j].y - h.y);
            l.z = tf32_round(v[j].z - h.z); l.w = tf32_round(v[j].w - h.w);
            lo[c] = l;
          }
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&xform[s]);
    }
  } else if (warp < 10) {
    // ------------------------------------------------------------------ TMEM flush (warps 6..9)
    const int q = warp & 3;
    const int64_t n_groups = (my_rb + a.flush_every - 1) / a.flush_every;
    for (int64_t g = 0; g < n_groups; ++g) {
      mbar_wait(acc_full, (uint32_t)g & 1u);
      tc_fence_after();
      for (int mt = 0; mt < a.n_mtiles; ++mt) {
        const int d = mt * 128 + q * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * NPAD);
        for (int c0 = 0; c0 < NPAD; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(taddr + c0, r);
          tmem_ld_wait();
          if (d < D) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int k = c0 + j;
              // fire-and-forget float64 add: element (k, d) of this CTA's table is only ever touched by THIS thread,
              // in program order (same-address reductions of one thread stay ordered) -> deterministic, and the
              // flush does not wait for a load round trip to L2 while the tensor core idles
              if (k < K)
                asm volatile("red.global.add.f64 [%0], %1;" ::"l"(part + (size_t)k * (D + 1) + d),
                             "d"((double)__uint_as_float(r[j]))
                             : "memory");
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(acc_empty);
    }
  } else {
    // ------------------------------------------------------------------ delta builder (warps 10..13)
    const int w4 = warp - 10;                       // rows 4*w4 .. 4*w4+3 of the row-block
    double cs[NKK];
#pragma unroll
    for (int kk = 0; kk < NKK; ++kk) cs[kk] = 0.0;
    double rc[NKK][4], rp[NKK][4];
    auto load = [&](int64_t rb) {
      const int64_t r0 = rb * GT_BR + 4 * w4;
#pragma unroll
      for (int kk = 0; kk < NKK; ++kk) {
        const int k = lane + 32 * kk;
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const int64_t r = r0 + rr;
          const bool ok = (k < K) && (r < a.n_rows);
          rc[kk][rr] = ok ? __ldg(a.cC + r * K + k) : 0.0;
          rp[kk][rr] = ok ? __ldg(a.pz + r * K + k) : 0.0;
        }
      }
    };
    if (my_rb > 0) load(rb_lo);
    for (int64_t i = 0; i < my_rb; ++i) {
      const int s = (int)(i % S);
      const uint32_t ph = (uint32_t)(i / S) & 1u;
      float4 h[NKK], l[NKK];
#pragma unroll
      for (int kk = 0; kk < NKK; ++kk) {
        float hv[4], lv[4];
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const double dl = rc[kk][rr] - rp[kk][rr];
          cs[kk] += dl;
          hv[rr] = tf32_round((float)dl);
          lv[rr] = tf32_round((float)(dl - (double)hv[rr]));
        }
        h[kk] = make_float4(hv[0], hv[1], hv[2], hv[3]);
        l[kk] = make_float4(lv[0], lv[1], lv[2], lv[3]);
      }
      if (i + 1 < my_rb) load(rb_lo + i + 1);       // in flight while this row-block is written
      mbar_wait(&empty[s], ph ^ 1u);
      uint8_t* sb = stage_base + (size_t)s * stage_bytes + 2 * a_bytes + (w4 >> 1) * (b_bytes / 2) + (w4 & 1) * 128;
#pragma unroll
      for (int kk = 0; kk < NKK; ++kk) {
        const int k = lane + 32 * kk;
        if (k < NPAD) {
          const int off = (k >> 3) * 256 + (k & 7) * 16;
          *reinterpret_cast<float4*>(sb + off) = h[kk];
          *reinterpret_cast<float4*>(sb + b_bytes + off) = l[kk];
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&xform[s]);
    }
#pragma unroll
    for (int kk = 0; kk < NKK; ++kk) colsum[w4 * 128 + lane + 32 * kk] = cs[kk];
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
  // bias column of the gradient = column sums of delta: the 4 builder warps' sums in fixed order
  if (threadIdx.x < K && my_rb > 0) {
    const int k = threadIdx.x;
    const double sum = ((colsum[k] + colsum[128 + k]) + colsum[256 + k]) + colsum[384 + k];
    part[(size_t)k * (D + 1) + D] += sum;
  }
}

__global__ void grad_tc_reduce_kernel(const double* __restrict__ partial, int splits, int64_t elems,
                                      double* __restrict__ grad) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= elems) return;
  double s = 0.0;
  for (int sp = 0; sp < splits; ++sp) s += partial[(size_t)sp * elems + e];
  grad[e] = s;
}

}  // namespace
}  // namespace mwd

using namespace mwd;

extern "C" int mwd_posterior_grad_tc_supported(int feat_is_f64, int feat_dim, int n_concepts) {
  if (feat_is_f64 || feat_dim % 4 != 0 || feat_dim < 32 || feat_dim > 512 || n_concepts < 1 || n_concepts > MWD_KMAX)
    return 0;
  const int NPAD = (n_concepts + 15) & ~15;
  if (((feat_dim + 127) / 128) * NPAD > 512) return 0;
  return GT_SMEM_BUDGET - 8192 >= gt_stage_bytes((feat_dim + 31) / 32, NPAD) ? 1 : 0;
}

extern "C" int64_t mwd_posterior_grad_tc_partials_len(int n_concepts, int feat_dim) {
  return (int64_t)sm_count() * n_concepts * (feat_dim + 1);
}

// partials must hold mwd_posterior_grad_tc_partials_len doubles; accumulate == 0 zeroes them first (one memset),
// accumulate != 0 adds this shard chunk's product to what earlier chunks left there.
extern "C" int mwd_ik_posterior_grad_tc_partial(const mwd_ik_problem* p, double* partials, int accumulate,
                                                int split_mode, void* stream) {
  cudaStream_t st = as_stream(stream);
  const int K = p->n_concepts, D = p->feat_dim;
  MWD_REQUIRE(mwd_posterior_grad_tc_supported(p->feat_is_f64, D, K),
              "mwd_ik_posterior_grad_tc: unsupported shape D=%d K=%d (fp32 features, D %% 4 == 0, 32 <= D <= 512)", D, K);
  const int nsm = sm_count();
  if (!accumulate)
    MWD_CHECK_CUDA(cudaMemsetAsync(partials, 0, (size_t)nsm * K * (D + 1) * sizeof(double), st));
  if (p->n_regions <= 0) return 0;
  MWD_REQUIRE(((uintptr_t)p->feats & 15) == 0, "mwd_ik_posterior_grad_tc: feats must be 16-byte aligned");
  GradTcArgs a;
  a.n_rows = p->n_regions;
  a.n_rb = (p->n_regions + GT_BR - 1) / GT_BR;
  a.rb_per_cta = (a.n_rb + nsm - 1) / nsm;
  a.D = D;
  a.K = K;
  a.NPAD = (K + 15) & ~15;
  a.n_dblk = (D + 31) / 32;
  a.n_mtiles = (D + 127) / 128;
  a.split_mode = split_mode;
  a.flush_every = 32;                  // 512 rows (64 truncating fp32 accumulates) per group, then float64
  a.prefetch = 0;          // measured: no gain at any distance, kept as a knob (MWD_TC_PREFETCH)
  if (const char* e = getenv("MWD_TC_PREFETCH")) a.prefetch = atoi(e) < 0 ? 0 : atoi(e);
  a.cC = p->concept_counts;
  a.pz = p->pz;
  a.partials = partials;
  const int misc = 4 * 128 * 8 + (3 * GT_MAX_STAGES + 2) * 8 + 16;
  int stages = (GT_SMEM_BUDGET - 1024 - misc) / gt_stage_bytes(a.n_dblk, a.NPAD);
  if (stages > GT_MAX_STAGES) stages = GT_MAX_STAGES;
  MWD_REQUIRE(stages >= 1, "mwd_ik_posterior_grad_tc: no room for one pipeline stage");
  a.stages = stages;
  const size_t smem = 1024 + (size_t)stages * gt_stage_bytes(a.n_dblk, a.NPAD) + misc;
  CUtensorMap tmA;
  if (int rc = tc::make_tmap_f32_2d(&tmA, p->feats, (uint64_t)p->n_regions, (uint64_t)D, (uint64_t)D * 4, GT_BR, /*atom32=*/true))
    return rc;
  CUtensorMap tmPF;
  if (int rc = tc::make_tmap_f32_2d_plain(&tmPF, p->feats, (uint64_t)p->n_regions, (uint64_t)D, (uint64_t)D * 4, GT_BR,
                                          (uint32_t)(D < 256 ? D : 256)))
    return rc;
  const int64_t grid = (a.n_rb + a.rb_per_cta - 1) / a.rb_per_cta;
  const int nkk = (a.NPAD + 31) / 32;
#define MWD_GT_LAUNCH(N)                                                                                        \
  case N: {                                                                                                     \
    static bool attr_set = false;                                                                               \
    if (!attr_set) {                                                                                            \
      MWD_CHECK_CUDA(cudaFuncSetAttribute(posterior_grad_tc_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          GT_SMEM_BUDGET));                                                     \
      attr_set = true;                                                                                          \
    }                                                                                                           \
    posterior_grad_tc_kernel<N><<<(unsigned)grid, GT_THREADS, smem, st>>>(tmA, tmPF, a);                             \
  } break;
  switch (nkk) {
    MWD_GT_LAUNCH(1) MWD_GT_LAUNCH(2) MWD_GT_LAUNCH(3) MWD_GT_LAUNCH(4)
    default: MWD_REQUIRE(false, "mwd_ik_posterior_grad_tc: bad concept padding %d", a.NPAD);
  }
#undef MWD_GT_LAUNCH
  MWD_CHECK_LAUNCH();
  return 0;
}

extern "C" int mwd_posterior_grad_tc_finish(int n_concepts, int feat_dim, const double* partials, double* grad,
                                            void* stream) {
  const int64_t elems = (int64_t)n_concepts * (feat_dim + 1);
  grad_tc_reduce_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, as_stream(stream)>>>(partials, sm_count(), elems,
                                                                                        grad);
  MWD_CHECK_LAUNCH();
  return 0;
}

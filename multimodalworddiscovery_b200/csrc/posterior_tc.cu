// K3-tc -- image-posterior emission GEMM + row softmax on the Blackwell tensor cores (tcgen05 / TMEM / TMA).
//
//   softmaxLayer, hmm_dnn/image_phone_hmm_word_discoverer.py:533-541:  pz = softmax_rows([V,1] W^T).
//
// Opt-in (mwd_ik_problem.mixed_precision & MWD_MIXED_POSTERIOR): softmaxLayer has no EPS floor (SURVEY 8a census),
// so it may leave the FP64 pipe as long as the north-star tolerance (1e-5) holds.  Arithmetic: split-TF32,
//     v = v_hi + v_lo (fp32 feature, both parts exact TF32 values),  w = w_hi + w_lo (float64 weight),
//     logit = [v_hi . w_hi]  +  [v_hi . w_lo + v_lo . w_hi]  + bias
// three tcgen05.mma kind::tf32 per 8 feature dims, the large and the small sum in SEPARATE fp32 TMEM accumulators,
// recombined with the bias in float64, then the max-shifted softmax in float64.  Dropped terms are O(2^-22) per
// product; measured error against the float64 kernel is in tests/test_gpu_posterior_tc.py.
//
// Structure (one persistent CTA per SM, 128-row tiles, 32-dim k-blocks, warp-specialised):
//   warp 0      TMA producer: cp.async.bulk.tensor of the fp32 feature tile (128 x 32, 128-byte swizzle) and of the
//               pre-split weight tiles (w_hi | w_lo, NPAD x 32 each) into an S-stage ring
//   warps 2-5   splitter: v -> v_hi (in place), v_lo (second buffer); element-wise, so the swizzle is irrelevant
//   warp 1      one thread issues the UMMAs: [main | corr] = v_hi . [w_hi ; w_lo]^T (N = 2 NPAD), corr += v_lo . w_hi^T;
//               tcgen05.commit releases the stage and, after the last k-block, publishes the accumulator
//   warps 6-9   epilogue: tcgen05.ld (thread = row), float64 recombination + softmax through a per-warp staging
//               buffer, written back with ONE bulk copy per warp (32 rows x K doubles are contiguous in pz)
//   TMEM        C = 512 / (2 NPAD) (at most 4) accumulator CHUNKS of 2 NPAD columns: the feature dimension is cut into
//               C ranges with an accumulator each, summed in float64 by the epilogue.  The tensor core truncates
//               (rounds toward zero) every fp32 accumulate -- measured: mean logit error = 0.5 ulp x #accumulates --
//               so the error shrinks ~1/C.  The accumulators are released as soon as the epilogue has copied them
//               to shared memory (first pass), which is what keeps a single buffer cheap.
#include <stdlib.h>

#include "mwd_common.cuh"
#include "tc_common.cuh"

namespace mwd {
namespace tc {

encode_tiled_fn get_encode_tiled() {
  static encode_tiled_fn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<encode_tiled_fn>(p);
  return fn;
}

int make_tmap_f32_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                     uint32_t box_rows, bool atom32) {
  encode_tiled_fn enc = get_encode_tiled();
  MWD_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {32u, box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MWD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for a %llu x %llu fp32 tensor", (int)r,
              (unsigned long long)rows, (unsigned long long)cols);
  return 0;
}

int make_tmap_f32_2d_plain(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                           uint32_t box_rows, uint32_t box_cols) {
  encode_tiled_fn enc = get_encode_tiled();
  MWD_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MWD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (prefetch map) failed (%d)", (int)r);
  return 0;
}

}  // namespace tc

namespace {

using namespace tc;

constexpr int PT_BM = 128;                       // rows per tile = UMMA M
constexpr int PT_BK = 32;                        // feature dims per k-block (one 128-byte swizzle row of fp32)
constexpr int PT_THREADS = 320;                  // 10 warps, see the role table above
constexpr int PT_A_BYTES = PT_BM * PT_BK * 4;    // 16 KB
constexpr int PT_MAX_STAGES = 8;
constexpr int PT_SMEM_BUDGET = 227 * 1024;

struct PostTcArgs {
  int64_t n_rows;
  int64_t n_tiles;
  int32_t n_kblocks;
  int32_t K;             // concepts
  int32_t NPAD;          // K rounded up to 16
  int32_t stages;
  int32_t split_mode;    // 0: v_hi = rn_tf32(v) written in place; 1: v_hi = the raw word (hardware truncation)
  int32_t chunks;        // accumulator chunks over the feature dimension
  int32_t bulk_ok;       // pz is 16-byte aligned: whole-warp bulk stores allowed
  int32_t prefetch;      // k-blocks of features prefetched into L2 ahead of the TMA loads
  int32_t ldw;           // D + 1 (row stride of W, bias in the last column)
  const double* W;
  double* pz;
};

__host__ __device__ inline int pt_stage_bytes(int NPAD) { return 2 * PT_A_BYTES + 2 * NPAD * PT_BK * 4; }
__host__ __device__ inline int pt_stage_row_doubles(int K) { return K | 1; }   // odd stride: conflict-free rows

// exp(x) for x <= 0 on the SFU: 2^(x log2 e) with the integer part moved into the float64 exponent and the fraction
// (|f| <= 1/2, float32-exact to 3e-8) through ex2.approx (relative error 2^-22).  The row softmax needs 65 of these per
// region; libm's float64 exp (~28 FP64-pipe instructions each) made the epilogue, not the tensor core, the critical path.
__device__ __forceinline__ double exp_nonpos(double x) {
  const double t = fmax(x, -708.0) * 1.4426950408889634;
  const double r = rint(t);
  float p;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"((float)(t - r)));
  const double v = __longlong_as_double(__double_as_longlong((double)p) + ((long long)(int)r << 52));
  return x < -708.0 ? 0.0 : v;
}

__global__ void __launch_bounds__(PT_THREADS, 1)
posterior_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                    const PostTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = a.stages, NPAD = a.NPAD, K = a.K;
  const int stage_bytes = pt_stage_bytes(NPAD);
  const int w_bytes = NPAD * PT_BK * 4;
  const int srow = pt_stage_row_doubles(K);

  uint8_t* stage_base = smem;
  double* staging = reinterpret_cast<double*>(smem + (size_t)S * stage_bytes);       // 4 warps x 32 rows x srow
  double* bias = staging + 4 * 32 * srow;                                           // K
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias + ((K + 1) & ~1));
  uint64_t* full_raw = bars;                      // [S] TMA landed
  uint64_t* xform = bars + PT_MAX_STAGES;         // [S] split written
  uint64_t* empty = bars + 2 * PT_MAX_STAGES;     // [S] UMMAs of the stage retired
  uint64_t* tmem_full = bars + 3 * PT_MAX_STAGES; // accumulators of a tile complete
  uint64_t* tmem_empty = tmem_full + 1;           // accumulators copied out by the epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_raw[s], 1);
      mbar_init(&xform[s], 128);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 128);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  for (int k = threadIdx.x; k < K; k += PT_THREADS) bias[k] = a.W[(size_t)k * a.ldw + (a.ldw - 1)];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const uint64_t pol_stream = policy_evict_first(), pol_keep = policy_evict_last();
      uint32_t it = 0;
      // L2 prefetch runs a.prefetch k-blocks ahead of the loads: the shared-memory ring (3 stages) alone keeps too few
      // bytes in flight to cover HBM latency (measured 41 % of the HBM peak without it)
      int64_t pf_tile = blockIdx.x;
      int pf_kb = 0;
      auto prefetch_next = [&]() {
        if (pf_tile < a.n_tiles) {
          tma_prefetch_l2_2d(&tmA, pf_kb * PT_BK, (int)(pf_tile * PT_BM));
          if (++pf_kb == a.n_kblocks) { pf_kb = 0; pf_tile += gridDim.x; }
        }
      };
      for (int i = 0; i < a.prefetch; ++i) prefetch_next();
      for (int64_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < a.n_kblocks; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1u;
          prefetch_next();
          mbar_wait(&empty[s], ph ^ 1u);
          uint8_t* st = stage_base + (size_t)s * stage_bytes;
          mbar_arrive_expect_tx(&full_raw[s], PT_A_BYTES + 2 * w_bytes);
          tma_load_2d_hint(st, &tmA, &full_raw[s], kb * PT_BK, (int)(tile * PT_BM), pol_stream);
          tma_load_2d_hint(st + 2 * PT_A_BYTES, &tmW, &full_raw[s], kb * PT_BK, 0, pol_keep);
          tma_load_2d_hint(st + 2 * PT_A_BYTES + w_bytes, &tmW, &full_raw[s], kb * PT_BK, NPAD, pol_keep);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ UMMA issuer
    if (lane == 0) {
      const uint32_t idesc_wide = umma_idesc_tf32(PT_BM, 2 * NPAD, 0, 0);
      const uint32_t idesc_narrow = umma_idesc_tf32(PT_BM, NPAD, 0, 0);
      uint32_t it = 0, tcount = 0;
      const int C = a.chunks;
      for (int64_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++tcount) {
        mbar_wait(tmem_empty, (tcount & 1u) ^ 1u);
        tc_fence_after();
        int prev_chunk = -1;
        for (int kb = 0; kb < a.n_kblocks; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1u;
          const int chunk = kb * C / a.n_kblocks;
          const uint32_t d_main = tmem_base + (uint32_t)(chunk * 2 * NPAD);
          const uint32_t d_corr = d_main + (uint32_t)NPAD;
          mbar_wait(&xform[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + (size_t)s * stage_bytes);
          const uint64_t d_hi = umma_desc(sa, 16, 1024, kSwizzle128);
          const uint64_t d_lo = umma_desc(sa + PT_A_BYTES, 16, 1024, kSwizzle128);
          const uint64_t d_w = umma_desc(sa + 2 * PT_A_BYTES, 16, 1024, kSwizzle128);
#pragma unroll
          for (int ks = 0; ks < PT_BK / 8; ++ks) {
            const uint64_t adv = (uint64_t)(ks * 32 >> 4);        // 8 tf32 = 32 bytes along K inside the swizzle row
            const uint32_t acc = (chunk != prev_chunk && ks == 0) ? 0u : 1u;
            umma_tf32(d_main, d_hi + adv, d_w + adv, idesc_wide, acc);      // [v_hi.w_hi | v_hi.w_lo]
            umma_tf32(d_corr, d_lo + adv, d_w + adv, idesc_narrow, 1u);     // corr += v_lo.w_hi
          }
          prev_chunk = chunk;
          umma_commit(&empty[s]);
        }
        umma_commit(tmem_full);
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------------------------ splitter (128 threads)
    const int t = threadIdx.x - 64;
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < a.n_kblocks; ++kb, ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1u;
        mbar_wait(&full_raw[s], ph);
        float4* hi = reinterpret_cast<float4*>(stage_base + (size_t)s * stage_bytes);
        float4* lo = reinterpret_cast<float4*>(stage_base + (size_t)s * stage_bytes + PT_A_BYTES);
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = hi[t + 128 * j];
        if (a.split_mode == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 h, l;
            h.x = tf32_round(v[j].x); h.y = tf32_round(v[j].y); h.z = tf32_round(v[j].z); h.w = tf32_round(v[j].w);
            l.x = tf32_round(v[j].x - h.x); l.y = tf32_round(v[j].y - h.y);
            l.z = tf32_round(v[j].z - h.z); l.w = tf32_round(v[j].w - h.w);
            hi[t + 128 * j] = h;
            lo[t + 128 * j] = l;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 l;
            l.x = v[j].x - __uint_as_float(__float_as_uint(v[j].x) & 0xffffe000u);
            l.y = v[j].y - __uint_as_float(__float_as_uint(v[j].y) & 0xffffe000u);
            l.z = v[j].z - __uint_as_float(__float_as_uint(v[j].z) & 0xffffe000u);
            l.w = v[j].w - __uint_as_float(__float_as_uint(v[j].w) & 0xffffe000u);
            lo[t + 128 * j] = l;
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(&xform[s]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 6..9)
    const int q = warp & 3;                                 // TMEM lane quadrant this warp may read
    double* my = staging + (size_t)(q * 32 + lane) * srow;  // this thread's row
    double* wbase = staging + (size_t)q * 32 * srow;
    uint32_t tcount = 0;
    bool store_pending = false;
    for (int64_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++tcount) {
      // staging of this warp must have been read out by the previous bulk store
      if (store_pending) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        store_pending = false;
      }
      mbar_wait(tmem_full, tcount & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
      double m = -1.0 / 0.0;
      for (int c0 = 0; c0 < NPAD; c0 += 16) {
        double x[16];
        float xc[16];                       // the small sums (2^-11 of the large ones) add up in float32
#pragma unroll
        for (int j = 0; j < 16; ++j) { x[j] = 0.0; xc[j] = 0.0f; }
        for (int ch = 0; ch < a.chunks; ++ch) {
          uint32_t rm[16], rc[16];
          tmem_ld16(taddr + ch * 2 * NPAD + c0, rm);
          tmem_ld16(taddr + ch * 2 * NPAD + NPAD + c0, rc);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            x[j] += (double)__uint_as_float(rm[j]);
            xc[j] += __uint_as_float(rc[j]);
          }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] += (double)xc[j];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int k = c0 + j;
          if (k < K) {
            const double xv = x[j] + bias[k];
            my[k] = xv;
            m = xv > m ? xv : m;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tmem_empty);              // the accumulators may be overwritten by the next tile
      double ssum = 0.0;
      for (int k = 0; k < K; ++k) {
        const double e = exp_nonpos(my[k] - m);
        my[k] = e;
        ssum += e;
      }
      const double inv = 1.0 / ssum;
      const int64_t row0 = tile * PT_BM + q * 32;
      const int64_t rows_left = a.n_rows - row0;
      if (rows_left >= 32 && srow == K && a.bulk_ok) {
        for (int k = 0; k < K; ++k) my[k] *= inv;
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          double* g = a.pz + row0 * K;
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(smem_u32(wbase)),
                       "r"((uint32_t)(32 * K * 8))
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        store_pending = true;
      } else if (rows_left > 0) {
        // even K (padded staging rows) or the ragged last tile: plain stores
        if (lane < rows_left) {
          double* g = a.pz + (row0 + lane) * K;
          for (int k = 0; k < K; ++k) g[k] = my[k] * inv;
        }
        __syncwarp();
      }
    }
    if (store_pending && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// W (K x (D+1) float64, bias last) -> [w_hi ; w_lo] as fp32 holding exact TF32 values, (2 NPAD) x D, zero rows >= K
__global__ void split_weights_kernel(const double* __restrict__ W, int K, int D, int NPAD, float* __restrict__ out) {
  const int64_t n = (int64_t)NPAD * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i / D), d = (int)(i % D);
    float hi = 0.f, lo = 0.f;
    if (k < K) {
      const double w = W[(size_t)k * (D + 1) + d];
      hi = tc::tf32_round((float)w);
      lo = tc::tf32_round((float)(w - (double)hi));
    }
    out[i] = hi;
    out[n + i] = lo;
  }
}

}  // namespace
}  // namespace mwd

using namespace mwd;

extern "C" int64_t mwd_posterior_tc_scratch_bytes(int n_concepts, int feat_dim) {
  const int NPAD = (n_concepts + 15) & ~15;
  return (int64_t)2 * NPAD * feat_dim * sizeof(float);
}

extern "C" int mwd_posterior_tc_supported(int feat_is_f64, int feat_dim, int n_concepts) {
  if (feat_is_f64 || feat_dim % 4 != 0 || feat_dim < 32 || n_concepts < 1 || n_concepts > MWD_KMAX) return 0;
  const int NPAD = (n_concepts + 15) & ~15;
  const int staging = 4 * 32 * pt_stage_row_doubles(n_concepts) * 8;
  return PT_SMEM_BUDGET - staging - 4096 >= pt_stage_bytes(NPAD) ? 1 : 0;
}

extern "C" int mwd_posterior_linear_tc(const float* feats, int64_t n_regions, int feat_dim, const double* W,
                                       int n_concepts, double* pz, void* w_split_scratch, int split_mode,
                                       void* stream) {
  if (n_regions <= 0) return 0;
  MWD_REQUIRE(mwd_posterior_tc_supported(0, feat_dim, n_concepts),
              "mwd_posterior_linear_tc: unsupported shape D=%d K=%d (needs fp32 features, D %% 4 == 0, K <= %d)",
              feat_dim, n_concepts, MWD_KMAX);
  MWD_REQUIRE(((uintptr_t)feats & 15) == 0 && ((uintptr_t)w_split_scratch & 15) == 0,
              "mwd_posterior_linear_tc: feats / scratch must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  const int K = n_concepts, D = feat_dim, NPAD = (K + 15) & ~15;
  float* wsplit = static_cast<float*>(w_split_scratch);
  split_weights_kernel<<<64, 256, 0, st>>>(W, K, D, NPAD, wsplit);
  MWD_CHECK_LAUNCH();

  CUtensorMap tmA, tmW;
  if (int rc = tc::make_tmap_f32_2d(&tmA, feats, (uint64_t)n_regions, (uint64_t)D, (uint64_t)D * 4, PT_BM)) return rc;
  if (int rc = tc::make_tmap_f32_2d(&tmW, wsplit, (uint64_t)2 * NPAD, (uint64_t)D, (uint64_t)D * 4, (uint32_t)NPAD))
    return rc;

  PostTcArgs a;
  a.n_rows = n_regions;
  a.n_tiles = (n_regions + PT_BM - 1) / PT_BM;
  a.n_kblocks = (D + PT_BK - 1) / PT_BK;
  a.K = K;
  a.NPAD = NPAD;
  a.split_mode = split_mode;
  a.chunks = 512 / (2 * NPAD) > 4 ? 4 : 512 / (2 * NPAD);
  if (a.chunks > a.n_kblocks) a.chunks = a.n_kblocks;
  a.bulk_ok = (((uintptr_t)pz & 15) == 0) ? 1 : 0;
  a.prefetch = 0;          // measured: no gain at any distance (the kernel is not HBM-latency bound), kept as a knob
  if (const char* e = getenv("MWD_TC_PREFETCH")) a.prefetch = atoi(e) < 0 ? 0 : atoi(e);
  a.ldw = D + 1;
  a.W = W;
  a.pz = pz;
  const int staging = 4 * 32 * pt_stage_row_doubles(K) * 8;
  const int misc = ((K + 1) & ~1) * 8 + (3 * PT_MAX_STAGES + 2) * 8 + 16;
  int stages = (PT_SMEM_BUDGET - 1024 - staging - misc) / pt_stage_bytes(NPAD);
  if (stages > PT_MAX_STAGES) stages = PT_MAX_STAGES;
  MWD_REQUIRE(stages >= 1, "mwd_posterior_linear_tc: no room for one pipeline stage at K=%d", K);
  a.stages = stages;
  const size_t smem = 1024 + (size_t)stages * pt_stage_bytes(NPAD) + staging + misc;
  static bool attr_set = false;
  if (!attr_set) {
    MWD_CHECK_CUDA(cudaFuncSetAttribute(posterior_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        PT_SMEM_BUDGET));
    attr_set = true;
  }
  const int64_t grid = a.n_tiles < sm_count() ? a.n_tiles : sm_count();
  posterior_tc_kernel<<<(unsigned)grid, PT_THREADS, smem, st>>>(tmA, tmW, a);
  MWD_CHECK_LAUNCH();
  return 0;
}

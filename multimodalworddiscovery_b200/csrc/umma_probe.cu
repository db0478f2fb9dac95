// Diagnostic: ONE tcgen05.mma kind::tf32 on caller-supplied shared-memory images and descriptors, accumulator dumped.
// Used by tools/umma_layout_probe.py to pin the shared-memory layouts / descriptor fields the tensor-core kernels
// rely on (K-major SWIZZLE_128B, MN-major SWIZZLE_128B, K-major without swizzle) against NumPy on the GPU box.
#include "mwd_common.cuh"
#include "tc_common.cuh"

namespace mwd {
namespace {
using namespace tc;

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const uint32_t* __restrict__ a_img, int a_words, const uint32_t* __restrict__ b_img, int b_words,
                  uint64_t adesc, uint64_t bdesc, uint32_t idesc, int n_cols, int n_mma, uint32_t a_step, uint32_t b_step,
                  float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint32_t* sa = reinterpret_cast<uint32_t*>(smem);
  uint32_t* sb = reinterpret_cast<uint32_t*>(smem + 96 * 1024);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < a_words; i += blockDim.x) sa[i] = a_img[i];
  for (int i = threadIdx.x; i < b_words; i += blockDim.x) sb[i] = b_img[i];
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) tmem_alloc(&slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = slot;
  if (threadIdx.x == 0) {
    const uint64_t ad = adesc + (uint64_t)((smem_u32(sa) >> 4) & 0x3fffu);
    const uint64_t bd = bdesc + (uint64_t)((smem_u32(sb) >> 4) & 0x3fffu);
    for (int i = 0; i < n_mma; ++i) umma_tf32(tbase, ad + (uint64_t)i * a_step, bd + (uint64_t)i * b_step, idesc, i ? 1u : 0u);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c0 = 0; c0 < n_cols; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j)
      if (c0 + j < n_cols) out[(size_t)(warp * 32 + lane) * n_cols + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tbase, 256);
}
}  // namespace
}  // namespace mwd

using namespace mwd;

// a_img / b_img [dev]: raw 32-bit words copied to shared memory (A at a 1024-byte aligned base, B 96 KB above it);
// adesc / bdesc: descriptors with a ZERO start-address field (the kernel adds the base); n_mma MMAs are issued, the
// descriptors advanced by a_step / b_step (16-byte units) each time; out [dev]: 128 x n_cols fp32.
extern "C" int mwd_umma_probe(const void* a_img, int a_words, const void* b_img, int b_words, uint64_t adesc,
                              uint64_t bdesc, uint32_t idesc, int n_cols, int n_mma, uint32_t a_step, uint32_t b_step,
                              float* out, void* stream) {
  MWD_REQUIRE(a_words * 4 <= 96 * 1024 && b_words * 4 <= 96 * 1024 && n_cols <= 256, "umma probe: image too large");
  static bool attr_set = false;
  if (!attr_set) {
    MWD_CHECK_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  umma_probe_kernel<<<1, 128, 193 * 1024 + 1024, as_stream(stream)>>>(
      static_cast<const uint32_t*>(a_img), a_words, static_cast<const uint32_t*>(b_img), b_words, adesc, bdesc, idesc,
      n_cols, n_mma, a_step, b_step, out);
  MWD_CHECK_LAUNCH();
  return 0;
}

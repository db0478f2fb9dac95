// Issue-rate probe for the FMA forms that matter to the concept-chain kernel (K2) on sm_100a:
//   ffma3   : FFMA with three register operands
//   ffma_c  : FFMA whose multiplier comes from __constant__ memory (compiler: UR / c[] operand)
//   ffma2   : packed fma.rn.f32x2 (two FP32 FMAs per lane per instruction)
//   dfma    : DFMA (the float64 kernel)
// Each thread runs NCH independent accumulator chains; 1024 threads x 2 CTAs per SM, all SMs.
// Build:  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fma_forms fma_forms.cu
#include <cstdio>
#include <cuda_runtime.h>

__constant__ float c_f[32];
__constant__ double c_d[32];
constexpr int NCH = 10;
constexpr int ITERS = 4096;

__global__ void k_ffma3(float* out, float a0, float b0) {
  float acc[NCH], b[NCH];
  for (int i = 0; i < NCH; ++i) { acc[i] = threadIdx.x * 1e-3f + i; b[i] = b0 + i * 1e-3f; }
  float a = a0;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) acc[i] = fmaf(acc[i], a, b[i]);
  }
  float s = 0; for (int i = 0; i < NCH; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma_c(float* out) {
  float acc[NCH];
  for (int i = 0; i < NCH; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) acc[i] = fmaf(acc[i], c_f[i], acc[(i + 1) % NCH]);
  }
  float s = 0; for (int i = 0; i < NCH; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, float a0, float b0) {
  unsigned long long acc[NCH], b[NCH], a;
  for (int i = 0; i < NCH; ++i) {
    float2 v = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i), w = make_float2(b0 + i * 1e-3f, b0 - i * 1e-3f);
    acc[i] = *reinterpret_cast<unsigned long long*>(&v);
    b[i] = *reinterpret_cast<unsigned long long*>(&w);
  }
  { float2 v = make_float2(a0, a0 * 0.999f); a = *reinterpret_cast<unsigned long long*>(&v); }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(a), "l"(b[i]));
  }
  float s = 0;
  for (int i = 0; i < NCH; ++i) { float2 v = *reinterpret_cast<float2*>(&acc[i]); s += v.x + v.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__constant__ float2 c_f2[32];
// packed FMA whose multiplier pair comes from __constant__ memory (compiler: FFMA2 R, R, UR, R)
__global__ void k_ffma2_c(float* out) {
  unsigned long long acc[NCH];
  for (int i = 0; i < NCH; ++i) {
    float2 v = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
    acc[i] = *reinterpret_cast<unsigned long long*>(&v);
  }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      float2 m = c_f2[i];
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(*reinterpret_cast<unsigned long long*>(&m)), "l"(acc[(i + 1) % NCH]));
    }
  }
  float s = 0;
  for (int i = 0; i < NCH; ++i) { float2 v = *reinterpret_cast<float2*>(&acc[i]); s += v.x + v.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// packed FMA interleaved 1:1 with an integer ALU instruction: does FFMA2 leave issue slots free?
__global__ void k_ffma2_mix(float* out, float a0, float b0, int q) {
  unsigned long long acc[NCH], b[NCH], a;
  int z[NCH];
  for (int i = 0; i < NCH; ++i) {
    float2 v = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i), w = make_float2(b0 + i * 1e-3f, b0 - i * 1e-3f);
    acc[i] = *reinterpret_cast<unsigned long long*>(&v);
    b[i] = *reinterpret_cast<unsigned long long*>(&w);
    z[i] = threadIdx.x + i;
  }
  { float2 v = make_float2(a0, a0 * 0.999f); a = *reinterpret_cast<unsigned long long*>(&v); }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(a), "l"(b[i]));
      asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(z[i]) : "r"(q), "r"(it));
    }
  }
  float s = 0;
  for (int i = 0; i < NCH; ++i) { float2 v = *reinterpret_cast<float2*>(&acc[i]); s += v.x + v.y + z[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dfma(double* out, double a0, double b0) {
  double acc[NCH], b[NCH];
  for (int i = 0; i < NCH; ++i) { acc[i] = threadIdx.x * 1e-3 + i; b[i] = b0 + i * 1e-3; }
  double a = a0;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) acc[i] = fma(acc[i], a, b[i]);
  }
  double s = 0; for (int i = 0; i < NCH; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dfma_c(double* out) {
  double acc[NCH];
  for (int i = 0; i < NCH; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) acc[i] = fma(acc[i], c_d[i], acc[(i + 1) % NCH]);
  }
  double s = 0; for (int i = 0; i < NCH; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F> static float time_ms(F launch) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); cudaDeviceSynchronize();
  cudaEventRecord(e0); for (int r = 0; r < 5; ++r) launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 5;
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float hf[32]; double hd[32];
  for (int i = 0; i < 32; ++i) { hf[i] = 0.999f + 1e-5f * i; hd[i] = 0.999 + 1e-5 * i; }
  cudaMemcpyToSymbol(c_f, hf, sizeof(hf)); cudaMemcpyToSymbol(c_d, hd, sizeof(hd));
  const int grid = sms * 2, block = 1024;
  void* buf; cudaMalloc(&buf, (size_t)grid * block * 8);
  const double fmas = (double)grid * block * NCH * ITERS;
  float t;
  t = time_ms([&] { k_ffma3<<<grid, block>>>((float*)buf, 0.9991f, 1e-3f); });
  printf("ffma3   %8.3f ms  %7.2f TFLOP/s  (%.1f FMA/clk/SM at 1965 MHz)\n", t, 2 * fmas / t / 1e9, fmas / (t * 1e-3) / sms / 1.965e9);
  t = time_ms([&] { k_ffma_c<<<grid, block>>>((float*)buf); });
  printf("ffma_c  %8.3f ms  %7.2f TFLOP/s  (%.1f FMA/clk/SM)\n", t, 2 * fmas / t / 1e9, fmas / (t * 1e-3) / sms / 1.965e9);
  t = time_ms([&] { k_ffma2<<<grid, block>>>((float*)buf, 0.9991f, 1e-3f); });
  printf("ffma2   %8.3f ms  %7.2f TFLOP/s  (%.1f FMA/clk/SM)\n", t, 4 * fmas / t / 1e9, 2 * fmas / (t * 1e-3) / sms / 1.965e9);
  { float2 h2[32]; for (int i = 0; i < 32; ++i) h2[i] = make_float2(hf[i], hf[i]); cudaMemcpyToSymbol(c_f2, h2, sizeof(h2)); }
  t = time_ms([&] { k_ffma2_c<<<grid, block>>>((float*)buf); });
  printf("ffma2_c %8.3f ms  %7.2f TFLOP/s  (%.1f FMA/clk/SM)\n", t, 4 * fmas / t / 1e9, 2 * fmas / (t * 1e-3) / sms / 1.965e9);
  t = time_ms([&] { k_ffma2_mix<<<grid, block>>>((float*)buf, 0.9991f, 1e-3f, 12345); });
  printf("ffma2+lop3 1:1 %8.3f ms  %7.2f TFLOP/s  (%.1f FMA/clk/SM; the same number of FFMA2 as ffma2)\n", t, 4 * fmas / t / 1e9, 2 * fmas / (t * 1e-3) / sms / 1.965e9);
  t = time_ms([&] { k_dfma<<<grid, block>>>((double*)buf, 0.9991, 1e-3); });
  printf("dfma    %8.3f ms  %7.2f TFLOP/s  (%.1f FMA/clk/SM)\n", t, 2 * fmas / t / 1e9, fmas / (t * 1e-3) / sms / 1.965e9);
  t = time_ms([&] { k_dfma_c<<<grid, block>>>((double*)buf); });
  printf("dfma_c  %8.3f ms  %7.2f TFLOP/s  (%.1f FMA/clk/SM)\n", t, 2 * fmas / t / 1e9, fmas / (t * 1e-3) / sms / 1.965e9);
  return 0;
}

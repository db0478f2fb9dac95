// Measures the FP64 issue ceilings of one B200: DFMA (SIMT) and DMMA m8n8k4 / m16n8k16 (tensor).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters) {
  double a[8], x = 1.0000001 + threadIdx.x * 1e-9, y = 0.9999999;
  for (int i = 0; i < 8; ++i) a[i] = i + threadIdx.x;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], x, y);
  double s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma_kernel(double* out, int iters) {
  double c[8][2], a = 1.0 + threadIdx.x * 1e-9, b = 1e-3;
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  double s = 0;
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma16_kernel(double* out, int iters) {
  double c[4][4], a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = 1.0 + (threadIdx.x + i) * 1e-9;
  for (int i = 0; i < 4; ++i) b[i] = 1e-3 * (i + 1);
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) c[i][j] = i + j;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                     "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
  double s = 0;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float timeit(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 16 * 1024);
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    int threads = warps * 32 > 1024 ? 1024 : warps * 32, blocks = sms * (warps * 32 / threads);
    float ms = timeit([&] { dfma_kernel<<<blocks, threads>>>(out, iters); });
    double fl = 2.0 * 8 * iters * (double)blocks * threads;
    printf("DFMA        warps/SM %2d : %7.2f TFLOP/s\n", warps, fl / ms / 1e9);
    ms = timeit([&] { dmma_kernel<<<blocks, threads>>>(out, iters); });
    fl = 2.0 * 256 * 8 * iters * (double)blocks * (threads / 32);
    printf("DMMA m8n8k4 warps/SM %2d : %7.2f TFLOP/s\n", warps, fl / ms / 1e9);
    ms = timeit([&] { dmma16_kernel<<<blocks, threads>>>(out, iters); });
    fl = 2.0 * 2048 * 4 * iters * (double)blocks * (threads / 32);
    printf("DMMA m16n8k16 warps/SM %2d : %7.2f TFLOP/s\n", warps, fl / ms / 1e9);
  }
  return 0;
}

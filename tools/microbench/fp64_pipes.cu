// Micro-benchmark: FP64 vector pipe (DFMA), FP64 tensor pipe (DMMA m8n8k4) and both at once on
// one B200.  Answers "are the two pipes additive?" before any gate kernel is reshaped for DMMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu && ./fp64_pipes
#include <cuda_runtime.h>
#include <stdio.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int MODE>  // 0: DFMA only, 1: DMMA only, 2: both interleaved
__global__ void __launch_bounds__(256) k(double* out, int iters, double s) {
  double f[16], c[8][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) f[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = i;
  const double a = 1.0 + s, b = 1e-9 + s;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 2) {
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = fma(f[i], a, b);
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = fma(f[i], a, b);
    }
    if (MODE == 1 || MODE == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a, b);
    }
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) r += f[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) r += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, int ctas_per_sm) {
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out;
  cudaMalloc(&out, sizeof(double) * 256 * sms * 8);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<sms * ctas_per_sm, 256>>>(out, 100, 0.0);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<sms * ctas_per_sm, 256>>>(out, iters, 0.0);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double thr = (double)sms * ctas_per_sm * 256;
  double fma_v = (MODE == 0 || MODE == 2) ? thr * iters * 32.0 : 0;           // DFMA lanes
  double fma_t = (MODE == 1 || MODE == 2) ? thr / 32 * iters * 8.0 * 256 : 0;  // 8x8x4 FMA per warp DMMA
  printf("%-28s ctas/sm=%d  %.3f ms  vector %.2f TFMA/s  tensor %.2f TFMA/s  total %.2f TFLOP/s\n", name, ctas_per_sm, ms,
         fma_v / ms / 1e9, fma_t / ms / 1e9, 2 * (fma_v + fma_t) / ms / 1e9);
  cudaFree(out);
}

int main() {
  for (int c = 1; c <= 4; c *= 2) {
    run<0>("DFMA only", c);
    run<1>("DMMA m8n8k4 only", c);
    run<2>("DFMA + DMMA interleaved", c);
  }
  return 0;
}

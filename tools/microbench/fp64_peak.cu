// fp64 peak microbenchmark for the GP stage's roofline: register-resident DFMA chains and DMMA (mma.sync.m8n8k4.f64)
// chains, 8 independent accumulators per thread, no memory traffic.  nvcc -arch=sm_100a -O3 fp64_peak.cu -o fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double x[8];
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma_kernel(double* out, int iters, double a, double b) {
  double c[8][2];
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int blocks = p.multiProcessorCount * 4, threads = 256, iters = 20000;
  double* out; cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int which = 0; which < 2; ++which) {
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      if (which == 0) dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
      else dmma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    // DFMA: 8 fma per thread per iter; DMMA: 8 mma per warp per iter, 8*8*4 = 256 fma each
    const double fma = which == 0 ? (double)blocks * threads * iters * 8.0 : (double)blocks * (threads / 32) * iters * 8.0 * 256.0;
    printf("{\"op\": \"%s\", \"ms\": %.3f, \"tflops\": %.2f, \"sms\": %d}\n", which == 0 ? "dfma" : "dmma_m8n8k4", best,
           2.0 * fma / (best * 1e-3) / 1e12, p.multiProcessorCount);
  }
  return 0;
}

// fp64 FMA peak and latency on this GPU ("measure an fp64 FMA peak and report K0 against both roofs").
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak profiles/fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_fma(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
static double run(int blocks, int threads, int iters, double* d) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k_fma<ILP><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  k_fma<ILP><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double* d;
  cudaMalloc(&d, sizeof(double) * 148 * 32 * 1024);
  const int iters = 1 << 16;
  // latency: one warp, one dependent chain
  double ms = run<1>(1, 32, iters, d);
  printf("{\"sms\": %d, \"clock_mhz\": %.0f, \"dfma_latency_cycles\": %.2f", p.multiProcessorCount, clk / 1e3,
         ms * 1e-3 * clk * 1e3 / iters);
  // throughput: all SMs, many warps, 8 chains per thread
  const int blocks = p.multiProcessorCount * 4, threads = 512;
  ms = run<8>(blocks, threads, iters, d);
  double fmas = (double)blocks * threads * 8 * iters;
  printf(", \"dfma_per_s\": %.4g, \"fp64_tflops\": %.2f, \"dfma_per_clk_per_sm\": %.1f", fmas / (ms * 1e-3), 2 * fmas / (ms * 1e-3) / 1e12,
         fmas / (ms * 1e-3) / (clk * 1e3) / p.multiProcessorCount);
  // what one scheduler sustains with W warps x ILP chains: warps per SM = 4 * W
  for (int w = 1; w <= 8; w *= 2) {
    double m1 = run<1>(p.multiProcessorCount, 128 * w, iters, d), m2 = run<2>(p.multiProcessorCount, 128 * w, iters, d),
           m4 = run<4>(p.multiProcessorCount, 128 * w, iters, d);
    double base = (double)p.multiProcessorCount * 128 * w * iters / (clk * 1e3) / p.multiProcessorCount;
    printf(", \"w%d_ilp1\": %.1f, \"w%d_ilp2\": %.1f, \"w%d_ilp4\": %.1f", w, base / (m1 * 1e-3), w, 2 * base / (m2 * 1e-3), w,
           4 * base / (m4 * 1e-3));
  }
  printf("}\n");
  return 0;
}

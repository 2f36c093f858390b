// Measured FP64 FMA issue rate of the device (developer tool; the roofline cross-check of K1 / K9 in DESIGN.md).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tools/fp64_peak tools/fp64_peak.cu && build/tools/fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void k_dfma(double* out, double a, double b, int iters) {
  double acc[CHAINS];
#pragma unroll
  for (int q = 0; q < CHAINS; ++q) acc[q] = threadIdx.x * 1e-9 + q;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int q = 0; q < CHAINS; ++q) acc[q] = fma(acc[q], a, b);
  }
  double s = 0;
#pragma unroll
  for (int q = 0; q < CHAINS; ++q) s += acc[q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// three distinct register-pair sources per DFMA (what a register-tiled FIR / butterfly issues), 2 * CHAINS operands live
template <int CHAINS>
__global__ void k_dfma3(double* out, const double* in, int iters) {
  double acc[CHAINS], t[CHAINS], w[CHAINS];
#pragma unroll
  for (int q = 0; q < CHAINS; ++q) { acc[q] = threadIdx.x * 1e-9 + q; t[q] = in[q]; w[q] = in[CHAINS + q]; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int q = 0; q < CHAINS; ++q) acc[q] = fma(t[q], w[(q + 1) % CHAINS], acc[q]);
#pragma unroll
    for (int q = 0; q < CHAINS; ++q) acc[q] = fma(t[(q + 3) % CHAINS], w[q], acc[q]);
  }
  double s = 0;
#pragma unroll
  for (int q = 0; q < CHAINS; ++q) s += acc[q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  double* out;
  const int threads = 256, blocks = p.multiProcessorCount * 8, iters = 20000;
  cudaMalloc(&out, sizeof(double) * threads * blocks);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k_dfma<12><<<blocks, threads>>>(out, 0.999999, 1e-7, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma = (double)threads * blocks * iters * 12;
    printf("%s: %d SMs, %.3f ms, %.2f TFMA/s FP64 = %.2f TFLOP/s, %.1f FMA/clk/SM at %d MHz\n", p.name,
           p.multiProcessorCount, ms, fma / ms / 1e9, 2 * fma / ms / 1e9,
           fma / (ms * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3), p.clockRate / 1000);
  }
  double* in;
  cudaMalloc(&in, sizeof(double) * 64);
  cudaMemset(in, 0, sizeof(double) * 64);
  for (int warps = 16; warps >= 1; warps /= 2) {  // occupancy sweep: CTAs of 32*warps threads, one CTA per SM
    cudaEventRecord(e0);
    k_dfma3<12><<<p.multiProcessorCount, 32 * warps>>>(out, in, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma = (double)32 * warps * p.multiProcessorCount * iters * 24;
    printf("3-operand DFMA, %2d warps/SM: %.3f ms, %.2f TFMA/s, %.1f FMA/clk/SM\n", warps, ms, fma / ms / 1e9,
           fma / (ms * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3));
  }
  return 0;
}

// Development microbenchmark: dependent-issue latency and per-SM throughput of DFMA on the box's GPU
// (what bounds a one-thread-per-voxel FP64 kernel at 2 warps per scheduler).  Not part of the product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_variants/fp64_latency tools/micro/fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void k_chain(double* out, long long* cyc, int iters, double a, double b) {
  double x[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 1e-3 + c;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 16; ++r)
#pragma unroll
      for (int c = 0; c < CHAINS; ++c) x[c] = fma(x[c], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int CHAINS>
void run(int warps_per_cta, int ctas, double* d_out, long long* d_cyc) {
  const int iters = 2000;
  k_chain<CHAINS><<<ctas, 32 * warps_per_cta>>>(d_out, d_cyc, iters, 0.999999, 1e-7);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k_chain<CHAINS><<<ctas, 32 * warps_per_cta>>>(d_out, d_cyc, iters, 0.999999, 1e-7);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long cyc; cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
  const double n = (double)iters * 16 * CHAINS;
  printf("chains %2d warps/CTA %2d CTAs %4d: %.2f cycles per dependent DFMA step (%.2f per DFMA per warp), %.2f TFLOP/s\n", CHAINS,
         warps_per_cta, ctas, (double)cyc / (iters * 16.0), (double)cyc / n, 2.0 * n * 32 * warps_per_cta * ctas / (ms * 1e-3) / 1e12);
}

int main() {
  double* d_out; long long* d_cyc;
  cudaMalloc(&d_out, sizeof(double) * 1024 * 1024 * 4); cudaMalloc(&d_cyc, 8);
  run<1>(1, 1, d_out, d_cyc);      // pure latency
  run<2>(1, 1, d_out, d_cyc);
  run<4>(1, 1, d_out, d_cyc);
  run<8>(1, 1, d_out, d_cyc);
  run<1>(4, 148, d_out, d_cyc);    // 1 warp per scheduler
  run<1>(8, 148, d_out, d_cyc);    // 2 warps per scheduler: the update kernel's occupancy
  run<2>(8, 148, d_out, d_cyc);
  run<4>(8, 148, d_out, d_cyc);
  run<8>(8, 148, d_out, d_cyc);
  run<4>(16, 148, d_out, d_cyc);
  run<8>(32, 148, d_out, d_cyc);
  return 0;
}

// fp64 issue rate of one SM on this GPU: N independent DFMA chains per thread, W warps per SM, measured in DFMA lanes / clk / SM.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_rate fp64_rate.cu ; run: ./fp64_rate
#include <cstdio>
#include <cuda_runtime.h>
template <int CHAINS>
__global__ void k(double* out, int iters, double a, double b)
{
  double x[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CHAINS>
void run(int threads, int blocks_per_sm, int sms, double clk_ghz)
{
  double* d;
  cudaMalloc(&d, sizeof(double) * threads * blocks_per_sm * sms);
  const int iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<CHAINS><<<blocks_per_sm * sms, threads>>>(d, 64, 0.999, 1e-3);
  cudaEventRecord(e0);
  k<CHAINS><<<blocks_per_sm * sms, threads>>>(d, iters, 0.999, 1e-3);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double dfma = (double)iters * 8 * CHAINS * threads * blocks_per_sm;   // per SM
  printf("chains %d  warps/SM %3d : %.1f DFMA lanes/clk/SM at %.3f GHz (%.2f ms)\n", CHAINS, threads * blocks_per_sm / 32,
         dfma / (ms * 1e-3 * clk_ghz * 1e9), clk_ghz, ms);
  cudaFree(d);
}
int main()
{
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const double ghz = p.clockRate * 1e-6;
  printf("%s, %d SMs, %.3f GHz nominal\n", p.name, p.multiProcessorCount, ghz);
  const int sms = p.multiProcessorCount;
  run<1>(256, 1, sms, ghz); run<1>(256, 2, sms, ghz); run<1>(256, 4, sms, ghz);
  run<2>(256, 1, sms, ghz); run<2>(256, 2, sms, ghz); run<2>(256, 4, sms, ghz);
  run<4>(256, 1, sms, ghz); run<4>(256, 2, sms, ghz); run<4>(128, 1, sms, ghz);
  run<8>(128, 1, sms, ghz); run<8>(256, 1, sms, ghz); run<8>(256, 2, sms, ghz);
  return 0;
}

// Does a non-fp64 instruction issue in the shadow of an fp64 one?  8 independent DFMA chains per thread (enough to saturate the
// fp64 pipe alone, see fp64_rate.cu) interleaved with M independent integer multiply-adds per DFMA.  Prints DFMA lanes / clk / SM
// and total warp-instructions / clk / SMSP for M = 0, 1, 2, 3.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_mix fp64_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int M>
__global__ void k(double* out, int* iout, int iters, double a, double b, int ia, int ib)
{
  double x[8];
  int y[8 * (M ? M : 1)];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
  for (int i = 0; i < 8 * (M ? M : 1); ++i) y[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        x[i] = fma(x[i], a, b);
#pragma unroll
        for (int m = 0; m < M; ++m) y[i * (M ? M : 1) + m] = y[i * (M ? M : 1) + m] * ia + ib;
      }
  }
  double s = 0;
  int t = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
#pragma unroll
  for (int i = 0; i < 8 * (M ? M : 1); ++i) t += y[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  iout[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
template <int M>
void run(int sms, double ghz)
{
  const int threads = 256, bps = 2;
  double* d; int* di;
  cudaMalloc(&d, sizeof(double) * threads * bps * sms);
  cudaMalloc(&di, sizeof(int) * threads * bps * sms);
  const int iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<M><<<bps * sms, threads>>>(d, di, 64, 0.999, 1e-3, 3, 7);
  cudaEventRecord(e0);
  k<M><<<bps * sms, threads>>>(d, di, iters, 0.999, 1e-3, 3, 7);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double clk = ms * 1e-3 * ghz * 1e9;
  const double dfma_lanes = (double)iters * 32 * threads * bps;              // per SM
  const double warp_instr = (double)iters * 32 * (1 + M) * (threads / 32) * bps / 4;   // per SMSP
  printf("M = %d integer ops per DFMA: %.1f DFMA lanes/clk/SM, %.2f warp-instructions/clk/SMSP (%.2f ms)\n", M, dfma_lanes / clk, warp_instr / clk, ms);
  cudaFree(d); cudaFree(di);
}
int main()
{
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const double ghz = p.clockRate * 1e-6;
  printf("%s, %d SMs, %.3f GHz nominal\n", p.name, p.multiProcessorCount, ghz);
  run<0>(p.multiProcessorCount, ghz); run<1>(p.multiProcessorCount, ghz); run<2>(p.multiProcessorCount, ghz); run<3>(p.multiProcessorCount, ghz);
  return 0;
}

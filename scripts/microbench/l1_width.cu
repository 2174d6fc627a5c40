// L1 -> register-file throughput of LDG.64 / .128 / .256 on sm_100a, coalesced and scattered inside an L1-resident
// window.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l1_width l1_width.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int W> struct Ld;
template <> struct Ld<8> {
  static __device__ __forceinline__ double ld(const char* p) { double r; asm volatile("ld.global.ca.f64 %0, [%1];" : "=d"(r) : "l"(p)); return r; }
};
template <> struct Ld<16> {
  static __device__ __forceinline__ double ld(const char* p) { double a, b; asm volatile("ld.global.ca.v2.f64 {%0,%1}, [%2];" : "=d"(a), "=d"(b) : "l"(p)); return a + b; }
};
template <> struct Ld<32> {
  static __device__ __forceinline__ double ld(const char* p) { double a, b, c, d; asm volatile("ld.global.ca.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p)); return (a + b) + (c + d); }
};
// window bytes per CTA; pattern 0 = coalesced, 1 = scattered (unit = W bytes), 2 = scattered 8 lanes per 128-B line group
template <int W, int PAT>
__global__ void __launch_bounds__(256) k(const char* base, int window, int iters, double* out, long long* cyc)
{
  const char* win = base + (size_t)blockIdx.x * window;
  const int units = window / W;
  unsigned s = threadIdx.x * 2654435761u + 12345u;
  double acc = 0;
  // warm the window into L1
  for (int i = threadIdx.x; i < units; i += 256) acc += Ld<W>::ld(win + (size_t)i * W);
  __syncthreads();
  long long t0 = clock64();
  int u = threadIdx.x % units;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (PAT == 0) { u += 256; if (u >= units) u -= units; }
      else { s = s * 1664525u + 1013904223u; u = (s >> 8) % units; }
      acc += Ld<W>::ld(win + (size_t)u * W);
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * 256 + threadIdx.x] = acc;
}
template <int W, int PAT>
void run(const char* d, double* o, long long* c, int window, int ctas_per_sm)
{
  int grid = 148 * ctas_per_sm, iters = 2000;
  k<W, PAT><<<grid, 256>>>(d, window, 10, o, c);
  cudaDeviceSynchronize();
  k<W, PAT><<<grid, 256>>>(d, window, iters, o, c);
  cudaDeviceSynchronize();
  long long h[148 * 4];
  cudaMemcpy(h, c, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < grid; ++i) mean += h[i];
  mean /= grid;
  double bytes_per_sm = (double)ctas_per_sm * 256 * iters * 8 * W;
  printf("W=%2d pat=%d window=%6d ctas/sm=%d : %.1f B/clk/SM  (%.2f clk per warp-LDG)\n", W, PAT, window, ctas_per_sm,
         bytes_per_sm / mean, mean / (iters * 8.0 * 8 * ctas_per_sm));
}
int main()
{
  char* d; double* o; long long* c;
  cudaMalloc(&d, 148 * 4 * 65536); cudaMemset(d, 0, 148 * 4 * 65536);
  cudaMalloc(&o, 148 * 4 * 256 * 8); cudaMalloc(&c, 148 * 4 * 8);
  for (int cps = 1; cps <= 2; ++cps)
    for (int window : {16384, 65536}) {
      run<8, 0>(d, o, c, window, cps); run<16, 0>(d, o, c, window, cps); run<32, 0>(d, o, c, window, cps);
      run<8, 1>(d, o, c, window, cps); run<16, 1>(d, o, c, window, cps); run<32, 1>(d, o, c, window, cps);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

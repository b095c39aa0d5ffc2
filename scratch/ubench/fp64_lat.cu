// DFMA latency / throughput probe: CHAINS independent dependent-chains per thread, warps per SM varied.
#include <cstdio>
#include <cuda_runtime.h>
template <int CHAINS>
__global__ void k(double *out, int iters, double a, double b)
{
  double x[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x + c;
  for (int i = 0; i < iters; ++i)
    {
#pragma unroll
      for (int u = 0; u < 16; ++u)
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) x[c] = fma(x[c], a, b);
    }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CHAINS> void run(int warps_per_sm, double *out)
{
  int sms = 148, iters = 2000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<CHAINS><<<sms, warps_per_sm * 32>>>(out, 10, 0.999, 1e-9);
  cudaEventRecord(e0);
  k<CHAINS><<<sms, warps_per_sm * 32>>>(out, iters, 0.999, 1e-9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int clk;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double n_per_warp = (double)iters * 16 * CHAINS;
  double cycles = ms * 1e-3 * 1.9e9;   // approximate clock
  printf("chains %d warps/SM %2d: %.3f ms  %.2f cyc per DFMA per warp (at 1.9GHz)  %.2f TFLOP/s\n", CHAINS, warps_per_sm, ms,
         cycles / n_per_warp, 2.0 * 32 * n_per_warp * warps_per_sm * sms / (ms * 1e-3) * 1e-12);
}
int main()
{
  double *out;
  cudaMalloc(&out, 148 * 1024 * 8);
  for (int w : {4, 8, 16, 32})
    {
      run<1>(w, out);
      run<2>(w, out);
      run<4>(w, out);
      run<8>(w, out);
    }
  return 0;
}

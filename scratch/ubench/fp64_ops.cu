// FP64 pipe throughput by operand shape: 16 warps/SM, 4 chains
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double *out, int iters, double a, double b)
{
  constexpr int C = 4;
  double x[C], y[C], z[C];
  for (int c = 0; c < C; ++c) { x[c] = threadIdx.x + c; y[c] = 0.999 + 1e-6 * (threadIdx.x + c); z[c] = 1e-9 * (threadIdx.x + 2 * c + 1); }
  for (int i = 0; i < iters; ++i)
    {
#pragma unroll
      for (int u = 0; u < 16; ++u)
#pragma unroll
        for (int c = 0; c < C; ++c)
          {
            if (MODE == 0) x[c] = fma(x[c], a, b);             // 1 reg + 2 uniform
            if (MODE == 1) x[c] = fma(x[c], y[c], b);          // 2 regs + uniform
            if (MODE == 2) x[c] = fma(x[c], y[c], z[c]);       // 3 distinct regs
            if (MODE == 3) x[c] = fma(y[c], z[c], x[c]);       // accumulate form
            if (MODE == 4) x[c] = x[c] + y[c];                 // DADD 2 regs
            if (MODE == 5) x[c] = x[c] * y[c];                 // DMUL 2 regs
            if (MODE == 6) x[c] = __fma_rz(fabs(x[c]), a, 0.5); // DFMA.RZ |x|
            if (MODE == 7) x[c] = fma(x[c], x[c], z[c]);       // 2 distinct regs
          }
    }
  double s = 0;
  for (int c = 0; c < C; ++c) s += x[c] + y[c] + z[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(double *out, const char *name)
{
  int sms = 148, iters = 1000, wps = 16;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<sms, wps * 32>>>(out, 10, 0.999, 1e-9);
  cudaEventRecord(e0);
  k<MODE><<<sms, wps * 32>>>(out, iters, 0.999, 1e-9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double ninst = (double)iters * 16 * 4 * wps * sms;      // warp-instructions
  printf("%-28s %.3f ms  %.3e warp-inst/s  (%.2f per clk per SM at 1.9GHz)\n", name, ms, ninst / (ms * 1e-3), ninst / (ms * 1e-3) / 148 / 1.9e9);
}
int main()
{
  double *out;
  cudaMalloc(&out, 148 * 1024 * 8);
  run<0>(out, "fma(x,U,U)");
  run<1>(out, "fma(x,y,U)");
  run<2>(out, "fma(x,y,z)");
  run<3>(out, "fma(y,z,x)");
  run<4>(out, "x+y");
  run<5>(out, "x*y");
  run<6>(out, "fma_rz(|x|,U,0.5)");
  run<7>(out, "fma(x,x,z)");
  return 0;
}

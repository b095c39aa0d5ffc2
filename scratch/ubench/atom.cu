// atomic throughput probe: random bins (1000), 16 warps/SM, 296 CTAs x 256 threads
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned lcg(unsigned &s) { s = s * 1664525u + 1013904223u; return s >> 8; }
template <int MODE>
__global__ void __launch_bounds__(256) k(double *ghist, int iters, int active_lanes)
{
  __shared__ unsigned long long sh64[1024];
  __shared__ unsigned sh32[1024];
  __shared__ unsigned sh32b[1024];
  __shared__ double shd[4][1024];
  for (int i = threadIdx.x; i < 1024; i += 256) { sh64[i] = 0; sh32[i] = 0; sh32b[i] = 0; for (int w = 0; w < 4; ++w) shd[w][i] = 0; }
  __syncthreads();
  unsigned s = threadIdx.x * 7919u + blockIdx.x * 104729u + 1;
  double *rep = ghist + (size_t)blockIdx.x * 1024;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool act = lane < active_lanes;
  double acc = 0;
  for (int i = 0; i < iters; ++i)
    {
      const int b = lcg(s) % 1000;
      const double v = 1.0 + 1e-3 * b;
      if (MODE == 0) { if (act) atomicAdd(rep + b, v); }
      if (MODE == 1) { if (act) atomicAdd(&sh64[b], (unsigned long long)(v * 1048576.0)); }
      if (MODE == 2) { if (act) atomicAdd(&sh32[b], 1u); }
      if (MODE == 3) { if (act) atomicAdd(&shd[0][b], v); }
      if (MODE == 4)
        { // per-warp private histogram, conflicts resolved with match.any
          unsigned m = __match_any_sync(0xffffffffu, act ? b : -1 - lane);
          int rank = __popc(m & ((1u << lane) - 1));
          int maxr = __reduce_max_sync(0xffffffffu, rank);
          for (int r = 0; r <= maxr; ++r)
            {
              if (act && rank == r) shd[warp & 3][b] += v;
              __syncwarp();
            }
        }
      if (MODE == 5) { acc += v; }   // baseline loop cost
      if (MODE == 6)
        { // two-limb fixed point with carry (the scheme of evaluate.cuh)
          if (act)
            {
              const long long x = (long long)(v * 68719476736.0);
              const unsigned xlo = (unsigned)(unsigned long long)x, xhi = (unsigned)((unsigned long long)x >> 32);
              const unsigned old = atomicAdd(&sh32[b], xlo);
              const unsigned carry = (old + xlo < old) ? 1u : 0u;
              atomicAdd(&sh32b[b], xhi + carry);
            }
        }
      if (MODE == 7)
        { // two independent limbs, no return value
          if (act)
            {
              const long long x = (long long)(v * 68719476736.0);
              atomicAdd(&sh32[b], (unsigned)(x & 0xfffff));
              atomicAdd(&sh32b[b], (unsigned)(x >> 20));
            }
        }
    }
  __syncthreads();
  if (MODE != 0)
    for (int i = threadIdx.x; i < 1000; i += 256) rep[i] += (double)sh64[i] + sh32[i] + sh32b[i] + shd[0][i] + shd[warp & 3][i] + acc;
}
template <int MODE> void run(double *g, const char *name, int active)
{
  int iters = 2000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<296, 256>>>(g, 10, active);
  cudaEventRecord(e0);
  k<MODE><<<296, 256>>>(g, iters, active);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double nwarp = (double)iters * 296 * 8;
  printf("%-34s lanes %2d: %.3f ms  %.1f cyc per warp-instr per SM (1.9GHz)  %.3e lane-atomics/s\n", name, active, ms,
         ms * 1e-3 * 1.9e9 / (nwarp / 148), nwarp * active / (ms * 1e-3));
}
int main()
{
  double *g;
  cudaMalloc(&g, 296 * 1024 * 8);
  cudaMemset(g, 0, 296 * 1024 * 8);
  for (int a : {32, 16})
    {
      run<5>(g, "baseline loop", a);
      run<0>(g, "RED.F64 global per-CTA replica", a);
      run<1>(g, "ATOMS u64 shared", a);
      run<2>(g, "ATOMS u32 shared", a);
      run<3>(g, "atomicAdd double shared (CAS)", a);
      run<6>(g, "2-limb fixed point with carry", a);
      run<7>(g, "2 independent limbs no return", a);
    }
  return 0;
}

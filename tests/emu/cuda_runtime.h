// tests/emu/cuda_runtime.h — TEST INFRASTRUCTURE, not part of the product.
//
// A stand-in for <cuda_runtime.h> that lets g++ compile the repository's CUDA sources
// (qmc4_b200/csrc/*.cu, *.cuh) as host C++ and run every kernel with one OS thread per CUDA thread
// (CTAs one after the other).  It exists so that the kernels' *logic* — barriers, shuffles, shared
// memory layout, parent maps, prefix sums, fixed-point histograms — can be exercised by the parity
// tests in a container without a GPU (tests/test_emulated_kernels.py).  It is thousands of times
// slower than one CPU core running the oracle, it is built only by tests/emu/Makefile into
// tests/emu/libqmc4b200_emu.so, and nothing in qmc4_b200/, bench.py or __graft_entry__.py can load
// it: the shipped library is the nvcc build and still fails loudly without a CUDA device.
//
// Semantics emulated: threadIdx/blockIdx/blockDim/gridDim, dynamic + static shared memory,
// __syncthreads, named barriers (bar.sync id, n), __syncwarp, full-mask warp shuffles, shared /
// global atomics, the integer/bit intrinsics used by the sources.  Streams are synchronous.
#pragma once
#define QMC_EMULATE 1

#include <algorithm>
#include <cfenv>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>

// ---- qualifiers ---------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static
#define __align__(n) alignas(n)

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

namespace emu {
struct ThreadCtx {
  uint3 tid{0, 0, 0}, bid{0, 0, 0};
  dim3 bdim, gdim;
  unsigned char *dyn_smem = nullptr;
};
extern thread_local ThreadCtx tls;
void barrier(int id, int nthreads);                 // bar.sync id, n   (id 0 = __syncthreads)
void warp_barrier();
uint32_t warp_exchange(uint32_t v, int src_lane);   // value of v on lane src_lane
typedef void (*kernel_tramp)(void **args);
}  // namespace emu

#define threadIdx (emu::tls.tid)
#define blockIdx (emu::tls.bid)
#define blockDim (emu::tls.bdim)
#define gridDim (emu::tls.gdim)

// ---- synchronisation and shuffles ---------------------------------------------------------------
static inline void __syncthreads() { emu::barrier(0, (int)emu::tls.bdim.x); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }
static inline int __shfl_sync(unsigned, int v, int src) { return (int)emu::warp_exchange((uint32_t)v, src & 31); }
static inline int __shfl_xor_sync(unsigned, int v, int m)
{
  return (int)emu::warp_exchange((uint32_t)v, ((int)(emu::tls.tid.x & 31) ^ m) & 31);
}
static inline int __shfl_up_sync(unsigned, int v, int d)
{
  const int lane = (int)(emu::tls.tid.x & 31);
  return (int)emu::warp_exchange((uint32_t)v, lane >= d ? lane - d : lane);
}

// ---- atomics ------------------------------------------------------------------------------------
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v)
{
  return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}
static inline unsigned atomicExch(unsigned *p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
static inline double atomicAdd(double *p, double v)
{
  uint64_t *q = reinterpret_cast<uint64_t *>(p);
  uint64_t old = __atomic_load_n(q, __ATOMIC_RELAXED);
  for (;;)
    {
      double cur;
      std::memcpy(&cur, &old, 8);
      double nxt = cur + v;
      uint64_t want;
      std::memcpy(&want, &nxt, 8);
      if (__atomic_compare_exchange_n(q, &old, want, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return cur;
    }
}

// ---- bit / arithmetic intrinsics ----------------------------------------------------------------
static inline int __double2hiint(double d)
{
  uint64_t u;
  std::memcpy(&u, &d, 8);
  return (int)(uint32_t)(u >> 32);
}
static inline int __double2loint(double d)
{
  uint64_t u;
  std::memcpy(&u, &d, 8);
  return (int)(uint32_t)u;
}
static inline double __hiloint2double(int hi, int lo)
{
  uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
  double d;
  std::memcpy(&d, &u, 8);
  return d;
}
static inline double __fma_rz(double a, double b, double c)
{ /* fused multiply-add under round-toward-zero (per-thread rounding mode; built with -frounding-math) */
  volatile double va = a, vb = b, vc = c;
  const int old = fegetround();
  fesetround(FE_TOWARDZERO);
  volatile double r = std::fma(va, vb, vc);
  fesetround(old);
  return r;
}
static inline double __dmul_rn(double a, double b)
{
  volatile double r = a * b;       /* rounded product: the build uses -ffp-contract=off */
  return r;
}
static inline double __dadd_rn(double a, double b)
{
  volatile double r = a + b;
  return r;
}
static inline long long __double_as_longlong(double d)
{
  long long u;
  std::memcpy(&u, &d, 8);
  return u;
}
static inline long long __double2ll_rn(double d)
{ /* cvt.rni.s64.f64 saturates */
  if (d >= 9223372036854775807.0) return 9223372036854775807LL;
  if (d <= -9223372036854775808.0) return (-9223372036854775807LL - 1);
  return std::llrint(d);
}
static inline double __ll2double_rn(long long v) { return (double)v; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline float emu_log2f(float x) { return log2f(x); }
static inline void emu_sincosf(float x, float *s, float *c)
{
  *s = sinf(x);
  *c = cosf(x);
}
#define __log2f emu_log2f     /* glibc owns the reserved names */
#define __sincosf emu_sincosf
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }

using std::max;
using std::min;

// ---- runtime API (synchronous, host memory) -----------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2 };
typedef struct emu_stream *cudaStream_t;
typedef struct emu_event *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
struct cudaFuncAttributes { int numRegs = 0; size_t sharedSizeBytes = 0; int maxThreadsPerBlock = 1024; };
struct cudaDeviceProp {
  char name[256];
  int major, minor, multiProcessorCount;
  size_t sharedMemPerBlockOptin, totalGlobalMem;
  int clockRate;
};

const char *cudaGetErrorString(cudaError_t);
cudaError_t cudaGetLastError();
cudaError_t cudaGetDeviceCount(int *);
cudaError_t cudaGetDevice(int *);
cudaError_t cudaSetDevice(int);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *, int);
cudaError_t cudaDeviceGetAttribute(int *, cudaDeviceAttr, int);
cudaError_t emu_malloc(void **, size_t);
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) { return emu_malloc((void **)p, n); }
cudaError_t cudaFree(void *);
cudaError_t cudaMemcpy(void *, const void *, size_t, cudaMemcpyKind);
cudaError_t cudaMemcpyAsync(void *, const void *, size_t, cudaMemcpyKind, cudaStream_t = nullptr);
cudaError_t cudaMemset(void *, int, size_t);
cudaError_t cudaMemsetAsync(void *, int, size_t, cudaStream_t = nullptr);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *, unsigned);
cudaError_t cudaStreamCreateWithPriority(cudaStream_t *, unsigned, int);
cudaError_t cudaDeviceGetStreamPriorityRange(int *, int *);
cudaError_t cudaHostAlloc(void **, size_t, unsigned);
cudaError_t cudaFreeHost(void *);
cudaError_t cudaStreamDestroy(cudaStream_t);
cudaError_t cudaStreamSynchronize(cudaStream_t);
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0);
cudaError_t cudaEventCreate(cudaEvent_t *);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *, unsigned);
cudaError_t cudaEventDestroy(cudaEvent_t);
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr);
cudaError_t cudaEventSynchronize(cudaEvent_t);
cudaError_t cudaEventElapsedTime(float *, cudaEvent_t, cudaEvent_t);
cudaError_t cudaFuncSetAttribute(const void *, cudaFuncAttribute, int);
cudaError_t cudaFuncGetAttributes(cudaFuncAttributes *, const void *);
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *, const void *, int, size_t);
// `func` must be an emu::kernel_tramp (the sources hand out trampolines under QMC_EMULATE)
cudaError_t cudaLaunchKernel(const void *func, dim3 grid, dim3 block, void **args, size_t smem, cudaStream_t);

namespace emu {
// run `body` once per CUDA thread of a grid x block launch with `smem` bytes of dynamic shared memory
void launch(dim3 grid, dim3 block, size_t smem, void (*body)(void *), void *closure);
template <class F> static inline void launch_lambda(dim3 grid, dim3 block, size_t smem, F &&f)
{
  launch(grid, block, smem, [](void *c) { (*static_cast<F *>(c))(); }, &f);
}
}  // namespace emu

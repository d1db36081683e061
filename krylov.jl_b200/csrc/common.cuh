// common.cuh -- shared device/host helpers for libkrylov_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

namespace kb {

struct CudaError : std::runtime_error {
  explicit CudaError(const std::string& m) : std::runtime_error(m) {}
};

#define KB_CUDA(call)                                                                 \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      throw ::kb::CudaError(std::string(#call) + " failed: " + cudaGetErrorString(e__) + \
                            " (" __FILE__ ":" + std::to_string(__LINE__) + ")");     \
    }                                                                                 \
  } while (0)

constexpr int kBlock = 256;          // threads per CTA for streaming kernels
constexpr int kMaxPartials = 2048;   // upper bound on reduction grid size

// ---- floating-point without contraction -----------------------------------
// The reference (Julia, no implicit FMA) rounds every product before the add.
// Vector updates and the SpMV use these so they agree bit-for-bit with the
// sequential CPU oracle given the same scalars.
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float  mul_rn(float a, float b)   { return __fmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float  add_rn(float a, float b)   { return __fadd_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float  div_rn(float a, float b)   { return __fdiv_rn(a, b); }
__device__ __forceinline__ double sqrt_rn(double a) { return __dsqrt_rn(a); }
__device__ __forceinline__ float  sqrt_rn(float a)  { return __fsqrt_rn(a); }

template <class T> struct Eps;
template <> struct Eps<double> { static constexpr double v = 2.220446049250313e-16; };
template <> struct Eps<float>  { static constexpr float  v = 1.1920929e-07f; };

// ---- deterministic block reduction ----------------------------------------
// Fixed shuffle tree inside a warp, fixed order across warps: the result
// depends only on the values and blockDim, never on scheduling.
template <class T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// All threads of the CTA must call.  Result valid in thread 0.
template <class T, int NWARPS_MAX = 32>
__device__ __forceinline__ T block_sum(T v, T* smem /* >= NWARPS_MAX */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  if (lane == 0) smem[w] = v;
  __syncthreads();
  T r = T(0);
  if (w == 0) {
    r = lane < nw ? smem[lane] : T(0);
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}

// Grid-wide deterministic sum of K values per CTA ("last block finalises").
// Each CTA stores its K partials at part[k * gridDim.x + blockIdx.x]; the CTA
// that draws the last ticket re-reads all partials in index order and reduces
// them with the same fixed tree, so the total is independent of which CTA
// happens to be last.  Returns true (in every thread of that last CTA) and the
// totals in out[0..K) (valid in thread 0).
template <class T, int K>
__device__ __forceinline__ bool grid_sum_last(const T (&mine)[K], T* part, unsigned* ticket, T* smem, T (&out)[K]) {
  __shared__ bool is_last;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; k++) part[k * gridDim.x + blockIdx.x] = mine[k];
    __threadfence();
    unsigned t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
    if (is_last) *ticket = 0u;  // re-arm for the next launch (stream-ordered)
  }
  __syncthreads();
  if (!is_last) return false;
  __threadfence();
#pragma unroll
  for (int k = 0; k < K; k++) {
    T acc = T(0);
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) acc += __ldcg(&part[k * gridDim.x + i]);
    out[k] = block_sum(acc, smem);
  }
  return true;
}

// Opt a kernel into > 48 KB of dynamic shared memory, once per (kernel, device): the attribute belongs to the
// device's context, so a process that drives several GPUs must set it on each (blas1.cu).
void ensure_dyn_smem(const void* func, int bytes);

inline int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    KB_CUDA(cudaGetDevice(&dev));
    KB_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  }
  return n;
}

// Grid for a streaming kernel over n elements, `per_thread` elements each:
// a whole number of CTAs per SM (148 SMs on B200), capped by the work.
inline int stream_grid(long long n, int per_thread, int ctas_per_sm) {
  long long need = (n + (long long)kBlock * per_thread - 1) / ((long long)kBlock * per_thread);
  long long cap = (long long)sm_count() * ctas_per_sm;
  if (cap > kMaxPartials) cap = kMaxPartials;
  long long g = need < cap ? need : cap;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace kb

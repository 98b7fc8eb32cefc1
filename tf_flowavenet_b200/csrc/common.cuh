// Shared helpers for libflowavenet_b200 (sm_100a only).
#pragma once
#include <stdlib.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace fwn {

// thread-local error string behind fwn_last_error()
void set_error(const char* fmt, ...);
const char* get_error();

#define FWN_CHECK(cond, ...)                 \
  do {                                       \
    if (!(cond)) {                           \
      ::fwn::set_error(__VA_ARGS__);         \
      return 1;                              \
    }                                        \
  } while (0)

#define FWN_CUDA(expr)                                                                       \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      ::fwn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return 1;                                                                              \
    }                                                                                        \
  } while (0)

// FWN_SYNC_DEBUG=1 synchronises after every launch so an asynchronous fault is reported at the kernel that caused it
inline cudaError_t launch_status() {
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("FWN_SYNC_DEBUG");
    dbg = (e && e[0] == '1') ? 1 : 0;
  }
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess && dbg) {
    e = cudaDeviceSynchronize();
  }
  return e;
}
#define FWN_LAUNCH_CHECK() FWN_CUDA(::fwn::launch_status())

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// FWN_PDL=0 launches the tensor-core GEMM chain without programmatic dependent launch (diagnostics / A-B timing)
inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FWN_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

int num_sms();

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; result valid in thread 0.  `red` must hold >= 32 values.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    v = lane < nw ? red[lane] : T(0);
    v = warp_sum(v);
  }
  return v;
}

// activation storage types
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }
// two floats -> one 32-bit word of the 16-bit storage type (low half = first element)
template <typename T> __device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <> __device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// precision modes whose activations / GEMM operands are 16-bit (bf16 or fp16) on the tcgen05 engine
inline bool is_mixed(int precision) { return precision == 1 || precision == 2; }   // FWN_MIXED_BF16, FWN_MIXED_FP16

}  // namespace fwn

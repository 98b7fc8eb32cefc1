// Elementwise kernels of the 16-bit (bf16) training mode: the backward pass of the pieces that are not GEMMs, with bf16 tape /
// gradient tensors and fp32 arithmetic.  Bandwidth-bound: 16-byte accesses, grid-stride loops.
// Reference: the tf.gradients of model.py:133-135 (coupling), modules.py:124 (gate), modules.py:165,177-179 (ReLUs); train.py:62-63.
#include <cuda_bf16.h>

#include <algorithm>

#include "common.cuh"
#include "train.h"

namespace fwn {

static int ew_grid16(int64_t n) { return (int)std::min<int64_t>(std::max<int64_t>(cdiv(n, 256), 1), (int64_t)num_sms() * 16); }

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[2 * j] = __uint_as_float(w[j] << 16);
    f[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ uint32_t pk(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// AffineCoupling backward (model.py:133-135), bf16 (d log_s, d t) for the zero-conv GEMMs.  See affine_bwd_kernel (train_kernels.cu).
__global__ void affine_bwd16_kernel(float* __restrict__ dX, const float* __restrict__ Xpost, const float* __restrict__ net, int64_t ldn,
                                    __nv_bfloat16* __restrict__ dNet, int64_t ldd, int64_t rows, int Cx, int nq, const int* __restrict__ b_off,
                                    float inv_n) {
  const int64_t n = rows * nq;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / nq;
    const int p = (int)(i - row * nq);
    const int ob = __ldg(b_off + p);
    const float g = dX[row * Cx + ob];
    const float outb = __ldg(Xpost + row * Cx + ob);
    const float2 lt = *reinterpret_cast<const float2*>(net + row * ldn + 2 * p);
    const float el = expf(-lt.x);
    *reinterpret_cast<uint32_t*>(dNet + row * ldd + 2 * p) = pk(-g * outb + inv_n, -g * el);
    dX[row * Cx + ob] = g * el;
  }
}
int affine_bwd16(float* dX, const float* Xpost, const float* net, int64_t ldn, void* dNet, int64_t ldd, int64_t rows, int Cx, int nq,
                 const int* b_off, double n_total, cudaStream_t st) {
  if (ldd != 2 * nq) FWN_CUDA(cudaMemsetAsync(dNet, 0, (size_t)rows * ldd * 2, st));   // pad columns meet zero weights but must be finite
  affine_bwd16_kernel<<<ew_grid16(rows * nq), 256, 0, st>>>(dX, Xpost, net, ldn, reinterpret_cast<__nv_bfloat16*>(dNet), ldd, rows, Cx, nq, b_off,
                                                             (float)(1.0 / n_total));
  FWN_LAUNCH_CHECK();
  return 0;
}

// Gate backward (modules.py:124): o = tanh(f) sigmoid(g).  The forward pass kept o and s = sigmoid(g) (bf16):
//   t = tanh(f) = o / s;   d f = d o * s (1 - t^2) = d o * (s - o^2 / s);   d g = d o * t s (1 - s) = d o * o (1 - s)
// Output columns (2c, 2c+1) = (d f_c, d g_c), the column order of the gate GEMM's weights.
__global__ void gate_bwd16_kernel(const uint4* __restrict__ dO, const uint4* __restrict__ O, const uint4* __restrict__ S, uint4* __restrict__ dFG,
                                  int64_t n8) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float d[8], o[8], s[8];
    unpack8(__ldg(dO + i), d);
    unpack8(__ldg(O + i), o);
    unpack8(__ldg(S + i), s);
    uint32_t r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float df = s[j] > 1e-20f ? d[j] * (s[j] - o[j] * o[j] / s[j]) : 0.f;
      const float dg = d[j] * o[j] * (1.f - s[j]);
      r[j] = pk(df, dg);
    }
    dFG[2 * i] = make_uint4(r[0], r[1], r[2], r[3]);
    dFG[2 * i + 1] = make_uint4(r[4], r[5], r[6], r[7]);
  }
}
int gate_bwd16(const void* dO, const void* O, const void* S, void* dFG, int64_t n, cudaStream_t st) {
  FWN_CHECK(n % 8 == 0, "gate_bwd16: element count must be a multiple of 8");
  gate_bwd16_kernel<<<ew_grid16(n / 8), 256, 0, st>>>(reinterpret_cast<const uint4*>(dO), reinterpret_cast<const uint4*>(O),
                                                       reinterpret_cast<const uint4*>(S), reinterpret_cast<uint4*>(dFG), n / 8);
  FWN_LAUNCH_CHECK();
  return 0;
}

// y = h > 0 ? y : 0 in place (gradient through tf.nn.relu of the front conv, modules.py:165)
__global__ void relu_mask16_kernel(uint4* __restrict__ Y, const uint4* __restrict__ H, int64_t n8) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    uint4 y = Y[i];
    const uint4 h = __ldg(H + i);
    uint32_t yw[4] = {y.x, y.y, y.z, y.w};
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // bf16 > 0  <=>  sign bit clear and magnitude non-zero
      const uint32_t lo = hw[j] & 0xFFFFu, hi = hw[j] >> 16;
      if (!(lo != 0 && lo < 0x8000u)) yw[j] &= 0xFFFF0000u;
      if (!(hi != 0 && hi < 0x8000u)) yw[j] &= 0x0000FFFFu;
    }
    Y[i] = make_uint4(yw[0], yw[1], yw[2], yw[3]);
  }
}
int relu_mask16(void* Y, const void* H, int64_t n, cudaStream_t st) {
  FWN_CHECK(n % 8 == 0, "relu_mask16: element count must be a multiple of 8");
  relu_mask16_kernel<<<ew_grid16(n / 8), 256, 0, st>>>(reinterpret_cast<uint4*>(Y), reinterpret_cast<const uint4*>(H), n / 8);
  FWN_LAUNCH_CHECK();
  return 0;
}

}  // namespace fwn

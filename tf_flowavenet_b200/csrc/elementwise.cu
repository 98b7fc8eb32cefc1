// Bandwidth-bound flow ops in the reference's own layout ([rows, C] fp32, channels-last).
// One kernel per TF op site of SURVEY 2.3; all are coalesced grid-stride kernels with 128-bit
// accesses where the shape allows, warp-shuffle reductions and one atomic per CTA.
#include "common.cuh"
#include "kernels.h"

namespace fwn {

static inline int ew_grid(int64_t n_items, int threads) {
  int64_t blocks = cdiv(n_items, threads);
  int64_t cap = (int64_t)num_sms() * 16;  // multiple of the SM count, enough waves for latency hiding
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ---------------------------------------------------------------- squeeze / unsqueeze / change_order
// squeeze: y[b,t,2c+k] = x[b,2t+k,c].  Row t of y occupies the same 2C floats as rows 2t,2t+1 of x,
// so this is a [2,C]->[C,2] transpose inside each 2C-float group: fully coalesced both ways.
__global__ void squeeze_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, int C, bool inverse) {
  const int C2 = 2 * C;  // group width; for unsqueeze C is the OUTPUT channel count (= Cin/2)
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t grp = i / C2;
    int r = (int)(i - grp * C2);
    int src;
    if (!inverse) {  // out index r = 2c+k  <- in index k*C + c
      src = (r & 1) * C + (r >> 1);
    } else {         // out index r = k*C + c <- in index 2c+k
      int k = r / C, c = r - k * C;
      src = 2 * c + k;
    }
    y[i] = __ldg(x + grp * C2 + src);
  }
}

// 128-bit variant (C % 4 == 0): one thread moves 8 floats of one 2C-float group.
//   squeeze:   a = in[c0..c0+3] (row parity 0), b = in[C+c0..] (parity 1)  ->  out[2c0..2c0+7] = a0 b0 a1 b1 a2 b2 a3 b3
//   unsqueeze: the inverse de-interleave.
template <bool INVERSE>
__global__ void squeeze_vec_kernel(const float4* __restrict__ x, float4* __restrict__ y, int64_t n8, int C) {
  const int q = C / 4;  // float4 per half group
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t grp = i / q;
    const int c4 = (int)(i - grp * q);
    const int64_t base = grp * 2 * q;  // float4 index of the group's first element
    if (!INVERSE) {
      const float4 a = __ldg(x + base + c4), b = __ldg(x + base + q + c4);
      y[base + 2 * c4] = make_float4(a.x, b.x, a.y, b.y);
      y[base + 2 * c4 + 1] = make_float4(a.z, b.z, a.w, b.w);
    } else {
      const float4 u = __ldg(x + base + 2 * c4), v = __ldg(x + base + 2 * c4 + 1);
      y[base + c4] = make_float4(u.x, u.z, v.x, v.z);
      y[base + q + c4] = make_float4(u.y, u.w, v.y, v.w);
    }
  }
}
// C == 2: a group is one float4 (k0c0 k0c1 k1c0 k1c1) <-> (c0k0 c0k1 c1k0 c1k1): swap the middle pair (self-inverse)
__global__ void squeeze_c2_kernel(const float4* __restrict__ x, float4* __restrict__ y, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(x + i);
    y[i] = make_float4(v.x, v.z, v.y, v.w);
  }
}
static bool aligned16(const void* a, const void* b) { return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0; }

int squeeze(const float* x, float* y, int B, int T, int C, cudaStream_t st) {
  FWN_CHECK(T % 2 == 0, "squeeze: T=%d must be even (model.py:226)", T);
  int64_t n = (int64_t)B * T * C;
  if (n == 0) return 0;
  if (C == 1) {  // out[b,t,k] = x[b,2t+k,0]: the identity on memory
    if (x != y) FWN_CUDA(cudaMemcpyAsync(y, x, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  if (C == 2 && aligned16(x, y)) squeeze_c2_kernel<<<ew_grid(n / 4, 256), 256, 0, st>>>((const float4*)x, (float4*)y, n / 4);
  else if (C % 4 == 0 && aligned16(x, y)) squeeze_vec_kernel<false><<<ew_grid(n / 8, 256), 256, 0, st>>>((const float4*)x, (float4*)y, n / 8, C);
  else squeeze_kernel<<<ew_grid(n, 256), 256, 0, st>>>(x, y, n, C, false);
  FWN_LAUNCH_CHECK();
  return 0;
}
int unsqueeze(const float* x, float* y, int B, int T, int C, cudaStream_t st) {
  FWN_CHECK(C % 2 == 0, "unsqueeze: C=%d must be even (model.py:260)", C);
  int64_t n = (int64_t)B * T * C;
  if (n == 0) return 0;
  if (C == 2) {
    if (x != y) FWN_CUDA(cudaMemcpyAsync(y, x, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  if (C == 4 && aligned16(x, y)) squeeze_c2_kernel<<<ew_grid(n / 4, 256), 256, 0, st>>>((const float4*)x, (float4*)y, n / 4);
  else if ((C / 2) % 4 == 0 && aligned16(x, y)) squeeze_vec_kernel<true><<<ew_grid(n / 8, 256), 256, 0, st>>>((const float4*)x, (float4*)y, n / 8, C / 2);
  else squeeze_kernel<<<ew_grid(n, 256), 256, 0, st>>>(x, y, n, C / 2, true);
  FWN_LAUNCH_CHECK();
  return 0;
}

__global__ void change_order_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, int C) {
  const int h = C / 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t row = i / C;
    int c = (int)(i - row * C);
    y[i] = __ldg(x + row * C + (c < h ? c + h : c - h));
  }
}
// 128-bit variant: C/2 % 4 == 0 -> whole float4s move; C == 2 / 4 -> swizzle inside one float4
__global__ void change_order_vec_kernel(const float4* __restrict__ x, float4* __restrict__ y, int64_t n4, int C) {
  const int q = C / 4, hq = C / 8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    if (C == 2) {
      const float4 v = __ldg(x + i);
      y[i] = make_float4(v.y, v.x, v.w, v.z);
    } else if (C == 4) {
      const float4 v = __ldg(x + i);
      y[i] = make_float4(v.z, v.w, v.x, v.y);
    } else {
      const int64_t row = i / q;
      const int c4 = (int)(i - row * q);
      y[i] = __ldg(x + row * q + (c4 < hq ? c4 + hq : c4 - hq));
    }
  }
}
int change_order(const float* x, float* y, int64_t rows, int C, cudaStream_t st) {
  FWN_CHECK(C % 2 == 0, "change_order: C=%d must be even", C);
  int64_t n = rows * C;
  if (n == 0) return 0;
  if ((C == 2 || C == 4 || C % 8 == 0) && n % 4 == 0 && aligned16(x, y)) {
    change_order_vec_kernel<<<ew_grid(n / 4, 256), 256, 0, st>>>((const float4*)x, (float4*)y, n / 4, C);
    FWN_LAUNCH_CHECK();
    return 0;
  }
  change_order_kernel<<<ew_grid(n, 256), 256, 0, st>>>(x, y, n, C);
  FWN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- ActNorm
// y = (x + b) * exp(3 logs)  /  y = x * exp(-3 logs) - b.   float4 path when C % 4 == 0 or 4 % C == 0.
template <bool REV>
__global__ void actnorm_kernel(const float* __restrict__ x, const float* __restrict__ b, const float* __restrict__ logs,
                               float* __restrict__ y, int64_t n, int C) {
  extern __shared__ float sm[];  // [C] bias, [C] scale
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    sm[c] = b[c];
    sm[C + c] = expf(REV ? -3.f * logs[c] : 3.f * logs[c]);
  }
  __syncthreads();
  const bool vec = ((n & 3) == 0) && ((C % 4 == 0) || (4 % C == 0)) && ((((uintptr_t)x | (uintptr_t)y) & 15) == 0);
  if (vec) {
    const int64_t n4 = n >> 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
      float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
      int c0 = (int)((i * 4) % C);
      float* pv = reinterpret_cast<float*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int c = c0 + j;
        if (c >= C) c -= C * (c / C);
        pv[j] = REV ? pv[j] * sm[C + c] - sm[c] : (pv[j] + sm[c]) * sm[C + c];
      }
      reinterpret_cast<float4*>(y)[i] = v;
    }
  } else {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      int c = (int)(i % C);
      float v = __ldg(x + i);
      y[i] = REV ? v * sm[C + c] - sm[c] : (v + sm[c]) * sm[C + c];
    }
  }
}
__global__ void actnorm_logdet_kernel(const float* __restrict__ logs, float* out, int C) {
  __shared__ float red[32];
  float s = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) s += 3.f * logs[c];
  s = block_sum(s, red);
  if (threadIdx.x == 0) *out = s / (float)C;
}
int actnorm(const float* x, const float* b, const float* logs, float* y, float* logdet_out, int64_t rows, int C, bool rev,
            cudaStream_t st) {
  int64_t n = rows * C;
  if (n > 0) {
    int grid = ew_grid(cdiv(n, 4), 256);
    size_t smem = 2 * (size_t)C * sizeof(float);
    if (rev) actnorm_kernel<true><<<grid, 256, smem, st>>>(x, b, logs, y, n, C);
    else actnorm_kernel<false><<<grid, 256, smem, st>>>(x, b, logs, y, n, C);
    FWN_LAUNCH_CHECK();
  }
  if (logdet_out) {
    actnorm_logdet_kernel<<<1, 256, 0, st>>>(logs, logdet_out, C);
    FWN_LAUNCH_CHECK();
  }
  return 0;
}

// DDI: two-pass (mean, then centred second moment) with double accumulation, like the reference's
// reduce_mean of (x+b)^2 (model.py:65 is evaluated on the already centred x).
__global__ void colsum_kernel(const float* __restrict__ x, const float* __restrict__ shift, double* __restrict__ acc, int64_t rows,
                              int C, bool square) {
  __shared__ double part[256];
  if ((int)blockDim.x >= C) {
    const int tpr = blockDim.x / C;  // threads sharing one channel, striding over rows
    const int c = threadIdx.x % C, sub = threadIdx.x / C;
    double s = 0.0;
    if (sub < tpr) {
      const float sh = shift ? shift[c] : 0.f;
      for (int64_t r = (int64_t)blockIdx.x * tpr + sub; r < rows; r += (int64_t)gridDim.x * tpr) {
        float v = __ldg(x + r * C + c) + sh;
        s += square ? (double)v * v : (double)v;
      }
    }
    part[threadIdx.x] = s;
    __syncthreads();
    if ((int)threadIdx.x < C) {
      double t = 0.0;
      for (int j = 0; j < tpr; ++j) t += part[j * C + threadIdx.x];
      atomicAdd(acc + threadIdx.x, t);
    }
  } else {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float sh = shift ? shift[c] : 0.f;
      double s = 0.0;
      for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
        float v = __ldg(x + r * C + c) + sh;
        s += square ? (double)v * v : (double)v;
      }
      atomicAdd(acc + c, s);
    }
  }
}
__global__ void ddi_finish_kernel(const double* acc, float* b_out, float* logs_out, int64_t rows, int C, int phase) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (phase == 0) {
    b_out[c] = (float)(-acc[c] / (double)rows);
  } else {
    double var = acc[C + c] / (double)rows;
    logs_out[c] = (float)(log(1.0 / (sqrt(var) + 1e-7)) / 3.0);
  }
}
int actnorm_ddi(const float* x, float* b_out, float* logs_out, int64_t rows, int C, double* scratch2C, cudaStream_t st) {
  FWN_CUDA(cudaMemsetAsync(scratch2C, 0, 2 * (size_t)C * sizeof(double), st));
  int grid = (int)std::min<int64_t>((int64_t)num_sms() * 4, std::max<int64_t>(1, rows / 8));
  colsum_kernel<<<grid, 256, 0, st>>>(x, nullptr, scratch2C, rows, C, false);
  FWN_LAUNCH_CHECK();
  ddi_finish_kernel<<<(int)cdiv(C, 128), 128, 0, st>>>(scratch2C, b_out, logs_out, rows, C, 0);
  FWN_LAUNCH_CHECK();
  colsum_kernel<<<grid, 256, 0, st>>>(x, b_out, scratch2C + C, rows, C, true);
  FWN_LAUNCH_CHECK();
  ddi_finish_kernel<<<(int)cdiv(C, 128), 128, 0, st>>>(scratch2C, b_out, logs_out, rows, C, 1);
  FWN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- affine coupling (elementwise part)
// x [rows,C], net [rows,NC] with NC = C (affine: log_s | t) or C/2 (additive).
template <bool REV, bool AFFINE>
__global__ void affine_kernel(const float* __restrict__ x, const float* __restrict__ net, float* __restrict__ y, double* __restrict__ acc,
                              int64_t rows, int C) {
  __shared__ double red[32];
  const int h = C / 2;
  const int64_t n = rows * C;
  double ls = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t row = i / C;
    int c = (int)(i - row * C);
    float v = __ldg(x + i);
    if (c >= h) {
      if (AFFINE) {
        float log_s = __ldg(net + row * C + (c - h));
        float t = __ldg(net + row * C + c);
        if (REV) v = v * expf(log_s) + t;
        else { v = (v - t) * expf(-log_s); ls += (double)log_s; }
      } else {
        float t = __ldg(net + row * h + (c - h));
        v = REV ? v - t : v + t;
      }
    }
    y[i] = v;
  }
  if (!REV && AFFINE && acc) {
    ls = block_sum(ls, red);
    if (threadIdx.x == 0) atomicAdd(acc, ls);
  }
}
// 128-bit variant of the affine coupling.  MODE 0: (C/2) % 4 == 0 (one float4 of the transformed half + the matching
// pass-through float4 per thread); MODE 2: C == 2 (two rows per float4); MODE 4: C == 4 (one row per float4); MODE 8: C == 8 (a row = two
// float4: consecutive lanes take consecutive float4 of x / net / y -- fully coalesced -- and the odd lane, which owns the transformed half,
// gets log_s from its even neighbour's net float4 by shuffle; the generic MODE 0 mapping makes every lane stride by 32 bytes there).
template <bool REV, int MODE>
__global__ void affine_vec_kernel(const float4* __restrict__ x, const float4* __restrict__ net, float4* __restrict__ y, double* __restrict__ acc,
                                  int64_t nwork, int C) {
  __shared__ double red[32];
  float ls = 0.f;
  auto tr = [&](float v, float log_s, float t) -> float {
    if (REV) return v * expf(log_s) + t;
    ls += log_s;
    return (v - t) * expf(-log_s);
  };
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nwork; i += (int64_t)gridDim.x * blockDim.x) {
    if (MODE == 2) {         // x = (a0 b0 a1 b1), net = (ls0 t0 ls1 t1)
      const float4 v = __ldg(x + i), w = __ldg(net + i);
      y[i] = make_float4(v.x, tr(v.y, w.x, w.y), v.z, tr(v.w, w.z, w.w));
    } else if (MODE == 4) {  // x = (a0 a1 b0 b1), net = (ls0 ls1 t0 t1)
      const float4 v = __ldg(x + i), w = __ldg(net + i);
      y[i] = make_float4(v.x, v.y, tr(v.z, w.x, w.z), tr(v.w, w.y, w.w));
    } else if (MODE == 8) {  // float4 2r = pass-through half of row r, 2r+1 = transformed half; net float4 2r = log_s, 2r+1 = t
      // (the grid-stride keeps whole warps inside or outside the loop: nwork and the stride are multiples of 32)
      float4 v = __ldg(x + i);
      const float4 w = __ldg(net + i);
      float4 l;
      l.x = __shfl_up_sync(0xffffffffu, w.x, 1);
      l.y = __shfl_up_sync(0xffffffffu, w.y, 1);
      l.z = __shfl_up_sync(0xffffffffu, w.z, 1);
      l.w = __shfl_up_sync(0xffffffffu, w.w, 1);
      if (i & 1) v = make_float4(tr(v.x, l.x, w.x), tr(v.y, l.y, w.y), tr(v.z, l.z, w.z), tr(v.w, l.w, w.w));
      y[i] = v;
    } else {
      const int hq = C / 8, q = C / 4;  // float4 per half row / per row
      const int64_t row = i / hq;
      const int c4 = (int)(i - row * hq);
      const float4 a = __ldg(x + row * q + c4), b = __ldg(x + row * q + hq + c4);
      const float4 l = __ldg(net + row * q + c4), t = __ldg(net + row * q + hq + c4);
      y[row * q + c4] = a;
      y[row * q + hq + c4] = make_float4(tr(b.x, l.x, t.x), tr(b.y, l.y, t.y), tr(b.z, l.z, t.z), tr(b.w, l.w, t.w));
    }
  }
  if (!REV && acc) {
    double d = block_sum((double)ls, red);
    if (threadIdx.x == 0) atomicAdd(acc, d);
  }
}
__global__ void affine_logdet_finish(const double* acc, float* out, double denom) { *out = (float)(-(*acc) / denom / 2.0); }

int affine(const float* x, const float* net, float* y, float* logdet_out, int64_t rows, int C, bool affine_, bool rev, double* scratch,
           cudaStream_t st) {
  FWN_CHECK(C % 2 == 0, "affine: C=%d must be even", C);
  int64_t n = rows * C;
  if (n == 0) return 0;
  int grid = ew_grid(n, 256);
  if (scratch) FWN_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double), st));
  int mode = C == 2 ? 2 : (C == 4 ? 4 : ((C / 2) % 4 == 0 ? 0 : -1));
  if (C == 8 && (n / 4) % 32 == 0) mode = 8;   // whole warps only (the shuffle needs both lanes of a row)
  const bool vec = affine_ && mode >= 0 && n % 4 == 0 && aligned16(x, y) && aligned16(net, net);
  if (vec) {
    const int64_t nwork = mode == 0 ? rows * (C / 8) : n / 4;   // float4 of the transformed half (mode 0) or all float4
    const int g2 = ew_grid(nwork, 256);
    const float4 *x4 = (const float4*)x, *n4 = (const float4*)net;
    float4* y4 = (float4*)y;
#define FWN_AFF(R, M) affine_vec_kernel<R, M><<<g2, 256, 0, st>>>(x4, n4, y4, R ? nullptr : scratch, nwork, C)
    if (rev) { if (mode == 2) FWN_AFF(true, 2); else if (mode == 4) FWN_AFF(true, 4); else if (mode == 8) FWN_AFF(true, 8); else FWN_AFF(true, 0); }
    else { if (mode == 2) FWN_AFF(false, 2); else if (mode == 4) FWN_AFF(false, 4); else if (mode == 8) FWN_AFF(false, 8); else FWN_AFF(false, 0); }
#undef FWN_AFF
  } else if (affine_) {
    if (rev) affine_kernel<true, true><<<grid, 256, 0, st>>>(x, net, y, nullptr, rows, C);
    else affine_kernel<false, true><<<grid, 256, 0, st>>>(x, net, y, scratch, rows, C);
  } else {
    if (rev) affine_kernel<true, false><<<grid, 256, 0, st>>>(x, net, y, nullptr, rows, C);
    else affine_kernel<false, false><<<grid, 256, 0, st>>>(x, net, y, nullptr, rows, C);
  }
  FWN_LAUNCH_CHECK();
  if (!rev && affine_ && logdet_out) {
    affine_logdet_finish<<<1, 1, 0, st>>>(scratch, logdet_out, (double)rows * (C / 2));
    FWN_LAUNCH_CHECK();
  }
  return 0;
}

// ---------------------------------------------------------------- small elementwise helpers
__global__ void gated_kernel(const float* __restrict__ f, const float* __restrict__ g, float* __restrict__ y, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = tanhf(__ldg(f + i)) * (1.f / (1.f + expf(-__ldg(g + i))));
}
__global__ void residual_kernel(const float* __restrict__ x, const float* __restrict__ r, float* __restrict__ y, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = (__ldg(x + i) + __ldg(r + i)) * 0.70710678118654752440f;
}
__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, int64_t n, bool relu) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = __ldg(a + i) + __ldg(b + i);
    y[i] = relu ? fmaxf(v, 0.f) : v;
  }
}
int gated_activation(const float* f, const float* g, float* y, int64_t n, cudaStream_t st) {
  if (n == 0) return 0;
  gated_kernel<<<ew_grid(n, 256), 256, 0, st>>>(f, g, y, n);
  FWN_LAUNCH_CHECK();
  return 0;
}
int residual_scale(const float* x, const float* r, float* y, int64_t n, cudaStream_t st) {
  if (n == 0) return 0;
  residual_kernel<<<ew_grid(n, 256), 256, 0, st>>>(x, r, y, n);
  FWN_LAUNCH_CHECK();
  return 0;
}
int add(const float* a, const float* b, float* y, int64_t n, bool relu, cudaStream_t st) {
  if (n == 0) return 0;
  add_kernel<<<ew_grid(n, 256), 256, 0, st>>>(a, b, y, n, relu);
  FWN_LAUNCH_CHECK();
  return 0;
}

// log_p = mean(0.5 (-log 2pi - z^2))
__global__ void sumsq_kernel(const float* __restrict__ z, double* __restrict__ acc, int64_t n) {
  __shared__ double red[32];
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = __ldg(z + i);
    s += (double)v * v;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(acc, s);
}
__global__ void logp_finish(const double* acc, float* out, double n) {
  *out = (float)(0.5 * (-1.8378770664093454835606594728112 - (*acc) / n));
}
int sumsq(const float* z, double* acc, int64_t n, cudaStream_t st) {
  if (n == 0) return 0;
  sumsq_kernel<<<ew_grid(n, 256), 256, 0, st>>>(z, acc, n);
  FWN_LAUNCH_CHECK();
  return 0;
}
int log_p(const float* z, float* out, int64_t n, double* scratch, cudaStream_t st) {
  FWN_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double), st));
  if (sumsq(z, scratch, n, st)) return 1;
  logp_finish<<<1, 1, 0, st>>>(scratch, out, (double)n);
  FWN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- upsampler (transposed conv stage)
// out[b,i,m] = lrelu( bias + sum_{kh in {r, r+s}} sum_{kw} in[b, j(kh), m+kw-1] * w[kh,kw] ),
// r = (i + s/2) mod s, j(r) = (i + s/2) / s, j(r+s) = j(r) - 1.  6 MAC per output, write-bound.
// SPLIT: write the two mel halves to separate [B*T, mels/2] planes (the fused path's cond layout).
template <typename TOut, bool SPLIT>
__global__ void upsample_kernel(const float* __restrict__ in, const float* __restrict__ w /*[2s,3] weight-normed*/, const float* __restrict__ bias_p,
                                TOut* __restrict__ out0, TOut* __restrict__ out1, int B, int Tm, int mels, int s) {
  extern __shared__ float sw[];  // [2s*3]
  for (int i = threadIdx.x; i < 2 * s * 3; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int To = Tm * s;
  const int64_t n = (int64_t)B * To * mels;
  const int half = mels / 2;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    int m = (int)(idx % mels);
    int64_t bt = idx / mels;
    int i = (int)(bt % To);
    int b = (int)(bt / To);
    int q = i + s / 2;
    int r = q % s, j0 = q / s;
    float acc = __ldg(bias_p);
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      int j = j0 - a, kh = r + a * s;
      if (j < 0 || j >= Tm) continue;
      const float* row = in + ((int64_t)b * Tm + j) * mels;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        int mm = m + 1 - kw;  // in[m'] with m' + kw - 1 = m
        if (mm >= 0 && mm < mels) acc = fmaf(__ldg(row + mm), sw[kh * 3 + kw], acc);
      }
    }
    acc = fmaxf(acc, 0.4f * acc);
    if (SPLIT) {
      if (m < half) out0[bt * half + m] = from_f<TOut>(acc);
      else out1[bt * half + (m - half)] = from_f<TOut>(acc);
    } else {
      out0[idx] = from_f<TOut>(acc);
    }
  }
}
// Warp-cooperative upsampler (mels % 16 == 0): one warp owns one input-frame pair (j0-1, j0) of one utterance and
// produces every output that reads it -- the s time steps i = j0*s - s/2 .. j0*s + s/2 - 1 -- for all mel bins (or one
// mel half when SPLIT).  Outputs are enumerated as 8-mel chunks in memory order, so the warp's stores are contiguous
// 16-byte (bf16) / 32-byte (fp32) pieces: each input is read once from global memory, each output written once.
template <typename TOut, bool SPLIT>
__global__ void upsample_warp_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias_p,
                                     TOut* __restrict__ out0, TOut* __restrict__ out1, int B, int Tm, int mels, int s) {
  extern __shared__ float sm[];
  float* sw = sm;                                   // [2s*3]
  const int wpb = blockDim.x >> 5;
  const int rowlen = mels + 2;
  float* sin = sm + 2 * s * 3 + (threadIdx.x >> 5) * 2 * rowlen;   // per warp: rows j0-1 and j0, with a zero halo column each side
  for (int i = threadIdx.x; i < 2 * s * 3; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int To = Tm * s, half = mels / 2;
  const int planes = SPLIT ? 2 : 1;
  const int gpr = (SPLIT ? half : mels) / 8;        // 8-mel chunks per output row (of one plane)
  const float bias = __ldg(bias_p);
  const int64_t ntask = (int64_t)B * (Tm + 1);
  for (int64_t task = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); task < ntask; task += (int64_t)gridDim.x * wpb) {
    const int b = (int)(task / (Tm + 1)), j0 = (int)(task - (int64_t)b * (Tm + 1));
    __syncwarp();
    for (int i = lane; i < 2 * rowlen; i += 32) {
      const int rr = i / rowlen, mm = i - rr * rowlen - 1;   // rr = 0 -> frame j0-1, 1 -> frame j0
      const int j = j0 - 1 + rr;
      sin[i] = (j >= 0 && j < Tm && mm >= 0 && mm < mels) ? __ldg(in + ((int64_t)b * Tm + j) * mels + mm) : 0.f;
    }
    __syncwarp();
    for (int pl = 0; pl < planes; ++pl) {
      for (int k = lane; k < s * gpr; k += 32) {
        const int r = k / gpr, gi = k - r * gpr;
        const int i = j0 * s + r - s / 2;            // output time step; its taps are kh = r (frame j0) and r + s (frame j0-1)
        if (i < 0 || i >= To) continue;
        const int m0 = pl * half * (SPLIT ? 1 : 0) + gi * 8;
        const float* x1 = sin + rowlen + m0;        // frame j0, element m0-1 at x1[0]
        const float* x0 = sin + m0;                 // frame j0-1
        const float a0 = sw[r * 3], a1 = sw[r * 3 + 1], a2 = sw[r * 3 + 2];
        const float c0 = sw[(r + s) * 3], c1 = sw[(r + s) * 3 + 1], c2 = sw[(r + s) * 3 + 2];
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {  // out[m] = sum_kw in[m + 1 - kw] * w[kh][kw]
          float v = bias;
          v = fmaf(x1[j + 2], a0, fmaf(x1[j + 1], a1, fmaf(x1[j], a2, v)));
          v = fmaf(x0[j + 2], c0, fmaf(x0[j + 1], c1, fmaf(x0[j], c2, v)));
          acc[j] = fmaxf(v, 0.4f * v);
        }
        const int64_t bt = (int64_t)b * To + i;
        TOut* dst = SPLIT ? (pl == 0 ? out0 : out1) + bt * half + gi * 8 : out0 + bt * mels + gi * 8;
        if constexpr (sizeof(TOut) == 2) {
          *reinterpret_cast<uint4*>(dst) = make_uint4(pack2<TOut>(acc[0], acc[1]), pack2<TOut>(acc[2], acc[3]),
                                                      pack2<TOut>(acc[4], acc[5]), pack2<TOut>(acc[6], acc[7]));
        } else {
          reinterpret_cast<float4*>(dst)[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
          reinterpret_cast<float4*>(dst)[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
        }
      }
    }
  }
}
// Chunk-per-thread upsampler (mels % 8 == 0, and (mels/2) % 8 == 0 when SPLIT): one thread = 8 consecutive mel bins of one output
// time step = one 16-byte (16-bit) or 32-byte (fp32) store; consecutive threads write consecutive chunks of a plane, so every warp
// store instruction covers 512 contiguous bytes.  The 2 x 10 inputs a thread needs come from the (tiny, L1-resident) stage input:
// an input frame is re-used by the s output steps it feeds.  No shared-memory staging, no warp synchronisation: the kernel runs at
// the write bandwidth of its output.
// IT = index type of the flat chunk index: 32-bit whenever the tensor allows it (64-bit divisions cost ~100 instructions each, and
// five of them per 16-byte store made the kernel issue-bound at 0.2 of the HBM rate).
template <typename TOut, bool SPLIT, typename IT>
__global__ void __launch_bounds__(256) upsample_chunk_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias_p,
                                                             TOut* __restrict__ out0, TOut* __restrict__ out1, int B, int Tm, int mels, int s) {
  extern __shared__ float sw[];  // [2s*3]
  for (int i = threadIdx.x; i < 2 * s * 3; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const IT To = (IT)Tm * (IT)s;
  const int half = mels / 2;
  const int pm = SPLIT ? half : mels;            // mel bins per output row of one plane
  const IT gpr = (IT)(pm / 8);                   // 8-mel chunks per row
  const IT rows = (IT)B * To;
  const IT per_plane = rows * gpr;
  const IT total = per_plane * (SPLIT ? 2 : 1);
  const float bias = __ldg(bias_p);
  for (IT idx = (IT)blockIdx.x * (IT)blockDim.x + (IT)threadIdx.x; idx < total; idx += (IT)gridDim.x * (IT)blockDim.x) {
    const int pl = (SPLIT && idx >= per_plane) ? 1 : 0;
    const IT rem = idx - (pl ? per_plane : (IT)0);
    const IT bt = rem / gpr;
    const int gi = (int)(rem - bt * gpr);
    const IT bb = bt / To;
    const int b = (int)bb, i = (int)(bt - bb * To);
    const int q = i + s / 2, j0 = q / s, r = q - j0 * s;           // taps kh = r (frame j0) and r + s (frame j0 - 1)
    const int m0 = pl * half + gi * 8;                             // first mel bin of the chunk
    float x1[10], x0[10];                                          // frames j0 / j0-1, mel bins m0-1 .. m0+8 (zero outside)
    const float* f1 = in + ((int64_t)b * Tm + j0) * mels;
    const float* f0 = f1 - mels;
    const bool v1 = j0 < Tm, v0 = j0 >= 1;
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      const int mm = m0 - 1 + k;
      const bool ok = mm >= 0 && mm < mels;
      x1[k] = (ok && v1) ? __ldg(f1 + mm) : 0.f;
      x0[k] = (ok && v0) ? __ldg(f0 + mm) : 0.f;
    }
    const float a0 = sw[r * 3], a1 = sw[r * 3 + 1], a2 = sw[r * 3 + 2];
    const float c0 = sw[(r + s) * 3], c1 = sw[(r + s) * 3 + 1], c2 = sw[(r + s) * 3 + 2];
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {  // out[m] = sum_kw in[m + 1 - kw] * w[kh][kw]
      float v = bias;
      v = fmaf(x1[j + 2], a0, fmaf(x1[j + 1], a1, fmaf(x1[j], a2, v)));
      v = fmaf(x0[j + 2], c0, fmaf(x0[j + 1], c1, fmaf(x0[j], c2, v)));
      acc[j] = fmaxf(v, 0.4f * v);
    }
    TOut* dst = (SPLIT && pl == 1 ? out1 : out0) + (int64_t)bt * pm + gi * 8;
    if constexpr (sizeof(TOut) == 2) {
      *reinterpret_cast<uint4*>(dst) = make_uint4(pack2<TOut>(acc[0], acc[1]), pack2<TOut>(acc[2], acc[3]),
                                                  pack2<TOut>(acc[4], acc[5]), pack2<TOut>(acc[6], acc[7]));
    } else {
      reinterpret_cast<float4*>(dst)[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      reinterpret_cast<float4*>(dst)[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
  }
}
// Frame-per-thread upsampler: one thread = 8 consecutive mel bins of the s output time steps fed by one pair of input frames
// (j0 - 1, j0): the 2 x 10 inputs are loaded once and every step costs 48 FMA + one 16-byte store, 4x fewer instructions per byte
// than one thread per store (which was issue-bound at 0.2 of the HBM rate).  Lanes walk the chunks of a row first, so a store
// instruction writes runs of pm * 2 contiguous bytes, s rows apart.
template <typename TOut, bool SPLIT>
__global__ void __launch_bounds__(256) upsample_frame_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias_p,
                                                             TOut* __restrict__ out0, TOut* __restrict__ out1, int B, int Tm, int mels, int s) {
  extern __shared__ float sw[];  // [2s*3]
  for (int i = threadIdx.x; i < 2 * s * 3; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int half = mels / 2;
  const int pm = SPLIT ? half : mels;            // mel bins per output row of one plane
  const uint32_t gpr = (uint32_t)(pm / 8);       // 8-mel chunks per row
  const uint32_t per_b = (uint32_t)(Tm + 1) * gpr, per_plane = (uint32_t)B * per_b;
  const uint32_t total = per_plane * (SPLIT ? 2u : 1u);
  const int To = Tm * s;
  const float bias = __ldg(bias_p);
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int pl = (SPLIT && idx >= per_plane) ? 1 : 0;
    const uint32_t rem = idx - (pl ? per_plane : 0u);
    const uint32_t b = rem / per_b, rb = rem - b * per_b;
    const int j0 = (int)(rb / gpr), gi = (int)(rb - (uint32_t)j0 * gpr);
    const int m0 = pl * half + gi * 8;
    float x1[10], x0[10];                          // frames j0 / j0-1, mel bins m0-1 .. m0+8 (zero outside)
    const float* f1 = in + ((int64_t)b * Tm + j0) * mels;
    const float* f0 = f1 - mels;
    const bool v1 = j0 < Tm, v0 = j0 >= 1;
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      const int mm = m0 - 1 + k;
      const bool ok = mm >= 0 && mm < mels;
      x1[k] = (ok && v1) ? __ldg(f1 + mm) : 0.f;
      x0[k] = (ok && v0) ? __ldg(f0 + mm) : 0.f;
    }
    TOut* plane = (SPLIT && pl == 1 ? out1 : out0) + gi * 8;
    const int i0 = j0 * s - s / 2;                 // output steps i0 .. i0 + s - 1 use taps kh = r (frame j0) and r + s (frame j0 - 1)
    for (int r = 0; r < s; ++r) {
      const int i = i0 + r;
      if (i < 0 || i >= To) continue;
      const float a0 = sw[r * 3], a1 = sw[r * 3 + 1], a2 = sw[r * 3 + 2];
      const float c0 = sw[(r + s) * 3], c1 = sw[(r + s) * 3 + 1], c2 = sw[(r + s) * 3 + 2];
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {  // out[m] = sum_kw in[m + 1 - kw] * w[kh][kw]
        float v = bias;
        v = fmaf(x1[j + 2], a0, fmaf(x1[j + 1], a1, fmaf(x1[j], a2, v)));
        v = fmaf(x0[j + 2], c0, fmaf(x0[j + 1], c1, fmaf(x0[j], c2, v)));
        acc[j] = fmaxf(v, 0.4f * v);
      }
      TOut* dst = plane + ((int64_t)b * To + i) * pm;
      if constexpr (sizeof(TOut) == 2) {
        *reinterpret_cast<uint4*>(dst) = make_uint4(pack2<TOut>(acc[0], acc[1]), pack2<TOut>(acc[2], acc[3]),
                                                    pack2<TOut>(acc[4], acc[5]), pack2<TOut>(acc[6], acc[7]));
      } else {
        reinterpret_cast<float4*>(dst)[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        reinterpret_cast<float4*>(dst)[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
      }
    }
  }
}
// weight norm of the [2s,3,1,1] kernel over axes [0,2] => per kw column (convolutional.py:186)
__global__ void upsample_wn_kernel(const float* __restrict__ v, const float* __restrict__ g, float* __restrict__ w, int s) {
  int kw = threadIdx.x;
  if (kw >= 3) return;
  float ss = 0.f;
  for (int kh = 0; kh < 2 * s; ++kh) ss += v[kh * 3 + kw] * v[kh * 3 + kw];
  float sc = rsqrtf(fmaxf(ss, 1e-12f)) * g[0];
  for (int kh = 0; kh < 2 * s; ++kh) w[kh * 3 + kw] = v[kh * 3 + kw] * sc;
}
int upsample_weight_norm(const float* v, const float* g, float* w, int s, cudaStream_t st) {
  upsample_wn_kernel<<<1, 32, 0, st>>>(v, g, w, s);
  FWN_LAUNCH_CHECK();
  return 0;
}
template <typename TOut>
int upsample_stage_t(const float* in, const float* w, const float* bias, TOut* out0, TOut* out1, int B, int Tm, int mels, int s, bool split,
                     cudaStream_t st) {
  int64_t n = (int64_t)B * Tm * s * mels;
  if (n == 0) return 0;
  int grid = ew_grid(n, 256);
  size_t smem = 2 * (size_t)s * 3 * sizeof(float);
  const bool al16 = (reinterpret_cast<uintptr_t>(out0) % 16) == 0 && (!split || reinterpret_cast<uintptr_t>(out1) % 16 == 0);
  static const bool use_warp_kernel = getenv("FWN_UPSAMPLE_WARP") != nullptr;   // the round-1 kernel, kept for A/B timing
  if (!use_warp_kernel && al16 && mels % 8 == 0 && (!split || (mels / 2) % 8 == 0)) {
    const int64_t chunks = n / 8;
    const int g2 = (int)std::min<int64_t>(cdiv(chunks, 256), (int64_t)num_sms() * 32);
    static const bool use_chunk_kernel = getenv("FWN_UPSAMPLE_CHUNK") != nullptr;   // the per-store kernel, kept for A/B timing
    const int64_t frame_threads = (int64_t)B * (Tm + 1) * (mels / 8);
    if (!use_chunk_kernel && s <= 32 && frame_threads + (int64_t)num_sms() * 32 * 256 < (int64_t(1) << 31)) {
      const int g3 = (int)std::min<int64_t>(cdiv(frame_threads, 256), (int64_t)num_sms() * 32);
      if (split) upsample_frame_kernel<TOut, true><<<g3, 256, smem, st>>>(in, w, bias, out0, out1, B, Tm, mels, s);
      else upsample_frame_kernel<TOut, false><<<g3, 256, smem, st>>>(in, w, bias, out0, out1, B, Tm, mels, s);
      FWN_LAUNCH_CHECK();
      return 0;
    }
    const bool small = chunks + (int64_t)g2 * 256 < (int64_t(1) << 31);   // flat chunk index (plus one grid stride) fits 32 bits
    if (small) {
      if (split) upsample_chunk_kernel<TOut, true, uint32_t><<<g2, 256, smem, st>>>(in, w, bias, out0, out1, B, Tm, mels, s);
      else upsample_chunk_kernel<TOut, false, uint32_t><<<g2, 256, smem, st>>>(in, w, bias, out0, out1, B, Tm, mels, s);
    } else {
      if (split) upsample_chunk_kernel<TOut, true, int64_t><<<g2, 256, smem, st>>>(in, w, bias, out0, out1, B, Tm, mels, s);
      else upsample_chunk_kernel<TOut, false, int64_t><<<g2, 256, smem, st>>>(in, w, bias, out0, out1, B, Tm, mels, s);
    }
    FWN_LAUNCH_CHECK();
    return 0;
  }
  if (mels % 16 == 0 && al16) {
    const int wpb = 8;
    const size_t sm2 = smem + (size_t)wpb * 2 * (mels + 2) * sizeof(float);
    const int64_t ntask = (int64_t)B * (Tm + 1);
    const int g2 = (int)std::min<int64_t>(cdiv(ntask, wpb), (int64_t)num_sms() * 16);
    if (split) upsample_warp_kernel<TOut, true><<<g2, wpb * 32, sm2, st>>>(in, w, bias, out0, out1, B, Tm, mels, s);
    else upsample_warp_kernel<TOut, false><<<g2, wpb * 32, sm2, st>>>(in, w, bias, out0, out1, B, Tm, mels, s);
    FWN_LAUNCH_CHECK();
    return 0;
  }
  if (split) upsample_kernel<TOut, true><<<grid, 256, smem, st>>>(in, w, bias, out0, out1, B, Tm, mels, s);
  else upsample_kernel<TOut, false><<<grid, 256, smem, st>>>(in, w, bias, out0, out1, B, Tm, mels, s);
  FWN_LAUNCH_CHECK();
  return 0;
}
int upsample_stage(const float* in, const float* w, const float* bias, void* out0, void* out1, int B, int Tm, int mels, int s, bool split,
                   int out_kind, cudaStream_t st) {
  if (out_kind == 1) return upsample_stage_t<__nv_bfloat16>(in, w, bias, (__nv_bfloat16*)out0, (__nv_bfloat16*)out1, B, Tm, mels, s, split, st);
  if (out_kind == 2) return upsample_stage_t<__half>(in, w, bias, (__half*)out0, (__half*)out1, B, Tm, mels, s, split, st);
  return upsample_stage_t<float>(in, w, bias, (float*)out0, (float*)out1, B, Tm, mels, s, split, st);
}

}  // namespace fwn

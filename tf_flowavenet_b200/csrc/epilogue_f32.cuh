// fp32 epilogues of the implicit GEMM, shared by the two engines of the fp32 parity mode (CUDA-core engine in conv_simt.cu and the
// 3-way-split tcgen05 engine in gemm_tc3.cu).  Called with 4 consecutive accumulator columns (col % 4 == 0) of one row.
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace fwn {

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.f / (1.f + expf(-x)); }

// ---------------------------------------------------------------- epilogues (fp32 activations)
// Called with 4 consecutive columns (col % 4 == 0) of one row.
template <int EPI>
struct Epilogue {
  // LINEAR accumulates in place (in0 == out0 is allowed), so the compiler must keep every load of in0 behind the preceding store:
  // callers that walk many columns per row fetch the addends of a whole batch first (prefetch), then call apply with `pre`.
  __device__ static __forceinline__ void prefetch(const GemmArgs& g, int64_t row, int col, float pre[4]) {
    const EpiArgs& e = g.e;
    pre[0] = pre[1] = pre[2] = pre[3] = 0.f;
    if ((EPI != EPI_LINEAR && EPI != EPI_GATE) || !e.in0) return;
    const int64_t ld = EPI == EPI_GATE ? 2 * e.F : e.ld;
    const float* a0 = reinterpret_cast<const float*>(e.in0) + row * ld + col;
    if (col + 3 < g.N && (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(a0) & 15) == 0) {
      const float4 v = *reinterpret_cast<const float4*>(a0);
      pre[0] = v.x; pre[1] = v.y; pre[2] = v.z; pre[3] = v.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) if (col + j < g.N) pre[j] = a0[j];
    }
  }
  __device__ static __forceinline__ void apply(const GemmArgs& g, int64_t row, int t, int col, const float acc[4], double& ls_sum,
                                               const float* pre = nullptr) {
    const EpiArgs& e = g.e;
    if (EPI == EPI_PLAIN) {
      float* y = reinterpret_cast<float*>(e.out0) + row * e.ld;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = col + j;
        v[j] = acc[j];
        if (n < g.N) {
          if (e.colscale) v[j] *= __ldg(e.colscale + n);
          v[j] += __ldg(e.bias + n);
          if (e.relu) v[j] = fmaxf(v[j], 0.f);
        }
      }
      if (col + 3 < g.N && (e.ld & 3) == 0 && (reinterpret_cast<uintptr_t>(y + col) & 15) == 0) {
        *reinterpret_cast<float4*>(y + col) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (col + j < g.N) y[col + j] = v[j];
      }
    } else if (EPI == EPI_GATE) {
      // columns (2c, 2c+1) = (filter_c, gate_c)  -> o[row, c] = tanh(f) * sigmoid(g)   (modules.py:124)
      float* o = reinterpret_cast<float*>(e.out0) + row * e.F;
      if (col + 3 < g.N) {
        float f0 = acc[0] + __ldg(e.bias + col), g0 = acc[1] + __ldg(e.bias + col + 1);
        float f1 = acc[2] + __ldg(e.bias + col + 2), g1 = acc[3] + __ldg(e.bias + col + 3);
        if (pre) {    // training, deep blocks: the conditioning projection was computed ahead (it does not depend on the flow state)
          f0 += pre[0]; g0 += pre[1]; f1 += pre[2]; g1 += pre[3];
        } else if (e.in0) {   // (in0 may alias out1: batch callers prefetch, see above)
          const float4 y = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(e.in0) + row * (2 * e.F) + col);
          f0 += y.x; g0 += y.y; f1 += y.z; g1 += y.w;
        }
        float2 v = make_float2(tanhf(f0) * sigmoidf_acc(g0), tanhf(f1) * sigmoidf_acc(g1));
        *reinterpret_cast<float2*>(o + col / 2) = v;
        if (e.out1)  // training: keep the pre-activations for the backward pass
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out1) + row * (2 * e.F) + col) = make_float4(f0, g0, f1, g1);
      }
    } else if (EPI == EPI_RES_SKIP) {
      // [0,F): h_out = (h_in + res) * sqrt(.5) (modules.py:128); skip columns: skip (+ running sum) (modules.py:127,176)
      const int F = e.F;
      if (col + 3 < g.N) {
        float4 b4 = __ldg(reinterpret_cast<const float4*>(e.bias + col));
        float v[4] = {acc[0] + b4.x, acc[1] + b4.y, acc[2] + b4.z, acc[3] + b4.w};
        if (e.has_res && col < F) {
          const float4 h = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(e.in0) + row * F + col));
          const float s = 0.70710678118654752440f;
          float4 r = make_float4((h.x + v[0]) * s, (h.y + v[1]) * s, (h.z + v[2]) * s, (h.w + v[3]) * s);
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out0) + row * F + col) = r;
        } else {
          const int c = e.has_res ? col - F : col;
          if (e.in1) {
            const float4 s4 = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(e.in1) + row * F + c));
            v[0] += s4.x; v[1] += s4.y; v[2] += s4.z; v[3] += s4.w;
          }
          if (e.relu) {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out1) + row * F + c) = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
    } else if (EPI == EPI_AFFINE) {
      // columns (2q, 2q+1) = (log_s, t) of transformed element q; also applies ActNorm to both halves.
      float* xr = e.X + row * e.Cx;
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const int q = col / 2 + p;
        if (q >= e.nq) continue;
        const float log_s = acc[2 * p] + __ldg(e.bias + col + 2 * p);
        const float tt = acc[2 * p + 1] + __ldg(e.bias + col + 2 * p + 1);
        const int oa = __ldg(e.a_off + q), ob = __ldg(e.b_off + q);
        if (e.out1) *reinterpret_cast<float2*>(reinterpret_cast<float*>(e.out1) + row * e.ld + 2 * q) = make_float2(log_s, tt);
        float xa = xr[oa], xb = xr[ob];
        if (!e.reverse) {  // Flow.forward: ActNorm, then out_b = (in_b - t) exp(-log_s)   (model.py:188-189,134)
          xa = (xa + __ldg(e.an_b + oa)) * __ldg(e.an_s + oa);
          xb = (xb + __ldg(e.an_b + ob)) * __ldg(e.an_s + ob);
          xb = (xb - tt) * expf(-log_s);
          ls_sum += (double)log_s;
        } else {           // Flow.reverse: in_b = out_b exp(log_s) + t, then ActNorm.reverse   (model.py:156,201)
          xb = xb * expf(log_s) + tt;
          xa = xa * __ldg(e.an_s + oa) - __ldg(e.an_b + oa);
          xb = xb * __ldg(e.an_s + ob) - __ldg(e.an_b + ob);
        }
        xr[oa] = xa;
        xr[ob] = xb;
      }
    } else if (EPI == EPI_LINEAR) {
      float* y = reinterpret_cast<float*>(e.out0) + row * e.ld;
      const float* a0 = (e.in0 && !pre) ? reinterpret_cast<const float*>(e.in0) + row * e.ld : nullptr;
      const float* mk = e.in1 ? reinterpret_cast<const float*>(e.in1) + row * e.ld : nullptr;
      float v[4];
      const bool vec = col + 3 < g.N && (e.ld & 3) == 0 && (reinterpret_cast<uintptr_t>(y + col) & 15) == 0;
      float mv[4] = {1.f, 1.f, 1.f, 1.f};
      if (mk) {   // one 16-byte load per row instead of four 4-byte ones: row-per-thread accesses cost a cache line each
        if (vec && (reinterpret_cast<uintptr_t>(mk + col) & 15) == 0) {
          const float4 t4 = __ldg(reinterpret_cast<const float4*>(mk + col));
          mv[0] = t4.x; mv[1] = t4.y; mv[2] = t4.z; mv[3] = t4.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) if (col + j < g.N) mv[j] = __ldg(mk + col + j);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = col + j;
        v[j] = acc[j];
        if (n < g.N) {
          if (e.bias) v[j] += __ldg(e.bias + n);
          if (pre) v[j] += pre[j];
          else if (a0) v[j] += a0[n];
          v[j] *= e.alpha;
          if (!(mv[j] > 0.f)) v[j] = 0.f;
        }
      }
      if (vec) {
        *reinterpret_cast<float4*>(y + col) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (col + j < g.N) y[col + j] = v[j];
      }
    } else if (EPI == EPI_GATE_BWD) {
      // o = tanh(f) sigmoid(g):  df = do sigmoid(g) (1 - tanh^2 f),  dg = do tanh(f) sigmoid(g) (1 - sigmoid(g))
      if (col + 3 < g.N) {
        const float* fg = reinterpret_cast<const float*>(e.in0) + row * (2 * e.F) + 2 * col;
        float* d = reinterpret_cast<float*>(e.out0) + row * (2 * e.F) + 2 * col;
        const float4 p0 = __ldg(reinterpret_cast<const float4*>(fg)), p1 = __ldg(reinterpret_cast<const float4*>(fg + 4));
        const float pf[4] = {p0.x, p0.z, p1.x, p1.z}, pg[4] = {p0.y, p0.w, p1.y, p1.w};
        float r[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float th = tanhf(pf[j]), sg = sigmoidf_acc(pg[j]);
          r[2 * j] = acc[j] * sg * (1.f - th * th);
          r[2 * j + 1] = acc[j] * th * sg * (1.f - sg);
        }
        *reinterpret_cast<float4*>(d) = make_float4(r[0], r[1], r[2], r[3]);
        *reinterpret_cast<float4*>(d + 4) = make_float4(r[4], r[5], r[6], r[7]);
      }
    }
  }
};

}  // namespace fwn

// fp32 CUDA-core engine: implicit-GEMM dilated convolution with fused flow epilogues.
//
// This is the FP32 PARITY MODE of the path (BASELINE north_star: z / log-det within 1e-4 relative of the
// reference graph in fp32).  A 10-bit-mantissa tensor-core pass cannot hold that bound through 48 flows,
// so the parity mode runs fp32 FMAs; the throughput mode is the tcgen05 engine in gemm_tc.cu.
//
// Kernel: 128x128x16 tile, 256 threads, 8x8 register micro-tile, double-buffered shared memory,
// 128-bit global and shared accesses.  A is gathered per segment with a time shift and zero fill
// (= tf.pad of modules.py:27); see GemmArgs in kernels.h.
#include "common.cuh"
#include "kernels.h"

namespace fwn {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256;

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.f / (1.f + expf(-x)); }

// ---------------------------------------------------------------- epilogues (fp32 activations)
// Called with 4 consecutive columns (col % 4 == 0) of one row.
template <int EPI>
struct Epilogue {
  __device__ static __forceinline__ void apply(const GemmArgs& g, int64_t row, int t, int col, const float acc[4], double& ls_sum) {
    const EpiArgs& e = g.e;
    if (EPI == EPI_PLAIN) {
      float* y = reinterpret_cast<float*>(e.out0) + row * e.ld;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int n = col + j;
        if (n < g.N) {
          float v = acc[j];
          if (e.colscale) v *= __ldg(e.colscale + n);
          v += __ldg(e.bias + n);
          y[n] = e.relu ? fmaxf(v, 0.f) : v;
        }
      }
    } else if (EPI == EPI_GATE) {
      // columns (2c, 2c+1) = (filter_c, gate_c)  -> o[row, c] = tanh(f) * sigmoid(g)   (modules.py:124)
      float* o = reinterpret_cast<float*>(e.out0) + row * e.F;
      if (col + 3 < g.N) {
        float f0 = acc[0] + __ldg(e.bias + col), g0 = acc[1] + __ldg(e.bias + col + 1);
        float f1 = acc[2] + __ldg(e.bias + col + 2), g1 = acc[3] + __ldg(e.bias + col + 3);
        float2 v = make_float2(tanhf(f0) * sigmoidf_acc(g0), tanhf(f1) * sigmoidf_acc(g1));
        *reinterpret_cast<float2*>(o + col / 2) = v;
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
    }
  }
};

// ---------------------------------------------------------------- main kernel
template <int EPI>
__global__ void __launch_bounds__(NT, 2) simt_gemm_kernel(const GemmArgs g) {
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  __shared__ double red[32];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int tiles_per_utt = (g.Ti + BM - 1) / BM;
  const int mt = blockIdx.x;               // m tile: (utterance, time tile)
  const int ub = mt / tiles_per_utt;
  const int t0 = (mt - ub * tiles_per_utt) * BM;
  const int n0 = blockIdx.y * BN;

  // A-load mapping: one row, 8 consecutive k
  const int a_row = tid & 127, a_k = (tid >> 7) * 8;
  // B-load mapping: one k, 8 consecutive n
  const int b_k = tid >> 4, b_n = (tid & 15) * 8;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  // flattened chunk list over segments
  int nchunk = 0;
  for (int s = 0; s < g.nseg; ++s) nchunk += (g.seg[s].K + BK - 1) / BK;

  float ra[8], rb[8];
  auto load_chunk = [&](int chunk) {
    int s = 0, c = chunk;
    while (true) {
      int nc = (g.seg[s].K + BK - 1) / BK;
      if (c < nc) break;
      c -= nc;
      ++s;
    }
    const Seg& sg = g.seg[s];
    const int k0 = c * BK;
    // ---- A
    {
      const int t = t0 + a_row + sg.shift;
      const bool row_ok = (t0 + a_row < g.Ti) && t >= 0 && t < g.Ti;
      const float* ap = reinterpret_cast<const float*>(sg.A) + ((int64_t)ub * g.Ti + t) * sg.lda + k0 + a_k;
      const bool vec = row_ok && (k0 + a_k + 7 < sg.K) && ((sg.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(ap) & 15) == 0);
      if (vec) {
        float4 v0 = __ldg(reinterpret_cast<const float4*>(ap));
        float4 v1 = __ldg(reinterpret_cast<const float4*>(ap) + 1);
        ra[0] = v0.x; ra[1] = v0.y; ra[2] = v0.z; ra[3] = v0.w;
        ra[4] = v1.x; ra[5] = v1.y; ra[6] = v1.z; ra[7] = v1.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) ra[j] = (row_ok && (k0 + a_k + j < sg.K)) ? __ldg(ap + j) : 0.f;
      }
    }
    // ---- B (weights)
    {
      const int k = k0 + b_k;
      const bool k_ok = k < sg.K;
      const float* wp = reinterpret_cast<const float*>(g.W) + (int64_t)(sg.koff + k) * g.ldw + n0 + b_n;
      const bool vec = k_ok && (n0 + b_n + 7 < g.N) && ((g.ldw & 3) == 0) && ((reinterpret_cast<uintptr_t>(wp) & 15) == 0);
      if (vec) {
        float4 v0 = __ldg(reinterpret_cast<const float4*>(wp));
        float4 v1 = __ldg(reinterpret_cast<const float4*>(wp) + 1);
        rb[0] = v0.x; rb[1] = v0.y; rb[2] = v0.z; rb[3] = v0.w;
        rb[4] = v1.x; rb[5] = v1.y; rb[6] = v1.z; rb[7] = v1.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) rb[j] = (k_ok && (n0 + b_n + j < g.N)) ? __ldg(wp + j) : 0.f;
      }
    }
  };
  auto store_chunk = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 8; ++j) As[buf][a_k + j][a_row] = ra[j];
    *reinterpret_cast<float4*>(&Bs[buf][b_k][b_n]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
    *reinterpret_cast<float4*>(&Bs[buf][b_k][b_n + 4]) = make_float4(rb[4], rb[5], rb[6], rb[7]);
  };

  load_chunk(0);
  store_chunk(0);
  __syncthreads();
  for (int ch = 0; ch < nchunk; ++ch) {
    const int buf = ch & 1;
    if (ch + 1 < nchunk) load_chunk(ch + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (ch + 1 < nchunk) store_chunk(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue
  double ls_sum = 0.0;
#pragma unroll
  for (int ih = 0; ih < 2; ++ih)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = t0 + ih * 64 + ty * 4 + i;
      if (t >= g.Ti) continue;
      const int64_t row = (int64_t)ub * g.Ti + t;
#pragma unroll
      for (int jh = 0; jh < 2; ++jh) {
        const int col = n0 + jh * 64 + tx * 4;
        if (col >= g.N) continue;
        Epilogue<EPI>::apply(g, row, t, col, &acc[ih * 4 + i][jh * 4], ls_sum);
      }
    }
  if (EPI == EPI_AFFINE) {
    if (!g.e.reverse && g.e.logdet_acc) {
      ls_sum = block_sum(ls_sum, red);
      if (tid == 0) atomicAdd(g.e.logdet_acc, ls_sum);
    }
  }
}

int simt_gemm(const GemmArgs& a, EpiKind kind, cudaStream_t st) {
  if (a.B <= 0 || a.Ti <= 0 || a.N <= 0) return 0;
  FWN_CHECK(a.nseg >= 1 && a.nseg <= 4, "simt_gemm: bad segment count %d", a.nseg);
  if (kind == EPI_GATE || kind == EPI_RES_SKIP) FWN_CHECK(a.N % 4 == 0 && a.e.F % 4 == 0, "simt_gemm: N and F must be multiples of 4");
  const int tiles_per_utt = (a.Ti + BM - 1) / BM;
  dim3 grid((unsigned)(a.B * tiles_per_utt), (unsigned)((a.N + BN - 1) / BN));
  switch (kind) {
    case EPI_PLAIN: simt_gemm_kernel<EPI_PLAIN><<<grid, NT, 0, st>>>(a); break;
    case EPI_GATE: simt_gemm_kernel<EPI_GATE><<<grid, NT, 0, st>>>(a); break;
    case EPI_RES_SKIP: simt_gemm_kernel<EPI_RES_SKIP><<<grid, NT, 0, st>>>(a); break;
    case EPI_AFFINE: simt_gemm_kernel<EPI_AFFINE><<<grid, NT, 0, st>>>(a); break;
  }
  FWN_LAUNCH_CHECK();
  return 0;
}

// scale[o] = g[o] * rsqrt(max(sum_{k,i} v[k,i,o]^2, 1e-12))  -- weight norm as an epilogue column scale
// (l2_normalize over axes [0,1] is per output channel, convolutional.py:80)
__global__ void wn_scale_kernel(const float* __restrict__ v, const float* __restrict__ g, float* __restrict__ scale, int K, int Cout) {
  int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= Cout) return;
  float ss = 0.f;
  for (int k = 0; k < K; ++k) {
    float w = __ldg(v + (int64_t)k * Cout + o);
    ss = fmaf(w, w, ss);
  }
  scale[o] = __ldg(g + o) * rsqrtf(fmaxf(ss, 1e-12f));
}
int weight_norm_scale(const float* v, const float* g, float* scale, int K, int Cout, cudaStream_t st) {
  wn_scale_kernel<<<(int)cdiv(Cout, 128), 128, 0, st>>>(v, g, scale, K, Cout);
  FWN_LAUNCH_CHECK();
  return 0;
}
__global__ void exp3_kernel(const float* __restrict__ s, float* __restrict__ out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = expf(3.f * s[i]);
}
int exp3(const float* s, float* out, int n, cudaStream_t st) {
  exp3_kernel<<<(int)cdiv(n, 128), 128, 0, st>>>(s, out, n);
  FWN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- front conv on the flow variable
// h0[row, ch] = relu(bias[ch] + sum_tap sum_q a(row + shift_tap, q) * W[tap][q][ch]),
// a(r, q) = ActNorm(X[r, a_off[q]]) (forward) or X[r, a_off[q]] (reverse); zero outside the utterance.
// K = 3*Cx/2 is 3..384: 384 MAC per audio sample in every block, <0.3% of the pass -> CUDA cores.
constexpr int FR = 32;  // rows per CTA
template <typename TOut>
__global__ void __launch_bounds__(256) front_kernel(const FrontArgs a) {
  extern __shared__ float xs[];  // [3][nq][FR]  (one aligned copy per tap so reads are 128-bit)
  const int nq = a.nq;
  const int tiles_per_utt = (a.Ti + FR - 1) / FR;
  const int ub = blockIdx.x / tiles_per_utt;
  const int t0 = (blockIdx.x - ub * tiles_per_utt) * FR;
  for (int i = threadIdx.x; i < 3 * nq * FR; i += blockDim.x) {
    const int r = i % FR;
    const int q = (i / FR) % nq;
    const int tap = i / (FR * nq);
    const int t = t0 + r + a.shift[tap];
    float v = 0.f;
    if (t >= 0 && t < a.Ti) {
      const int o = __ldg(a.a_off + q);
      v = __ldg(a.X + ((int64_t)ub * a.Ti + t) * a.Cx + o);
      if (a.an_b) v = (v + __ldg(a.an_b + o)) * __ldg(a.an_s + o);
    }
    xs[i] = v;
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < a.F; ch += blockDim.x) {
    float acc[FR];
    const float b = __ldg(a.bias + ch);
#pragma unroll
    for (int r = 0; r < FR; ++r) acc[r] = b;
    for (int kq = 0; kq < 3 * nq; ++kq) {
      const float w = __ldg(a.W + (int64_t)kq * a.F + ch);
      const float4* xp = reinterpret_cast<const float4*>(xs + kq * FR);
#pragma unroll
      for (int r4 = 0; r4 < FR / 4; ++r4) {
        const float4 x4 = xp[r4];
        acc[r4 * 4 + 0] = fmaf(x4.x, w, acc[r4 * 4 + 0]);
        acc[r4 * 4 + 1] = fmaf(x4.y, w, acc[r4 * 4 + 1]);
        acc[r4 * 4 + 2] = fmaf(x4.z, w, acc[r4 * 4 + 2]);
        acc[r4 * 4 + 3] = fmaf(x4.w, w, acc[r4 * 4 + 3]);
      }
    }
    TOut* H = reinterpret_cast<TOut*>(a.H);
#pragma unroll
    for (int r = 0; r < FR; ++r) {
      const int t = t0 + r;
      if (t < a.Ti) H[((int64_t)ub * a.Ti + t) * a.F + ch] = from_f<TOut>(fmaxf(acc[r], 0.f));
    }
  }
}
int front_conv(const FrontArgs& a, bool bf16_out, cudaStream_t st) {
  if (a.B <= 0 || a.Ti <= 0) return 0;
  const int tiles_per_utt = (a.Ti + FR - 1) / FR;
  const size_t smem = (size_t)3 * a.nq * FR * sizeof(float);
  FWN_CHECK(smem <= 48 * 1024, "front_conv: Cx/2=%d too large for the shared-memory tile", a.nq);
  if (bf16_out) front_kernel<__nv_bfloat16><<<a.B * tiles_per_utt, 256, smem, st>>>(a);
  else front_kernel<float><<<a.B * tiles_per_utt, 256, smem, st>>>(a);
  FWN_LAUNCH_CHECK();
  return 0;
}

}  // namespace fwn

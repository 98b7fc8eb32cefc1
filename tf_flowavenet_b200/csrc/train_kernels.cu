// Kernels of the training step (train.py:56-81) that are not GEMM epilogues:
//   fold / unfold      raw variables <-> folded parameter vector (weight norm, ZeroConv scale, bias sums) and its gradient
//   gather / scatter   folded vector <-> packed GEMM operands (signed permutation built by model_prepack)
//   make_planes        fp32 operand -> three bf16 planes of the split tensor-core engine (also transposed, for dgrad)
//   wgrad / colsum     dW[koff + k, n] = sum_rows A[row + shift, k] dY[row, n], bias gradients
//   affine_bwd, actnorm_bwd, logp_bwd, upsampler backward, Adam + global-norm clip
#include <algorithm>

#include "common.cuh"
#include "train.h"

namespace fwn {

// ---------------------------------------------------------------- fold (raw -> What) and its transpose
// One CTA (32 column lanes x 8 row slices) per work item (descriptor, column tile).  VEC = 4: a lane owns 4 consecutive columns
// (16-byte accesses, 512 contiguous bytes per warp and row); VEC = 1 for operands whose row length is not a multiple of 4.
template <int VEC>
struct VecF { float v[VEC]; };
template <int VEC>
__device__ __forceinline__ VecF<VEC> ldv(const float* p) {
  VecF<VEC> r;
  if (VEC == 4) { const float4 t = *reinterpret_cast<const float4*>(p); r.v[0] = t.x; r.v[1 % VEC] = t.y; r.v[2 % VEC] = t.z; r.v[3 % VEC] = t.w; }
  else r.v[0] = *p;
  return r;
}
template <int VEC>
__device__ __forceinline__ void stv(float* p, const VecF<VEC>& r) {
  if (VEC == 4) *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1 % VEC], r.v[2 % VEC], r.v[3 % VEC]);
  else *p = r.v[0];
}

template <int VEC>
__device__ __forceinline__ void fold_body(const float* __restrict__ raw, float* __restrict__ what, const FoldDesc& d, int col0,
                                          int64_t raw_floats, double (*red)[8][33 * 4]) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int o = col0 + tx * VEC;
  const bool ok = o < d.N;   // N % VEC == 0, so a lane is either fully inside or fully outside
  if (d.kind == FOLD_SUM2) {
    if (ok && ty == 0)
      for (int j = 0; j < VEC; ++j) what[raw_floats + d.d + o + j] = raw[d.a + o + j] + raw[d.b + o + j];
    return;
  }
  const float* v = raw + d.a;
  float sc[VEC];
  if (d.kind == FOLD_WN) {
    double ss[VEC];
    for (int j = 0; j < VEC; ++j) ss[j] = 0.0;
    if (ok)
      for (int k = ty; k < d.K; k += 8) {
        const VecF<VEC> x = ldv<VEC>(v + (int64_t)k * d.N + o);
        for (int j = 0; j < VEC; ++j) ss[j] += (double)x.v[j] * x.v[j];
      }
    for (int j = 0; j < VEC; ++j) red[0][ty][tx * VEC + j] = ss[j];
    __syncthreads();
    for (int j = 0; j < VEC; ++j) {
      double t = 0.0;
      for (int q = 0; q < 8; ++q) t += red[0][q][tx * VEC + j];
      sc[j] = ok ? (float)((double)raw[d.b + o + j] / sqrt(fmax(t, 1e-12))) : 0.f;
    }
  } else {
    for (int j = 0; j < VEC; ++j) sc[j] = ok ? expf(3.f * raw[d.c + o + j]) : 0.f;
    if (ok && ty == 0)
      for (int j = 0; j < VEC; ++j) what[d.b + o + j] = raw[d.b + o + j] * sc[j];
  }
  if (ok)
    for (int k = ty; k < d.K; k += 8) {
      VecF<VEC> x = ldv<VEC>(v + (int64_t)k * d.N + o);
      for (int j = 0; j < VEC; ++j) x.v[j] *= sc[j];
      stv<VEC>(what + d.a + (int64_t)k * d.N + o, x);
    }
}
__global__ void fold_kernel(const float* __restrict__ raw, float* __restrict__ what, const FoldDesc* __restrict__ descs,
                            const FoldWork* __restrict__ work, int64_t raw_floats) {
  __shared__ double red[2][8][33 * 4];
  const FoldWork wk = work[blockIdx.x];
  const FoldDesc d = descs[wk.desc];
  if (wk.vec) fold_body<4>(raw, what, d, wk.col0, raw_floats, red);
  else fold_body<1>(raw, what, d, wk.col0, raw_floats, red);
}

// In place on G (gradient w.r.t. What, raw layout + ext) -> gradient w.r.t. the raw variables.
template <int VEC>
__device__ __forceinline__ void unfold_body(const float* __restrict__ raw, float* __restrict__ G, const FoldDesc& d, int col0,
                                            int64_t raw_floats, double (*red)[8][33 * 4]) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int o = col0 + tx * VEC;
  const bool ok = o < d.N;
  if (d.kind == FOLD_SUM2) {
    if (ok && ty == 0)
      for (int j = 0; j < VEC; ++j) {
        const float g = G[raw_floats + d.d + o + j];
        G[d.a + o + j] = g;
        G[d.b + o + j] = g;
      }
    return;
  }
  const float* v = raw + d.a;
  float* gv = G + d.a;
  double ss[VEC], ds[VEC];  // sum v^2, sum G v
  for (int j = 0; j < VEC; ++j) ss[j] = ds[j] = 0.0;
  if (ok)
    for (int k = ty; k < d.K; k += 8) {
      const VecF<VEC> x = ldv<VEC>(v + (int64_t)k * d.N + o), gg = ldv<VEC>(gv + (int64_t)k * d.N + o);
      for (int j = 0; j < VEC; ++j) { ss[j] += (double)x.v[j] * x.v[j]; ds[j] += (double)gg.v[j] * x.v[j]; }
    }
  for (int j = 0; j < VEC; ++j) { red[0][ty][tx * VEC + j] = ss[j]; red[1][ty][tx * VEC + j] = ds[j]; }
  __syncthreads();
  if (!ok) return;
  float s[VEC], c2[VEC];
  for (int j = 0; j < VEC; ++j) {
    double tss = 0.0, tds = 0.0;
    for (int q = 0; q < 8; ++q) { tss += red[0][q][tx * VEC + j]; tds += red[1][q][tx * VEC + j]; }
    if (d.kind == FOLD_WN) {
      // What = v g / n, n = sqrt(max(sum v^2, 1e-12)):  dg = dS / n,  dv = G g / n - v dS g / n^3   (dS = sum_k G v)
      const double g = raw[d.b + o + j];
      const bool clamped = tss <= 1e-12;
      const double n = sqrt(fmax(tss, 1e-12));
      s[j] = (float)(g / n);
      c2[j] = clamped ? 0.f : (float)(tds * g / (n * n * n));
      if (ty == 0) G[d.b + o + j] = (float)(tds / n);
    } else {
      // What_w = W e, What_b = b e, e = exp(3 scale):  dW = G e, db = Gb e, dscale = 3 e (sum_k G W + Gb b)
      const float e = expf(3.f * raw[d.c + o + j]);
      s[j] = e;
      c2[j] = 0.f;
      if (ty == 0) {
        const float gb = G[d.b + o + j];
        G[d.c + o + j] = (float)(3.0 * (double)e * (tds + (double)gb * raw[d.b + o + j]));
        G[d.b + o + j] = gb * e;
      }
    }
  }
  for (int k = ty; k < d.K; k += 8) {
    const int64_t i = (int64_t)k * d.N + o;
    VecF<VEC> gg = ldv<VEC>(gv + i);
    const VecF<VEC> x = ldv<VEC>(v + i);
    for (int j = 0; j < VEC; ++j) gg.v[j] = gg.v[j] * s[j] - x.v[j] * c2[j];
    stv<VEC>(gv + i, gg);
  }
}
__global__ void unfold_kernel(const float* __restrict__ raw, float* __restrict__ G, const FoldDesc* __restrict__ descs,
                              const FoldWork* __restrict__ work, int64_t raw_floats) {
  __shared__ double red[2][8][33 * 4];
  const FoldWork wk = work[blockIdx.x];
  const FoldDesc d = descs[wk.desc];
  if (wk.vec) unfold_body<4>(raw, G, d, wk.col0, raw_floats, red);
  else unfold_body<1>(raw, G, d, wk.col0, raw_floats, red);
}

int fold_forward(const float* raw, float* what, const FoldDesc* descs, const FoldWork* work, int nwork, int64_t raw_floats, cudaStream_t st) {
  FWN_CUDA(cudaMemcpyAsync(what, raw, (size_t)raw_floats * 4, cudaMemcpyDeviceToDevice, st));
  if (nwork) fold_kernel<<<nwork, 256, 0, st>>>(raw, what, descs, work, raw_floats);
  FWN_LAUNCH_CHECK();
  return 0;
}
int fold_backward(const float* raw, float* G, const FoldDesc* descs, const FoldWork* work, int nwork, int64_t raw_floats, cudaStream_t st) {
  if (nwork) unfold_kernel<<<nwork, 256, 0, st>>>(raw, G, descs, work, raw_floats);
  FWN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- gather / scatter through the prepack map
__global__ void gather_kernel(const float* __restrict__ what, const int32_t* __restrict__ map, float* __restrict__ P, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t c = __ldg(map + i);
    if (c < 0) continue;  // constant (zero padding)
    const float v = __ldg(what + (c & 0x3FFFFFFF));
    P[i] = (c & (1 << 30)) ? -v : v;
  }
}
__global__ void scatter_kernel(const float* __restrict__ gP, const int32_t* __restrict__ map, float* __restrict__ G, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t c = __ldg(map + i);
    if (c < 0) continue;
    const float v = __ldg(gP + i);
    G[c & 0x3FFFFFFF] = (c & (1 << 30)) ? -v : v;   // every folded element feeds exactly one packed element
  }
}
static int ew_grid_t(int64_t n) { return (int)std::min<int64_t>(cdiv(n, 256), (int64_t)num_sms() * 16); }
int gather_pack(const float* what, const int32_t* map, float* P, int64_t n, cudaStream_t st) {
  if (n) gather_kernel<<<ew_grid_t(n), 256, 0, st>>>(what, map, P, n);
  FWN_LAUNCH_CHECK();
  return 0;
}
int scatter_grad(const float* gP, const int32_t* map, float* G, int64_t n, cudaStream_t st) {
  if (n) scatter_kernel<<<ew_grid_t(n), 256, 0, st>>>(gP, map, G, n);
  FWN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- fp32 operand -> bf16x3 planes [3][Npad][Kpad]
__global__ void make_planes_kernel(const PlaneDesc* __restrict__ descs, const PlaneWork* __restrict__ work, int nplanes) {
  __shared__ float tile[32][33];
  const PlaneWork wk = work[blockIdx.x];
  const PlaneDesc d = descs[wk.desc];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int k0 = wk.kt * 32, n0 = wk.nt * 32;
  // read: coalesce along whichever index is contiguous in the source
  if (d.sn == 1) {
    for (int j = ty; j < 32; j += 8) {
      const int k = k0 + j, n = n0 + tx;
      tile[j][tx] = (k < d.K && n < d.N) ? __ldg(d.src + (int64_t)k * d.sk + n) : 0.f;
    }
  } else {
    for (int j = ty; j < 32; j += 8) {
      const int k = k0 + tx, n = n0 + j;
      tile[tx][j] = (k < d.K && n < d.N) ? __ldg(d.src + (int64_t)k * d.sk + (int64_t)n * d.sn) : 0.f;
    }
  }
  __syncthreads();
  const size_t plane = (size_t)d.Kpad * d.Npad;
  for (int j = ty; j < 32; j += 8) {
    const int n = n0 + j, k = k0 + tx;
    if (n >= d.N || k >= d.K) continue;
    const float x = tile[tx][j];
    const __nv_bfloat16 h1 = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(h1);
    const __nv_bfloat16 h2 = __float2bfloat16_rn(r1);
    const __nv_bfloat16 h3 = __float2bfloat16_rn(r1 - __bfloat162float(h2));
    const size_t i = (size_t)n * d.Kpad + d.k0 + k;
    d.dst[i] = h1;
    if (nplanes == 3) {   // the bf16 training mode reads plane 0 only
      d.dst[plane + i] = h2;
      d.dst[2 * plane + i] = h3;
    }
  }
}
int make_planes(const PlaneDesc* descs, const PlaneWork* work, int nwork, int nplanes, cudaStream_t st) {
  if (nwork) make_planes_kernel<<<nwork, 256, 0, st>>>(descs, work, nplanes);
  FWN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- ActNorm vectors in physical order + sum of log-dets
__global__ void actnorm_pack_kernel(const ActnormDesc* __restrict__ descs, int nflows, double* __restrict__ an_logdet) {
  const ActnormDesc d = descs[blockIdx.x];
  double ld = 0.0;
  for (int o = threadIdx.x; o < d.Cx; o += blockDim.x) {
    const int l = d.off2log[o];
    const float logs = d.raw_logs[l];
    d.an_b[o] = d.raw_b[l];
    d.an_s[o] = (float)exp(3.0 * (double)logs);
    d.an_is[o] = (float)exp(-3.0 * (double)logs);
    ld += 3.0 * (double)logs;
  }
  __shared__ double red[32];
  ld = block_sum(ld, red);
  if (threadIdx.x == 0) atomicAdd(an_logdet, ld / d.Cx);
}
int actnorm_pack(const ActnormDesc* descs, int nflows, double* an_logdet, cudaStream_t st) {
  FWN_CUDA(cudaMemsetAsync(an_logdet, 0, sizeof(double), st));
  if (nflows) actnorm_pack_kernel<<<nflows, 256, 0, st>>>(descs, nflows, an_logdet);
  FWN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- wgrad (CUDA cores, fp32): dW[koff + k, n] += sum_rows A[row + shift, k] dY[row, n]
// 128 x 128 output tile per CTA (8 x 8 per thread), rows streamed 8 at a time through shared memory; the row range is split over
// blockIdx.z and partial tiles are combined with fp32 atomics (dW is zeroed by the caller).
constexpr int WG_T = 128, WG_R = 8;
__global__ void __launch_bounds__(256) wgrad_kernel(const WgradArgs a) {
  __shared__ __align__(16) float As[2][WG_R][WG_T];
  __shared__ __align__(16) float Ys[2][WG_R][WG_T];
  // which K tile (segment + offset)
  int kt = blockIdx.x, sidx = 0;
  while (sidx < a.nseg - 1 && kt >= (a.seg[sidx].K + WG_T - 1) / WG_T) { kt -= (a.seg[sidx].K + WG_T - 1) / WG_T; ++sidx; }
  const Seg sg = a.seg[sidx];
  const float* A = reinterpret_cast<const float*>(sg.A);
  const int k0 = kt * WG_T, n0 = blockIdx.y * WG_T;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  // row slab of this CTA: rows are (b, t); slab = contiguous t range of one utterance
  const int slabs_per_utt = a.slabs_per_utt;
  const int ub = blockIdx.z / slabs_per_utt, sl = blockIdx.z - ub * slabs_per_utt;
  const int t_begin = (int)((int64_t)a.Ti * sl / slabs_per_utt), t_end = (int)((int64_t)a.Ti * (sl + 1) / slabs_per_utt);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  // loader mapping: 256 threads x 4 floats = 8 rows x 128 columns
  const int lr = tid >> 5, lc = (tid & 31) * 4;
  auto load = [&](int buf, int t) {
    const int tr = t + lr;
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f), yv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tr < t_end) {
      const int ta = tr + sg.shift;
      if (ta >= 0 && ta < a.Ti) {
        const float* p = A + ((int64_t)ub * a.Ti + ta) * sg.lda + k0 + lc;
        if (k0 + lc + 3 < sg.K && (sg.lda & 3) == 0) av = __ldg(reinterpret_cast<const float4*>(p));
        else {
          float t4[4] = {0.f, 0.f, 0.f, 0.f};
          for (int j = 0; j < 4; ++j) if (k0 + lc + j < sg.K) t4[j] = __ldg(p + j);
          av = make_float4(t4[0], t4[1], t4[2], t4[3]);
        }
      }
      float t4[4] = {0.f, 0.f, 0.f, 0.f};
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + lc + j;
        if (n < a.N) t4[j] = n < a.n0cols ? __ldg(a.dY0 + ((int64_t)ub * a.Ti + tr) * a.ld0 + n)
                                         : __ldg(a.dY1 + ((int64_t)ub * a.Ti + tr) * a.ld1 + (n - a.n0cols));
      }
      yv = make_float4(t4[0], t4[1], t4[2], t4[3]);
    }
    *reinterpret_cast<float4*>(&As[buf][lr][lc]) = av;
    *reinterpret_cast<float4*>(&Ys[buf][lr][lc]) = yv;
  };
  int buf = 0;
  if (t_begin < t_end) load(0, t_begin);
  __syncthreads();
  for (int t = t_begin; t < t_end; t += WG_R) {
    if (t + WG_R < t_end) load(buf ^ 1, t + WG_R);
#pragma unroll
    for (int r = 0; r < WG_R; ++r) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][r][ty * 8]), a1 = *reinterpret_cast<const float4*>(&As[buf][r][ty * 8 + 4]);
      const float4 y0 = *reinterpret_cast<const float4*>(&Ys[buf][r][tx * 8]), y1 = *reinterpret_cast<const float4*>(&Ys[buf][r][tx * 8 + 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float yv[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], yv[j], acc[i][j]);
    }
    __syncthreads();
    buf ^= 1;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = k0 + ty * 8 + i;
    if (k >= sg.K) continue;
    float* drow = a.dW + (int64_t)(sg.koff + k) * a.ldw;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + tx * 8 + j;
      if (n < a.N) atomicAdd(drow + n, acc[i][j]);
    }
  }
}
int wgrad(const WgradArgs& a_in, cudaStream_t st) {
  WgradArgs a = a_in;
  if (a.B <= 0 || a.Ti <= 0 || a.N <= 0) return 0;
  int ktiles = 0;
  for (int s = 0; s < a.nseg; ++s) ktiles += (a.seg[s].K + WG_T - 1) / WG_T;
  const int ntiles = (a.N + WG_T - 1) / WG_T;
  // enough row slabs to fill the machine ~2x, but at least 64 rows per slab
  const int64_t tiles = (int64_t)ktiles * ntiles;
  int64_t want = std::max<int64_t>(1, ((int64_t)num_sms() * 2 + tiles - 1) / tiles);
  int per_utt = (int)std::max<int64_t>(1, std::min<int64_t>((want + a.B - 1) / a.B, std::max(1, a.Ti / 64)));
  a.slabs_per_utt = per_utt;
  FWN_CHECK((int64_t)a.B * per_utt <= 65535, "wgrad: too many row slabs");
  dim3 grid((unsigned)ktiles, (unsigned)ntiles, (unsigned)(a.B * per_utt));
  wgrad_kernel<<<grid, 256, 0, st>>>(a);
  FWN_LAUNCH_CHECK();
  return 0;
}

// column sums (bias gradients): out[n] += sum_rows dY[row, n]; two column segments like wgrad
__global__ void colsum_kernel(const float* __restrict__ dY0, int64_t ld0, int n0cols, const float* __restrict__ dY1, int64_t ld1, int N,
                              int64_t rows, float* __restrict__ out) {
  // blockDim = (32, 8): 32 columns x 8 row lanes
  __shared__ double red[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  double s = 0.0;
  if (n < N) {
    const float* p = n < n0cols ? dY0 + n : dY1 + (n - n0cols);
    const int64_t ld = n < n0cols ? ld0 : ld1;
    for (int64_t r = blockIdx.y * 8 + threadIdx.y; r < rows; r += (int64_t)gridDim.y * 8) s += (double)__ldg(p + r * ld);
  }
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    for (int j = 1; j < 8; ++j) s += red[j][threadIdx.x];
    atomicAdd(out + n, (float)s);
  }
}
int colsum(const float* dY0, int64_t ld0, int n0cols, const float* dY1, int64_t ld1, int N, int64_t rows, float* out, cudaStream_t st) {
  if (rows <= 0 || N <= 0) return 0;
  const int gx = (N + 31) / 32;
  const int gy = (int)std::max<int64_t>(1, std::min<int64_t>(cdiv(rows, 64), std::max<int64_t>(1, (int64_t)num_sms() * 4 / gx)));
  colsum_kernel<<<dim3(gx, gy), dim3(32, 8), 0, st>>>(dY0, ld0, n0cols, dY1, ld1, N, rows, out);
  FWN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- flow-variable gradients
// loss = -(log_p + logdet) (train.py:59): d loss / d z = z / (B T)   (model.py:343)
__global__ void logp_bwd_kernel(const float* __restrict__ z, float* __restrict__ dX, int64_t n, float inv_n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dX[i] = z[i] * inv_n;
}
int logp_bwd(const float* z, float* dX, int64_t n, cudaStream_t st) {
  logp_bwd_kernel<<<ew_grid_t(n), 256, 0, st>>>(z, dX, n, (float)(1.0 / (double)n));
  FWN_LAUNCH_CHECK();
  return 0;
}

// AffineCoupling backward (model.py:133-135): out_b = (b - t) exp(-log_s), logdet = mean(-log_s)/2.
//   d log_s = -g out_b + 1/(B T),  d t = -g exp(-log_s),  d b = g exp(-log_s)      (g = d loss / d out_b)
__global__ void affine_bwd_kernel(float* __restrict__ dX, const float* __restrict__ Xpost, const float* __restrict__ net, int64_t ldn,
                                  float* __restrict__ dNet, int64_t rows, int Cx, int nq, const int* __restrict__ b_off, float inv_n) {
  const int64_t n = rows * nq;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / nq;
    const int p = (int)(i - row * nq);
    const int ob = __ldg(b_off + p);
    const float g = dX[row * Cx + ob];
    const float outb = __ldg(Xpost + row * Cx + ob);
    const float2 lt = *reinterpret_cast<const float2*>(net + row * ldn + 2 * p);
    const float el = expf(-lt.x);
    *reinterpret_cast<float2*>(dNet + row * ldn + 2 * p) = make_float2(-g * outb + inv_n, -g * el);
    dX[row * Cx + ob] = g * el;
  }
}
int affine_bwd(float* dX, const float* Xpost, const float* net, int64_t ldn, float* dNet, int64_t rows, int Cx, int nq, const int* b_off,
               double n_total, cudaStream_t st) {
  if (ldn != 2 * nq) FWN_CUDA(cudaMemsetAsync(dNet, 0, (size_t)rows * ldn * 4, st));
  affine_bwd_kernel<<<ew_grid_t(rows * nq), 256, 0, st>>>(dX, Xpost, net, ldn, dNet, rows, Cx, nq, b_off, (float)(1.0 / n_total));
  FWN_LAUNCH_CHECK();
  return 0;
}

// ActNorm backward (model.py:86-94) fused with the gradient arriving through the WaveNet input (pass-through half):
//   dy = dX + da0 (a-half only);  y = (x + b) s, s = exp(3 logs):  dx = dy s,  db = sum dy s,  dlogs = 3 sum dy y - 3/Cx
__global__ void actnorm_bwd_kernel(float* __restrict__ dX, const float* __restrict__ da0, int ld_a0, const float* __restrict__ xpre,
                                   const float* __restrict__ an_b, const float* __restrict__ an_s, const int* __restrict__ off2log,
                                   int64_t rows, int Cx, int nq, float* __restrict__ g_b, float* __restrict__ g_logs) {
  extern __shared__ double part[];  // [2][blockDim]
  const int nth = blockDim.x;
  if (Cx <= nth) {
    const int o = threadIdx.x % Cx;
    const int l = __ldg(off2log + o);
    const float b = __ldg(an_b + o), s = __ldg(an_s + o);
    double sb = 0.0, sl = 0.0;
    const int64_t n = rows * Cx;
    for (int64_t i = (int64_t)blockIdx.x * nth + threadIdx.x; i < n; i += (int64_t)gridDim.x * nth) {
      const int64_t row = i / Cx;
      float dy = dX[i];
      if (l < nq) dy += __ldg(da0 + row * ld_a0 + l);
      const float y = (__ldg(xpre + i) + b) * s;
      dX[i] = dy * s;
      sb += (double)dy * s;
      sl += (double)dy * y;
    }
    part[threadIdx.x] = sb;
    part[nth + threadIdx.x] = sl;
    __syncthreads();
    if ((int)threadIdx.x < Cx) {
      double tb = 0.0, tl = 0.0;
      for (int j = threadIdx.x; j < nth; j += Cx) { tb += part[j]; tl += part[nth + j]; }
      if (blockIdx.x == 0) tl -= 1.0 / Cx;   // the log-det term, once
      atomicAdd(g_b + l, (float)tb);
      atomicAdd(g_logs + l, (float)(3.0 * tl));
    }
  } else {
    for (int o = threadIdx.x; o < Cx; o += nth) {
      const int l = __ldg(off2log + o);
      const float b = __ldg(an_b + o), s = __ldg(an_s + o);
      double sb = 0.0, sl = 0.0;
      for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
        const int64_t i = row * Cx + o;
        float dy = dX[i];
        if (l < nq) dy += __ldg(da0 + row * ld_a0 + l);
        const float y = (__ldg(xpre + i) + b) * s;
        dX[i] = dy * s;
        sb += (double)dy * s;
        sl += (double)dy * y;
      }
      if (blockIdx.x == 0) sl -= 1.0 / Cx;
      atomicAdd(g_b + l, (float)sb);
      atomicAdd(g_logs + l, (float)(3.0 * sl));
    }
  }
}
int actnorm_bwd(float* dX, const float* da0, int ld_a0, const float* xpre, const float* an_b, const float* an_s, const int* off2log,
                int64_t rows, int Cx, int nq, float* g_b, float* g_logs, cudaStream_t st) {
  const int64_t n = rows * Cx;
  int grid = (int)std::min<int64_t>((int64_t)num_sms() * 4, std::max<int64_t>(1, n / 1024));
  if (Cx > 256) grid = (int)std::min<int64_t>(rows, (int64_t)num_sms() * 2);
  actnorm_bwd_kernel<<<grid, 256, 2 * 256 * sizeof(double), st>>>(dX, da0, ld_a0, xpre, an_b, an_s, off2log, rows, Cx, nq, g_b, g_logs);
  FWN_LAUNCH_CHECK();
  return 0;
}

// fp32 twin of front_pack (conv_simt.cu): A0[row, q] = ActNorm(x)[row, logical channel q], q < nq; row pitch kq, zero padded
__global__ void front_pack_f32_kernel(const float* __restrict__ X, int Cx, int nq, int kq, const int* __restrict__ off2log,
                                      const float* __restrict__ an_b, const float* __restrict__ an_s, float* __restrict__ A0, int64_t rows) {
  const int64_t n = rows * Cx;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / Cx;
    const int o = (int)(i - row * Cx);
    const int q = __ldg(off2log + o);
    if (q >= nq) continue;
    float v = __ldg(X + i);
    if (an_b) v = (v + __ldg(an_b + o)) * __ldg(an_s + o);   // forward direction: ActNorm on load; reverse: identity
    A0[row * kq + q] = v;
  }
}
int front_pack_f32(const float* X, int Cx, int nq, int kq, const int* off2log, const float* an_b, const float* an_s, float* A0, int64_t rows,
                   cudaStream_t st) {
  if (kq != nq) FWN_CUDA(cudaMemsetAsync(A0, 0, (size_t)rows * kq * 4, st));
  front_pack_f32_kernel<<<ew_grid_t(rows * Cx), 256, 0, st>>>(X, Cx, nq, kq, off2log, an_b, an_s, A0, rows);
  FWN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- upsampler backward (model.py:398-404, convolutional.py:155-201)
// Forward of one stage: out[b,i,m] = lrelu(bias + sum_{a<2, kw<3} in[b, j0-a, m+1-kw] w[r + a s, kw]),  q = i + s/2, r = q % s, j0 = q / s.
// Gradient buffers of the last stage are the two mel-half planes [B*To, mels/2] (SPLIT), else one tensor [B*To, mels].
struct PlanePair {
  const float* p0; const float* p1; int mels, half; bool split;
  __device__ __forceinline__ float at(int64_t bt, int m) const {
    if (!split) return __ldg(p0 + bt * mels + m);
    return m < half ? __ldg(p0 + bt * half + m) : __ldg(p1 + bt * half + (m - half));
  }
};
// d pre-activation = dout * (out > 0 ? 1 : 0.4)
__device__ __forceinline__ float dpre_at(const PlanePair& dout, const PlanePair& out, int64_t bt, int m) {
  const float o = out.at(bt, m);
  return dout.at(bt, m) * (o > 0.f ? 1.f : 0.4f);
}
// weight + bias gradient: one warp per output row (b, i); lanes stride the mel axis
__global__ void upsample_bwd_w_kernel(PlanePair dout, PlanePair out, const float* __restrict__ in, int B, int Tm, int mels, int s,
                                      double* __restrict__ dw /*[2s*3]*/, double* __restrict__ dbias) {
  // these 6s+1 sums run over every output element and are then differenced by the weight-norm chain: accumulate in double
  extern __shared__ double sacc[];  // [2s*3 + 1]
  const int nw = 2 * s * 3 + 1;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) sacc[i] = 0.0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, warp = threadIdx.x >> 5;
  const int To = Tm * s;
  const int64_t nrows = (int64_t)B * To;
  for (int64_t bt = (int64_t)blockIdx.x * wpb + warp; bt < nrows; bt += (int64_t)gridDim.x * wpb) {
    const int i = (int)(bt % To), b = (int)(bt / To);
    const int q = i + s / 2, r = q % s, j0 = q / s;
    float part[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int m = lane; m < mels; m += 32) {
      const float dp = dpre_at(dout, out, bt, m);
      part[6] += dp;
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int j = j0 - a;
        if (j < 0 || j >= Tm) continue;
        const float* row = in + ((int64_t)b * Tm + j) * mels;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int mm = m + 1 - kw;
          if (mm >= 0 && mm < mels) part[a * 3 + kw] = fmaf(dp, __ldg(row + mm), part[a * 3 + kw]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) part[k] = warp_sum(part[k]);
    if (lane == 0) {
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) atomicAdd(&sacc[(r + a * s) * 3 + kw], (double)part[a * 3 + kw]);
      atomicAdd(&sacc[nw - 1], (double)part[6]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nw - 1; i += blockDim.x) atomicAdd(dw + i, sacc[i]);
  if (threadIdx.x == 0) atomicAdd(dbias, sacc[nw - 1]);
}
// input gradient: din[b,j,mm] = sum_{a,r,kw} dpre[b, (j+a) s + r - s/2, mm-1+kw] w[r + a s, kw]
__global__ void upsample_bwd_in_kernel(PlanePair dout, PlanePair out, const float* __restrict__ w, float* __restrict__ din, int B, int Tm,
                                       int mels, int s) {
  extern __shared__ float sw[];
  for (int i = threadIdx.x; i < 2 * s * 3; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int To = Tm * s;
  const int64_t n = (int64_t)B * Tm * mels;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int mm = (int)(idx % mels);
    const int64_t bj = idx / mels;
    const int j = (int)(bj % Tm), b = (int)(bj / Tm);
    float acc = 0.f;
    for (int a = 0; a < 2; ++a)
      for (int r = 0; r < s; ++r) {
        const int i = (j + a) * s + r - s / 2;
        if (i < 0 || i >= To) continue;
        const int64_t bt = (int64_t)b * To + i;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int m = mm - 1 + kw;
          if (m >= 0 && m < mels) acc = fmaf(dpre_at(dout, out, bt, m), sw[(r + a * s) * 3 + kw], acc);
        }
      }
    din[idx] = acc;
  }
}
// weight norm of the [2s,3,1,1] kernel over axes [0,2] (per kw column, convolutional.py:186) -- backward
__global__ void upsample_wn_bwd_kernel(const float* __restrict__ v, const float* __restrict__ g, const double* __restrict__ dw, int s,
                                       float* __restrict__ gv, float* __restrict__ gg, float* __restrict__ gbias) {
  // one warp; lanes 0..2 own a kw column each
  const int kw = threadIdx.x;
  float dg = 0.f;
  if (kw < 3) {
    double ss = 0.0, ds = 0.0;
    for (int kh = 0; kh < 2 * s; ++kh) { const float x = v[kh * 3 + kw]; ss += (double)x * x; ds += dw[kh * 3 + kw] * x; }
    const bool clamped = ss <= 1e-12;
    const double n = sqrt(fmax(ss, 1e-12));
    const double sc = g[0] / n, c2 = clamped ? 0.0 : ds * g[0] / (n * n * n);
    for (int kh = 0; kh < 2 * s; ++kh) gv[kh * 3 + kw] = (float)(dw[kh * 3 + kw] * sc - v[kh * 3 + kw] * c2);
    dg = (float)(ds / n);
  }
  dg = warp_sum(dg);
  if (threadIdx.x == 0) { gg[0] = dg; gbias[0] = (float)dw[2 * s * 3]; }
}
int upsample_bwd_stage(const float* dout0, const float* dout1, const float* out0, const float* out1, bool split, const float* in,
                       const float* w, double* dw_scratch /*[2s*3+1], zeroed here*/, float* din /*nullable*/, int B, int Tm, int mels, int s,
                       cudaStream_t st) {
  PlanePair d{dout0, dout1, mels, mels / 2, split}, o{out0, out1, mels, mels / 2, split};
  const int nw = 2 * s * 3 + 1;
  FWN_CUDA(cudaMemsetAsync(dw_scratch, 0, nw * sizeof(double), st));
  const int64_t nrows = (int64_t)B * Tm * s;
  const int grid = (int)std::min<int64_t>(cdiv(nrows, 8), (int64_t)num_sms() * 8);
  upsample_bwd_w_kernel<<<grid, 256, nw * sizeof(double), st>>>(d, o, in, B, Tm, mels, s, dw_scratch, dw_scratch + nw - 1);
  FWN_LAUNCH_CHECK();
  if (din) {
    const int64_t n = (int64_t)B * Tm * mels;
    upsample_bwd_in_kernel<<<ew_grid_t(n), 256, 2 * s * 3 * sizeof(float), st>>>(d, o, w, din, B, Tm, mels, s);
    FWN_LAUNCH_CHECK();
  }
  return 0;
}
int upsample_wn_bwd(const float* v, const float* g, const double* dw, int s, float* gv, float* gg, float* gbias, cudaStream_t st) {
  upsample_wn_bwd_kernel<<<1, 32, 0, st>>>(v, g, dw, s, gv, gg, gbias);
  FWN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- optimizer (train.py:15-32,76-81)
__global__ void sumsq_f_kernel(const float* __restrict__ x, int64_t n, double* __restrict__ acc) {
  __shared__ double red[32];
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = __ldg(x + i);
    s += (double)v * v;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(acc, s);
}
__global__ void finish_norm_kernel(const double* acc, float* norm_out) { *norm_out = (float)sqrt(*acc); }
int grad_global_norm(const float* g, int64_t n, double* scratch, float* norm_out, cudaStream_t st) {
  FWN_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double), st));
  sumsq_f_kernel<<<ew_grid_t(n), 256, 0, st>>>(g, n, scratch);
  FWN_LAUNCH_CHECK();
  finish_norm_kernel<<<1, 1, 0, st>>>(scratch, norm_out);
  FWN_LAUNCH_CHECK();
  return 0;
}
// tf.clip_by_global_norm(grads, clip): g * clip / max(norm, clip);  tf.train.AdamOptimizer (beta1 .9, beta2 .999, eps 1e-8):
//   lr_t = lr sqrt(1 - b2^t) / (1 - b1^t);  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr_t m / (sqrt(v) + eps)
__global__ void adam_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ g,
                            const float* __restrict__ norm, float clip, float lr_t, float b1, float b2, float eps, int64_t n) {
  float scale = 1.f;
  if (clip > 0.f) scale = clip / fmaxf(*norm, clip);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * scale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}
int adam_update(float* p, float* m, float* v, const float* g, const float* norm, float clip, float lr, float b1, float b2, float eps,
                int64_t step, int64_t n, cudaStream_t st) {
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)b2, (double)step)) / (1.0 - pow((double)b1, (double)step));
  adam_kernel<<<ew_grid_t(n), 256, 0, st>>>(p, m, v, g, norm, clip, (float)lr_t, b1, b2, eps, n);
  FWN_LAUNCH_CHECK();
  return 0;
}

}  // namespace fwn

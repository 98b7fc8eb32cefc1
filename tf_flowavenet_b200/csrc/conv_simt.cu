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
#include "epilogue_f32.cuh"

namespace fwn {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256;

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
    case EPI_LINEAR: simt_gemm_kernel<EPI_LINEAR><<<grid, NT, 0, st>>>(a); break;
    case EPI_GATE_BWD: simt_gemm_kernel<EPI_GATE_BWD><<<grid, NT, 0, st>>>(a); break;
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
// K = 3*Cx/2 is 3..384: 384 MAC per audio sample in every block, 0.2% of the pass -> CUDA cores, fp32.
// CTA = 64 rows x 256 channels; a warp owns 8 rows, a lane 8 consecutive channels (64 accumulators), so every
// output row is written as one 512-byte (bf16) coalesced store.  X is read as ONE contiguous span per CTA.
constexpr int FR = 64;          // rows per CTA
constexpr int FRW = FR / 8;     // rows per warp
template <typename TOut>
__global__ void __launch_bounds__(256) front_kernel(const FrontArgs a) {
  extern __shared__ __align__(16) float xs[];  // [3][nq][FR]: one copy per tap, already shifted
  const int nq = a.nq, Cx = a.Cx;
  const int tiles_per_utt = (a.Ti + FR - 1) / FR;
  const int ub = blockIdx.x / tiles_per_utt;
  const int t0 = (blockIdx.x - ub * tiles_per_utt) * FR;
  int smin = a.shift[0], smax = a.shift[0];
#pragma unroll
  for (int k = 1; k < 3; ++k) { smin = min(smin, a.shift[k]); smax = max(smax, a.shift[k]); }
  for (int i = threadIdx.x; i < 3 * nq * FR; i += blockDim.x) xs[i] = 0.f;
  __syncthreads();
  // contiguous span of X covering rows t0+smin .. t0+FR-1+smax of this utterance
  const int r_lo = max(t0 + smin, 0), r_hi = min(t0 + FR - 1 + smax, a.Ti - 1);
  const float* xbase = a.X + (int64_t)ub * a.Ti * Cx;
  for (int64_t i = (int64_t)r_lo * Cx + threadIdx.x; i < (int64_t)(r_hi + 1) * Cx; i += blockDim.x) {
    const int t = (int)(i / Cx), o = (int)(i - (int64_t)t * Cx);
    const int q = __ldg(a.off2log + o);  // logical channel; the pass-through half is q < nq
    if (q >= nq) continue;
    float v = __ldg(xbase + i);
    if (a.an_b) v = (v + __ldg(a.an_b + o)) * __ldg(a.an_s + o);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int r = t - a.shift[k] - t0;  // output row that reads input row t through tap k
      if (r >= 0 && r < FR) xs[(k * nq + q) * FR + r] = v;
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int ch0 = lane * 8; ch0 < a.F; ch0 += 256) {
    float acc[FRW][8];
    {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.bias + ch0)), b1 = __ldg(reinterpret_cast<const float4*>(a.bias + ch0 + 4));
#pragma unroll
      for (int r = 0; r < FRW; ++r) {
        acc[r][0] = b0.x; acc[r][1] = b0.y; acc[r][2] = b0.z; acc[r][3] = b0.w;
        acc[r][4] = b1.x; acc[r][5] = b1.y; acc[r][6] = b1.z; acc[r][7] = b1.w;
      }
    }
    for (int kq = 0; kq < 3 * nq; ++kq) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(a.W + (int64_t)kq * a.F + ch0));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(a.W + (int64_t)kq * a.F + ch0 + 4));
      const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      const float4 x0 = *reinterpret_cast<const float4*>(xs + kq * FR + warp * FRW);
      const float4 x1 = *reinterpret_cast<const float4*>(xs + kq * FR + warp * FRW + 4);
      const float x[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
      for (int r = 0; r < FRW; ++r)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[r][j] = fmaf(x[r], w[j], acc[r][j]);
    }
#pragma unroll
    for (int r = 0; r < FRW; ++r) {
      const int t = t0 + warp * FRW + r;
      if (t >= a.Ti) continue;
      const int64_t off = ((int64_t)ub * a.Ti + t) * a.F + ch0;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(acc[r][j], 0.f);
      if (sizeof(TOut) == 2) {
        __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
        uint4 u = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1), *reinterpret_cast<uint32_t*>(&p2),
                             *reinterpret_cast<uint32_t*>(&p3));
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.H) + off) = u;
      } else {
        float4* hp = reinterpret_cast<float4*>(reinterpret_cast<float*>(a.H) + off);
        hp[0] = make_float4(v[0], v[1], v[2], v[3]);
        hp[1] = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
  }
}
int front_conv(const FrontArgs& a, bool bf16_out, cudaStream_t st) {
  if (a.B <= 0 || a.Ti <= 0) return 0;
  FWN_CHECK(a.F % 8 == 0, "front_conv: filter size must be a multiple of 8");
  const int tiles_per_utt = (a.Ti + FR - 1) / FR;
  const size_t smem = (size_t)3 * a.nq * FR * sizeof(float);
  FWN_CHECK(smem <= 200 * 1024, "front_conv: Cx/2=%d too large for the shared-memory tile", a.nq);
  static bool configured = false;
  if (!configured) {
    FWN_CUDA(cudaFuncSetAttribute(front_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    FWN_CUDA(cudaFuncSetAttribute(front_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  if (bf16_out) front_kernel<__nv_bfloat16><<<a.B * tiles_per_utt, 256, smem, st>>>(a);
  else front_kernel<float><<<a.B * tiles_per_utt, 256, smem, st>>>(a);
  FWN_LAUNCH_CHECK();
  return 0;
}

// Front conv of the shallow blocks (nq = C_x/2 <= 4 input channels: 6 .. 24 MAC per output) for the mixed modes, straight from the
// flow variable: h0[row, ch] = relu(b[ch] + sum_{k,q} a(row + shift_k, q) W[k,q,ch]) with a = ActNorm'd pass-through half of x (fp32,
// zero outside the utterance = tf.pad, modules.py:27).  Write-bound (512 B per row); replaces front_pack + a K = 3 x 16 tensor-core
// GEMM whose operand tiles are 94 % padding.  A block walks tiles of FD_ROWS consecutive rows of one utterance: the tile's inputs
// (+ halo) are gathered once into shared memory -- the NEXT tile's inputs are already in flight in registers while this tile is
// computed --, then each warp produces 8 rows: a lane owns 8 channels (its 3 nq x 8 weights live in registers), the 3 nq inputs of
// a row are broadcast shared-memory reads, and a row leaves as 32 contiguous 16-byte stores, 8 independent rows in flight per warp.
// Rows per tile: 256 / 128 / 64 for nq = 1 / 2 / 4 -- one tile's compute (~3 k cycles) then covers the DRAM latency of the next tile's
// inputs, which are requested one tile ahead (with 64-row tiles every tile waited ~1.5 k cycles for them: 0.48 of the HBM write rate).
template <int NQ, typename T16, int FD_ROWS>
__global__ void __launch_bounds__(256, 2) front_direct_kernel(const FrontArgs a, int lo, int hi, int tiles_per_utt, int n_tiles) {
  constexpr int SPAN = FD_ROWS + 64;   // rows [t0 + lo, t0 + FD_ROWS + hi) with -32 <= lo <= 0 <= hi <= 32
  constexpr int PER = (SPAN * NQ + 255) / 256;
  __shared__ float xs[SPAN * NQ];
  __shared__ int s_off[NQ];
  __shared__ float s_ab[NQ], s_as[NQ];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ch0 = lane * 8;
  float w[3 * NQ][8], bias[8];
#pragma unroll
  for (int kq = 0; kq < 3 * NQ; ++kq) {
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(a.W + (int64_t)kq * a.F + ch0));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(a.W + (int64_t)kq * a.F + ch0 + 4));
    w[kq][0] = w0.x; w[kq][1] = w0.y; w[kq][2] = w0.z; w[kq][3] = w0.w;
    w[kq][4] = w1.x; w[kq][5] = w1.y; w[kq][6] = w1.z; w[kq][7] = w1.w;
  }
  {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.bias + ch0)), b1 = __ldg(reinterpret_cast<const float4*>(a.bias + ch0 + 4));
    bias[0] = b0.x; bias[1] = b0.y; bias[2] = b0.z; bias[3] = b0.w; bias[4] = b1.x; bias[5] = b1.y; bias[6] = b1.z; bias[7] = b1.w;
  }
  // programmatic dependent launch: the weights above do not depend on the previous kernel of the chain (the tail of the previous flow),
  // everything below does (ActNorm parameters during the data-dependent init pass, x always)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // physical offset of logical pass-through channel q inside a row of X, and its ActNorm (identity in the reverse direction)
  if (threadIdx.x < a.Cx) {
    const int o = threadIdx.x, l = __ldg(a.off2log + o);
    if (l < NQ) {
      s_off[l] = o;
      s_ab[l] = a.an_b ? __ldg(a.an_b + o) : 0.f;
      s_as[l] = a.an_b ? __ldg(a.an_s + o) : 1.f;
    }
  }
  __syncthreads();
  const int Ti = a.Ti, Cx = a.Cx, span = FD_ROWS + hi - lo;
  // element i of a tile's input window: row t0 + lo + i / NQ, channel i % NQ; 0 outside the utterance
  auto fetch = [&](int tile, float* v) {
    const int ub = tile / tiles_per_utt, t0 = (tile - ub * tiles_per_utt) * FD_ROWS;
#pragma unroll
    for (int p = 0; p < PER; ++p) {
      const int i = threadIdx.x + p * 256, r = i / NQ, q = i - r * NQ, t = t0 + lo + r;
      v[p] = 0.f;
      if (tile < n_tiles && r < span && t >= 0 && t < Ti) v[p] = (__ldg(a.X + ((int64_t)ub * Ti + t) * Cx + s_off[q]) + s_ab[q]) * s_as[q];
    }
  };
  float nxt[PER];
  int tile = blockIdx.x;
  fetch(tile, nxt);
  for (; tile < n_tiles; tile += gridDim.x) {
    __syncthreads();   // every warp is done with the previous tile's window
#pragma unroll
    for (int p = 0; p < PER; ++p)
      if (threadIdx.x + p * 256 < SPAN * NQ) xs[threadIdx.x + p * 256] = nxt[p];
    __syncthreads();
    fetch(tile + gridDim.x, nxt);
    const int ub = tile / tiles_per_utt, t0 = (tile - ub * tiles_per_utt) * FD_ROWS;
#pragma unroll
    for (int j8 = 0; j8 < FD_ROWS / 8; ++j8) {
      const int r = warp + 8 * j8, t = t0 + r;
      if (t >= Ti) break;
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = bias[j];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float* xr = xs + (r + a.shift[k] - lo) * NQ;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const float v = xr[q];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(v, w[k * NQ + q][j], acc[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaxf(acc[j], 0.f);
      *reinterpret_cast<uint4*>(reinterpret_cast<T16*>(a.H) + ((int64_t)ub * Ti + t) * a.F + ch0) =
          make_uint4(pack2<T16>(acc[0], acc[1]), pack2<T16>(acc[2], acc[3]), pack2<T16>(acc[4], acc[5]), pack2<T16>(acc[6], acc[7]));
    }
  }
}
bool front_direct_supported(const FrontArgs& a) {
  if (!(a.F == 256 && (a.nq == 1 || a.nq == 2 || a.nq == 4) && a.Cx <= 256)) return false;
  for (int k = 0; k < 3; ++k)
    if (a.shift[k] < -32 || a.shift[k] > 32) return false;
  return true;
}
int front_direct(const FrontArgs& a, bool fp16, cudaStream_t st) {
  if (a.B <= 0 || a.Ti <= 0) return 0;
  FWN_CHECK(front_direct_supported(a), "front_direct: needs F = 256, nq in {1, 2, 4} and |shift| <= 32");
  int lo = 0, hi = 0;
  for (int k = 0; k < 3; ++k) { lo = std::min(lo, a.shift[k]); hi = std::max(hi, a.shift[k]); }
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(256);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (pdl_enabled()) {   // see the kernel: its weight loads overlap the tail of the previous kernel of the chain
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
#define FWN_FD(NQ, ROWS)                                                                                             \
  {                                                                                                                  \
    const int tiles_per_utt = (a.Ti + ROWS - 1) / ROWS;                                                              \
    const int n_tiles = a.B * tiles_per_utt;                                                                         \
    cfg.gridDim = dim3((unsigned)std::min(n_tiles, num_sms() * 2));                                                  \
    if (fp16) FWN_CUDA(cudaLaunchKernelEx(&cfg, front_direct_kernel<NQ, __half, ROWS>, a, lo, hi, tiles_per_utt, n_tiles));        \
    else FWN_CUDA(cudaLaunchKernelEx(&cfg, front_direct_kernel<NQ, __nv_bfloat16, ROWS>, a, lo, hi, tiles_per_utt, n_tiles));      \
  }
  // big tiles only when they still give every SM several tiles; short inputs keep 64-row tiles (more blocks in flight)
  bool big = (int64_t)a.B * a.Ti >= (int64_t)256 * 4 * num_sms();
  if (const char* e = getenv("FWN_FRONT_TILE")) big = e[0] == 'b';   // "big" / "small": tests run both tile sizes on the same input
  if (a.nq == 1) { if (big) FWN_FD(1, 256) else FWN_FD(1, 64) }
  else if (a.nq == 2) { if (big) FWN_FD(2, 128) else FWN_FD(2, 64) }
  else FWN_FD(4, 64)
#undef FWN_FD
  FWN_LAUNCH_CHECK();
  return 0;
}

// Gather the pass-through half of the flow variable, apply ActNorm (forward direction) and cast to bf16: the A operand of the
// tensor-core front conv.  One thread per (row, 8 output columns): X rows are contiguous so a warp reads whole rows.
template <typename T16>
__global__ void front_pack_kernel(const float* __restrict__ X, int Cx, int nq, int kq, const int* __restrict__ off2log,
                                  const float* __restrict__ an_b, const float* __restrict__ an_s, T16* __restrict__ A0, int64_t rows) {
  const int64_t n = rows * Cx;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / Cx;
    const int o = (int)(i - row * Cx);
    const int q = __ldg(off2log + o);
    if (q >= nq) continue;
    float v = __ldg(X + i);
    if (an_b) v = (v + __ldg(an_b + o)) * __ldg(an_s + o);
    A0[row * kq + q] = from_f<T16>(v);
  }
}
int front_pack(const float* X, int Cx, int nq, int kq, const int* off2log, const float* an_b, const float* an_s, void* A0, int64_t rows,
               bool fp16, cudaStream_t st) {
  if (rows <= 0) return 0;
  if (kq != nq) FWN_CUDA(cudaMemsetAsync(A0, 0, (size_t)rows * kq * 2, st));  // padding columns must be finite (they meet zero weights)
  const int64_t n = rows * Cx;
  int grid = (int)std::min<int64_t>(cdiv(n, 256), (int64_t)num_sms() * 16);
  if (fp16) front_pack_kernel<__half><<<grid, 256, 0, st>>>(X, Cx, nq, kq, off2log, an_b, an_s, reinterpret_cast<__half*>(A0), rows);
  else front_pack_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(X, Cx, nq, kq, off2log, an_b, an_s, reinterpret_cast<__nv_bfloat16*>(A0), rows);
  FWN_LAUNCH_CHECK();
  return 0;
}

}  // namespace fwn

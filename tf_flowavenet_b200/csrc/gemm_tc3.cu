// fp32 parity mode on the tensor cores: implicit GEMM with a 3-way bf16 split ("bf16x3").
//
//   a = a1 + a2 + a3 (bf16 pieces, 24 significand bits),  w = w1 + w2 + w3   =>
//   a.w ~= a1w1 + a1w2 + a2w1 + a2w2 + a1w3 + a3w1          (dropped terms <= 2^-24 relative)
//
// Products of bf16 pairs are exact and tcgen05 accumulates in fp32 (TMEM), so the result carries fp32-GEMM accuracy
// (measured 1e-6 relative on this path's shapes) at 1/6 of the bf16 tensor rate -- still ~10x the CUDA-core engine.
// Activations stay fp32 in HBM; weights are split into three bf16 planes at prepack.
//
// Pipeline per 64-wide K chunk (warp-specialised, persistent grid):
//   warp 0      TMA: the chunk of fp32 activations (two [128 x 32-float] boxes, time-shifted, zero-filled = tf.pad) and the three
//               weight planes of this column tile -> shared memory stage
//   warps 2-5   converter: one row per thread; split the 64 floats into three bf16 pieces and write three K-major, 128B-swizzled
//               [128 x 64] operand tiles (the layout TMA itself would have produced), fence.proxy.async, signal
//   warp 1      MMA: six tcgen05.mma groups (one per split term) into the TMEM accumulator, commit frees stage + operand tiles
//   warps 6-13  epilogue: two groups alternate tiles; tcgen05.ld, then the SAME fp32 epilogue functor as the CUDA-core engine
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "epilogue_f32.cuh"
#include "model.h"
#include "tc_ptx.cuh"

namespace fwn {
namespace tc3 {
using namespace tc;

constexpr int BN3 = 64;
constexpr int NPL = 3;                             // bf16 planes per fp32 value
constexpr int F32_BOX = BM * 32 * 4;               // [128 rows x 32 floats] = 16 KB
constexpr int A32_BYTES = 2 * F32_BOX;             // one 64-wide K chunk of fp32 activations
constexpr int WPL_BYTES = BN3 * BK * 2;            // one weight plane tile [64 x 64] bf16 = 8 KB
constexpr int W_BYTES = NPL * WPL_BYTES;
constexpr int STAGE_BYTES = A32_BYTES + W_BYTES;   // 56 KB
constexpr int NST = 2;                             // TMA stages
constexpr int ABF_BYTES = NPL * A_BYTES;           // three bf16 operand tiles = 48 KB
constexpr int NAB = 2;                             // converted-operand buffers
constexpr int NACC = 4;                            // TMEM accumulator stages of 64 columns
constexpr int THREADS = 64 + 128 + 256;
constexpr size_t SMEM = 1024 + (size_t)NST * STAGE_BYTES + (size_t)NAB * ABF_BYTES + 256;
constexpr int NTERMS = 6;

struct alignas(64) Tc3Args {
  CUtensorMap mapA[4];  // fp32 activations of each K segment: (C, Ti, B), box (32, 128, 1), 128B swizzle, zero OOB fill
  CUtensorMap mapW;     // bf16 weight planes: (Kpad, Npad, 3), box (64, 64, 1)
  int shift[4], nchunk[4], last_ksteps[4], wk0[4];
  int nseg, tiles_per_utt, n_tiles;
  GemmArgs g;
};

template <int EPI>
__global__ void __launch_bounds__(THREADS, 1) tc3_gemm_kernel(const __grid_constant__ Tc3Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint8_t* abf_base = smem + (size_t)NST * STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(abf_base + (size_t)NAB * ABF_BYTES);
  uint64_t* empty_bar = full_bar + NST;
  uint64_t* conv_full = empty_bar + NST;
  uint64_t* conv_empty = conv_full + NAB;
  uint64_t* tmem_full = conv_empty + NAB;
  uint64_t* tmem_empty = tmem_full + NACC;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + NACC);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const GemmArgs& g = a.g;
  const int num_m_tiles = g.B * a.tiles_per_utt;
  const int total = num_m_tiles * a.n_tiles;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < a.nseg; ++s) prefetch_tmap(&a.mapA[s]);
    prefetch_tmap(&a.mapW);
    for (int i = 0; i < NST; ++i) { mbar_init(full_bar + i, 1); mbar_init(empty_bar + i, 1); }
    for (int i = 0; i < NAB; ++i) { mbar_init(conv_full + i, 4); mbar_init(conv_empty + i, 1); }
    for (int i = 0; i < NACC; ++i) { mbar_init(tmem_full + i, 1); mbar_init(tmem_empty + i, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<NACC * BN3>(tmem_ptr);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int m_tile = tile / a.n_tiles, n_tile = tile - m_tile * a.n_tiles;
        const int ub = m_tile / a.tiles_per_utt;
        const int t0 = (m_tile - ub * a.tiles_per_utt) * BM;
        for (int s = 0; s < a.nseg; ++s)
          for (int ch = 0; ch < a.nchunk[s]; ++ch) {
            mbar_wait(empty_bar + st, ph ^ 1);
            uint8_t* sa = stage_base + (size_t)st * STAGE_BYTES;
            mbar_expect_tx(full_bar + st, STAGE_BYTES);
            tma_load_3d(sa, &a.mapA[s], full_bar + st, ch * BK, t0 + a.shift[s], ub);
            tma_load_3d(sa + F32_BOX, &a.mapA[s], full_bar + st, ch * BK + 32, t0 + a.shift[s], ub);
#pragma unroll
            for (int p = 0; p < NPL; ++p)
              tma_load_3d(sa + A32_BYTES + p * WPL_BYTES, &a.mapW, full_bar + st, a.wk0[s] + ch * BK, n_tile * BN3, p);
            if (++st == NST) { st = 0; ph ^= 1; }
          }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc<BN3>();
    const uint64_t desc_hi = make_smem_desc(0);
    const uint32_t stage0 = smem_u32(stage_base), abf0 = smem_u32(abf_base);
    int st = 0, ab = 0, as = 0;
    uint32_t ph = 0, abph = 0, aph = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      mbar_wait(tmem_empty + as, aph ^ 1);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN3);
      uint32_t accumulate = 0;
      for (int s = 0; s < a.nseg; ++s) {
        const int nch = a.nchunk[s];
        const uint32_t last_ks = (uint32_t)a.last_ksteps[s];
        for (int ch = 0; ch < nch; ++ch) {
          mbar_wait(full_bar + st, ph);       // weight planes of this stage (and the fp32 chunk) have landed
          mbar_wait(conv_full + ab, abph);    // the converter has written the three bf16 operand tiles
          tcgen05_fence_after();
          const uint32_t ksteps = (ch == nch - 1) ? last_ks : (uint32_t)(BK / UMMA_K);
          const uint32_t sa = abf0 + (uint32_t)ab * ABF_BYTES;
          const uint32_t sw = stage0 + (uint32_t)st * STAGE_BYTES + A32_BYTES;
          // the six split terms (pa, pw), most significant first
#pragma unroll
          for (int tm = 0; tm < NTERMS; ++tm) {
            const int pa = (tm == 0 || tm == 1 || tm == 4) ? 0 : (tm == 5 ? 2 : 1);
            const int pw = (tm == 0 || tm == 2 || tm == 5) ? 0 : (tm == 4 ? 2 : 1);
            const uint64_t adesc = desc_hi | (uint64_t)(((sa + pa * A_BYTES) >> 4) & 0x3FFF);
            const uint64_t bdesc = desc_hi | (uint64_t)(((sw + pw * WPL_BYTES) >> 4) & 0x3FFF);
            umma_chunk(tmem_d, adesc, bdesc, idesc, accumulate, ksteps);
            accumulate = 1;
          }
          umma_commit_elect<false>(smem_u32(empty_bar + st));     // stage (weights) may be refilled
          umma_commit_elect<false>(smem_u32(conv_empty + ab));    // operand tiles may be overwritten
          if (++st == NST) { st = 0; ph ^= 1; }
          if (++ab == NAB) { ab = 0; abph ^= 1; }
        }
      }
      umma_commit_elect<false>(smem_u32(tmem_full + as));
      if (++as == NACC) { as = 0; aph ^= 1; }
    }
  } else if (warp < 6) {
    // ===================== converter (warps 2..5): fp32 chunk -> three bf16 operand tiles =====================
    const int r = (warp - 2) * 32 + lane;   // the row this thread converts
    const uint32_t stage0 = smem_u32(stage_base), abf0 = smem_u32(abf_base);
    int st = 0, ab = 0;
    uint32_t ph = 0, abph = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      for (int s = 0; s < a.nseg; ++s)
        for (int ch = 0; ch < a.nchunk[s]; ++ch) {
          mbar_wait(full_bar + st, ph);
          mbar_wait(conv_empty + ab, abph ^ 1);
          const uint32_t src = stage0 + (uint32_t)st * STAGE_BYTES + (uint32_t)r * 128;
          const uint32_t dst = abf0 + (uint32_t)ab * ABF_BYTES + (uint32_t)r * 128;
          const uint32_t sw = (uint32_t)(r & 7);
#pragma unroll
          for (int c8 = 0; c8 < 8; ++c8) {        // 8 output chunks of 8 bf16 (16 bytes) per plane = 64 K values
            // the 8 floats of this chunk: float4 chunks 2*c8, 2*c8+1 of the 64-float row = box (c8 >> 2), chunk-in-box ((2*c8) & 7) + {0,1}
            const uint32_t box = (uint32_t)(c8 >> 2) * F32_BOX;
            const uint32_t q0 = (uint32_t)((2 * c8) & 7), q1 = q0 + 1;
            const uint4 u0 = lds128(src + box + ((q0 ^ sw) << 4));
            const uint4 u1 = lds128(src + box + ((q1 ^ sw) << 4));
            const float f[8] = {__uint_as_float(u0.x), __uint_as_float(u0.y), __uint_as_float(u0.z), __uint_as_float(u0.w),
                                __uint_as_float(u1.x), __uint_as_float(u1.y), __uint_as_float(u1.z), __uint_as_float(u1.w)};
            uint32_t p1[4], p2[4], p3[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float h[2][3];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const float x = f[2 * j + e];
                const float a1 = __bfloat162float(__float2bfloat16_rn(x));
                const float r1 = x - a1;
                const float a2 = __bfloat162float(__float2bfloat16_rn(r1));
                const float a3 = r1 - a2;   // rounded to bf16 by the pack below
                h[e][0] = a1; h[e][1] = a2; h[e][2] = a3;
              }
              p1[j] = pack_bf16(h[0][0], h[1][0]);
              p2[j] = pack_bf16(h[0][1], h[1][1]);
              p3[j] = pack_bf16(h[0][2], h[1][2]);
            }
            const uint32_t off = ((uint32_t)c8 ^ sw) << 4;
            sts128(dst + off, make_uint4(p1[0], p1[1], p1[2], p1[3]));
            sts128(dst + A_BYTES + off, make_uint4(p2[0], p2[1], p2[2], p2[3]));
            sts128(dst + 2 * A_BYTES + off, make_uint4(p3[0], p3[1], p3[2], p3[3]));
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to tcgen05.mma
          __syncwarp();
          if (lane == 0) mbar_arrive(conv_full + ab);
          if (++st == NST) { st = 0; ph ^= 1; }
          if (++ab == NAB) { ab = 0; abph ^= 1; }
        }
    }
  } else {
    // ===================== epilogue (warps 6..13): two groups of four warps alternate tiles =====================
    const int lg = warp & 3;
    const int grp = (warp - 6) >> 2;
    const int r = lg * 32 + lane;
    double ls_sum = 0.0;
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      if ((it & 1) != grp) continue;
      const int m_tile = tile / a.n_tiles, n_tile = tile - m_tile * a.n_tiles;
      const int ub = m_tile / a.tiles_per_utt;
      const int t = (m_tile - ub * a.tiles_per_utt) * BM + r;
      const bool row_ok = t < g.Ti;
      const int64_t row = (int64_t)ub * g.Ti + t;
      const int as = it % NACC;
      mbar_wait(tmem_full + as, (uint32_t)(it / NACC) & 1);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * BN3);
      uint32_t v[BN3];
#pragma unroll
      for (int j = 0; j < BN3; j += 16) tmem_ld_x16(taddr + j, v + j);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty + as);   // accumulator is in registers: release the TMEM stage right away
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < BN3; j += 4) {
          const int col = n_tile * BN3 + j;
          if (col < g.N) {
            const float acc[4] = {__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])};
            Epilogue<EPI>::apply(g, row, t, col, acc, ls_sum);
          }
        }
      }
    }
    if (EPI == EPI_AFFINE && !g.e.reverse && g.e.logdet_acc) {
      ls_sum = warp_sum(ls_sum);
      if (lane == 0 && ls_sum != 0.0) atomicAdd(g.e.logdet_acc, ls_sum);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<NACC * BN3>(tmem_base);
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
  }
  return fn;
}

template <int EPI>
static int launch(const Tc3Args& a, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    FWN_CUDA(cudaFuncSetAttribute(tc3_gemm_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    configured = true;
  }
  const int total = a.g.B * a.tiles_per_utt * a.n_tiles;
  tc3_gemm_kernel<EPI><<<std::min(total, num_sms()), THREADS, SMEM, st>>>(a);
  FWN_LAUNCH_CHECK();
  return 0;
}

}  // namespace tc3

// Can this GEMM run on the split engine?  (TMA needs 16-byte aligned fp32 rows; the per-op weight-norm column scale is not supported.)
bool tc3_supported(const GemmArgs& g) {
  if (g.e.colscale) return false;
  for (int s = 0; s < g.nseg; ++s)
    if ((g.seg[s].lda & 3) || (reinterpret_cast<uintptr_t>(g.seg[s].A) & 15) || g.seg[s].K <= 0 || (g.seg[s].koff & 7)) return false;
  return true;
}

// w3: bf16 planes [3][Npad][Kpad] of the [Ktot][N] weight matrix (K-major), Kpad % 64 == 0, Npad % 16 == 0
int tc3_gemm(const GemmArgs& g, EpiKind kind, const void* w3, int Kpad, int Npad, cudaStream_t st) {
  if (g.B <= 0 || g.Ti <= 0 || g.N <= 0) return 0;
  tc3::EncodeFn enc = tc3::get_encode();
  FWN_CHECK(enc, "cuTensorMapEncodeTiled unavailable (driver too old?)");
  tc3::Tc3Args a;
  memset(&a, 0, sizeof(a));
  a.g = g;
  a.nseg = g.nseg;
  for (int s = 0; s < g.nseg; ++s) {
    const Seg& sg = g.seg[s];
    cuuint64_t dims[3] = {(cuuint64_t)sg.K, (cuuint64_t)g.Ti, (cuuint64_t)g.B};
    cuuint64_t strides[2] = {(cuuint64_t)sg.lda * 4, (cuuint64_t)sg.lda * 4 * (cuuint64_t)g.Ti};
    cuuint32_t box[3] = {32, tc::BM, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&a.mapA[s], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(sg.A), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FWN_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(fp32 activations K=%d Ti=%d B=%d lda=%lld) failed: %d", sg.K, g.Ti, g.B,
              (long long)sg.lda, (int)r);
    a.shift[s] = sg.shift;
    const int K16 = (sg.K + 15) / 16 * 16;
    a.nchunk[s] = (K16 + tc::BK - 1) / tc::BK;
    a.last_ksteps[s] = (K16 - (a.nchunk[s] - 1) * tc::BK) / tc::UMMA_K;
    a.wk0[s] = sg.koff;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)Kpad, (cuuint64_t)Npad, (cuuint64_t)tc3::NPL};
    cuuint64_t strides[2] = {(cuuint64_t)Kpad * 2, (cuuint64_t)Kpad * 2 * (cuuint64_t)Npad};
    cuuint32_t box[3] = {tc::BK, tc3::BN3, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&a.mapW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w3), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FWN_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weight planes N=%d K=%d) failed: %d", Npad, Kpad, (int)r);
  }
  a.tiles_per_utt = (g.Ti + tc::BM - 1) / tc::BM;
  a.n_tiles = (g.N + tc3::BN3 - 1) / tc3::BN3;
  if (getenv("FWN_TC3_TRACE")) {
    fprintf(stderr, "tc3 kind=%d B=%d Ti=%d N=%d Kpad=%d Npad=%d nseg=%d:", (int)kind, g.B, g.Ti, g.N, Kpad, Npad, g.nseg);
    for (int s = 0; s < g.nseg; ++s)
      fprintf(stderr, " [K=%d lda=%lld sh=%d koff=%d nch=%d lk=%d]", g.seg[s].K, (long long)g.seg[s].lda, g.seg[s].shift, g.seg[s].koff,
              a.nchunk[s], a.last_ksteps[s]);
    fprintf(stderr, "\n");
  }
  switch (kind) {
    case EPI_PLAIN: return tc3::launch<EPI_PLAIN>(a, st);
    case EPI_GATE: return tc3::launch<EPI_GATE>(a, st);
    case EPI_RES_SKIP: return tc3::launch<EPI_RES_SKIP>(a, st);
    case EPI_AFFINE: return tc3::launch<EPI_AFFINE>(a, st);
    case EPI_LINEAR: return tc3::launch<EPI_LINEAR>(a, st);
    case EPI_GATE_BWD: return tc3::launch<EPI_GATE_BWD>(a, st);
  }
  return 1;
}

}  // namespace fwn

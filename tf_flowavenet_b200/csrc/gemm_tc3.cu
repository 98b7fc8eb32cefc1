// fp32 parity mode on the tensor cores: implicit GEMM with a 3-way bf16 split ("bf16x3").
//
//   a = a1 + a2 + a3 (bf16 pieces, round-to-nearest),  w = w1 + w2 + w3   =>
//   a.w ~= a1w1 + a1w2 + a2w1 [+ a2w2 + a1w3 + a3w1]        3 terms: ~2^-16..2^-18 per product; 6 terms: 2^-24
//
// Products of bf16 pairs are exact and tcgen05 accumulates in fp32 (TMEM).  Measured against the CUDA-core fp32 GEMM: 5e-6
// relative with 6 terms on K = 848 (the accumulator truncates where FFMA rounds); ~4.5x the CUDA-core engine's throughput.
// Activations stay fp32 in HBM; weights are split into three bf16 planes at prepack (model.cu) / re-pack (train_kernels.cu).
//
// Pipeline per 64-wide K chunk (warp-specialised, persistent grid, column tile BN = 128 or 64):
//   warp 0       TMA: the chunk of fp32 activations (two [128 x 32-float] boxes, time-shifted, zero-filled = tf.pad) and the two or
//                three weight planes of this column tile -> shared memory stage (2 stages)
//   warps 2-9    converters: thread = (row, half): 8 LDS.128, split, 12 STS.128 into three K-major, 128B-swizzled [128 x 64]
//                operand tiles (the layout TMA itself would have produced), fence.proxy.async, signal.  Loads + split of chunk
//                i+1 overlap the MMAs of chunk i (single operand buffer); only the stores wait.
//   warp 1       MMA: 3 or 6 groups of tcgen05.mma (one per split term) into the TMEM accumulator; commits free stage + operands
//   warps 10-17  epilogue: two groups alternate tiles over 4 TMEM stages; tcgen05.ld, per-warp shared-memory transpose (coalesced
//                global accesses), then the SAME fp32 epilogue functor as the CUDA-core engine (epilogue_f32.cuh)
// GEMMs without a time shift run on one flat row axis (B*Ti rows), so short utterances still fill 128-row tiles.
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

constexpr int NPL = 3;                             // bf16 planes per fp32 value
constexpr int F32_BOX = BM * 32 * 4;               // [128 rows x 32 floats] = 16 KB
constexpr int A32_BYTES = 2 * F32_BOX;             // one 64-wide K chunk of fp32 activations
constexpr int NST = 2;                             // TMA stages
constexpr int ABF_BYTES = NPL * A_BYTES;           // three bf16 operand tiles = 48 KB (single buffer, see the converter)
constexpr int CONV_WARPS = 8, EPI_WARPS = 8;
constexpr int THREADS = 64 + 32 * CONV_WARPS + 32 * EPI_WARPS;
constexpr int NTERMS = 6;
constexpr int EPI_STG_BYTES = EPI_WARPS * 32 * 16 * 4;   // per epilogue warp: a [32 rows x 16 cols] fp32 transpose tile (16 KB in all)

template <int BN>
struct Cfg3 {
  static constexpr int WPL_BYTES = BN * BK * 2;            // one weight plane tile [BN x 64] bf16
  static constexpr int W_BYTES = NPL * WPL_BYTES;
  static constexpr int STAGE_BYTES = A32_BYTES + W_BYTES;  // 80 KB (BN 128) / 56 KB (BN 64)
  static constexpr int NACC = 4;                           // TMEM accumulator stages
  static constexpr size_t SMEM = 1024 + (size_t)NST * STAGE_BYTES + ABF_BYTES + EPI_STG_BYTES + 256;
};

struct alignas(64) Tc3Args {
  CUtensorMap mapA[4];  // fp32 activations of each K segment: (C, Ti, B), box (32, 128, 1), 128B swizzle, zero OOB fill
  CUtensorMap mapW;     // bf16 weight planes: (Kpad, Npad, 3), box (64, BN, 1)
  int shift[4], nchunk[4], last_ksteps[4], wk0[4];
  int nseg, tiles_per_utt, n_tiles;
  int nterms;           // 6: all products down to 2^-24; 3: a1w1 + a1w2 + a2w1 (~2^-16 per product)
  GemmArgs g;
};

template <int EPI, int BN>
__global__ void __launch_bounds__(THREADS, 1) tc3_gemm_kernel(const __grid_constant__ Tc3Args a) {
  using C = Cfg3<BN>;
  constexpr int NACC = C::NACC;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint8_t* abf_base = smem + (size_t)NST * C::STAGE_BYTES;
  uint8_t* epi_stg = abf_base + ABF_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_stg + EPI_STG_BYTES);
  uint64_t* empty_bar = full_bar + NST;
  uint64_t* conv_full = empty_bar + NST;
  uint64_t* conv_empty = conv_full + 1;
  uint64_t* tmem_full = conv_empty + 1;
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
    mbar_init(conv_full, CONV_WARPS);
    mbar_init(conv_empty, 1);
    for (int i = 0; i < NACC; ++i) { mbar_init(tmem_full + i, 1); mbar_init(tmem_empty + i, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<NACC * BN>(tmem_ptr);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_launch_dependents();   // programmatic dependent launch: see tc_ptx.cuh

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      pdl_wait();
      const int nplanes = a.nterms > 3 ? 3 : 2;   // the third weight plane only feeds the 2^-16 terms
      int st = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int m_tile = tile / a.n_tiles, n_tile = tile - m_tile * a.n_tiles;
        const int ub = m_tile / a.tiles_per_utt;
        const int t0 = (m_tile - ub * a.tiles_per_utt) * BM;
        for (int s = 0; s < a.nseg; ++s)
          for (int ch = 0; ch < a.nchunk[s]; ++ch) {
            mbar_wait(empty_bar + st, ph ^ 1);
            uint8_t* sa = stage_base + (size_t)st * C::STAGE_BYTES;
            mbar_expect_tx(full_bar + st, A32_BYTES + nplanes * C::WPL_BYTES);
            tma_load_3d(sa, &a.mapA[s], full_bar + st, ch * BK, t0 + a.shift[s], ub);
            tma_load_3d(sa + F32_BOX, &a.mapA[s], full_bar + st, ch * BK + 32, t0 + a.shift[s], ub);
            for (int p = 0; p < nplanes; ++p)
              tma_load_3d(sa + A32_BYTES + p * C::WPL_BYTES, &a.mapW, full_bar + st, a.wk0[s] + ch * BK, n_tile * BN, p);
            if (++st == NST) { st = 0; ph ^= 1; }
          }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc<BN>();
    const uint64_t desc_hi = make_smem_desc(0);
    const uint32_t stage0 = smem_u32(stage_base), sa = smem_u32(abf_base);
    const int nterms = a.nterms;
    int st = 0, as = 0;
    uint32_t ph = 0, cph = 0, aph = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      mbar_wait(tmem_empty + as, aph ^ 1);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
      uint32_t accumulate = 0;
      for (int s = 0; s < a.nseg; ++s) {
        const int nch = a.nchunk[s];
        const uint32_t last_ks = (uint32_t)a.last_ksteps[s];
        for (int ch = 0; ch < nch; ++ch) {
          mbar_wait(full_bar + st, ph);     // weight planes of this stage have landed
          mbar_wait(conv_full, cph);        // the converters have written the three bf16 operand tiles
          cph ^= 1;
          tcgen05_fence_after();
          const uint32_t ksteps = (ch == nch - 1) ? last_ks : (uint32_t)(BK / UMMA_K);
          const uint32_t sw = stage0 + (uint32_t)st * C::STAGE_BYTES + A32_BYTES;
          // the six split terms (pa, pw), most significant first
          for (int tm = 0; tm < nterms; ++tm) {
            const int pa = (0x201100 >> (4 * tm)) & 0xF;   // (pa, pw) = (0,0) (0,1) (1,0) | (1,1) (0,2) (2,0)
            const int pw = (0x021010 >> (4 * tm)) & 0xF;
            const uint64_t adesc = desc_hi | (uint64_t)(((sa + pa * A_BYTES) >> 4) & 0x3FFF);
            const uint64_t bdesc = desc_hi | (uint64_t)(((sw + pw * C::WPL_BYTES) >> 4) & 0x3FFF);
            umma_chunk(tmem_d, adesc, bdesc, idesc, accumulate, ksteps);
            accumulate = 1;
          }
          umma_commit_elect<false>(smem_u32(empty_bar + st));   // stage may be refilled
          umma_commit_elect<false>(smem_u32(conv_empty));       // operand tiles may be overwritten
          if (++st == NST) { st = 0; ph ^= 1; }
        }
      }
      umma_commit_elect<false>(smem_u32(tmem_full + as));
      if (++as == NACC) { as = 0; aph ^= 1; }
    }
  } else if (warp < 2 + CONV_WARPS) {
    // ===================== converters: fp32 chunk -> three bf16 operand tiles =====================
    // Thread = (row, half): 32 floats = one TMA box row.  The loads and the split of chunk i+1 run while the MMAs of chunk i
    // still read the (single) operand buffer; only the stores wait for them.
    const int j = threadIdx.x - 64;
    const int r = j & 127, half = j >> 7;
    const uint32_t sw = (uint32_t)(r & 7);
    const uint32_t src_off = (uint32_t)half * F32_BOX + (uint32_t)r * 128;
    const uint32_t dst = smem_u32(abf_base) + (uint32_t)r * 128;
    const uint32_t stage0 = smem_u32(stage_base);
    int st = 0;
    uint32_t ph = 0, ce = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      for (int s = 0; s < a.nseg; ++s)
        for (int ch = 0; ch < a.nchunk[s]; ++ch) {
          mbar_wait(full_bar + st, ph);
          const uint32_t src = stage0 + (uint32_t)st * C::STAGE_BYTES + src_off;
          uint32_t p1[16], p2[16], p3[16];
#pragma unroll
          for (int q = 0; q < 8; ++q) {     // 8 float4 chunks of this half row
            const uint4 u = lds128(src + (((uint32_t)q ^ sw) << 4));
            split3_pair(__uint_as_float(u.x), __uint_as_float(u.y), p1[2 * q], p2[2 * q], p3[2 * q]);
            split3_pair(__uint_as_float(u.z), __uint_as_float(u.w), p1[2 * q + 1], p2[2 * q + 1], p3[2 * q + 1]);
          }
          mbar_wait(conv_empty, ce ^ 1);    // MMAs of the previous chunk are done with the operand tiles
          ce ^= 1;
#pragma unroll
          for (int c = 0; c < 4; ++c) {     // 4 output chunks of 8 bf16 per plane: K columns 32*half + 8c ..
            const uint32_t off = ((uint32_t)(4 * half + c) ^ sw) << 4;
            sts128(dst + off, make_uint4(p1[4 * c], p1[4 * c + 1], p1[4 * c + 2], p1[4 * c + 3]));
            sts128(dst + A_BYTES + off, make_uint4(p2[4 * c], p2[4 * c + 1], p2[4 * c + 2], p2[4 * c + 3]));
            if (a.nterms > 3) sts128(dst + 2 * A_BYTES + off, make_uint4(p3[4 * c], p3[4 * c + 1], p3[4 * c + 2], p3[4 * c + 3]));
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to tcgen05.mma
          __syncwarp();
          if (lane == 0) mbar_arrive(conv_full);
          if (++st == NST) { st = 0; ph ^= 1; }
        }
    }
  } else {
    // ===================== epilogue: two groups of four warps alternate tiles =====================
    // tcgen05.ld hands every thread one ROW of the accumulator; touching global memory row-per-thread costs a cache line per lane
    // and instruction.  Each warp therefore transposes 16-column slabs through a private swizzled shared-memory tile, after which
    // a lane owns 4 consecutive columns of 8 different rows per step: 64 contiguous bytes per row, 8 rows per memory instruction.
    const int lg = warp & 3;
    const int ew = warp - (2 + CONV_WARPS);
    const int grp = ew >> 2;
    const uint32_t stg = smem_u32(epi_stg) + (uint32_t)ew * (32 * 16 * 4);
    const uint32_t wsw = ((uint32_t)lane >> 1) & 3;                       // write side: my row = lane
    const int rsub = lane >> 2, c4 = lane & 3;                           // read side: row 8k + rsub, column chunk c4
    double ls_sum = 0.0;
    int it = 0;
    pdl_wait();
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      if ((it & 1) != grp) continue;
      const int m_tile = tile / a.n_tiles, n_tile = tile - m_tile * a.n_tiles;
      const int ub = m_tile / a.tiles_per_utt;
      const int t_warp = (m_tile - ub * a.tiles_per_utt) * BM + lg * 32;  // first row of this warp's 32-row slab
      const int as = it % NACC;
      mbar_wait(tmem_full + as, (uint32_t)(it / NACC) & 1);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * BN);
#pragma unroll 1
      for (int cc = 0; cc < BN; cc += 32) {
        uint32_t v[32];
        tmem_ld_x16(taddr + cc, v);
        tmem_ld_x16(taddr + cc + 16, v + 16);
        tmem_ld_wait();
        if (cc + 32 == BN) {   // accumulator fully in registers: release the TMEM stage
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty + as);
        }
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
          __syncwarp();        // the previous slab has been read
#pragma unroll
          for (int q = 0; q < 4; ++q)
            sts128(stg + (uint32_t)lane * 64 + ((((uint32_t)q) ^ wsw) << 4),
                   make_uint4(v[16 * hb + 4 * q], v[16 * hb + 4 * q + 1], v[16 * hb + 4 * q + 2], v[16 * hb + 4 * q + 3]));
          __syncwarp();
          const int col = n_tile * BN + cc + 16 * hb + 4 * c4;
          float acc[4][4], pre[4][4];
          bool ok[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int rl = 8 * k + rsub;
            const uint4 u = lds128(stg + (uint32_t)rl * 64 + ((((uint32_t)c4) ^ (((uint32_t)rl >> 1) & 3)) << 4));
            acc[k][0] = __uint_as_float(u.x); acc[k][1] = __uint_as_float(u.y); acc[k][2] = __uint_as_float(u.z); acc[k][3] = __uint_as_float(u.w);
            ok[k] = (t_warp + rl < g.Ti) && col < g.N;
          }
          if (EPI == EPI_LINEAR || EPI == EPI_GATE) {   // all addends first (see Epilogue::prefetch)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (ok[k]) Epilogue<EPI>::prefetch(g, (int64_t)ub * g.Ti + t_warp + 8 * k + rsub, col, pre[k]);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int tt = t_warp + 8 * k + rsub;
            if (ok[k]) Epilogue<EPI>::apply(g, (int64_t)ub * g.Ti + tt, tt, col, acc[k], ls_sum, ((EPI == EPI_LINEAR || EPI == EPI_GATE) && g.e.in0) ? pre[k] : nullptr);
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
    tmem_dealloc<NACC * BN>(tmem_base);
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

template <int EPI, int BN>
static int launch_bn(const Tc3Args& a, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    FWN_CUDA(cudaFuncSetAttribute(tc3_gemm_kernel<EPI, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg3<BN>::SMEM));
    configured = true;
  }
  const int total = a.g.B * a.tiles_per_utt * a.n_tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)std::min(total, num_sms()));
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = Cfg3<BN>::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  FWN_CUDA(cudaLaunchKernelEx(&cfg, tc3_gemm_kernel<EPI, BN>, a));
  FWN_LAUNCH_CHECK();
  return 0;
}
template <int EPI>
static int launch(const Tc3Args& a, int bn, cudaStream_t st) {
  return bn == 128 ? launch_bn<EPI, 128>(a, st) : launch_bn<EPI, 64>(a, st);
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
int tc3_gemm(const GemmArgs& g, EpiKind kind, const void* w3, int Kpad, int Npad, int nterms, cudaStream_t st) {
  if (g.B <= 0 || g.Ti <= 0 || g.N <= 0) return 0;
  tc3::EncodeFn enc = tc3::get_encode();
  FWN_CHECK(enc, "cuTensorMapEncodeTiled unavailable (driver too old?)");
  const int bn = g.N > 64 ? 128 : 64;   // column tile: 128 halves the conversion work per FLOP
  tc3::Tc3Args a;
  memset(&a, 0, sizeof(a));
  a.g = g;
  bool shifted = false;
  for (int s = 0; s < g.nseg; ++s) shifted = shifted || g.seg[s].shift != 0;
  if (!shifted && (int64_t)g.B * g.Ti < (int64_t(1) << 31)) {  // 1x1 convs: utterance boundaries do not matter -> one flat row axis, full tiles
    a.g.Ti = g.B * g.Ti;
    a.g.B = 1;
  }
  a.nterms = nterms == 3 ? 3 : 6;
  a.nseg = g.nseg;
  for (int s = 0; s < g.nseg; ++s) {
    const Seg& sg = g.seg[s];
    cuuint64_t dims[3] = {(cuuint64_t)sg.K, (cuuint64_t)a.g.Ti, (cuuint64_t)a.g.B};
    cuuint64_t strides[2] = {(cuuint64_t)sg.lda * 4, (cuuint64_t)sg.lda * 4 * (cuuint64_t)a.g.Ti};
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
    cuuint32_t box[3] = {tc::BK, (cuuint32_t)bn, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&a.mapW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w3), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FWN_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weight planes N=%d K=%d) failed: %d", Npad, Kpad, (int)r);
  }
  a.tiles_per_utt = (a.g.Ti + tc::BM - 1) / tc::BM;
  a.n_tiles = (g.N + bn - 1) / bn;
  if (getenv("FWN_TC3_TRACE")) {
    fprintf(stderr, "tc3 kind=%d B=%d Ti=%d N=%d Kpad=%d Npad=%d nseg=%d:", (int)kind, g.B, g.Ti, g.N, Kpad, Npad, g.nseg);
    for (int s = 0; s < g.nseg; ++s)
      fprintf(stderr, " [K=%d lda=%lld sh=%d koff=%d nch=%d lk=%d]", g.seg[s].K, (long long)g.seg[s].lda, g.seg[s].shift, g.seg[s].koff,
              a.nchunk[s], a.last_ksteps[s]);
    fprintf(stderr, "\n");
  }
  switch (kind) {
    case EPI_PLAIN: return tc3::launch<EPI_PLAIN>(a, bn, st);
    case EPI_GATE: return tc3::launch<EPI_GATE>(a, bn, st);
    case EPI_RES_SKIP: return tc3::launch<EPI_RES_SKIP>(a, bn, st);
    case EPI_AFFINE: return tc3::launch<EPI_AFFINE>(a, bn, st);
    case EPI_LINEAR: return tc3::launch<EPI_LINEAR>(a, bn, st);
    case EPI_GATE_BWD: return tc3::launch<EPI_GATE_BWD>(a, bn, st);
  }
  return 1;
}

}  // namespace fwn

// Weight gradients on the tensor cores in the fp32 parity mode (3-way bf16 split of BOTH operands, fp32 accumulation in TMEM):
//   dW[koff + k, n] += sum_{b,t} A[b, t + shift, k] * dY[b, t, n]
// The reduction runs over time, so both operands are needed "transposed" (reduction index contiguous).  TMA brings fp32 tiles
// [64 time steps x 128 channels] of A and dY into shared memory; eight converter warps read them COLUMN-wise (one channel per
// thread, conflict-free), split each value into three bf16 pieces (round-to-nearest, unbiased) and write the
// K-major, 128B-swizzled operand tiles [128 channels x 64 time steps] x 3 planes that tcgen05.mma consumes.  Six MMA groups per
// chunk (a1y1, a1y2, a2y1, a2y2, a1y3, a3y1).  A CTA owns one [128 k x 128 n] tile of dW and a slab of whole 64-step chunks of
// one utterance; partial tiles are combined with fp32 atomics (coalesced through a shared-memory transpose).
// The converter loads + splits the next chunk into registers while the MMAs of the current chunk run; only its shared-memory
// stores wait for them (single operand buffer: shared memory holds 2 x 64 KB of fp32 staging + 96 KB of bf16 planes).
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "train.h"

namespace fwn {
namespace wg3 {
using namespace tc;

constexpr int BNW = 128;                       // n tile
constexpr int MC = 64;                         // time steps per chunk
constexpr int BOX_BYTES = MC * 128;            // one TMA box [64 rows x 32 floats] = 8 KB
constexpr int STG_A = 4 * BOX_BYTES, STG_Y = 4 * BOX_BYTES, STAGE_BYTES = STG_A + STG_Y;   // 64 KB
constexpr int NST = 2;
constexpr int PL_A = 3 * A_BYTES, PL_Y = 3 * A_BYTES;                                      // 48 KB each
constexpr int CONV_WARPS = 16;                 // thread = (channel, half of the 64 time steps)
constexpr int THREADS = 64 + CONV_WARPS * 32 + 128;
constexpr size_t SMEM = 1024 + (size_t)NST * STAGE_BYTES + PL_A + PL_Y + 256;
constexpr int TSTRIDE = 129;                   // fp32 transpose buffer row pitch (floats)
static_assert((size_t)128 * TSTRIDE * 4 <= (size_t)NST * STAGE_BYTES, "transpose buffer must fit in the staging area");

struct alignas(64) Wg3Args {
  CUtensorMap mapA[4];   // fp32 activations per K segment: (K, Ti, B), box (32, 64, 1), 128B swizzle
  CUtensorMap mapY[2];   // fp32 output gradients per column segment: (ncols, Ti, B), same box
  int shift[4], K[4], koff[4], ktiles[4];
  int nseg, n0cols, N;
  float* dW;
  int64_t ldw;
  float* dbias;          // nullable: column sums of dY (bias gradient), accumulated by the CTAs of the first K tile
  int B, Ti, chunks_per_utt, slabs;   // slab = contiguous range of (utterance, chunk) pairs
  int nterms;   // 6 or 3, as in gemm_tc3.cu
};

__global__ void __launch_bounds__(THREADS, 1) wgrad_tc3_kernel(const __grid_constant__ Wg3Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint8_t* planes = smem + (size_t)NST * STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(planes + PL_A + PL_Y);
  uint64_t* empty_bar = full_bar + NST;
  uint64_t* conv_full = empty_bar + NST;
  uint64_t* planes_free = conv_full + 1;
  uint64_t* tmem_full = planes_free + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tile / slab of this CTA
  int kt = blockIdx.x, sidx = 0;
  while (sidx < a.nseg - 1 && kt >= a.ktiles[sidx]) { kt -= a.ktiles[sidx]; ++sidx; }
  const int k0 = kt * 128, n0 = blockIdx.y * BNW;
  const int64_t total_chunks = (int64_t)a.B * a.chunks_per_utt;
  const int c_begin = (int)(total_chunks * blockIdx.z / a.slabs), c_end = (int)(total_chunks * (blockIdx.z + 1) / a.slabs);
  const int nchunks = c_end - c_begin;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&a.mapA[sidx]);
    prefetch_tmap(&a.mapY[n0 < a.n0cols ? 0 : 1]);
    for (int i = 0; i < NST; ++i) { mbar_init(full_bar + i, 1); mbar_init(empty_bar + i, CONV_WARPS); }
    mbar_init(conv_full, CONV_WARPS);
    mbar_init(planes_free, 1);
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<BNW>(tmem_ptr);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (nchunks > 0) {
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (lane == 0) {
        const CUtensorMap* my = &a.mapY[n0 < a.n0cols ? 0 : 1];
        const int ny = n0 < a.n0cols ? n0 : n0 - a.n0cols;
        int st = 0;
        uint32_t ph = 0;
        for (int ch = c_begin; ch < c_end; ++ch) {
          const int ub = ch / a.chunks_per_utt;
          const int t = (ch - ub * a.chunks_per_utt) * MC;
          mbar_wait(empty_bar + st, ph ^ 1);
          uint8_t* sa = stage_base + (size_t)st * STAGE_BYTES;
          mbar_expect_tx(full_bar + st, STAGE_BYTES);
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_3d(sa + j * BOX_BYTES, &a.mapA[sidx], full_bar + st, k0 + 32 * j, t + a.shift[sidx], ub);
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_load_3d(sa + STG_A + j * BOX_BYTES, my, full_bar + st, ny + 32 * j, t, ub);
          if (++st == NST) { st = 0; ph ^= 1; }
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = make_idesc<BNW>();
      const uint64_t desc_hi = make_smem_desc(0);
      const uint32_t pa0 = smem_u32(planes), py0 = pa0 + PL_A;
      uint32_t ph = 0, accumulate = 0;
      for (int ch = 0; ch < nchunks; ++ch) {
        mbar_wait(conv_full, ph);
        ph ^= 1;
        tcgen05_fence_after();
        for (int tm = 0; tm < a.nterms; ++tm) {
          const int pa = (0x201100 >> (4 * tm)) & 0xF;
          const int py = (0x021010 >> (4 * tm)) & 0xF;
          const uint64_t adesc = desc_hi | (uint64_t)(((pa0 + pa * A_BYTES) >> 4) & 0x3FFF);
          const uint64_t bdesc = desc_hi | (uint64_t)(((py0 + py * A_BYTES) >> 4) & 0x3FFF);
          umma_chunk(tmem_base, adesc, bdesc, idesc, accumulate, MC / UMMA_K);
          accumulate = 1;
        }
        umma_commit_elect<false>(smem_u32(planes_free));
      }
      umma_commit_elect<false>(smem_u32(tmem_full));
    } else if (warp < 2 + CONV_WARPS) {
      // ===================== converters: fp32 column -> three bf16 rows =====================
      const int j = threadIdx.x - 64;          // 0..511
      const bool isA = (j & 255) < 128;
      const int row = j & 127;                 // channel (k for A, n for dY) = row of the operand tile
      const int mh = j >> 8;                   // which half of the chunk's 64 time steps
      const uint32_t src_off = (isA ? 0u : (uint32_t)STG_A) + (uint32_t)(row >> 5) * BOX_BYTES + (uint32_t)(row & 3) * 4;
      const uint32_t c4 = (uint32_t)(row & 31) >> 2;
      const uint32_t dst0 = smem_u32(planes) + (isA ? 0u : (uint32_t)PL_A) + (uint32_t)row * 128;
      const uint32_t sw = (uint32_t)(row & 7);
      const uint32_t stage0 = smem_u32(stage_base);
      double bsum = 0.0;  // dY converters: running column sum = bias gradient (double: it runs over every row of the slab)
      int st = 0;
      uint32_t ph = 0, pf = 0;
      for (int ch = 0; ch < nchunks; ++ch) {
        mbar_wait(full_bar + st, ph);
        const uint32_t src = stage0 + (uint32_t)st * STAGE_BYTES + src_off;
        uint32_t p1[16], p2[16], p3[16];
#pragma unroll
        for (int m = 0; m < MC / 2; m += 2) {
          float x[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const uint32_t mm = (uint32_t)(32 * mh + m + e);
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x[e]) : "r"(src + mm * 128 + ((c4 ^ (mm & 7)) << 4)));
          }
          bsum += (double)(x[0] + x[1]);
          split3_pair(x[0], x[1], p1[m >> 1], p2[m >> 1], p3[m >> 1]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar + st);    // staging slot consumed: the producer may refill it
        mbar_wait(planes_free, pf ^ 1);                // MMAs of the previous chunk have read the operand tiles
        pf ^= 1;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t off = ((uint32_t)(4 * mh + g) ^ sw) << 4;
          sts128(dst0 + off, make_uint4(p1[4 * g], p1[4 * g + 1], p1[4 * g + 2], p1[4 * g + 3]));
          sts128(dst0 + A_BYTES + off, make_uint4(p2[4 * g], p2[4 * g + 1], p2[4 * g + 2], p2[4 * g + 3]));
          if (a.nterms > 3) sts128(dst0 + 2 * A_BYTES + off, make_uint4(p3[4 * g], p3[4 * g + 1], p3[4 * g + 2], p3[4 * g + 3]));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(conv_full);
        if (++st == NST) { st = 0; ph ^= 1; }
      }
      if (!isA && a.dbias && blockIdx.x == 0 && n0 + row < a.N) atomicAdd(a.dbias + n0 + row, (float)bsum);
    } else {
      // ===================== epilogue: TMEM -> shared (transpose) -> coalesced fp32 atomics =====================
      const int lg = warp & 3;
      float* tile = reinterpret_cast<float*>(stage_base);
      mbar_wait(tmem_full, 0);
      tcgen05_fence_after();
      const int r = lg * 32 + lane;
#pragma unroll 1
      for (int cc = 0; cc < BNW; cc += 32) {
        uint32_t v[32];
        tmem_ld_x16(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)cc, v);
        tmem_ld_x16(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(cc + 16), v + 16);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 32; ++q) tile[r * TSTRIDE + cc + q] = __uint_as_float(v[q]);
      }
      tcgen05_fence_before();
      asm volatile("bar.sync 1, 128;" ::: "memory");   // the four epilogue warps only
      const int c = (warp - (2 + CONV_WARPS)) * 32 + lane;   // column owned for the atomics
      const int n = n0 + c;
      const int krows = min(128, a.K[sidx] - k0);
      if (n < a.N) {
        float* dst = a.dW + (int64_t)(a.koff[sidx] + k0) * a.ldw + n;
        for (int rr = 0; rr < krows; ++rr) atomicAdd(dst + (int64_t)rr * a.ldw, tile[rr * TSTRIDE + c]);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<BNW>(tmem_base);
  }
}

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
static int make_map(CUtensorMap* m, const float* base, int cols, int64_t ld, int Ti, int B) {
  EncodeFn enc = get_encode();
  FWN_CHECK(enc, "cuTensorMapEncodeTiled unavailable");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)Ti, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)ld * 4 * (cuuint64_t)Ti};
  cuuint32_t box[3] = {32, MC, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FWN_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(wgrad operand cols=%d ld=%lld Ti=%d B=%d) failed: %d", cols, (long long)ld, Ti, B, (int)r);
  return 0;
}

}  // namespace wg3

bool wgrad_tc3_supported(const WgradArgs& a) {
  auto ok = [](const void* p, int64_t ld) { return (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  for (int s = 0; s < a.nseg; ++s)
    if (!ok(a.seg[s].A, a.seg[s].lda) || a.seg[s].K <= 0) return false;
  if (!ok(a.dY0, a.ld0)) return false;
  if (a.n0cols < a.N && (!ok(a.dY1, a.ld1) || a.n0cols % wg3::BNW != 0)) return false;
  return true;
}

int wgrad_tc3(const WgradArgs& w, float* dbias, int nterms, cudaStream_t st) {
  if (w.B <= 0 || w.Ti <= 0 || w.N <= 0) return 0;
  static bool configured = false;
  if (!configured) {
    FWN_CUDA(cudaFuncSetAttribute(wg3::wgrad_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wg3::SMEM));
    configured = true;
  }
  wg3::Wg3Args a;
  memset(&a, 0, sizeof(a));
  int ktiles = 0;
  for (int s = 0; s < w.nseg; ++s) {
    if (wg3::make_map(&a.mapA[s], reinterpret_cast<const float*>(w.seg[s].A), w.seg[s].K, w.seg[s].lda, w.Ti, w.B)) return 1;
    a.shift[s] = w.seg[s].shift; a.K[s] = w.seg[s].K; a.koff[s] = w.seg[s].koff;
    a.ktiles[s] = (w.seg[s].K + 127) / 128;
    ktiles += a.ktiles[s];
  }
  const int n0cols = std::min(w.n0cols, w.N);
  if (wg3::make_map(&a.mapY[0], w.dY0, n0cols, w.ld0, w.Ti, w.B)) return 1;
  if (n0cols < w.N) {
    if (wg3::make_map(&a.mapY[1], w.dY1, w.N - n0cols, w.ld1, w.Ti, w.B)) return 1;
  } else {
    a.mapY[1] = a.mapY[0];
  }
  a.dbias = dbias;
  a.nseg = w.nseg; a.n0cols = n0cols; a.N = w.N; a.dW = w.dW; a.ldw = w.ldw; a.B = w.B; a.Ti = w.Ti;
  a.nterms = nterms == 3 ? 3 : 6;
  a.chunks_per_utt = (w.Ti + wg3::MC - 1) / wg3::MC;
  const int ntiles = (w.N + wg3::BNW - 1) / wg3::BNW;
  const int64_t tiles = (int64_t)ktiles * ntiles;
  // one wave of CTAs if possible, but keep >= 4 chunks per slab so the epilogue (atomics) stays amortised
  const int64_t total_chunks = (int64_t)w.B * a.chunks_per_utt;
  int64_t slabs = std::max<int64_t>(1, (int64_t)num_sms() / tiles);                 // at most one wave of CTAs
  slabs = std::max<int64_t>(1, std::min<int64_t>(slabs, total_chunks / 4));          // >= 4 chunks per slab
  a.slabs = (int)slabs;
  FWN_CHECK(slabs <= 65535 && ntiles <= 65535, "wgrad: grid too large");
  dim3 grid((unsigned)ktiles, (unsigned)ntiles, (unsigned)slabs);
  wg3::wgrad_tc3_kernel<<<grid, wg3::THREADS, wg3::SMEM, st>>>(a);
  FWN_LAUNCH_CHECK();
  return 0;
}

}  // namespace fwn

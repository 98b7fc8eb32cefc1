// Weight gradients of the 16-bit (bf16) training mode on the tensor cores:
//   dW[koff + k, n] += sum_{b,t} A[b, t + shift, k] * dY[b, t, n]          (A, dY bf16 [B, Ti, C]; dW fp32)
// The reduction index is TIME, which is the slow (row) index of both operands in HBM.  tcgen05.mma accepts such operands directly:
// with MN-major shared-memory descriptors the tile TMA writes -- rows = time steps, 64 channels = 128 bytes per row, 128B swizzle --
// IS the canonical layout ((8,8,m),(8,k)):((1,8,LBO),(64,SBO)) of an MN-major operand (channels contiguous, 8-row groups 1024 B
// apart, 64-channel groups LBO apart).  So, unlike the fp32 split engine (wgrad_tc3.cu), no converter warps and no transposes:
// TMA -> tcgen05.mma -> TMEM, one [128 k x 128 n] tile of dW per CTA over a slab of (utterance, 64-step chunk) pairs, partial tiles
// combined with coalesced fp32 atomics through a shared-memory transpose.  The bias gradient (column sums of dY) is taken from the
// staged dY tiles by the otherwise idle epilogue warps of the first k-tile's CTAs.
// Reference: the tf.gradients of every conv kernel / bias in modules.py:24-33,117-127 (train.py:62-63).
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "kernels.h"

namespace fwn {
namespace wg16 {
using namespace tc;

constexpr int BNW = 128;                         // n tile (dY channels)
constexpr int MC = 64;                           // time steps per chunk = 4 MMAs of K = 16
constexpr int BOX_BYTES = MC * 128;              // one TMA box [64 time steps x 64 channels] bf16 = 8 KB
constexpr int OP_BYTES = 2 * BOX_BYTES;          // 128 channels of one operand
constexpr int STAGE_BYTES = 2 * OP_BYTES;        // A tile + dY tile = 32 KB
constexpr int NST = 5;
constexpr int THREADS = 64 + 128;                // warp 0 TMA, warp 1 MMA, warps 2-5 bias sums + epilogue
constexpr int TSTRIDE = 129;                     // fp32 transpose buffer row pitch (floats)
constexpr size_t SMEM = 1024 + (size_t)NST * STAGE_BYTES + 256;
static_assert((size_t)128 * TSTRIDE * 4 <= (size_t)NST * STAGE_BYTES, "transpose buffer must fit in the staging area");

struct alignas(64) Wg16Args {
  CUtensorMap mapA[4];   // per K segment: (K, Ti, B) bf16, box (64, 64, 1), 128B swizzle, zero OOB fill (= tf.pad for shifted taps)
  CUtensorMap mapY[2];   // per column segment of dY
  int shift[4], K[4], koff[4], ktiles[4];
  int nseg, n0cols, N;
  float* dW;
  int64_t ldw;
  float* dbias;          // nullable
  int B, Ti, chunks_per_utt, slabs;
};

// MN-major, 128-byte swizzle descriptor: start address, LBO = distance between 64-channel groups, SBO = 1024 (8 time steps)
__device__ __forceinline__ uint64_t make_mn_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((BOX_BYTES >> 4) & 0x3FFF) << 16;   // leading byte offset
  d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset
  d |= (uint64_t)1 << 46;                             // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
  return d;
}

__global__ void __launch_bounds__(THREADS, 1) wgrad_tc_kernel(const __grid_constant__ Wg16Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)NST * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + NST;
  uint64_t* tmem_full = empty_bar + NST;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int kt = blockIdx.x, sidx = 0;
  while (sidx < a.nseg - 1 && kt >= a.ktiles[sidx]) { kt -= a.ktiles[sidx]; ++sidx; }
  const int k0 = kt * 128, n0 = blockIdx.y * BNW;
  const int64_t total_chunks = (int64_t)a.B * a.chunks_per_utt;
  const int c_begin = (int)(total_chunks * blockIdx.z / a.slabs), c_end = (int)(total_chunks * (blockIdx.z + 1) / a.slabs);
  const int nchunks = c_end - c_begin;
  const bool do_bias = a.dbias != nullptr && blockIdx.x == 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&a.mapA[sidx]);
    prefetch_tmap(&a.mapY[n0 < a.n0cols ? 0 : 1]);
    for (int i = 0; i < NST; ++i) { mbar_init(full_bar + i, 1); mbar_init(empty_bar + i, do_bias ? 5 : 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<BNW>(tmem_ptr);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_launch_dependents();

  if (nchunks > 0) {
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (lane == 0) {
        pdl_wait();
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
          tma_load_3d(sa, &a.mapA[sidx], full_bar + st, k0, t + a.shift[sidx], ub);
          tma_load_3d(sa + BOX_BYTES, &a.mapA[sidx], full_bar + st, k0 + 64, t + a.shift[sidx], ub);
          tma_load_3d(sa + OP_BYTES, my, full_bar + st, ny, t, ub);
          tma_load_3d(sa + OP_BYTES + BOX_BYTES, my, full_bar + st, ny + 64, t, ub);
          if (++st == NST) { st = 0; ph ^= 1; }
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      // D[k, n] += A^T[k, t] * dY[t, n]: both operands MN-major (bits 15 / 16 of the instruction descriptor)
      constexpr uint32_t idesc = make_idesc<BNW>() | (1u << 15) | (1u << 16);
      const uint32_t stage0 = smem_u32(stage_base);
      int st = 0;
      uint32_t ph = 0, accumulate = 0;
      for (int ch = 0; ch < nchunks; ++ch) {
        mbar_wait(full_bar + st, ph);
        tcgen05_fence_after();
        const uint32_t sa = stage0 + (uint32_t)st * STAGE_BYTES;
        const uint64_t adesc = make_mn_desc(sa), bdesc = make_mn_desc(sa + OP_BYTES);
        // four K = 16 steps: 16 time steps = 2048 bytes (+128 in 16-byte units) further down the tile
        asm volatile(
            "{\n"
            ".reg .pred pe, pacc, pt;\n"
            ".reg .b64 da, db;\n"
            "elect.sync _|pe, 0xffffffff;\n"
            "setp.ne.b32 pacc, %4, 0;\n"
            "setp.eq.u32 pt, 0, 0;\n"
            "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pacc;\n"
            "add.u64 da, %1, 128;\n add.u64 db, %2, 128;\n"
            "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n"
            "add.u64 da, %1, 256;\n add.u64 db, %2, 256;\n"
            "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n"
            "add.u64 da, %1, 384;\n add.u64 db, %2, 384;\n"
            "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n"
            "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n"
            "}\n" ::"r"(tmem_base),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(smem_u32(empty_bar + st))
            : "memory");
        accumulate = 1;
        if (++st == NST) { st = 0; ph ^= 1; }
      }
      umma_commit_elect<false>(smem_u32(tmem_full));
    } else {
      // ===================== warps 2-5: bias gradient during the main loop, then the epilogue =====================
      const int ew = warp - 2;
      const int c = ew * 32 + lane;                 // the dY column (bias sums) / dW column (atomics) this thread owns
      if (do_bias) {
        // column c of the staged dY tile: box c / 64, row t at t * 128 bytes, 16-byte chunk ((c % 64) / 8) ^ (t % 8)
        const uint32_t col_off = (uint32_t)OP_BYTES + (uint32_t)(c >> 6) * BOX_BYTES + (uint32_t)(c & 7) * 2;
        const uint32_t cch = (uint32_t)(c & 63) >> 3;
        const uint32_t stage0 = smem_u32(stage_base);
        float bsum = 0.f;
        int st = 0;
        uint32_t ph = 0;
        for (int ch = 0; ch < nchunks; ++ch) {
          mbar_wait(full_bar + st, ph);
          const uint32_t src = stage0 + (uint32_t)st * STAGE_BYTES + col_off;
          float part = 0.f;
#pragma unroll 8
          for (uint32_t t = 0; t < MC; ++t) {
            uint16_t h;
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(src + t * 128 + ((cch ^ (t & 7)) << 4)));
            part += __uint_as_float((uint32_t)h << 16);
          }
          bsum += part;
          __syncwarp();
          if (lane == 0) mbar_arrive(empty_bar + st);
          if (++st == NST) { st = 0; ph ^= 1; }
        }
        if (n0 + c < a.N) atomicAdd(a.dbias + n0 + c, bsum);
      }
      const int lg = warp & 3;                      // TMEM lane group this warp may read
      float* tile = reinterpret_cast<float*>(stage_base);
      mbar_wait(tmem_full, 0);                      // every MMA has completed: the staging area is free for the tensor pipe ...
      tcgen05_fence_after();
      asm volatile("bar.sync 1, 128;" ::: "memory");   // ... and every epilogue warp has finished its bias sums over the last stages
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

}  // namespace wg16

int tc_map_3d(CUtensorMap* out, const void* base, int C, int Ti, int B, int64_t ld, int box_c, int box_rows, bool fp16);   // gemm_tc.cu (cached)

bool wgrad_tc_supported(const Wgrad16Args& a) {
  auto ok = [](const void* p, int64_t ld) { return (ld & 7) == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  for (int s = 0; s < a.nseg; ++s)
    if (!ok(a.seg[s].A, a.seg[s].lda) || a.seg[s].K <= 0) return false;
  if (!ok(a.dY0, a.ld0)) return false;
  if (a.n0cols < a.N && (!ok(a.dY1, a.ld1) || a.n0cols % wg16::BNW != 0)) return false;
  return true;
}

int wgrad_tc(const Wgrad16Args& w, float* dbias, cudaStream_t st) {
  if (w.B <= 0 || w.Ti <= 0 || w.N <= 0) return 0;
  FWN_CHECK(wgrad_tc_supported(w), "wgrad_tc: operands must be 16-byte aligned with row pitches that are multiples of 8 elements");
  static bool configured = false;
  if (!configured) {
    FWN_CUDA(cudaFuncSetAttribute(wg16::wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wg16::SMEM));
    configured = true;
  }
  wg16::Wg16Args a;
  memset(&a, 0, sizeof(a));
  int ktiles = 0;
  for (int s = 0; s < w.nseg; ++s) {
    if (tc_map_3d(&a.mapA[s], w.seg[s].A, w.seg[s].K, w.Ti, w.B, w.seg[s].lda, 64, wg16::MC, false)) return 1;
    a.shift[s] = w.seg[s].shift; a.K[s] = w.seg[s].K; a.koff[s] = w.seg[s].koff;
    a.ktiles[s] = (w.seg[s].K + 127) / 128;
    ktiles += a.ktiles[s];
  }
  const int n0cols = std::min(w.n0cols, w.N);
  if (tc_map_3d(&a.mapY[0], w.dY0, n0cols, w.Ti, w.B, w.ld0, 64, wg16::MC, false)) return 1;
  if (n0cols < w.N) {
    if (tc_map_3d(&a.mapY[1], w.dY1, w.N - n0cols, w.Ti, w.B, w.ld1, 64, wg16::MC, false)) return 1;
  } else {
    a.mapY[1] = a.mapY[0];
  }
  a.dbias = dbias;
  a.nseg = w.nseg; a.n0cols = n0cols; a.N = w.N; a.dW = w.dW; a.ldw = w.ldw; a.B = w.B; a.Ti = w.Ti;
  a.chunks_per_utt = (w.Ti + wg16::MC - 1) / wg16::MC;
  const int ntiles = (w.N + wg16::BNW - 1) / wg16::BNW;
  const int64_t tiles = (int64_t)ktiles * ntiles;
  const int64_t total_chunks = (int64_t)w.B * a.chunks_per_utt;
  int64_t slabs = std::max<int64_t>(1, (int64_t)num_sms() / tiles);                 // at most one wave of CTAs
  slabs = std::max<int64_t>(1, std::min<int64_t>(slabs, total_chunks / 4));          // >= 4 chunks per slab: the atomics stay amortised
  a.slabs = (int)slabs;
  FWN_CHECK(slabs <= 65535 && ntiles <= 65535, "wgrad: grid too large");
  if (getenv("FWN_TRACE")) {
    fprintf(stderr, "wgrad_tc B=%d Ti=%d N=%d n0=%d ktiles=%d ntiles=%d slabs=%d bias=%d:", w.B, w.Ti, w.N, n0cols, ktiles, ntiles, (int)slabs, dbias != nullptr);
    for (int s = 0; s < w.nseg; ++s) fprintf(stderr, " [K=%d lda=%lld sh=%d koff=%d]", w.seg[s].K, (long long)w.seg[s].lda, w.seg[s].shift, w.seg[s].koff);
    fprintf(stderr, "\n");
    fflush(stderr);
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)ktiles, (unsigned)ntiles, (unsigned)slabs);
  cfg.blockDim = dim3(wg16::THREADS);
  cfg.dynamicSmemBytes = wg16::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  FWN_CUDA(cudaLaunchKernelEx(&cfg, wg16::wgrad_tc_kernel, a));
  FWN_LAUNCH_CHECK();
  return 0;
}

}  // namespace fwn

// One ResBlock layer of the coupling WaveNet in ONE launch (mixed-precision inference passes, modules.py:113-128):
//
//   o      = tanh(f) * sigmoid(g),  (f, g) = dilated conv k=3 of h_in  +  1x1 of the conditioning   (gate GEMM, N = 2F = 512)
//   h_out  = (h_in + o . W_res + b_res) * sqrt(1/2)                                                 (res GEMM,  N = F,  K = F)
//   skip  += o . W_skip + b_skip            (last layer: relu(skip), no residual)                   (skip GEMM, N = F,  K = F)
//
// The unfused pair (tc_gemm_kernel<GATE> + tc_gemm_kernel<RES_SKIP>) writes o [rows, 256] to HBM and reads it straight back; here the
// gate epilogue writes its 16-bit tile into shared memory in the K-major 128B-swizzled layout tcgen05.mma consumes, and the 1x1 GEMMs
// take it from there as their A operand.  Only their weights (256 KB per row-tile pair, from L2) stream through the TMA pipeline.
//
// CTA pair (cta_group::2, M = 256 rows across the two SMs), per pair of row tiles a fixed program of accumulator "ops", each
// N = 256 wide, alternating between the two 256-column halves of TMEM:
//     G0 (gate columns 0..255 -> o channels 0..127), G1 (gate columns 256..511 -> o channels 128..255), [R (residual)], K (skip)
// software-pipelined by one gate half so the 1x1 MMAs never wait for a gate epilogue:
//     G0(0) G1(0) | G0(1) [R(0)] K(0) G1(1) | G0(2) [R(1)] K(1) G1(2) | ... | [R(n-1)] K(n-1)
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer (leader CTA only), warps 2..17 epilogue.  All 16 epilogue warps work on every
// op in program order (thread = one row x 64 accumulator columns, one tcgen05.wait::ld) and hand the TMEM half back to the MMA warp
// BEFORE the arithmetic.  G0(j+1)'s half of the o tile is still the A operand of R(j) / K(j), which are issued right after it: its
// store is deferred (16 registers) into the next op's epilogue, after that op's accumulator has been fetched and K(j) has completed.
// Shared memory (227 KB): 3 pipeline stages x (16 KB A + 16 KB weight half-box) | o tile 64 KB | staging tile 64 KB | gate bias.
// The 1x1 ops pack two weight K chunks into one stage.  The staging tile carries their in-place epilogue I/O: TMA loads h_in (or the
// running skip sum) one gate half ahead, the residual op updates each warp's 32 x 64 region in place and TMA-stores h_out, then the
// skip op reuses the region for the skip tile.
// Last layer with LayerArgs::tail: the WaveNet tail rides in the same launch as two more op kinds -- F (final 1x1 on relu(skip sum), which
// the skip op leaves in the staging tile instead of storing it) and Z (zero conv on relu(final), written in place over it; epilogue =
// ActNorm + affine coupling on x, model.py:7-105,121-161) -- see for_each_op for the program.  Neither the skip sum nor u reaches HBM.
// Measurements, timelines and the rejected variants: profiles/r2_ncu_fused.md.
#include <cuda.h>
#include <stdio.h>

#include "common.cuh"
#include "layer_tc.cuh"
#include "tc_ptx.cuh"

namespace fwn {
namespace tc {

constexpr int L_STAGES = 3;
constexpr int L_STAGE_BYTES = 2 * A_BYTES;   // activation chunk + this CTA's half of the weight box
constexpr int L_TILE_BYTES = 4 * A_BYTES;    // [128 rows x 256 channels] 16-bit = four swizzled [128 x 64] chunks
constexpr int L_EPI_WARPS = 16;
constexpr int L_THREADS = 64 + 32 * L_EPI_WARPS;
constexpr int L_BIAS_BYTES = 2048;           // 512 gate biases
constexpr size_t L_SMEM = (size_t)L_STAGES * L_STAGE_BYTES + 2 * L_TILE_BYTES + L_BIAS_BYTES + 256;
static_assert(L_SMEM <= 227 * 1024, "shared memory budget exceeded");

// byte offset of the 16-byte chunk holding channels [c, c+8) of row r inside a [128 x 256] tile (TMA SWIZZLE_128B, 64-channel chunks)
__device__ __forceinline__ uint32_t tile_off(int r, int c) {
  return (uint32_t)((c >> 6) * A_BYTES + r * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4));
}
// arrive (release at cluster scope) on the barrier at this offset in CTA `cta` of the cluster: orders this thread's earlier writes
// (made visible to the async proxy by fence.proxy.async) before the MMA the leader issues after its acquire
__device__ __forceinline__ void mbar_arrive_remote_release(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}

// umma_chunk_commit<true> / umma_commit_elect<true> (tc_ptx.cuh) with an explicit CTA mask for the commit's arrive
__device__ __forceinline__ void umma_chunk_commit_mask(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                                       uint32_t ksteps, uint32_t bar, uint16_t mask) {
  asm volatile(
      "{\n"
      ".reg .pred pe, pacc, pt, p1, p2, p3;\n"
      ".reg .b64 da, db;\n"
      "elect.sync _|pe, 0xffffffff;\n"
      "setp.ne.b32 pacc, %4, 0;\n"
      "setp.eq.u32 pt, 0, 0;\n"
      "setp.gt.u32 p1, %5, 1;\n and.pred p1, p1, pe;\n"
      "setp.gt.u32 p2, %5, 2;\n and.pred p2, p2, pe;\n"
      "setp.gt.u32 p3, %5, 3;\n and.pred p3, p3, pe;\n"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, pacc;\n"
      "add.u64 da, %1, 2;\n add.u64 db, %2, 2;\n"
      "@p1 tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, pt;\n"
      "add.u64 da, %1, 4;\n add.u64 db, %2, 4;\n"
      "@p2 tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, pt;\n"
      "add.u64 da, %1, 6;\n add.u64 db, %2, 6;\n"
      "@p3 tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, pt;\n"
      "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%6], %7;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(ksteps), "r"(bar), "h"(mask)
      : "memory");
}
// the same four MMAs without the commit (two weight boxes share one pipeline stage in the 1x1 ops)
__device__ __forceinline__ void umma_chunk_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred pe, pacc, pt;\n"
      ".reg .b64 da, db;\n"
      "elect.sync _|pe, 0xffffffff;\n"
      "setp.ne.b32 pacc, %4, 0;\n"
      "setp.eq.u32 pt, 0, 0;\n"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, pacc;\n"
      "add.u64 da, %1, 2;\n add.u64 db, %2, 2;\n"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, pt;\n"
      "add.u64 da, %1, 4;\n add.u64 db, %2, 4;\n"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, pt;\n"
      "add.u64 da, %1, 6;\n add.u64 db, %2, 6;\n"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, pt;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mask(uint32_t bar, uint16_t mask) {
  asm volatile(
      "{\n.reg .pred pe;\nelect.sync _|pe, 0xffffffff;\n"
      "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n}\n" ::"r"(bar),
      "h"(mask)
      : "memory");
}

// L2 prefetch of one tensor-map box (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

enum LOp { OP_G0 = 0, OP_G1 = 1, OP_RES = 2, OP_SKIP = 3, OP_FIN = 4, OP_ZERO = 5 };

// The op program of one CTA pair over its n row-tile pairs -- see layer_kernel.  f(tile, kind, gop) is called once per op, in order:
//   layer only:   G0(0) G1(0) | G0(1) [R(0)] K(0) G1(1) | G0(2) [R(1)] K(1) G1(2) | ... | [R(n-1)] K(n-1)
//   with tail:    G0(0) G1(0) | G0(1) K(0) G1(1) F(0) | G0(2) Z(0) K(1) G1(2) F(1) | ... | Z(n-2) K(n-1) F(n-1) | Z(n-1)
// (every op that consumes an epilogue's shared-memory tile is issued one gate half after the op that produces it)
template <typename F>
__device__ __forceinline__ void for_each_op(int n, bool has_res, bool tail, F&& f) {
  int gop = 0;
  for (int jb = -1; jb < n + (tail ? 1 : 0); ++jb)
    for (int step = 0; step < 6; ++step) {
      int it, kind;
      if (step == 0) { if (jb + 1 >= n) continue; it = jb + 1; kind = OP_G0; }
      else if (step == 1) { if (!tail || jb < 1) continue; it = jb - 1; kind = OP_ZERO; }
      else if (step == 2) { if (jb < 0 || jb >= n || !has_res) continue; it = jb; kind = OP_RES; }
      else if (step == 3) { if (jb < 0 || jb >= n) continue; it = jb; kind = OP_SKIP; }
      else if (step == 4) { if (jb + 1 >= n) continue; it = jb + 1; kind = OP_G1; }
      else { if (!tail || jb < 0 || jb >= n) continue; it = jb; kind = OP_FIN; }
      f(it, kind, gop);
      ++gop;
    }
}

// diagnostics: slot (it, opi, k) of the timeline, leader CTA of cluster 0, first L_TRACE_TILES tiles
constexpr int L_TRACE_TILES = 8, L_TRACE_K = 8, L_TRACE_T0 = 30;
#define L_TRACE(k)                                                                                             \
  do {                                                                                                         \
    if (a.trace && blockIdx.x == 0 && it >= L_TRACE_T0 && it < L_TRACE_T0 + L_TRACE_TILES && lane == 0) {       \
      a.trace[((it - L_TRACE_T0) * 6 + kind) * L_TRACE_K + (k)] = clock64();                                   \
      if ((k) == 1 && kind == OP_G0) {                                                                         \
        unsigned long long gt_;                                                                                \
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));                                                \
        a.trace[L_TRACE_TILES * 6 * L_TRACE_K + (it - L_TRACE_T0)] = (long long)gt_;                           \
      }                                                                                                        \
    }                                                                                                          \
  } while (0)

__global__ void __launch_bounds__(L_THREADS, 1) layer_kernel(const __grid_constant__ LayerArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* stage_base = smem;
  uint8_t* o_base = smem + (size_t)L_STAGES * L_STAGE_BYTES;
  uint8_t* stg = o_base + L_TILE_BYTES;
  float* sbias = reinterpret_cast<float*>(stg + L_TILE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sbias) + L_BIAS_BYTES);
  uint64_t* empty_bar = full_bar + L_STAGES;
  uint64_t* tmem_full = empty_bar + L_STAGES;   // [2] accumulator half complete (own CTA; the leader's commit is multicast)
  uint64_t* tmem_empty = tmem_full + 2;         // [2] leader: 16 warps x 2 CTAs hold the half's accumulators in registers
  uint64_t* o_full = tmem_empty + 2;            // [2] leader: o channels [0,128) / [128,256) of both CTAs written
  uint64_t* stg_full = o_full + 2;              // staging tile pre-loaded (TMA)
  uint64_t* stg_empty = stg_full + 1;          // the skip op's stores (16 warps) have read the staging tile
  uint64_t* su_full = stg_empty + 1;            // [2] tail, leader: relu(skip) / relu(final) tile of both CTAs written (16 warps x 2)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(su_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const uint32_t lead_cta = 0;                      // cluster rank of the pair's leader
  const bool leader = rank == 0;
  const uint16_t all_mask = 3, pair_mask = 3;       // commit arrives on both CTAs of the pair
  const int num_m_tiles = a.B * a.tiles_per_utt;
  const int n_pairs = (num_m_tiles + 1) / 2;
  const int cl = (int)blockIdx.x >> 1, ncl = (int)gridDim.x >> 1;
  const int n_units = n_pairs;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) {
      printf("layer_kernel: dynamic shared memory is not 1024-byte aligned\n");
      asm volatile("trap;");
    }
    for (int s = 0; s < a.nseg; ++s) prefetch_tmap(&a.mapA[s]);
    prefetch_tmap(&a.mapWg);
    prefetch_tmap(&a.mapWr);
    if (a.tail) {
      prefetch_tmap(&a.mapWf);
      prefetch_tmap(&a.mapWz);
    }
    for (int i = 0; i < L_STAGES; ++i) {
      mbar_init(full_bar + i, 1);
      mbar_init(empty_bar + i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tmem_full + i, 1);
      mbar_init(tmem_empty + i, 32);
      mbar_init(o_full + i, 32);
    }
    mbar_init(stg_full, 1);
    mbar_init(stg_empty, 16);
    mbar_init(su_full, 32);
    mbar_init(su_full + 1, 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 512; i += L_THREADS) sbias[i] = __ldg(a.gate_bias + i);
  if (warp == 1) tmem_alloc_2sm<512>(tmem_ptr);
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();   // peer barriers are initialised before any remote arrive / multicast commit
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_launch_dependents();

  // The op program of this CTA pair over its n_my row-tile pairs, software-pipelined by one gate half so that the MMA pipe never
  // waits for a gate epilogue:   G0(0) G1(0) | G0(1) [R(0)] K(0) G1(1) | G0(2) [R(1)] K(1) G1(2) | ... | [R(n-1)] K(n-1)
  // (the 1x1 ops of tile j are issued after the first gate half of tile j+1: by then G1(j)'s epilogue has had a whole gate half to
  // finish the o tile).  Every role walks the same sequence; op number gop uses TMEM half gop & 1.
  const int n_my = (n_units - cl + ncl - 1) / ncl;
  const bool tail = a.tail != 0;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      pdl_wait();
      auto coords = [&](int it, int& ub, int& t0) {
        const int m_tile = 2 * (cl + it * ncl) + rank;
        ub = m_tile / a.tiles_per_utt;
        t0 = (m_tile - ub * a.tiles_per_utt) * BM;
      };
      // staging tile of tile js (h_in for the residual op / running skip sum of the last layer), one 16 KB sub-tile per call
      auto stage_in = [&](int js, int sub) {
        if (!a.has_in || tail || (a.dbg & 2)) return;   // with the tail fused in, the epilogue warps reload the staging tile themselves
        int ub, t0;
        coords(js, ub, t0);
        if (sub == 0) {
          if (js > 0) mbar_wait(stg_empty, (uint32_t)(js - 1) & 1);   // the previous tile's skip stores have read the tile
          mbar_expect_tx(stg_full, L_TILE_BYTES);
        }
        tma_load_3d(stg + (size_t)sub * A_BYTES, &a.mapIn, stg_full, sub * 64, t0, ub);
      };
      for_each_op(n_my, a.has_res != 0, tail, [&](int it, int kind, int gop) {
        (void)gop;
        int ub, t0;
        coords(it, ub, t0);
        if (kind <= OP_G1) {
          int ub2 = 0, t02 = 0;
          const bool pf = kind == OP_G1 && it + 1 < n_my;   // the next tile's activation boxes -> L2 while this half streams
          if (pf) coords(it + 1, ub2, t02);
          int c = 0;
          for (int s = 0; s < a.nseg; ++s) {
            for (int ch = 0; ch < a.nchunk[s]; ++ch, ++c) {
              mbar_wait(empty_bar + stage, phase ^ 1);
              uint8_t* sa = stage_base + (size_t)stage * L_STAGE_BYTES;
              const bool skip_a = (a.dbg & 8) && kind == OP_G1;   // diagnostics: how much of the time is the activation traffic?
              if (leader) mbar_expect_tx(full_bar + stage, skip_a ? 2 * A_BYTES : 2 * L_STAGE_BYTES);   // both CTAs' loads complete on the leader's barrier
              if (!skip_a) tma_load_3d_2sm(sa, &a.mapA[s], full_bar + stage, ch * BK, t0 + a.shift[s], ub);
              tma_load_2d_2sm(sa + A_BYTES, &a.mapWg, full_bar + stage, a.wk0[s] + ch * BK, kind * 256 + rank * 128);
              if (++stage == L_STAGES) { stage = 0; phase ^= 1; }
              if (pf && !(a.dbg & 16)) tma_prefetch_3d(&a.mapA[s], ch * BK, t02 + a.shift[s], ub2);
              // the 1x1 ops of the previous tile follow this gate half: their staging tile is fetched now, one sub-tile per chunk
              if (kind == OP_G0 && it >= 1 && c < 4) stage_in(it - 1, c);
            }
          }
        } else if (kind == OP_ZERO) {
          // zero conv: four K chunks of this CTA's NzBox/2 weight rows, all in one stage
          const uint32_t zb = (uint32_t)(a.NzBox / 2) * 128u;
          mbar_wait(empty_bar + stage, phase ^ 1);
          uint8_t* sa = stage_base + (size_t)stage * L_STAGE_BYTES;
          if (leader) mbar_expect_tx(full_bar + stage, 2u * 4u * zb);
          for (int kc = 0; kc < 4; ++kc) tma_load_2d_2sm(sa + (size_t)kc * zb, &a.mapWz, full_bar + stage, kc * BK, rank * (a.NzBox / 2));
          if (++stage == L_STAGES) { stage = 0; phase ^= 1; }
        } else {
          if (it == n_my - 1 && !tail && (kind == OP_RES || (kind == OP_SKIP && !a.has_res)))   // last tile: no gate half precedes its 1x1 ops
            for (int sub = 0; sub < 4; ++sub) stage_in(it, sub);
          // 1x1 ops (residual / skip on the o tile, final conv on the staging tile): only the weight half-boxes stream
          const CUtensorMap* wm = kind == OP_FIN ? &a.mapWf : &a.mapWr;
          const int n0 = ((a.has_res && kind == OP_SKIP) ? 256 : 0) + rank * 128;
          for (int kc = 0; kc < 4; kc += 2) {   // two K chunks of weight half-boxes per stage (a stage has room for an A and a B box)
            mbar_wait(empty_bar + stage, phase ^ 1);
            uint8_t* sa = stage_base + (size_t)stage * L_STAGE_BYTES;
            if (leader) mbar_expect_tx(full_bar + stage, 2 * L_STAGE_BYTES);
            tma_load_2d_2sm(sa, wm, full_bar + stage, kc * BK, n0);
            tma_load_2d_2sm(sa + A_BYTES, wm, full_bar + stage, (kc + 1) * BK, n0);
            if (++stage == L_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      });
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA of the pair) =====================
    if (leader) {
      constexpr uint32_t idesc_bf16 = (make_idesc<256>() & ~(0x1Fu << 24)) | ((uint32_t)(256 >> 4) << 24);   // M = 256 across the pair
      const uint32_t idesc = a.fp16 ? (idesc_bf16 & ~IDESC_BF16_BITS) : idesc_bf16;
      const uint32_t stage0 = smem_u32(stage_base), o0 = smem_u32(o_base);
      const uint64_t desc_hi = make_smem_desc(0);
      int stage = 0;
      uint32_t phase = 0;
      for_each_op(n_my, a.has_res != 0, tail, [&](int it, int kind, int gop) {
        const int slot = gop & 1;
        L_TRACE(0);
        mbar_wait(tmem_empty + slot, (((uint32_t)gop >> 1) & 1) ^ 1);
        tcgen05_fence_after();
        L_TRACE(1);
        const uint32_t tmem_d = tmem_base + (uint32_t)(slot * 256);
        uint32_t accumulate = 0;
        if (kind <= OP_G1) {
          for (int s = 0; s < a.nseg; ++s) {
            const int nch = a.nchunk[s];
            const uint32_t last_ks = (uint32_t)a.last_ksteps[s];
            for (int ch = 0; ch < nch; ++ch) {
              mbar_wait(full_bar + stage, phase);
              tcgen05_fence_after();
              const uint32_t sa = stage0 + (uint32_t)stage * L_STAGE_BYTES, sb = sa + A_BYTES;
              const uint64_t adesc = desc_hi | (uint64_t)((sa >> 4) & 0x3FFF), bdesc = desc_hi | (uint64_t)((sb >> 4) & 0x3FFF);
              umma_chunk_commit_mask(tmem_d, adesc, bdesc, idesc, accumulate, (ch == nch - 1) ? last_ks : 4u, smem_u32(empty_bar + stage), all_mask);
              accumulate = 1;
              if (++stage == L_STAGES) { stage = 0; phase ^= 1; }
            }
          }
        } else if (kind == OP_ZERO) {
          // zero conv: A = relu(final) in the staging tile (u_full), B = the stage's four small weight boxes, N = NzBox
          mbar_wait_cluster(su_full + 1, (uint32_t)it & 1);
          mbar_wait(full_bar + stage, phase);
          tcgen05_fence_after();
          const uint32_t idesc_z = (idesc & ~(0x3Fu << 17)) | ((uint32_t)(a.NzBox >> 3) << 17);
          const uint32_t zb = (uint32_t)(a.NzBox / 2) * 128u, sb0 = stage0 + (uint32_t)stage * L_STAGE_BYTES, st0 = smem_u32(stg);
          for (int kc = 0; kc < 3; ++kc)
            umma_chunk_pair(tmem_d, desc_hi | (uint64_t)(((st0 + (uint32_t)kc * A_BYTES) >> 4) & 0x3FFF), desc_hi | (uint64_t)(((sb0 + (uint32_t)kc * zb) >> 4) & 0x3FFF),
                            idesc_z, kc > 0 ? 1u : 0u);
          umma_chunk_commit_mask(tmem_d, desc_hi | (uint64_t)(((st0 + 3u * A_BYTES) >> 4) & 0x3FFF), desc_hi | (uint64_t)(((sb0 + 3u * zb) >> 4) & 0x3FFF), idesc_z, 1u,
                                 4u, smem_u32(empty_bar + stage), all_mask);
          if (++stage == L_STAGES) { stage = 0; phase ^= 1; }
        } else {
          // residual / skip: A = the o tile (channels [0,128) from G0's epilogue, [128,256) from G1's); final conv: A = relu(skip sum) in
          // the staging tile (s_full)
          const uint32_t a0 = kind == OP_FIN ? smem_u32(stg) : o0;
          if (kind == OP_FIN) {
            mbar_wait_cluster(su_full, (uint32_t)it & 1);
            tcgen05_fence_after();
          }
          for (int kc = 0; kc < 4; kc += 2) {
            if (kind != OP_FIN) {
              mbar_wait_cluster(o_full + (kc >> 1), (uint32_t)it & 1);
              tcgen05_fence_after();
              L_TRACE(2 + (kc >> 1));
            }
            mbar_wait(full_bar + stage, phase);
            tcgen05_fence_after();
            const uint32_t sa = a0 + (uint32_t)kc * A_BYTES, sb = stage0 + (uint32_t)stage * L_STAGE_BYTES;
            umma_chunk_pair(tmem_d, desc_hi | (uint64_t)((sa >> 4) & 0x3FFF), desc_hi | (uint64_t)((sb >> 4) & 0x3FFF), idesc, accumulate);
            umma_chunk_commit_mask(tmem_d, desc_hi | (uint64_t)(((sa + A_BYTES) >> 4) & 0x3FFF), desc_hi | (uint64_t)(((sb + A_BYTES) >> 4) & 0x3FFF),
                                   idesc, 1u, 4u, smem_u32(empty_bar + stage), all_mask);
            accumulate = 1;
            if (++stage == L_STAGES) { stage = 0; phase ^= 1; }
          }
        }
        umma_commit_mask(smem_u32(tmem_full + slot), pair_mask);
        L_TRACE(4);
      });
    }
  } else {
    // ===================== epilogue: 16 warps, every op in program order =====================
    // warp = (TMEM lane group lg, column quarter q): thread = one row x 64 of the op's 256 accumulator columns.  All 64 columns are
    // fetched with one tcgen05.wait::ld and the TMEM half is handed back to the MMA warp BEFORE the math, so the next op that needs
    // the half never waits for an epilogue's arithmetic -- only o_full (the gated tile the 1x1 MMAs read) does.
    const int lg = warp & 3;                      // TMEM lane group this warp may access
    const int qtr = (warp - 2) >> 2;              // which 64 of the op's 256 accumulator columns
    const int r = lg * 32 + lane;
    const int cbeg = qtr * 64;
    const bool fp16 = a.fp16 != 0;
    const uint32_t o_u32 = smem_u32(o_base), stg_u32 = smem_u32(stg);
    bool store_pending = false;                   // lane 0: this warp's last TMA store may still be reading the staging tile
    // G0's gated outputs (32 channels of this thread's row) whose store into the o tile is deferred: channels [0,128) of the tile are
    // still the A operand of the previous tile's 1x1 MMAs, which are issued right AFTER this gate half.  The store happens inside the
    // next op's epilogue, after that op's accumulator has been fetched -- so the residual op's TMEM half is handed back to the MMA
    // warp (which needs it for G1) without waiting for the skip MMAs to finish.
    uint4 held[4];
    int held_gk = -1;                             // op number of the K op the deferred store waits for, -1 = nothing held
    double ls_sum = 0.0;                          // tail: sum of log_s over this thread's rows (forward direction)
    pdl_wait();
    // tail: the staging tile cycles  running skip sum (TMA) -> relu(skip sum) (K epilogue) -> relu(final) (F epilogue) -> free once the
    // zero conv's MMAs have read it; the reload for the next tile is issued from here, by one thread, right after that point
    auto load_skip_in = [&](int js) {
      const int m_tile = 2 * (cl + js * ncl) + rank;
      const int ub = m_tile / a.tiles_per_utt;
      const int t0 = (m_tile - ub * a.tiles_per_utt) * BM;
      mbar_expect_tx(stg_full, L_TILE_BYTES);
#pragma unroll
      for (int sub = 0; sub < 4; ++sub) tma_load_3d(stg + (size_t)sub * A_BYTES, &a.mapIn, stg_full, sub * 64, t0, ub);
    };
    if (tail && warp == 2 && lane == 0 && n_my > 0) load_skip_in(0);
    for_each_op(n_my, a.has_res != 0, tail, [&](int it, int kind, int gop) {
      const int m_tile = 2 * (cl + it * ncl) + rank;
      const int ub = m_tile / a.tiles_per_utt;
      const int t0 = (m_tile - ub * a.tiles_per_utt) * BM;
      const int t = t0 + r;
      const bool row_ok = t < a.Ti && ub < a.B;
      const int64_t row = (int64_t)ub * a.Ti + t;
      const int slot = gop & 1;
      const bool have_in = (kind == OP_RES || (kind == OP_SKIP && !a.has_res && a.has_in)) && !(a.dbg & 2);
      if (warp == 2) L_TRACE(5);
      mbar_wait(tmem_full + slot, ((uint32_t)gop >> 1) & 1);
      tcgen05_fence_after();
      if (warp == 2) L_TRACE(6);
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(slot * 256 + cbeg);
      uint32_t v[64];
      tmem_ld_x16(taddr, v);
      tmem_ld_x16(taddr + 16, v + 16);
      tmem_ld_x16(taddr + 32, v + 32);
      tmem_ld_x16(taddr + 48, v + 48);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(tmem_empty + slot, lead_cta);   // accumulator in registers: the MMA warp may reuse the half
      if (held_gk >= 0 && kind != OP_ZERO) {   // deferred store of the previous G0 (see above): the last MMA that reads the old tile is K = op held_gk
        mbar_wait(tmem_full + (held_gk & 1), ((uint32_t)held_gk >> 1) & 1);
#pragma unroll
        for (int j = 0; j < 4; ++j) sts128(o_u32 + tile_off(r, (cbeg + 16 * j) >> 1), held[j]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive_remote_release(o_full + OP_G0, lead_cta);
        held_gk = -1;
      }
      if (kind <= OP_G1) {
        uint4 outv[4];
#pragma unroll
        for (int j = 0; j < 64; j += 16) {
          const int col = kind * 256 + cbeg + j;   // gate column of v[j]: (2c, 2c+1) = (filter_c, gate_c)
          float acc[16];
          const float* bp = sbias + col;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float4 b4 = *reinterpret_cast<const float4*>(bp + 4 * k);
            acc[4 * k] = __uint_as_float(v[j + 4 * k]) + b4.x;
            acc[4 * k + 1] = __uint_as_float(v[j + 4 * k + 1]) + b4.y;
            acc[4 * k + 2] = __uint_as_float(v[j + 4 * k + 2]) + b4.z;
            acc[4 * k + 3] = __uint_as_float(v[j + 4 * k + 3]) + b4.w;
          }
          if (a.pc && row_ok) {   // deep blocks: conditioning projection of this layer computed ahead
            const float4* pp = reinterpret_cast<const float4*>(a.pc + row * a.pc_ld + col);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float4 q = __ldg(pp + k);
              acc[4 * k] += q.x; acc[4 * k + 1] += q.y; acc[4 * k + 2] += q.z; acc[4 * k + 3] += q.w;
            }
          }
          uint32_t pk[4];
          if (fp16) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              pk[k] = pack16(gate_accurate(acc[4 * k], acc[4 * k + 1]), gate_accurate(acc[4 * k + 2], acc[4 * k + 3]), true);
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // tanh(f) * sigmoid(g) = t + t*tanh(g/2) with t = tanh(f)/2 : 2 MUFU + 3 FP32 ops per output (modules.py:124)
              const float t0_ = 0.5f * tanh_fast(acc[4 * k]), t1_ = 0.5f * tanh_fast(acc[4 * k + 2]);
              const float h0_ = tanh_fast(0.5f * acc[4 * k + 1]), h1_ = tanh_fast(0.5f * acc[4 * k + 3]);
              pk[k] = pack_bf16(fmaf(h0_, t0_, t0_), fmaf(h1_, t1_, t1_));
            }
          }
          outv[j >> 4] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
        if (kind == OP_G0 && it >= 1) {
#pragma unroll
          for (int j = 0; j < 4; ++j) held[j] = outv[j];
          held_gk = gop + 1 + (a.has_res ? 1 : 0) + ((tail && it >= 2) ? 1 : 0);   // K(it-1): after [Z(it-2)] [R(it-1)]
        } else {
#pragma unroll
          for (int j = 0; j < 64; j += 16) sts128(o_u32 + tile_off(r, (kind * 256 + cbeg + j) >> 1), outv[j >> 4]);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // o tile (generic-proxy writes) -> visible to the MMA's async-proxy reads
          __syncwarp();
          if (lane == 0) mbar_arrive_remote_release(o_full + kind, lead_cta);
        }
      } else if (kind == OP_ZERO) {
        // the zero conv's MMAs have completed (tmem_full): the staging tile is free -> fetch the next tile's running skip sum
        if (warp == 2 && lane == 0 && it + 1 < n_my) load_skip_in(it + 1);
        if (qtr == 0 && row_ok) {   // 4 warps: (log_s, t) pairs of this row = accumulator columns [0, Nz); ActNorm + coupling in place on x
          const EpiArgs& e = a.ez;
          float* xr = e.X + row * e.Cx;
          const bool xfast = e.pairs_adjacent && e.Cx >= 16;
#pragma unroll
          for (int c0 = 0; c0 < 32; c0 += 16) {
            if (c0 >= a.Nz) break;
            if (xfast) {   // pairs ordered by physical position: 16 columns = 16 consecutive floats of the row
              const float4* bp = reinterpret_cast<const float4*>(e.an_b + c0);
              const float4* sp = reinterpret_cast<const float4*>(e.an_s + c0);
              const int bo = e.b_odd;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float4 b4 = __ldg(bp + k), s4 = __ldg(sp + k);
                const float4 zb4 = __ldg(reinterpret_cast<const float4*>(e.bias + c0 + 4 * k));
                const float4 xq = *reinterpret_cast<const float4*>(xr + c0 + 4 * k);
                float x[4] = {xq.x, xq.y, xq.z, xq.w};
                const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w};
                const float ac[4] = {__uint_as_float(v[c0 + 4 * k]) + zb4.x, __uint_as_float(v[c0 + 4 * k + 1]) + zb4.y,
                                     __uint_as_float(v[c0 + 4 * k + 2]) + zb4.z, __uint_as_float(v[c0 + 4 * k + 3]) + zb4.w};
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const float log_s = ac[2 * h], tt = ac[2 * h + 1];
                  const int ib = 2 * h + bo, ia = 2 * h + 1 - bo;
                  if (!e.reverse) {
                    x[ia] = (x[ia] + bb[ia]) * ss[ia];
                    x[ib] = ((x[ib] + bb[ib]) * ss[ib] - tt) * __expf(-log_s);
                    ls_sum += (double)log_s;
                  } else {
                    x[ib] = (x[ib] * __expf(log_s) + tt) * ss[ib] - bb[ib];
                    x[ia] = x[ia] * ss[ia] - bb[ia];
                  }
                }
                *reinterpret_cast<float4*>(xr + c0 + 4 * k) = make_float4(x[0], x[1], x[2], x[3]);
              }
            } else {
#pragma unroll
              for (int p = 0; p < 8; ++p) {
                const int q = c0 / 2 + p;
                if (q >= e.nq) break;
                const float log_s = __uint_as_float(v[c0 + 2 * p]) + __ldg(e.bias + c0 + 2 * p);
                const float tt = __uint_as_float(v[c0 + 2 * p + 1]) + __ldg(e.bias + c0 + 2 * p + 1);
                const int oa = __ldg(e.a_off + q), ob = __ldg(e.b_off + q);
                float xa = xr[oa], xb = xr[ob];
                if (!e.reverse) {
                  xa = (xa + __ldg(e.an_b + oa)) * __ldg(e.an_s + oa);
                  xb = (xb + __ldg(e.an_b + ob)) * __ldg(e.an_s + ob);
                  xb = (xb - tt) * __expf(-log_s);
                  ls_sum += (double)log_s;
                } else {
                  xb = xb * __expf(log_s) + tt;
                  xa = xa * __ldg(e.an_s + oa) - __ldg(e.an_b + oa);
                  xb = xb * __ldg(e.an_s + ob) - __ldg(e.an_b + ob);
                }
                xr[oa] = xa;
                xr[ob] = xb;
              }
            }
          }
        }
      } else if (kind == OP_FIN) {
        // relu(final conv) over this thread's 64 columns, in place over relu(skip sum) in the staging tile (the final conv's MMAs have
        // completed: tmem_full) -- the A operand of the zero conv
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.final_bias + cbeg + 8 * k));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(a.final_bias + cbeg + 8 * k + 4));
          const uint32_t p0 = pack16(fmaxf(__uint_as_float(v[8 * k]) + b0.x, 0.f), fmaxf(__uint_as_float(v[8 * k + 1]) + b0.y, 0.f), fp16);
          const uint32_t p1 = pack16(fmaxf(__uint_as_float(v[8 * k + 2]) + b0.z, 0.f), fmaxf(__uint_as_float(v[8 * k + 3]) + b0.w, 0.f), fp16);
          const uint32_t p2 = pack16(fmaxf(__uint_as_float(v[8 * k + 4]) + b1.x, 0.f), fmaxf(__uint_as_float(v[8 * k + 5]) + b1.y, 0.f), fp16);
          const uint32_t p3 = pack16(fmaxf(__uint_as_float(v[8 * k + 6]) + b1.z, 0.f), fmaxf(__uint_as_float(v[8 * k + 7]) + b1.w, 0.f), fp16);
          sts128(stg_u32 + tile_off(r, cbeg + 8 * k), make_uint4(p0, p1, p2, p3));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive_remote_release(su_full + 1, lead_cta);
      } else {
        const bool is_res = kind == OP_RES;
        const float* bias = a.rs_bias + ((!is_res && a.has_res) ? 256 : 0) + cbeg;
        const bool relu = !is_res && a.relu;
        // the staging tile: pre-loaded input (residual: h_in; last layer: running skip sum) or, for the skip op after a residual op,
        // this warp's own region once its previous store has read it
        if (have_in) mbar_wait(stg_full, (uint32_t)it & 1);
        if (lane == 0 && store_pending) {
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          store_pending = false;
        }
        __syncwarp();
        if (!(a.dbg & 4))
#pragma unroll
        for (int k = 0; k < 8; ++k) {   // 8 columns = one 16-byte chunk of the row
          const uint32_t addr = stg_u32 + tile_off(r, cbeg + 8 * k);
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + 8 * k));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + 8 * k + 4));
          float y[8] = {__uint_as_float(v[8 * k]) + b0.x,     __uint_as_float(v[8 * k + 1]) + b0.y, __uint_as_float(v[8 * k + 2]) + b0.z,
                        __uint_as_float(v[8 * k + 3]) + b0.w, __uint_as_float(v[8 * k + 4]) + b1.x, __uint_as_float(v[8 * k + 5]) + b1.y,
                        __uint_as_float(v[8 * k + 6]) + b1.z, __uint_as_float(v[8 * k + 7]) + b1.w};
          if (have_in) {
            const uint4 in4 = lds128(addr);
            const uint32_t hu[4] = {in4.x, in4.y, in4.z, in4.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float lo, hi;
              unpack16(hu[q], lo, hi, fp16);
              y[2 * q] += lo;
              y[2 * q + 1] += hi;
            }
          }
          uint32_t pk[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float lo = y[2 * q], hi = y[2 * q + 1];
            if (is_res) {   // h_out = (h_in + res) * sqrt(.5)   (modules.py:128)
              lo *= 0.70710678118654752440f;
              hi *= 0.70710678118654752440f;
            } else if (relu) {
              lo = fmaxf(lo, 0.f);
              hi = fmaxf(hi, 0.f);
            }
            pk[q] = pack16(lo, hi, fp16);
          }
          sts128(addr, make_uint4(pk[0], pk[1], pk[2], pk[3]));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (tail) {   // relu(skip sum) stays in the staging tile: the A operand of the final conv
          if (lane == 0) mbar_arrive_remote_release(su_full, lead_cta);
        } else if (lane == 0) {
          // this warp's 32 rows x 64 columns leave as one TMA store (rows >= Ti and the dummy tile of an odd pair are clipped)
          if (!(a.dbg & 1))
            tma_store_3d(is_res ? &a.mapOutH : &a.mapOutS, stg + (size_t)qtr * A_BYTES + (size_t)lg * 32 * 128, qtr * 64, t0 + lg * 32, ub);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          if (is_res) {
            store_pending = true;   // the skip op follows at once and waits before it overwrites the region
          } else {                  // last 1x1 op of the tile: hand the staging tile back to the producer
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            mbar_arrive(stg_empty);
          }
        }
        __syncwarp();
      }
      if (warp == 2) L_TRACE(7);
    });
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (tail && qtr == 0 && !a.ez.reverse && a.ez.logdet_acc) {
      ls_sum = warp_sum(ls_sum);
      if (lane == 0 && ls_sum != 0.0) atomicAdd(a.ez.logdet_acc, ls_sum);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();   // no CTA exits (or frees TMEM) while its peer's MMAs / arrives may still target it
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc_2sm<512>(tmem_base);
  }
}

int launch_layer(const LayerArgs& a0, cudaStream_t st) {
  LayerArgs a = a0;
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("FWN_LAYER_DBG");
    dbg = e ? atoi(e) : 0;
  }
  a.dbg = dbg;
  static long long* trace_buf = nullptr;
  static int trace_state = -1;   // -1 unknown, 0 off, 1 armed, 2 printed
  if (trace_state < 0) {
    const char* e = getenv("FWN_LAYER_TRACE");
    trace_state = (e && e[0] == '1') ? 1 : 0;
    if (trace_state == 1) FWN_CUDA(cudaMalloc(&trace_buf, (L_TRACE_TILES * 6 * L_TRACE_K + L_TRACE_TILES) * sizeof(long long)));
  }
  static int printed[2] = {0, 0};   // one residual layer and one last layer
  static int trace_skip = -1, seen[2] = {0, 0};
  if (trace_skip < 0) {
    const char* e = getenv("FWN_LAYER_TRACE_SKIP");
    trace_skip = e ? atoi(e) : 0;
  }
  bool tracing = trace_state == 1 && a.B * a.tiles_per_utt >= 8000 && !printed[a.has_res ? 1 : 0];
  if (tracing && seen[a.has_res ? 1 : 0]++ < trace_skip) tracing = false;
  if (tracing) {
    FWN_CUDA(cudaMemsetAsync(trace_buf, 0, (L_TRACE_TILES * 6 * L_TRACE_K + L_TRACE_TILES) * sizeof(long long), st));
    a.trace = trace_buf;
  }
  static bool configured = false;
  if (!configured) {
    FWN_CUDA(cudaFuncSetAttribute(layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L_SMEM));
    configured = true;
  }
  const int num_m = a.B * a.tiles_per_utt;
  const int pairs = (num_m + 1) / 2;
  const int grid = 2 * std::min(pairs, num_sms() / 2);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(L_THREADS);
  cfg.dynamicSmemBytes = L_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = 2;
  attr[na].val.clusterDim.y = 1;
  attr[na].val.clusterDim.z = 1;
  ++na;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  FWN_CUDA(cudaLaunchKernelEx(&cfg, layer_kernel, a));
  FWN_LAUNCH_CHECK();
  if (tracing) {   // diagnostics only (run with FWN_GRAPH=0): timeline of the first tiles of cluster 0, cycles relative to the first event
    long long h[L_TRACE_TILES * 6 * L_TRACE_K + L_TRACE_TILES];
    FWN_CUDA(cudaStreamSynchronize(st));
    FWN_CUDA(cudaMemcpy(h, trace_buf, sizeof(h), cudaMemcpyDeviceToHost));
    const long long t00 = h[0];
    {
      const long long* gt = h + L_TRACE_TILES * 6 * L_TRACE_K;
      const double cyc = (double)(h[((L_TRACE_TILES - 1) * 6 + OP_G0) * L_TRACE_K + 1] - h[OP_G0 * L_TRACE_K + 1]);
      const double ns = (double)(gt[L_TRACE_TILES - 1] - gt[0]);
      fprintf(stderr, "layer trace: tiles %d..%d of cluster 0: %.0f cycles in %.0f ns = %.3f GHz, %.0f cycles / %.2f us per tile\n", L_TRACE_T0,
              L_TRACE_T0 + L_TRACE_TILES - 1, cyc, ns, cyc / ns, cyc / (L_TRACE_TILES - 1), ns / (L_TRACE_TILES - 1) * 1e-3);
    }
    fprintf(stderr, "layer trace: has_res=%d nseg=%d chunks=%d+%d+%d+%d  [mma: wait_tmem_empty, go, o_lo, o_hi, issued | epi: wait_full, full, done]\n", a.has_res,
            a.nseg, a.nchunk[0], a.nchunk[1], a.nchunk[2], a.nchunk[3]);
    for (int it = 0; it < L_TRACE_TILES; ++it)
      for (int opi = 0; opi < 6; ++opi) {
        if ((opi == 2 && !a.has_res) || (opi >= 4 && !a.tail)) continue;
        fprintf(stderr, "  tile %d op %d:", it, opi);
        for (int k = 0; k < L_TRACE_K; ++k) {
          const long long v = h[(it * 6 + opi) * L_TRACE_K + k];
          fprintf(stderr, " %8lld", v ? v - t00 : -1);
        }
        fprintf(stderr, "\n");
      }
    printed[a.has_res ? 1 : 0] = 1;
  }
  return 0;
}

}  // namespace tc
}  // namespace fwn

// Inline-PTX wrappers shared by the tcgen05 engines (sm_100a): mbarrier, TMA (cp.async.bulk.tensor), tcgen05.mma / commit /
// ld / alloc, CTA-pair (cta_group::2) variants, shared-memory matrix descriptors.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace fwn {
namespace tc {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2;  // one [128 x 64] bf16 operand tile: 16 KB

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(smem_u32(smem)),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// Multicast load: ONE request fetches the box from L2 and writes it to the same shared-memory offset of every CTA in `cta_mask`
// of the cluster, signalling complete_tx on the barrier at the same offset in each of them (each CTA posts its own expect_tx).
__device__ __forceinline__ void tma_load_3d_mc(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(
          smem_u32(smem)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(cta_mask)
      : "memory");
}
// ---- CTA-pair (cta_group::2) variants.  Both CTAs issue their own loads into their own shared memory, but the
// transaction bytes complete on the LEADER's (even CTA's) barrier: clearing bit 24 of a shared::cluster address selects
// the even CTA of the pair at the same offset.
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_3d_2sm(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem)),
      "l"(map), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem)),
      "l"(map), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1)
      : "memory");
}
// plain arrive on the barrier at this offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Programmatic dependent launch.  A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while its
// predecessor in the stream is still draining: everything before pdl_wait() (barrier init, TMEM allocation, tensor-map prefetch,
// bias / resident-weight staging -- none of which touches data the predecessor produces) overlaps the predecessor's tail;
// pdl_wait() returns once the predecessor grid has completed and its writes are visible.  Without the launch attribute both are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
// Issue the (up to four) K=16 MMAs of one 64-wide K chunk and commit them to `bar`, from ONE convergent asm block predicated
// by elect.sync.  Issuing each tcgen05.mma from inside a divergent `if (lane == 0)` makes ptxas wrap every UTCHMMA in an
// ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall: ~100 dependent instructions per chunk on one warp, which was slower than the
// 512 cycles of tensor work it feeds.  Descriptors advance by 32 bytes (+2 in 16-byte units) per K step inside the swizzle atom.
template <bool PAIR>
__device__ __forceinline__ void umma_chunk_commit(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                                  uint32_t ksteps, uint32_t bar) {
  if (PAIR) {
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
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(ksteps), "r"(bar), "h"((uint16_t)3)
        : "memory");
  } else {
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
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pacc;\n"
        "add.u64 da, %1, 2;\n add.u64 db, %2, 2;\n"
        "@p1 tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n"
        "add.u64 da, %1, 4;\n add.u64 db, %2, 4;\n"
        "@p2 tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n"
        "add.u64 da, %1, 6;\n add.u64 db, %2, 6;\n"
        "@p3 tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(ksteps), "r"(bar)
        : "memory");
  }
}
// Same, single-CTA, WITHOUT the commit: several operand pairs (the bf16x3 split terms) accumulate into one tile before one commit.
__device__ __forceinline__ void umma_chunk(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate, uint32_t ksteps) {
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
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pacc;\n"
      "add.u64 da, %1, 2;\n add.u64 db, %2, 2;\n"
      "@p1 tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n"
      "add.u64 da, %1, 4;\n add.u64 db, %2, 4;\n"
      "@p2 tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n"
      "add.u64 da, %1, 6;\n add.u64 db, %2, 6;\n"
      "@p3 tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(ksteps)
      : "memory");
}
// commit only (accumulator-complete signal), convergent + elect-predicated
template <bool PAIR>
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
  if (PAIR) {
    asm volatile(
        "{\n.reg .pred pe;\nelect.sync _|pe, 0xffffffff;\n"
        "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n}\n" ::"r"(bar),
        "h"((uint16_t)3)
        : "memory");
  } else {
    asm volatile(
        "{\n.reg .pred pe;\nelect.sync _|pe, 0xffffffff;\n"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(bar)
        : "memory");
  }
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte swizzle shared-memory matrix descriptor (rows of 64 bf16 = 128 B, 8-row groups 1024 B apart).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);  // start address, 16-byte units
  d |= (uint64_t)1 << 16;                      // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;            // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                      // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                      // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, M=128, N=BN
template <int BN>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
// same with fp16 operands (a_format = b_format = 0); the accumulator stays fp32
constexpr uint32_t IDESC_BF16_BITS = (1u << 7) | (1u << 10);

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void unpack_bf16(uint32_t u, float& lo, float& hi) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  lo = __low2float(v);
  hi = __high2float(v);
}
// 16-bit activation storage of the mixed modes: bf16, or fp16 when `fp16` (a launch-uniform flag) is set
__device__ __forceinline__ uint32_t pack16(float lo, float hi, bool fp16) {
  if (fp16) {
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  return pack_bf16(lo, hi);
}
__device__ __forceinline__ void unpack16(uint32_t u, float& lo, float& hi, bool fp16) {
  if (fp16) {
    const float2 f = __half22float2(*reinterpret_cast<__half2*>(&u));
    lo = f.x;
    hi = f.y;
  } else {
    unpack_bf16(u, lo, hi);
  }
}
// tanh(f) * sigmoid(g) to fp32 accuracy from 2 ex2 + 1 rcp (the fp16 mode's gate: tanh.approx's 2^-11 absolute error would
// otherwise dominate the 2^-12 relative rounding of fp16 storage):  (E - 1) / ((E + 1)(1 + G)),  E = e^{2f}, G = e^{-g}
__device__ __forceinline__ float gate_accurate(float f, float g) {
  f = fminf(fmaxf(f, -15.f), 15.f);       // tanh saturates to +-1 in fp32 beyond |f| ~ 9; keeps E finite
  g = fmaxf(g, -60.f);                    // G <= e^60: (E+1)(1+G) stays finite
  float E, G;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(E) : "f"(f * 2.8853900817779268f));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(G) : "f"(g * -1.4426950408889634f));
  return __fdividef(E - 1.f, (E + 1.f) * (1.f + G));
}


// fp32 pair -> three packed bf16 pairs with x = a1 + a2 + a3 (round-to-nearest pieces: |a2| <= 2^-9 |x|, |a3| <= 2^-18 |x|; unbiased).
// Truncating instead of rounding would be cheaper but biases every piece towards zero, and the bias survives the cancellation in
// gradient sums (measured: 5% error on a weight gradient with 3 product terms).
__device__ __forceinline__ void split3_pair(float x0, float x1, uint32_t& p1, uint32_t& p2, uint32_t& p3) {
  p1 = pack_bf16(x0, x1);
  const float r0 = x0 - __uint_as_float(p1 << 16), r1 = x1 - __uint_as_float(p1 & 0xFFFF0000u);
  p2 = pack_bf16(r0, r1);
  p3 = pack_bf16(r0 - __uint_as_float(p2 << 16), r1 - __uint_as_float(p2 & 0xFFFF0000u));
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

}  // namespace tc
}  // namespace fwn

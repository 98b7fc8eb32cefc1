// tcgen05 / TMEM / TMA implicit-GEMM engine (bf16 operands, fp32 accumulate) -- the throughput mode.
//
//   acc[128 rows, BN cols] (TMEM, fp32) = sum over K segments of  A_seg[rows shifted in time, 64-wide K chunk] x W[K chunk, BN]
//
// * A operand: activations [B, T_i, C] bf16 in HBM, fetched by TMA through a 3-D tensor map (C, T_i, B) with a
//   (64, 128, 1) box.  A dilated-conv tap is the same load at time coordinate t0 + (k-1)*d; TMA's out-of-bounds
//   zero fill IS the tf.pad of modules.py:27 and can never bleed across utterances.
// * B operand: prepacked weights [N, Kpad] bf16 (K-major), 2-D tensor map, (64, BN) box.
// * Both land in shared memory in the 128-byte-swizzled K-major layout tcgen05.mma consumes directly.
// * One elected thread issues tcgen05.mma (M=128, N=BN, K=16) into one of two TMEM accumulator stages, so the
//   epilogue of tile i overlaps the MMAs of tile i+1.  tcgen05.commit releases shared-memory stages / publishes
//   accumulators through mbarriers.
// * Epilogue warps read TMEM with tcgen05.ld (each thread owns one output row) and apply the fused flow op:
//   tanh*sigmoid gate, residual + skip accumulation, bias+ReLU, or ActNorm + affine coupling + log-det in place on x.
// Persistent grid (<= one CTA per SM), warp-specialised: warp 0 TMA producer, warp 1 MMA issuer + TMEM owner,
// warps 2-9 epilogue: two groups of four warps (one per TMEM lane group) that ALTERNATE tiles, each group owning one TMEM
// accumulator stage and one staging buffer, so the latency chains of two consecutive tiles overlap.  Tiles move through
// 128B-swizzled shared memory: residual / skip inputs arrive by TMA bulk loads, outputs leave by per-warp TMA bulk stores.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include <unordered_map>

#include "common.cuh"
#include "layer_tc.cuh"
#include "tail_tc.cuh"
#include "model.h"
#include "tc_ptx.cuh"

namespace fwn {
namespace tc {

constexpr int MAX_SEG = 4;

struct alignas(64) TcArgs {
  CUtensorMap mapA[MAX_SEG];
  CUtensorMap mapB;
  CUtensorMap mapIn[2];      // RES_SKIP: [0] residual input h_in, [1] running skip sum (same (64,128,1) swizzled boxes)
  CUtensorMap mapOut[2];     // GATE: [0]=o; RES_SKIP: [0]=h_out, [1]=skip; PLAIN: [0]=y
  int has_in[2];
  int shift[MAX_SEG];
  int nchunk[MAX_SEG];       // 64-wide K chunks in the segment
  int last_ksteps[MAX_SEG];  // valid 16-wide MMA steps in the segment's last chunk (1..4)
  int wk0[MAX_SEG];          // first W column (k) of the segment
  int nseg;
  int multicast;             // weight map has (64, BN/2) boxes and the kernel runs as cta_group::2 CTA pairs
  int fp16;                  // operands / activations are fp16 instead of bf16 (FWN_MIXED_FP16)
  int B, Ti, tiles_per_utt, n_tiles, N;
  EpiArgs e;
};

// Per-(epilogue, tile-width) configuration.  STAGED epilogues move their tiles through 128B-swizzled shared memory:
// outputs leave with TMA bulk stores and the residual / running-skip inputs arrive with TMA bulk loads, so no
// warp ever issues row-strided (32 lines per instruction) global accesses.
// WS (weight-stationary): the CTA owns ONE column tile for its whole life; that tile's weights (K <= 256, i.e. <= 4
// chunks) are loaded once into shared memory and only activations stream through the (A-only, deeper) pipeline.
// Used by the K=256 1x1 GEMMs (res|skip, final) whose per-tile weight traffic otherwise dwarfs their MMA time.
// PAIR: CTAs are launched as 2-CTA clusters (one SM pair) working on two neighbouring row tiles of the SAME column tile
// with ONE tcgen05.mma.cta_group::2 (M=256): each CTA stages its own A tile and only HALF of the weight box, which halves
// both the L2->SM weight traffic and -- the actual limiter of the single-CTA kernel -- the shared-memory bandwidth per MMA
// (TMA fill + operand read drop from 192 to 128 bytes per cycle per SM).
template <int EPI, int BN, bool WS, bool PAIR = false>
struct Cfg {
  static constexpr bool STAGED = (EPI == EPI_GATE) || (EPI == EPI_RES_SKIP) || (EPI == EPI_LINEAR) || (EPI == EPI_PLAIN && BN == 128);
  // staging tile is TMA-loaded, updated in place, TMA-stored.  LINEAR (backward pass of the 16-bit training mode) is RES_SKIP with
  // every column a "residual" column: y = alpha (acc + in0), or y = in0 > 0 ? alpha acc : 0 when in0 is a ReLU mask
  static constexpr bool IN_PLACE = (EPI == EPI_RES_SKIP) || (EPI == EPI_LINEAR);
  static constexpr int OUT_COLS = (EPI == EPI_GATE) ? BN / 2 : BN;        // bf16 output columns per tile
  static constexpr int SUBTILES = STAGED ? OUT_COLS / 64 : 0;             // [128 rows x 64 cols] 16 KB boxes
  static constexpr int STG_BYTES = SUBTILES * BM * 128;
  // staging buffers: one per epilogue group (the two groups of four warps alternate tiles); in-place tiles use a ring of 3
  // so a buffer is refilled a full tile after the group that stored from it moved on (lazy release, nobody waits on a
  // store it just issued)
  static constexpr int NSTG = STAGED ? (IN_PLACE ? 3 : 2) : 0;
  // epilogue organisation: GROUPS groups of WPG warps; the groups alternate tiles (each owns one TMEM stage and one staging
  // buffer) so the latency chains of consecutive tiles overlap; with WPG == 8 the two warps of a TMEM lane group split the
  // tile's columns.  The MUFU-heavy gate epilogue is issue-latency-bound and gets 16 warps, the others 8.
  static constexpr int GROUPS = 2;
  static constexpr int WPG = (EPI == EPI_GATE) ? 8 : 4;
  static constexpr int THREADS = 64 + 32 * GROUPS * WPG;
  static constexpr int LDW_MAX = (EPI == EPI_GATE) ? 32 : 64;   // accumulator columns fetched per tcgen05.wait::ld
  static constexpr int BIAS_BYTES = WS ? BN * 4 : 2048;   // bias in smem: the CTA's own column tile (WS) or all <= 512 columns
  static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * BK * 2;   // per-CTA share of the weight box
  static constexpr int WS_CHUNKS = 4;
  static constexpr int W_BYTES = WS ? WS_CHUNKS * B_BYTES : 0;
  static constexpr int STAGE_BYTES = WS ? A_BYTES : A_BYTES + B_BYTES;
  static constexpr int BUDGET = 223 * 1024 - NSTG * STG_BYTES - W_BYTES;
  static constexpr int STAGES = BUDGET / STAGE_BYTES > 8 ? 8 : BUDGET / STAGE_BYTES;
  // TMEM accumulator stages: 4 where they fit (BN <= 128) so the MMA warp can run several tiles ahead of the epilogue groups
  static constexpr int NACC = (BN <= 128) ? 4 : 2;
  static constexpr int TMEM_COLS = NACC * BN < 32 ? 32 : NACC * BN;
  static constexpr size_t SMEM = 1024 + (size_t)STAGES * STAGE_BYTES + (size_t)W_BYTES + (size_t)NSTG * STG_BYTES + BIAS_BYTES + 320;
  static_assert(STAGES >= 2, "pipeline needs at least two stages");
  static_assert(SMEM <= 227 * 1024, "shared memory budget exceeded");
};

// byte offset of the 16-byte chunk holding columns [c, c+8) of row r inside a staged tile (TMA SWIZZLE_128B layout)
__device__ __forceinline__ uint32_t stg_off(int r, int c) {
  return (uint32_t)((c >> 6) * (BM * 128) + r * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4));
}

// ---------------------------------------------------------------- epilogue on 16 accumulator columns of one row
// `stg` = this tile's staging buffer (STAGED kinds); for RES_SKIP it already holds the residual input / running skip sum.
template <int EPI, int BN, bool WS>
__device__ __forceinline__ void epilogue16(const TcArgs& a, int64_t row, bool row_ok, int r, int n_tile, int c0, const uint32_t* v,
                                           uint8_t* stg, bool have_in, const uint4* inp, const float* sbias, double& ls_sum) {
  using C = Cfg<EPI, BN, WS, false>;
  const EpiArgs& e = a.e;
  const bool fp16 = a.fp16 != 0;
  const int col = n_tile * BN + c0;  // global column of v[0]
  float acc[16];
  if (EPI == EPI_PLAIN_F32) {   // no bias (and N may exceed the staged bias range)
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(v[j]);
  } else {  // bias from shared memory (broadcast 128-bit reads): WS kernels stage their own column tile, the others all columns
    const float* bp = sbias + (WS ? c0 : col);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 b4 = *reinterpret_cast<const float4*>(bp + 4 * j);
      acc[4 * j] = __uint_as_float(v[4 * j]) + b4.x;
      acc[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b4.y;
      acc[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b4.z;
      acc[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b4.w;
    }
  }
  if (EPI == EPI_GATE) {
    // (2c, 2c+1) = (filter_c, gate_c): o = tanh(f) * sigmoid(g)   (modules.py:124); 16 columns -> 8 channels = one 16-byte chunk
    if (e.in0 && row_ok) {   // deep blocks: the conditioning projection of this layer was computed ahead (fp32 [rows, ld])
      const float4* pp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(e.in0) + row * e.ld + col);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 q = __ldg(pp + k);
        acc[4 * k] += q.x; acc[4 * k + 1] += q.y; acc[4 * k + 2] += q.z; acc[4 * k + 3] += q.w;
      }
    }
    uint32_t p[4];
    if (fp16) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        p[j] = pack16(gate_accurate(acc[4 * j], acc[4 * j + 1]), gate_accurate(acc[4 * j + 2], acc[4 * j + 3]), true);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // tanh(f) * sigmoid(g) = t + t*tanh(g/2) with t = tanh(f)/2 : 2 MUFU + 3 FP32 ops per output
        const float t0 = 0.5f * tanh_fast(acc[4 * j]), t1 = 0.5f * tanh_fast(acc[4 * j + 2]);
        const float h0 = tanh_fast(0.5f * acc[4 * j + 1]), h1 = tanh_fast(0.5f * acc[4 * j + 3]);
        const float o0 = fmaf(h0, t0, t0);
        const float o1 = fmaf(h1, t1, t1);
        p[j] = pack_bf16(o0, o1);
        if (e.tape) acc[4 * j] = fmaf(h0, 0.5f, 0.5f), acc[4 * j + 2] = fmaf(h1, 0.5f, 0.5f);   // sigmoid(g), kept for the backward pass
      }
      if (e.tape && row_ok)   // training forward: 8 sigmoids = 16 bytes of this thread's row
        *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(e.tape) + row * e.F + col / 2) =
            make_uint4(pack_bf16(acc[0], acc[2]), pack_bf16(acc[4], acc[6]), pack_bf16(acc[8], acc[10]), pack_bf16(acc[12], acc[14]));
    }
    sts128(smem_u32(stg) + stg_off(r, c0 / 2), make_uint4(p[0], p[1], p[2], p[3]));
  } else if (EPI == EPI_RES_SKIP) {
    const uint32_t sbase = smem_u32(stg);
    const bool is_res = e.has_res && col < e.F;
    if (have_in) {   // residual input / running skip sum: this thread's two 16-byte chunks, pre-loaded from the staging tile
      const uint4 h0 = inp[0], h1 = inp[1];
      const uint32_t hu[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float lo, hi;
        unpack16(hu[j], lo, hi, fp16);
        acc[2 * j] += lo;
        acc[2 * j + 1] += hi;
      }
    }
    uint32_t p[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float lo = acc[2 * j], hi = acc[2 * j + 1];
      if (is_res) {  // h_out = (h_in + res) * sqrt(.5)   (modules.py:128)
        lo *= 0.70710678118654752440f;
        hi *= 0.70710678118654752440f;
      } else if (e.relu) {  // last layer: relu(sum of skips) feeds Conv_final (modules.py:176-177)
        lo = fmaxf(lo, 0.f);
        hi = fmaxf(hi, 0.f);
      }
      p[j] = pack16(lo, hi, fp16);
    }
    sts128(sbase + stg_off(r, c0), make_uint4(p[0], p[1], p[2], p[3]));
    sts128(sbase + stg_off(r, c0 + 8), make_uint4(p[4], p[5], p[6], p[7]));
  } else if (EPI == EPI_LINEAR) {
    const uint32_t sbase = smem_u32(stg);
    float in[16];
    if (have_in) {
      const uint4 h0 = inp[0], h1 = inp[1];
      const uint32_t hu[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) unpack16(hu[j], in[2 * j], in[2 * j + 1], fp16);
    }
    uint32_t p[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float lo = acc[2 * j], hi = acc[2 * j + 1];
      if (have_in && !e.mask_mode) { lo += in[2 * j]; hi += in[2 * j + 1]; }
      lo *= e.alpha;
      hi *= e.alpha;
      if (have_in && e.mask_mode) {
        lo = in[2 * j] > 0.f ? lo : 0.f;
        hi = in[2 * j + 1] > 0.f ? hi : 0.f;
      }
      p[j] = pack16(lo, hi, fp16);
    }
    sts128(sbase + stg_off(r, c0), make_uint4(p[0], p[1], p[2], p[3]));
    sts128(sbase + stg_off(r, c0 + 8), make_uint4(p[4], p[5], p[6], p[7]));
  } else if (EPI == EPI_PLAIN_F32) {
    if (!row_ok) return;
    float* o = reinterpret_cast<float*>(e.out0) + row * e.ld + col;
    // accumulate (y += acc: the conditioning gradient, one update per layer): vector reductions instead of load-add-store -- every
    // element is updated by exactly one thread per launch, so the sum is the same fp32 addition, but the epilogue no longer waits
    // ~1 us per 16-column group for the old values (these launches averaged 37 us for a K = 512 GEMM)
    if (col + 15 < a.N && (e.ld & 3) == 0) {
      float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (e.accum)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o4 + k), "f"(acc[4 * k]), "f"(acc[4 * k + 1]), "f"(acc[4 * k + 2]),
                       "f"(acc[4 * k + 3])
                       : "memory");
        else
          o4[k] = make_float4(acc[4 * k], acc[4 * k + 1], acc[4 * k + 2], acc[4 * k + 3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (col + j < a.N) {
          if (e.accum) atomicAdd(o + j, acc[j]);
          else o[j] = acc[j];
        }
    }
  } else if (EPI == EPI_PLAIN) {
    uint32_t p[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float lo = acc[2 * j], hi = acc[2 * j + 1];
      if (e.relu) { lo = fmaxf(lo, 0.f); hi = fmaxf(hi, 0.f); }
      p[j] = pack16(lo, hi, fp16);
    }
    if (C::STAGED) {
      sts128(smem_u32(stg) + stg_off(r, c0), make_uint4(p[0], p[1], p[2], p[3]));
      sts128(smem_u32(stg) + stg_off(r, c0 + 8), make_uint4(p[4], p[5], p[6], p[7]));
    } else if (row_ok) {
      if (col + 15 < a.N) {
        uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(e.out0) + row * e.ld + col);
        op[0] = make_uint4(p[0], p[1], p[2], p[3]);
        op[1] = make_uint4(p[4], p[5], p[6], p[7]);
      } else {
        uint16_t* o = reinterpret_cast<uint16_t*>(e.out0) + row * e.ld + col;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (col + j < a.N) o[j] = (uint16_t)(pack16(e.relu ? fmaxf(acc[j], 0.f) : acc[j], 0.f, fp16) & 0xFFFFu);
      }
    }
  } else if (EPI == EPI_AFFINE) {
    // (2q, 2q+1) = (log_s, t) of transformed element q; ActNorm + coupling in place on the fp32 flow variable
    if (!row_ok) return;
    float* xr = e.X + row * e.Cx;
    const int q0 = col / 2;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const int q = q0 + p;
      if (q >= e.nq) break;
      const float log_s = acc[2 * p], tt = acc[2 * p + 1];
      if (e.out1) *reinterpret_cast<float2*>(reinterpret_cast<float*>(e.out1) + row * e.ld + 2 * q) = make_float2(log_s, tt);   // tape
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

// AFFINE fast path on registers: 16 accumulator columns = 8 adjacent (pass-through, transformed) pairs = 16 consecutive floats
// of the thread's x row, already loaded into x4[0..3]; ActNorm + coupling applied in place in registers.
__device__ __forceinline__ void affine16_regs(const EpiArgs& e, const float* sbias, int col, const uint32_t* v, float4* x4, bool row_ok,
                                              double& ls_sum, int64_t row) {
  const float4* bp = reinterpret_cast<const float4*>(e.an_b + col);
  const float4* sp = reinterpret_cast<const float4*>(e.an_s + col);
  const int bo = e.b_odd;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float4 b4 = __ldg(bp + k), s4 = __ldg(sp + k);
    const float4 bias4 = *reinterpret_cast<const float4*>(sbias + col + 4 * k);
    float x[4] = {x4[k].x, x4[k].y, x4[k].z, x4[k].w};
    const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w};
    const float ac[4] = {__uint_as_float(v[4 * k]) + bias4.x, __uint_as_float(v[4 * k + 1]) + bias4.y,
                         __uint_as_float(v[4 * k + 2]) + bias4.z, __uint_as_float(v[4 * k + 3]) + bias4.w};
    if (e.out1 && row_ok)   // training forward: (log_s, t) pairs of this row, in column order (the tape of affine_bwd)
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out1) + row * e.ld + col + 4 * k) = make_float4(ac[0], ac[1], ac[2], ac[3]);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float log_s = ac[2 * h], tt = ac[2 * h + 1];
      const int ib = 2 * h + bo, ia = 2 * h + 1 - bo;
      if (!e.reverse) {
        x[ia] = (x[ia] + bb[ia]) * ss[ia];
        x[ib] = ((x[ib] + bb[ib]) * ss[ib] - tt) * __expf(-log_s);
        if (row_ok) ls_sum += (double)log_s;   // rows past the utterance end see bias-only accumulators: not part of the log-det
      } else {
        x[ib] = (x[ib] * __expf(log_s) + tt) * ss[ib] - bb[ib];
        x[ia] = x[ia] * ss[ia] - bb[ia];
      }
    }
    x4[k] = make_float4(x[0], x[1], x[2], x[3]);
  }
}

// ---------------------------------------------------------------- the kernel
// CL > 1 (weight-stationary kernels only): the CL CTAs of a cluster own the CL column tiles of the SAME row tiles.  Their A operand
// is identical, so the leader fetches every activation chunk once and TMA-multicasts it into all CL shared memories: L2->SM
// activation traffic drops by CL (the K = 256 GEMMs re-read their A tile once per column tile and were bound by the L2 fabric).
template <int EPI, int BN, bool WS, bool PAIR, int CL = 1>
__global__ void __launch_bounds__((Cfg<EPI, BN, WS, PAIR>::THREADS), 1) tc_gemm_kernel(const __grid_constant__ TcArgs a) {
  using C = Cfg<EPI, BN, WS, PAIR>;
  static_assert(!(WS && PAIR), "the CTA-pair variant is for the streaming kernel");
  static_assert(CL == 1 || WS, "activation multicast is for the weight-stationary kernels");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint8_t* w_base = smem + (size_t)C::STAGES * C::STAGE_BYTES;   // WS: resident weight chunks
  uint8_t* stg_base = w_base + C::W_BYTES;
  float* sbias = reinterpret_cast<float*>(stg_base + (size_t)C::NSTG * C::STG_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sbias) + C::BIAS_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tmem_full = empty_bar + C::STAGES;
  uint64_t* tmem_empty = tmem_full + 4;
  uint64_t* in_full = tmem_empty + 4;   // staged inputs landed (RES_SKIP)
  uint64_t* in_empty = in_full + 3;     // staging buffer may be overwritten by the next input load
  uint64_t* w_full = in_empty + 3;      // WS: resident weights landed
  uint64_t* cl_empty = w_full + 1;      // CL > 1, leader: every CTA of the cluster has freed the stage and expects the next chunk
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(cl_empty + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m_tiles = a.B * a.tiles_per_utt;
  // tile schedule.  streaming: tile = blockIdx.x + it*gridDim.x over (m, n) with n fastest.
  // weight-stationary: n fixed per CTA, m strides by gridDim.x / n_tiles (host guarantees gridDim.x % n_tiles == 0).
  const int ws_groups = WS ? (int)gridDim.x / a.n_tiles : 1;
  const int ws_n = WS ? (int)blockIdx.x % a.n_tiles : 0;
  const int ws_m0 = WS ? (int)blockIdx.x / a.n_tiles : 0;
  const int mc_rank = (PAIR || CL > 1) ? (int)cluster_ctarank() : 0;
  const bool leader = mc_rank == 0;
  auto tile_of = [&](int it, int& m_tile, int& n_tile) -> bool {
    if (PAIR) {  // cluster c handles row-tile pairs; an odd tail pair's second tile is a dummy (TMA zero fill in, clipped out)
      const int p = (int)(blockIdx.x >> 1) + it * (int)(gridDim.x >> 1);
      const int m_pair = p / a.n_tiles;
      n_tile = p - m_pair * a.n_tiles;
      m_tile = 2 * m_pair + mc_rank;
      return m_pair < (num_m_tiles + 1) / 2;
    }
    if (WS) {
      m_tile = ws_m0 + it * ws_groups;
      n_tile = ws_n;
      return m_tile < num_m_tiles;
    }
    const int tile = (int)blockIdx.x + it * (int)gridDim.x;
    m_tile = tile / a.n_tiles;
    n_tile = tile - m_tile * a.n_tiles;
    return tile < num_m_tiles * a.n_tiles;
  };

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < a.nseg; ++s) prefetch_tmap(&a.mapA[s]);
    prefetch_tmap(&a.mapB);
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(full_bar + i, 1);
      mbar_init(empty_bar + i, 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(tmem_full + i, 1);
      mbar_init(tmem_empty + i, C::WPG * (PAIR ? 2 : 1));   // one arrive per warp draining this stage
    }
    for (int i = 0; i < 3; ++i) {
      mbar_init(in_full + i, 1);
      mbar_init(in_empty + i, 4);   // the four warps of the group that stored from this staging buffer
    }
    mbar_init(w_full, 1);
    for (int i = 0; i < 8; ++i) mbar_init(cl_empty + i, CL);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < C::BIAS_BYTES / 4; i += C::THREADS) {
    const int col = (WS ? ws_n * BN : 0) + i;
    sbias[i] = (col < a.N && a.e.bias) ? __ldg(a.e.bias + col) : 0.f;
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_2sm<C::TMEM_COLS>(tmem_ptr);
    else tmem_alloc<C::TMEM_COLS>(tmem_ptr);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (PAIR || CL > 1) cluster_sync_all();  // peer barriers are initialised before any remote arrive
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // the next kernel of the chain may be scheduled as SMs free up (its own prologue overlaps our tail) ...
  pdl_launch_dependents();

  // which input map (if any) feeds the epilogue of column tile n_tile (RES_SKIP only)
  auto input_map_of = [&](int n_tile) -> int {
    if (EPI != EPI_RES_SKIP && EPI != EPI_LINEAR) return -1;
    const int which = (a.e.has_res && n_tile * BN < a.e.F) ? 0 : 1;
    return a.has_in[which] ? which : -1;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int m_tile, n_tile;
      if (WS) {  // resident weights: every K chunk of this CTA's column tile, once
        int nch = 0;
        for (int s = 0; s < a.nseg; ++s) nch += a.nchunk[s];
        mbar_expect_tx(w_full, (uint32_t)nch * C::B_BYTES);
        int c = 0;
        for (int s = 0; s < a.nseg; ++s)
          for (int ch = 0; ch < a.nchunk[s]; ++ch, ++c)
            tma_load_2d(w_base + (size_t)c * C::B_BYTES, &a.mapB, w_full, a.wk0[s] + ch * BK, ws_n * BN);
      }
      pdl_wait();   // ... and we touch activations only once the previous kernel of the chain has completed
      for (int it = 0; tile_of(it, m_tile, n_tile); ++it) {
        const int ub = m_tile / a.tiles_per_utt;
        const int t0 = (m_tile - ub * a.tiles_per_utt) * BM;
        if (C::IN_PLACE) {
          // epilogue input tile -> staging buffer it % NSTG; the buffer is free once the TMA store that last used it has read it
          const int b = it % 3;
          mbar_wait(in_empty + b, ((it / 3) & 1) ^ 1);
          const int im = input_map_of(n_tile);
          if (im >= 0) {
            const int cin0 = (n_tile * BN) % a.e.F;  // column of this tile inside the [rows, F] input tensor
            mbar_expect_tx(in_full + b, C::STG_BYTES);
#pragma unroll
            for (int j = 0; j < C::SUBTILES; ++j)
              tma_load_3d(stg_base + (size_t)b * C::STG_BYTES + (size_t)j * BM * 128, &a.mapIn[im], in_full + b, cin0 + j * 64, t0, ub);
          } else {
            mbar_arrive(in_full + b);
          }
        }
        for (int s = 0; s < a.nseg; ++s) {
          for (int ch = 0; ch < a.nchunk[s]; ++ch) {
            mbar_wait(empty_bar + stage, phase ^ 1);
            uint8_t* sa = stage_base + (size_t)stage * C::STAGE_BYTES;
            if (PAIR) {
              // both CTAs' loads complete on the leader's full barrier, which therefore expects both shares
              if (leader) mbar_expect_tx(full_bar + stage, 2 * C::STAGE_BYTES);
              tma_load_3d_2sm(sa, &a.mapA[s], full_bar + stage, ch * BK, t0 + a.shift[s], ub);
              tma_load_2d_2sm(sa + A_BYTES, &a.mapB, full_bar + stage, a.wk0[s] + ch * BK, n_tile * BN + mc_rank * (BN / 2));
            } else if (CL > 1) {
              // this CTA's stage is free (waited above) and now expects the chunk; once all CL CTAs said so, the leader multicasts it
              mbar_expect_tx(full_bar + stage, C::STAGE_BYTES);
              mbar_arrive_remote(cl_empty + stage, 0);
              if (leader) {
                mbar_wait(cl_empty + stage, phase);
                tma_load_3d_mc(sa, &a.mapA[s], full_bar + stage, ch * BK, t0 + a.shift[s], ub, (uint16_t)((1u << CL) - 1));
              }
            } else {
              mbar_expect_tx(full_bar + stage, C::STAGE_BYTES);
              tma_load_3d(sa, &a.mapA[s], full_bar + stage, ch * BK, t0 + a.shift[s], ub);
              if (!WS) tma_load_2d(sa + A_BYTES, &a.mapB, full_bar + stage, a.wk0[s] + ch * BK, n_tile * BN);
            }
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_bf16 = PAIR ? (make_idesc<BN>() & ~(0x1Fu << 24)) | ((uint32_t)(256 >> 4) << 24) : make_idesc<BN>();
    const uint32_t idesc = a.fp16 ? (idesc_bf16 & ~IDESC_BF16_BITS) : idesc_bf16;
    int stage = 0, as = 0;
    uint32_t phase = 0, aphase = 0;
    int m_tile, n_tile;
    const uint32_t stage0 = smem_u32(stage_base), w0 = smem_u32(w_base);
    const uint64_t desc_hi = make_smem_desc(0);   // everything but the start-address field
    if (WS) {
      mbar_wait(w_full, 0);
      tcgen05_fence_after();
    }
    for (int it = 0; (!PAIR || leader) && tile_of(it, m_tile, n_tile); ++it) {   // in a cta_group::2 pair only the leader CTA issues MMAs
      mbar_wait(tmem_empty + as, aphase ^ 1);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
      uint32_t accumulate = 0;
      int c = 0;
      for (int s = 0; s < a.nseg; ++s) {
        const int nch = a.nchunk[s];
        const uint32_t last_ks = (uint32_t)a.last_ksteps[s];
        for (int ch = 0; ch < nch; ++ch, ++c) {
          mbar_wait(full_bar + stage, phase);
          tcgen05_fence_after();
          const uint32_t sa = stage0 + (uint32_t)stage * C::STAGE_BYTES;
          const uint32_t sb = WS ? w0 + (uint32_t)c * C::B_BYTES : sa + A_BYTES;
          const uint64_t adesc = desc_hi | (uint64_t)((sa >> 4) & 0x3FFF), bdesc = desc_hi | (uint64_t)((sb >> 4) & 0x3FFF);
          const uint32_t ksteps = (ch == nch - 1) ? last_ks : (uint32_t)(BK / UMMA_K);
          // all 32 lanes stay convergent; one elected lane issues the chunk's MMAs and the commit that frees the stage
          umma_chunk_commit<PAIR>(tmem_d, adesc, bdesc, idesc, accumulate, ksteps, smem_u32(empty_bar + stage));
          accumulate = 1;
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
      umma_commit_elect<PAIR>(smem_u32(tmem_full + as));   // accumulator complete -> epilogue (of both CTAs for a pair)
      if (++as == C::NACC) { as = 0; aphase ^= 1; }
    }
  } else {
    // ===================== epilogue (warps 2..9): two groups of four warps alternate tiles =====================
    const int lg = warp & 3;            // TMEM lane group this warp may access
    const int grp = (warp - 2) / C::WPG;                     // group g drains TMEM stage g (tiles it % 2 == g)
    const int half = C::WPG == 8 ? ((warp - 2) % 8) >> 2 : 0;   // WPG==8: which half of the tile's columns this warp owns
    constexpr int CWID = C::WPG == 8 ? BN / 2 : BN;          // accumulator columns per warp
    const int cbeg = half * CWID;
    const int r = lg * 32 + lane;
    double ls_sum = 0.0;
    int m_tile, n_tile;
    pdl_wait();   // epilogue warps read / write global memory themselves (x rows, clipped stores): same rule as the producer
    for (int it = grp; tile_of(it, m_tile, n_tile); it += C::GROUPS) {
      const int as = it % C::NACC;
      const uint32_t aphase = (uint32_t)(it / C::NACC) & 1;
      const int ub = m_tile / a.tiles_per_utt;
      const int t0 = (m_tile - ub * a.tiles_per_utt) * BM;
      const int t = t0 + r;
      const bool row_ok = t < a.Ti && ub < a.B;
      const int64_t row = (int64_t)ub * a.Ti + t;
      const int sb = C::IN_PLACE ? it % 3 : grp;
      uint8_t* stg = stg_base + (size_t)sb * C::STG_BYTES;
      bool have_in = false;
      if (C::STAGED) {
        // the store this warp issued from this staging slice (tile it-2) has had a whole tile to read it
        if (lane == 0) {
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          if (C::IN_PLACE && it >= 2) mbar_arrive(in_empty + ((it - 2) % 3));   // lazy release -> input of tile it+1 may land there
        }
        __syncwarp();
      }
      if (C::IN_PLACE) {
        have_in = input_map_of(n_tile) >= 0;
        mbar_wait(in_full + sb, (uint32_t)(it / 3) & 1);
      }
      mbar_wait(tmem_full + as, aphase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * BN);
      constexpr int LDW = CWID >= C::LDW_MAX ? C::LDW_MAX : CWID;
#pragma unroll
      for (int cc = cbeg; cc < cbeg + CWID; cc += LDW) {
        uint32_t v[LDW];
        // AFFINE fast path: this batch's LDW columns are LDW consecutive floats of the thread's x row -> fetch them (independent
        // 128-bit loads, issued before the TMEM wait), update in registers, store once: one memory round trip per batch
        const int gcol = n_tile * BN + cc;
        const bool xfast = (EPI == EPI_AFFINE) && a.e.pairs_adjacent && (gcol + LDW <= a.e.Cx) && LDW >= 16;
        float4 xv[LDW >= 4 ? LDW / 4 : 1];
        float4* xp = nullptr;
        if (EPI == EPI_AFFINE && xfast) {
          xp = reinterpret_cast<float4*>(a.e.X + row * a.e.Cx + gcol);
          if (row_ok) {
#pragma unroll
            for (int k = 0; k < LDW / 4; ++k) xv[k] = xp[k];
          } else {
#pragma unroll
            for (int k = 0; k < LDW / 4; ++k) xv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        // RES_SKIP: this thread's slice of the staged input tile (already landed: in_full was waited for), fetched with
        // explicit shared-space loads BEFORE the accumulator wait so their latency overlaps it
        uint4 inp[LDW >= 8 ? LDW / 8 : 1];
        if ((EPI == EPI_RES_SKIP || EPI == EPI_LINEAR) && have_in) {
          const uint32_t sbase = smem_u32(stg);
#pragma unroll
          for (int k = 0; k < LDW / 8; ++k) inp[k] = lds128(sbase + stg_off(r, cc + 8 * k));
        }
#pragma unroll
        for (int j = 0; j < LDW; j += 16) tmem_ld_x16(taddr + cc + j, v + j);
        tmem_ld_wait();
        if (EPI == EPI_AFFINE && xfast) {
#pragma unroll
          for (int j = 0; j < LDW; j += 16) affine16_regs(a.e, sbias, gcol + j, v + j, xv + j / 4, row_ok, ls_sum, row);
          if (row_ok) {
#pragma unroll
            for (int k = 0; k < LDW / 4; ++k) xp[k] = xv[k];
          }
        } else {
#pragma unroll
          for (int j = 0; j < LDW; j += 16)
            epilogue16<EPI, BN, WS>(a, row, row_ok, r, n_tile, cc + j, v + j, stg, have_in, inp + j / 8, sbias, ls_sum);
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {                // accumulator drained: the (leader's) MMA warp may start tile it+2 in this TMEM stage
        if (PAIR) mbar_arrive_remote(tmem_empty + as, 0);
        else mbar_arrive(tmem_empty + as);
      }
      if (C::STAGED) {
        // generic-proxy writes -> visible to the async proxy; then this warp stores its own 32 rows (rows >= Ti are clipped)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          int om = 0, ocol = n_tile * C::OUT_COLS;
          if (EPI == EPI_RES_SKIP || EPI == EPI_LINEAR) {
            om = (a.e.has_res && n_tile * BN < a.e.F) ? 0 : 1;
            ocol = (n_tile * BN) % a.e.F;
          }
          constexpr int SPW = C::SUBTILES / (C::WPG == 8 ? 2 : 1);   // 64-column sub-tiles per warp
#pragma unroll
          for (int jj = 0; jj < SPW; ++jj) {
            const int j = half * SPW + jj;
            tma_store_3d(&a.mapOut[om], stg + (size_t)j * BM * 128 + (size_t)lg * 32 * 128, ocol + j * 64, t0 + lg * 32, ub);
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (C::STAGED && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (EPI == EPI_AFFINE) {
      if (!a.e.reverse && a.e.logdet_acc) {
        ls_sum = warp_sum(ls_sum);
        if (lane == 0 && ls_sum != 0.0) atomicAdd(a.e.logdet_acc, ls_sum);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (PAIR || CL > 1) cluster_sync_all();  // no CTA exits (or frees TMEM) while its peer's MMAs / arrives / multicasts may still target it
  if (warp == 1) {
    tcgen05_fence_after();
    if (PAIR) tmem_dealloc_2sm<C::TMEM_COLS>(tmem_base);
    else tmem_dealloc<C::TMEM_COLS>(tmem_base);
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

// activations [B, Ti, C] bf16 (row stride ld elements) -> 3-D map, box (64, 128, 1), 128B swizzle, zero OOB fill
static bool g_map_fp16 = false;   // element type the next tensor maps are encoded with (set by tc_prepare; both are 2-byte types)
static CUtensorMapDataType map_dtype() { return g_map_fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; }
int make_act_map(CUtensorMap* map, const void* base, int B, int Ti, int C, int64_t ld) {
  EncodeFn enc = get_encode();
  FWN_CHECK(enc, "cuTensorMapEncodeTiled unavailable (driver too old?)");
  FWN_CHECK((ld * 2) % 16 == 0 && (reinterpret_cast<uintptr_t>(base) % 16) == 0, "TMA needs 16-byte aligned rows (ld=%lld)", (long long)ld);
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)Ti, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * (cuuint64_t)Ti};
  cuuint32_t box[3] = {BK, BM, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, map_dtype(), 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FWN_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activations B=%d Ti=%d C=%d) failed: %d", B, Ti, C, (int)r);
  return 0;
}
// same tensor, (64, 32, 1) boxes: the per-warp TMA stores of the epilogue (one TMEM lane group = 32 rows)
int make_store_map(CUtensorMap* map, const void* base, int B, int Ti, int C, int64_t ld) {
  EncodeFn enc = get_encode();
  FWN_CHECK(enc, "cuTensorMapEncodeTiled unavailable (driver too old?)");
  FWN_CHECK((ld * 2) % 16 == 0 && (reinterpret_cast<uintptr_t>(base) % 16) == 0, "TMA needs 16-byte aligned rows (ld=%lld)", (long long)ld);
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)Ti, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * (cuuint64_t)Ti};
  cuuint32_t box[3] = {BK, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, map_dtype(), 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FWN_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(store map B=%d Ti=%d C=%d) failed: %d", B, Ti, C, (int)r);
  return 0;
}
// weights [Npad, Kpad] bf16 -> 2-D map, box (64, bn)
int make_w_map(CUtensorMap* map, const void* base, int Npad, int Kpad, int bn) {
  EncodeFn enc = get_encode();
  FWN_CHECK(enc, "cuTensorMapEncodeTiled unavailable (driver too old?)");
  cuuint64_t dims[2] = {(cuuint64_t)Kpad, (cuuint64_t)Npad};
  cuuint64_t strides[1] = {(cuuint64_t)Kpad * 2};
  cuuint32_t box[2] = {BK, (cuuint32_t)bn};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, map_dtype(), 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FWN_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weights N=%d K=%d box=%d) failed: %d", Npad, Kpad, bn, (int)r);
  return 0;
}

// FWN_WS_MULTICAST=1 runs the weight-stationary GEMMs as clusters that receive their activation chunks by TMA multicast.  Measured
// on C3 (profiles/r2_ws_multicast.md): res|skip 12.7 -> 15.0 ms per pass, final conv 3.1 -> 3.2 ms -- the lock-step of the cluster
// (a stage is re-armed only after every CTA of the cluster released it) costs more than the L2->SM traffic it saves.  Off by default.
static bool ws_multicast() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FWN_WS_MULTICAST");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

template <int EPI, int BN, bool WS, bool PAIR = false, int CL = 1>
static int launch(const TcArgs& a, cudaStream_t st) {
  using C = Cfg<EPI, BN, WS, PAIR>;
  static bool configured = false;
  static int max_clusters = 0;   // CL > 1: clusters of CL CTAs that can be resident at once
  if (!configured) {
    FWN_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<EPI, BN, WS, PAIR, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    if (CL > 1) {
      cudaLaunchConfig_t q = {};
      q.gridDim = dim3((unsigned)(num_sms() / CL * CL));
      q.blockDim = dim3(C::THREADS);
      q.dynamicSmemBytes = C::SMEM;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = CL;
      qa[0].val.clusterDim.y = 1;
      qa[0].val.clusterDim.z = 1;
      q.attrs = qa;
      q.numAttrs = 1;
      if (cudaOccupancyMaxActiveClusters(&max_clusters, tc_gemm_kernel<EPI, BN, WS, PAIR, CL>, &q) != cudaSuccess || max_clusters <= 0) {
        cudaGetLastError();
        max_clusters = num_sms() / CL;
      }
    }
    configured = true;
  }
  const int num_m = a.B * a.tiles_per_utt;
  int grid;
  if (WS) {
    int nch = 0;
    for (int s = 0; s < a.nseg; ++s) nch += a.nchunk[s];
    FWN_CHECK(nch <= C::WS_CHUNKS, "weight-stationary GEMM needs K <= 256");
    FWN_CHECK(CL == 1 || CL == a.n_tiles, "internal: cluster size must equal the number of column tiles");
    const int cap = CL > 1 ? std::min(max_clusters, num_sms() / CL) : num_sms() / a.n_tiles;
    const int groups = std::max(1, std::min(num_m, cap));
    grid = groups * a.n_tiles;
  } else if (PAIR) {
    const int pairs = ((num_m + 1) / 2) * a.n_tiles;
    grid = 2 * std::min(pairs, num_sms() / 2);
  } else {
    grid = std::min(num_m * a.n_tiles, num_sms());
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(C::THREADS);
  cfg.dynamicSmemBytes = C::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (PAIR || CL > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = PAIR ? 2 : CL;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {   // programmatic dependent launch: see pdl_wait() in tc_ptx.cuh
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  FWN_CUDA(cudaLaunchKernelEx(&cfg, tc_gemm_kernel<EPI, BN, WS, PAIR, CL>, a));
  FWN_LAUNCH_CHECK();
  return 0;
}

// tile width per GEMM family: the two 1x1 K=256 GEMMs of the model path run weight-stationary at BN=128
int block_n_for(EpiKind kind, int N, bool model_path) {
  if (kind == EPI_GATE) return 256;
  if (kind == EPI_RES_SKIP) return 128;
  if (kind == EPI_PLAIN && model_path) return 128;
  int bn = 16;
  while (bn < N && bn < 256) bn *= 2;
  return bn;
}

int tc_launch(TcArgs& a, EpiKind kind, int bn, bool ws, cudaStream_t st) {
  FWN_CHECK(kind == EPI_PLAIN_F32 || a.n_tiles * bn <= 512, "tc_gemm: N=%d too wide", a.N);
  switch (kind) {
    case EPI_LINEAR:
      FWN_CHECK(bn == 128, "linear runs at BN=128");
      if (ws && a.n_tiles == 2 && ws_multicast()) return launch<EPI_LINEAR, 128, true, false, 2>(a, st);
      return ws ? launch<EPI_LINEAR, 128, true>(a, st) : launch<EPI_LINEAR, 128, false>(a, st);
    case EPI_PLAIN_F32:
      switch (bn) {
        case 16: return launch<EPI_PLAIN_F32, 16, false>(a, st);
        case 32: return launch<EPI_PLAIN_F32, 32, false>(a, st);
        case 64: return launch<EPI_PLAIN_F32, 64, false>(a, st);
        case 128: return launch<EPI_PLAIN_F32, 128, false>(a, st);
        default: return launch<EPI_PLAIN_F32, 256, false>(a, st);
      }
    case EPI_GATE_BWD: break;
    case EPI_GATE:
      FWN_CHECK(bn == 256, "gate needs BN=256");
      return a.multicast ? launch<EPI_GATE, 256, false, true>(a, st) : launch<EPI_GATE, 256, false, false>(a, st);
    case EPI_RES_SKIP:
      FWN_CHECK(bn == 128 && ws, "res/skip runs weight-stationary at BN=128");
      // the column tiles of one row tile share their A operand: clusters of n_tiles CTAs receive it by TMA multicast
      if (a.n_tiles == 4 && ws_multicast()) return launch<EPI_RES_SKIP, 128, true, false, 4>(a, st);
      if (a.n_tiles == 2 && ws_multicast()) return launch<EPI_RES_SKIP, 128, true, false, 2>(a, st);
      return launch<EPI_RES_SKIP, 128, true>(a, st);
    case EPI_PLAIN:
      if (ws) {
        FWN_CHECK(bn == 128, "weight-stationary plain GEMM needs BN=128");
        if (a.n_tiles == 2 && ws_multicast()) return launch<EPI_PLAIN, 128, true, false, 2>(a, st);
        return launch<EPI_PLAIN, 128, true>(a, st);
      }
      switch (bn) {
        case 16: return launch<EPI_PLAIN, 16, false>(a, st);
        case 32: return launch<EPI_PLAIN, 32, false>(a, st);
        case 64: return launch<EPI_PLAIN, 64, false>(a, st);
        case 128: return launch<EPI_PLAIN, 128, false>(a, st);
        default: return launch<EPI_PLAIN, 256, false>(a, st);
      }
    case EPI_AFFINE:
      switch (bn) {
        case 16: return launch<EPI_AFFINE, 16, false>(a, st);
        case 32: return launch<EPI_AFFINE, 32, false>(a, st);
        case 64: return launch<EPI_AFFINE, 64, false>(a, st);
        case 128: return launch<EPI_AFFINE, 128, false>(a, st);
        default: return launch<EPI_AFFINE, 256, false>(a, st);
      }
  }
  return 1;
}

}  // namespace tc

// ---------------------------------------------------------------- engine glue
// FWN_GATE_PAIR=0 runs the gate GEMM as single-CTA MMAs instead of cta_group::2 pairs (diagnostics / A-B timing)
static bool gate_multicast() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FWN_GATE_PAIR");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

struct TcPlan {
  Workspace w;
  // activation maps per block: h0, h1, o, s, u (C = F), cA, cB (C = Kc) and a0 (C = ceil8(nq))
  std::vector<CUtensorMap> act;  // [n_block * 8]
  std::vector<CUtensorMap> st32; // [n_block * 5] store maps (32-row boxes) of h0, h1, o, s, u
  std::vector<CUtensorMap> wmap; // [n_flows * GEMM_IDS]
  std::vector<int> wbn;          // BN each weight map was built for
  std::vector<char> wws;         // runs weight-stationary
};

static int act_index(const Workspace& w, const void* p) {
  if (p == w.h0) return 0;
  if (p == w.h1) return 1;
  if (p == w.o) return 2;
  if (p == w.s) return 3;
  if (p == w.u) return 4;
  if (p == w.cA) return 5;
  if (p == w.cB) return 6;
  if (p == w.a0) return 7;
  return -1;
}

int tc_prepare(Model* m, const Workspace& w, int B, int T, cudaStream_t st) {
  if (m->tc && m->plan_B == B && m->plan_T == T && m->plan_ws == (const void*)w.sums) return 0;
  const fwn_config& c = m->cfg;
  const int F = c.filter_size, L = c.n_layer, H = c.num_mels / 2;
  if (!m->tc) m->tc = new TcPlan();
  TcPlan* p = m->tc;
  tc::g_map_fp16 = c.precision == FWN_MIXED_FP16;
  p->w = w;
  p->act.assign((size_t)c.n_block * 8, CUtensorMap());
  p->st32.assign((size_t)c.n_block * 5, CUtensorMap());
  p->wmap.assign(m->flows.size() * GEMM_IDS, CUtensorMap());
  p->wbn.assign(m->flows.size() * GEMM_IDS, 0);
  p->wws.assign(m->flows.size() * GEMM_IDS, 0);
  for (int i = 0; i < c.n_block; ++i) {
    const int Ti = T >> (i + 1), Kc = H << (i + 1);
    void* bufs[8] = {w.h0, w.h1, w.o, w.s, w.u, w.cA, w.cB, w.a0};
    const int kq = ((1 << i) + 7) / 8 * 8;  // nq = 2^i pass-through channels, row pitch padded to 8
    for (int k = 0; k < 8; ++k) {
      const int C = k < 5 ? F : (k < 7 ? Kc : kq);
      if (tc::make_act_map(&p->act[(size_t)i * 8 + k], bufs[k], B, Ti, C, C)) return 1;
      if (k < 5 && tc::make_store_map(&p->st32[(size_t)i * 5 + k], bufs[k], B, Ti, C, C)) return 1;
    }
  }
  for (size_t f = 0; f < m->flows.size(); ++f) {
    const FlowPack& fp = m->flows[f];
    auto mk = [&](int id, const void* wptr, int N, int Kpad, EpiKind kind, int chunks = 4) {
      // the 1x1 / front GEMMs run weight-stationary at BN=128 when their K fits 4 resident chunks, else streaming at BN=256
      const bool ws = (kind == EPI_RES_SKIP || kind == EPI_PLAIN) && chunks <= 4;
      const int bn = (kind == EPI_PLAIN && !ws) ? 256 : tc::block_n_for(kind, N, true);
      p->wws[f * GEMM_IDS + id] = ws ? 1 : 0;
      const int Npad = (N + 15) / 16 * 16;
      p->wbn[f * GEMM_IDS + id] = bn;
      const int box = (kind == EPI_GATE && gate_multicast()) ? bn / 2 : std::min(bn, Npad);
      return tc::make_w_map(&p->wmap[f * GEMM_IDS + id], wptr, Npad, Kpad, box);
    };
    for (int n = 0; n < L; ++n) {
      if (mk(GEMM_GATE0 + n, fp.gate_w[n], 2 * F, fp.gate_ld, EPI_GATE)) return 1;
      if (mk(GEMM_RS0 + n, fp.rs_w[n], n == L - 1 ? F : 2 * F, fp.rs_ld[n], EPI_RES_SKIP)) return 1;
    }
    if (mk(GEMM_FINAL, fp.final_w, F, fp.final_ld, EPI_PLAIN)) return 1;
    if (mk(GEMM_FRONT, fp.front_wtc, F, fp.front_ld, EPI_PLAIN, 3 * ((fp.front_k16 + 63) / 64))) return 1;
    if (mk(GEMM_ZERO, fp.zero_w, 2 * fp.nq, fp.zero_ld, EPI_AFFINE)) return 1;
  }
  m->plan_B = B;
  m->plan_T = T;
  m->plan_ws = (const void*)w.sums;
  return 0;
}

int tc_run(Model* m, const GemmArgs& g, EpiKind kind, int gemm_id, const FlowPack& fp, cudaStream_t st) {
  TcPlan* p = m->tc;
  FWN_CHECK(p, "tcgen05 engine not prepared");
  const size_t f = (size_t)(&fp - m->flows.data());
  const int block = (int)(f / m->cfg.n_flow);
  tc::TcArgs a;
  memset(&a, 0, sizeof(a));
  a.nseg = g.nseg;
  for (int s = 0; s < g.nseg; ++s) {
    const int ai = act_index(p->w, g.seg[s].A);
    FWN_CHECK(ai >= 0, "tc_run: segment %d does not read a planned workspace buffer", s);
    a.mapA[s] = p->act[(size_t)block * 8 + ai];
    a.shift[s] = g.seg[s].shift;
    const int K16 = (g.seg[s].K + 15) / 16 * 16;
    a.nchunk[s] = (K16 + tc::BK - 1) / tc::BK;
    a.last_ksteps[s] = (K16 - (a.nchunk[s] - 1) * tc::BK) / tc::UMMA_K;
    a.wk0[s] = g.seg[s].koff;
  }
  a.mapB = p->wmap[f * GEMM_IDS + gemm_id];
  auto act_map = [&](const void* ptr, CUtensorMap* dst) -> bool {
    const int ai = ptr ? act_index(p->w, ptr) : -1;
    if (ai < 0) return false;
    *dst = p->act[(size_t)block * 8 + ai];
    return true;
  };
  auto store_map = [&](const void* ptr, CUtensorMap* dst) -> bool {
    const int ai = ptr ? act_index(p->w, ptr) : -1;
    if (ai < 0 || ai >= 5) return false;
    *dst = p->st32[(size_t)block * 5 + ai];
    return true;
  };
  if (kind == EPI_GATE) {
    FWN_CHECK(store_map(g.e.out0, &a.mapOut[0]), "tc_run: gate output is not a planned buffer");
  } else if (kind == EPI_RES_SKIP) {
    if (g.e.has_res) {
      FWN_CHECK(store_map(g.e.out0, &a.mapOut[0]) && act_map(g.e.in0, &a.mapIn[0]), "tc_run: residual buffers are not planned buffers");
      a.has_in[0] = 1;
    }
    FWN_CHECK(store_map(g.e.out1, &a.mapOut[1]), "tc_run: skip output is not a planned buffer");
    a.has_in[1] = act_map(g.e.in1, &a.mapIn[1]) ? 1 : 0;
  } else if (kind == EPI_PLAIN) {
    FWN_CHECK(store_map(g.e.out0, &a.mapOut[0]), "tc_run: output is not a planned buffer");
  }
  const int bn = p->wbn[f * GEMM_IDS + gemm_id];
  a.B = g.B;
  a.Ti = g.Ti;
  a.tiles_per_utt = (g.Ti + tc::BM - 1) / tc::BM;
  a.N = g.N;
  a.n_tiles = (g.N + bn - 1) / bn;
  a.e = g.e;
  const bool ws = p->wws[f * GEMM_IDS + gemm_id] != 0;
  a.multicast = (kind == EPI_GATE && gate_multicast()) ? 1 : 0;
  a.fp16 = m->cfg.precision == FWN_MIXED_FP16 ? 1 : 0;
  return tc::tc_launch(a, kind, bn, ws, st);
}

int tc_map_3d(CUtensorMap* out, const void* base, int C, int Ti, int B, int64_t ld, int box_c, int box_rows, bool fp16);   // below (cached)

// Fused ResBlock layer (layer_tc.cu): the gate GEMM `g` and the res|skip 1x1 `r` of one layer in one launch; o never leaves the SM.
// Supported: the CTA-pair gate path, F = 256, and a 1x1 whose staged epilogue input is single (first layer: h_in; last layer: the
// running skip sum) -- a middle layer of a deeper WaveNet (residual AND running skip) takes the two-launch path.
bool tc_layer_supported(const Model* m, const GemmArgs& g, const GemmArgs& r) {
  if (!m->tc || !gate_multicast() || m->cfg.filter_size != 256) return false;
  if (g.e.tape || g.e.out1) return false;
  if (r.e.has_res && r.e.in1) return false;
  return true;
}
// f, z (optional, last layer only): the final 1x1 and the zero conv + affine coupling are folded into the same launch (the tail ops of
// layer_tc.cu); the caller has checked tc_tail_supported and that the layer's staged input is the running skip sum.
int tc_run_layer(Model* m, const GemmArgs& g, const GemmArgs& r, int layer, const FlowPack& fp, cudaStream_t st, const GemmArgs* fin, const GemmArgs* zero) {
  TcPlan* p = m->tc;
  FWN_CHECK(p, "tcgen05 engine not prepared");
  const size_t fi_ = (size_t)(&fp - m->flows.data());
  const int block = (int)(fi_ / m->cfg.n_flow);
  tc::LayerArgs a;
  memset(&a, 0, sizeof(a));
  a.nseg = g.nseg;
  for (int s = 0; s < g.nseg; ++s) {
    const int ai = act_index(p->w, g.seg[s].A);
    FWN_CHECK(ai >= 0, "tc_run_layer: segment %d does not read a planned workspace buffer", s);
    a.mapA[s] = p->act[(size_t)block * 8 + ai];
    a.shift[s] = g.seg[s].shift;
    const int K16 = (g.seg[s].K + 15) / 16 * 16;
    a.nchunk[s] = (K16 + tc::BK - 1) / tc::BK;
    a.last_ksteps[s] = (K16 - (a.nchunk[s] - 1) * tc::BK) / tc::UMMA_K;
    a.wk0[s] = g.seg[s].koff;
  }
  a.mapWg = p->wmap[fi_ * GEMM_IDS + GEMM_GATE0 + layer];
  a.mapWr = p->wmap[fi_ * GEMM_IDS + GEMM_RS0 + layer];
  auto act_map = [&](const void* ptr, CUtensorMap* dst) -> bool {
    const int ai = ptr ? act_index(p->w, ptr) : -1;
    if (ai < 0) return false;
    *dst = p->act[(size_t)block * 8 + ai];
    return true;
  };
  auto store_map = [&](const void* ptr, CUtensorMap* dst) -> bool {
    const int ai = ptr ? act_index(p->w, ptr) : -1;
    if (ai < 0 || ai >= 5) return false;
    *dst = p->st32[(size_t)block * 5 + ai];
    return true;
  };
  a.has_res = r.e.has_res ? 1 : 0;
  if (a.has_res) {
    FWN_CHECK(store_map(r.e.out0, &a.mapOutH) && act_map(r.e.in0, &a.mapIn), "tc_run_layer: residual buffers are not planned buffers");
    a.has_in = 1;
  } else if (r.e.in1) {
    FWN_CHECK(act_map(r.e.in1, &a.mapIn), "tc_run_layer: skip input is not a planned buffer");
    a.has_in = 1;
  }
  FWN_CHECK(store_map(r.e.out1, &a.mapOutS), "tc_run_layer: skip output is not a planned buffer");
  a.relu = r.e.relu;
  a.fp16 = m->cfg.precision == FWN_MIXED_FP16 ? 1 : 0;
  a.B = g.B;
  a.Ti = g.Ti;
  a.tiles_per_utt = (g.Ti + tc::BM - 1) / tc::BM;
  a.gate_bias = g.e.bias;
  a.rs_bias = r.e.bias;
  a.pc = reinterpret_cast<const float*>(g.e.in0);
  a.pc_ld = g.e.ld;
  if (fin && zero) {
    FWN_CHECK(!a.has_res && a.has_in, "tc_run_layer: the tail can only follow a last layer with a running skip sum");
    a.tail = 1;
    a.mapWf = p->wmap[fi_ * GEMM_IDS + GEMM_FINAL];
    FWN_CHECK(p->wbn[fi_ * GEMM_IDS + GEMM_FINAL] == 128, "tc_run_layer: final-conv weight map has the wrong box");
    a.Nz = zero->N;
    a.NzBox = std::min(p->wbn[fi_ * GEMM_IDS + GEMM_ZERO], (zero->N + 15) / 16 * 16);
    FWN_CHECK(a.NzBox == 16 || a.NzBox == 32, "tc_run_layer: zero conv too wide (%d)", a.NzBox);
    const int Npad = (zero->N + 15) / 16 * 16;
    if (tc_map_3d(&a.mapWz, fp.zero_w, fp.zero_ld, Npad, 0, fp.zero_ld, tc::BK, a.NzBox / 2, a.fp16 != 0)) return 1;
    a.final_bias = fin->e.bias;
    a.ez = zero->e;
  }
  return tc::launch_layer(a, st);
}

// Fused WaveNet tail (tail_tc.cu): final 1x1 `f` (+ReLU) and the zero conv + affine coupling `z` in one launch; u never leaves the SM.
// Supported: F = 256 and at most 32 zero-conv columns (blocks with C_x <= 32); deeper blocks take the two-launch path.
bool tc_tail_supported(const Model* m, const GemmArgs& f, const GemmArgs& z) {
  if (!m->tc || m->cfg.filter_size != 256) return false;
  if (z.e.out1 || z.N > 32 || !f.e.relu) return false;
  return true;
}
int tc_run_tail(Model* m, const GemmArgs& f, const GemmArgs& z, const FlowPack& fp, cudaStream_t st) {
  TcPlan* p = m->tc;
  FWN_CHECK(p, "tcgen05 engine not prepared");
  const size_t fi = (size_t)(&fp - m->flows.data());
  const int block = (int)(fi / m->cfg.n_flow);
  tc::TailArgs a;
  memset(&a, 0, sizeof(a));
  const int ai = act_index(p->w, f.seg[0].A);
  FWN_CHECK(ai >= 0, "tc_run_tail: the final conv does not read a planned workspace buffer");
  a.mapS = p->act[(size_t)block * 8 + ai];
  a.mapWf = p->wmap[fi * GEMM_IDS + GEMM_FINAL];
  a.mapWz = p->wmap[fi * GEMM_IDS + GEMM_ZERO];
  FWN_CHECK(p->wbn[fi * GEMM_IDS + GEMM_FINAL] == 128, "tc_run_tail: final-conv weight map has the wrong box");
  a.B = f.B;
  a.Ti = f.Ti;
  a.tiles_per_utt = (f.Ti + tc::BM - 1) / tc::BM;
  a.Nz = z.N;
  a.NzBox = std::min(p->wbn[fi * GEMM_IDS + GEMM_ZERO], (z.N + 15) / 16 * 16);
  FWN_CHECK(a.NzBox == 16 || a.NzBox == 32, "tc_run_tail: zero conv too wide (%d)", a.NzBox);
  a.fp16 = m->cfg.precision == FWN_MIXED_FP16 ? 1 : 0;
  a.final_bias = f.e.bias;
  a.e = z.e;
  return tc::launch_tail(a, st);
}

// Stand-alone mixed-precision conv (per-op entry fwn_conv1d_bf16): y = [relu](conv(x, w) + bias), bf16 in/out.
// x [B,T,Cin] bf16, w prepacked [ceil16(Cout)][ceil64(ksize*ceil16(Cin))] bf16 (tap-major K), bias fp32 [Cout].
int tc_conv1d(const void* x, const void* w, const float* bias, void* y, int B, int T, int Cin, int Cout, int ksize, int dilation, int causal,
              int relu, cudaStream_t st) {
  FWN_CHECK(ksize >= 1 && ksize <= 4 && Cin % 8 == 0 && Cout % 8 == 0, "conv1d_bf16: need ksize<=4 and channel counts that are multiples of 8");
  tc::TcArgs a;
  memset(&a, 0, sizeof(a));
  tc::g_map_fp16 = false;
  const int Cin16 = (Cin + 15) / 16 * 16, Kpad = (ksize * Cin16 + 63) / 64 * 64, Npad = (Cout + 15) / 16 * 16;
  const int bn = tc::block_n_for(EPI_PLAIN, Cout, false);
  const int pad = causal ? dilation * (ksize - 1) : dilation * (ksize - 1) / 2;
  a.nseg = ksize;
  for (int k = 0; k < ksize; ++k) {
    if (tc::make_act_map(&a.mapA[k], x, B, T, Cin, Cin)) return 1;
    a.shift[k] = k * dilation - pad;
    a.nchunk[k] = (Cin16 + tc::BK - 1) / tc::BK;
    a.last_ksteps[k] = (Cin16 - (a.nchunk[k] - 1) * tc::BK) / tc::UMMA_K;
    a.wk0[k] = k * Cin16;
  }
  if (tc::make_w_map(&a.mapB, w, Npad, Kpad, std::min(bn, Npad))) return 1;
  a.B = B; a.Ti = T; a.tiles_per_utt = (T + tc::BM - 1) / tc::BM; a.N = Cout; a.n_tiles = (Cout + bn - 1) / bn;
  a.e.bias = bias; a.e.out0 = y; a.e.ld = Cout; a.e.relu = relu; a.e.F = Cout;
  if (bn == 128 && tc::make_store_map(&a.mapOut[0], y, B, T, Cout, Cout)) return 1;  // staged TMA-store epilogue
  return tc::tc_launch(a, EPI_PLAIN, bn, false, st);
}

// ---------------------------------------------------------------- generic 16-bit GEMM (training step): cached tensor maps
namespace {
struct MapKey {
  const void* base;
  int64_t ld;
  int C, Ti, B, box_c, box_rows, fp16;
  bool operator==(const MapKey& o) const {
    return base == o.base && ld == o.ld && C == o.C && Ti == o.Ti && B == o.B && box_c == o.box_c && box_rows == o.box_rows && fp16 == o.fp16;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = reinterpret_cast<uintptr_t>(k.base) * 0x9E3779B97F4A7C15ull;
    auto mix = [&](uint64_t v) { h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); };
    mix((uint64_t)k.ld); mix((uint64_t)k.C); mix((uint64_t)k.Ti); mix((uint64_t)k.B);
    mix((uint64_t)k.box_c * 4096 + (uint64_t)k.box_rows * 2 + (uint64_t)k.fp16);
    return (size_t)h;
  }
};
std::unordered_map<MapKey, CUtensorMap, MapKeyHash>& map_cache() {
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> c;
  return c;
}
}  // namespace

// 16-bit tensor [B, Ti, C] with row pitch ld (elements) -> 3-D map with a (box_c, box_rows, 1) box, 128B swizzle, zero OOB fill.
// B == 0 encodes a 2-D weight matrix [Ti = rows(N), C = K] with a (box_c, box_rows) box.  Encoded once per distinct key.
int tc_map_3d(CUtensorMap* out, const void* base, int C, int Ti, int B, int64_t ld, int box_c, int box_rows, bool fp16) {
  const MapKey key{base, ld, C, Ti, B, box_c, box_rows, fp16 ? 1 : 0};
  auto& cache = map_cache();
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return 0;
  }
  if (cache.size() > 200000) cache.clear();
  tc::EncodeFn enc = tc::get_encode();
  FWN_CHECK(enc, "cuTensorMapEncodeTiled unavailable (driver too old?)");
  FWN_CHECK((ld * 2) % 16 == 0 && (reinterpret_cast<uintptr_t>(base) % 16) == 0, "TMA needs 16-byte aligned rows (ld=%lld)", (long long)ld);
  const CUtensorMapDataType dt = fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r;
  if (B > 0) {
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)Ti, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * (cuuint64_t)Ti};
    cuuint32_t box[3] = {(cuuint32_t)box_c, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    r = enc(out, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)Ti};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_c, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    r = enc(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  FWN_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(B=%d Ti=%d C=%d ld=%lld box=%dx%d) failed: %d", B, Ti, C, (long long)ld, box_c, box_rows, (int)r);
  cache.emplace(key, *out);
  return 0;
}

int tc_gemm16(const GemmArgs& g0, EpiKind kind, const void* W, int Kpad, int Npad, bool fp16, cudaStream_t st) {
  if (g0.B <= 0 || g0.Ti <= 0 || g0.N <= 0) return 0;
  GemmArgs g = g0;
  bool shifted = false;
  for (int s = 0; s < g.nseg; ++s) shifted = shifted || g.seg[s].shift != 0;
  if (!shifted && (int64_t)g.B * g.Ti < (int64_t(1) << 31)) {   // 1x1 convs: utterance borders do not matter -> one flat row axis, full tiles
    g.Ti = g.B * g.Ti;
    g.B = 1;
  }
  const int F = g.e.F;
  int total_chunks = 0;
  tc::TcArgs a;
  memset(&a, 0, sizeof(a));
  a.nseg = g.nseg;
  for (int s = 0; s < g.nseg; ++s) {
    const Seg& sg = g.seg[s];
    if (tc_map_3d(&a.mapA[s], sg.A, sg.K, g.Ti, g.B, sg.lda, tc::BK, tc::BM, fp16)) return 1;
    a.shift[s] = sg.shift;
    const int K16 = (sg.K + 15) / 16 * 16;
    a.nchunk[s] = (K16 + tc::BK - 1) / tc::BK;
    a.last_ksteps[s] = (K16 - (a.nchunk[s] - 1) * tc::BK) / tc::UMMA_K;
    a.wk0[s] = sg.koff;
    total_chunks += a.nchunk[s];
  }
  int bn;
  bool ws = false, pair = false;
  switch (kind) {
    case EPI_GATE: bn = 256; pair = gate_multicast(); break;
    case EPI_RES_SKIP: bn = 128; ws = true; FWN_CHECK(total_chunks <= 4, "res|skip GEMM needs K <= 256"); break;
    case EPI_LINEAR: bn = 128; ws = total_chunks <= 4; break;
    case EPI_PLAIN: bn = 128; ws = total_chunks <= 4; FWN_CHECK(g.N % 128 == 0, "16-bit PLAIN GEMM needs N %% 128 == 0 (N=%d)", g.N); break;
    case EPI_PLAIN_F32:
    case EPI_AFFINE: bn = tc::block_n_for(EPI_AFFINE, g.N, false); break;
    default: FWN_CHECK(false, "tc_gemm16: unsupported epilogue %d", (int)kind);
  }
  // weights [Npad][Kpad]: 2-D map, (64, box_n) box.  The box is always the full column tile: the kernel expects BN x 128 bytes per
  // chunk on its barrier, and rows past Npad are zero-filled by TMA (they still count towards the transaction bytes)
  const int box_n = pair ? bn / 2 : bn;
  if (tc_map_3d(&a.mapB, W, Kpad, Npad, 0, Kpad, tc::BK, box_n, fp16)) return 1;
  auto act = [&](const void* p, CUtensorMap* dst, int C, int rows) { return p ? tc_map_3d(dst, p, C, g.Ti, g.B, C, tc::BK, rows, fp16) : 0; };
  if (kind == EPI_GATE) {
    FWN_CHECK(!(fp16 && g.e.tape), "the gate tape is a bf16 training feature");
    if (act(g.e.out0, &a.mapOut[0], F, 32)) return 1;
  } else if (kind == EPI_RES_SKIP) {
    if (g.e.has_res) {
      if (act(g.e.out0, &a.mapOut[0], F, 32) || act(g.e.in0, &a.mapIn[0], F, tc::BM)) return 1;
      a.has_in[0] = 1;
    }
    if (act(g.e.out1, &a.mapOut[1], F, 32)) return 1;
    if (g.e.in1) {
      if (act(g.e.in1, &a.mapIn[1], F, tc::BM)) return 1;
      a.has_in[1] = 1;
    }
  } else if (kind == EPI_LINEAR) {
    FWN_CHECK(g.N % 128 == 0, "16-bit LINEAR GEMM needs N %% 128 == 0 (N=%d)", g.N);
    g.e.has_res = 1;      // every column tile takes the "residual" route of the staged epilogue: map 0 in, map 0 out
    g.e.F = g.N;
    if (act(g.e.out0, &a.mapOut[0], g.N, 32)) return 1;
    if (g.e.in0) {
      if (act(g.e.in0, &a.mapIn[0], g.N, tc::BM)) return 1;
      a.has_in[0] = 1;
    }
  } else if (kind == EPI_PLAIN) {
    if (act(g.e.out0, &a.mapOut[0], g.N, 32)) return 1;
  }
  a.B = g.B;
  a.Ti = g.Ti;
  a.tiles_per_utt = (g.Ti + tc::BM - 1) / tc::BM;
  a.N = g.N;
  a.n_tiles = (g.N + bn - 1) / bn;
  a.e = g.e;
  a.multicast = pair ? 1 : 0;
  a.fp16 = fp16 ? 1 : 0;
  if (getenv("FWN_TRACE")) {
    fprintf(stderr, "tc_gemm16 kind=%d B=%d Ti=%d N=%d bn=%d ws=%d pair=%d Kpad=%d Npad=%d nseg=%d:", (int)kind, g.B, g.Ti, g.N, bn, (int)ws, (int)pair,
            Kpad, Npad, g.nseg);
    for (int s = 0; s < g.nseg; ++s)
      fprintf(stderr, " [K=%d lda=%lld sh=%d koff=%d nch=%d]", g.seg[s].K, (long long)g.seg[s].lda, g.seg[s].shift, g.seg[s].koff, a.nchunk[s]);
    fprintf(stderr, "\n");
    fflush(stderr);
  }
  return tc::tc_launch(a, kind, bn, ws, st);
}

void tc_free(Model* m) {
  delete m->tc;
  m->tc = nullptr;
}

}  // namespace fwn

// The tail of a coupling WaveNet in ONE launch (mixed-precision inference passes):
//
//   u          = relu(Conv_final(s))                       1x1, K = N = F = 256                 (modules.py:176-179)
//   (log_s, t) = ZeroConv1d(u)                             1x1, K = 256, N = 2 nq               (modules.py:39-59, 180)
//   x          <- ActNorm + affine coupling in place        + sum(log_s) for the log-det         (model.py:7-105, 121-161)
//
// The two-launch path (tc_gemm_kernel<PLAIN,WS> + tc_gemm_kernel<AFFINE>) writes u [rows, 256] to HBM and reads it straight back --
// two thirds of the pair's HBM traffic.  Here the first GEMM's epilogue writes relu(u) in place over its own A tile (s) in shared
// memory, in the K-major 128B-swizzled layout, and the zero conv takes it from there.
//
// One CTA per SM, persistent over 128-row tiles.  Warp roles: warp 0 streams the activation tiles (two 64 KB tile buffers: tile j+1
// is in flight while tile j is computed), warp 1 streams the final-conv weights (L2-resident, two 32 KB slots) after loading the
// zero-conv weights once, warp 2 issues the MMAs  F(0) | F(1) Z(0) | F(2) Z(1) | ...  (the zero conv of tile j is issued after the
// final conv of tile j+1, so the pipe never waits for the ReLU epilogue), warps 3..18 are the epilogue: all 16 drain F (one row x
// 64 columns per thread, the TMEM columns handed back before the arithmetic), four of them drain Z and update x.
// TMEM: columns [0,256) F accumulator, [256, 256 + Nz) Z accumulator.
#include <cuda.h>
#include <stdio.h>

#include "common.cuh"
#include "tail_tc.cuh"
#include "tc_ptx.cuh"

namespace fwn {
namespace tc {

constexpr int T_TILE_BYTES = 4 * A_BYTES;         // [128 rows x 256 ch] 16-bit
constexpr int T_W_BYTES = 2 * A_BYTES;            // one K chunk of the final-conv weights: [256 n x 64 k]
constexpr int T_WZ_BYTES = 4 * 32 * 128;          // zero-conv weights, up to 32 columns: four K chunks of [32 n x 64 k]
constexpr int T_THREADS = 32 * 19;
constexpr size_t T_SMEM = 2 * T_TILE_BYTES + 2 * T_W_BYTES + T_WZ_BYTES + 1024 + 256 + 512;
static_assert(T_SMEM <= 227 * 1024, "shared memory budget exceeded");

__device__ __forceinline__ uint32_t t_tile_off(int r, int c) {
  return (uint32_t)((c >> 6) * A_BYTES + r * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4));
}

__global__ void __launch_bounds__(T_THREADS, 1) tail_kernel(const __grid_constant__ TailArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* tile_base = smem;                                   // two activation / u tiles
  uint8_t* w_base = smem + 2 * T_TILE_BYTES;                   // two final-weight slots
  uint8_t* wz_base = w_base + 2 * T_W_BYTES;                   // resident zero-conv weights
  float* sbias_f = reinterpret_cast<float*>(wz_base + T_WZ_BYTES);   // [256]
  float* sbias_z = sbias_f + 256;                              // [64] (32 used)
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sbias_z + 64);      // [2][4] activation chunk landed
  uint64_t* tile_free = a_full + 8;                            // [2] Z of the tile in this buffer completed
  uint64_t* w_full = tile_free + 2;                            // [2]
  uint64_t* w_empty = w_full + 2;                              // [2]
  uint64_t* wz_full = w_empty + 2;
  uint64_t* f_full = wz_full + 1;                              // F accumulator complete
  uint64_t* f_empty = f_full + 1;                              // 16 warps hold it in registers
  uint64_t* u_full = f_empty + 1;                              // [2] relu(u) tile written (16 warps)
  uint64_t* z_full = u_full + 2;
  uint64_t* z_empty = z_full + 1;                              // 4 warps
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(z_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = a.B * a.tiles_per_utt;
  const int n_my = ((int)blockIdx.x < n_tiles) ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  auto coords = [&](int j, int& ub, int& t0) {
    const int m_tile = (int)blockIdx.x + j * (int)gridDim.x;
    ub = m_tile / a.tiles_per_utt;
    t0 = (m_tile - ub * a.tiles_per_utt) * BM;
  };

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) {
      printf("tail_kernel: dynamic shared memory is not 1024-byte aligned\n");
      asm volatile("trap;");
    }
    prefetch_tmap(&a.mapS);
    prefetch_tmap(&a.mapWf);
    prefetch_tmap(&a.mapWz);
    for (int i = 0; i < 8; ++i) mbar_init(a_full + i, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(tile_free + i, 1);
      mbar_init(w_full + i, 1);
      mbar_init(w_empty + i, 1);
      mbar_init(u_full + i, 16);
    }
    mbar_init(wz_full, 1);
    mbar_init(f_full, 1);
    mbar_init(f_empty, 16);
    mbar_init(z_full, 1);
    mbar_init(z_empty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 256; i += T_THREADS) sbias_f[i] = __ldg(a.final_bias + i);
  for (int i = threadIdx.x; i < 64; i += T_THREADS) sbias_z[i] = (i < a.Nz && a.e.bias) ? __ldg(a.e.bias + i) : 0.f;
  if (warp == 2) tmem_alloc<512>(tmem_ptr);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================== activation tiles =====================
    if (lane == 0) {
      pdl_wait();
      for (int j = 0; j < n_my; ++j) {
        const int b = j & 1;
        int ub, t0;
        coords(j, ub, t0);
        if (j >= 2) mbar_wait(tile_free + b, (uint32_t)((j >> 1) - 1) & 1);   // Z(j-2) has read the u tile in this buffer
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          mbar_expect_tx(a_full + b * 4 + c, A_BYTES);
          tma_load_3d(tile_base + (size_t)b * T_TILE_BYTES + (size_t)c * A_BYTES, &a.mapS, a_full + b * 4 + c, c * BK, t0, ub);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== weights =====================
    if (lane == 0) {
      mbar_expect_tx(wz_full, 4u * (uint32_t)a.NzBox * 128u);
      for (int c = 0; c < 4; ++c) tma_load_2d(wz_base + (size_t)c * a.NzBox * 128, &a.mapWz, wz_full, c * BK, 0);
      int slot = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_my; ++j)
        for (int c = 0; c < 4; ++c) {
          mbar_wait(w_empty + slot, phase ^ 1);
          mbar_expect_tx(w_full + slot, T_W_BYTES);
          tma_load_2d(w_base + (size_t)slot * T_W_BYTES, &a.mapWf, w_full + slot, c * BK, 0);
          tma_load_2d(w_base + (size_t)slot * T_W_BYTES + A_BYTES, &a.mapWf, w_full + slot, c * BK, 128);
          if (++slot == 2) { slot = 0; phase ^= 1; }
        }
    }
  } else if (warp == 2) {
    // ===================== MMA issuer:  F(0) | F(1) Z(0) | F(2) Z(1) | ... | Z(n-1) =====================
    const uint32_t bf = a.fp16 ? 0u : IDESC_BF16_BITS;
    const uint32_t idesc_f = ((1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24)) | bf;
    const uint32_t idesc_z = ((1u << 4) | ((uint32_t)(a.NzBox >> 3) << 17) | ((uint32_t)(BM >> 4) << 24)) | bf;
    const uint64_t desc_hi = make_smem_desc(0);
    const uint32_t tile0 = smem_u32(tile_base), w0 = smem_u32(w_base), wz0 = smem_u32(wz_base);
    int slot = 0;
    uint32_t phase = 0;
    mbar_wait(wz_full, 0);
    auto issue_f = [&](int j) {
      const int b = j & 1;
      mbar_wait(f_empty, ((uint32_t)j & 1) ^ 1);
      tcgen05_fence_after();
      for (int c = 0; c < 4; ++c) {
        mbar_wait(a_full + b * 4 + c, (uint32_t)(j >> 1) & 1);
        mbar_wait(w_full + slot, phase);
        tcgen05_fence_after();
        const uint32_t sa = tile0 + (uint32_t)b * T_TILE_BYTES + (uint32_t)c * A_BYTES, sb = w0 + (uint32_t)slot * T_W_BYTES;
        umma_chunk_commit<false>(tmem_base, desc_hi | (uint64_t)((sa >> 4) & 0x3FFF), desc_hi | (uint64_t)((sb >> 4) & 0x3FFF), idesc_f,
                                 c > 0 ? 1u : 0u, 4u, smem_u32(w_empty + slot));
        if (++slot == 2) { slot = 0; phase ^= 1; }
      }
      umma_commit_elect<false>(smem_u32(f_full));
    };
    auto issue_z = [&](int j) {
      const int b = j & 1;
      mbar_wait(z_empty, ((uint32_t)j & 1) ^ 1);
      mbar_wait(u_full + b, (uint32_t)(j >> 1) & 1);
      tcgen05_fence_after();
      for (int c = 0; c < 4; ++c) {
        const uint32_t sa = tile0 + (uint32_t)b * T_TILE_BYTES + (uint32_t)c * A_BYTES, sb = wz0 + (uint32_t)c * a.NzBox * 128;
        umma_chunk(tmem_base + 256, desc_hi | (uint64_t)((sa >> 4) & 0x3FFF), desc_hi | (uint64_t)((sb >> 4) & 0x3FFF), idesc_z, c > 0 ? 1u : 0u, 4u);
      }
      umma_commit_elect<false>(smem_u32(z_full));
      umma_commit_elect<false>(smem_u32(tile_free + b));
    };
    if (n_my > 0) issue_f(0);
    for (int j = 0; j < n_my; ++j) {
      if (j + 1 < n_my) issue_f(j + 1);
      issue_z(j);
    }
  } else {
    // ===================== epilogue (warps 3..18) =====================
    const int lg = warp & 3;                    // TMEM lane group this warp may access
    const int qtr = (warp - 3) >> 2;            // F: which 64 of the 256 accumulator columns; Z: only quarter 0 takes part
    const int r = lg * 32 + lane;
    const bool fp16 = a.fp16 != 0;
    const EpiArgs& e = a.e;
    const uint32_t tile_u32 = smem_u32(tile_base);
    double ls_sum = 0.0;
    pdl_wait();
    auto epi_f = [&](int j) {
      const int b = j & 1;
      mbar_wait(f_full, (uint32_t)j & 1);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(qtr * 64);
      uint32_t v[64];
      tmem_ld_x16(taddr, v);
      tmem_ld_x16(taddr + 16, v + 16);
      tmem_ld_x16(taddr + 32, v + 32);
      tmem_ld_x16(taddr + 48, v + 48);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(f_empty);
      // relu(u) over this thread's 64 columns of the A tile (the final conv's MMAs have completed: f_full), in the operand layout
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4 b0 = *reinterpret_cast<const float4*>(sbias_f + qtr * 64 + 8 * k);
        const float4 b1 = *reinterpret_cast<const float4*>(sbias_f + qtr * 64 + 8 * k + 4);
        const uint32_t p0 = pack16(fmaxf(__uint_as_float(v[8 * k]) + b0.x, 0.f), fmaxf(__uint_as_float(v[8 * k + 1]) + b0.y, 0.f), fp16);
        const uint32_t p1 = pack16(fmaxf(__uint_as_float(v[8 * k + 2]) + b0.z, 0.f), fmaxf(__uint_as_float(v[8 * k + 3]) + b0.w, 0.f), fp16);
        const uint32_t p2 = pack16(fmaxf(__uint_as_float(v[8 * k + 4]) + b1.x, 0.f), fmaxf(__uint_as_float(v[8 * k + 5]) + b1.y, 0.f), fp16);
        const uint32_t p3 = pack16(fmaxf(__uint_as_float(v[8 * k + 6]) + b1.z, 0.f), fmaxf(__uint_as_float(v[8 * k + 7]) + b1.w, 0.f), fp16);
        sts128(tile_u32 + (uint32_t)b * T_TILE_BYTES + t_tile_off(r, qtr * 64 + 8 * k), make_uint4(p0, p1, p2, p3));
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(u_full + b);
    };
    auto epi_z = [&](int j) {
      if (qtr != 0) return;
      int ub, t0;
      coords(j, ub, t0);
      const int t = t0 + r;
      const bool row_ok = t < a.Ti && ub < a.B;
      const int64_t row = (int64_t)ub * a.Ti + t;
      float* xr = e.X + row * e.Cx;
      // fast path: the zero conv's column pairs are ordered by physical position, so 16 accumulator columns are 16 consecutive
      // floats of the thread's x row -- fetched before the accumulator wait
      const bool xfast = e.pairs_adjacent && e.Cx >= 16;
      float4 xv[8];
      if (xfast && row_ok) {
#pragma unroll
        for (int c0 = 0; c0 < 32; c0 += 16)
          if (c0 < a.Nz)
#pragma unroll
          for (int k = 0; k < 4; ++k) xv[(c0 >> 2) + k] = *reinterpret_cast<const float4*>(xr + c0 + 4 * k);
      }
      mbar_wait(z_full, (uint32_t)j & 1);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + 256u;
      uint32_t v[32];
      tmem_ld_x16(taddr, v);
      if (a.NzBox > 16) tmem_ld_x16(taddr + 16, v + 16);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(z_empty);
      if (!row_ok) return;
#pragma unroll
      for (int c0 = 0; c0 < 32; c0 += 16) {
        if (c0 >= a.Nz) break;
        if (xfast) {
          const float4* bp = reinterpret_cast<const float4*>(e.an_b + c0);
          const float4* sp = reinterpret_cast<const float4*>(e.an_s + c0);
          const int bo = e.b_odd;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float4 b4 = __ldg(bp + k), s4 = __ldg(sp + k);
            const float4 bias4 = *reinterpret_cast<const float4*>(sbias_z + c0 + 4 * k);
            const float4 xq = xv[(c0 >> 2) + k];
            float x[4] = {xq.x, xq.y, xq.z, xq.w};
            const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w};
            const float ac[4] = {__uint_as_float(v[c0 + 4 * k]) + bias4.x, __uint_as_float(v[c0 + 4 * k + 1]) + bias4.y,
                                 __uint_as_float(v[c0 + 4 * k + 2]) + bias4.z, __uint_as_float(v[c0 + 4 * k + 3]) + bias4.w};
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
          // (2q, 2q+1) = (log_s, t) of transformed element q; ActNorm + coupling in place on the fp32 flow variable
#pragma unroll
          for (int p = 0; p < 8; ++p) {
            const int q = c0 / 2 + p;
            if (q >= e.nq) break;
            const float log_s = __uint_as_float(v[c0 + 2 * p]) + sbias_z[c0 + 2 * p], tt = __uint_as_float(v[c0 + 2 * p + 1]) + sbias_z[c0 + 2 * p + 1];
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
    };
    if (n_my > 0) epi_f(0);
    for (int j = 0; j < n_my; ++j) {
      if (j + 1 < n_my) epi_f(j + 1);
      epi_z(j);
    }
    if (!e.reverse && e.logdet_acc && qtr == 0) {
      ls_sum = warp_sum(ls_sum);
      if (lane == 0 && ls_sum != 0.0) atomicAdd(e.logdet_acc, ls_sum);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

int launch_tail(const TailArgs& a, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    FWN_CUDA(cudaFuncSetAttribute(tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T_SMEM));
    configured = true;
  }
  const int n_tiles = a.B * a.tiles_per_utt;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)std::min(n_tiles, num_sms()));
  cfg.blockDim = dim3(T_THREADS);
  cfg.dynamicSmemBytes = T_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  int na = 0;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  FWN_CUDA(cudaLaunchKernelEx(&cfg, tail_kernel, a));
  FWN_LAUNCH_CHECK();
  return 0;
}

}  // namespace tc
}  // namespace fwn

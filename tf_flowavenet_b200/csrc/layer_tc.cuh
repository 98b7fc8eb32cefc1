// Arguments of the fused ResBlock-layer kernel (layer_tc.cu): gate GEMM -> tanh*sigmoid -> res|skip 1x1 in ONE launch.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace fwn {
namespace tc {

struct alignas(64) LayerArgs {
  CUtensorMap mapA[4];     // gate K segments: three time-shifted taps of h_in (+ the conditioning plane), (64, 128, 1) boxes
  CUtensorMap mapWg;       // gate weights [2F][Kpad] K-major, (64, 128) boxes (one CTA's half of a 256-column MMA)
  CUtensorMap mapWr;       // res|skip weights [2F or F][F] K-major, (64, 128) boxes
  CUtensorMap mapIn;       // staged epilogue input: h_in (has_res) or the running skip sum (last layer), (64, 128, 1) boxes
  CUtensorMap mapOutH;     // h_out, (64, 32, 1) store boxes
  CUtensorMap mapOutS;     // skip sum out, (64, 32, 1) store boxes
  int shift[4], nchunk[4], last_ksteps[4], wk0[4];
  int nseg;
  int has_res;             // columns [0, F) of the 1x1 are the residual conv (every layer but the last, modules.py:126-128)
  int has_in;              // the staging tile is pre-loaded (has_res: h_in; else: running skip sum)
  int relu;                // last layer: relu(sum of skips) feeds Conv_final (modules.py:176-177)
  int fp16;
  int B, Ti, tiles_per_utt;
  const float* gate_bias;  // [2F] (filter, gate) interleaved
  const float* rs_bias;    // [2F] or [F]
  // last layer only: the WaveNet tail in the same launch -- relu(skip sum) stays in the staging tile as the A operand of the final 1x1,
  // relu(final) replaces it in place as the A operand of the zero conv, whose epilogue is ActNorm + affine coupling on x
  int tail;                // 1: ops F (final conv) and Z (zero conv + affine) follow every tile's skip op; nothing is stored by the skip op
  CUtensorMap mapWf;       // final-conv weights [F][F] K-major, (64, 128) boxes
  CUtensorMap mapWz;       // zero-conv weights [Npad][F] K-major, (64, NzBox / 2) boxes (one CTA's half of the second MMA's B operand)
  int Nz, NzBox;           // 2 nq (log_s, t) columns; N of the zero-conv MMA (16 or 32)
  const float* final_bias; // [F]
  EpiArgs ez;              // the affine epilogue's arguments (bias = zero-conv bias)
  const float* pc;         // deep blocks: this layer's slice of the conditioning projection computed ahead, fp32 [rows, pc_ld]; else null
  int64_t pc_ld;
  int dbg;                 // diagnostics (FWN_LAYER_DBG bitmask): 1 no 1x1 TMA stores, 2 no staging loads, 4 no 1x1 epilogue math, 8 no activation loads for the second gate half, 16 no L2 prefetch (1..8: WRONG results)
  long long* trace;        // diagnostics (FWN_LAYER_TRACE=1): clock64 timeline of cluster 0's first tiles, else null
};

int launch_layer(const LayerArgs& a, cudaStream_t st);

}  // namespace tc
}  // namespace fwn

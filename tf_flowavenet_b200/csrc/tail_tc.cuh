// Arguments of the fused WaveNet-tail kernel (tail_tc.cu): final 1x1 + ReLU -> ZeroConv1d -> ActNorm + affine coupling on x.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace fwn {
namespace tc {

struct alignas(64) TailArgs {
  CUtensorMap mapS;        // relu(sum of skips) [B, Ti, F], (64, 128, 1) boxes
  CUtensorMap mapWf;       // final-conv weights [F][F] K-major, (64, 128) boxes
  CUtensorMap mapWz;       // zero-conv weights [Npad][F] K-major, (64, NzBox) boxes
  int B, Ti, tiles_per_utt;
  int Nz;                  // 2 nq zero-conv columns: (log_s, t) pairs
  int NzBox;               // rows of the weight box = N of the second MMA (16 or 32)
  int fp16;
  const float* final_bias; // [F]
  EpiArgs e;               // the AFFINE epilogue's arguments (bias = zero-conv bias)
};

int launch_tail(const TailArgs& a, cudaStream_t st);

}  // namespace tc
}  // namespace fwn

// Internal (C++) interface between the translation units of libflowavenet_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fwn {

// ---- elementwise.cu
int squeeze(const float* x, float* y, int B, int T, int C, cudaStream_t st);
int unsqueeze(const float* x, float* y, int B, int T, int C, cudaStream_t st);
int change_order(const float* x, float* y, int64_t rows, int C, cudaStream_t st);
int actnorm(const float* x, const float* b, const float* logs, float* y, float* logdet_out, int64_t rows, int C, bool rev, cudaStream_t st);
int actnorm_ddi(const float* x, float* b_out, float* logs_out, int64_t rows, int C, double* scratch2C, cudaStream_t st);
int affine(const float* x, const float* net, float* y, float* logdet_out, int64_t rows, int C, bool affine_, bool rev, double* scratch,
           cudaStream_t st);
int gated_activation(const float* f, const float* g, float* y, int64_t n, cudaStream_t st);
int residual_scale(const float* x, const float* r, float* y, int64_t n, cudaStream_t st);
int add(const float* a, const float* b, float* y, int64_t n, bool relu, cudaStream_t st);
int sumsq(const float* z, double* acc, int64_t n, cudaStream_t st);
int log_p(const float* z, float* out, int64_t n, double* scratch, cudaStream_t st);
int upsample_weight_norm(const float* v, const float* g, float* w, int s, cudaStream_t st);
int upsample_stage(const float* in, const float* w, const float* bias, void* out0, void* out1, int B, int Tm, int mels, int s, bool split,
                   int out_kind /* 0 fp32, 1 bf16, 2 fp16 */, cudaStream_t st);

// ---- implicit-GEMM description shared by the CUDA-core (fp32) and tcgen05 (bf16) engines.
//   acc[m, n] = sum_seg sum_{k < seg.K} A_seg[row(m) shifted by seg.shift in time, k] * W[seg.koff + k, n]
// Rows are (b, t) pairs, t in [0, Ti); a shifted row outside [0, Ti) reads as zero -- that IS the
// tf.pad of modules.py:27, applied per utterance.
struct Seg {
  const void* A;   // [B*Ti, lda] activations (float or bf16)
  int64_t lda;
  int shift;       // time shift in rows
  int K;           // valid reduction length
  int koff;        // first W row of this segment
};

enum EpiKind { EPI_PLAIN = 0, EPI_GATE = 1, EPI_RES_SKIP = 2, EPI_AFFINE = 3,
               // backward pass (fp32 engines only):
               EPI_LINEAR = 4,    // y = alpha * (acc + bias? + in0?), optionally masked by (in1 > 0)  (dgrad of the 1x1 / dilated convs)
               EPI_GATE_BWD = 5   // acc = dL/do -> (dL/df, dL/dg) interleaved, from the saved pre-activations (modules.py:124)
};

struct EpiArgs {
  const float* bias;       // [N]
  const float* colscale;   // [N] or null (PLAIN: acc*colscale + bias)
  void* out0;              // PLAIN: y; GATE: o [rows,F]; RES_SKIP: h_out [rows,F]
  void* out1;              // RES_SKIP: skip_out [rows,F]; training: GATE saves pre-activations [rows,2F], AFFINE saves (log_s,t) [rows,2nq]
  const void* in0;         // RES_SKIP: h_in
  const void* in1;         // RES_SKIP: skip_in (nullable -> no accumulate)
  int64_t ld;              // leading dim of PLAIN output
  int relu;                // PLAIN: relu; RES_SKIP: relu on skip output
  int has_res;             // RES_SKIP: columns [0,F) are the residual conv
  int F;                   // filter size
  float alpha;             // LINEAR: output scale
  // AFFINE (zero-conv epilogue == ActNorm + coupling in place on the flow variable)
  float* X;                // [rows, Cx] fp32, physical (time-ordered) layout
  int Cx, nq;
  const int* a_off;        // [nq] physical offsets of the pass-through half
  const int* b_off;        // [nq] physical offsets of the transformed half
  const float* an_b;       // [Cx] ActNorm bias (physical order)
  const float* an_s;       // [Cx] exp(3 logs) (forward) or exp(-3 logs) (reverse)
  double* logdet_acc;      // sum of log_s (forward)
  int reverse;
  int pairs_adjacent, b_odd;  // fast path: pair p of the zero conv's columns = x offsets (2p, 2p+1)
};

struct GemmArgs {
  Seg seg[4];
  int nseg;
  const void* W;     // fp32 engine: [Ktot, ldw] row-major fp32
  int64_t ldw;
  int N;
  int B, Ti;
  EpiArgs e;
};

// ---- conv_simt.cu (fp32 CUDA-core engine)
int simt_gemm(const GemmArgs& a, EpiKind kind, cudaStream_t st);
int weight_norm_scale(const float* v, const float* g, float* scale, int K, int Cout, cudaStream_t st);
int exp3(const float* s, float* out, int n, cudaStream_t st);
// front conv of the coupling WaveNet, reading the flow variable X directly (optional ActNorm on load)
struct FrontArgs {
  const float* X; int Cx; int nq; const int* a_off; const int* off2log;
  const float* an_b; const float* an_s;  // null -> identity (reverse direction)
  const float* W;   // [3][nq][F] fp32
  const float* bias; // [F]
  void* H;          // [rows, F] float or bf16
  int B, Ti, F;
  int shift[3];
};
int front_conv(const FrontArgs& a, bool bf16_out, cudaStream_t st);
// mixed mode: gather + ActNorm + bf16 cast of the pass-through half: A0[row, q] = a(row, q), q < nq, row pitch kq (zero padded)
int front_pack(const float* X, int Cx, int nq, int kq, const int* off2log, const float* an_b, const float* an_s, void* A0, int64_t rows,
               bool fp16, cudaStream_t st);

}  // namespace fwn

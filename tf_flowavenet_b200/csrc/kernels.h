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
               EPI_GATE_BWD = 5,  // acc = dL/do -> (dL/df, dL/dg) interleaved, from the saved pre-activations (modules.py:124)
               // 16-bit training mode (tcgen05 engine only):
               EPI_PLAIN_F32 = 6  // y (fp32, row pitch ld) = acc or y += acc: narrow / accumulating outputs (front-conv dgrad, conditioning gradient)
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
  // 16-bit training mode (gemm_tc.cu)
  void* tape;              // GATE: sigmoid(g) [rows, F] saved for the backward pass (with o it determines both gate derivatives)
  int mask_mode;           // LINEAR: in0 is a ReLU mask (y = in0 > 0 ? alpha acc : 0) instead of an addend (y = alpha (acc + in0))
  int accum;               // PLAIN_F32: y += acc
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

// ---- gemm_tc.cu: one 16-bit implicit GEMM on the tcgen05 engine with operands anywhere in device memory (the training step's tape;
// the inference passes go through the planned workspace of tc_prepare / tc_run).  Tensor maps are cached per (pointer, shape).
//   kind: GATE (out0 = o, e.tape = sigmoid(g)), RES_SKIP, PLAIN (16-bit out, bias / relu), LINEAR (16-bit out, staged addend or mask),
//         PLAIN_F32 (fp32 out, optional accumulate), AFFINE (in place on X; e.out1 = (log_s, t) fp32 with row pitch e.ld)
//   W: 16-bit [Npad][Kpad] K-major (plane 0 of the split engine's operand planes)
int tc_gemm16(const GemmArgs& g, EpiKind kind, const void* W, int Kpad, int Npad, bool fp16, cudaStream_t st);

// ---- wgrad_tc.cu: dW[koff + k, n] += sum_{b,t} A[b, t + shift, k] dY[b, t, n] with bf16 operands (see the file header)
struct Wgrad16Args {
  Seg seg[4];                                  // A operands (bf16), K = channels, koff = first dW row
  int nseg;
  const void* dY0; int64_t ld0; int n0cols;    // dY columns [0, n0cols) come from dY0, [n0cols, N) from dY1 (bf16)
  const void* dY1; int64_t ld1;
  int N;
  float* dW; int64_t ldw;                      // [Ktot, ldw] fp32, accumulated with atomics
  int B, Ti;
};
bool wgrad_tc_supported(const Wgrad16Args& a);
int wgrad_tc(const Wgrad16Args& a, float* dbias, cudaStream_t st);

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
// mixed modes, nq <= 4: the same conv straight from X on the CUDA cores, 16-bit output (write-bound; no operand packing)
bool front_direct_supported(const FrontArgs& a);
int front_direct(const FrontArgs& a, bool fp16, cudaStream_t st);
// mixed mode: gather + ActNorm + bf16 cast of the pass-through half: A0[row, q] = a(row, q), q < nq, row pitch kq (zero padded)
int front_pack(const float* X, int Cx, int nq, int kq, const int* off2log, const float* an_b, const float* an_s, void* A0, int64_t rows,
               bool fp16, cudaStream_t st);

}  // namespace fwn

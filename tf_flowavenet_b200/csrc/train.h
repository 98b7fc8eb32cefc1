// Training-step internals (train.cu, train_kernels.cu).  Reference: train.py:15-32,56-81, utils.py:34-60.
#pragma once
#include <cuda_bf16.h>

#include <vector>

#include "model.h"

namespace fwn {

struct FoldWork { int desc; int col0; int vec; };   // vec: 128-column tile with 16-byte accesses, else 32-column tile

// one fp32 operand (or a K range of one) -> bf16x3 planes:  dst[p][n][k0 + k] = split_p(src[k sk + n sn])
struct PlaneDesc {
  const float* src;
  int64_t sk, sn;
  int K, N;
  __nv_bfloat16* dst;
  int Kpad, Npad, k0;
};
struct PlaneWork { int desc, kt, nt; };

struct ActnormDesc {
  const float *raw_b, *raw_logs;
  float *an_b, *an_s, *an_is;
  const int* off2log;
  int Cx;
};

struct WgradArgs {
  Seg seg[4];
  int nseg;
  const float* dY0; int64_t ld0; int n0cols;   // dY columns [0, n0cols) come from dY0, [n0cols, N) from dY1
  const float* dY1; int64_t ld1;
  int N;
  float* dW; int64_t ldw;                      // [Ktot, ldw] fp32, accumulated with atomics
  int B, Ti;
  int slabs_per_utt;                           // set by wgrad()
};

// ---- training state (train.cu builds it; train16.cu, the 16-bit mode, shares it)
struct TrainFlow {
  W3 zero_T, final_T, front_T;
  W3 rs_T[MAX_LAYERS], gate_T[MAX_LAYERS], cond_T[MAX_LAYERS];
};

struct TrainState {
  float* what = nullptr;      // [raw + ext] folded parameters
  int32_t* wmap = nullptr;    // [wall] gather map
  float* gwall = nullptr;     // [wall] gradients of the packed fp32 operands / bias vectors
  FoldDesc* d_folds = nullptr;
  FoldWork* d_fwork = nullptr;
  int n_fwork = 0;
  PlaneDesc* d_pdesc = nullptr;
  PlaneWork* d_pwork = nullptr;
  int n_pwork = 0;
  ActnormDesc* d_an = nullptr;
  char* planes = nullptr;     // transposed bf16x3 planes (dgrad operands)
  std::vector<TrainFlow> flows;
  float *adam_m = nullptr, *adam_v = nullptr;
  double* scratch = nullptr;  // [8]
  float* norm = nullptr;      // [1]
  double* up_dw = nullptr;    // [max 2s*3 + 1]
  // backward pass: weight gradients and the conditioning gradient are off the critical chain of dgrads -> low-priority side stream
  cudaStream_t side = nullptr;
  std::vector<cudaEvent_t> ev;
  size_t ev_next = 0;
  cudaEvent_t set_done[2] = {nullptr, nullptr};
  std::vector<cudaEvent_t> cond_ready;   // per block: the conditioning projections of its flows are in the tape
  // Gradient buckets in PRODUCTION order (block n-1 first, ..., block 0, then the upsampler / speaker-embedding rest): contiguous
  // ranges of the flat gradient whose values are final once `ready` has fired, so the tower average of bucket k (an all-reduce on
  // the caller's communication stream, utils.py:34-60) overlaps the backward pass of the blocks still to come.
  struct Bucket { int64_t off, count; int64_t wall0, wall1; int work0, work1; cudaEvent_t ready; };
  std::vector<Bucket> buckets;
  // 16-bit mode: the whole loss-and-gradients pass (~1 400 launches on two streams) captured once per (shape, buffers) and replayed.
  // While capturing, the bucket events are recorded as EXTERNAL event nodes so that the caller's communication stream can still wait
  // on them after every replay.
  struct StepGraph { int B, T, dual; const void *ws, *grads, *lp, *ld; int state; cudaGraphExec_t exec; int64_t launches; };
  std::vector<StepGraph> step_graphs;
  bool capturing = false;
  cudaEvent_t next_event() {
    if (ev.empty()) {
      ev.resize(64);
      for (auto& e : ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    }
    cudaEvent_t e = ev[ev_next];
    ev_next = (ev_next + 1) % ev.size();
    return e;
  }
};

// ---- train_kernels.cu
int fold_forward(const float* raw, float* what, const FoldDesc* descs, const FoldWork* work, int nwork, int64_t raw_floats, cudaStream_t st);
int fold_backward(const float* raw, float* G, const FoldDesc* descs, const FoldWork* work, int nwork, int64_t raw_floats, cudaStream_t st);
int gather_pack(const float* what, const int32_t* map, float* P, int64_t n, cudaStream_t st);
int scatter_grad(const float* gP, const int32_t* map, float* G, int64_t n, cudaStream_t st);
int make_planes(const PlaneDesc* descs, const PlaneWork* work, int nwork, int nplanes, cudaStream_t st);
int actnorm_pack(const ActnormDesc* descs, int nflows, double* an_logdet, cudaStream_t st);
int wgrad(const WgradArgs& a, cudaStream_t st);
int colsum(const float* dY0, int64_t ld0, int n0cols, const float* dY1, int64_t ld1, int N, int64_t rows, float* out, cudaStream_t st);
int logp_bwd(const float* z, float* dX, int64_t n, cudaStream_t st);
int affine_bwd(float* dX, const float* Xpost, const float* net, int64_t ldn, float* dNet, int64_t rows, int Cx, int nq, const int* b_off,
               double n_total, cudaStream_t st);
int actnorm_bwd(float* dX, const float* da0, int ld_a0, const float* xpre, const float* an_b, const float* an_s, const int* off2log,
                int64_t rows, int Cx, int nq, float* g_b, float* g_logs, cudaStream_t st);
int front_pack_f32(const float* X, int Cx, int nq, int kq, const int* off2log, const float* an_b, const float* an_s, float* A0, int64_t rows,
                   cudaStream_t st);
int upsample_bwd_stage(const float* dout0, const float* dout1, const float* out0, const float* out1, bool split, const float* in,
                       const float* w, double* dw_scratch, float* din, int B, int Tm, int mels, int s, cudaStream_t st);
int upsample_wn_bwd(const float* v, const float* g, const double* dw, int s, float* gv, float* gg, float* gbias, cudaStream_t st);
int grad_global_norm(const float* g, int64_t n, double* scratch, float* norm_out, cudaStream_t st);
int adam_update(float* p, float* m, float* v, const float* g, const float* norm, float clip, float lr, float b1, float b2, float eps,
                int64_t step, int64_t n, cudaStream_t st);

// ---- wgrad_tc3.cu (tensor-core wgrad, fp32-accurate)
bool wgrad_tc3_supported(const WgradArgs& a);
int wgrad_tc3(const WgradArgs& a, float* dbias, int nterms, cudaStream_t st);

// ---- gemm_tc3.cu
bool tc3_supported(const GemmArgs& g);
int tc3_gemm(const GemmArgs& g, EpiKind kind, const void* w3, int Kpad, int Npad, int nterms, cudaStream_t st);

// ---- model.cu helpers shared with the training pass
int run_upsample(const Model* m, const Workspace& w, const float* c_in, int B, int T, cudaStream_t st);
int finish_forward(const double* sums, const double* an_logdet, float* logp_out, float* logdet_out, double n, cudaStream_t st);

// ---- train_kernels16.cu (bf16 tape / gradients)
int affine_bwd16(float* dX, const float* Xpost, const float* net, int64_t ldn, void* dNet, int64_t ldd, int64_t rows, int Cx, int nq,
                 const int* b_off, double n_total, cudaStream_t st);
int gate_bwd16(const void* dO, const void* O, const void* S, void* dFG, int64_t n, cudaStream_t st);
int relu_mask16(void* Y, const void* H, int64_t n, cudaStream_t st);

// ---- train16.cu: the training step with bf16 operands / tape on the tcgen05 engine (fp32 master variables, fp32 accumulation)
int64_t train16_workspace_bytes(const Model* m, int B, int T);
int train16_loss_and_grads(Model* m, const float* x, const float* c, const int32_t* g, int B, int T, float* logp_out, float* logdet_out,
                           float* grads, int64_t grad_floats, void* ws, int64_t ws_bytes, cudaStream_t st);
// shared pieces of the two modes (train.cu)
float* train_gw(const Model* m, const void* P);        // gradient slot mirroring a packed fp32 operand / bias vector
bool train_dual_stream();
int train_finish_block(Model* m, int block, float* grads, cudaStream_t st);   // scatter + unfold of one block, bucket event
int train_upsampler_backward(Model* m, const float* cmel, const float* up0, float* dup0, const float* cA, const float* cB, const float* dcA,
                             const float* dcB, int B, int T, float* grads, cudaStream_t st);

// ---- train.cu
int train_enable(Model* m, cudaStream_t st);
int train_after_prepack(Model* m);
void train_free(Model* m);
int64_t train_workspace_bytes(const Model* m, int B, int T);
int64_t train_grad_floats(const Model* m);
int train_loss_and_grads(Model* m, const float* x, const float* c, const int32_t* g, int B, int T, float* logp_out, float* logdet_out,
                         float* grads, int64_t grad_floats, void* ws, int64_t ws_bytes, cudaStream_t st);
int train_bucket_count(const Model* m);
int train_bucket_range(const Model* m, int k, int64_t* off, int64_t* count);
int train_bucket_wait(const Model* m, int k, cudaStream_t consumer);
int train_grad_norm(Model* m, const float* grads, float* norm_out, cudaStream_t st);
int train_apply(Model* m, const float* grads, float lr, float beta1, float beta2, float eps, float clip_norm, int64_t step, cudaStream_t st);
int train_repack(Model* m, cudaStream_t st, bool plane0_only = false);
// the optimizer step of the bf16 mode refreshes plane 0 of the operand planes only; anything that reads all three (fp32 passes, the
// fp32 training mode) calls this first
int train_ensure_full_planes(Model* m, cudaStream_t st);
int train_state_copy(Model* m, int which, float* dst, const float* src, int64_t numel, cudaStream_t st);

}  // namespace fwn

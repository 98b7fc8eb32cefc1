// Internal model structures (host side).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/flowavenet_b200.h"
#include "kernels.h"

namespace fwn {

struct ParamDesc {
  std::string name;
  std::vector<int64_t> shape;
  int64_t numel;
  int64_t offset;  // floats into Model::raw
};

constexpr int MAX_LAYERS = 8;
constexpr int AHEAD_MIN_KC = 2560;   // blocks whose conditioning half has at least this many channels run their projections ahead

enum GemmId { GEMM_GATE0 = 0, GEMM_RS0 = MAX_LAYERS, GEMM_FINAL = 2 * MAX_LAYERS, GEMM_ZERO = 2 * MAX_LAYERS + 1, GEMM_FRONT = 2 * MAX_LAYERS + 2,
              GEMM_IDS = 2 * MAX_LAYERS + 3 };

// fp32 mode, split engine: the [K][N] matrix as three bf16 planes [3][Npad][Kpad] (w = w1 + w2 + w3), see gemm_tc3.cu
struct W3 { void* p; int Kpad, Npad; };

// Prepacked weights of one flow (ActNorm + AffineCoupling/WaveNet), device pointers.
struct FlowPack {
  int Cx, nq, Kc, cond_half;
  int pairs_adjacent, b_odd;   // pair p = physical offsets (2p, 2p+1); b_odd: the transformed element is the odd one
  int *a_off, *b_off, *off2log;
  float *an_b, *an_s, *an_is;      // ActNorm bias / exp(3 logs) / exp(-3 logs), physical order
  float *raw_b, *raw_logs;         // the reference-named variables (updated by DDI)
  float *front_w, *front_b;
  void* front_wtc;   // mixed mode: bf16 [F][Kpad], K index = tap*ceil16(nq) + q (tcgen05 B operand)
  int front_ld, front_k16;  // Kpad and ceil16(nq)
  void* gate_w[MAX_LAYERS]; float* gate_b[MAX_LAYERS];
  void* rs_w[MAX_LAYERS];   float* rs_b[MAX_LAYERS];
  void* final_w; float* final_b;
  void* zero_w;  float* zero_b;
  int gate_ld, rs_ld[MAX_LAYERS], final_ld, zero_ld;
  // mixed modes, deep blocks (K_c >= AHEAD_MIN_KC): the conditioning projections c_a . W_c of all flows / layers of the block do not
  // depend on the flow state, so one wide GEMM per (block, mel half) computes them ahead of the dependent chain (cond_w below) and
  // the gate GEMMs reduce over the 768 conv inputs only, adding their slice [rows, 2F] of the projection in the epilogue.
  int ahead;        // 1: this flow's gate GEMMs take the projection from Workspace::pc[cond_half]
  int ahead_slot;   // first of this flow's n_layer slices (of 2F columns) in the projection
  W3 w3[GEMM_IDS];   // indexed by GemmId; p == nullptr when absent (mixed mode, front conv)
};

struct Workspace {
  double* sums;   // [0] sum log_s, [1] sum z^2
  double* ddi;
  float* x;
  float* up[2];
  void *cA, *cB, *h0, *h1, *o, *s, *u;
  void* a0;       // mixed mode: bf16 [rows, ceil8(nq)] ActNorm'd pass-through half of x (A operand of the front conv)
  float* pc[2];   // mixed modes: conditioning projections of the current deep block, one per mel half: fp32 [rows_i, slices * 2F]
  size_t bytes;
};


struct TcPlan;  // tensor maps of the tcgen05 engine (gemm_tc.cu)
struct TrainState;  // training-step state (train.cu)

// How the folded parameter vector What (raw layout + `ext` slots) derives from the raw variables.  The packed GEMM operands are pure
// signed gathers of What (model_prepack builds the gather map), so the same two steps -- fold, gather -- run on the host at
// fwn_prepack and on the device after every optimizer step, and their transposes carry the gradients back (train.cu).
enum FoldKind {
  FOLD_WN = 0,    // weight norm (convolutional.py:73-83): What[a + k N + o] = v[k,o] g[o] rsqrt(max(sum_k v[k,o]^2, 1e-12));  a = kernel, b = g
  FOLD_ZERO = 1,  // ZeroConv1d (modules.py:51-56): What[a + k N + o] = W[k,o] e[o], What[b + o] = bias[o] e[o], e = exp(3 scale[o]); c = scale
  FOLD_SUM2 = 2   // What[raw_floats + d + j] = raw[a + j] + raw[b + j]   (conv bias + conditioning-conv bias of one gate column)
};
struct FoldDesc { int kind; int64_t a, b, c, d; int K, N; };

struct Model {
  fwn_config cfg;
  int hop = 1;
  std::vector<ParamDesc> params;
  std::unordered_map<std::string, int> index;
  float* raw = nullptr;
  int64_t raw_floats = 0;
  std::vector<FoldDesc> folds;
  std::unordered_map<std::string, int64_t> ext_of;  // first bias name of a FOLD_SUM2 -> its ext slot
  int64_t ext_floats = 0;
  int64_t wall_floats = 0;          // leading fp32 region of `pack`: every packed fp32 GEMM operand and bias vector, contiguous
  bool keep_map = false;            // fwn_train_enable: keep the gather map of that region
  std::vector<int32_t> host_wmap;   // [wall_floats] What index | sign << 30, -1 = constant
  TrainState* train = nullptr;
  // bf16 split terms per fp32 product on the tensor cores: 6 = every product down to 2^-24, 3 = ~2^-16 per product (fwn_set_split_terms)
  int terms_infer = 6, terms_train = 3, cur_terms = 6;
  // fp32 training, parity setting: run the FORWARD GEMMs of the training step on the CUDA-core engine (round-to-nearest FFMA chains).
  // The tensor cores' accumulator truncates, a bias the backward pass amplifies on cancellation-heavy gradients.
  bool train_exact_fwd = false, force_simt = false;
  // training compute dtype: false = fp32-accurate (split engine), true = bf16 operands / tape on the tcgen05 engine (train16.cu)
  bool train_bf16 = false;
  bool planes_partial = false;   // planes 1, 2 of the split engine's operand planes are stale (bf16 optimizer steps refresh plane 0 only)
  // mixed inference passes: 1 = fused ResBlock-layer kernel (layer_tc.cu), 0 = gate GEMM + res|skip GEMM as two launches, -1 = FWN_FUSE_LAYER / default
  int fuse_layer = -1;
  bool packed = false, rev_ok = false;
  char* pack = nullptr;
  size_t pack_bytes = 0;
  std::vector<FlowPack> flows;
  float* up_w[4] = {nullptr, nullptr, nullptr, nullptr};
  float* up_b[4] = {nullptr, nullptr, nullptr, nullptr};
  double* d_an_logdet = nullptr;
  // conditioning-ahead operands (mixed modes): per (block, mel half) 16-bit [N = slices * 2F][Kpad], K in the physical order of cA / cB
  struct CondAhead { void* w = nullptr; int N = 0, Kpad = 0, Kc = 0; };
  std::vector<CondAhead> cond_w;   // [n_block * 2]; w == nullptr where the block keeps the projection inside its gate GEMMs
  // tcgen05 engine state
  TcPlan* tc = nullptr;
  int plan_B = -1, plan_T = -1;
  const void* plan_ws = nullptr;
  // buffers owned by the *_host entry points
  void* host_ws = nullptr;
  int64_t host_ws_bytes = 0;
  void* host_io = nullptr;
  int64_t host_io_bytes = 0;
  cudaStream_t host_stream = nullptr;  // non-blocking stream of the *_host entry points (capturable, unlike stream 0)
  int64_t launches = 0;  // kernels launched by the last pass (bench.py's gpu_launches)
  // CUDA graphs of the per-pass flow chain, keyed by (direction, B, T, workspace): the chain is ~200-400 dependent
  // launches, so replaying one graph removes the per-launch gaps (decisive for short utterances)
  struct PassGraph { int reverse, B, T; const void* ws; int state; cudaGraphExec_t exec; int64_t launches; };
  std::vector<PassGraph> graphs;
  // optional per-kernel-family CUDA-event timing (fwn_profile_*): bench.py's roofline numbers
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_ev;
  std::vector<int> prof_kind;
  std::vector<double> prof_work;
  size_t prof_used = 0;
};

enum ProfKind { PROF_FRONT = 0, PROF_GATE = 1, PROF_RES_SKIP = 2, PROF_FINAL = 3, PROF_ZERO_AFFINE = 4, PROF_UPSAMPLE = 5, PROF_OTHER = 6, PROF_KINDS = 8 };
void prof_begin(Model* m, int kind, double work, cudaStream_t st);
void prof_end(Model* m, cudaStream_t st);
int prof_read(Model* m, double* ms, int64_t* launches, double* work);

int model_create(const fwn_config* cfg, Model** out);
void model_destroy(Model* m);
int model_prepack(Model* m, cudaStream_t st);
int model_plan(const Model* m, int B, int T, Workspace* w, char* base);
int model_forward(Model* m, const float* x, const float* c, const int32_t* g, int B, int T, float* z_out, float* logp_out,
                  float* logdet_out, int ddi, void* ws, int64_t ws_bytes, cudaStream_t st);
int model_reverse(Model* m, const float* z, const float* c, const int32_t* g, int B, int T, float* x_out, void* ws, int64_t ws_bytes,
                  cudaStream_t st);
int model_receptive_halo(const Model* m);
void model_drop_graphs(Model* m);

// engine dispatch (engine.cu): fp32 -> CUDA-core implicit GEMM, mixed -> tcgen05
int prepare_engine(Model* m, const Workspace& w, int B, int T, cudaStream_t st);
int run_gemm(Model* m, const GemmArgs& g, EpiKind kind, int gemm_id, const FlowPack& fp, cudaStream_t st);
void engine_free(Model* m);
// fused ResBlock layer on the tcgen05 engine (gemm_tc.cu / layer_tc.cu): gate GEMM `g` + res|skip 1x1 `r` of layer `layer`
bool tc_layer_supported(const Model* m, const GemmArgs& g, const GemmArgs& r);
int train_ensure_full_planes(Model* m, cudaStream_t st);   // train.cu
int tc_run_layer(Model* m, const GemmArgs& g, const GemmArgs& r, int layer, const FlowPack& fp, cudaStream_t st, const GemmArgs* fin = nullptr,
                 const GemmArgs* zero = nullptr);   // fin, zero: fold the WaveNet tail into the last layer's launch
// fused WaveNet tail (gemm_tc.cu / tail_tc.cu): final 1x1 `f` + zero conv / affine coupling `z`
bool tc_tail_supported(const Model* m, const GemmArgs& f, const GemmArgs& z);
int tc_run_tail(Model* m, const GemmArgs& f, const GemmArgs& z, const FlowPack& fp, cudaStream_t st);

}  // namespace fwn

// Training step on the fp32 parity engines: forward with a tape, hand-written backward, device-side re-pack, Adam.
// Reference: train.py:56-81 (loss = -(log_p + logdet); tf.gradients over all trainable variables; average over towers;
// clip_by_global_norm 1; AdamOptimizer), train.py:15-24 (learning-rate schedule), model.py / modules.py for the graph itself.
//
// Gradient flow.  The packed GEMM operands are signed gathers of the folded parameter vector What (model_prepack), so
//   raw --fold--> What --gather--> packed operands --(forward, tape)--> loss
//   d raw <--unfold-- d What <--scatter-- d packed <--(wgrad / colsum)-- backward
// dgrad GEMMs run on the same implicit-GEMM engines as the forward pass (time shifts negated, transposed bf16x3 weight planes),
// with two extra epilogues (EPI_LINEAR, EPI_GATE_BWD).  The conditioning gradient accumulates into two mel-half planes that
// mirror the forward cA / cB layout, so squeeze and change_order need no backward kernels either.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "train.h"

namespace fwn {

void train_free(Model* m) {
  TrainState* t = m->train;
  if (!t) return;
  cudaFree(t->what); cudaFree(t->wmap); cudaFree(t->gwall); cudaFree(t->d_folds); cudaFree(t->d_fwork); cudaFree(t->d_pdesc);
  cudaFree(t->d_pwork); cudaFree(t->d_an); cudaFree(t->planes); cudaFree(t->adam_m); cudaFree(t->adam_v); cudaFree(t->scratch);
  cudaFree(t->norm); cudaFree(t->up_dw);
  for (auto e : t->ev) cudaEventDestroy(e);
  for (auto e : t->set_done) if (e) cudaEventDestroy(e);
  for (auto e : t->cond_ready) cudaEventDestroy(e);
  for (auto& b : t->buckets) if (b.ready) cudaEventDestroy(b.ready);
  for (auto& g : t->step_graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
  if (t->side) cudaStreamDestroy(t->side);
  delete t;
  m->train = nullptr;
}

template <class T>
static int upload(T** dst, const std::vector<T>& v) {
  *dst = nullptr;
  if (v.empty()) return 0;
  FWN_CUDA(cudaMalloc(dst, v.size() * sizeof(T)));
  FWN_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

static inline int ceil4(int x) { return (x + 3) & ~3; }

// Build the training state for the CURRENT pack buffer (called at the end of model_prepack when keep_map is set).
static int train_build(Model* m) {
  const fwn_config& c = m->cfg;
  FWN_CHECK(c.precision == FWN_FP32, "training runs on the fp32 engines: create the model with precision FWN_FP32");
  FWN_CHECK(c.n_upsample <= 2, "training supports at most two upsampling stages");
  FWN_CHECK((int64_t)m->host_wmap.size() == m->wall_floats, "internal: gather map missing");
  const int F = c.filter_size, L = c.n_layer;
  // keep optimizer state across a re-prepack
  float *keep_m = nullptr, *keep_v = nullptr;
  if (m->train) { keep_m = m->train->adam_m; keep_v = m->train->adam_v; m->train->adam_m = m->train->adam_v = nullptr; train_free(m); }
  TrainState* t = new TrainState();
  m->train = t;
  const int64_t nwhat = m->raw_floats + m->ext_floats;
  FWN_CUDA(cudaMalloc(&t->what, (size_t)nwhat * 4));
  if (upload(&t->wmap, m->host_wmap)) return 1;
  m->host_wmap.clear();
  m->host_wmap.shrink_to_fit();
  FWN_CUDA(cudaMalloc(&t->gwall, (size_t)m->wall_floats * 4));
  if (upload(&t->d_folds, m->folds)) return 1;
  // first variable (flat offset) of every block; blocks are contiguous in the flat layout: [upsampler | Block_0 | ... | Block_n-1 | rest]
  std::vector<int64_t> blk_off(c.n_block + 1);
  for (int i = 0; i < c.n_block; ++i) blk_off[i] = m->params[m->index.at("Block_" + std::to_string(i) + "/Flow_0/ActNorm/b")].offset;
  blk_off[c.n_block] = c.gin_channels > 0 ? m->params[m->index.at("speaker_embeddings")].offset : m->raw_floats;
  auto block_of = [&](int64_t off) { return (int)(std::upper_bound(blk_off.begin(), blk_off.end(), off) - blk_off.begin()) - 1; };
  std::vector<FoldWork> fw;
  std::vector<int> fw_begin(c.n_block + 1, 0);
  for (int blk = 0; blk < c.n_block; ++blk) {
    const size_t first = fw.size();
    for (size_t i = 0; i < m->folds.size(); ++i) {
      if (block_of(m->folds[i].a) != blk) continue;
      // wide (16-byte) tiles for the many K <= 768 operands; the few long ones (conditioning convs of the deep blocks, K up to 5120)
      // keep 32-column tiles so that their CTAs do not become the tail of the launch
      const int vec = ((m->folds[i].N & 3) == 0 && m->folds[i].K <= 1024) ? 1 : 0;
      for (int c0 = 0; c0 < m->folds[i].N; c0 += vec ? 128 : 32) fw.push_back(FoldWork{(int)i, c0, vec});
    }
    std::stable_sort(fw.begin() + first, fw.end(), [&](const FoldWork& x, const FoldWork& y) { return m->folds[x.desc].K > m->folds[y.desc].K; });   // longest first
    fw_begin[blk + 1] = (int)fw.size();
  }
  t->n_fwork = (int)fw.size();
  if (upload(&t->d_fwork, fw)) return 1;
  {
    // wall (packed fp32 operand) range of every block: flows are packed in order, each starting with its front-conv kernel
    std::vector<int64_t> wall_begin(c.n_block + 1, m->wall_floats);
    for (int i = 0; i < c.n_block; ++i)
      wall_begin[i] = reinterpret_cast<const float*>(m->flows[(size_t)i * c.n_flow].front_w) - reinterpret_cast<const float*>(m->pack);
    for (int i = c.n_block - 1; i >= 0; --i) {
      TrainState::Bucket b{blk_off[i], blk_off[i + 1] - blk_off[i], wall_begin[i], wall_begin[i + 1], fw_begin[i], fw_begin[i + 1], nullptr};
      t->buckets.push_back(b);
    }
    t->buckets.push_back(TrainState::Bucket{0, blk_off[0], 0, 0, 0, 0, nullptr});                                     // upsampler variables
    if (blk_off[c.n_block] < m->raw_floats)
      t->buckets.push_back(TrainState::Bucket{blk_off[c.n_block], m->raw_floats - blk_off[c.n_block], 0, 0, 0, 0, nullptr});   // speaker embeddings
    for (auto& b : t->buckets) FWN_CUDA(cudaEventCreateWithFlags(&b.ready, cudaEventDisableTiming));
  }

  // ---- plane descriptors: forward planes (refreshed after every optimizer step) + transposed planes for dgrad
  std::vector<PlaneDesc> pd;
  size_t pbytes = 0;
  auto fwd = [&](const void* P, int K, int N, const W3& w3) {
    if (!w3.p) return;
    pd.push_back(PlaneDesc{reinterpret_cast<const float*>(P), N, 1, K, N, reinterpret_cast<__nv_bfloat16*>(w3.p), w3.Kpad, w3.Npad, 0});
  };
  auto alloc_T = [&](int Kd, int Nd, W3* out) {
    out->Kpad = (Kd + 63) / 64 * 64;
    out->Npad = (Nd + 15) / 16 * 16;
    const size_t off = pbytes;
    pbytes += ((size_t)3 * out->Kpad * out->Npad * 2 + 255) & ~size_t(255);
    out->p = reinterpret_cast<void*>(off);   // fixed up below
    return off;
  };
  auto tr = [&](const float* src, int64_t sn, int K, int N, const W3& w3, int k0) {
    pd.push_back(PlaneDesc{src, 1, sn, K, N, reinterpret_cast<__nv_bfloat16*>(w3.p), w3.Kpad, w3.Npad, k0});
  };
  t->flows.resize(m->flows.size());
  for (size_t f = 0; f < m->flows.size(); ++f) {
    const FlowPack& fp = m->flows[f];
    const int nq = fp.nq, Kc = fp.Kc, Kg = 3 * F + (Kc + 15) / 16 * 16;
    for (int k = 0; k < 3; ++k)   // tap k of the front conv lives at K rows k * front_k16 of its planes
      pd.push_back(PlaneDesc{fp.front_w + (size_t)k * nq * F, F, 1, nq, F, reinterpret_cast<__nv_bfloat16*>(fp.w3[GEMM_FRONT].p),
                             fp.w3[GEMM_FRONT].Kpad, fp.w3[GEMM_FRONT].Npad, k * fp.front_k16});
    for (int n = 0; n < L; ++n) {
      fwd(fp.gate_w[n], Kg, 2 * F, fp.w3[GEMM_GATE0 + n]);
      fwd(fp.rs_w[n], F, n == L - 1 ? F : 2 * F, fp.w3[GEMM_RS0 + n]);
    }
    fwd(fp.final_w, F, F, fp.w3[GEMM_FINAL]);
    fwd(fp.zero_w, F, 2 * nq, fp.w3[GEMM_ZERO]);
  }
  const size_t n_fwd = pd.size();
  for (size_t f = 0; f < m->flows.size(); ++f) {
    const FlowPack& fp = m->flows[f];
    TrainFlow& tf = t->flows[f];
    const int nq = fp.nq, Kc = fp.Kc;
    alloc_T(2 * nq, F, &tf.zero_T);
    tr(reinterpret_cast<const float*>(fp.zero_w), 2 * nq, 2 * nq, F, tf.zero_T, 0);
    alloc_T(F, F, &tf.final_T);
    tr(reinterpret_cast<const float*>(fp.final_w), F, F, F, tf.final_T, 0);
    for (int n = 0; n < L; ++n) {
      const int Nr = n == L - 1 ? F : 2 * F;
      alloc_T(Nr, F, &tf.rs_T[n]);
      tr(reinterpret_cast<const float*>(fp.rs_w[n]), Nr, Nr, F, tf.rs_T[n], 0);
      const float* Pg = reinterpret_cast<const float*>(fp.gate_w[n]);
      alloc_T(6 * F, F, &tf.gate_T[n]);
      for (int k = 0; k < 3; ++k) tr(Pg + (size_t)k * F * 2 * F, 2 * F, 2 * F, F, tf.gate_T[n], k * 2 * F);
      alloc_T(2 * F, Kc, &tf.cond_T[n]);
      tr(Pg + (size_t)3 * F * 2 * F, 2 * F, 2 * F, Kc, tf.cond_T[n], 0);
    }
    alloc_T(3 * F, nq, &tf.front_T);
    for (int k = 0; k < 3; ++k) tr(fp.front_w + (size_t)k * nq * F, F, F, nq, tf.front_T, k * F);
  }
  FWN_CUDA(cudaMalloc(&t->planes, std::max<size_t>(pbytes, 256)));
  FWN_CUDA(cudaMemset(t->planes, 0, std::max<size_t>(pbytes, 256)));
  auto fix = [&](W3& w) { w.p = t->planes + reinterpret_cast<size_t>(w.p); };
  for (size_t i = n_fwd; i < pd.size(); ++i) pd[i].dst = reinterpret_cast<__nv_bfloat16*>(t->planes + reinterpret_cast<size_t>(pd[i].dst));
  for (auto& tf : t->flows) {
    fix(tf.zero_T); fix(tf.final_T); fix(tf.front_T);
    for (int n = 0; n < L; ++n) { fix(tf.rs_T[n]); fix(tf.gate_T[n]); fix(tf.cond_T[n]); }
  }
  std::vector<PlaneWork> pw;
  for (size_t i = 0; i < pd.size(); ++i)
    for (int kt = 0; kt < (pd[i].K + 31) / 32; ++kt)
      for (int nt = 0; nt < (pd[i].N + 31) / 32; ++nt) pw.push_back(PlaneWork{(int)i, kt, nt});
  t->n_pwork = (int)pw.size();
  if (upload(&t->d_pdesc, pd)) return 1;
  if (upload(&t->d_pwork, pw)) return 1;

  std::vector<ActnormDesc> an;
  for (const FlowPack& fp : m->flows) an.push_back(ActnormDesc{fp.raw_b, fp.raw_logs, fp.an_b, fp.an_s, fp.an_is, fp.off2log, fp.Cx});
  if (upload(&t->d_an, an)) return 1;

  if (keep_m) { t->adam_m = keep_m; t->adam_v = keep_v; }
  else {
    FWN_CUDA(cudaMalloc(&t->adam_m, (size_t)m->raw_floats * 4));
    FWN_CUDA(cudaMalloc(&t->adam_v, (size_t)m->raw_floats * 4));
    FWN_CUDA(cudaMemset(t->adam_m, 0, (size_t)m->raw_floats * 4));
    FWN_CUDA(cudaMemset(t->adam_v, 0, (size_t)m->raw_floats * 4));
  }
  FWN_CUDA(cudaMalloc(&t->scratch, 8 * sizeof(double)));
  FWN_CUDA(cudaMalloc(&t->norm, sizeof(float)));
  int smax = 2;
  for (int i = 0; i < c.n_upsample; ++i) smax = std::max(smax, c.upsample_scales[i]);
  FWN_CUDA(cudaMalloc(&t->up_dw, (size_t)(2 * smax * 3 + 1) * sizeof(double)));
  int lo = 0, hi = 0;
  FWN_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));   // lo = numerically greatest = lowest priority
  FWN_CUDA(cudaStreamCreateWithPriority(&t->side, cudaStreamNonBlocking, lo));
  for (auto& e : t->set_done) FWN_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  t->cond_ready.resize(c.n_block);
  for (auto& e : t->cond_ready) FWN_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  return 0;
}

int train_enable(Model* m, cudaStream_t st) {
  FWN_CHECK(m, "null handle");
  FWN_CHECK(m->cfg.precision == FWN_FP32, "training runs on the fp32 engines: create the model with precision FWN_FP32");
  m->keep_map = true;
  // builds the gather map and the training state, then re-derives every operand on the device (train_repack): the transposed dgrad
  // planes do not exist on the host
  return model_prepack(m, st);
}
int train_after_prepack(Model* m) { return m->keep_map ? train_build(m) : 0; }

// raw variables -> every derived operand, on the device (what fwn_prepack does on the host)
int train_repack(Model* m, cudaStream_t st, bool plane0_only) {
  TrainState* t = m->train;
  FWN_CHECK(t, "training not enabled: call fwn_train_enable first");
  const fwn_config& c = m->cfg;
  m->launches += 4 + 1 * c.n_upsample;
  if (fold_forward(m->raw, t->what, t->d_folds, t->d_fwork, t->n_fwork, m->raw_floats, st)) return 1;
  if (gather_pack(t->what, t->wmap, reinterpret_cast<float*>(m->pack), m->wall_floats, st)) return 1;
  if (make_planes(t->d_pdesc, t->d_pwork, t->n_pwork, plane0_only ? 1 : 3, st)) return 1;
  m->planes_partial = plane0_only;
  if (actnorm_pack(t->d_an, (int)m->flows.size(), m->d_an_logdet, st)) return 1;
  for (int i = 0; i < c.n_upsample; ++i) {
    const std::string n = i == 0 ? "conv2d_transpose" : "conv2d_transpose_" + std::to_string(i);
    const float* v = m->raw + m->params[m->index.at(n + "/kernel")].offset;
    const float* g = m->raw + m->params[m->index.at(n + "/wn/g")].offset;
    const float* b = m->raw + m->params[m->index.at(n + "/bias")].offset;
    if (upsample_weight_norm(v, g, m->up_w[i], c.upsample_scales[i], st)) return 1;
    FWN_CUDA(cudaMemcpyAsync(m->up_b[i], b, sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}
int train_ensure_full_planes(Model* m, cudaStream_t st) {
  if (!m->planes_partial || !m->train) return 0;
  TrainState* t = m->train;
  m->launches += 1;
  if (make_planes(t->d_pdesc, t->d_pwork, t->n_pwork, 3, st)) return 1;
  m->planes_partial = false;
  return 0;
}

// ---------------------------------------------------------------- workspace
struct Tape {   // saved activations of one flow
  float *xpre, *a0, *s, *u, *net;
  float *h[MAX_LAYERS], *fg[MAX_LAYERS], *o[MAX_LAYERS];
};
struct TrainWs {
  double* sums;
  double* ddi;
  float *X, *dX, *up0, *dup0, *cA, *cB, *dcA, *dcB;
  struct BwdSet { float *dnet, *da0, *du, *ds; float* dfg[MAX_LAYERS]; float* r[MAX_LAYERS]; } set[2];   // alternate per flow
  std::vector<Tape> tape;
  size_t bytes;
};
static inline size_t al256(size_t x) { return (x + 255) & ~size_t(255); }

static int train_plan(const Model* m, int B, int T, TrainWs* w, char* base) {
  const fwn_config& c = m->cfg;
  FWN_CHECK(B > 0 && T > 0, "empty input: B=%d T=%d", B, T);
  FWN_CHECK(T % m->hop == 0, "T=%d is not a multiple of the hop size %d", T, m->hop);
  FWN_CHECK(T % (1 << c.n_block) == 0, "T=%d is not a multiple of 2^n_block=%d", T, 1 << c.n_block);
  const int F = c.filter_size, H = c.num_mels / 2, L = c.n_layer;
  const size_t BT = (size_t)B * T, M0 = BT / 2;
  size_t off = 0;
  auto take = [&](size_t floats) {
    size_t o = off;
    off = al256(off + floats * 4);
    return base ? reinterpret_cast<float*>(base + o) : (float*)nullptr;
  };
  w->sums = reinterpret_cast<double*>(take(16));
  w->ddi = reinterpret_cast<double*>(take(2 * 4096 * 2));
  w->X = take(BT);
  w->dX = take(BT);
  const int s_last = c.upsample_scales[c.n_upsample - 1];
  const size_t up_elems = c.n_upsample > 1 ? (size_t)B * (T / s_last) * c.num_mels : 0;
  w->up0 = take(up_elems);
  w->dup0 = take(up_elems);
  w->cA = take(BT * H); w->cB = take(BT * H); w->dcA = take(BT * H); w->dcB = take(BT * H);
  for (auto& bs : w->set) {
    bs.dnet = take(2 * BT); bs.da0 = take(2 * BT);
    bs.du = take(M0 * F); bs.ds = take(M0 * F);
    for (int n = 0; n < L; ++n) { bs.dfg[n] = take(M0 * 2 * F); bs.r[n] = take(M0 * F); }
  }
  w->tape.assign(m->flows.size(), Tape{});
  for (int i = 0; i < c.n_block; ++i) {
    const size_t M = BT >> (i + 1);
    const int nq = 1 << i;
    for (int j = 0; j < c.n_flow; ++j) {
      Tape& tp = w->tape[(size_t)i * c.n_flow + j];
      tp.xpre = take(BT);
      tp.a0 = take(M * ceil4(nq));
      for (int n = 0; n < L; ++n) { tp.h[n] = take(M * F); tp.fg[n] = take(M * 2 * F); tp.o[n] = take(M * F); }
      tp.s = take(M * F); tp.u = take(M * F);
      tp.net = take(M * ceil4(2 * nq));
    }
  }
  w->bytes = off;
  return 0;
}
int64_t train_workspace_bytes(const Model* m, int B, int T) {
  if (m->train_bf16) return train16_workspace_bytes(m, B, T);
  TrainWs w;
  if (train_plan(m, B, T, &w, nullptr)) return -1;
  return (int64_t)w.bytes;
}
int64_t train_grad_floats(const Model* m) { return m->raw_floats + m->ext_floats; }

// ---------------------------------------------------------------- forward with tape / backward of one flow
static int bw_gemm(Model* m, const GemmArgs& g, EpiKind kind, const W3& w3, cudaStream_t st) {
  m->launches++;
  FWN_CHECK(w3.p && tc3_supported(g), "internal: backward GEMM operand not supported by the split engine");
  return tc3_gemm(g, kind, w3.p, w3.Kpad, w3.Npad, m->cur_terms, st);
}
// FWN_WGRAD = tc3 (default) | simt
static bool wgrad_on_tensor_cores() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FWN_WGRAD");
    v = (e && !strcmp(e, "simt")) ? 0 : 1;
  }
  return v == 1;
}
static inline int shift_of(const fwn_config& c, int k, int d) { return c.causal ? (k - 2) * d : (k - 1) * d; }

// Deep blocks (few rows per utterance, K_c in the thousands): the conditioning projections c_a . W_c of every (flow, layer) do not
// depend on the flow state, so they are computed ahead of the dependent chain -- on the side stream, over a flat row axis (full
// 128-row tiles instead of one mostly empty tile per utterance) -- straight into the pre-activation tape; the gate GEMM then
// reduces over the 768 conv inputs only and adds the projection in its epilogue.
static bool g_exact_fwd = false;   // set for the duration of a forward pass that runs on the CUDA-core engine
static bool cond_ahead(int Ti, const FlowPack& fp) {
  const char* e = getenv("FWN_COND_AHEAD");   // 0 disables (diagnostics)
  if ((e && e[0] == '0') || g_exact_fwd) return false;
  return Ti <= 256 && (fp.Kc & 3) == 0 && fp.w3[GEMM_GATE0].p != nullptr;
}
static int cond_forward(Model* m, const TrainWs& w, const FlowPack& fp, const Tape& tp, int B, int Ti, cudaStream_t st) {
  const int F = m->cfg.filter_size, L = m->cfg.n_layer;
  const float* cond = fp.cond_half == 0 ? w.cA : w.cB;
  for (int n = 0; n < L; ++n) {
    GemmArgs g = {};
    g.B = B; g.Ti = Ti;
    g.seg[0] = Seg{cond, fp.Kc, 0, fp.Kc, 3 * F};
    g.nseg = 1; g.N = 2 * F;
    g.e.out0 = tp.fg[n]; g.e.ld = 2 * F; g.e.alpha = 1.f; g.e.F = F;
    m->launches++;
    FWN_CHECK(fp.w3[GEMM_GATE0 + n].p && tc3_supported(g), "internal: conditioning projection not supported by the split engine");
    if (tc3_gemm(g, EPI_LINEAR, fp.w3[GEMM_GATE0 + n].p, fp.w3[GEMM_GATE0 + n].Kpad, fp.w3[GEMM_GATE0 + n].Npad, m->cur_terms, st)) return 1;
  }
  return 0;
}

static int flow_forward(Model* m, const TrainWs& w, const FlowPack& fp, const Tape& tp, int B, int Ti, cudaStream_t st) {
  const fwn_config& c = m->cfg;
  const int F = c.filter_size, L = c.n_layer, nq = fp.nq, nq4 = ceil4(nq);
  const int64_t rows = (int64_t)B * Ti;
  FWN_CUDA(cudaMemcpyAsync(tp.xpre, w.X, (size_t)rows * fp.Cx * 4, cudaMemcpyDeviceToDevice, st));
  m->launches++;
  if (front_pack_f32(w.X, fp.Cx, nq, nq4, fp.off2log, fp.an_b, fp.an_s, tp.a0, rows, st)) return 1;
  {
    GemmArgs g = {};
    g.B = B; g.Ti = Ti;
    for (int k = 0; k < 3; ++k) g.seg[k] = Seg{tp.a0, nq4, shift_of(c, k, 1), nq, k * fp.front_k16};
    g.nseg = 3;
    g.W = fp.front_w; g.ldw = F; g.N = F;
    g.e.bias = fp.front_b; g.e.out0 = tp.h[0]; g.e.ld = F; g.e.relu = 1; g.e.F = F;
    if (run_gemm(m, g, EPI_PLAIN, GEMM_FRONT, fp, st)) return 1;
  }
  const float* cond = fp.cond_half == 0 ? w.cA : w.cB;
  int d = 1;
  for (int n = 0; n < L; ++n, d *= 3) {
    GemmArgs g = {};
    g.B = B; g.Ti = Ti;
    for (int k = 0; k < 3; ++k) g.seg[k] = Seg{tp.h[n], F, shift_of(c, k, d), F, k * F};
    g.seg[3] = Seg{cond, fp.Kc, 0, fp.Kc, 3 * F};
    g.nseg = cond_ahead(Ti, fp) ? 3 : 4;
    g.W = fp.gate_w[n]; g.ldw = fp.gate_ld; g.N = 2 * F;
    g.e.bias = fp.gate_b[n]; g.e.out0 = tp.o[n]; g.e.out1 = tp.fg[n]; g.e.F = F;
    if (cond_ahead(Ti, fp)) g.e.in0 = tp.fg[n];   // cond_forward left the projection there
    if (run_gemm(m, g, EPI_GATE, GEMM_GATE0 + n, fp, st)) return 1;
    const bool last = n == L - 1;
    GemmArgs r = {};
    r.B = B; r.Ti = Ti;
    r.seg[0] = Seg{tp.o[n], F, 0, F, 0};
    r.nseg = 1;
    r.W = fp.rs_w[n]; r.ldw = fp.rs_ld[n]; r.N = last ? F : 2 * F;
    r.e.bias = fp.rs_b[n]; r.e.F = F; r.e.has_res = !last; r.e.relu = last;
    r.e.in0 = tp.h[n]; r.e.out0 = last ? nullptr : tp.h[n + 1]; r.e.in1 = n > 0 ? tp.s : nullptr; r.e.out1 = tp.s;
    if (run_gemm(m, r, EPI_RES_SKIP, GEMM_RS0 + n, fp, st)) return 1;
  }
  {
    GemmArgs g = {};
    g.B = B; g.Ti = Ti;
    g.seg[0] = Seg{tp.s, F, 0, F, 0};
    g.nseg = 1;
    g.W = fp.final_w; g.ldw = fp.final_ld; g.N = F;
    g.e.bias = fp.final_b; g.e.out0 = tp.u; g.e.ld = F; g.e.relu = 1; g.e.F = F;
    if (run_gemm(m, g, EPI_PLAIN, GEMM_FINAL, fp, st)) return 1;
  }
  {
    GemmArgs g = {};
    g.B = B; g.Ti = Ti;
    g.seg[0] = Seg{tp.u, F, 0, F, 0};
    g.nseg = 1;
    g.W = fp.zero_w; g.ldw = fp.zero_ld; g.N = 2 * nq;
    g.e.bias = fp.zero_b; g.e.F = F;
    g.e.X = w.X; g.e.Cx = fp.Cx; g.e.nq = nq; g.e.a_off = fp.a_off; g.e.b_off = fp.b_off;
    g.e.an_b = fp.an_b; g.e.an_s = fp.an_s;
    g.e.logdet_acc = w.sums;
    g.e.reverse = 0;
    g.e.pairs_adjacent = fp.pairs_adjacent;
    g.e.b_odd = fp.b_odd;
    g.e.out1 = tp.net; g.e.ld = ceil4(2 * nq);
    if (run_gemm(m, g, EPI_AFFINE, GEMM_ZERO, fp, st)) return 1;
  }
  return 0;
}

static float* gw(const Model* m, const void* P) {  // gradient slot mirroring a packed fp32 operand / bias vector
  return m->train->gwall + (reinterpret_cast<const float*>(P) - reinterpret_cast<const float*>(m->pack));
}
float* train_gw(const Model* m, const void* P) { return gw(m, P); }

// FWN_TRAIN_STREAMS=1 keeps the whole backward pass on the caller's stream (diagnostics / A-B timing)
static bool dual_stream() {   // read per call: tests flip it between two passes of one process
  const char* e = getenv("FWN_TRAIN_STREAMS");
  return !(e && e[0] == '1');
}
bool train_dual_stream() { return dual_stream(); }

// Block `block` is done: packed-operand gradients -> folded vector -> raw variables for THIS block, behind its weight gradients on
// the side stream (the dgrad chain of the next block does not wait for it); then the block's bucket of the flat gradient is final.
int train_finish_block(Model* m, int block, float* grads, cudaStream_t st) {
  TrainState* t = m->train;
  TrainState::Bucket& bk = t->buckets[(size_t)(m->cfg.n_block - 1 - block)];
  cudaStream_t s1 = dual_stream() ? t->side : st;
  if (dual_stream()) {   // the ActNorm gradients of the block were written on the main stream
    cudaEvent_t e = t->next_event();
    FWN_CUDA(cudaEventRecord(e, st));
    FWN_CUDA(cudaStreamWaitEvent(s1, e, 0));
  }
  m->launches += 2;
  if (scatter_grad(t->gwall + bk.wall0, t->wmap + bk.wall0, grads, bk.wall1 - bk.wall0, s1)) return 1;
  if (fold_backward(m->raw, grads, t->d_folds, t->d_fwork + bk.work0, bk.work1 - bk.work0, m->raw_floats, s1)) return 1;
  FWN_CUDA(cudaEventRecordWithFlags(bk.ready, s1, t->capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
  return 0;
}

// Gradient of the two transposed-conv stages (model.py:398-404) from the accumulated conditioning gradient; closes the flat gradient
// (records the events of the buckets behind the blocks).  c planes / up0 are the fp32 forward outputs (leaky-relu masks).
int train_upsampler_backward(Model* m, const float* cmel, const float* up0, float* dup0, const float* cA, const float* cB, const float* dcA,
                             const float* dcB, int B, int T, float* grads, cudaStream_t st) {
  const fwn_config& c = m->cfg;
  TrainState* t = m->train;
  int Tm = T / m->hop;
  const int last = c.n_upsample - 1;
  auto poff = [&](const std::string& n) { return m->params[m->index.at(n)].offset; };
  auto pname = [&](int i) { return i == 0 ? std::string("conv2d_transpose") : "conv2d_transpose_" + std::to_string(i); };
  for (int i = last; i >= 0; --i) {
    const int s = c.upsample_scales[i];
    int Tin = Tm;
    for (int k = 0; k < i; ++k) Tin *= c.upsample_scales[k];
    const float* in = i == 0 ? cmel : up0;
    const bool split = i == last;
    const float *d0 = split ? dcA : dup0, *d1 = split ? dcB : nullptr;
    const float *o0 = split ? cA : up0, *o1 = split ? cB : nullptr;
    m->launches += 3;
    if (upsample_bwd_stage(d0, d1, o0, o1, split, in, m->up_w[i], t->up_dw, i > 0 ? dup0 : nullptr, B, Tin, c.num_mels, s, st)) return 1;
    const std::string n = pname(i);
    if (upsample_wn_bwd(m->raw + poff(n + "/kernel"), m->raw + poff(n + "/wn/g"), t->up_dw, s, grads + poff(n + "/kernel"),
                        grads + poff(n + "/wn/g"), grads + poff(n + "/bias"), st))
      return 1;
  }
  // the upsampler variables (and the speaker embeddings, whose gradient is identically zero: SURVEY F6) close the gradient
  for (size_t k = (size_t)c.n_block; k < t->buckets.size(); ++k)
    FWN_CUDA(cudaEventRecordWithFlags(t->buckets[k].ready, st, t->capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
  return 0;
}

static int flow_backward(Model* m, const TrainWs& w, const FlowPack& fp, const TrainFlow& tf, const Tape& tp, const float* Xpost, int B,
                         int Ti, float* G, int parity, cudaStream_t st) {
  TrainState* t = m->train;
  const bool dual = dual_stream();
  cudaStream_t s1 = dual ? t->side : st;   // weight gradients + conditioning gradient
  const TrainWs::BwdSet& bs = w.set[parity];
  // everything issued so far on the main stream becomes visible to the side stream
  auto fork = [&]() -> int {
    if (!dual) return 0;
    cudaEvent_t e = t->next_event();
    FWN_CUDA(cudaEventRecord(e, st));
    FWN_CUDA(cudaStreamWaitEvent(s1, e, 0));
    return 0;
  };
  // this flow reuses the scratch set of the flow two steps back: its side-stream readers must be done
  if (dual) FWN_CUDA(cudaStreamWaitEvent(st, t->set_done[parity], 0));
  const fwn_config& c = m->cfg;
  const int F = c.filter_size, L = c.n_layer, nq = fp.nq, nq4 = ceil4(nq), ldn = ceil4(2 * nq);
  const int64_t rows = (int64_t)B * Ti;
  const double n_total = (double)rows * fp.Cx;
  auto linear = [&](GemmArgs& g, float* out, int ld, const float* add, const float* mask, float alpha) {
    g.B = B; g.Ti = Ti;
    g.e.out0 = out; g.e.ld = ld; g.e.in0 = add; g.e.in1 = mask; g.e.alpha = alpha; g.e.F = F;
  };
  auto wg = [&](const Seg* segs, int nseg, const float* y0, int64_t ld0, int n0, const float* y1, int64_t ld1, int N, const void* P,
                int64_t ldw, const float* bias_slot) {
    WgradArgs a = {};
    for (int i = 0; i < nseg; ++i) a.seg[i] = segs[i];
    a.nseg = nseg;
    a.dY0 = y0; a.ld0 = ld0; a.n0cols = n0; a.dY1 = y1; a.ld1 = ld1; a.N = N;
    a.dW = gw(m, P); a.ldw = ldw; a.B = B; a.Ti = Ti;
    if (wgrad_on_tensor_cores() && wgrad_tc3_supported(a)) {   // bias gradient (column sums of dY) fused into the same kernel
      m->launches += 1;
      return wgrad_tc3(a, gw(m, bias_slot), m->cur_terms, s1);
    }
    m->launches += 2;
    if (wgrad(a, s1)) return 1;
    return colsum(y0, ld0, n0, y1, ld1, N, rows, gw(m, bias_slot), s1);
  };

  // 1. coupling: d out_b -> (d log_s, d t), d b
  m->launches++;
  if (affine_bwd(w.dX, Xpost, tp.net, ldn, bs.dnet, rows, fp.Cx, nq, fp.b_off, n_total, st)) return 1;
  if (fork()) return 1;
  // 2. zero conv
  {
    Seg s0{tp.u, F, 0, F, 0};
    if (wg(&s0, 1, bs.dnet, ldn, 2 * nq, nullptr, 0, 2 * nq, fp.zero_w, 2 * nq, fp.zero_b)) return 1;
    GemmArgs g = {};
    g.seg[0] = Seg{bs.dnet, ldn, 0, 2 * nq, 0};
    g.nseg = 1; g.N = F;
    linear(g, bs.du, F, nullptr, tp.u, 1.f);   // through relu(final(..))
    if (bw_gemm(m, g, EPI_LINEAR, tf.zero_T, st)) return 1;
    if (fork()) return 1;
  }
  // 3. final conv
  {
    Seg s0{tp.s, F, 0, F, 0};
    if (wg(&s0, 1, bs.du, F, F, nullptr, 0, F, fp.final_w, F, fp.final_b)) return 1;
    GemmArgs g = {};
    g.seg[0] = Seg{bs.du, F, 0, F, 0};
    g.nseg = 1; g.N = F;
    linear(g, bs.ds, F, nullptr, tp.s, 1.f);   // through relu(sum of skips): the gradient of every layer's skip output
    if (bw_gemm(m, g, EPI_LINEAR, tf.final_T, st)) return 1;
  }
  // 4. residual layers, last to first.  r = gradient of the layer's residual-conv output = sqrt(.5) * d h_{n+1}
  const float* cond = fp.cond_half == 0 ? w.cA : w.cB;
  float* dcond = fp.cond_half == 0 ? w.dcA : w.dcB;
  int d = 1;
  for (int n = 1; n < L; ++n) d *= 3;
  for (int n = L - 1; n >= 0; --n, d /= 3) {
    const bool last = n == L - 1;
    const float* r = last ? nullptr : bs.r[n + 1];   // written by the previous iteration (layer n+1)
    float* rnext = bs.r[n];
    float* dfg = bs.dfg[n];
    if (fork()) return 1;                             // ds (and r) are ready for the res|skip weight gradient
    {
      Seg s0{tp.o[n], F, 0, F, 0};
      if (last) { if (wg(&s0, 1, bs.ds, F, F, nullptr, 0, F, fp.rs_w[n], F, fp.rs_b[n])) return 1; }
      else if (wg(&s0, 1, r, F, F, bs.ds, F, 2 * F, fp.rs_w[n], 2 * F, fp.rs_b[n])) return 1;
      GemmArgs g = {};
      if (last) { g.seg[0] = Seg{bs.ds, F, 0, F, 0}; g.nseg = 1; }
      else { g.seg[0] = Seg{r, F, 0, F, 0}; g.seg[1] = Seg{bs.ds, F, 0, F, F}; g.nseg = 2; }
      g.N = F; g.B = B; g.Ti = Ti;
      g.e.in0 = tp.fg[n]; g.e.out0 = dfg; g.e.F = F;
      if (bw_gemm(m, g, EPI_GATE_BWD, tf.rs_T[n], st)) return 1;
      if (fork()) return 1;
    }
    {
      Seg sg[4];
      for (int k = 0; k < 3; ++k) sg[k] = Seg{tp.h[n], F, shift_of(c, k, d), F, k * F};
      sg[3] = Seg{cond, fp.Kc, 0, fp.Kc, 3 * F};
      if (wg(sg, 4, dfg, 2 * F, 2 * F, nullptr, 0, 2 * F, fp.gate_w[n], 2 * F, fp.gate_b[n])) return 1;
      GemmArgs gc = {};
      gc.seg[0] = Seg{dfg, 2 * F, 0, 2 * F, 0};
      gc.nseg = 1; gc.N = fp.Kc;
      linear(gc, dcond, fp.Kc, dcond, nullptr, 1.f);   // accumulate the conditioning gradient in place
      if (bw_gemm(m, gc, EPI_LINEAR, tf.cond_T[n], s1)) return 1;
      GemmArgs gh = {};
      for (int k = 0; k < 3; ++k) gh.seg[k] = Seg{dfg, 2 * F, -shift_of(c, k, d), 2 * F, k * 2 * F};
      gh.nseg = 3; gh.N = F;
      linear(gh, rnext, F, last ? nullptr : r, n == 0 ? tp.h[0] : nullptr, n > 0 ? 0.70710678118654752440f : 1.f);
      if (bw_gemm(m, gh, EPI_LINEAR, tf.gate_T[n], st)) return 1;
    }
  }
  const float* dh0 = bs.r[0];   // gradient of the front conv's pre-activation
  if (fork()) return 1;
  // 5. front conv
  {
    Seg sg[3];
    for (int k = 0; k < 3; ++k) sg[k] = Seg{tp.a0, nq4, shift_of(c, k, 1), nq, k * nq};
    if (wg(sg, 3, dh0, F, F, nullptr, 0, F, fp.front_w, F, fp.front_b)) return 1;
    GemmArgs g = {};
    for (int k = 0; k < 3; ++k) g.seg[k] = Seg{dh0, F, -shift_of(c, k, 1), F, k * F};
    g.nseg = 3; g.N = nq;
    if (nq4 != nq) FWN_CUDA(cudaMemsetAsync(bs.da0, 0, (size_t)rows * nq4 * 4, st));
    linear(g, bs.da0, nq4, nullptr, nullptr, 1.f);
    if (bw_gemm(m, g, EPI_LINEAR, tf.front_T, st)) return 1;
  }
  // 6. ActNorm (+ the WaveNet-input gradient on the pass-through half)
  m->launches++;
  if (actnorm_bwd(w.dX, bs.da0, nq4, tp.xpre, fp.an_b, fp.an_s, fp.off2log, rows, fp.Cx, nq, G + (fp.raw_b - m->raw),
                     G + (fp.raw_logs - m->raw), st)) return 1;
  if (dual) FWN_CUDA(cudaEventRecord(t->set_done[parity], s1));   // side-stream readers of this scratch set
  return 0;
}

int train_loss_and_grads(Model* m, const float* x, const float* cmel, const int32_t* gspk, int B, int T, float* logp_out, float* logdet_out,
                         float* grads, int64_t grad_floats, void* ws, int64_t ws_bytes, cudaStream_t st) {
  FWN_CHECK(m && m->packed && m->train, "training not enabled: call fwn_train_enable after the last fwn_set_param");
  FWN_CHECK(x && cmel && grads, "null pointer");
  FWN_CHECK(!(m->cfg.gin_channels > 0 && gspk == nullptr), "g is None");
  FWN_CHECK(grad_floats >= train_grad_floats(m), "gradient buffer too small: need %lld floats", (long long)train_grad_floats(m));
  if (m->train_bf16) return train16_loss_and_grads(m, x, cmel, gspk, B, T, logp_out, logdet_out, grads, grad_floats, ws, ws_bytes, st);
  if (train_ensure_full_planes(m, st)) return 1;
  TrainWs w;
  if (train_plan(m, B, T, &w, (char*)ws)) return 1;
  FWN_CHECK(ws && ws_bytes >= (int64_t)w.bytes, "workspace too small: need %lld bytes, got %lld", (long long)w.bytes, (long long)ws_bytes);
  const fwn_config& c = m->cfg;
  TrainState* t = m->train;
  m->launches = 0;
  struct TermsGuard { Model* m; ~TermsGuard() { m->cur_terms = m->terms_infer; m->force_simt = false; g_exact_fwd = false; } } terms_guard{m};
  m->cur_terms = m->terms_train;
  m->force_simt = g_exact_fwd = m->train_exact_fwd;   // forward phase only; reset before the backward pass
  const size_t BT = (size_t)B * T;
  const int H = c.num_mels / 2;
  FWN_CUDA(cudaMemcpyAsync(w.X, x, BT * 4, cudaMemcpyDeviceToDevice, st));
  FWN_CUDA(cudaMemsetAsync(w.sums, 0, 8 * sizeof(double), st));
  FWN_CUDA(cudaMemsetAsync(grads, 0, (size_t)train_grad_floats(m) * 4, st));
  FWN_CUDA(cudaMemsetAsync(t->gwall, 0, (size_t)m->wall_floats * 4, st));
  FWN_CUDA(cudaMemsetAsync(w.dcA, 0, BT * H * 4, st));
  FWN_CUDA(cudaMemsetAsync(w.dcB, 0, BT * H * 4, st));
  // ---- forward
  Workspace iw = {};
  iw.up[0] = w.up0; iw.cA = w.cA; iw.cB = w.cB;
  if (run_upsample(m, iw, cmel, B, T, st)) return 1;
  {
    const bool dual = dual_stream();
    cudaStream_t s1 = dual ? t->side : st;
    if (dual) {   // the side stream sees the upsampled conditioning
      cudaEvent_t e = t->next_event();
      FWN_CUDA(cudaEventRecord(e, st));
      FWN_CUDA(cudaStreamWaitEvent(s1, e, 0));
    }
    for (int i = 0; i < c.n_block; ++i) {   // conditioning projections of the deep blocks, ahead of the chain
      if (!cond_ahead(T >> (i + 1), m->flows[(size_t)i * c.n_flow])) continue;
      for (int j = 0; j < c.n_flow; ++j) {
        const size_t f = (size_t)i * c.n_flow + j;
        if (cond_forward(m, w, m->flows[f], w.tape[f], B, T >> (i + 1), s1)) return 1;
      }
      if (dual) FWN_CUDA(cudaEventRecord(t->cond_ready[i], s1));
    }
    for (int i = 0; i < c.n_block; ++i) {
      if (dual && cond_ahead(T >> (i + 1), m->flows[(size_t)i * c.n_flow])) FWN_CUDA(cudaStreamWaitEvent(st, t->cond_ready[i], 0));
      for (int j = 0; j < c.n_flow; ++j) {
        const size_t f = (size_t)i * c.n_flow + j;
        if (flow_forward(m, w, m->flows[f], w.tape[f], B, T >> (i + 1), st)) return 1;
      }
    }
  }
  if (sumsq(w.X, w.sums + 1, (int64_t)BT, st)) return 1;
  if (finish_forward(w.sums, m->d_an_logdet, logp_out, logdet_out, (double)BT, st)) return 1;
  m->force_simt = false;
  // ---- backward
  if (dual_stream()) {
    for (auto& e : t->set_done) FWN_CUDA(cudaEventRecord(e, st));   // nothing pending on either scratch set; also orders the memsets above
    cudaEvent_t e = t->next_event();
    FWN_CUDA(cudaEventRecord(e, st));
    FWN_CUDA(cudaStreamWaitEvent(t->side, e, 0));
  }
  m->launches++;
  if (logp_bwd(w.X, w.dX, (int64_t)BT, st)) return 1;
  for (int i = c.n_block - 1; i >= 0; --i) {
    for (int j = c.n_flow - 1; j >= 0; --j) {
      const size_t f = (size_t)i * c.n_flow + j;
      const float* Xpost = f + 1 < m->flows.size() ? w.tape[f + 1].xpre : w.X;
      if (flow_backward(m, w, m->flows[f], t->flows[f], w.tape[f], Xpost, B, T >> (i + 1), grads, (int)(f & 1), st)) return 1;
    }
    if (train_finish_block(m, i, grads, st)) return 1;
  }
  if (dual_stream()) {   // join: the conditioning gradient and all weight gradients are complete
    cudaEvent_t e = t->next_event();
    FWN_CUDA(cudaEventRecord(e, t->side));
    FWN_CUDA(cudaStreamWaitEvent(st, e, 0));
  }
  // ---- upsampler
  return train_upsampler_backward(m, cmel, w.up0, w.dup0, w.cA, w.cB, w.dcA, w.dcB, B, T, grads, st);
}

int train_bucket_count(const Model* m) { return m->train ? (int)m->train->buckets.size() : -1; }
int train_bucket_range(const Model* m, int k, int64_t* off, int64_t* count) {
  FWN_CHECK(m->train && k >= 0 && k < (int)m->train->buckets.size(), "bad gradient bucket %d", k);
  *off = m->train->buckets[(size_t)k].off;
  *count = m->train->buckets[(size_t)k].count;
  return 0;
}
int train_bucket_wait(const Model* m, int k, cudaStream_t consumer) {
  FWN_CHECK(m->train && k >= 0 && k < (int)m->train->buckets.size(), "bad gradient bucket %d", k);
  FWN_CUDA(cudaStreamWaitEvent(consumer, m->train->buckets[(size_t)k].ready, 0));
  return 0;
}

// Checkpoint / resume (train.py:190,199-210: tf.train.Saver over tf.global_variables() = variables + Adam slots + global_step).
// which: 0 = the flat variable vector, 1 = Adam first moments, 2 = Adam second moments; same flat layout for all three.
int train_state_copy(Model* m, int which, float* dst, const float* src, int64_t numel, cudaStream_t st) {
  FWN_CHECK(m && m->train, "training not enabled");
  FWN_CHECK(which >= 0 && which <= 2, "unknown state vector %d", which);
  FWN_CHECK(numel == m->raw_floats, "state vector has %lld floats, expected %lld", (long long)numel, (long long)m->raw_floats);
  float* mine = which == 0 ? m->raw : which == 1 ? m->train->adam_m : m->train->adam_v;
  if (dst) FWN_CUDA(cudaMemcpyAsync(dst, mine, (size_t)numel * 4, cudaMemcpyDeviceToDevice, st));
  if (src) {
    FWN_CUDA(cudaMemcpyAsync(mine, src, (size_t)numel * 4, cudaMemcpyDeviceToDevice, st));
    if (which == 0) return train_repack(m, st);   // every derived operand follows the variables
  }
  return 0;
}

int train_grad_norm(Model* m, const float* grads, float* norm_out, cudaStream_t st) {
  FWN_CHECK(m && m->train, "training not enabled");
  return grad_global_norm(grads, m->raw_floats, m->train->scratch, norm_out, st);
}

// clip_by_global_norm + Adam + device re-pack (train.py:76-81).  `grads` already averaged over towers by the caller.
int train_apply(Model* m, const float* grads, float lr, float beta1, float beta2, float eps, float clip_norm, int64_t step, cudaStream_t st) {
  FWN_CHECK(m && m->train, "training not enabled");
  FWN_CHECK(step >= 1, "Adam step counter starts at 1");
  TrainState* t = m->train;
  m->launches += 3;
  if (grad_global_norm(grads, m->raw_floats, t->scratch, t->norm, st)) return 1;
  if (adam_update(m->raw, t->adam_m, t->adam_v, grads, t->norm, clip_norm, lr, beta1, beta2, eps, step, m->raw_floats, st)) return 1;
  return train_repack(m, st, m->train_bf16);   // the bf16 step reads plane 0 of the operand planes only
}

}  // namespace fwn

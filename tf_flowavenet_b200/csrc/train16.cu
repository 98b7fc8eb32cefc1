// Training step in the 16-bit compute mode (BASELINE config C5: "bf16"): the reference's mixed-precision training (utils.py:3-31:
// fp32 master variables, low-precision compute copies; train.py:53-81) on the tcgen05 engine.
//
//   variables, Adam moments, gradients, log-det / log_p reductions, the flow variable x and its gradient   fp32
//   GEMM operands: weights (plane 0 of the split engine's operand planes = bf16(W), refreshed by the device re-pack after every
//   optimizer step), WaveNet activations and their gradients (the tape)                                       bf16
//   accumulation (TMEM) and every epilogue                                                                   fp32
//
// Forward = the inference path's kernels (gemm_tc.cu) writing a tape: per flow x before the flow, the front conv's A operand, and
// per layer h_n, o_n, sigmoid(g_n); sum of skips, final activation, (log_s, t).  Backward: dgrads = the same implicit-GEMM kernel
// with transposed weight planes and negated time shifts (EPI_LINEAR / PLAIN / PLAIN_F32 epilogues), wgrads = wgrad_tc.cu (MN-major
// tcgen05 operands straight from the [B, T_i, C] tape), elementwise pieces in train_kernels16.cu.  bf16 has fp32's exponent range,
// so the reference's static loss scale (hparams.scale = 64 for fp16, train.py:62,75-77) is not needed: scale = 1.
// Streams, scratch double-buffering, gradient buckets, fold / scatter / unfold and the optimizer are shared with train.cu.
#include <algorithm>
#include <cstring>

#include "common.cuh"
#include "train.h"

namespace fwn {

namespace {

struct Tape16 {   // saved activations of one flow (16-bit unless noted)
  float* xpre;    // fp32 x before the flow
  void* a0;       // [rows, ceil8(nq)] ActNorm'd pass-through half (front conv operand)
  void *h[MAX_LAYERS], *o[MAX_LAYERS], *sg[MAX_LAYERS];   // layer input, gated output, sigmoid(g)   [rows, F]
  void *s, *u;    // relu(sum of skips), relu(final conv)                                            [rows, F]
  float* net;     // fp32 (log_s, t) [rows, ceil4(2 nq)]
  float* pc[MAX_LAYERS];   // deep blocks: fp32 [rows, 2F] conditioning projection of each layer, computed ahead on the side stream
};
struct Ws16 {
  double* sums;
  float *X, *dX, *cmel, *up0, *dup0, *cA32, *cB32, *dcA, *dcB;   // cmel: the step's own copy of the mel input (fixed address)
  void *cA, *cB;   // bf16 conditioning planes (GEMM operands)
  struct BwdSet { void* dnet; float* da0; void *du, *ds, *dobuf; void* dfg[MAX_LAYERS]; void* r[MAX_LAYERS]; } set[2];   // alternate per flow
  std::vector<Tape16> tape;
  size_t bytes;
};
inline size_t al256(size_t x) { return (x + 255) & ~size_t(255); }
inline int ceil4(int x) { return (x + 3) & ~3; }
inline int ceil8(int x) { return (x + 7) & ~7; }
inline int shift_of(const fwn_config& c, int k, int d) { return c.causal ? (k - 2) * d : (k - 1) * d; }

int plan16(const Model* m, int B, int T, Ws16* w, char* base) {
  const fwn_config& c = m->cfg;
  FWN_CHECK(B > 0 && T > 0, "empty input: B=%d T=%d", B, T);
  FWN_CHECK(T % m->hop == 0, "T=%d is not a multiple of the hop size %d", T, m->hop);
  FWN_CHECK(T % (1 << c.n_block) == 0, "T=%d is not a multiple of 2^n_block=%d", T, 1 << c.n_block);
  const int F = c.filter_size, H = c.num_mels / 2, L = c.n_layer;
  const size_t BT = (size_t)B * T, M0 = BT / 2;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = al256(off + bytes);
    return base ? base + o : (char*)nullptr;
  };
  w->sums = reinterpret_cast<double*>(take(16 * 8));
  w->X = (float*)take(BT * 4);
  w->dX = (float*)take(BT * 4);
  w->cmel = (float*)take((size_t)B * (T / m->hop) * c.num_mels * 4);
  const int s_last = c.upsample_scales[c.n_upsample - 1];
  const size_t up_elems = c.n_upsample > 1 ? (size_t)B * (T / s_last) * c.num_mels : 0;
  w->up0 = (float*)take(up_elems * 4);
  w->dup0 = (float*)take(up_elems * 4);
  w->cA32 = (float*)take(BT * H * 4); w->cB32 = (float*)take(BT * H * 4);
  w->dcA = (float*)take(BT * H * 4); w->dcB = (float*)take(BT * H * 4);
  w->cA = take(BT * H * 2); w->cB = take(BT * H * 2);
  for (auto& bs : w->set) {
    bs.dnet = take(4 * BT * 2);     // rows_i * ceil8(2 nq_i) <= 4 B T
    bs.da0 = (float*)take(2 * BT * 4);   // rows_i * ceil4(nq_i) <= 2 B T
    bs.du = take(M0 * F * 2); bs.ds = take(M0 * F * 2); bs.dobuf = take(M0 * F * 2);
    for (int n = 0; n < L; ++n) { bs.dfg[n] = take(M0 * 2 * F * 2); bs.r[n] = take(M0 * F * 2); }
  }
  w->tape.assign(m->flows.size(), Tape16{});
  for (int i = 0; i < c.n_block; ++i) {
    const size_t M = BT >> (i + 1);
    const int nq = 1 << i;
    for (int j = 0; j < c.n_flow; ++j) {
      Tape16& tp = w->tape[(size_t)i * c.n_flow + j];
      tp.xpre = (float*)take(BT * 4);
      tp.a0 = take(M * ceil8(nq) * 2);
      for (int n = 0; n < L; ++n) { tp.h[n] = take(M * F * 2); tp.o[n] = take(M * F * 2); tp.sg[n] = take(M * F * 2); }
      tp.s = take(M * F * 2); tp.u = take(M * F * 2);
      tp.net = (float*)take(M * ceil4(2 * nq) * 4);
      for (int n = 0; n < L; ++n) tp.pc[n] = H * (2 << i) >= AHEAD_MIN_KC ? (float*)take(M * 2 * F * 4) : nullptr;
    }
  }
  w->bytes = off;
  return 0;
}

// one 16-bit GEMM of the step; W = plane 0 of a split-engine operand (bf16 [Npad][Kpad])
int gemm16(Model* m, const GemmArgs& g, EpiKind kind, const W3& w, cudaStream_t st) {
  m->launches++;
  FWN_CHECK(w.p, "internal: missing operand planes");
  return tc_gemm16(g, kind, w.p, w.Kpad, w.Npad, false, st);
}

int flow_forward16(Model* m, const Ws16& w, const FlowPack& fp, const Tape16& tp, int B, int Ti, cudaStream_t st) {
  const fwn_config& c = m->cfg;
  const int F = c.filter_size, L = c.n_layer, nq = fp.nq, kq = ceil8(nq);
  const int64_t rows = (int64_t)B * Ti;
  FWN_CUDA(cudaMemcpyAsync(tp.xpre, w.X, (size_t)rows * fp.Cx * 4, cudaMemcpyDeviceToDevice, st));
  m->launches++;
  if (front_pack(w.X, fp.Cx, nq, kq, fp.off2log, fp.an_b, fp.an_s, tp.a0, rows, false, st)) return 1;
  {
    GemmArgs g = {};
    g.B = B; g.Ti = Ti;
    for (int k = 0; k < 3; ++k) g.seg[k] = Seg{tp.a0, kq, shift_of(c, k, 1), nq, k * fp.front_k16};
    g.nseg = 3; g.N = F;
    g.e.bias = fp.front_b; g.e.out0 = tp.h[0]; g.e.relu = 1; g.e.F = F;
    if (gemm16(m, g, EPI_PLAIN, fp.w3[GEMM_FRONT], st)) return 1;
  }
  const void* cond = fp.cond_half == 0 ? w.cA : w.cB;
  int d = 1;
  for (int n = 0; n < L; ++n, d *= 3) {
    GemmArgs g = {};
    g.B = B; g.Ti = Ti;
    for (int k = 0; k < 3; ++k) g.seg[k] = Seg{tp.h[n], F, shift_of(c, k, d), F, k * F};
    g.seg[3] = Seg{cond, fp.Kc, 0, fp.Kc, 3 * F};
    g.nseg = tp.pc[n] ? 3 : 4; g.N = 2 * F;
    g.e.bias = fp.gate_b[n]; g.e.out0 = tp.o[n]; g.e.tape = tp.sg[n]; g.e.F = F;
    if (tp.pc[n]) { g.e.in0 = tp.pc[n]; g.e.ld = 2 * F; }   // the projection was computed ahead (cond_forward16)
    if (gemm16(m, g, EPI_GATE, fp.w3[GEMM_GATE0 + n], st)) return 1;
    const bool last = n == L - 1;
    GemmArgs r = {};
    r.B = B; r.Ti = Ti;
    r.seg[0] = Seg{tp.o[n], F, 0, F, 0};
    r.nseg = 1; r.N = last ? F : 2 * F;
    r.e.bias = fp.rs_b[n]; r.e.F = F; r.e.has_res = !last; r.e.relu = last;
    r.e.in0 = tp.h[n]; r.e.out0 = last ? nullptr : tp.h[n + 1]; r.e.in1 = n > 0 ? tp.s : nullptr; r.e.out1 = tp.s;
    if (gemm16(m, r, EPI_RES_SKIP, fp.w3[GEMM_RS0 + n], st)) return 1;
  }
  {
    GemmArgs g = {};
    g.B = B; g.Ti = Ti;
    g.seg[0] = Seg{tp.s, F, 0, F, 0};
    g.nseg = 1; g.N = F;
    g.e.bias = fp.final_b; g.e.out0 = tp.u; g.e.relu = 1; g.e.F = F;
    if (gemm16(m, g, EPI_PLAIN, fp.w3[GEMM_FINAL], st)) return 1;
  }
  {
    GemmArgs g = {};
    g.B = B; g.Ti = Ti;
    g.seg[0] = Seg{tp.u, F, 0, F, 0};
    g.nseg = 1; g.N = 2 * nq;
    g.e.bias = fp.zero_b; g.e.F = F;
    g.e.X = w.X; g.e.Cx = fp.Cx; g.e.nq = nq; g.e.a_off = fp.a_off; g.e.b_off = fp.b_off;
    g.e.an_b = fp.an_b; g.e.an_s = fp.an_s;
    g.e.logdet_acc = w.sums;
    g.e.reverse = 0;
    g.e.pairs_adjacent = fp.pairs_adjacent;
    g.e.b_odd = fp.b_odd;
    g.e.out1 = tp.net; g.e.ld = ceil4(2 * nq);
    if (gemm16(m, g, EPI_AFFINE, fp.w3[GEMM_ZERO], st)) return 1;
  }
  return 0;
}

// Deep blocks (K_c >= AHEAD_MIN_KC: thousands of conditioning channels against a few hundred rows): the projections c_a . W_c do not
// depend on the flow state, so they run ahead of the dependent chain -- on the side stream, over a flat row axis -- into fp32
// buffers the gate GEMMs add in their epilogue; the chain's gate GEMMs then reduce over the 768 conv inputs only.
int cond_forward16(Model* m, const Ws16& w, const FlowPack& fp, const Tape16& tp, int B, int Ti, cudaStream_t st) {
  const int F = m->cfg.filter_size, L = m->cfg.n_layer;
  const void* cond = fp.cond_half == 0 ? w.cA : w.cB;
  for (int n = 0; n < L; ++n) {
    if (!tp.pc[n]) continue;
    GemmArgs g = {};
    g.B = B; g.Ti = Ti;
    g.seg[0] = Seg{cond, fp.Kc, 0, fp.Kc, 3 * F};
    g.nseg = 1; g.N = 2 * F;
    g.e.out0 = tp.pc[n]; g.e.ld = 2 * F; g.e.F = F;
    if (gemm16(m, g, EPI_PLAIN_F32, fp.w3[GEMM_GATE0 + n], st)) return 1;
  }
  return 0;
}

int flow_backward16(Model* m, const Ws16& w, const FlowPack& fp, const TrainFlow& tf, const Tape16& tp, const float* Xpost, int B, int Ti,
                    float* G, int parity, cudaStream_t st) {
  TrainState* t = m->train;
  const bool dual = train_dual_stream();
  cudaStream_t s1 = dual ? t->side : st;   // weight gradients + conditioning gradient
  const Ws16::BwdSet& bs = w.set[parity];
  auto fork = [&]() -> int {   // everything issued so far on the main stream becomes visible to the side stream
    if (!dual) return 0;
    cudaEvent_t e = t->next_event();
    FWN_CUDA(cudaEventRecord(e, st));
    FWN_CUDA(cudaStreamWaitEvent(s1, e, 0));
    return 0;
  };
  // this flow reuses the scratch set of the flow two steps back: its side-stream readers must be done
  if (dual) FWN_CUDA(cudaStreamWaitEvent(st, t->set_done[parity], 0));
  const fwn_config& c = m->cfg;
  const int F = c.filter_size, L = c.n_layer, nq = fp.nq, kq = ceil8(nq), nq4 = ceil4(nq), ldn = ceil4(2 * nq), ldd = ceil8(2 * nq);
  const int64_t rows = (int64_t)B * Ti;
  const double n_total = (double)rows * fp.Cx;
  auto wg = [&](const Seg* segs, int nseg, const void* y0, int64_t ld0, int n0, const void* y1, int64_t ld1, int N, const void* P,
                int64_t ldw, const float* bias_slot) {
    Wgrad16Args a = {};
    for (int i = 0; i < nseg; ++i) a.seg[i] = segs[i];
    a.nseg = nseg;
    a.dY0 = y0; a.ld0 = ld0; a.n0cols = n0; a.dY1 = y1; a.ld1 = ld1; a.N = N;
    a.dW = train_gw(m, P); a.ldw = ldw; a.B = B; a.Ti = Ti;
    m->launches++;
    return wgrad_tc(a, train_gw(m, bias_slot), s1);   // bias gradient (column sums of dY) inside the same kernel
  };
  auto dgrad = [&](GemmArgs& g, EpiKind kind, const W3& wT, void* out, const void* in, int mask, float alpha, cudaStream_t stream) {
    g.B = B; g.Ti = Ti;
    g.e.out0 = out; g.e.in0 = in; g.e.mask_mode = mask; g.e.alpha = alpha; g.e.F = F;
    return gemm16(m, g, kind, wT, stream);
  };

  // 1. coupling: d out_b -> (d log_s, d t), d b
  m->launches++;
  if (affine_bwd16(w.dX, Xpost, tp.net, ldn, bs.dnet, ldd, rows, fp.Cx, nq, fp.b_off, n_total, st)) return 1;
  if (fork()) return 1;
  // 2. zero conv
  {
    Seg s0{tp.u, F, 0, F, 0};
    if (wg(&s0, 1, bs.dnet, ldd, 2 * nq, nullptr, 0, 2 * nq, fp.zero_w, 2 * nq, fp.zero_b)) return 1;
    GemmArgs g = {};
    g.seg[0] = Seg{bs.dnet, ldd, 0, 2 * nq, 0};
    g.nseg = 1; g.N = F;
    if (dgrad(g, EPI_LINEAR, tf.zero_T, bs.du, tp.u, 1, 1.f, st)) return 1;   // through relu(final(..))
    if (fork()) return 1;
  }
  // 3. final conv
  {
    Seg s0{tp.s, F, 0, F, 0};
    if (wg(&s0, 1, bs.du, F, F, nullptr, 0, F, fp.final_w, F, fp.final_b)) return 1;
    GemmArgs g = {};
    g.seg[0] = Seg{bs.du, F, 0, F, 0};
    g.nseg = 1; g.N = F;
    if (dgrad(g, EPI_LINEAR, tf.final_T, bs.ds, tp.s, 1, 1.f, st)) return 1;   // through relu(sum of skips): every layer's skip gradient
  }
  // 4. residual layers, last to first.  r = gradient of the layer's residual-conv output = sqrt(.5) * d h_{n+1}
  const void* cond = fp.cond_half == 0 ? w.cA : w.cB;
  float* dcond = fp.cond_half == 0 ? w.dcA : w.dcB;
  int d = 1;
  for (int n = 1; n < L; ++n) d *= 3;
  for (int n = L - 1; n >= 0; --n, d /= 3) {
    const bool last = n == L - 1;
    const void* r = last ? nullptr : bs.r[n + 1];
    void* rnext = bs.r[n];
    void* dfg = bs.dfg[n];
    if (fork()) return 1;
    {
      Seg s0{tp.o[n], F, 0, F, 0};
      if (last) { if (wg(&s0, 1, bs.ds, F, F, nullptr, 0, F, fp.rs_w[n], F, fp.rs_b[n])) return 1; }
      else if (wg(&s0, 1, r, F, F, bs.ds, F, 2 * F, fp.rs_w[n], 2 * F, fp.rs_b[n])) return 1;
      GemmArgs g = {};
      if (last) { g.seg[0] = Seg{bs.ds, F, 0, F, 0}; g.nseg = 1; }
      else { g.seg[0] = Seg{r, F, 0, F, 0}; g.seg[1] = Seg{bs.ds, F, 0, F, F}; g.nseg = 2; }
      g.N = F;
      if (dgrad(g, EPI_PLAIN, tf.rs_T[n], bs.dobuf, nullptr, 0, 1.f, st)) return 1;      // d o
      m->launches++;
      if (gate_bwd16(bs.dobuf, tp.o[n], tp.sg[n], dfg, rows * F, st)) return 1;          // -> (d f, d g)
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
      gc.e.ld = fp.Kc; gc.e.accum = 1;                                                   // accumulate the conditioning gradient in place
      if (dgrad(gc, EPI_PLAIN_F32, tf.cond_T[n], dcond, nullptr, 0, 1.f, s1)) return 1;
      GemmArgs gh = {};
      for (int k = 0; k < 3; ++k) gh.seg[k] = Seg{dfg, 2 * F, -shift_of(c, k, d), 2 * F, k * 2 * F};
      gh.nseg = 3; gh.N = F;
      if (dgrad(gh, EPI_LINEAR, tf.gate_T[n], rnext, last ? nullptr : r, 0, n > 0 ? 0.70710678118654752440f : 1.f, st)) return 1;
      if (n == 0) {   // through relu(front conv)
        m->launches++;
        if (relu_mask16(rnext, tp.h[0], rows * F, st)) return 1;
      }
    }
  }
  const void* dh0 = bs.r[0];   // gradient of the front conv's pre-activation
  if (fork()) return 1;
  // 5. front conv
  {
    Seg sg[3];
    for (int k = 0; k < 3; ++k) sg[k] = Seg{tp.a0, kq, shift_of(c, k, 1), nq, k * nq};
    if (wg(sg, 3, dh0, F, F, nullptr, 0, F, fp.front_w, F, fp.front_b)) return 1;
    GemmArgs g = {};
    for (int k = 0; k < 3; ++k) g.seg[k] = Seg{dh0, F, -shift_of(c, k, 1), F, k * F};
    g.nseg = 3; g.N = nq;
    if (nq4 != nq) FWN_CUDA(cudaMemsetAsync(bs.da0, 0, (size_t)rows * nq4 * 4, st));
    g.e.ld = nq4;
    if (dgrad(g, EPI_PLAIN_F32, tf.front_T, bs.da0, nullptr, 0, 1.f, st)) return 1;
  }
  // 6. ActNorm (+ the WaveNet-input gradient on the pass-through half)
  m->launches++;
  if (actnorm_bwd(w.dX, bs.da0, nq4, tp.xpre, fp.an_b, fp.an_s, fp.off2log, rows, fp.Cx, nq, G + (fp.raw_b - m->raw),
                  G + (fp.raw_logs - m->raw), st)) return 1;
  if (dual) FWN_CUDA(cudaEventRecord(t->set_done[parity], s1));   // side-stream readers of this scratch set
  return 0;
}

}  // namespace

int64_t train16_workspace_bytes(const Model* m, int B, int T) {
  Ws16 w;
  if (plan16(m, B, T, &w, nullptr)) return -1;
  return (int64_t)w.bytes;
}

// Everything of the step that depends only on buffers inside the workspace / the handle: capturable.
static int step_body(Model* m, const Ws16& w, int B, int T, float* logp_out, float* logdet_out, float* grads, cudaStream_t st) {
  const fwn_config& c = m->cfg;
  TrainState* t = m->train;
  const size_t BT = (size_t)B * T;
  const int H = c.num_mels / 2;
  const bool dual = train_dual_stream();
  const float* cmel = w.cmel;
  FWN_CUDA(cudaMemsetAsync(w.sums, 0, 8 * sizeof(double), st));
  FWN_CUDA(cudaMemsetAsync(grads, 0, (size_t)train_grad_floats(m) * 4, st));
  FWN_CUDA(cudaMemsetAsync(t->gwall, 0, (size_t)m->wall_floats * 4, st));
  FWN_CUDA(cudaMemsetAsync(w.dcA, 0, BT * H * 4, st));
  FWN_CUDA(cudaMemsetAsync(w.dcB, 0, BT * H * 4, st));
  // ---- forward: upsampler in fp32 (leaky-relu masks of its backward pass) and once more into the bf16 operand planes
  Workspace iw = {};
  iw.up[0] = w.up0; iw.cA = w.cA32; iw.cB = w.cB32;
  if (run_upsample(m, iw, cmel, B, T, st)) return 1;
  {
    int Tm = T / m->hop;
    for (int i = 0; i + 1 < c.n_upsample; ++i) Tm *= c.upsample_scales[i];
    const int last = c.n_upsample - 1;
    m->launches++;
    if (upsample_stage(last == 0 ? cmel : w.up0, m->up_w[last], m->up_b[last], w.cA, w.cB, B, Tm, c.num_mels, c.upsample_scales[last], true, 1, st))
      return 1;
  }
  {
    cudaStream_t s1 = dual ? t->side : st;
    if (dual) {   // the side stream sees the upsampled conditioning
      cudaEvent_t e = t->next_event();
      FWN_CUDA(cudaEventRecord(e, st));
      FWN_CUDA(cudaStreamWaitEvent(s1, e, 0));
    }
    for (int i = 0; i < c.n_block; ++i) {   // conditioning projections of the deep blocks, ahead of the chain
      if (!w.tape[(size_t)i * c.n_flow].pc[0]) continue;
      for (int j = 0; j < c.n_flow; ++j) {
        const size_t f = (size_t)i * c.n_flow + j;
        if (cond_forward16(m, w, m->flows[f], w.tape[f], B, T >> (i + 1), s1)) return 1;
      }
      if (dual) FWN_CUDA(cudaEventRecord(t->cond_ready[i], s1));
    }
  }
  for (int i = 0; i < c.n_block; ++i) {
    if (dual && w.tape[(size_t)i * c.n_flow].pc[0]) FWN_CUDA(cudaStreamWaitEvent(st, t->cond_ready[i], 0));
    for (int j = 0; j < c.n_flow; ++j) {
      const size_t f = (size_t)i * c.n_flow + j;
      if (flow_forward16(m, w, m->flows[f], w.tape[f], B, T >> (i + 1), st)) return 1;
    }
  }
  if (sumsq(w.X, w.sums + 1, (int64_t)BT, st)) return 1;
  if (finish_forward(w.sums, m->d_an_logdet, logp_out, logdet_out, (double)BT, st)) return 1;
  // ---- backward
  if (dual) {
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
      if (flow_backward16(m, w, m->flows[f], t->flows[f], w.tape[f], Xpost, B, T >> (i + 1), grads, (int)(f & 1), st)) return 1;
    }
    if (train_finish_block(m, i, grads, st)) return 1;
  }
  if (dual) {   // join: the conditioning gradient and all weight gradients are complete
    cudaEvent_t e = t->next_event();
    FWN_CUDA(cudaEventRecord(e, t->side));
    FWN_CUDA(cudaStreamWaitEvent(st, e, 0));
  }
  return train_upsampler_backward(m, cmel, w.up0, w.dup0, w.cA32, w.cB32, w.dcA, w.dcB, B, T, grads, st);
}

// FWN_TRAIN_GRAPH=0 launches the step eagerly every time (diagnostics / A-B timing)
static bool step_graph_enabled() {
  const char* e = getenv("FWN_TRAIN_GRAPH");
  return !(e && e[0] == '0');
}

int train16_loss_and_grads(Model* m, const float* x, const float* cmel, const int32_t* gspk, int B, int T, float* logp_out, float* logdet_out,
                           float* grads, int64_t grad_floats, void* ws, int64_t ws_bytes, cudaStream_t st) {
  const fwn_config& c = m->cfg;
  FWN_CHECK(c.filter_size == 256 && c.num_mels % 8 == 0, "bf16 training needs filter_size 256 and num_mels %% 8 == 0 (TMA strides / tile shape)");
  FWN_CHECK(c.n_upsample <= 2, "training supports at most two upsampling stages");
  Ws16 w;
  if (plan16(m, B, T, &w, (char*)ws)) return 1;
  FWN_CHECK(ws && ws_bytes >= (int64_t)w.bytes, "workspace too small: need %lld bytes, got %lld", (long long)w.bytes, (long long)ws_bytes);
  TrainState* t = m->train;
  m->launches = 0;
  // the caller's inputs go into the workspace first: everything after this touches fixed addresses only
  FWN_CUDA(cudaMemcpyAsync(w.X, x, (size_t)B * T * 4, cudaMemcpyDeviceToDevice, st));
  FWN_CUDA(cudaMemcpyAsync(w.cmel, cmel, (size_t)B * (T / m->hop) * c.num_mels * 4, cudaMemcpyDeviceToDevice, st));
  const int dual = train_dual_stream() ? 1 : 0;
  TrainState::StepGraph* g = nullptr;
  if (step_graph_enabled()) {
    for (auto& e : t->step_graphs)
      if (e.B == B && e.T == T && e.dual == dual && e.ws == ws && e.grads == grads && e.lp == logp_out && e.ld == logdet_out) g = &e;
    if (!g) {   // first step with these buffers: eager (kernel attributes, tensor-map cache); the next one is captured
      if (t->step_graphs.size() >= 8) {
        for (auto& e : t->step_graphs) if (e.exec) cudaGraphExecDestroy(e.exec);
        t->step_graphs.clear();
      }
      t->step_graphs.push_back(TrainState::StepGraph{B, T, dual, ws, grads, logp_out, logdet_out, 0, nullptr, 0});
      return step_body(m, w, B, T, logp_out, logdet_out, grads, st);
    }
    if (g->state == 0) {
      cudaGraph_t graph = nullptr;
      if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();   // e.g. the legacy default stream: stay eager on this stream
        g->state = -1;
        return step_body(m, w, B, T, logp_out, logdet_out, grads, st);
      }
      t->capturing = true;
      const int64_t before = m->launches;
      const int rc = step_body(m, w, B, T, logp_out, logdet_out, grads, st);
      t->capturing = false;
      cudaError_t e = cudaStreamEndCapture(st, &graph);
      if (rc || e != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        g->state = -1;
        if (rc) return 1;
        m->launches = before;
        return step_body(m, w, B, T, logp_out, logdet_out, grads, st);
      }
      g->launches = m->launches - before;
      m->launches = before;
      e = cudaGraphInstantiate(&g->exec, graph, 0);
      cudaGraphDestroy(graph);
      if (e != cudaSuccess) {
        cudaGetLastError();
        g->exec = nullptr;
        g->state = -1;
        return step_body(m, w, B, T, logp_out, logdet_out, grads, st);
      }
      g->state = 1;
    }
    if (g->state == 1) {
      FWN_CUDA(cudaGraphLaunch(g->exec, st));
      m->launches += g->launches;
      return 0;
    }
  }
  return step_body(m, w, B, T, logp_out, logdet_out, grads, st);
}

}  // namespace fwn

// Model handle, weight prepack, workspace plan and the two whole passes (forward log-likelihood,
// inverse synthesis) of FloWaveNet (model.py:282-404) on one GPU / one stream.
//
// Data layout in HBM:
//   X      fp32 [B, T]            the flow variable, kept in natural time order for the WHOLE pass.
//                                 Block i merely views it as [B*T/2^(i+1), 2^(i+1)]: the squeeze
//                                 (model.py:226-228) is a within-row permutation, and it together with
//                                 change_order (model.py:166-174) is absorbed into the weights at prepack.
//   cA,cB  act  [B, T, mels/2]    upsampled conditioning, split in its two mel halves (= the c_a / c_b of
//                                 AffineCoupling, model.py:125); block i views them as [B*T_i, K_c].
//   h,o,s,u act [B*T_i, 256]      WaveNet hidden state / gated / skip-sum / final activations.
// No squeeze, unsqueeze, split, concat or change_order kernel ever runs on the model path.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "model.h"
#include "train.h"

namespace fwn {

// ---------------------------------------------------------------- errors / misc
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

static float bf2f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static uint16_t f2h(float f) {   // IEEE half, round-to-nearest-even (host side of __float2half_rn)
  const __half h = __float2half_rn(f);
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}
static uint16_t f2bf(float f) {  // round-to-nearest-even, like __float2bfloat16_rn
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

// ---------------------------------------------------------------- parameter schema
// Names follow the reference's variable scopes (model.py:13,110,178,209,284; modules.py:8,41,65,139;
// convolutional.py:65-93) with Keras layer scopes in first-call order; verified against a
// run of the reference's own Python in tests/golden/make_golden.py.
static void add_param(Model* m, const std::string& name, std::vector<int64_t> shape) {
  ParamDesc d;
  d.name = name;
  d.shape = shape;
  d.numel = 1;
  for (auto s : shape) d.numel *= s;
  d.offset = m->raw_floats;
  m->raw_floats += (d.numel + 3) & ~int64_t(3);
  m->index[name] = (int)m->params.size();
  m->params.push_back(d);
}
static int64_t off_of(const Model* m, const std::string& n) { return m->params[m->index.at(n)].offset; }
static void add_conv(Model* m, const std::string& p, int k, int cin, int cout) {
  add_param(m, p + "/kernel", {k, cin, cout});
  add_param(m, p + "/wn/g", {cout});
  add_param(m, p + "/bias", {cout});
  m->folds.push_back(FoldDesc{FOLD_WN, off_of(m, p + "/kernel"), off_of(m, p + "/wn/g"), 0, 0, k * cin, cout});
}
static void add_sum2(Model* m, const std::string& a, const std::string& b, int n) {
  m->ext_of[a] = m->ext_floats;
  m->folds.push_back(FoldDesc{FOLD_SUM2, off_of(m, a), off_of(m, b), 0, m->ext_floats, 1, n});
  m->ext_floats += n;
}

int model_create(const fwn_config* cfg, Model** out) {
  FWN_CHECK(cfg && out, "fwn_create: null argument");
  FWN_CHECK(cfg->n_block >= 1 && cfg->n_block <= 12, "n_block=%d out of range", cfg->n_block);
  FWN_CHECK(cfg->n_flow >= 1 && cfg->n_layer >= 1 && cfg->n_layer <= 8, "bad n_flow/n_layer");
  FWN_CHECK(cfg->num_mels >= 2 && cfg->num_mels % 2 == 0, "num_mels=%d must be even (c is split in halves, model.py:125)", cfg->num_mels);
  FWN_CHECK(cfg->filter_size >= 16 && cfg->filter_size % 16 == 0, "filter_size=%d must be a multiple of 16", cfg->filter_size);
  FWN_CHECK(cfg->n_upsample >= 1 && cfg->n_upsample <= 4, "n_upsample out of range");
  FWN_CHECK(cfg->precision == FWN_FP32 || is_mixed(cfg->precision), "unknown precision %d", cfg->precision);
  if (is_mixed(cfg->precision))
    FWN_CHECK(cfg->num_mels % 8 == 0 && cfg->filter_size == 256, "mixed precision needs num_mels %% 8 == 0 and filter_size 256 (TMA strides / tile shape)");
  Model* m = new Model();
  m->cfg = *cfg;
  m->hop = 1;
  for (int i = 0; i < cfg->n_upsample; ++i) {
    FWN_CHECK(cfg->upsample_scales[i] >= 2 && cfg->upsample_scales[i] % 2 == 0, "upsample scale %d must be even", cfg->upsample_scales[i]);
    m->hop *= cfg->upsample_scales[i];
    std::string n = i == 0 ? "conv2d_transpose" : "conv2d_transpose_" + std::to_string(i);
    add_param(m, n + "/kernel", {2 * cfg->upsample_scales[i], 3, 1, 1});
    add_param(m, n + "/wn/g", {1});
    add_param(m, n + "/bias", {1});
  }
  const int F = cfg->filter_size;
  int cx = 1, cc = cfg->num_mels;
  for (int i = 0; i < cfg->n_block; ++i) {
    cx *= 2;
    cc *= 2;
    const int out_ch = cfg->affine ? cx : cx / 2;
    for (int j = 0; j < cfg->n_flow; ++j) {
      std::string pre = "Block_" + std::to_string(i) + "/Flow_" + std::to_string(j);
      add_param(m, pre + "/ActNorm/b", {1, 1, cx});
      add_param(m, pre + "/ActNorm/logs", {1, 1, cx});
      std::string w = pre + "/AffineCoupling/WaveNet";
      add_conv(m, w + "/Conv_front/conv1d", 3, cx / 2, F);
      for (int n = 0; n < cfg->n_layer; ++n) {
        std::string r = w + "/ResBlock_0_" + std::to_string(n);
        add_conv(m, r + "/Conv_filter/conv1d", 3, F, F);
        add_conv(m, r + "/Conv_gate/conv1d", 3, F, F);
        add_conv(m, r + "/conv1d", 1, cc / 2, F);    // _filter_conv_c (modules.py:117)
        add_conv(m, r + "/conv1d_1", 1, cc / 2, F);  // _gate_conv_c   (modules.py:118)
        add_conv(m, r + "/conv1d_2", 1, F, F);       // _res_conv      (modules.py:126)
        add_conv(m, r + "/conv1d_3", 1, F, F);       // _skip_conv     (modules.py:127)
        add_sum2(m, r + "/Conv_filter/conv1d/bias", r + "/conv1d/bias", F);
        add_sum2(m, r + "/Conv_gate/conv1d/bias", r + "/conv1d_1/bias", F);
      }
      add_conv(m, w + "/Conv_final/conv1d", 1, F, F);
      add_param(m, w + "/ZeroConv1d/conv1d/kernel", {1, F, out_ch});
      add_param(m, w + "/ZeroConv1d/conv1d/bias", {out_ch});
      add_param(m, w + "/ZeroConv1d/scale", {1, 1, out_ch});
      m->folds.push_back(FoldDesc{FOLD_ZERO, off_of(m, w + "/ZeroConv1d/conv1d/kernel"), off_of(m, w + "/ZeroConv1d/conv1d/bias"),
                                  off_of(m, w + "/ZeroConv1d/scale"), 0, F, out_ch});
    }
  }
  if (cfg->gin_channels > 0) add_param(m, "speaker_embeddings", {cfg->n_speakers, cfg->gin_channels});
  if (m->raw_floats + m->ext_floats >= (int64_t(1) << 30)) {
    set_error("model too large for the 30-bit gather map (%lld parameters)", (long long)m->raw_floats);
    delete m;
    return 1;
  }
  cudaError_t e = cudaMalloc(&m->raw, (size_t)m->raw_floats * sizeof(float));
  if (e != cudaSuccess) {
    set_error("cudaMalloc(%lld B) for parameters failed: %s", (long long)m->raw_floats * 4, cudaGetErrorString(e));
    delete m;
    return 1;
  }
  cudaMemset(m->raw, 0, (size_t)m->raw_floats * sizeof(float));
  *out = m;
  return 0;
}

void model_drop_graphs(Model* m) {
  for (auto& g : m->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  m->graphs.clear();
}

void model_destroy(Model* m) {
  if (!m) return;
  model_drop_graphs(m);
  train_free(m);
  for (auto e : m->prof_ev) cudaEventDestroy(e);
  cudaFree(m->raw);
  cudaFree(m->pack);
  cudaFree(m->host_ws);
  cudaFree(m->host_io);
  if (m->host_stream) cudaStreamDestroy(m->host_stream);
  delete m;
}

// ---------------------------------------------------------------- prepack (host side)
namespace {

struct Bump {  // bump allocator over a host staging buffer mirrored 1:1 on the device
  std::vector<char> buf;
  size_t alloc(size_t bytes) {
    size_t off = (buf.size() + 255) & ~size_t(255);
    buf.resize(off + bytes, 0);
    return off;
  }
};

// value + provenance of one packed element: `code` = index into the folded parameter vector | sign << 30, -1 = constant zero
struct Elem { double v; int32_t code; };
static inline Elem operator-(Elem e) { return Elem{-e.v, e.code < 0 ? e.code : (e.code ^ (1 << 30))}; }
struct Ref {  // a run of the folded parameter vector
  const double* w;
  int64_t base;
  Elem operator[](size_t i) const { return Elem{w[base + (int64_t)i], (int32_t)(base + (int64_t)i)}; }
};
struct PMat {  // a packed matrix / vector being assembled
  std::vector<double> v;
  std::vector<int32_t> code;
  explicit PMat(size_t n = 0) : v(n, 0.0), code(n, -1) {}
  struct Proxy {
    PMat* m; size_t i;
    void operator=(Elem e) { m->v[i] = e.v; m->code[i] = e.code; }
  };
  Proxy operator[](size_t i) { return Proxy{this, i}; }
  Elem at(size_t i) const { return Elem{v[i], code[i]}; }
  size_t size() const { return v.size(); }
};
static PMat pm_from(Ref r, size_t n) {
  PMat p(n);
  for (size_t i = 0; i < n; ++i) p[i] = r[i];
  return p;
}

struct HostParams {
  const Model* m;
  std::vector<float> raw;
  std::vector<double> what;  // folded parameter vector (host_fold)
  const float* p(const std::string& n) const {
    auto it = m->index.find(n);
    if (it == m->index.end()) abort();
    return raw.data() + m->params[it->second].offset;
  }
  Ref ref(const std::string& n) const { return Ref{what.data(), m->params[m->index.at(n)].offset}; }
  Ref ext(const std::string& first_bias) const { return Ref{what.data(), m->raw_floats + m->ext_of.at(first_bias)}; }
};

// The fold step on the host, in double precision (the device twin is fold_kernel in train_kernels.cu).
static void host_fold(HostParams& hp) {
  const Model* m = hp.m;
  hp.what.assign((size_t)(m->raw_floats + m->ext_floats), 0.0);
  for (int64_t i = 0; i < m->raw_floats; ++i) hp.what[i] = hp.raw[i];
  for (const FoldDesc& d : m->folds) {
    if (d.kind == FOLD_WN) {  // convolutional.py:80: v * rsqrt(max(sum_{k,i} v^2, 1e-12)) * g
      const float *v = hp.raw.data() + d.a, *g = hp.raw.data() + d.b;
      std::vector<double> ss(d.N, 0.0);
      for (int64_t r = 0; r < d.K; ++r)
        for (int o = 0; o < d.N; ++o) ss[o] += (double)v[r * d.N + o] * v[r * d.N + o];
      for (int o = 0; o < d.N; ++o) ss[o] = (double)g[o] / sqrt(std::max(ss[o], 1e-12));
      for (int64_t r = 0; r < d.K; ++r)
        for (int o = 0; o < d.N; ++o) hp.what[d.a + r * d.N + o] = (double)v[r * d.N + o] * ss[o];
    } else if (d.kind == FOLD_ZERO) {
      for (int o = 0; o < d.N; ++o) {
        const double e = exp(3.0 * (double)hp.raw[d.c + o]);
        for (int64_t r = 0; r < d.K; ++r) hp.what[d.a + r * d.N + o] = (double)hp.raw[d.a + r * d.N + o] * e;
        hp.what[d.b + o] = (double)hp.raw[d.b + o] * e;
      }
    } else {
      for (int j = 0; j < d.N; ++j) hp.what[m->raw_floats + d.d + j] = (double)hp.raw[d.a + j] + (double)hp.raw[d.b + j];
    }
  }
}
static Ref wn_kernel(const HostParams& hp, const std::string& pre, int, int, int) { return hp.ref(pre + "/kernel"); }

// fp32 region of the pack buffer: values + gather codes, float-indexed
struct WBump {
  std::vector<float> v;
  std::vector<int32_t> code;
  size_t alloc(size_t n) {
    size_t off = (v.size() + 63) & ~size_t(63);
    v.resize(off + n, 0.f);
    code.resize(off + n, -1);
    return off;
  }
};
constexpr size_t IN_WALL = size_t(1) << 62;  // tag on offsets that point into the fp32 region

// Store a [K][N] matrix either as fp32 [K][N] in the fp32 region (fp32 engines) or bf16 [Npad][Kpad] (tcgen05 engine, K-major B operand).
// `bf16`: 0 = fp32 operand, 1 = bf16, 2 = fp16 (the precision enum)
static size_t store_matrix(WBump& bw, Bump& b, const PMat& w, int K, int N, int bf16, int* ld_out) {
  if (!bf16) {
    size_t off = bw.alloc((size_t)K * N);
    for (size_t i = 0; i < (size_t)K * N; ++i) { bw.v[off + i] = (float)w.v[i]; bw.code[off + i] = w.code[i]; }
    *ld_out = N;
    return off | IN_WALL;
  }
  const int Kpad = (K + 63) / 64 * 64, Npad = (N + 15) / 16 * 16;
  size_t off = b.alloc((size_t)Kpad * Npad * 2);
  uint16_t* d = reinterpret_cast<uint16_t*>(b.buf.data() + off);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) d[(size_t)n * Kpad + k] = bf16 == 2 ? f2h((float)w.v[(size_t)k * N + n]) : f2bf((float)w.v[(size_t)k * N + n]);
  *ld_out = Kpad;
  return off;
}
// fp32 mode: additionally the three bf16 planes of the split engine, [3][Npad][Kpad] (K-major B operands).
struct W3Off { size_t off; int Kpad, Npad; bool set = false; };
static W3Off store_planes(Bump& b, const PMat& w, int K, int N) {
  W3Off o;
  o.Kpad = (K + 63) / 64 * 64;
  o.Npad = (N + 15) / 16 * 16;
  const size_t plane = (size_t)o.Kpad * o.Npad;
  o.off = b.alloc(3 * plane * 2);
  uint16_t* d = reinterpret_cast<uint16_t*>(b.buf.data() + o.off);
  memset(d, 0, 3 * plane * 2);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      const float x = (float)w.v[(size_t)k * N + n];   // the fp32 value the CUDA-core engine would use
      const uint16_t h1 = f2bf(x);
      const float r1 = x - bf2f(h1);
      const uint16_t h2 = f2bf(r1);
      const float r2 = r1 - bf2f(h2);
      const uint16_t h3 = f2bf(r2);
      const size_t i = (size_t)n * o.Kpad + k;
      d[i] = h1; d[plane + i] = h2; d[2 * plane + i] = h3;
    }
  o.set = true;
  return o;
}
// a (trainable-derived) float vector in the fp32 region
static size_t store_pvec(WBump& bw, const PMat& v, int pad_to = 0) {
  size_t n = std::max<size_t>(v.size(), (size_t)pad_to);
  size_t off = bw.alloc(n);
  for (size_t i = 0; i < v.size(); ++i) { bw.v[off + i] = (float)v.v[i]; bw.code[off + i] = v.code[i]; }
  return off | IN_WALL;
}
static size_t store_floats(Bump& b, const std::vector<double>& v, int pad_to = 0) {
  size_t n = std::max<size_t>(v.size(), (size_t)pad_to);
  size_t off = b.alloc(n * 4);
  float* d = reinterpret_cast<float*>(b.buf.data() + off);
  for (size_t i = 0; i < v.size(); ++i) d[i] = (float)v[i];
  return off;
}
static size_t store_ints(Bump& b, const std::vector<int>& v) {
  size_t off = b.alloc(v.size() * 4);
  memcpy(b.buf.data() + off, v.data(), v.size() * 4);
  return off;
}

struct CMap { int m, o; };

}  // namespace

int model_prepack(Model* m, cudaStream_t st) {
  const fwn_config& c = m->cfg;
  const int F = c.filter_size, L = c.n_layer, H = c.num_mels / 2;
  const int bf16 = is_mixed(c.precision) ? c.precision : 0;   // 0 fp32 operands, 1 bf16, 2 fp16
  HostParams hp;
  hp.m = m;
  hp.raw.resize((size_t)m->raw_floats);
  FWN_CUDA(cudaStreamSynchronize(st));
  FWN_CUDA(cudaMemcpy(hp.raw.data(), m->raw, (size_t)m->raw_floats * 4, cudaMemcpyDeviceToHost));

  host_fold(hp);
  Bump b;
  WBump bw;
  struct FlowOff {  // offsets into the staging buffer, turned into device pointers after upload
    size_t a_off, b_off, off2log, an_b, an_s, an_is, front_w, front_b, front_wtc, final_w, final_b, zero_w, zero_b;
    std::vector<size_t> gate_w, gate_b, rs_w, rs_b;
    W3Off w3[GEMM_IDS];
  };
  std::vector<FlowOff> offs;
  struct CondOff { int block, half; size_t off; int N, Kpad, Kc; };
  std::vector<CondOff> cond_offs;
  m->flows.clear();
  double an_logdet = 0.0;

  // ---- upsampler: fold weight norm over axes [0,2] of [2s,3,1,1] => per kw column (convolutional.py:186)
  std::vector<size_t> up_w, up_b;
  for (int i = 0; i < c.n_upsample; ++i) {
    const int s = c.upsample_scales[i];
    std::string n = i == 0 ? "conv2d_transpose" : "conv2d_transpose_" + std::to_string(i);
    const float* v = hp.p(n + "/kernel");
    const double g = hp.p(n + "/wn/g")[0];
    std::vector<double> w((size_t)2 * s * 3);
    for (int kw = 0; kw < 3; ++kw) {
      double ss = 0;
      for (int kh = 0; kh < 2 * s; ++kh) ss += (double)v[kh * 3 + kw] * v[kh * 3 + kw];
      const double sc = g / sqrt(std::max(ss, 1e-12));
      for (int kh = 0; kh < 2 * s; ++kh) w[kh * 3 + kw] = v[kh * 3 + kw] * sc;
    }
    up_w.push_back(store_floats(b, w));
    up_b.push_back(store_floats(b, {(double)hp.p(n + "/bias")[0]}));
  }

  // ---- permutation tracking: logical channel -> physical (time-ordered) offset
  std::vector<int> x_off = {0};
  std::vector<CMap> c_map(c.num_mels);
  for (int mm = 0; mm < c.num_mels; ++mm) c_map[mm] = {mm, 0};
  int cx = 1;
  for (int i = 0; i < c.n_block; ++i) {
    // squeeze (model.py:226-233): new channel 2c+k <- (row parity k, old channel c)
    std::vector<int> nx(2 * x_off.size());
    for (size_t ch = 0; ch < x_off.size(); ++ch)
      for (int k = 0; k < 2; ++k) nx[2 * ch + k] = k * cx + x_off[ch];
    std::vector<CMap> nc(2 * c_map.size());
    for (size_t ch = 0; ch < c_map.size(); ++ch)
      for (int k = 0; k < 2; ++k) nc[2 * ch + k] = {c_map[ch].m, k * cx + c_map[ch].o};
    x_off.swap(nx);
    c_map.swap(nc);
    cx *= 2;
    const int nq = cx / 2, Kc = H * cx;
    // conditioning-ahead operands of a deep block (mixed modes): one [slices * 2F][Kpad] matrix per mel half, filled flow by flow
    const bool ahead = bf16 != 0 && Kc >= AHEAD_MIN_KC;
    const int ca_kpad = (Kc + 63) / 64 * 64;
    size_t ca_off[2] = {0, 0};
    int ca_used[2] = {0, 0};
    if (ahead) {
      const int half0 = c_map[0].m / H;
      int cnt[2] = {0, 0};
      for (int j = 0; j < c.n_flow; ++j) cnt[(half0 + j) & 1]++;
      for (int h = 0; h < 2; ++h) {
        const size_t bytes = (size_t)cnt[h] * L * 2 * F * ca_kpad * 2;
        ca_off[h] = b.alloc(std::max<size_t>(bytes, 16));
        cond_offs.push_back(CondOff{i, h, ca_off[h], cnt[h] * L * 2 * F, ca_kpad, Kc});
      }
    }
    for (int j = 0; j < c.n_flow; ++j) {
      FlowPack fp;
      FlowOff fo;
      fp.Cx = cx;
      fp.nq = nq;
      fp.Kc = Kc;
      std::string pre = "Block_" + std::to_string(i) + "/Flow_" + std::to_string(j);
      std::string wpre = pre + "/AffineCoupling/WaveNet";
      // a / b halves.  Pair q couples pass-through channel q with transformed channel nq+q; every such pair is physically
      // adjacent (offsets 2p, 2p+1 in some order), so pairs are enumerated in PHYSICAL order p: the zero conv's column pairs
      // (and a_off / b_off) follow that order and a thread's consecutive columns touch consecutive bytes of its x row.
      std::vector<int> qperm(nq);
      for (int q = 0; q < nq; ++q) qperm[q] = q;
      std::sort(qperm.begin(), qperm.end(), [&](int u, int v) { return x_off[nq + u] < x_off[nq + v]; });
      std::vector<int> a_off(nq), b_off(nq);
      bool adjacent = true;
      for (int p2 = 0; p2 < nq; ++p2) {
        a_off[p2] = x_off[qperm[p2]];
        b_off[p2] = x_off[nq + qperm[p2]];
        adjacent = adjacent && ((a_off[p2] ^ b_off[p2]) == 1) && ((b_off[p2] >> 1) == p2);
      }
      fp.pairs_adjacent = adjacent ? 1 : 0;
      fp.b_odd = (b_off[0] & 1);
      std::vector<int> off2log(cx);
      for (int l = 0; l < cx; ++l) off2log[x_off[l]] = l;
      fo.a_off = store_ints(b, a_off);
      fo.b_off = store_ints(b, b_off);
      fo.off2log = store_ints(b, off2log);
      // ActNorm params in physical order
      const float* ab = hp.p(pre + "/ActNorm/b");
      const float* al = hp.p(pre + "/ActNorm/logs");
      std::vector<double> vb(cx), vs(cx), vis(cx);
      double ld = 0;
      for (int l = 0; l < cx; ++l) {
        vb[x_off[l]] = ab[l];
        vs[x_off[l]] = exp(3.0 * (double)al[l]);
        vis[x_off[l]] = exp(-3.0 * (double)al[l]);
        ld += 3.0 * (double)al[l];
      }
      an_logdet += ld / cx;
      fo.an_b = store_floats(b, vb);
      fo.an_s = store_floats(b, vs);
      fo.an_is = store_floats(b, vis);
      fp.raw_b = m->raw + m->params[m->index[pre + "/ActNorm/b"]].offset;
      fp.raw_logs = m->raw + m->params[m->index[pre + "/ActNorm/logs"]].offset;
      // conditioning half + K order
      const int half = c_map[0].m / H;
      std::vector<int> cpos(Kc);
      for (int l = 0; l < Kc; ++l) {
        FWN_CHECK(c_map[l].m / H == half, "internal: c_a is not a whole mel half");
        cpos[l] = c_map[l].o * H + (c_map[l].m % H);
      }
      fp.cond_half = half;
      fp.ahead = ahead ? 1 : 0;
      fp.ahead_slot = ca_used[half];
      if (ahead) ca_used[half] += L;
      // front conv [3][nq][F] (input channel q = logical channel q of x_a)
      {
        Ref w = wn_kernel(hp, wpre + "/Conv_front/conv1d", 3, nq, F);
        fo.front_w = store_pvec(bw, pm_from(w, (size_t)3 * nq * F));
        if (!bf16) {  // split-engine planes: tap k at K rows k*kq8 .. (TMA needs 16-byte aligned K starts)
          const int kq8 = (nq + 7) / 8 * 8;
          PMat w3((size_t)3 * kq8 * F);
          for (int k = 0; k < 3; ++k)
            for (int q = 0; q < nq; ++q)
              for (int ch = 0; ch < F; ++ch) w3[((size_t)k * kq8 + q) * F + ch] = w[((size_t)k * nq + q) * F + ch];
          fo.w3[GEMM_FRONT] = store_planes(b, w3, 3 * kq8, F);
          fp.front_k16 = kq8;
        }
        fo.front_wtc = 0;
        if (bf16) {  // tensor-core layout: [3*k16][F] with tap k at rows k*k16 .. k*k16+nq, then transposed to [F][Kpad]
          const int k16 = (nq + 15) / 16 * 16;
          PMat wt((size_t)3 * k16 * F);
          for (int k = 0; k < 3; ++k)
            for (int q = 0; q < nq; ++q)
              for (int ch = 0; ch < F; ++ch) wt[((size_t)k * k16 + q) * F + ch] = w[((size_t)k * nq + q) * F + ch];
          int ld;
          fo.front_wtc = store_matrix(bw, b, wt, 3 * k16, F, bf16, &ld);
          fp.front_ld = ld;
          fp.front_k16 = k16;
        }
        fo.front_b = store_pvec(bw, pm_from(hp.ref(wpre + "/Conv_front/conv1d/bias"), F));
      }
      for (int n = 0; n < L; ++n) {
        std::string r = wpre + "/ResBlock_0_" + std::to_string(n);
        Ref wf = wn_kernel(hp, r + "/Conv_filter/conv1d", 3, F, F), wg = wn_kernel(hp, r + "/Conv_gate/conv1d", 3, F, F);
        Ref wcf = wn_kernel(hp, r + "/conv1d", 1, Kc, F), wcg = wn_kernel(hp, r + "/conv1d_1", 1, Kc, F);
        const int Kc16 = (Kc + 15) / 16 * 16;
        const int Kg = ahead ? 3 * F : 3 * F + Kc16, Ng = 2 * F;   // ahead: the conditioning rows live in the block's cond_w instead
        PMat W((size_t)Kg * Ng), B(Ng);
        for (int k = 0; k < 3 * F; ++k)
          for (int ch = 0; ch < F; ++ch) {
            W[(size_t)k * Ng + 2 * ch] = wf[(size_t)k * F + ch];
            W[(size_t)k * Ng + 2 * ch + 1] = wg[(size_t)k * F + ch];
          }
        if (!ahead) {
          for (int l = 0; l < Kc; ++l)
            for (int ch = 0; ch < F; ++ch) {
              W[(size_t)(3 * F + cpos[l]) * Ng + 2 * ch] = wcf[(size_t)l * F + ch];
              W[(size_t)(3 * F + cpos[l]) * Ng + 2 * ch + 1] = wcg[(size_t)l * F + ch];
            }
        } else {   // rows (slot * 2F + column) of the block's projection operand, K = physical conditioning channel
          uint16_t* dst = reinterpret_cast<uint16_t*>(b.buf.data() + ca_off[half]) + (size_t)(fp.ahead_slot + n) * 2 * F * ca_kpad;
          for (int l = 0; l < Kc; ++l)
            for (int ch = 0; ch < F; ++ch) {
              const float vf = (float)wcf[(size_t)l * F + ch].v, vg = (float)wcg[(size_t)l * F + ch].v;
              dst[(size_t)(2 * ch) * ca_kpad + cpos[l]] = bf16 == 2 ? f2h(vf) : f2bf(vf);
              dst[(size_t)(2 * ch + 1) * ca_kpad + cpos[l]] = bf16 == 2 ? f2h(vg) : f2bf(vg);
            }
        }
        // bias of a gate column = conv bias + conditioning-conv bias: an `ext` slot of the folded vector (FOLD_SUM2)
        Ref bfs = hp.ext(r + "/Conv_filter/conv1d/bias"), bgs = hp.ext(r + "/Conv_gate/conv1d/bias");
        for (int ch = 0; ch < F; ++ch) {
          B[2 * ch] = bfs[ch];
          B[2 * ch + 1] = bgs[ch];
        }
        int ld;
        fo.gate_w.push_back(store_matrix(bw, b, W, Kg, Ng, bf16, &ld));
        if (!bf16) fo.w3[GEMM_GATE0 + n] = store_planes(b, W, Kg, Ng);
        fp.gate_ld = ld;
        fo.gate_b.push_back(store_pvec(bw, B));
        // res | skip 1x1 (the last layer's residual output is dead in the reference graph: modules.py:170-176)
        const bool last = n == L - 1;
        const int Nr = last ? F : 2 * F;
        Ref wr = wn_kernel(hp, r + "/conv1d_2", 1, F, F), ws = wn_kernel(hp, r + "/conv1d_3", 1, F, F);
        Ref br = hp.ref(r + "/conv1d_2/bias"), bs = hp.ref(r + "/conv1d_3/bias");
        PMat W2((size_t)F * Nr), B2(Nr);
        for (int k = 0; k < F; ++k)
          for (int ch = 0; ch < F; ++ch) {
            if (!last) {
              W2[(size_t)k * Nr + ch] = wr[(size_t)k * F + ch];
              W2[(size_t)k * Nr + F + ch] = ws[(size_t)k * F + ch];
            } else {
              W2[(size_t)k * Nr + ch] = ws[(size_t)k * F + ch];
            }
          }
        for (int ch = 0; ch < F; ++ch) {
          if (!last) { B2[ch] = br[ch]; B2[F + ch] = bs[ch]; }
          else B2[ch] = bs[ch];
        }
        fo.rs_w.push_back(store_matrix(bw, b, W2, F, Nr, bf16, &ld));
        if (!bf16) fo.w3[GEMM_RS0 + n] = store_planes(b, W2, F, Nr);
        fp.rs_ld[n] = ld;
        fo.rs_b.push_back(store_pvec(bw, B2));
      }
      {
        PMat w = pm_from(wn_kernel(hp, wpre + "/Conv_final/conv1d", 1, F, F), (size_t)F * F);
        int ld;
        fo.final_w = store_matrix(bw, b, w, F, F, bf16, &ld);
        if (!bf16) fo.w3[GEMM_FINAL] = store_planes(b, w, F, F);
        fp.final_ld = ld;
        fo.final_b = store_pvec(bw, pm_from(hp.ref(wpre + "/Conv_final/conv1d/bias"), F));
      }
      {
        // ZeroConv1d (modules.py:51-56): (u.W + b) * exp(3 scale); exp folded into W and b.
        // Columns (2q, 2q+1) = (log_s, t) of the q-th transformed channel.  Additive coupling
        // (model.py:136-139,157-159) is expressed as log_s = 0, t = -net.
        const int out_ch = c.affine ? cx : nq;
        Ref zk = hp.ref(wpre + "/ZeroConv1d/conv1d/kernel");   // exp(3 scale) already folded in (FOLD_ZERO)
        Ref zb = hp.ref(wpre + "/ZeroConv1d/conv1d/bias");
        const int Nz = 2 * nq;
        PMat W((size_t)F * Nz), B(Nz);
        for (int p2 = 0; p2 < nq; ++p2) {   // column pair p2 <- logical transformed channel q = qperm[p2]
          const int q = qperm[p2];
          if (c.affine) {
            for (int k = 0; k < F; ++k) {
              W[(size_t)k * Nz + 2 * p2] = zk[(size_t)k * out_ch + q];
              W[(size_t)k * Nz + 2 * p2 + 1] = zk[(size_t)k * out_ch + nq + q];
            }
            B[2 * p2] = zb[q];
            B[2 * p2 + 1] = zb[nq + q];
          } else {
            for (int k = 0; k < F; ++k) W[(size_t)k * Nz + 2 * p2 + 1] = -zk[(size_t)k * out_ch + q];
            B[2 * p2 + 1] = -zb[q];
          }
        }
        int ld;
        fo.zero_w = store_matrix(bw, b, W, F, Nz, bf16, &ld);
        if (!bf16) fo.w3[GEMM_ZERO] = store_planes(b, W, F, Nz);
        fp.zero_ld = ld;
        fo.zero_b = store_pvec(bw, B, (Nz + 15) / 16 * 16);
      }
      m->flows.push_back(fp);
      offs.push_back(fo);
      // change_order (model.py:190): swap halves of x and c
      std::rotate(x_off.begin(), x_off.begin() + nq, x_off.end());
      std::rotate(c_map.begin(), c_map.begin() + c_map.size() / 2, c_map.end());
    }
  }
  // The reverse pass (model.py:374-396) visits the same (x, c) channel orders iff n_flow is even.
  m->rev_ok = (c.n_flow % 2 == 0);

  // ---- upload + pointer fix-up
  size_t scal = b.alloc(8 * sizeof(double));
  reinterpret_cast<double*>(b.buf.data() + scal)[0] = an_logdet;
  if (m->pack) FWN_CUDA(cudaFree(m->pack));
  m->pack = nullptr;
  // device layout: [fp32 region (wall_floats) | everything else]
  const size_t wall_bytes = (bw.v.size() * 4 + 255) & ~size_t(255);
  FWN_CUDA(cudaMalloc(&m->pack, wall_bytes + b.buf.size()));
  m->pack_bytes = wall_bytes + b.buf.size();
  m->wall_floats = (int64_t)bw.v.size();
  FWN_CUDA(cudaMemcpy(m->pack, bw.v.data(), bw.v.size() * 4, cudaMemcpyHostToDevice));
  FWN_CUDA(cudaMemcpy(m->pack + wall_bytes, b.buf.data(), b.buf.size(), cudaMemcpyHostToDevice));
  if (m->keep_map) m->host_wmap = bw.code; else m->host_wmap.clear();
  struct BasePtr {  // resolves a tagged staging offset to its device address
    char* wall; char* rest;
    char* operator+(size_t off) const { return (off & IN_WALL) ? wall + (off & ~IN_WALL) * 4 : rest + off; }
  } base{m->pack, m->pack + wall_bytes};
  m->d_an_logdet = reinterpret_cast<double*>(base + scal);
  for (int i = 0; i < c.n_upsample; ++i) {
    m->up_w[i] = reinterpret_cast<float*>(base + up_w[i]);
    m->up_b[i] = reinterpret_cast<float*>(base + up_b[i]);
  }
  m->cond_w.assign((size_t)c.n_block * 2, Model::CondAhead());
  for (const CondOff& co : cond_offs) {
    Model::CondAhead& ca = m->cond_w[(size_t)co.block * 2 + co.half];
    ca.w = co.N > 0 ? (void*)(base + co.off) : nullptr;
    ca.N = co.N; ca.Kpad = co.Kpad; ca.Kc = co.Kc;
  }
  for (size_t f = 0; f < m->flows.size(); ++f) {
    FlowPack& fp = m->flows[f];
    const FlowOff& fo = offs[f];
    fp.a_off = reinterpret_cast<int*>(base + fo.a_off);
    fp.b_off = reinterpret_cast<int*>(base + fo.b_off);
    fp.off2log = reinterpret_cast<int*>(base + fo.off2log);
    fp.an_b = reinterpret_cast<float*>(base + fo.an_b);
    fp.an_s = reinterpret_cast<float*>(base + fo.an_s);
    fp.an_is = reinterpret_cast<float*>(base + fo.an_is);
    fp.front_w = reinterpret_cast<float*>(base + fo.front_w);
    fp.front_b = reinterpret_cast<float*>(base + fo.front_b);
    fp.front_wtc = bf16 ? base + fo.front_wtc : nullptr;
    fp.final_w = base + fo.final_w;
    fp.final_b = reinterpret_cast<float*>(base + fo.final_b);
    fp.zero_w = base + fo.zero_w;
    fp.zero_b = reinterpret_cast<float*>(base + fo.zero_b);
    for (int i = 0; i < GEMM_IDS; ++i) fp.w3[i] = W3{fo.w3[i].set ? (void*)(base + fo.w3[i].off) : nullptr, fo.w3[i].Kpad, fo.w3[i].Npad};
    for (int n = 0; n < L; ++n) {
      fp.gate_w[n] = base + fo.gate_w[n];
      fp.gate_b[n] = reinterpret_cast<float*>(base + fo.gate_b[n]);
      fp.rs_w[n] = base + fo.rs_w[n];
      fp.rs_b[n] = reinterpret_cast<float*>(base + fo.rs_b[n]);
    }
  }
  model_drop_graphs(m);
  m->packed = true;
  m->plan_B = m->plan_T = -1;  // tensor maps (tcgen05 engine) must be rebuilt
  // training enabled: rebuild the state that points into the new pack buffer, then produce the operands that exist only on the
  // device (the transposed dgrad planes) -- a re-prepack after fwn_train_enable must leave the handle ready for fwn_loss_and_grads
  if (train_after_prepack(m)) return 1;
  return m->keep_map ? train_repack(m, st) : 0;
}

// ---------------------------------------------------------------- optional CUDA-event profiling
void prof_begin(Model* m, int kind, double work, cudaStream_t st) {
  if (!m->prof_on) return;
  if (m->prof_used * 2 + 2 > m->prof_ev.size()) {
    for (int i = 0; i < 2; ++i) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      m->prof_ev.push_back(e);
    }
    m->prof_kind.push_back(0);
    m->prof_work.push_back(0);
  }
  m->prof_kind[m->prof_used] = kind;
  m->prof_work[m->prof_used] = work;
  cudaEventRecord(m->prof_ev[m->prof_used * 2], st);
}
void prof_end(Model* m, cudaStream_t st) {
  if (!m->prof_on) return;
  cudaEventRecord(m->prof_ev[m->prof_used * 2 + 1], st);
  m->prof_used++;
}
int prof_read(Model* m, double* ms, int64_t* launches, double* work) {
  for (int k = 0; k < PROF_KINDS; ++k) ms[k] = 0, launches[k] = 0, work[k] = 0;
  if (m->prof_used) FWN_CUDA(cudaEventSynchronize(m->prof_ev[m->prof_used * 2 - 1]));
  for (size_t i = 0; i < m->prof_used; ++i) {
    float t = 0;
    FWN_CUDA(cudaEventElapsedTime(&t, m->prof_ev[2 * i], m->prof_ev[2 * i + 1]));
    ms[m->prof_kind[i]] += t;
    launches[m->prof_kind[i]] += 1;
    work[m->prof_kind[i]] += m->prof_work[i];
  }
  m->prof_used = 0;
  return 0;
}

// ---------------------------------------------------------------- workspace plan
static inline size_t al256(size_t x) { return (x + 255) & ~size_t(255); }

int model_plan(const Model* m, int B, int T, Workspace* w, char* base) {
  const fwn_config& c = m->cfg;
  FWN_CHECK(B > 0 && T > 0, "empty input: B=%d T=%d", B, T);
  FWN_CHECK(T % m->hop == 0, "T=%d is not a multiple of the hop size %d (upsample_scales product; tfrecord.py:53)", T, m->hop);
  FWN_CHECK(T % (1 << c.n_block) == 0, "T=%d is not a multiple of 2^n_block=%d (squeeze, model.py:226)", T, 1 << c.n_block);
  const size_t as = is_mixed(c.precision) ? 2 : 4;
  const int F = c.filter_size, H = c.num_mels / 2;
  const size_t M0 = (size_t)B * T / 2;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = al256(off + bytes);
    return base ? base + o : (char*)nullptr;
  };
  w->sums = (double*)take(8 * sizeof(double));
  w->ddi = (double*)take(2 * 4096 * sizeof(double));
  w->x = (float*)take((size_t)B * T * 4);
  const int s_last = c.upsample_scales[c.n_upsample - 1];
  const size_t up_elems = c.n_upsample > 1 ? (size_t)B * (T / s_last) * c.num_mels : 0;
  w->up[0] = (float*)take(up_elems * 4);
  w->up[1] = (float*)take(c.n_upsample > 2 ? up_elems * 4 : 0);
  w->cA = take((size_t)B * T * H * as);
  w->cB = take((size_t)B * T * H * as);
  w->h0 = take(M0 * F * as);
  w->h1 = take(M0 * F * as);
  w->o = take(M0 * F * as);
  w->s = take(M0 * F * as);
  w->u = take(M0 * F * as);
  w->a0 = take((size_t)B * T * 8);  // mixed: rows_i * ceil8(nq_i) * 2 bytes; fp32: rows_i * ceil4(nq_i) * 4 bytes; both <= 8*B*T
  w->pc[0] = w->pc[1] = nullptr;
  if (is_mixed(c.precision)) {   // conditioning projections of the first (largest) deep block: rows_i x (flows of the half) x L x 2F, fp32
    for (int i = 0; i < c.n_block; ++i)
      if (H * (2 << i) >= AHEAD_MIN_KC) {
        const size_t elems = ((size_t)B * T >> (i + 1)) * (size_t)((c.n_flow + 1) / 2) * c.n_layer * 2 * F;
        w->pc[0] = (float*)take(elems * 4);
        w->pc[1] = (float*)take(elems * 4);
        break;
      }
  }
  w->bytes = off;
  return 0;
}

// ---------------------------------------------------------------- passes
__global__ void ddi_phys_finish_kernel(const double* acc, int64_t rows, int Cx, int phase, float* an_b, float* an_s, float* an_is,
                                       float* raw_b, float* raw_logs, const int* off2log, double* an_logdet) {
  int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= Cx) return;
  const int l = off2log[o];
  if (phase == 0) {
    float b = (float)(-acc[o] / (double)rows);
    an_b[o] = b;
    raw_b[l] = b;
  } else {
    double var = acc[Cx + o] / (double)rows;
    float logs = (float)(log(1.0 / (sqrt(var) + 1e-7)) / 3.0);   // model.py:69
    raw_logs[l] = logs;
    an_s[o] = (float)exp(3.0 * (double)logs);
    an_is[o] = (float)exp(-3.0 * (double)logs);
    atomicAdd(an_logdet, 3.0 * (double)logs / Cx);
  }
}
__global__ void colsum_phys_kernel(const float* __restrict__ x, const float* __restrict__ shift, double* __restrict__ acc, int64_t n, int Cx,
                                   bool square) {
  // Cx is a power of two <= 4096; each thread owns offset (tid % Cx) when blockDim % Cx == 0, else strides
  extern __shared__ double part[];
  const int nth = blockDim.x;
  if (Cx <= nth) {
    const int o = threadIdx.x % Cx;
    const float sh = shift ? shift[o] : 0.f;
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * nth + threadIdx.x; i < n; i += (int64_t)gridDim.x * nth) {
      float v = __ldg(x + i) + sh;
      s += square ? (double)v * v : (double)v;
    }
    part[threadIdx.x] = s;
    __syncthreads();
    if ((int)threadIdx.x < Cx) {
      double t = 0.0;
      for (int j = threadIdx.x; j < nth; j += Cx) t += part[j];
      atomicAdd(acc + threadIdx.x, t);
    }
  } else {
    for (int o = threadIdx.x; o < Cx; o += nth) {
      const float sh = shift ? shift[o] : 0.f;
      double s = 0.0;
      for (int64_t r = blockIdx.x; r * Cx < n; r += gridDim.x) {
        float v = __ldg(x + r * Cx + o) + sh;
        s += square ? (double)v * v : (double)v;
      }
      atomicAdd(acc + o, s);
    }
  }
}
static int ddi_flow(const Model* m, const FlowPack& fp, float* X, int64_t rows, double* scratch, cudaStream_t st) {
  const int Cx = fp.Cx;
  FWN_CHECK(Cx <= 4096, "DDI: Cx too large");
  const int64_t n = rows * Cx;
  FWN_CUDA(cudaMemsetAsync(scratch, 0, 2 * (size_t)Cx * sizeof(double), st));
  int grid = (int)std::min<int64_t>((int64_t)num_sms() * 4, std::max<int64_t>(1, n / 1024));
  colsum_phys_kernel<<<grid, 256, 256 * sizeof(double), st>>>(X, nullptr, scratch, n, Cx, false);
  FWN_LAUNCH_CHECK();
  ddi_phys_finish_kernel<<<(int)cdiv(Cx, 128), 128, 0, st>>>(scratch, rows, Cx, 0, fp.an_b, fp.an_s, fp.an_is, fp.raw_b, fp.raw_logs,
                                                            fp.off2log, m->d_an_logdet);
  FWN_LAUNCH_CHECK();
  colsum_phys_kernel<<<grid, 256, 256 * sizeof(double), st>>>(X, fp.an_b, scratch + Cx, n, Cx, true);
  FWN_LAUNCH_CHECK();
  ddi_phys_finish_kernel<<<(int)cdiv(Cx, 128), 128, 0, st>>>(scratch, rows, Cx, 1, fp.an_b, fp.an_s, fp.an_is, fp.raw_b, fp.raw_logs,
                                                            fp.off2log, m->d_an_logdet);
  FWN_LAUNCH_CHECK();
  return 0;
}

__global__ void finish_forward_kernel(const double* sums, const double* an_logdet, float* logp_out, float* logdet_out, double n) {
  // log_p = mean(0.5(-log 2pi - z^2)) (model.py:343); logdet = sum_flows [mean_c(3 logs) + mean(-log_s)/2] (model.py:80,135,342)
  if (logp_out) *logp_out = (float)(0.5 * (-1.8378770664093454835606594728112 - sums[1] / n));
  if (logdet_out) *logdet_out = (float)(*an_logdet - sums[0] / n);
}

int run_upsample(const Model* m, const Workspace& w, const float* c_in, int B, int T, cudaStream_t st) {
  const fwn_config& c = m->cfg;
  const bool bf16 = is_mixed(c.precision);
  int Tm = T / m->hop;
  const float* in = c_in;
  for (int i = 0; i < c.n_upsample; ++i) {
    const int s = c.upsample_scales[i];
    const bool last = i == c.n_upsample - 1;
    // algorithmic bytes: read the stage input once, write the stage output once
    const double bytes = (double)B * Tm * c.num_mels * 4.0 + (double)B * Tm * s * c.num_mels * (last && bf16 ? 2.0 : 4.0);
    prof_begin(const_cast<Model*>(m), PROF_UPSAMPLE, bytes, st);
    const_cast<Model*>(m)->launches++;
    if (last) {
      if (upsample_stage(in, m->up_w[i], m->up_b[i], w.cA, w.cB, B, Tm, c.num_mels, s, true, bf16 ? c.precision : 0, st)) return 1;
    } else {
      float* out = w.up[i & 1];
      if (upsample_stage(in, m->up_w[i], m->up_b[i], out, nullptr, B, Tm, c.num_mels, s, false, 0, st)) return 1;
      in = out;
    }
    prof_end(const_cast<Model*>(m), st);
    Tm *= s;
  }
  return 0;
}

// FWN_FRONT_DIRECT=0 keeps the shallow blocks' front conv on the tensor-core path (diagnostics / A-B timing)
static bool front_direct_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FWN_FRONT_DIRECT");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}
// FWN_FUSE_LAYER=0 runs every ResBlock layer as two launches (gate GEMM, res|skip GEMM) instead of the fused layer kernel
static bool layer_fusion_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FWN_FUSE_LAYER");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}
// FWN_FUSE_TAIL=0 runs the WaveNet tail as two launches (final conv, zero conv + affine) instead of the fused tail kernel
static bool tail_fusion_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FWN_FUSE_TAIL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}
// smallest launch (in 256-row tile pairs) that takes the fused layer kernel; FWN_FUSE_MIN_PAIRS overrides (A/B timing)
static int64_t layer_fusion_min_pairs() {
  static int64_t v = -1;
  if (v < 0) {
    const char* e = getenv("FWN_FUSE_MIN_PAIRS");
    v = e ? atoll(e) : 2 * (num_sms() / 2);
  }
  return v;
}
// FWN_LAYER_TAIL=0 keeps the tail out of the last layer's fused launch (it then runs as the separate tail kernel)
static bool layer_tail_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FWN_LAYER_TAIL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}
static bool fp32_front_on_tensor_cores() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FWN_FP32_ENGINE");
    v = (e && !strcmp(e, "simt")) ? 0 : 1;
  }
  return v == 1;
}

// One coupling WaveNet + the in-place flow update of X (ActNorm + AffineCoupling [+ change_order absorbed]).
int finish_forward(const double* sums, const double* an_logdet, float* logp_out, float* logdet_out, double n, cudaStream_t st) {
  finish_forward_kernel<<<1, 1, 0, st>>>(sums, an_logdet, logp_out, logdet_out, n);
  FWN_LAUNCH_CHECK();
  return 0;
}

static int run_flow(Model* m, const Workspace& w, const FlowPack& fp, float* X, int B, int Ti, bool reverse, cudaStream_t st) {
  const fwn_config& c = m->cfg;
  const int F = c.filter_size, L = c.n_layer;
  const bool bf16 = is_mixed(c.precision);
  auto shift_of = [&](int k, int d) { return c.causal ? (k - 2) * d : (k - 1) * d; };  // modules.py:12-15,27

  FrontArgs fa;
  fa.X = X; fa.Cx = fp.Cx; fa.nq = fp.nq; fa.a_off = fp.a_off; fa.off2log = fp.off2log;
  fa.an_b = reverse ? nullptr : fp.an_b;
  fa.an_s = reverse ? nullptr : fp.an_s;
  fa.W = fp.front_w; fa.bias = fp.front_b; fa.H = w.h0; fa.B = B; fa.Ti = Ti; fa.F = F;
  for (int k = 0; k < 3; ++k) fa.shift[k] = shift_of(k, 1);
  const double rows = (double)B * Ti;
  prof_begin(m, PROF_FRONT, 2.0 * rows * 3 * fp.nq * F, st);
  if (!bf16 && fp.w3[GEMM_FRONT].p && fp32_front_on_tensor_cores()) {
    // fp32 mode: gather (+ActNorm) the pass-through half, then the front conv is three time-shifted K segments on the split engine
    const int kq = (fp.nq + 3) / 4 * 4;
    m->launches++;
    if (front_pack_f32(X, fp.Cx, fp.nq, kq, fp.off2log, fa.an_b, fa.an_s, reinterpret_cast<float*>(w.a0), (int64_t)B * Ti, st)) return 1;
    GemmArgs g = {};
    g.B = B; g.Ti = Ti;
    for (int k = 0; k < 3; ++k) g.seg[k] = Seg{w.a0, kq, fa.shift[k], fp.nq, k * fp.front_k16};
    g.nseg = 3;
    g.W = fp.front_w; g.ldw = F; g.N = F;
    g.e.bias = fp.front_b; g.e.out0 = w.h0; g.e.ld = F; g.e.relu = 1; g.e.F = F;
    if (run_gemm(m, g, EPI_PLAIN, GEMM_FRONT, fp, st)) return 1;
  } else if (!bf16) {
    m->launches++;
    if (front_conv(fa, false, st)) return 1;
  } else if (front_direct_supported(fa) && front_direct_enabled()) {
    // mixed modes, shallow blocks: 6 .. 24 MAC per output straight from X on the CUDA cores (write-bound)
    m->launches++;
    if (front_direct(fa, c.precision == FWN_MIXED_FP16, st)) return 1;
  } else {
    // mixed mode: tiny gather/cast kernel, then the front conv is three time-shifted K segments on the tcgen05 engine
    const int kq = (fp.nq + 7) / 8 * 8;
    m->launches++;
    if (front_pack(X, fp.Cx, fp.nq, kq, fp.off2log, fa.an_b, fa.an_s, w.a0, (int64_t)B * Ti, c.precision == FWN_MIXED_FP16, st)) return 1;
    GemmArgs g = {};
    g.B = B; g.Ti = Ti;
    for (int k = 0; k < 3; ++k) g.seg[k] = Seg{w.a0, kq, fa.shift[k], fp.nq, k * fp.front_k16};
    g.nseg = 3;
    g.W = fp.front_wtc; g.ldw = fp.front_ld; g.N = F;
    g.e.bias = fp.front_b; g.e.out0 = w.h0; g.e.ld = F; g.e.relu = 1; g.e.F = F;
    if (run_gemm(m, g, EPI_PLAIN, GEMM_FRONT, fp, st)) return 1;
  }
  prof_end(m, st);

  // the WaveNet tail (final 1x1 + ReLU, zero conv + ActNorm / affine coupling): described here because the last layer's fused launch
  // can carry it (layer_tc.cu, tail ops)
  GemmArgs tf_ = {};
  tf_.B = B; tf_.Ti = Ti;
  tf_.seg[0] = Seg{w.s, F, 0, F, 0};
  tf_.nseg = 1;
  tf_.W = fp.final_w; tf_.ldw = fp.final_ld; tf_.N = F;
  tf_.e.bias = fp.final_b; tf_.e.out0 = w.u; tf_.e.ld = F; tf_.e.relu = 1; tf_.e.F = F;
  GemmArgs tz_ = {};
  tz_.B = B; tz_.Ti = Ti;
  tz_.seg[0] = Seg{w.u, F, 0, F, 0};
  tz_.nseg = 1;
  tz_.W = fp.zero_w; tz_.ldw = fp.zero_ld; tz_.N = 2 * fp.nq;
  tz_.e.bias = fp.zero_b; tz_.e.F = F;
  tz_.e.X = X; tz_.e.Cx = fp.Cx; tz_.e.nq = fp.nq; tz_.e.a_off = fp.a_off; tz_.e.b_off = fp.b_off;
  tz_.e.an_b = fp.an_b; tz_.e.an_s = reverse ? fp.an_is : fp.an_s;
  tz_.e.logdet_acc = reverse ? nullptr : w.sums;
  tz_.e.reverse = reverse;
  tz_.e.pairs_adjacent = fp.pairs_adjacent;
  tz_.e.b_odd = fp.b_odd;
  const double final_flop = 2.0 * rows * F * F, zero_flop = 2.0 * rows * F * (c.affine ? fp.Cx : fp.nq);
  const bool fuse_tail = m->fuse_layer < 0 ? tail_fusion_enabled() : m->fuse_layer != 0;
  const bool tail_ok = bf16 && fuse_tail && tc_tail_supported(m, tf_, tz_);
  bool tail_done = false;

  void* hin = w.h0;
  void* hout = w.h1;
  const void* cond = fp.cond_half == 0 ? w.cA : w.cB;
  int d = 1;
  for (int n = 0; n < L; ++n, d *= 3) {
    GemmArgs g = {};
    g.B = B; g.Ti = Ti;
    for (int k = 0; k < 3; ++k) g.seg[k] = Seg{hin, F, shift_of(k, d), F, k * F};
    g.seg[3] = Seg{cond, fp.Kc, 0, fp.Kc, 3 * F};
    g.nseg = fp.ahead ? 3 : 4;
    g.W = fp.gate_w[n]; g.ldw = fp.gate_ld; g.N = 2 * F;
    g.e.bias = fp.gate_b[n]; g.e.out0 = w.o; g.e.F = F;
    if (fp.ahead) {   // this layer's slice of the block's conditioning projection (run_cond_ahead)
      const Model::CondAhead& ca = m->cond_w[(size_t)(&fp - m->flows.data()) / c.n_flow * 2 + fp.cond_half];
      g.e.in0 = w.pc[fp.cond_half] + (size_t)(fp.ahead_slot + n) * 2 * F;
      g.e.ld = ca.N;
    }
    const bool last = n == L - 1;
    GemmArgs r = {};
    r.B = B; r.Ti = Ti;
    r.seg[0] = Seg{w.o, F, 0, F, 0};
    r.nseg = 1;
    r.W = fp.rs_w[n]; r.ldw = fp.rs_ld[n]; r.N = last ? F : 2 * F;
    r.e.bias = fp.rs_b[n]; r.e.F = F; r.e.has_res = !last; r.e.relu = last;
    r.e.in0 = hin; r.e.out0 = hout; r.e.in1 = n > 0 ? w.s : nullptr; r.e.out1 = w.s;
    const double gate_flop = 2.0 * rows * (3 * F + (fp.ahead ? 0 : fp.Kc)) * 2 * F, rs_flop = 2.0 * rows * F * r.N;
    // default: fuse when the launch has at least two waves of row-tile pairs -- below that the fused kernel (one CTA pair walks
    // G0, G1, R, K of its rows in sequence) loses to two launches that spread the column tiles over twice as many SMs
    // (C1, 1 s / batch 1: 4.42 ms fused vs 3.62 ms per pass)
    const int64_t row_pairs = ((int64_t)B * ((Ti + 127) / 128) + 1) / 2;
    const bool fuse = m->fuse_layer < 0 ? (layer_fusion_enabled() && row_pairs >= layer_fusion_min_pairs()) : m->fuse_layer != 0;
    if (bf16 && fuse && tc_layer_supported(m, g, r)) {
      // gate GEMM -> tanh*sigmoid -> res|skip 1x1 in one launch: o stays in shared memory (layer_tc.cu); the last layer's launch
      // also carries the tail (relu(skip sum) and relu(final) stay in shared memory too) when it has a running skip sum to stage
      const bool with_tail = last && tail_ok && n > 0 && layer_tail_enabled();
      prof_begin(m, PROF_GATE, gate_flop + rs_flop + (with_tail ? final_flop + zero_flop : 0.0), st);
      m->launches++;
      if (tc_run_layer(m, g, r, n, fp, st, with_tail ? &tf_ : nullptr, with_tail ? &tz_ : nullptr)) return 1;
      prof_end(m, st);
      tail_done = with_tail;
    } else {
      prof_begin(m, PROF_GATE, gate_flop, st);
      if (run_gemm(m, g, EPI_GATE, GEMM_GATE0 + n, fp, st)) return 1;
      prof_end(m, st);
      prof_begin(m, PROF_RES_SKIP, rs_flop, st);
      if (run_gemm(m, r, EPI_RES_SKIP, GEMM_RS0 + n, fp, st)) return 1;
      prof_end(m, st);
    }
    std::swap(hin, hout);
  }
  if (tail_done) return 0;
  if (tail_ok) {
    // final 1x1 + ReLU -> zero conv -> ActNorm / affine coupling on x in one launch: u stays in shared memory (tail_tc.cu)
    prof_begin(m, PROF_FINAL, final_flop + zero_flop, st);
    m->launches++;
    if (tc_run_tail(m, tf_, tz_, fp, st)) return 1;
    prof_end(m, st);
  } else {
    prof_begin(m, PROF_FINAL, final_flop, st);
    if (run_gemm(m, tf_, EPI_PLAIN, GEMM_FINAL, fp, st)) return 1;
    prof_end(m, st);
    prof_begin(m, PROF_ZERO_AFFINE, zero_flop, st);
    if (run_gemm(m, tz_, EPI_AFFINE, GEMM_ZERO, fp, st)) return 1;
    prof_end(m, st);
  }
  return 0;
}

// The dependent chain of flows of one pass, on the workspace's X buffer (all pointers inside are workspace / pack pointers,
// so the captured graph stays valid as long as the workspace and the prepacked weights do).
// Conditioning projections of a deep block (FlowPack::ahead): P_h[rows, slices * 2F] = c_h . W_c for both mel halves, all flows and
// layers of the block in one wide GEMM each (modules.py:117-118: _filter_conv_c / _gate_conv_c do not depend on the flow state).
static int run_cond_ahead(Model* m, const Workspace& w, int block, int B, int Ti, cudaStream_t st) {
  if (m->cond_w.empty()) return 0;
  for (int h = 0; h < 2; ++h) {
    const Model::CondAhead& ca = m->cond_w[(size_t)block * 2 + h];
    if (!ca.w) continue;
    FWN_CHECK(w.pc[h], "internal: workspace has no projection buffer");
    GemmArgs g = {};
    g.B = B; g.Ti = Ti;
    g.seg[0] = Seg{h == 0 ? w.cA : w.cB, ca.Kc, 0, ca.Kc, 0};
    g.nseg = 1; g.N = ca.N;
    g.e.out0 = w.pc[h]; g.e.ld = ca.N; g.e.F = m->cfg.filter_size;
    prof_begin(m, PROF_GATE, 2.0 * B * Ti * ca.Kc * ca.N, st);
    m->launches++;
    if (tc_gemm16(g, EPI_PLAIN_F32, ca.w, ca.Kpad, ca.N, m->cfg.precision == FWN_MIXED_FP16, st)) return 1;
    prof_end(m, st);
  }
  return 0;
}

static int run_chain(Model* m, const Workspace& w, int B, int T, bool reverse, cudaStream_t st) {
  const fwn_config& cf = m->cfg;
  if (!reverse) {
    for (int i = 0; i < cf.n_block; ++i) {
      if (run_cond_ahead(m, w, i, B, T >> (i + 1), st)) return 1;
      for (int j = 0; j < cf.n_flow; ++j)
        if (run_flow(m, w, m->flows[(size_t)i * cf.n_flow + j], w.x, B, T >> (i + 1), false, st)) return 1;
    }
  } else {
    for (int i = cf.n_block - 1; i >= 0; --i) {
      if (run_cond_ahead(m, w, i, B, T >> (i + 1), st)) return 1;
      for (int j = cf.n_flow - 1; j >= 0; --j)
        if (run_flow(m, w, m->flows[(size_t)i * cf.n_flow + j], w.x, B, T >> (i + 1), true, st)) return 1;
    }
  }
  return 0;
}

static bool graphs_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FWN_GRAPH");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// First call for a (direction, shape, workspace): run eagerly (also sets kernel attributes, builds tensor maps).
// Second call: capture the chain into a graph and launch it.  Later calls: replay.
static int run_chain_graphed(Model* m, const Workspace& w, int B, int T, bool reverse, cudaStream_t st) {
  if (!graphs_enabled() || m->prof_on) return run_chain(m, w, B, T, reverse, st);
  Model::PassGraph* g = nullptr;
  for (auto& e : m->graphs)
    if (e.reverse == (int)reverse && e.B == B && e.T == T && e.ws == (const void*)w.sums) g = &e;
  if (!g) {
    if (m->graphs.size() >= 16) model_drop_graphs(m);
    m->graphs.push_back(Model::PassGraph{(int)reverse, B, T, (const void*)w.sums, 0, nullptr, 0});
    return run_chain(m, w, B, T, reverse, st);
  }
  if (g->state == 0) {
    const int64_t before = m->launches;
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();  // e.g. the legacy default stream cannot be captured: stay eager on this stream/shape
      g->state = -1;
      return run_chain(m, w, B, T, reverse, st);
    }
    const int rc = run_chain(m, w, B, T, reverse, st);
    cudaError_t e = cudaStreamEndCapture(st, &graph);
    if (rc || e != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      g->state = -1;  // capture not possible: stay eager for this shape
      if (rc) return 1;
      return run_chain(m, w, B, T, reverse, st);
    }
    g->launches = m->launches - before;
    m->launches = before;
    e = cudaGraphInstantiate(&g->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
      cudaGetLastError();
      g->exec = nullptr;
      g->state = -1;
      return run_chain(m, w, B, T, reverse, st);
    }
    g->state = 1;
  }
  if (g->state == 1) {
    FWN_CUDA(cudaGraphLaunch(g->exec, st));
    m->launches += g->launches;
    return 0;
  }
  return run_chain(m, w, B, T, reverse, st);
}

static int check_pass_args(Model* m, const void* x, const void* c, const int32_t* g, int B, int T, void* ws, int64_t ws_bytes,
                           Workspace* w) {
  FWN_CHECK(m && m->packed, "model not prepacked: call fwn_prepack after fwn_set_param");
  FWN_CHECK(x && c, "null input pointer");
  // model.py:320-321 / 353-354: `if g is None and gin_channels > 0: raise ValueError('g is None')`
  FWN_CHECK(!(m->cfg.gin_channels > 0 && g == nullptr), "g is None");
  if (model_plan(m, B, T, w, (char*)ws)) return 1;
  FWN_CHECK(ws && ws_bytes >= (int64_t)w->bytes, "workspace too small: need %lld bytes, got %lld", (long long)w->bytes, (long long)ws_bytes);
  return 0;
}

int model_forward(Model* m, const float* x, const float* c, const int32_t* g, int B, int T, float* z_out, float* logp_out,
                  float* logdet_out, int ddi, void* ws, int64_t ws_bytes, cudaStream_t st) {
  Workspace w;
  if (check_pass_args(m, x, c, g, B, T, ws, ws_bytes, &w)) return 1;
  if (prepare_engine(m, w, B, T, st)) return 1;
  if (train_ensure_full_planes(m, st)) return 1;   // after bf16 optimizer steps planes 1, 2 of the fp32 operand planes are stale
  // the flow variable lives in the workspace for the whole pass (fixed address -> the flow chain can be a CUDA graph)
  float* X = w.x;
  FWN_CUDA(cudaMemcpyAsync(X, x, (size_t)B * T * 4, cudaMemcpyDeviceToDevice, st));
  FWN_CUDA(cudaMemsetAsync(w.sums, 0, 8 * sizeof(double), st));
  if (ddi) FWN_CUDA(cudaMemsetAsync(m->d_an_logdet, 0, sizeof(double), st));
  if (run_upsample(m, w, c, B, T, st)) return 1;
  // The speaker embedding g is looked up, tiled and squeezed by the reference (model.py:330-336) but never
  // reaches a kernel: WaveNet.__call__ drops it (modules.py:188-189).  Outputs do not depend on g.
  const fwn_config& cf = m->cfg;
  if (ddi) {
    for (int i = 0; i < cf.n_block; ++i) {
      const int Ti = T >> (i + 1);
      if (run_cond_ahead(m, w, i, B, Ti, st)) return 1;
      for (int j = 0; j < cf.n_flow; ++j) {
        const FlowPack& fp = m->flows[(size_t)i * cf.n_flow + j];
        if (ddi_flow(m, fp, X, (int64_t)B * Ti, w.ddi, st)) return 1;
        if (run_flow(m, w, fp, X, B, Ti, false, st)) return 1;
      }
    }
  } else if (run_chain_graphed(m, w, B, T, false, st)) {
    return 1;
  }
  if (sumsq(X, w.sums + 1, (int64_t)B * T, st)) return 1;
  finish_forward_kernel<<<1, 1, 0, st>>>(w.sums, m->d_an_logdet, logp_out, logdet_out, (double)B * T);
  FWN_LAUNCH_CHECK();
  if (z_out) FWN_CUDA(cudaMemcpyAsync(z_out, X, (size_t)B * T * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}

int model_reverse(Model* m, const float* z, const float* c, const int32_t* g, int B, int T, float* x_out, void* ws, int64_t ws_bytes,
                  cudaStream_t st) {
  Workspace w;
  if (check_pass_args(m, z, c, g, B, T, ws, ws_bytes, &w)) return 1;
  FWN_CHECK(x_out, "null output pointer");
  FWN_CHECK(m->rev_ok, "fused reverse needs an even n_flow: with odd n_flow the reference's reverse is not the inverse of forward "
                       "(change_order parity, model.py:199,359) -- use the per-op Block/Flow API for that case");
  if (prepare_engine(m, w, B, T, st)) return 1;
  if (train_ensure_full_planes(m, st)) return 1;
  float* X = w.x;
  FWN_CUDA(cudaMemcpyAsync(X, z, (size_t)B * T * 4, cudaMemcpyDeviceToDevice, st));
  if (run_upsample(m, w, c, B, T, st)) return 1;
  if (run_chain_graphed(m, w, B, T, true, st)) return 1;
  if (x_out != X) FWN_CUDA(cudaMemcpyAsync(x_out, X, (size_t)B * T * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}

int model_receptive_halo(const Model* m) {
  const fwn_config& c = m->cfg;
  int rw = 1, d = 1;  // front conv (k=3, d=1) + layers d = 3^n; causal nets look back twice as far
  for (int n = 0; n < c.n_layer; ++n, d *= 3) rw += d;
  if (c.causal) rw *= 2;
  int64_t halo = 0;
  for (int i = 0; i < c.n_block; ++i) halo += (int64_t)c.n_flow * rw * (2 << i);
  halo += m->hop;  // the transposed-conv upsampler sees +-1 input frame per stage
  int64_t q = std::max(m->hop, 1 << c.n_block);
  while (q % m->hop || q % (1 << c.n_block)) ++q;
  return (int)((halo + q - 1) / q * q);
}

}  // namespace fwn

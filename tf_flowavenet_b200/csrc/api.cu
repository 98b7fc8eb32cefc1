// extern "C" surface of libflowavenet_b200.so -- see include/flowavenet_b200.h for the contract and the
// reference interface (file:line) each entry point replaces.
#include <string.h>

#include "common.cuh"
#include "model.h"
#include "train.h"

using namespace fwn;

namespace fwn {
int tc_conv1d(const void* x, const void* w, const float* bias, void* y, int B, int T, int Cin, int Cout, int ksize, int dilation, int causal,
              int relu, cudaStream_t st);
}

struct fwn_model {
  Model* m;
};

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

static int require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    set_error("no CUDA device available (%s); libflowavenet_b200 has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

// stream-ordered scratch for the per-op entry points
struct Scratch {
  void* p = nullptr;
  cudaStream_t st;
  int get(size_t bytes, cudaStream_t s) {
    st = s;
    FWN_CUDA(cudaMallocAsync(&p, bytes, s));
    return 0;
  }
  ~Scratch() {
    if (p) cudaFreeAsync(p, st);
  }
};

__global__ void mul_vec_kernel(const float* a, const float* b, float* y, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = a[i] * b[i];
}
static int mul_vec(const float* a, const float* b, float* y, int n, cudaStream_t st) {
  mul_vec_kernel<<<(n + 127) / 128, 128, 0, st>>>(a, b, y, n);
  if (cudaGetLastError() != cudaSuccess) {
    fwn::set_error("mul_vec launch failed");
    return 1;
  }
  return 0;
}

extern "C" {

const char* fwn_last_error(void) { return get_error(); }
int fwn_abi_version(void) { return FWN_ABI_VERSION; }

int fwn_create(const fwn_config* cfg, fwn_handle* out) {
  if (require_device()) return 1;
  FWN_CHECK(out, "fwn_create: null out pointer");
  Model* m = nullptr;
  if (model_create(cfg, &m)) return 1;
  *out = new fwn_model{m};
  return 0;
}

int fwn_destroy(fwn_handle h) {
  if (!h) return 0;
  engine_free(h->m);
  model_destroy(h->m);
  delete h;
  return 0;
}

int fwn_num_params(fwn_handle h) { return h ? (int)h->m->params.size() : -1; }

int fwn_param_info(fwn_handle h, int index, const char** name, int64_t shape[4], int* rank) {
  FWN_CHECK(h && index >= 0 && index < (int)h->m->params.size(), "fwn_param_info: bad index %d", index);
  const ParamDesc& d = h->m->params[index];
  if (name) *name = d.name.c_str();
  if (rank) *rank = (int)d.shape.size();
  if (shape)
    for (size_t i = 0; i < 4; ++i) shape[i] = i < d.shape.size() ? d.shape[i] : 1;
  return 0;
}

static int find_param(fwn_handle h, const char* name, int64_t numel, const ParamDesc** out) {
  FWN_CHECK(h && name, "null handle or name");
  auto it = h->m->index.find(name);
  FWN_CHECK(it != h->m->index.end(), "unknown variable '%s'", name);
  const ParamDesc& d = h->m->params[it->second];
  FWN_CHECK(d.numel == numel, "variable '%s' has %lld elements, got %lld", name, (long long)d.numel, (long long)numel);
  *out = &d;
  return 0;
}

int fwn_set_param(fwn_handle h, const char* name, const float* dev_src, int64_t numel, void* stream) {
  const ParamDesc* d;
  if (find_param(h, name, numel, &d)) return 1;
  FWN_CUDA(cudaMemcpyAsync(h->m->raw + d->offset, dev_src, (size_t)numel * 4, cudaMemcpyDeviceToDevice, S(stream)));
  h->m->packed = false;
  return 0;
}

int fwn_get_param(fwn_handle h, const char* name, float* dev_dst, int64_t numel, void* stream) {
  const ParamDesc* d;
  if (find_param(h, name, numel, &d)) return 1;
  FWN_CUDA(cudaMemcpyAsync(dev_dst, h->m->raw + d->offset, (size_t)numel * 4, cudaMemcpyDeviceToDevice, S(stream)));
  return 0;
}

int fwn_prepack(fwn_handle h, void* stream) {
  FWN_CHECK(h, "null handle");
  return model_prepack(h->m, S(stream));
}

int64_t fwn_workspace_bytes(fwn_handle h, int B, int T) {
  if (!h) {
    set_error("null handle");
    return -1;
  }
  Workspace w;
  if (model_plan(h->m, B, T, &w, nullptr)) return -1;
  return (int64_t)w.bytes;
}

int fwn_forward(fwn_handle h, const float* x, const float* c, const int32_t* g, int B, int T, float* z_out, float* logp_out,
                float* logdet_out, int ddi, void* workspace, int64_t workspace_bytes, void* stream) {
  FWN_CHECK(h, "null handle");
  h->m->launches = 0;
  return model_forward(h->m, x, c, g, B, T, z_out, logp_out, logdet_out, ddi, workspace, workspace_bytes, S(stream));
}

int fwn_reverse(fwn_handle h, const float* z, const float* c, const int32_t* g, int B, int T, float* x_out, void* workspace,
                int64_t workspace_bytes, void* stream) {
  FWN_CHECK(h, "null handle");
  h->m->launches = 0;
  return model_reverse(h->m, z, c, g, B, T, x_out, workspace, workspace_bytes, S(stream));
}

// ---- training step (train.cu)
int fwn_train_enable(fwn_handle h, void* stream) {
  FWN_CHECK(h, "null handle");
  return train_enable(h->m, S(stream));
}
int64_t fwn_train_workspace_bytes(fwn_handle h, int B, int T) {
  if (!h) {
    set_error("null handle");
    return -1;
  }
  return train_workspace_bytes(h->m, B, T);
}
int64_t fwn_param_floats(fwn_handle h) { return h ? h->m->raw_floats : -1; }
int64_t fwn_grad_floats(fwn_handle h) { return h ? train_grad_floats(h->m) : -1; }
int64_t fwn_param_offset(fwn_handle h, int index) {
  if (!h || index < 0 || index >= (int)h->m->params.size()) {
    set_error("fwn_param_offset: bad handle or index");
    return -1;
  }
  return h->m->params[index].offset;
}
int fwn_params_ptr(fwn_handle h, float** dev_ptr) {
  FWN_CHECK(h && dev_ptr, "null argument");
  *dev_ptr = h->m->raw;
  return 0;
}
int fwn_loss_and_grads(fwn_handle h, const float* x, const float* c, const int32_t* g, int B, int T, float* logp_out, float* logdet_out,
                       float* grads, int64_t grad_floats, void* workspace, int64_t workspace_bytes, void* stream) {
  FWN_CHECK(h, "null handle");
  return train_loss_and_grads(h->m, x, c, g, B, T, logp_out, logdet_out, grads, grad_floats, workspace, workspace_bytes, S(stream));
}
int fwn_grad_bucket_count(fwn_handle h) {
  if (!h) {
    set_error("null handle");
    return -1;
  }
  const int n = train_bucket_count(h->m);
  if (n < 0) set_error("training not enabled");
  return n;
}
int fwn_grad_bucket_range(fwn_handle h, int bucket, int64_t* offset, int64_t* count) {
  FWN_CHECK(h && offset && count, "null argument");
  return train_bucket_range(h->m, bucket, offset, count);
}
int fwn_grad_bucket_wait(fwn_handle h, int bucket, void* consumer_stream) {
  FWN_CHECK(h, "null handle");
  return train_bucket_wait(h->m, bucket, S(consumer_stream));
}
int fwn_grad_global_norm(fwn_handle h, const float* grads, float* norm_out, void* stream) {
  FWN_CHECK(h && grads && norm_out, "null argument");
  return train_grad_norm(h->m, grads, norm_out, S(stream));
}
int fwn_apply_gradients(fwn_handle h, const float* grads, float lr, float beta1, float beta2, float eps, float clip_norm, int64_t step,
                        void* stream) {
  FWN_CHECK(h && grads, "null argument");
  return train_apply(h->m, grads, lr, beta1, beta2, eps, clip_norm, step, S(stream));
}
int fwn_set_split_terms(fwn_handle h, int inference_terms, int training_terms) {
  FWN_CHECK(h, "null handle");
  FWN_CHECK((inference_terms == 3 || inference_terms == 6) && (training_terms == 3 || training_terms == 6), "split terms must be 3 or 6");
  h->m->terms_infer = h->m->cur_terms = inference_terms;
  h->m->terms_train = training_terms;
  model_drop_graphs(h->m);   // captured launches carry the old setting
  return 0;
}
int fwn_set_layer_fusion(fwn_handle h, int mode) {
  FWN_CHECK(h, "null handle");
  FWN_CHECK(mode >= -1 && mode <= 1, "layer fusion mode must be -1 (default), 0 or 1");
  h->m->fuse_layer = mode;
  model_drop_graphs(h->m);   // captured launches carry the old setting
  return 0;
}
int fwn_set_train_compute(fwn_handle h, int precision) {
  FWN_CHECK(h, "null handle");
  FWN_CHECK(precision == FWN_FP32 || precision == FWN_MIXED_BF16, "training compute precision must be FWN_FP32 or FWN_MIXED_BF16");
  if (precision == FWN_MIXED_BF16)
    FWN_CHECK(h->m->cfg.filter_size == 256 && h->m->cfg.num_mels % 8 == 0,
              "bf16 training needs filter_size 256 and num_mels %% 8 == 0 (TMA strides / tile shape)");
  h->m->train_bf16 = precision == FWN_MIXED_BF16;
  return 0;
}
int fwn_wgrad_bf16(const void* a, const void* dy, float* dw, float* dbias, int B, int T, int K, int N, int shift, void* stream) {
  FWN_CHECK(a && dy && dw, "null argument");
  Wgrad16Args w = {};
  w.seg[0] = Seg{a, K, shift, K, 0};
  w.nseg = 1;
  w.dY0 = dy; w.ld0 = N; w.n0cols = N; w.N = N;
  w.dW = dw; w.ldw = N; w.B = B; w.Ti = T;
  return wgrad_tc(w, dbias, S(stream));
}
int fwn_set_train_exact_forward(fwn_handle h, int on) {
  FWN_CHECK(h, "null handle");
  h->m->train_exact_fwd = on != 0;
  return 0;
}
int fwn_get_train_state(fwn_handle h, int which, float* dev_dst, int64_t numel, void* stream) {
  FWN_CHECK(h && dev_dst, "null argument");
  return train_state_copy(h->m, which, dev_dst, nullptr, numel, S(stream));
}
int fwn_set_train_state(fwn_handle h, int which, const float* dev_src, int64_t numel, void* stream) {
  FWN_CHECK(h && dev_src, "null argument");
  return train_state_copy(h->m, which, nullptr, dev_src, numel, S(stream));
}
int fwn_repack(fwn_handle h, void* stream) {
  FWN_CHECK(h, "null handle");
  return train_repack(h->m, S(stream));
}

int64_t fwn_last_launches(fwn_handle h) { return h ? h->m->launches : -1; }

int fwn_profile_enable(fwn_handle h, int on) {
  FWN_CHECK(h, "null handle");
  h->m->prof_on = on != 0;
  h->m->prof_used = 0;
  return 0;
}
int fwn_profile_read(fwn_handle h, double ms[8], int64_t launches[8], double work[8]) {
  FWN_CHECK(h && ms && launches && work, "null argument");
  return prof_read(h->m, ms, launches, work);
}

// ---- host-buffer convenience: H2D, pass, D2H (what synthesize.py:44-46's sess.run does end to end)
static int host_buffers(Model* m, int B, int T, float** d_x, float** d_c, float** d_out, float** d_scal) {
  Workspace w;
  if (model_plan(m, B, T, &w, nullptr)) return 1;
  if ((int64_t)w.bytes > m->host_ws_bytes) {
    if (m->host_ws) cudaFree(m->host_ws);
    m->host_ws = nullptr;
    m->host_ws_bytes = 0;
    FWN_CUDA(cudaMalloc(&m->host_ws, w.bytes));
    m->host_ws_bytes = (int64_t)w.bytes;
  }
  const int Tm = T / m->hop;
  const size_t nx = ((size_t)B * T + 63) & ~size_t(63), nc = ((size_t)B * Tm * m->cfg.num_mels + 63) & ~size_t(63);
  const int64_t need = (int64_t)(2 * nx + nc + 64) * 4;
  if (need > m->host_io_bytes) {
    if (m->host_io) cudaFree(m->host_io);
    m->host_io = nullptr;
    m->host_io_bytes = 0;
    FWN_CUDA(cudaMalloc(&m->host_io, need));
    m->host_io_bytes = need;
  }
  if (!m->host_stream) FWN_CUDA(cudaStreamCreateWithFlags(&m->host_stream, cudaStreamNonBlocking));
  float* p = (float*)m->host_io;
  *d_x = p;
  *d_out = p + nx;
  *d_c = p + 2 * nx;
  *d_scal = p + 2 * nx + nc;
  return 0;
}

int fwn_forward_host(fwn_handle h, const float* x, const float* c, const int32_t* g, int B, int T, float* z_out, float* logp_out,
                     float* logdet_out) {
  FWN_CHECK(h && x && c, "null argument");
  Model* m = h->m;
  float *d_x, *d_c, *d_out, *d_s;
  if (host_buffers(m, B, T, &d_x, &d_c, &d_out, &d_s)) return 1;
  cudaStream_t st = m->host_stream;
  FWN_CUDA(cudaMemcpyAsync(d_x, x, (size_t)B * T * 4, cudaMemcpyHostToDevice, st));
  FWN_CUDA(cudaMemcpyAsync(d_c, c, (size_t)B * (T / m->hop) * m->cfg.num_mels * 4, cudaMemcpyHostToDevice, st));
  m->launches = 0;
  // g only needs to be non-null when gin_channels > 0 (it never reaches a kernel)
  if (model_forward(m, d_x, d_c, g, B, T, d_out, d_s, d_s + 1, 0, m->host_ws, m->host_ws_bytes, st)) return 1;
  float sc[2];
  FWN_CUDA(cudaMemcpyAsync(sc, d_s, 8, cudaMemcpyDeviceToHost, st));
  if (z_out) FWN_CUDA(cudaMemcpyAsync(z_out, d_out, (size_t)B * T * 4, cudaMemcpyDeviceToHost, st));
  FWN_CUDA(cudaStreamSynchronize(st));
  if (logp_out) *logp_out = sc[0];
  if (logdet_out) *logdet_out = sc[1];
  return 0;
}

int fwn_reverse_host(fwn_handle h, const float* z, const float* c, const int32_t* g, int B, int T, float* x_out) {
  FWN_CHECK(h && z && c && x_out, "null argument");
  Model* m = h->m;
  float *d_x, *d_c, *d_out, *d_s;
  if (host_buffers(m, B, T, &d_x, &d_c, &d_out, &d_s)) return 1;
  cudaStream_t st = m->host_stream;
  FWN_CUDA(cudaMemcpyAsync(d_x, z, (size_t)B * T * 4, cudaMemcpyHostToDevice, st));
  FWN_CUDA(cudaMemcpyAsync(d_c, c, (size_t)B * (T / m->hop) * m->cfg.num_mels * 4, cudaMemcpyHostToDevice, st));
  m->launches = 0;
  if (model_reverse(m, d_x, d_c, g, B, T, d_out, m->host_ws, m->host_ws_bytes, st)) return 1;
  FWN_CUDA(cudaMemcpyAsync(x_out, d_out, (size_t)B * T * 4, cudaMemcpyDeviceToHost, st));
  FWN_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int fwn_receptive_halo(fwn_handle h) {
  if (!h) {
    set_error("null handle");
    return -1;
  }
  return model_receptive_halo(h->m);
}

__global__ void copy_interior_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int T_ext, int halo_l, int Tl) {
  const int64_t n = (int64_t)B * Tl;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = i / Tl;
    int t = (int)(i - b * Tl);
    dst[i] = src[b * T_ext + halo_l + t];
  }
}

int fwn_reverse_chunk(fwn_handle h, const float* z_ext, const float* c_ext, int B, int T_ext, int halo_l, int halo_r, float* x_out,
                      void* workspace, int64_t workspace_bytes, void* stream) {
  FWN_CHECK(h, "null handle");
  Model* m = h->m;
  FWN_CHECK(halo_l >= 0 && halo_r >= 0 && halo_l + halo_r < T_ext, "bad halo sizes");
  Workspace w;
  if (model_plan(m, B, T_ext, &w, (char*)workspace)) return 1;
  FWN_CHECK(workspace && workspace_bytes >= (int64_t)w.bytes, "workspace too small");
  // overlap-recompute: run the pass on the extended chunk in the workspace's X buffer, keep the interior.
  m->launches = 0;
  if (model_reverse(m, z_ext, c_ext, nullptr, B, T_ext, w.x, workspace, workspace_bytes, S(stream))) return 1;
  const int Tl = T_ext - halo_l - halo_r;
  int64_t n = (int64_t)B * Tl;
  copy_interior_kernel<<<(int)std::min<int64_t>(cdiv(n, 256), 148 * 8), 256, 0, S(stream)>>>(w.x, x_out, B, T_ext, halo_l, Tl);
  FWN_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------- per-op entry points
int fwn_squeeze(const float* x, float* y, int B, int T, int C, void* stream) { return squeeze(x, y, B, T, C, S(stream)); }
int fwn_unsqueeze(const float* x, float* y, int B, int T, int C, void* stream) { return unsqueeze(x, y, B, T, C, S(stream)); }
int fwn_change_order(const float* x, float* y, int64_t rows, int C, void* stream) { return change_order(x, y, rows, C, S(stream)); }

int fwn_actnorm_fwd(const float* x, const float* b, const float* logs, float* y, float* logdet_out, int64_t rows, int C, void* stream) {
  return actnorm(x, b, logs, y, logdet_out, rows, C, false, S(stream));
}
int fwn_actnorm_rev(const float* x, const float* b, const float* logs, float* y, int64_t rows, int C, void* stream) {
  return actnorm(x, b, logs, y, nullptr, rows, C, true, S(stream));
}
int fwn_actnorm_ddi(const float* x, float* b_out, float* logs_out, int64_t rows, int C, void* stream) {
  FWN_CHECK(rows > 0 && C > 0, "actnorm_ddi: empty input");
  Scratch sc;
  if (sc.get(2 * (size_t)C * sizeof(double), S(stream))) return 1;
  return actnorm_ddi(x, b_out, logs_out, rows, C, (double*)sc.p, S(stream));
}
int fwn_affine_fwd(const float* x, const float* net, float* y, float* logdet_out, int64_t rows, int C, int affine_, void* stream) {
  Scratch sc;
  if (sc.get(sizeof(double), S(stream))) return 1;
  return affine(x, net, y, logdet_out, rows, C, affine_ != 0, false, (double*)sc.p, S(stream));
}
int fwn_affine_rev(const float* x, const float* net, float* y, int64_t rows, int C, int affine_, void* stream) {
  return affine(x, net, y, nullptr, rows, C, affine_ != 0, true, nullptr, S(stream));
}

int fwn_upsample_stage(const float* c_in, const float* kernel, const float* g, const float* bias, float* c_out, int B, int Tm, int mels,
                       int s, void* stream) {
  FWN_CHECK(s >= 2 && s % 2 == 0, "upsample scale %d must be even", s);
  Scratch sc;
  if (sc.get((size_t)2 * s * 3 * sizeof(float), S(stream))) return 1;
  if (upsample_weight_norm(kernel, g, (float*)sc.p, s, S(stream))) return 1;
  return upsample_stage(c_in, (const float*)sc.p, bias, c_out, nullptr, B, Tm, mels, s, false, 0, S(stream));
}

int fwn_conv1d(const float* x, const float* kernel, const float* wn_g, const float* bias, float* y, int B, int T, int Cin, int Cout,
               int ksize, int dilation, int causal, int relu, void* stream) {
  FWN_CHECK(ksize >= 1 && ksize <= 4, "conv1d: kernel_size %d unsupported (1..4)", ksize);
  FWN_CHECK(causal || ksize % 2 == 1, "conv1d: non-causal padding needs an odd kernel_size");
  Scratch sc;
  GemmArgs g = {};
  g.B = B; g.Ti = T;
  const int pad = causal ? dilation * (ksize - 1) : dilation * (ksize - 1) / 2;  // modules.py:12-15
  for (int k = 0; k < ksize; ++k) g.seg[k] = Seg{x, Cin, k * dilation - pad, Cin, k * Cin};
  g.nseg = ksize;
  g.W = kernel; g.ldw = Cout; g.N = Cout;
  g.e.bias = bias; g.e.out0 = y; g.e.ld = Cout; g.e.relu = relu; g.e.F = Cout;
  if (wn_g) {
    if (sc.get((size_t)Cout * sizeof(float), S(stream))) return 1;
    if (weight_norm_scale(kernel, wn_g, (float*)sc.p, ksize * Cin, Cout, S(stream))) return 1;
    g.e.colscale = (const float*)sc.p;
  }
  return simt_gemm(g, EPI_PLAIN, S(stream));
}

int fwn_conv1d_bf16(const void* x, const void* w_packed, const float* bias, void* y, int B, int T, int Cin, int Cout, int ksize,
                    int dilation, int causal, int relu, void* stream) {
  return tc_conv1d(x, w_packed, bias, y, B, T, Cin, Cout, ksize, dilation, causal, relu, S(stream));
}

int fwn_zero_conv1d(const float* x, const float* kernel, const float* bias, const float* scale, float* y, int64_t rows, int Cin, int Cout,
                    void* stream) {
  // (x.W + b) * e = x.W * e + b * e : column scale e plus a pre-scaled bias
  FWN_CHECK(rows < (int64_t)1 << 31, "zero_conv1d: too many rows");
  Scratch sc;
  if (sc.get((size_t)2 * Cout * sizeof(float), S(stream))) return 1;
  float* e = (float*)sc.p;
  float* be = e + Cout;
  if (exp3(scale, e, Cout, S(stream))) return 1;
  if (mul_vec(bias, e, be, Cout, S(stream))) return 1;  // be = bias * e
  GemmArgs g = {};
  g.B = 1; g.Ti = (int)rows;
  g.seg[0] = Seg{x, Cin, 0, Cin, 0};
  g.nseg = 1;
  g.W = kernel; g.ldw = Cout; g.N = Cout;
  g.e.out0 = y; g.e.ld = Cout; g.e.F = Cout;
  g.e.colscale = e;
  g.e.bias = be;
  return simt_gemm(g, EPI_PLAIN, S(stream));
}

int fwn_gated_activation(const float* f, const float* g, float* y, int64_t n, void* stream) { return gated_activation(f, g, y, n, S(stream)); }
int fwn_residual_scale(const float* x, const float* res, float* y, int64_t n, void* stream) { return residual_scale(x, res, y, n, S(stream)); }
int fwn_add(const float* a, const float* b, float* y, int64_t n, int relu, void* stream) { return fwn::add(a, b, y, n, relu != 0, S(stream)); }
int fwn_log_p(const float* z, float* out, int64_t n, void* stream) {
  FWN_CHECK(n > 0, "log_p: empty input");
  Scratch sc;
  if (sc.get(sizeof(double), S(stream))) return 1;
  return log_p(z, out, n, (double*)sc.p, S(stream));
}

}  // extern "C"


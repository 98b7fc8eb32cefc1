/*
 * flowavenet_b200.h -- C ABI of libflowavenet_b200.so (sm_100a).
 *
 * Drop-in boundary for the FloWaveNet flow pass of ryhorv/tf-flowavenet.  The reference has no
 * FFI of its own: its boundary is the Python class surface of model.py / modules.py, whose
 * arithmetic runs inside TensorFlow-1.12 kernels.  Each entry point below names the reference
 * interface (file:line under /root/reference) whose TF op sites it replaces.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / TF types.
 *   - Unless a name ends in _host, every data pointer is a DEVICE pointer owned by the caller.
 *   - Tensors are channels-last [B, T, C] float32, exactly the reference's layout.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); calls are asynchronous.
 *   - Return value: 0 on success, non-zero on error; fwn_last_error() gives a thread-local message.
 *   - No CPU fallback exists: without a CUDA device every compute entry point fails.
 */
#ifndef FLOWAVENET_B200_H_
#define FLOWAVENET_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FWN_ABI_VERSION 1

enum fwn_precision {
  FWN_FP32 = 0,        /* fp32 storage, fp32-accurate GEMMs on tcgen05 via 3-way bf16 operand split (or CUDA cores):
                          the parity mode (1e-4 rel vs oracle)                                                   */
  FWN_MIXED_BF16 = 1,  /* bf16 operands/activations, fp32 accumulate (tcgen05), fp32 flow variable x and log-det  */
  FWN_MIXED_FP16 = 2   /* fp16 operands/activations -- the reference's own mixed dtype (hparams.py:9, utils.py:3-31) --
                          fp32 accumulate (tcgen05), fp32 flow variable x and log-det; 8x finer operand rounding than bf16 */
};

/* hparams.py:6-50 / hparams8000.py -- only the fields FloWaveNet.__init__ reads (model.py:288-314). */
typedef struct fwn_config {
  int32_t n_block;            /* hparams.n_block  (8; 5 for the 8 kHz config)           */
  int32_t n_flow;             /* hparams.n_flow   (6)                                   */
  int32_t n_layer;            /* hparams.n_layer  (2) -> dilations 3^n                  */
  int32_t num_mels;           /* hparams.num_mels (80), must be even                    */
  int32_t filter_size;        /* hard-coded 256 in Block (model.py:217)                 */
  int32_t affine;             /* hparams.affine                                         */
  int32_t causal;             /* hparams.causality                                      */
  int32_t n_upsample;         /* len(hparams.upsample_scales) <= 4                      */
  int32_t upsample_scales[4]; /* [16,16] / [8,12]                                       */
  int32_t gin_channels;       /* <=0: no speaker embedding                              */
  int32_t n_speakers;
  int32_t precision;          /* enum fwn_precision                                     */
} fwn_config;

typedef struct fwn_model* fwn_handle;

const char* fwn_last_error(void);
int fwn_abi_version(void);

/* ---- model lifetime: replaces FloWaveNet.__init__ (model.py:283-314) + tf.train.Saver restore ---- */
int fwn_create(const fwn_config* cfg, fwn_handle* out);
int fwn_destroy(fwn_handle h);
/* Number of variables the reference graph owns for this config, and the i-th one's name/shape
 * (names relative to the model scope, e.g. "Block_0/Flow_1/ActNorm/logs"). */
int fwn_num_params(fwn_handle h);
int fwn_param_info(fwn_handle h, int index, const char** name, int64_t shape[4], int* rank);
/* Copy one variable in (device fp32 source) / out (device fp32 destination). */
int fwn_set_param(fwn_handle h, const char* name, const float* dev_src, int64_t numel, void* stream);
int fwn_get_param(fwn_handle h, const char* name, float* dev_dst, int64_t numel, void* stream);
/* Fold weight-norm (convolutional.py:73-83,179-188), absorb squeeze/change_order permutations into
 * the weights, cast to the compute dtype.  Must be called after the last fwn_set_param. */
int fwn_prepack(fwn_handle h, void* stream);

/* ---- whole pass ---- */
int64_t fwn_workspace_bytes(fwn_handle h, int B, int T);
/* FloWaveNet.forward (model.py:317-347).  x [B,T,1], c [B,T/hop,num_mels], g [B] int32 or NULL.
 * Writes the two scalars the reference returns; z_out (nullable) receives z as [B,T,1]
 * (= unsqueeze^n of the reference's internal `out`).  ddi != 0 performs the ActNorm data-dependent
 * initialisation pass (train.py:221,229 feeding init=True; model.py:30-41) while running. */
int fwn_forward(fwn_handle h, const float* x, const float* c, const int32_t* g, int B, int T,
                float* z_out, float* logp_out, float* logdet_out, int ddi,
                void* workspace, int64_t workspace_bytes, void* stream);
/* FloWaveNet.reverse (model.py:350-396).  z [B,T,1] -> x_out [B,T,1]. */
int fwn_reverse(fwn_handle h, const float* z, const float* c, const int32_t* g, int B, int T,
                float* x_out, void* workspace, int64_t workspace_bytes, void* stream);
/* Same, with HOST buffers (pinned or pageable): H2D of inputs, pass, D2H of results, synchronous.
 * This is the call synthesize.py:44-46 (`sess.run(predictions, feed_dict={lc: mel})`) maps to. */
int fwn_forward_host(fwn_handle h, const float* x, const float* c, const int32_t* g, int B, int T,
                     float* z_out, float* logp_out, float* logdet_out);
int fwn_reverse_host(fwn_handle h, const float* z, const float* c, const int32_t* g, int B, int T,
                     float* x_out);
/* Time-chunk sharding support: run the pass on rows [t0, t0+Tl) of an utterance of total
 * length T_total, where the caller supplies z/c for the chunk extended by `halo_l`/`halo_r` samples of
 * real neighbour data (0 at a true utterance edge).  Only the Tl interior samples are written. */
int fwn_reverse_chunk(fwn_handle h, const float* z_ext, const float* c_ext, int B, int T_ext, int halo_l,
                      int halo_r, float* x_out, void* workspace, int64_t workspace_bytes, void* stream);
/* Number of kernels the last fwn_forward / fwn_reverse on this handle launched (bench.py gpu_launches). */
int64_t fwn_last_launches(fwn_handle h);
/* Optional per-kernel-family timing with CUDA events on the launching stream.  Families (index): 0 front conv,
 * 1 gate GEMM (dilated conv + cond), 2 res/skip GEMM, 3 final conv, 4 zero-conv + ActNorm/affine, 5 upsampler.
 * fwn_profile_read sums, since the last read, elapsed ms, launch counts and algorithmic work (FLOPs for 0-4,
 * bytes for 5) per family, then resets. */
int fwn_profile_enable(fwn_handle h, int on);
int fwn_profile_read(fwn_handle h, double ms[8], int64_t launches[8], double work[8]);
int fwn_receptive_halo(fwn_handle h); /* samples of halo per side needed for exact chunked synthesis */

/* ---- training step: train.py:56-81 (loss, tf.gradients, tower average, clip, Adam), train.py:15-24 (lr) ----
 * fp32 models only.  The variable vector is flat: variable i of fwn_param_info occupies floats
 * [fwn_param_offset(i), +numel) of a buffer of fwn_param_floats(h) floats; gradients use the same layout. */
/* Build the training state (gather maps, transposed operands, Adam moments).  Call after the last fwn_set_param, instead of or
 * after fwn_prepack. */
int fwn_train_enable(fwn_handle h, void* stream);
int64_t fwn_train_workspace_bytes(fwn_handle h, int B, int T);
int64_t fwn_param_floats(fwn_handle h);
int64_t fwn_param_offset(fwn_handle h, int index);
int64_t fwn_grad_floats(fwn_handle h); /* >= fwn_param_floats: the tail is scratch of the backward pass */
int fwn_params_ptr(fwn_handle h, float** dev_ptr); /* the flat variable vector itself (device, owned by the handle) */
/* One tower of build_model (train.py:56-66): log_p, logdet = model.forward(x, c, g); grads = d(-(log_p + logdet))/d variables.
 * grads: device buffer of grad_floats >= fwn_grad_floats(h) floats, overwritten.  Variables the loss does not depend on
 * (speaker_embeddings: the reference's tf.gradients returns None for them, train.py:75 filters them out) get zero. */
int fwn_loss_and_grads(fwn_handle h, const float* x, const float* c, const int32_t* g, int B, int T, float* logp_out,
                       float* logdet_out, float* grads, int64_t grad_floats, void* workspace, int64_t workspace_bytes,
                       void* stream);
/* Gradient buckets for overlapping the tower average (utils.average_gradients, utils.py:34-60; train.py:75-77) with the backward
 * pass.  The flat gradient is produced back to front: bucket 0 = the variables of the LAST block (39 % of all parameters for
 * hparams.py), ..., bucket n_block-1 = block 0, then the upsampler variables (and the speaker embeddings when present).  Each
 * bucket is one contiguous range [offset, offset+count) of the flat gradient; the ranges tile [0, fwn_param_floats).
 * fwn_grad_bucket_wait makes `consumer_stream` wait (cudaStreamWaitEvent) until bucket k of the LAST fwn_loss_and_grads call is
 * final, so an all-reduce enqueued there runs while the backward pass of the earlier blocks is still executing. */
int fwn_grad_bucket_count(fwn_handle h);
int fwn_grad_bucket_range(fwn_handle h, int bucket, int64_t* offset, int64_t* count);
int fwn_grad_bucket_wait(fwn_handle h, int bucket, void* consumer_stream);
/* tf.global_norm of the gradient vector (train.py:29) -> device scalar */
int fwn_grad_global_norm(fwn_handle h, const float* grads, float* norm_out, void* stream);
/* clip_by_global_norm(clip_norm) + tf.train.AdamOptimizer.apply_gradients (train.py:27-31,76-81) + device-side re-pack of every
 * derived operand.  `grads` = the tower average (utils.py:34-60), computed by the caller (all-reduce).  step counts from 1. */
int fwn_apply_gradients(fwn_handle h, const float* grads, float lr, float beta1, float beta2, float eps, float clip_norm,
                        int64_t step, void* stream);
/* fp32 models run their GEMMs on the tensor cores as sums of bf16 x bf16 products of 3-way split operands (fp32 accumulate).
 * 6 terms keep every product down to 2^-24 (the parity mode, default for fwn_forward / fwn_reverse; measured 5e-6 relative
 * against an fp32 FMA chain on K = 848 -- the tensor-core accumulator truncates -- and z within 4e-6 of the float64 oracle);
 * 3 terms (a1w1 + a1w2 + a2w1) keep ~2^-16..2^-18 per product at half the tensor work (default for the training step; gradients
 * that are small differences of large sums then carry errors up to ~1e-3 of the model's largest gradient entry). */
int fwn_set_split_terms(fwn_handle h, int inference_terms, int training_terms);
/* Mixed-precision passes: how the coupling WaveNet (modules.py:113-128, 161-186) is launched.  1 = fused kernels: one per ResBlock
 * layer (gate GEMM -> tanh*sigmoid -> res|skip 1x1, the gated tile kept in shared memory), the last layer's launch also carrying the tail
 * (final 1x1 + ReLU -> ZeroConv1d -> ActNorm / affine coupling on x; a separate tail kernel where the layer is not fused); 0 = one launch per GEMM with the intermediate activations round-tripping through
 * HBM; -1 = default (fused where it pays: large launches for the layer kernel; FWN_FUSE_LAYER=0 / FWN_FUSE_TAIL=0 disable).
 * The layer kernel is bit-identical to its two launches; the tail kernel rounds the same 16-bit u. */
int fwn_set_layer_fusion(fwn_handle h, int mode);
/* Compute precision of the training step (the model keeps fp32 master variables either way, utils.py:3-31):
 *   FWN_FP32        fp32-accurate GEMMs (3-way bf16 split on the tensor cores) -- the parity mode (default)
 *   FWN_MIXED_BF16  bf16 operands and tape, fp32 accumulation / reductions / gradients: BASELINE config 5 ("bf16").  bf16 keeps
 *                   fp32's exponent range, so the reference's static loss scale (hparams.scale, train.py:62,75-77) is 1. */
int fwn_set_train_compute(fwn_handle h, int precision);
/* Per-op entry of the bf16 weight-gradient kernel (unit parity): dw[k, n] += sum_{b,t} a[b, t + shift, k] dy[b, t, n];
 * dbias[n] += sum_{b,t} dy[b, t, n] (nullable).  a [B,T,K], dy [B,T,N] bf16 (K, N multiples of 8); dw [K,N], dbias [N] fp32, accumulated.
 * Reference: tf.gradients of a Conv1D kernel / bias (modules.py:24-33; train.py:62-63). */
int fwn_wgrad_bf16(const void* a, const void* dy, float* dw, float* dbias, int B, int T, int K, int N, int shift, void* stream);
/* fp32 training, parity setting: on != 0 runs the FORWARD GEMMs of fwn_loss_and_grads on the CUDA-core engine (round-to-nearest
 * FFMA chains) and only the backward GEMMs on the split tensor-core engine.  The tensor cores' fp32 accumulator truncates; the
 * resulting bias (5e-6 on the forward activations) is amplified by the backward pass on gradients that are small differences of
 * large sums.  With it every variable's gradient stays within 2e-4 of its own max-abs against the float64 oracle on long sequences. */
int fwn_set_train_exact_forward(fwn_handle h, int on);
/* Checkpoint / resume of a training run (train.py:190,199-210: tf.train.Saver over the variables and the Adam slots).
 * which: 0 = the flat variable vector, 1 = Adam first moments, 2 = Adam second moments; numel = fwn_param_floats(h).
 * Setting the variables re-derives every packed operand on the device.  The step counter lives with the caller. */
int fwn_get_train_state(fwn_handle h, int which, float* dev_dst, int64_t numel, void* stream);
int fwn_set_train_state(fwn_handle h, int which, const float* dev_src, int64_t numel, void* stream);
/* Device-side twin of fwn_prepack (needs fwn_train_enable): re-derive all operands from the current variables. */
int fwn_repack(fwn_handle h, void* stream);

/* ---- per-op entry points (reference layout, fp32): one per TF op site of SURVEY 2.3 ---- */
/* Block.forward squeeze model.py:226-228: y[b,t,2c+k] = x[b,2t+k,c];  x [B,T,C] -> y [B,T/2,2C] */
int fwn_squeeze(const float* x, float* y, int B, int T, int C, void* stream);
/* Block.reverse unsqueeze model.py:260-262: y[b,2t+k,c] = x[b,t,2c+k]; x [B,T,C] -> y [B,2T,C/2] */
int fwn_unsqueeze(const float* x, float* y, int B, int T, int C, void* stream);
/* change_order model.py:166-174 on one tensor [rows, C]: y = concat(x[:, C/2:], x[:, :C/2]) */
int fwn_change_order(const float* x, float* y, int64_t rows, int C, void* stream);
/* ActNorm.forward model.py:86-94: y=(x+b)*exp(3 logs); *logdet_out = mean_c(3 logs) (device scalar) */
int fwn_actnorm_fwd(const float* x, const float* b, const float* logs, float* y, float* logdet_out,
                    int64_t rows, int C, void* stream);
/* ActNorm.reverse model.py:97-102: y = x*exp(-3 logs) - b */
int fwn_actnorm_rev(const float* x, const float* b, const float* logs, float* y, int64_t rows, int C,
                    void* stream);
/* ActNorm DDI values model.py:55-56,65-70: b=-mean(x), logs=log(1/(sqrt(mean((x+b)^2))+1e-7))/3 */
int fwn_actnorm_ddi(const float* x, float* b_out, float* logs_out, int64_t rows, int C, void* stream);
/* AffineCoupling elementwise part model.py:133-141 / 155-161.  x [rows,C], net [rows,C]=(log_s|t)
 * (affine) or [rows,C/2] (additive).  y=[x_a,(x_b-t)exp(-log_s)], *logdet_out=mean(-log_s)/2. */
int fwn_affine_fwd(const float* x, const float* net, float* y, float* logdet_out, int64_t rows, int C,
                   int affine, void* stream);
int fwn_affine_rev(const float* x, const float* net, float* y, int64_t rows, int C, int affine,
                   void* stream);
/* One upsampling stage, model.py:398-404 + convolutional.py:155-201: weight-normed Conv2DTranspose
 * (kernel (2s,3), strides (s,1), SAME, 1->1 channel) + bias + leaky_relu(0.4).
 * c_in [B,Tm,mels] -> c_out [B,Tm*s,mels]; kernel [2s,3], g [1], bias [1] raw variables. */
int fwn_upsample_stage(const float* c_in, const float* kernel, const float* g, const float* bias,
                       float* c_out, int B, int Tm, int mels, int s, void* stream);
/* Conv.forward modules.py:24-33 + Conv1D.build convolutional.py:53-109: zero pad, VALID dilated
 * cross-correlation with the weight-normed kernel (wn_g NULL => no weight norm, as ZeroConv1d),
 * + bias.  x [B,T,Cin] -> y [B,T,Cout]; kernel [k,Cin,Cout].  relu != 0 fuses tf.nn.relu. */
int fwn_conv1d(const float* x, const float* kernel, const float* wn_g, const float* bias, float* y,
               int B, int T, int Cin, int Cout, int ksize, int dilation, int causal, int relu,
               void* stream);
/* Mixed-precision twin of fwn_conv1d on the tcgen05 engine (weight norm already folded): x,y bf16 [B,T,C];
 * w_packed bf16 [ceil16(Cout)][ceil64(ksize*ceil16(Cin))], K index = tap*ceil16(Cin) + cin; bias fp32. */
int fwn_conv1d_bf16(const void* x, const void* w_packed, const float* bias, void* y, int B, int T, int Cin,
                    int Cout, int ksize, int dilation, int causal, int relu, void* stream);
/* ZeroConv1d.forward modules.py:51-56: (x.W + b) * exp(3 scale).  x [rows,Cin] -> y [rows,Cout] */
int fwn_zero_conv1d(const float* x, const float* kernel, const float* bias, const float* scale, float* y,
                    int64_t rows, int Cin, int Cout, void* stream);
/* ResBlock gate modules.py:124: y = tanh(f) * sigmoid(g) */
int fwn_gated_activation(const float* f, const float* g, float* y, int64_t n, void* stream);
/* ResBlock residual modules.py:128: y = (x + res) * sqrt(0.5) */
int fwn_residual_scale(const float* x, const float* res, float* y, int64_t n, void* stream);
/* y = a + b (conditioning add modules.py:117-118; skip add_n modules.py:176), optional relu */
int fwn_add(const float* a, const float* b, float* y, int64_t n, int relu, void* stream);
/* mean(0.5(-log 2pi - z^2)) model.py:343 -> device scalar */
int fwn_log_p(const float* z, float* out, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FLOWAVENET_B200_H_ */

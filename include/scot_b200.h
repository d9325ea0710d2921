/*
 * scot_b200.h — C ABI of the B200-native scOT forward/backward engine (libscot_b200.so).
 *
 * This is the drop-in boundary of the hot path named in BASELINE.json: everything behind
 * `ScOT.forward` / autograd backward of the reference (scOT/model.py:1318-1509 and the HuggingFace
 * swinv2 pieces it imports, scOT/model.py:39-47) runs inside the functions declared here.
 * The reference has no FFI of its own (it is pure Python on torch ops); the Python host module
 * poseidon_b200/scOT/model.py binds these entry points with ctypes exactly the way a maintainer of the
 * reference would.
 *
 * Conventions
 *  - every pointer is a raw CUDA device pointer unless the name ends in `_host`;
 *  - the library never allocates, frees or retains device memory: the caller (PyTorch) owns every
 *    buffer, including the workspace/arena handed to the engine;
 *  - all work is enqueued on the `stream` argument (a cudaStream_t passed as void*), no internal
 *    synchronisation, so the calls are CUDA-graph capturable;
 *  - return value 0 = ok, 1 = bad argument / unsupported configuration, 2 = CUDA error;
 *    `scot_last_error()` returns a human readable description (thread local);
 *  - no CPU fallback exists: a missing GPU is an error.
 */
#ifndef SCOT_B200_H_
#define SCOT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCOT_ABI_VERSION 2

/* ---- library ------------------------------------------------------------------------------------ */
int scot_abi_version(void);
const char* scot_last_error(void);
/* number of kernels launched by this library since process start (bench.py's `gpu_launches`) */
unsigned long long scot_launch_count(void);
/* "parity" precision (split-bf16) for the per-op entry points below: when bytes != 0 every bf16 tensor T passed to them is
 * a pair -- T itself holds hi = bf16(x), and the tensor `bytes` bytes after T holds lo = bf16(x - hi) -- GEMMs accumulate
 * A_hi B_hi + A_hi B_lo + A_lo B_hi on the tensor cores, attention runs in fp32. Thread local; 0 restores plain bf16.
 * The whole-model engine selects the mode through ScotModelDesc.precision and manages the offset itself. */
void scot_set_split_offset(size_t bytes);

/* ---- GEMM (tcgen05 / TMA) ------------------------------------------------------------------------
 * D[m,n] = sum_k A(m,k) * B(n,k), bf16 operands, fp32 accumulation.
 *   a_mn_major == 0: A(m,k) = A[m*lda + k]   (activations; reduction dimension contiguous)
 *   a_mn_major == 1: A(m,k) = A[k*lda + m]   (wgrad: token dimension is the reduction)
 *   same for B with n.
 * Replaces nn.Linear forward/backward (cuBLAS) at: HF swinv2 query/key/value/dense (:416-418,:531,
 * :571,:586), scOT/model.py:669 (merging), :725-726 (unmerging), :186-190 (ConvNeXt pw convs). */
enum {
  SCOT_EPI_BF16 = 0,        /* out0(bf16) = acc + bias                                               */
  SCOT_EPI_F32 = 1,         /* out0(f32)  = acc + bias                                               */
  SCOT_EPI_GELU = 2,        /* h = acc + bias: out0(bf16) = gelu_erf'(h) (may be NULL), out1(bf16) = gelu_erf(h) */
  SCOT_EPI_GELU_BWD = 3,    /* out0(bf16) = acc * aux(bf16) (aux = saved gelu'); colsum(f32)[n] += column sums */
  SCOT_EPI_RMW_F32 = 4,     /* out0(f32) += acc                                                      */
  SCOT_EPI_ATOMIC_F32 = 5,  /* red.add out0(f32) += acc (split reduction; wgrad)                     */
  SCOT_EPI_ADD_F32_BF16 = 6 /* out0(f32) = acc + bias + aux(f32); out1(bf16, may be NULL) = same     */
};
enum { SCOT_GEMM_TCGEN05 = 0, SCOT_GEMM_SIMT = 1 /* CUDA-core cross-check path used by the tests */ };

typedef struct ScotEpilogue {
  int mode;
  const float* bias; /* [N] or NULL */
  void* out0;
  long ld0;
  void* out1;
  long ld1;
  const void* aux;
  long ldaux;
  float* colsum; /* [N] or NULL */
} ScotEpilogue;

int scot_gemm_bf16(const void* A, long lda, int a_mn_major, const void* B, long ldb, int b_mn_major, int M, int N,
                   int K, const ScotEpilogue* epi, int impl, void* stream);

/* weight gradients of up to 4 Linear layers in one launch: dW[n_out, n_in] (f32, ld_dw) += dY[tokens, n_out]^T X[tokens, n_in]
 * (dY, X bf16). Replaces the four autograd wgrad GEMMs of a ScOTLayer (q/k/v, attention output, intermediate, output). */
typedef struct ScotWgradProblem {
  const void* dY;
  long ld_dy;
  const void* X;
  long ld_x;
  float* dW;
  long ld_dw;
  long tokens;
  int n_out, n_in;
} ScotWgradProblem;
int scot_gemm_wgrad_group(const ScotWgradProblem* problems, int n, int impl, void* stream);


/* ---- per-op entry points (used by the parity tests; the engine below calls the same launchers) ------
 * Layouts: activations are token-major [batch*res*res, C]; "f32"/"bf16" name the element type. */

/* ConditionalLayerNorm / LayerNorm forward (scOT/model.py:135-160) fused with the residual add and the
 * bf16 down-cast. y = (aw*t+ab) * zhat + (cw*t+cb) [+ residual]; aw/cw NULL => plain LayerNorm(ab, cb).
 * perm_res > 0 applies ScOTPatchUnmerging's pixel-shuffle row permutation (model.py:748-754). */
int scot_cln_fwd(const float* z, const float* residual, const float* time, const float* aw, const float* ab,
                 const float* cw, const float* cb, float* x_out, void* xb_out, void* zhat, float* rstd, long rows, int C,
                 int rows_per_sample, int perm_res, float eps, void* stream);
/* backward: dz (bf16 or f32) and atomically accumulated parameter gradients; g_bias_prev += colsum(dz) */
int scot_cln_bwd(const float* dy, const void* zhat, const float* rstd, const float* time, const float* aw,
                 const float* ab, void* dz, int dz_is_f32, float* g_aw, float* g_ab, float* g_cw, float* g_cb,
                 float* g_bias_prev, long rows, int C, int rows_per_sample, int perm_res, void* stream);

/* continuous relative position bias (HF modeling_swinv2.py:450-460,489-510), all attention layers in one launch:
 * tab2[r,h] = 16*sigmoid(mlp(coords[r]))[h]*log2(e), alpha[h] = exp(min(logit_scale[h], ln 100)).
 * Parameters are addressed by element offsets into the flat fp32 parameter (and gradient) buffer, the
 * outputs by 256-byte-unit offsets into the caller's arena. */
#define SCOT_CPB_MAX_LAYERS 64
typedef struct ScotCpbLayer {
  int w1, b1, w2, ls;            /* continuous_position_bias_mlp.0.{weight,bias}, .2.weight, logit_scale */
  int tab2, alpha;               /* forward outputs: [(2ws-1)^2, heads] and [heads] floats */
  int dtab, dalpha, dpre;        /* backward: inputs dtab/dalpha (from scot_attn_bwd), scratch dpre [(2ws-1)^2*heads];
                                  * scot_cpb_bwd also reads tab2 as left by scot_cpb_fwd (sigmoid = tab2 / (16 log2 e)) */
  short ws, heads;
} ScotCpbLayer;
typedef struct ScotCpbTable {
  int n;
  ScotCpbLayer layer[SCOT_CPB_MAX_LAYERS];
} ScotCpbTable;
int scot_cpb_fwd(const ScotCpbTable* table, const float* params, void* arena, void* stream);
int scot_cpb_bwd(const ScotCpbTable* table, const float* params, float* grads, void* arena, void* stream);
/* shifted-window cosine attention (HF:421-487 + scOT/model.py:522-559) on qkv [tokens, 3C] bf16 */
int scot_attn_fwd(const void* qkv, void* out, float* lse, const float* tab2, const float* alpha, int batch, int res, int ws,
                  int shift, int heads, int head_dim, void* stream);
/* `partial` / scot_attn_bwd_partial_bytes(): reserved (kept for ABI stability). The relative-position-bias gradient is
 * folded onto `dtab` inside the dq kernel; `partial` may be NULL. `dtab` / `dalpha` are accumulated into (+=). */
size_t scot_attn_bwd_partial_bytes(int ws, int heads, int total_windows);
int scot_attn_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, const float* tab2, const float* alpha,
                  void* dqkv, float* partial, size_t partial_bytes, float* dtab, float* dalpha, float* g_qbias,
                  float* g_vbias, int batch, int res, int ws, int shift, int heads, int head_dim, void* stream);

/* ---- glue ops around the blocks (embedding, patch merging / unmerging, ConvNeXt skips, patch recovery, loss) ----------
 * One entry point per kernel; the engine below calls the same launchers in ScOT.forward order. Index-only ops (im2col,
 * merge gather / scatter, pixel unshuffle) are exact permutations (bit-exact tests in tests/test_gpu_glue.py). */
/* out(bf16)[n] = in(f32)[n], n % 4 == 0 */
int scot_cast_f32_bf16(const float* in, void* out, long n, void* stream);
/* patch embedding as GEMM input (ScOTPatchEmbeddings, scOT/model.py:295-310): x [B,Cin,H,W] f32 ->
 * out [B*(H/ps)*(W/ps), Cin*ps*ps] bf16, k = (c, di, dj); the projection itself is scot_gemm_bf16 */
int scot_embed_im2col(const float* x, void* out, int B, int Cin, int H, int W, int ps, void* stream);
/* ScOTPatchMerging gather (model.py:694-704, order (0,0),(1,0),(0,1),(1,1)): out [B*(res/2)^2, 4C] bf16 = x (+ inp) */
int scot_merge_gather(const float* x, const float* inp, void* out, int B, int res, int C, void* stream);
/* its backward: g_out [B*res^2, C] f32 = (g_in ? g_in : 0) + dG [B*(res/2)^2, 4C] scattered back */
int scot_merge_scatter(const float* dG, const float* g_in, float* g_out, int B, int res, int C, void* stream);
/* ConvNeXtBlock pieces (model.py:198-217): depthwise 7x7 (NHWC f32, pad 3) forward / backward, layer-scale residual.
 * dwconv7_bwd: g_out = g_in + conv^T(dout), g_w += weight gradient. scale_add: out = in + gamma * z, zb = bf16(z);
 * backward: dz(bf16) = gamma * g, g_gamma += sum g*z, g_bias += sum dz. */
int scot_convnext_dwconv7_fwd(const float* x, const float* w, const float* bias, float* out, int B, int res, int C, void* stream);
int scot_convnext_dwconv7_bwd(const float* x, const float* w, const float* dout, const float* g_in, float* g_out, float* g_w,
                              int B, int res, int C, void* stream);
int scot_convnext_scale_add_fwd(const float* in, const float* z, const float* gamma, float* out, void* zb, long rows, int C,
                                void* stream);
int scot_convnext_scale_add_bwd(const float* g, const void* zb, const float* gamma, void* dz, float* g_gamma, float* g_bias,
                                long rows, int C, void* stream);
/* ScOTPatchRecovery tail (model.py:639-647): D [tokens, OC*ps*ps] f32 (ConvTranspose2d as GEMM) -> planar P [B,OC,H,W];
 * 5x5 mixing conv (+ learn_residual input, + pixel_mask overwrite: mask_mode 1 = [B,OC], 2 = [B,OC,H,W] uint8);
 * backward: dP scratch (planar f32), dD [tokens, OC*ps*ps] bf16, g_w += mixup weight gradient, g_bias += projection bias */
int scot_recovery_unshuffle(const float* D, float* P, int B, int OC, int H, int W, int ps, void* stream);
int scot_recovery_conv5_fwd(const float* P, const float* w, const float* resid, int resid_channels, const float* labels,
                            const uint8_t* mask, int mask_mode, float* pred, int B, int OC, int H, int W, void* stream);
int scot_recovery_conv5_bwd(const float* P, const float* w, const float* dpred, float* dP_scratch, void* dD, float* g_w,
                            float* g_bias, int B, int OC, int H, int W, int ps, void* stream);
/* loss (model.py:1425-1484): p = 1 | 2; slices_host = channel_slice_list_normalized_loss (host array, n_slices entries)
 * or NULL for the plain mean; sums = 20 floats of scratch (kept for backward). loss_bwd: dpred = gscale[0] * dloss/dpred
 * (+ extra), zero where the mask overwrote the prediction. */
int scot_loss_fwd(const float* pred, const float* labels, float* sums, float* loss, const int* slices_host, int n_slices, int p,
                  int B, int OC, long HW, void* stream);
int scot_loss_bwd(const float* pred, const float* labels, const float* sums, const float* gscale, const float* extra,
                  const uint8_t* mask, int mask_mode, float* dpred, const int* slices_host, int n_slices, int p, int B, int OC,
                  long HW, void* stream);

/* ---- one ScOTLayer (scOT/model.py:500-581: shifted-window cosine attention + res-post-norm + MLP) ------------------------
 * Stand-alone form of the block the engine sequences. x, y, dy, dx: [batch*res*res, C] f32.
 * `params` / `grads`: scot_layer_num_params() device pointers to fp32 tensors in the reference's state_dict order of a layer
 * (attention.self.logit_scale, continuous_position_bias_mlp.0.{weight,bias}, .2.weight, query.{weight,bias}, key.weight,
 * value.{weight,bias}, attention.output.dense.{weight,bias}, layernorm_before.*, intermediate.dense.{weight,bias},
 * output.dense.{weight,bias}, layernorm_after.*; a norm is {weight.weight, weight.bias, bias.weight, bias.bias} when
 * conditioned, {weight, bias} otherwise). The workspace (scot_layer_workspace_bytes, 256 B aligned, caller-owned) keeps the
 * parameters' flat copy and everything scot_layer_bwd needs; bwd overwrites grads[i] (NULL entries are skipped). */
typedef struct ScotLayerDesc {
  int batch, res, C, heads;
  int window;  /* effective window: min(window_size, res) */
  int shift;   /* 0 or window / 2 */
  float mlp_ratio;
  int use_conditioning;
  float layer_norm_eps;
  int precision; /* 0 bf16, 1 parity (split-bf16) */
} ScotLayerDesc;
int scot_layer_num_params(const ScotLayerDesc* desc);
size_t scot_layer_workspace_bytes(const ScotLayerDesc* desc);
int scot_layer_fwd(const ScotLayerDesc* desc, const void* const* params, const float* x, const float* time, float* y,
                   void* workspace, size_t ws_bytes, void* stream);
int scot_layer_bwd(const ScotLayerDesc* desc, void* const* grads, const float* time, const float* dy, float* dx, void* workspace,
                   size_t ws_bytes, void* stream);

/* ---- whole-model engine --------------------------------------------------------------------------
 * Mirrors ScOTConfig (scOT/model.py:66-132); replaces ScOT.forward (:1318-1509) + autograd backward. */
typedef struct ScotModelDesc {
  int image_size, patch_size, num_channels, num_out_channels, embed_dim, num_stages;
  int depths[4], num_heads[4], skip_blocks[4];
  int window_size;
  float mlp_ratio;
  int use_conditioning, learn_residual, loss_p;
  int n_slices;   /* 0: plain l1/mse; else len(channel_slice_list_normalized_loss) */
  int slices[10];
  float layer_norm_eps;
  int precision; /* 0: bf16 operands (speed mode). 1: "parity" mode -- split-bf16 operands (3 tensor-core passes per GEMM),
                  * fp32 attention: the reference trains in fp32 (scOT/train.py:311) and this mode meets 1e-3 against it.
                  * The workspace doubles (every bf16 tensor gets its lo twin). */
} ScotModelDesc;

typedef struct ScotEngine ScotEngine;

int scot_engine_create(const ScotModelDesc* desc, int batch, ScotEngine** out);
void scot_engine_destroy(ScotEngine* e);
/* parameter table: names are the reference's state_dict keys; offsets index one flat fp32 buffer of
 * scot_engine_param_elems() elements that the caller allocates (zero filled) and the parameters view. */
long scot_engine_num_params(const ScotEngine* e);
long scot_engine_param_elems(const ScotEngine* e);
int scot_engine_param_info(const ScotEngine* e, long i, char* name, int name_cap, long* offset, long* numel, int* ndim,
                           long* shape4);
size_t scot_engine_workspace_bytes(const ScotEngine* e);
/* forward: pixel_values [B,Cin,H,W] f32, time [B] f32 or NULL, labels [B,Cout,H,W] f32 or NULL,
 * mask: uint8 [B,Cout] (mask_mode 1) or [B,Cout,H,W] (mask_mode 2) or NULL (0); pred_out [B,Cout,H,W],
 * loss_out [1]. The arena (scot_engine_workspace_bytes, 256 B aligned) keeps the activations for backward. */
int scot_engine_forward(ScotEngine* e, const float* params, void* arena, const float* pixel_values, const float* time,
                        const float* labels, const uint8_t* mask, int mask_mode, float* pred_out, float* loss_out,
                        int gemm_impl, void* stream);
/* backward of the last forward: grads (same layout as params) += d(grad_loss*loss + <grad_pred, pred>)/dparams.
 * grad_loss: device scalar or NULL; grad_pred: [B,Cout,H,W] or NULL. Gradients are accumulated, never zeroed. */
int scot_engine_backward(ScotEngine* e, const float* params, float* grads, void* arena, const float* grad_loss,
                         const float* grad_pred, int gemm_impl, void* stream);
/* The backward pass in two parts, for overlapping the data-parallel gradient exchange with compute: part 1 (loss, patch
 * recovery, decoder, ConvNeXt skips, deepest encoder stage) leaves the gradients of flat elements
 * [scot_engine_grad_split(), scot_engine_param_elems()) final — ~85 % of the bytes of the shipped models —, part 2 (remaining
 * encoder stages, embeddings) the rest. part 0 = both (== scot_engine_backward). */
int scot_engine_backward_part(ScotEngine* e, const float* params, float* grads, void* arena, const float* grad_loss,
                              const float* grad_pred, int gemm_impl, int part, void* stream);
long scot_engine_grad_split(const ScotEngine* e);
/* ---- fused optimizer step on the flat buffers (replaces accelerate clip_grad_norm_ + torch.optim.AdamW,
 * scOT/train.py:286, scOT/trainer.py:295-445) ----------------------------------------------------------------------
 * out[0] = sum(grads^2) (cleared first). */
int scot_grad_sq_norm(const float* grads, long n_elems, float* out, void* stream);
/* AdamW (decoupled weight decay, bias correction; torch.optim.AdamW semantics) over the whole flat buffer.
 * chunk_group[i/64] = parameter group of elements [64*(i/64), +64) or 255 (padding / frozen: untouched).
 * group_hp[g*8 ..] = {lr, weight_decay, beta1, beta2, eps, 1-beta1^t, sqrt(1-beta2^t), 0} (device memory, so that a
 * captured CUDA graph follows a learning-rate schedule). grad_sq_norm (nullable) + max_norm > 0: gradients are scaled
 * by min(1, max_norm / (sqrt(grad_sq_norm)*grad_scale + 1e-6)) on the fly (clip_grad_norm_); grad_scale multiplies every
 * gradient first (1/world_size after a summed all-reduce). params_bf16 (nullable): bf16 copy refreshed in the same pass. */
int scot_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, void* params_bf16,
                    const uint8_t* chunk_group, long n_elems, const float* group_hp, int n_groups, const float* grad_sq_norm,
                    float max_norm, float grad_scale, void* stream);
/* Evaluation metric reductions (scOT/metrics.py:12-36): out[(b*C + c)*2 + {0,1}] = sum_pixels |pred - labels|^p, |labels|^p
 * for every (sample, channel) plane of `hw` pixels; `planes` = B*C. */
int scot_lp_plane_sums(const float* pred, const float* labels, float* out, int p, long planes, long hw, void* stream);
/* Re-binds the tensors the next scot_engine_backward reads (the inputs / prediction of a forward that was replayed
 * from a CUDA graph, where the host-side bookkeeping of scot_engine_forward did not run). Host-only, no launch. */
int scot_engine_bind_io(ScotEngine* e, const float* pixel_values, const float* time, const float* labels,
                        const uint8_t* mask, int mask_mode, float* pred);

#ifdef __cplusplus
}
#endif
#endif /* SCOT_B200_H_ */

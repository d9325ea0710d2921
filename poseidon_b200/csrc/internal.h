// Cross-translation-unit declarations of the launch functions behind the C ABI.
#pragma once
#include <cuda_runtime.h>
#include "scot_b200.h"

int scot_gemm_launch(const void* A, long lda, int a_mn_major, const void* B, long ldb, int b_mn_major, int M, int N,
                     int K, const ScotEpilogue* e, int impl, cudaStream_t stream);
int scot_gemm_wgrad_group_launch(const ScotWgradProblem* probs, int n, int impl, cudaStream_t stream);

// norm.cu
int scot_cln_fwd_launch(const float* z, const float* residual, const float* time, const float* aw, const float* ab,
                        const float* cw, const float* cb, float* x_out, void* xb_out, void* zhat, float* rstd,
                        long rows, int C, int rows_per_sample, int perm_res, float eps, cudaStream_t st);
int scot_cln_bwd_launch(const float* dy, const void* zhat, const float* rstd, const float* time, const float* aw,
                        const float* ab, void* dz, int dz_is_f32, float* g_aw, float* g_ab, float* g_cw, float* g_cb,
                        float* g_bias_prev, long rows, int C, int rows_per_sample, int perm_res, cudaStream_t st);
// attention.cu
size_t scot_attn_bwd_partial_bytes(int ws, int heads, int total_windows);
int scot_cpb_fwd_launch(const ScotCpbTable* tab, const float* params, void* arena, cudaStream_t st);
int scot_cpb_bwd_launch(const ScotCpbTable* tab, const float* params, float* grads, void* arena, cudaStream_t st);
int scot_attn_fwd_launch(const void* qkv, void* out, float* lse, const float* tab2, const float* alpha, int batch,
                         int res, int ws, int shift, int heads, int hd, cudaStream_t st);
int scot_attn_bwd_launch(const void* qkv, const void* o, const void* d_o, const float* lse, const float* tab2,
                         const float* alpha, void* dqkv, float* partial, size_t partial_bytes, float* dtab, float* dalpha,
                         float* g_qbias, float* g_vbias, int batch, int res, int ws, int shift, int heads, int hd,
                         cudaStream_t st);
// same, with the dk/dv kernel forked onto a second stream (events are recorded / waited inside; capturable)
struct ScotAttnBwdFork {
  cudaStream_t stream;
  cudaEvent_t fork, join;
};
int scot_attn_bwd_launch2(const void* qkv, const void* o, const void* d_o, const float* lse, const float* tab2,
                          const float* alpha, void* dqkv, float* partial, size_t partial_bytes, float* dtab, float* dalpha,
                          float* g_qbias, float* g_vbias, int batch, int res, int ws, int shift, int heads, int hd,
                          cudaStream_t st, const ScotAttnBwdFork* fk);
// attention_tc.cu: tcgen05 / TMEM / TMA kernels for 16 x 16 windows
int scot_attn_tc_fwd_launch(const void* qkv, void* out, float* lse, const float* tab2, const float* alpha, int batch, int res,
                            int shift, int heads, int hd, cudaStream_t st);
// returns -1 when the shape is not covered (head_dim 64): the caller falls back to the mma.sync kernels
int scot_attn_tc_bwd_launch(const void* qkv, const void* o, const void* d_o, const float* lse, const float* tab2,
                            const float* alpha, void* dqkv, float* dtab, float* dalpha, float* g_qbias, float* g_vbias,
                            int batch, int res, int shift, int heads, int hd, cudaStream_t st);
// attention_f32.cu: fp32 CUDA-core attention of the "parity" precision mode (operands are hi + lo bf16 pairs, the lo
// tensors `lo_off` bytes after the hi ones; lo_off = 0 reads / writes plain bf16)
int scot_attn32_fwd_launch(const void* qkv, void* out, float* lse, const float* tab2, const float* alpha, int batch, int res,
                           int ws, int shift, int heads, int hd, size_t lo_off, cudaStream_t st);
int scot_attn32_bwd_launch(const void* qkv, const void* o, const void* d_o, const float* lse, const float* tab2,
                           const float* alpha, void* dqkv, float* dtab, float* dalpha, float* g_qbias, float* g_vbias,
                           int batch, int res, int ws, int shift, int heads, int hd, size_t lo_off, cudaStream_t st);
// misc.cu
int scot_cast_f32_bf16_launch(const float* in, void* out, long n, cudaStream_t st);
int scot_expand_bias_launch(const float* bias, float* out, int n, int rep, cudaStream_t st);
int scot_im2col_patch_launch(const float* x, void* out, int B, int Cin, int H, int W, int ps, cudaStream_t st);
int scot_merge_gather_launch(const float* x, const float* inp, void* out, int B, int res, int C, cudaStream_t st);
int scot_merge_scatter_launch(const float* dG, const float* g_in, float* g_out, int B, int res, int C, cudaStream_t st);
int scot_scale_add_fwd_launch(const float* in, const float* z, const float* gamma, float* out, void* zb, long rows, int C,
                              cudaStream_t st);
int scot_scale_add_bwd_launch(const float* g, const void* zb, const float* gamma, void* dz, float* g_gamma, float* g_bias,
                              long rows, int C, cudaStream_t st);
int scot_dwconv7_fwd_launch(const float* x, const float* w, const float* bias, float* out, int B, int res, int C,
                            cudaStream_t st);
int scot_dwconv7_bwd_launch(const float* x, const float* w, const float* dout, const float* g_in, float* g_out, float* g_w,
                            int B, int res, int C, cudaStream_t st);
int scot_unshuffle_launch(const float* D, float* P, int B, int OC, int H, int W, int ps, cudaStream_t st);
int scot_conv5_fwd_launch(const float* P, const float* w, const float* resid, int resid_channels, const float* labels,
                          const uint8_t* mask, int mask_mode, float* pred, int B, int OC, int H, int W, cudaStream_t st);
int scot_conv5_bwd_launch(const float* P, const float* w, const float* dpred, float* dP_scratch, void* dD, float* g_w,
                          float* g_bias, int B, int OC, int H, int W, int ps, cudaStream_t st);
int scot_loss_fwd_launch(const float* pred, const float* labels, float* sums, float* loss, const int* slices_host,
                         int n_slices, int p, int B, int OC, long HW, cudaStream_t st);
int scot_loss_bwd_launch(const float* pred, const float* labels, const float* sums, const float* gscale, const float* extra,
                         const uint8_t* mask, int mask_mode, float* dpred, const int* slices_host, int n_slices, int p, int B,
                         int OC, long HW, cudaStream_t st);

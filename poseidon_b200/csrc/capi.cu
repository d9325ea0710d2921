// extern "C" surface of libscot_b200.so (see include/scot_b200.h) + error / launch-count plumbing.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "internal.h"
#include "scot_b200.h"

static thread_local char g_err[1024] = "";
static std::atomic<unsigned long long> g_launches{0};

void scot_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void scot_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
static thread_local size_t g_split_off = 0;
size_t scot_split_off() { return g_split_off; }
void scot_set_split_off(size_t bytes) { g_split_off = bytes; }
bool scot_pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SCOT_PDL");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

extern "C" {

int scot_abi_version(void) { return SCOT_ABI_VERSION; }
const char* scot_last_error(void) { return g_err; }
unsigned long long scot_launch_count(void) { return g_launches.load(); }
void scot_set_split_offset(size_t bytes) { scot_set_split_off(bytes); }

int scot_gemm_bf16(const void* A, long lda, int a_mn_major, const void* B, long ldb, int b_mn_major, int M, int N,
                   int K, const ScotEpilogue* epi, int impl, void* stream) {
  return scot_gemm_launch(A, lda, a_mn_major, B, ldb, b_mn_major, M, N, K, epi, impl, (cudaStream_t)stream);
}

int scot_gemm_wgrad_group(const ScotWgradProblem* problems, int n, int impl, void* stream) {
  return scot_gemm_wgrad_group_launch(problems, n, impl, (cudaStream_t)stream);
}
int scot_cln_fwd(const float* z, const float* residual, const float* time, const float* aw, const float* ab,
                 const float* cw, const float* cb, float* x_out, void* xb_out, void* zhat, float* rstd, long rows, int C,
                 int rows_per_sample, int perm_res, float eps, void* stream) {
  return scot_cln_fwd_launch(z, residual, time, aw, ab, cw, cb, x_out, xb_out, zhat, rstd, rows, C, rows_per_sample,
                             perm_res, eps, (cudaStream_t)stream);
}
int scot_cln_bwd(const float* dy, const void* zhat, const float* rstd, const float* time, const float* aw,
                 const float* ab, void* dz, int dz_is_f32, float* g_aw, float* g_ab, float* g_cw, float* g_cb,
                 float* g_bias_prev, long rows, int C, int rows_per_sample, int perm_res, void* stream) {
  return scot_cln_bwd_launch(dy, zhat, rstd, time, aw, ab, dz, dz_is_f32, g_aw, g_ab, g_cw, g_cb, g_bias_prev, rows, C,
                             rows_per_sample, perm_res, (cudaStream_t)stream);
}
int scot_cpb_fwd(const ScotCpbTable* table, const float* params, void* arena, void* stream) {
  return scot_cpb_fwd_launch(table, params, arena, (cudaStream_t)stream);
}
int scot_cpb_bwd(const ScotCpbTable* table, const float* params, float* grads, void* arena, void* stream) {
  return scot_cpb_bwd_launch(table, params, grads, arena, (cudaStream_t)stream);
}
int scot_attn_fwd(const void* qkv, void* out, float* lse, const float* tab2, const float* alpha, int batch, int res, int ws,
                  int shift, int heads, int head_dim, void* stream) {
  return scot_attn_fwd_launch(qkv, out, lse, tab2, alpha, batch, res, ws, shift, heads, head_dim, (cudaStream_t)stream);
}
int scot_attn_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, const float* tab2, const float* alpha,
                  void* dqkv, float* partial, size_t partial_bytes, float* dtab, float* dalpha, float* g_qbias,
                  float* g_vbias, int batch, int res, int ws, int shift, int heads, int head_dim, void* stream) {
  return scot_attn_bwd_launch(qkv, o, d_o, lse, tab2, alpha, dqkv, partial, partial_bytes, dtab, dalpha, g_qbias, g_vbias,
                              batch, res, ws, shift, heads, head_dim, (cudaStream_t)stream);
}

int scot_cast_f32_bf16(const float* in, void* out, long n, void* stream) {
  return scot_cast_f32_bf16_launch(in, out, n, (cudaStream_t)stream);
}
int scot_embed_im2col(const float* x, void* out, int B, int Cin, int H, int W, int ps, void* stream) {
  return scot_im2col_patch_launch(x, out, B, Cin, H, W, ps, (cudaStream_t)stream);
}
int scot_merge_gather(const float* x, const float* inp, void* out, int B, int res, int C, void* stream) {
  return scot_merge_gather_launch(x, inp, out, B, res, C, (cudaStream_t)stream);
}
int scot_merge_scatter(const float* dG, const float* g_in, float* g_out, int B, int res, int C, void* stream) {
  return scot_merge_scatter_launch(dG, g_in, g_out, B, res, C, (cudaStream_t)stream);
}
int scot_convnext_dwconv7_fwd(const float* x, const float* w, const float* bias, float* out, int B, int res, int C, void* stream) {
  return scot_dwconv7_fwd_launch(x, w, bias, out, B, res, C, (cudaStream_t)stream);
}
int scot_convnext_dwconv7_bwd(const float* x, const float* w, const float* dout, const float* g_in, float* g_out, float* g_w,
                              int B, int res, int C, void* stream) {
  return scot_dwconv7_bwd_launch(x, w, dout, g_in, g_out, g_w, B, res, C, (cudaStream_t)stream);
}
int scot_convnext_scale_add_fwd(const float* in, const float* z, const float* gamma, float* out, void* zb, long rows, int C,
                                void* stream) {
  return scot_scale_add_fwd_launch(in, z, gamma, out, zb, rows, C, (cudaStream_t)stream);
}
int scot_convnext_scale_add_bwd(const float* g, const void* zb, const float* gamma, void* dz, float* g_gamma, float* g_bias,
                                long rows, int C, void* stream) {
  return scot_scale_add_bwd_launch(g, zb, gamma, dz, g_gamma, g_bias, rows, C, (cudaStream_t)stream);
}
int scot_recovery_unshuffle(const float* D, float* P, int B, int OC, int H, int W, int ps, void* stream) {
  return scot_unshuffle_launch(D, P, B, OC, H, W, ps, (cudaStream_t)stream);
}
int scot_recovery_conv5_fwd(const float* P, const float* w, const float* resid, int resid_channels, const float* labels,
                            const uint8_t* mask, int mask_mode, float* pred, int B, int OC, int H, int W, void* stream) {
  return scot_conv5_fwd_launch(P, w, resid, resid_channels, labels, mask, mask_mode, pred, B, OC, H, W, (cudaStream_t)stream);
}
int scot_recovery_conv5_bwd(const float* P, const float* w, const float* dpred, float* dP_scratch, void* dD, float* g_w,
                            float* g_bias, int B, int OC, int H, int W, int ps, void* stream) {
  return scot_conv5_bwd_launch(P, w, dpred, dP_scratch, dD, g_w, g_bias, B, OC, H, W, ps, (cudaStream_t)stream);
}
int scot_loss_fwd(const float* pred, const float* labels, float* sums, float* loss, const int* slices_host, int n_slices, int p,
                  int B, int OC, long HW, void* stream) {
  return scot_loss_fwd_launch(pred, labels, sums, loss, slices_host, n_slices, p, B, OC, HW, (cudaStream_t)stream);
}
int scot_loss_bwd(const float* pred, const float* labels, const float* sums, const float* gscale, const float* extra,
                  const uint8_t* mask, int mask_mode, float* dpred, const int* slices_host, int n_slices, int p, int B, int OC,
                  long HW, void* stream) {
  return scot_loss_bwd_launch(pred, labels, sums, gscale, extra, mask, mask_mode, dpred, slices_host, n_slices, p, B, OC, HW,
                              (cudaStream_t)stream);
}

}  // extern "C"

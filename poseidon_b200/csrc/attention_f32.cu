// fp32 CUDA-core window attention for the "parity" precision mode (split-bf16 operands, see common.cuh).
//
// Same arithmetic as attention.cu (Swinv2SelfAttention.forward, HF modeling_swinv2.py:421-487 + the cyclic shift /
// window partition of scOT/model.py:522-559), but every product is an fp32 FMA and P never leaves fp32: this is the
// attention path of `precision="parity"`, whose contract is the north-star tolerance (1e-3 relative L2 against the
// fp32/fp64 reference), not speed. One CTA per (window, head), one thread per token; q/k/v/o/dO arrive as hi + lo
// bf16 pairs. It is also the on-device cross-check of the tensor-core kernels in the tests.
#include "common.cuh"
#include "internal.h"

namespace {

constexpr float kLog2e = 1.4426950408889634f;

struct Geo32 {
  int res, shift, nws, heads, C, ws;
};

__device__ __forceinline__ long token_row32(const Geo32& g, int bw, int n) {
  const int nw = g.nws * g.nws;
  const int b = bw / nw, w = bw - b * nw;
  const int wi = w / g.nws, wj = w - wi * g.nws;
  int i = wi * g.ws + n / g.ws + g.shift;
  int j = wj * g.ws + n % g.ws + g.shift;
  if (i >= g.res) i -= g.res;
  if (j >= g.res) j -= g.res;
  return ((long)b * g.res + i) * g.res + j;
}
// region code of token n (scOT/model.py:448-472): bit0 = lower part of a last-row window, bit1 = right part of a
// last-column window of the shifted image
__device__ __forceinline__ int region32(const Geo32& g, int bw, int n) {
  if (g.shift == 0) return 0;
  const int w = bw % (g.nws * g.nws);
  const int wi = w / g.nws, wj = w - wi * g.nws;
  const int hm = (wi == g.nws - 1) && (n / g.ws >= g.ws - g.shift);
  const int wm = (wj == g.nws - 1) && (n % g.ws >= g.ws - g.shift);
  return hm | (wm << 1);
}
__device__ __forceinline__ int rowbase32(int ws, int m) {
  return (m / ws) * (2 * ws - 1) + (m % ws) + (ws - 1) * (2 * ws - 1) + (ws - 1);
}
__device__ __forceinline__ int coloff32(int ws, int n) { return (n / ws) * (2 * ws - 1) + (n % ws); }

template <int HD>
__device__ __forceinline__ void load_row(float* v, const bf16* p, size_t lo_off) {
#pragma unroll
  for (int c = 0; c < HD / 4; ++c) {
    const float4 t = ld_bf16x4(p + 4 * c, lo_off);
    v[4 * c] = t.x; v[4 * c + 1] = t.y; v[4 * c + 2] = t.z; v[4 * c + 3] = t.w;
  }
}
template <int HD>
__device__ __forceinline__ void store_row(bf16* p, size_t lo_off, const float* v) {
#pragma unroll
  for (int c = 0; c < HD / 4; ++c) st_bf16x4(p + 4 * c, lo_off, v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}
template <int HD>
__device__ __forceinline__ float normalize(float* v) {  // F.normalize, eps 1e-12 (HF:445); returns 1 / max(|v|, eps)
  float ss = 0.f;
#pragma unroll
  for (int c = 0; c < HD; ++c) ss = fmaf(v[c], v[c], ss);
  const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
  for (int c = 0; c < HD; ++c) v[c] *= inv;
  return inv;
}
template <int HD>
__device__ __forceinline__ float dot_s(const float* a, const float* srow) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < HD / 4; ++c) {
    const float4 t = *reinterpret_cast<const float4*>(srow + 4 * c);
    s = fmaf(a[4 * c], t.x, s); s = fmaf(a[4 * c + 1], t.y, s); s = fmaf(a[4 * c + 2], t.z, s); s = fmaf(a[4 * c + 3], t.w, s);
  }
  return s;
}
template <int HD>
__device__ __forceinline__ void axpy_s(float* acc, float a, const float* srow) {
#pragma unroll
  for (int c = 0; c < HD / 4; ++c) {
    const float4 t = *reinterpret_cast<const float4*>(srow + 4 * c);
    acc[4 * c] = fmaf(a, t.x, acc[4 * c]); acc[4 * c + 1] = fmaf(a, t.y, acc[4 * c + 1]);
    acc[4 * c + 2] = fmaf(a, t.z, acc[4 * c + 2]); acc[4 * c + 3] = fmaf(a, t.w, acc[4 * c + 3]);
  }
}

template <int HD>
__global__ void __launch_bounds__(256)
attn32_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, float* __restrict__ lse,
                  const float* __restrict__ tab2, const float* __restrict__ alpha, Geo32 g, size_t lo_off) {
  extern __shared__ __align__(16) float sm32[];
  const int N = g.ws * g.ws, TABN = (2 * g.ws - 1) * (2 * g.ws - 1);
  float* sk = sm32;            // [N][HD] normalised keys
  float* sv = sk + N * HD;     // [N][HD]
  float* stab = sv + N * HD;   // [TABN]
  const int unit = blockIdx.x, bw = unit / g.heads, h = unit - bw * g.heads;
  const int m = threadIdx.x;
  const long ld = 3L * g.C;
  float q[HD];
  long tr = 0;
  if (m < N) {
    tr = token_row32(g, bw, m);
    float t[HD];
    load_row<HD>(t, qkv + tr * ld + g.C + h * HD, lo_off);
    normalize<HD>(t);
#pragma unroll
    for (int c = 0; c < HD; ++c) sk[m * HD + c] = t[c];
    load_row<HD>(t, qkv + tr * ld + 2 * g.C + h * HD, lo_off);
#pragma unroll
    for (int c = 0; c < HD; ++c) sv[m * HD + c] = t[c];
    load_row<HD>(q, qkv + tr * ld + h * HD, lo_off);
    normalize<HD>(q);
  }
  for (int i = threadIdx.x; i < TABN; i += blockDim.x) stab[i] = tab2[i * g.heads + h];
  __syncthreads();
  if (m >= N) return;
  const float a2 = alpha[h] * kLog2e;
  const int rb = rowbase32(g.ws, m), code = region32(g, bw, m);
  float mx = -INFINITY;
  for (int n = 0; n < N; ++n) {
    float v = fmaf(dot_s<HD>(q, sk + n * HD), a2, stab[rb - coloff32(g.ws, n)]);
    if (g.shift && region32(g, bw, n) != code) v -= 200.0f * kLog2e;  // mask {0,-100} added twice (HF:465-468)
    mx = fmaxf(mx, v);
  }
  float l = 0.f, o[HD];
#pragma unroll
  for (int c = 0; c < HD; ++c) o[c] = 0.f;
  for (int n = 0; n < N; ++n) {
    float v = fmaf(dot_s<HD>(q, sk + n * HD), a2, stab[rb - coloff32(g.ws, n)]);
    if (g.shift && region32(g, bw, n) != code) v -= 200.0f * kLog2e;
    const float p = exp2f(v - mx);
    l += p;
    axpy_s<HD>(o, p, sv + n * HD);
  }
  const float il = 1.0f / l;
#pragma unroll
  for (int c = 0; c < HD; ++c) o[c] *= il;
  store_row<HD>(out + tr * g.C + h * HD, lo_off, o);
  lse[(long)unit * N + m] = mx + log2f(l);
}

// backward: phase 1 = every thread is a query row (dq, bias-table / logit-scale gradients),
// phase 2 = every thread is a key row (dk, dv); P is recomputed from the saved row log-sum-exp.
template <int HD>
__global__ void __launch_bounds__(256)
attn32_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ o, const bf16* __restrict__ d_o,
                  const float* __restrict__ lse, const float* __restrict__ tab2, const float* __restrict__ alpha,
                  bf16* __restrict__ dqkv, float* __restrict__ dtab, float* __restrict__ dalpha, float* __restrict__ g_qbias,
                  float* __restrict__ g_vbias, Geo32 g, size_t lo_off) {
  extern __shared__ __align__(16) float sm32[];
  const int N = g.ws * g.ws, TABN = (2 * g.ws - 1) * (2 * g.ws - 1);
  float* sa = sm32;               // phase 1: normalised keys   | phase 2: normalised queries
  float* sb = sa + N * HD;        // phase 1: values            | phase 2: dO rows
  float* stab = sb + N * HD;      // [TABN]
  float* sdtab = stab + TABN;     // [TABN] bias-table gradient of this (window, head)
  float* slse = sdtab + TABN;     // [N]
  float* sdelta = slse + N;       // [N]
  float* sred = sdelta + N;       // [2*HD + 1]: column sums of dq, dv; logit-scale gradient
  const int unit = blockIdx.x, bw = unit / g.heads, h = unit - bw * g.heads;
  const int t = threadIdx.x;
  const long ld = 3L * g.C;
  const bool act = t < N;
  long tr = 0;
  float qh[HD], dor[HD];
  float inv_q = 0.f, my_lse = 0.f, delta = 0.f;
  if (act) {
    tr = token_row32(g, bw, t);
    float tmp[HD];
    load_row<HD>(tmp, qkv + tr * ld + g.C + h * HD, lo_off);
    normalize<HD>(tmp);
#pragma unroll
    for (int c = 0; c < HD; ++c) sa[t * HD + c] = tmp[c];
    load_row<HD>(tmp, qkv + tr * ld + 2 * g.C + h * HD, lo_off);
#pragma unroll
    for (int c = 0; c < HD; ++c) sb[t * HD + c] = tmp[c];
    load_row<HD>(qh, qkv + tr * ld + h * HD, lo_off);
    inv_q = normalize<HD>(qh);
    load_row<HD>(dor, d_o + tr * g.C + h * HD, lo_off);
    load_row<HD>(tmp, o + tr * g.C + h * HD, lo_off);
#pragma unroll
    for (int c = 0; c < HD; ++c) delta = fmaf(dor[c], tmp[c], delta);
    my_lse = lse[(long)unit * N + t];
  }
  for (int i = t; i < TABN; i += blockDim.x) {
    stab[i] = tab2[i * g.heads + h];
    sdtab[i] = 0.f;
  }
  for (int i = t; i < 2 * HD + 1; i += blockDim.x) sred[i] = 0.f;
  __syncthreads();
  const float al = alpha[h], a2 = al * kLog2e;
  const int code = act ? region32(g, bw, t) : 0;
  if (act) {
    const int rb = rowbase32(g.ws, t);
    float dq[HD];
#pragma unroll
    for (int c = 0; c < HD; ++c) dq[c] = 0.f;
    float dal = 0.f;
    for (int n = 0; n < N; ++n) {
      const float s = dot_s<HD>(qh, sa + n * HD);
      const int idx = rb - coloff32(g.ws, n);
      float v = fmaf(s, a2, stab[idx]);
      if (g.shift && region32(g, bw, n) != code) v -= 200.0f * kLog2e;
      const float p = exp2f(v - my_lse);
      const float dp = dot_s<HD>(dor, sb + n * HD);
      const float ds = p * (dp - delta);
      axpy_s<HD>(dq, ds, sa + n * HD);
      atomicAdd(&sdtab[idx], ds);
      dal = fmaf(ds, s, dal);
    }
    // dq_hat = alpha * dq ; dq = (dq_hat - q_hat (q_hat . dq_hat)) / max(|q|, eps)
    float proj = 0.f;
#pragma unroll
    for (int c = 0; c < HD; ++c) {
      dq[c] *= al;
      proj = fmaf(qh[c], dq[c], proj);
    }
#pragma unroll
    for (int c = 0; c < HD; ++c) dq[c] = (dq[c] - qh[c] * proj) * inv_q;
    store_row<HD>(dqkv + tr * ld + h * HD, lo_off, dq);
#pragma unroll
    for (int c = 0; c < HD; ++c) atomicAdd(&sred[c], dq[c]);
    atomicAdd(&sred[2 * HD], dal);
  }
  __syncthreads();
  // phase 2: swap roles. Own key / value rows go to registers, smem takes the normalised queries and dO rows.
  float kh[HD], vr[HD];
  float inv_k = 0.f;
  if (act) {
#pragma unroll
    for (int c = 0; c < HD; ++c) {
      kh[c] = sa[t * HD + c];
      vr[c] = sb[t * HD + c];
    }
    float tmp[HD];
    load_row<HD>(tmp, qkv + tr * ld + g.C + h * HD, lo_off);
    inv_k = normalize<HD>(tmp);
  }
  __syncthreads();
  if (act) {
#pragma unroll
    for (int c = 0; c < HD; ++c) {
      sa[t * HD + c] = qh[c];
      sb[t * HD + c] = dor[c];
    }
    slse[t] = my_lse;
    sdelta[t] = delta;
  }
  __syncthreads();
  if (act) {
    const int co = coloff32(g.ws, t);
    float dk[HD], dv[HD];
#pragma unroll
    for (int c = 0; c < HD; ++c) dk[c] = dv[c] = 0.f;
    for (int m = 0; m < N; ++m) {
      const float s = dot_s<HD>(kh, sa + m * HD);
      float v = fmaf(s, a2, stab[rowbase32(g.ws, m) - co]);
      if (g.shift && region32(g, bw, m) != code) v -= 200.0f * kLog2e;
      const float p = exp2f(v - slse[m]);
      const float dp = dot_s<HD>(vr, sb + m * HD);
      const float ds = p * (dp - sdelta[m]);
      axpy_s<HD>(dv, p, sb + m * HD);
      axpy_s<HD>(dk, ds, sa + m * HD);
    }
    float proj = 0.f;
#pragma unroll
    for (int c = 0; c < HD; ++c) {
      dk[c] *= al;
      proj = fmaf(kh[c], dk[c], proj);
    }
#pragma unroll
    for (int c = 0; c < HD; ++c) dk[c] = (dk[c] - kh[c] * proj) * inv_k;
    store_row<HD>(dqkv + tr * ld + g.C + h * HD, lo_off, dk);
    store_row<HD>(dqkv + tr * ld + 2 * g.C + h * HD, lo_off, dv);
#pragma unroll
    for (int c = 0; c < HD; ++c) atomicAdd(&sred[HD + c], dv[c]);
  }
  __syncthreads();
  for (int i = t; i < TABN; i += blockDim.x) atomicAdd(dtab + i * g.heads + h, sdtab[i]);
  for (int i = t; i < HD; i += blockDim.x) {
    if (g_qbias != nullptr) atomicAdd(g_qbias + h * HD + i, sred[i]);
    if (g_vbias != nullptr) atomicAdd(g_vbias + h * HD + i, sred[HD + i]);
  }
  if (t == 0) atomicAdd(dalpha + h, sred[2 * HD]);
}

template <int HD>
int launch32(bool bwd, const void* qkv, const void* o, const void* d_o, float* lse, const float* tab2, const float* alpha,
             void* out_or_dqkv, float* dtab, float* dalpha, float* g_qbias, float* g_vbias, Geo32 g, int units,
             size_t lo_off, cudaStream_t st) {
  const int N = g.ws * g.ws, TABN = (2 * g.ws - 1) * (2 * g.ws - 1);
  const int threads = N < 32 ? 32 : N;
  if (!bwd) {
    const size_t smem = ((size_t)2 * N * HD + TABN) * sizeof(float);
    SCOT_CHECK_CUDA(cudaFuncSetAttribute(attn32_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attn32_fwd_kernel<HD><<<units, threads, smem, st>>>((const bf16*)qkv, (bf16*)out_or_dqkv, lse, tab2, alpha, g, lo_off);
  } else {
    const size_t smem = ((size_t)2 * N * HD + 2 * TABN + 2 * N + 2 * HD + 1) * sizeof(float);
    SCOT_CHECK_CUDA(cudaFuncSetAttribute(attn32_bwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attn32_bwd_kernel<HD><<<units, threads, smem, st>>>((const bf16*)qkv, (const bf16*)o, (const bf16*)d_o, lse, tab2, alpha,
                                                         (bf16*)out_or_dqkv, dtab, dalpha, g_qbias, g_vbias, g, lo_off);
  }
  SCOT_LAUNCH_CHECK();
  return 0;
}

int dispatch32(bool bwd, const void* qkv, const void* o, const void* d_o, float* lse, const float* tab2, const float* alpha,
               void* out_or_dqkv, float* dtab, float* dalpha, float* g_qbias, float* g_vbias, int batch, int res, int ws,
               int shift, int heads, int hd, size_t lo_off, cudaStream_t st) {
  SCOT_REQUIRE(ws >= 1 && ws <= 16 && res % ws == 0, "attention (fp32): window %d must divide the resolution %d", ws, res);
  SCOT_REQUIRE(shift == 0 || shift == ws / 2, "attention (fp32): shift must be 0 or ws/2");
  Geo32 g{res, shift, res / ws, heads, heads * hd, ws};
  const int units = batch * g.nws * g.nws * heads;
  switch (hd) {
    case 16: return launch32<16>(bwd, qkv, o, d_o, lse, tab2, alpha, out_or_dqkv, dtab, dalpha, g_qbias, g_vbias, g, units, lo_off, st);
    case 32: return launch32<32>(bwd, qkv, o, d_o, lse, tab2, alpha, out_or_dqkv, dtab, dalpha, g_qbias, g_vbias, g, units, lo_off, st);
    case 64: return launch32<64>(bwd, qkv, o, d_o, lse, tab2, alpha, out_or_dqkv, dtab, dalpha, g_qbias, g_vbias, g, units, lo_off, st);
  }
  SCOT_REQUIRE(false, "attention (fp32): head_dim %d must be 16/32/64", hd);
}

}  // namespace

int scot_attn32_fwd_launch(const void* qkv, void* out, float* lse, const float* tab2, const float* alpha, int batch, int res,
                           int ws, int shift, int heads, int hd, size_t lo_off, cudaStream_t st) {
  return dispatch32(false, qkv, nullptr, nullptr, lse, tab2, alpha, out, nullptr, nullptr, nullptr, nullptr, batch, res, ws,
                    shift, heads, hd, lo_off, st);
}
int scot_attn32_bwd_launch(const void* qkv, const void* o, const void* d_o, const float* lse, const float* tab2,
                           const float* alpha, void* dqkv, float* dtab, float* dalpha, float* g_qbias, float* g_vbias,
                           int batch, int res, int ws, int shift, int heads, int hd, size_t lo_off, cudaStream_t st) {
  return dispatch32(true, qkv, o, d_o, const_cast<float*>(lse), tab2, alpha, dqkv, dtab, dalpha, g_qbias, g_vbias, batch, res,
                    ws, shift, heads, hd, lo_off, st);
}

// (Conditional) LayerNorm forward / backward, fused with the residual add, the bf16 down-cast for the next
// GEMM, the pixel-shuffle row permutation of ScOTPatchUnmerging and the bias-gradient column sums.
//
// Reference: ConditionalLayerNorm.forward (scOT/model.py:150-160): mean, E[x^2]-mean^2, eps inside the
// sqrt, weight/bias = Linear(1, C)(time); LayerNorm (model.py:135-140) when use_conditioning=False.
// Call sites fused here: ScOTLayer res-post-norm adds (model.py:570,574), ScOTEmbeddings.norm (:352),
// ScOTPatchMerging.norm (:710), ScOTPatchUnmerging permute+norm (:748-759), ConvNeXtBlock.norm (:208).
// HBM-bound: one (sub-)warp per token row, 128-bit loads, shuffle reductions, row kept in registers.
// Two kernel families: cln_*_rows_kernel (a block's rows belong to one sample: lead time and scale / shift vectors are block
// constants, two rows in flight per (sub-)warp — the path of every shipped configuration up to C = 768) and cln_*_kernel
// (per-row lead time; blocks may span samples; C = 1536).

#include "common.cuh"
#include "internal.h"

namespace {

struct ClnFwdArgs {
  const float* z;        // [rows, C] input (row index = r_in)
  const float* residual; // [rows, C] or null (row index = r_out)
  const float* time;     // [B] or null
  const float* aw;       // scale slope  (weight.weight) [C] or null
  const float* ab;       // scale offset (weight.bias)   [C]
  const float* cw;       // shift slope  (bias.weight)   [C] or null
  const float* cb;       // shift offset (bias.bias)     [C]
  float* x_out;          // fp32 [rows, C] or null
  bf16* xb_out;          // bf16 [rows, C] or null
  bf16* zhat;            // bf16 normalised input (saved for backward) or null
  float* rstd;           // [rows] or null
  long rows;
  int C;
  int rows_per_sample;   // of the OUTPUT row index
  int perm_res;          // 0: r_out = r_in. >0: unmerge pixel shuffle with input grid perm_res x perm_res
  float eps;
  size_t lo_off;         // split-bf16 ("parity") mode: byte distance of the lo twins of xb_out / zhat (0 = none)
};

// r_in = ((b*res + i)*res + j)*4 + a*2 + c  ->  r_out = (b*2res + 2i+a)*2res + 2j+c   (model.py:748-754)
__device__ __forceinline__ long unmerge_row(long r_in, int res) {
  const int ac = (int)(r_in & 3);
  const long m = r_in >> 2;
  const int j = (int)(m % res);
  const long t = m / res;
  const int i = (int)(t % res);
  const long b = t / res;
  const int a = ac >> 1, c = ac & 1;
  return (b * (2 * res) + 2 * i + a) * (long)(2 * res) + 2 * j + c;
}

template <int LPR, int V>
__global__ void __launch_bounds__(256) cln_fwd_kernel(ClnFwdArgs p) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int RPW = 32 / LPR;  // rows per warp
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, sl = lane % LPR;
  const long warp_global = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long r_in = warp_global * RPW + sub;
  if (r_in >= p.rows) return;  // whole sub-warp exits together; shuffles below use the sub-group mask only
  const unsigned mask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (sub * LPR));
  const int nvec = p.C >> 2;
  float4 v[V];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int cv = sl + i * LPR;
    if (cv < nvec) {
      v[i] = *reinterpret_cast<const float4*>(p.z + r_in * p.C + cv * 4);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    } else {
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(mask, s, o);
  const float mean = s / (float)p.C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int cv = sl + i * LPR;
    if (cv < nvec) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(mask, q, o);
  const float rstd = rsqrtf(q / (float)p.C + p.eps);
  const long r_out = p.perm_res > 0 ? unmerge_row(r_in, p.perm_res) : r_in;
  const float t = p.time != nullptr ? p.time[r_out / p.rows_per_sample] : 0.f;
  if (sl == 0 && p.rstd != nullptr) p.rstd[r_out] = rstd;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int cv = sl + i * LPR;
    if (cv >= nvec) continue;
    const int c0 = cv * 4;
    float zh[4] = {(v[i].x - mean) * rstd, (v[i].y - mean) * rstd, (v[i].z - mean) * rstd, (v[i].w - mean) * rstd};
    // saved for the backward pass in bf16 (the forward output below uses the unrounded fp32 value)
    if (p.zhat != nullptr) st_bf16x4(p.zhat + r_out * p.C + c0, p.lo_off, zh[0], zh[1], zh[2], zh[3]);
    const float4 ab = *reinterpret_cast<const float4*>(p.ab + c0);
    const float4 cb = *reinterpret_cast<const float4*>(p.cb + c0);
    float sc[4] = {ab.x, ab.y, ab.z, ab.w}, sh[4] = {cb.x, cb.y, cb.z, cb.w};
    if (p.aw != nullptr) {
      const float4 aw = *reinterpret_cast<const float4*>(p.aw + c0);
      const float4 cw = *reinterpret_cast<const float4*>(p.cw + c0);
      sc[0] = fmaf(aw.x, t, sc[0]); sc[1] = fmaf(aw.y, t, sc[1]); sc[2] = fmaf(aw.z, t, sc[2]); sc[3] = fmaf(aw.w, t, sc[3]);
      sh[0] = fmaf(cw.x, t, sh[0]); sh[1] = fmaf(cw.y, t, sh[1]); sh[2] = fmaf(cw.z, t, sh[2]); sh[3] = fmaf(cw.w, t, sh[3]);
    }
    float y[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) y[k] = fmaf(sc[k], zh[k], sh[k]);
    if (p.residual != nullptr) {
      const float4 r = *reinterpret_cast<const float4*>(p.residual + r_out * p.C + c0);
      y[0] += r.x; y[1] += r.y; y[2] += r.z; y[3] += r.w;
    }
    if (p.x_out != nullptr) *reinterpret_cast<float4*>(p.x_out + r_out * p.C + c0) = make_float4(y[0], y[1], y[2], y[3]);
    if (p.xb_out != nullptr) st_bf16x4(p.xb_out + r_out * p.C + c0, p.lo_off, y[0], y[1], y[2], y[3]);
  }
}

// Forward for blocks whose rows all belong to one sample (rows_per_block divides rows_per_sample) or to an unconditioned
// norm: scale / shift vectors ab + aw t, cb + cw t are per-block constants in registers, every (sub-)warp loops over its rows
// with two rows (input and residual) in flight. The one-row-per-(sub-)warp kernel above remains the generic path.
template <int LPR, int V, bool EXACT>
__global__ void __launch_bounds__(256, (V <= 3 ? 2 : 1)) cln_fwd_rows_kernel(ClnFwdArgs p, int rows_per_block) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int RPW = 32 / LPR;
  constexpr int R = (V <= 3) ? 2 : 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int sub = lane / LPR, sl = lane % LPR;
  const unsigned mask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (sub * LPR));
  const int nvec = p.C >> 2;
  const long row_begin = (long)blockIdx.x * rows_per_block;
  const long left = p.rows - row_begin;
  const int nrows = left < (long)rows_per_block ? (int)left : rows_per_block;
  const float t = p.time != nullptr ? p.time[row_begin / p.rows_per_sample] : 0.f;
  float4 sc[V], sh[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int cv = sl + i * LPR;
    sc[i] = sh[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (EXACT || cv < nvec) {
      sc[i] = *reinterpret_cast<const float4*>(p.ab + cv * 4);
      sh[i] = *reinterpret_cast<const float4*>(p.cb + cv * 4);
      if (p.aw != nullptr) {
        const float4 aw = *reinterpret_cast<const float4*>(p.aw + cv * 4);
        const float4 cw = *reinterpret_cast<const float4*>(p.cw + cv * 4);
        sc[i].x = fmaf(aw.x, t, sc[i].x); sc[i].y = fmaf(aw.y, t, sc[i].y); sc[i].z = fmaf(aw.z, t, sc[i].z); sc[i].w = fmaf(aw.w, t, sc[i].w);
        sh[i].x = fmaf(cw.x, t, sh[i].x); sh[i].y = fmaf(cw.y, t, sh[i].y); sh[i].z = fmaf(cw.z, t, sh[i].z); sh[i].w = fmaf(cw.w, t, sh[i].w);
      }
    }
  }
  const float inv_c = 1.0f / (float)p.C;
  const int row_stride = nwarps * RPW;
  for (int rr0 = warp * RPW + sub; rr0 < nrows; rr0 += row_stride * R) {
    float4 v[R][V], res[R][V];
    long rout[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int rr = rr0 + k * row_stride;
      const bool in = rr < nrows;
      const long r_in = row_begin + rr;
      rout[k] = (p.perm_res > 0 && in) ? unmerge_row(r_in, p.perm_res) : r_in;
      const size_t ioff = (size_t)r_in * (size_t)p.C + (size_t)sl * 4;
      const size_t ooff = (size_t)rout[k] * (size_t)p.C + (size_t)sl * 4;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        v[k][i] = res[k][i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((EXACT || sl + i * LPR < nvec) && in) {
          v[k][i] = *reinterpret_cast<const float4*>(p.z + ioff + i * LPR * 4);
          if (p.residual != nullptr) res[k][i] = *reinterpret_cast<const float4*>(p.residual + ooff + i * LPR * 4);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int rr = rr0 + k * row_stride;
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) s += (v[k][i].x + v[k][i].y) + (v[k][i].z + v[k][i].w);
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(mask, s, o);
      const float mean = s * inv_c;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        if (EXACT || sl + i * LPR < nvec) {
          v[k][i].x -= mean; v[k][i].y -= mean; v[k][i].z -= mean; v[k][i].w -= mean;
          q = fmaf(v[k][i].x, v[k][i].x, q); q = fmaf(v[k][i].y, v[k][i].y, q);
          q = fmaf(v[k][i].z, v[k][i].z, q); q = fmaf(v[k][i].w, v[k][i].w, q);
        }
      }
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(mask, q, o);
      if (rr >= nrows) continue;  // uniform over the sub-warp that owns the row
      const float rstd = rsqrtf(q * inv_c + p.eps);
      if (sl == 0 && p.rstd != nullptr) p.rstd[rout[k]] = rstd;
      const size_t ooff = (size_t)rout[k] * (size_t)p.C + (size_t)sl * 4;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        if (EXACT || sl + i * LPR < nvec) {
          const float z0 = v[k][i].x * rstd, z1 = v[k][i].y * rstd, z2 = v[k][i].z * rstd, z3 = v[k][i].w * rstd;
          // saved for the backward pass in bf16 (the forward output below uses the unrounded fp32 value)
          if (p.zhat != nullptr) st_bf16x4(p.zhat + ooff + i * LPR * 4, p.lo_off, z0, z1, z2, z3);
          const float y0 = fmaf(sc[i].x, z0, sh[i].x) + res[k][i].x, y1 = fmaf(sc[i].y, z1, sh[i].y) + res[k][i].y;
          const float y2 = fmaf(sc[i].z, z2, sh[i].z) + res[k][i].z, y3 = fmaf(sc[i].w, z3, sh[i].w) + res[k][i].w;
          if (p.x_out != nullptr) *reinterpret_cast<float4*>(p.x_out + ooff + i * LPR * 4) = make_float4(y0, y1, y2, y3);
          if (p.xb_out != nullptr) st_bf16x4(p.xb_out + ooff + i * LPR * 4, p.lo_off, y0, y1, y2, y3);
        }
      }
    }
  }
}

struct ClnBwdArgs {
  const float* dy;   // [rows, C] gradient wrt the LN output (row index r_out)
  const bf16* zhat;  // [rows, C] (r_out)
  const float* rstd; // [rows]    (r_out)
  const float* time; // [B] or null
  const float* aw;   // [C] or null
  const float* ab;   // [C]
  void* dz;          // [rows, C] gradient wrt the LN input (row index r_in); bf16 or fp32
  int dz_is_f32;
  float* g_aw;       // param grads (atomic accumulate); g_aw / g_cw may be null (unconditioned)
  float* g_ab;
  float* g_cw;
  float* g_cb;
  float* g_bias_prev; // [C] or null: += column sums of dz (bias of the Linear that produced the LN input)
  long rows;
  int C;
  int rows_per_sample;
  int rows_per_block; // divides rows_per_sample
  int perm_res;
  size_t lo_off;      // split-bf16 ("parity") mode: lo twins of zhat (read) and of a bf16 dz (written)
};

// SPLIT: split-bf16 ("parity") mode — zhat is read as hi + lo, a bf16 dz is written as hi + lo
template <int LPR, int V, bool SPLIT = false>
__global__ void __launch_bounds__(256, ((V <= 3 && !SPLIT) ? 2 : 1)) cln_bwd_kernel(ClnBwdArgs p) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float red[];  // [warps][5][C]
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int sub = lane / LPR, sl = lane % LPR;
  const unsigned mask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (sub * LPR));
  const int nvec = p.C >> 2;
  const long row_begin = (long)blockIdx.x * p.rows_per_block;

  // scale = ab + aw * t(row): the block may span several samples, so t is looked up per row. Wide rows keep the two
  // scale vectors in registers; narrow-row variants (several rows per warp, V = 3) re-read them through L1 instead, which
  // keeps the kernel at two CTAs per SM.
  constexpr bool kScaleInRegs = (V <= 2);
  float4 sab[kScaleInRegs ? V : 1], saw[kScaleInRegs ? V : 1];
  if constexpr (kScaleInRegs) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int cv = sl + i * LPR;
      sab[i] = saw[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (cv < nvec) {
        sab[i] = *reinterpret_cast<const float4*>(p.ab + cv * 4);
        if (p.aw != nullptr) saw[i] = *reinterpret_cast<const float4*>(p.aw + cv * 4);
      }
    }
  }
  // column sums: dy*zhat, dy, dz and (conditioned norm) the same two weighted by t
  float4 acc_a[V], acc_c[V], acc_b[V], acc_at[V], acc_ct[V];
#pragma unroll
  for (int i = 0; i < V; ++i) acc_a[i] = acc_c[i] = acc_b[i] = acc_at[i] = acc_ct[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  // R rows per (sub-)warp are processed together so that several independent global loads are in flight
  constexpr int R = (V <= 1) ? 4 : (V <= 6 ? 2 : 1);
  const int row_stride = nwarps * RPW;
  for (int rr0 = warp * RPW + sub; rr0 < p.rows_per_block; rr0 += row_stride * R) {
    float4 dy[R][V];
    uint2 zp[R][V];  // normalised input, packed bf16 (unpacked at each use: two ALU ops instead of two more registers)
    uint2 zl[SPLIT ? R : 1][SPLIT ? V : 1];
    float rs[R], tt[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int rr = rr0 + k * row_stride;
      const long r_out = row_begin + rr;
      const bool in = rr < p.rows_per_block && r_out < p.rows;
      rs[k] = in ? p.rstd[r_out] : 0.f;
      tt[k] = (in && p.time != nullptr) ? p.time[(int)r_out / p.rows_per_sample] : 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int cv = sl + i * LPR;
        if (cv < nvec && in) {
          dy[k][i] = *reinterpret_cast<const float4*>(p.dy + r_out * p.C + cv * 4);
          zp[k][i] = *reinterpret_cast<const uint2*>(p.zhat + r_out * p.C + cv * 4);
          if constexpr (SPLIT)
            zl[k][i] = *reinterpret_cast<const uint2*>(reinterpret_cast<const char*>(p.zhat + r_out * p.C + cv * 4) + p.lo_off);
        } else {
          dy[k][i] = make_float4(0.f, 0.f, 0.f, 0.f);
          zp[k][i] = make_uint2(0u, 0u);
          if constexpr (SPLIT) zl[k][i] = make_uint2(0u, 0u);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int rr = rr0 + k * row_stride;  // rows beyond the block contribute zeros and are not stored
      const long r_out = row_begin + rr;
      float s1 = 0.f, s2 = 0.f;
      const float t = tt[k];
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float2 z01 = unpack_bf16x2(zp[k][i].x), z23 = unpack_bf16x2(zp[k][i].y);
        if constexpr (SPLIT) {
          const float2 l01 = unpack_bf16x2(zl[k][i].x), l23 = unpack_bf16x2(zl[k][i].y);
          z01.x += l01.x; z01.y += l01.y; z23.x += l23.x; z23.y += l23.y;
        }
        const float4 pz = make_float4(dy[k][i].x * z01.x, dy[k][i].y * z01.y, dy[k][i].z * z23.x, dy[k][i].w * z23.y);
        acc_a[i].x += pz.x; acc_a[i].y += pz.y; acc_a[i].z += pz.z; acc_a[i].w += pz.w;
        acc_c[i].x += dy[k][i].x; acc_c[i].y += dy[k][i].y; acc_c[i].z += dy[k][i].z; acc_c[i].w += dy[k][i].w;
        acc_at[i].x = fmaf(t, pz.x, acc_at[i].x); acc_at[i].y = fmaf(t, pz.y, acc_at[i].y);
        acc_at[i].z = fmaf(t, pz.z, acc_at[i].z); acc_at[i].w = fmaf(t, pz.w, acc_at[i].w);
        acc_ct[i].x = fmaf(t, dy[k][i].x, acc_ct[i].x); acc_ct[i].y = fmaf(t, dy[k][i].y, acc_ct[i].y);
        acc_ct[i].z = fmaf(t, dy[k][i].z, acc_ct[i].z); acc_ct[i].w = fmaf(t, dy[k][i].w, acc_ct[i].w);
        // dzhat = dy * scale(t)
        float4 b4, w4;
        if constexpr (kScaleInRegs) {
          b4 = sab[i];
          w4 = saw[i];
        } else {
          const int cv = sl + i * LPR;
          b4 = w4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (cv < nvec) {
            b4 = __ldg(reinterpret_cast<const float4*>(p.ab + cv * 4));
            if (p.aw != nullptr) w4 = __ldg(reinterpret_cast<const float4*>(p.aw + cv * 4));
          }
        }
        dy[k][i].x *= fmaf(w4.x, t, b4.x); dy[k][i].y *= fmaf(w4.y, t, b4.y);
        dy[k][i].z *= fmaf(w4.z, t, b4.z); dy[k][i].w *= fmaf(w4.w, t, b4.w);
        s1 += (dy[k][i].x + dy[k][i].y) + (dy[k][i].z + dy[k][i].w);
        s2 += (dy[k][i].x * z01.x + dy[k][i].y * z01.y) + (dy[k][i].z * z23.x + dy[k][i].w * z23.y);
      }
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(mask, s1, o);
        s2 += __shfl_xor_sync(mask, s2, o);
      }
      if (rr >= p.rows_per_block || r_out >= p.rows) continue;
      const float m1 = s1 / (float)p.C, m2 = s2 / (float)p.C;
      const float rstd = rs[k];
      long r_in = r_out;
      if (p.perm_res > 0) {
        // inverse of unmerge_row
        const int w2 = 2 * p.perm_res;
        const int x = (int)(r_out % w2);
        const long tt = r_out / w2;
        const int y = (int)(tt % w2);
        const long b = tt / w2;
        r_in = (((b * p.perm_res + (y >> 1)) * p.perm_res + (x >> 1)) << 2) + ((y & 1) << 1) + (x & 1);
      }
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int cv = sl + i * LPR;
        if (cv < nvec) {
          float2 z01 = unpack_bf16x2(zp[k][i].x), z23 = unpack_bf16x2(zp[k][i].y);
          if constexpr (SPLIT) {
            const float2 l01 = unpack_bf16x2(zl[k][i].x), l23 = unpack_bf16x2(zl[k][i].y);
            z01.x += l01.x; z01.y += l01.y; z23.x += l23.x; z23.y += l23.y;
          }
          float4 dz;
          dz.x = (dy[k][i].x - m1 - z01.x * m2) * rstd;
          dz.y = (dy[k][i].y - m1 - z01.y * m2) * rstd;
          dz.z = (dy[k][i].z - m1 - z23.x * m2) * rstd;
          dz.w = (dy[k][i].w - m1 - z23.y * m2) * rstd;
          if (p.dz_is_f32) {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.dz) + r_in * p.C + cv * 4) = dz;
          } else {
            dz = st_bf16x4(reinterpret_cast<bf16*>(p.dz) + r_in * p.C + cv * 4, SPLIT ? p.lo_off : 0, dz.x, dz.y, dz.z, dz.w);
          }
          acc_b[i].x += dz.x; acc_b[i].y += dz.y; acc_b[i].z += dz.z; acc_b[i].w += dz.w;
        }
      }
    }
  }
  // block reduction without atomics: fold the row sub-groups of a warp with shuffles, park one partial per warp in
  // smem ([warp][5][C]), then every thread sums a few columns over the warps and issues the global atomics
#define FOLD4(ACC)                                                              \
  _Pragma("unroll") for (int o = LPR; o < 32; o <<= 1) {                        \
    (ACC).x += __shfl_xor_sync(0xffffffffu, (ACC).x, o);                        \
    (ACC).y += __shfl_xor_sync(0xffffffffu, (ACC).y, o);                        \
    (ACC).z += __shfl_xor_sync(0xffffffffu, (ACC).z, o);                        \
    (ACC).w += __shfl_xor_sync(0xffffffffu, (ACC).w, o);                        \
  }
  float* wred = red + (size_t)warp * 5 * p.C;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    if (LPR < 32) {
      FOLD4(acc_a[i]) FOLD4(acc_c[i]) FOLD4(acc_b[i]) FOLD4(acc_at[i]) FOLD4(acc_ct[i])
    }
    const int cv = sl + i * LPR;
    if (cv < nvec && sub == 0) {
      const int c0 = cv * 4;
      *reinterpret_cast<float4*>(wred + 0 * p.C + c0) = acc_a[i];
      *reinterpret_cast<float4*>(wred + 1 * p.C + c0) = acc_c[i];
      *reinterpret_cast<float4*>(wred + 2 * p.C + c0) = acc_b[i];
      *reinterpret_cast<float4*>(wred + 3 * p.C + c0) = acc_at[i];
      *reinterpret_cast<float4*>(wred + 4 * p.C + c0) = acc_ct[i];
    }
  }
#undef FOLD4
  __syncthreads();
  for (int k = threadIdx.x; k < 5 * p.C; k += blockDim.x) {
    float v = 0.f;
    for (int w = 0; w < nwarps; ++w) v += red[(size_t)w * 5 * p.C + k];
    const int slot = k / p.C, c = k - slot * p.C;
    float* dst = slot == 0 ? p.g_ab : slot == 1 ? p.g_cb : slot == 2 ? p.g_bias_prev : slot == 3 ? p.g_aw : p.g_cw;
    if (dst != nullptr) atomicAdd(dst + c, v);
  }
}

// 16-byte vector reduction into global memory when the address allows it (the flat gradient buffer aligns every parameter
// to 256 B), four scalar ones otherwise
__device__ __forceinline__ void red_add4(float* dst, float4 v) {
  if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  } else {
    atomicAdd(dst + 0, v.x);
    atomicAdd(dst + 1, v.y);
    atomicAdd(dst + 2, v.z);
    atomicAdd(dst + 3, v.w);
  }
}

// Backward for blocks whose rows all belong to ONE sample (rows_per_block divides rows_per_sample) or to an unconditioned
// norm: the lead time, hence the scale vector ab + aw t, is a per-block constant kept in registers, the t-weighted column
// sums are t times the plain ones (formed once, in the block reduction), and the per-row work is ~11 instructions per
// element. The kernel above (per-row lead time, five accumulator sets) remains the path for blocks that span samples.
// EXACT: C == 4 * LPR * V, no column predicates.
template <int LPR, int V, bool SPLIT, bool EXACT>
__global__ void __launch_bounds__(256, ((V <= 3 && !SPLIT) ? 2 : 1)) cln_bwd_rows_kernel(ClnBwdArgs p) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float red[];  // [warps][3][C]
  constexpr int RPW = 32 / LPR;
  constexpr int R = (V <= 3 && !SPLIT) ? 2 : 1;  // rows in flight per (sub-)warp
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int sub = lane / LPR, sl = lane % LPR;
  const unsigned mask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (sub * LPR));
  const int nvec = p.C >> 2;
  const long row_begin = (long)blockIdx.x * p.rows_per_block;
  const long left = p.rows - row_begin;
  const int nrows = left < (long)p.rows_per_block ? (int)left : p.rows_per_block;
  const float t = p.time != nullptr ? p.time[row_begin / p.rows_per_sample] : 0.f;
  float4 sc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int cv = sl + i * LPR;
    sc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (EXACT || cv < nvec) {
      sc[i] = *reinterpret_cast<const float4*>(p.ab + cv * 4);
      if (p.aw != nullptr) {
        const float4 w = *reinterpret_cast<const float4*>(p.aw + cv * 4);
        sc[i].x = fmaf(w.x, t, sc[i].x); sc[i].y = fmaf(w.y, t, sc[i].y);
        sc[i].z = fmaf(w.z, t, sc[i].z); sc[i].w = fmaf(w.w, t, sc[i].w);
      }
    }
  }
  float4 acc_a[V], acc_c[V], acc_b[V];  // column sums of dy*zhat, dy, dz
#pragma unroll
  for (int i = 0; i < V; ++i) acc_a[i] = acc_c[i] = acc_b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float inv_c = 1.0f / (float)p.C;
  const int row_stride = nwarps * RPW;
  for (int rr0 = warp * RPW + sub; rr0 < nrows; rr0 += row_stride * R) {
    float4 g[R][V];   // dy on load, dy * scale afterwards
    uint2 zp[R][V];   // normalised input, packed bf16
    uint2 zl[SPLIT ? R : 1][SPLIT ? V : 1];
    float rs[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int rr = rr0 + k * row_stride;
      const bool in = rr < nrows;
      const size_t off = (size_t)(row_begin + rr) * (size_t)p.C + (size_t)sl * 4;
      rs[k] = in ? p.rstd[row_begin + rr] : 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        if ((EXACT || sl + i * LPR < nvec) && in) {
          g[k][i] = *reinterpret_cast<const float4*>(p.dy + off + i * LPR * 4);
          zp[k][i] = *reinterpret_cast<const uint2*>(p.zhat + off + i * LPR * 4);
          if constexpr (SPLIT)
            zl[k][i] = *reinterpret_cast<const uint2*>(reinterpret_cast<const char*>(p.zhat + off + i * LPR * 4) + p.lo_off);
        } else {
          g[k][i] = make_float4(0.f, 0.f, 0.f, 0.f);
          zp[k][i] = make_uint2(0u, 0u);
          if constexpr (SPLIT) zl[k][i] = make_uint2(0u, 0u);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int rr = rr0 + k * row_stride;
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float2 z01 = unpack_bf16x2(zp[k][i].x), z23 = unpack_bf16x2(zp[k][i].y);
        if constexpr (SPLIT) {
          const float2 l01 = unpack_bf16x2(zl[k][i].x), l23 = unpack_bf16x2(zl[k][i].y);
          z01.x += l01.x; z01.y += l01.y; z23.x += l23.x; z23.y += l23.y;
        }
        float4& d = g[k][i];
        acc_a[i].x = fmaf(d.x, z01.x, acc_a[i].x); acc_a[i].y = fmaf(d.y, z01.y, acc_a[i].y);
        acc_a[i].z = fmaf(d.z, z23.x, acc_a[i].z); acc_a[i].w = fmaf(d.w, z23.y, acc_a[i].w);
        acc_c[i].x += d.x; acc_c[i].y += d.y; acc_c[i].z += d.z; acc_c[i].w += d.w;
        d.x *= sc[i].x; d.y *= sc[i].y; d.z *= sc[i].z; d.w *= sc[i].w;  // dzhat = dy * scale(t)
        s1 += (d.x + d.y) + (d.z + d.w);
        s2 = fmaf(d.x, z01.x, s2); s2 = fmaf(d.y, z01.y, s2); s2 = fmaf(d.z, z23.x, s2); s2 = fmaf(d.w, z23.y, s2);
      }
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(mask, s1, o);
        s2 += __shfl_xor_sync(mask, s2, o);
      }
      if (rr >= nrows) continue;  // uniform over the sub-warp that owns the row
      const long r_out = row_begin + rr;
      const float rstd = rs[k];
      const float cb = -(s1 * inv_c) * rstd, cz = -(s2 * inv_c) * rstd;  // dz = (dzhat - m1 - zhat m2) rstd
      long r_in = r_out;
      if (p.perm_res > 0) {
        // inverse of unmerge_row
        const int w2 = 2 * p.perm_res;
        const int x = (int)(r_out % w2);
        const long tt = r_out / w2;
        const int y = (int)(tt % w2);
        const long b = tt / w2;
        r_in = (((b * p.perm_res + (y >> 1)) * p.perm_res + (x >> 1)) << 2) + ((y & 1) << 1) + (x & 1);
      }
      const size_t ooff = (size_t)r_in * (size_t)p.C + (size_t)sl * 4;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        if (EXACT || sl + i * LPR < nvec) {
          float2 z01 = unpack_bf16x2(zp[k][i].x), z23 = unpack_bf16x2(zp[k][i].y);
          if constexpr (SPLIT) {
            const float2 l01 = unpack_bf16x2(zl[k][i].x), l23 = unpack_bf16x2(zl[k][i].y);
            z01.x += l01.x; z01.y += l01.y; z23.x += l23.x; z23.y += l23.y;
          }
          const float4 d = g[k][i];
          float4 dz;
          dz.x = fmaf(z01.x, cz, fmaf(d.x, rstd, cb));
          dz.y = fmaf(z01.y, cz, fmaf(d.y, rstd, cb));
          dz.z = fmaf(z23.x, cz, fmaf(d.z, rstd, cb));
          dz.w = fmaf(z23.y, cz, fmaf(d.w, rstd, cb));
          if (p.dz_is_f32) {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.dz) + ooff + i * LPR * 4) = dz;
          } else {
            st_bf16x4(reinterpret_cast<bf16*>(p.dz) + ooff + i * LPR * 4, SPLIT ? p.lo_off : 0, dz.x, dz.y, dz.z, dz.w);
          }
          acc_b[i].x += dz.x; acc_b[i].y += dz.y; acc_b[i].z += dz.z; acc_b[i].w += dz.w;
        }
      }
    }
  }
  // block reduction: fold the row sub-groups of a warp with shuffles, park one partial per warp in smem ([warp][3][C]),
  // then one thread per float4 column sums over the warps and issues the (vector) reductions into the gradient buffers
#define FOLD4(ACC)                                                              \
  _Pragma("unroll") for (int o = LPR; o < 32; o <<= 1) {                        \
    (ACC).x += __shfl_xor_sync(0xffffffffu, (ACC).x, o);                        \
    (ACC).y += __shfl_xor_sync(0xffffffffu, (ACC).y, o);                        \
    (ACC).z += __shfl_xor_sync(0xffffffffu, (ACC).z, o);                        \
    (ACC).w += __shfl_xor_sync(0xffffffffu, (ACC).w, o);                        \
  }
  float* wred = red + (size_t)warp * 3 * p.C;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    if (LPR < 32) {
      FOLD4(acc_a[i]) FOLD4(acc_c[i]) FOLD4(acc_b[i])
    }
    const int cv = sl + i * LPR;
    if ((EXACT || cv < nvec) && sub == 0) {
      *reinterpret_cast<float4*>(wred + 0 * p.C + cv * 4) = acc_a[i];
      *reinterpret_cast<float4*>(wred + 1 * p.C + cv * 4) = acc_c[i];
      *reinterpret_cast<float4*>(wred + 2 * p.C + cv * 4) = acc_b[i];
    }
  }
#undef FOLD4
  __syncthreads();
  for (int cv = threadIdx.x; cv < nvec; cv += blockDim.x) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a, b = a;
    for (int w = 0; w < nwarps; ++w) {
      const float* r = red + (size_t)w * 3 * p.C + cv * 4;
      const float4 ra = *reinterpret_cast<const float4*>(r), rc = *reinterpret_cast<const float4*>(r + p.C),
                   rb = *reinterpret_cast<const float4*>(r + 2 * p.C);
      a.x += ra.x; a.y += ra.y; a.z += ra.z; a.w += ra.w;
      c.x += rc.x; c.y += rc.y; c.z += rc.z; c.w += rc.w;
      b.x += rb.x; b.y += rb.y; b.z += rb.z; b.w += rb.w;
    }
    red_add4(p.g_ab + cv * 4, a);
    red_add4(p.g_cb + cv * 4, c);
    if (p.g_bias_prev != nullptr) red_add4(p.g_bias_prev + cv * 4, b);
    if (p.g_aw != nullptr) {
      red_add4(p.g_aw + cv * 4, make_float4(t * a.x, t * a.y, t * a.z, t * a.w));
      red_add4(p.g_cw + cv * 4, make_float4(t * c.x, t * c.y, t * c.z, t * c.w));
    }
  }
}

int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

// generic path: a block sweeps a contiguous range of rows that may span samples (per-row lead time)
template <int LPR, int V, bool SPLIT>
int launch_bwd_span(const ClnBwdArgs& a_in, cudaStream_t st) {
  ClnBwdArgs a = a_in;
  const int warps = a.C > 768 ? 4 : 8;  // [warps][5][C] floats of smem must fit 227 KB (C = 1536: Poseidon-L stage 3)
  if (a.rows_per_block <= 0) {
    const long target = (long)num_sms() * (V <= 3 ? 2 : 1);
    const long unit = (long)warps * (32 / LPR);  // rows per sweep of the block's (sub-)warps
    long rpb = (a.rows + target - 1) / target;
    rpb = (rpb + unit - 1) / unit * unit;
    if (rpb < 8) rpb = (8 + unit - 1) / unit * unit;
    a.rows_per_block = (int)rpb;
  }
  const long blocks = (a.rows + a.rows_per_block - 1) / a.rows_per_block;
  const size_t smem = (size_t)warps * 5 * a.C * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    constexpr int kMaxC = 4 * LPR * V;
    constexpr int kMaxSmem = (kMaxC > 768 ? 4 : 8) * 5 * kMaxC * 4;
    SCOT_CHECK_CUDA(cudaFuncSetAttribute(cln_bwd_kernel<LPR, V, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    attr_done = true;
  }
  SCOT_CHECK_CUDA(scot_launch_pdl(cln_bwd_kernel<LPR, V, SPLIT>, dim3((unsigned)blocks), dim3(warps * 32), smem, st, a));
  SCOT_LAUNCH_CHECK();
  return 0;
}

template <int LPR, int V, bool SPLIT, bool EXACT>
int launch_bwd_rows(const ClnBwdArgs& a, cudaStream_t st) {
  const int warps = 8;
  const long blocks = (a.rows + a.rows_per_block - 1) / a.rows_per_block;
  const size_t smem = (size_t)warps * 3 * a.C * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    constexpr int kMaxSmem = 8 * 3 * (4 * LPR * V) * 4;
    SCOT_CHECK_CUDA(cudaFuncSetAttribute(cln_bwd_rows_kernel<LPR, V, SPLIT, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    attr_done = true;
  }
  SCOT_CHECK_CUDA(scot_launch_pdl(cln_bwd_rows_kernel<LPR, V, SPLIT, EXACT>, dim3((unsigned)blocks), dim3(warps * 32), smem, st, a));
  SCOT_LAUNCH_CHECK();
  return 0;
}

// Rows per block for the one-sample-per-block kernels: about one resident wave of CTAs (`target` blocks), a whole number of
// (sub-)warp sweeps, and — for a conditioned norm — a divisor of rows_per_sample so that no block spans two lead times.
// 0: no such size (the generic kernels take over).
int rows_kernel_rpb(long rows, long T, bool conditioned, long sweep, long unit, long target) {
  long rpb = (rows + target - 1) / target;
  rpb = (rpb + sweep - 1) / sweep * sweep;
  if (!conditioned) return (int)rpb;
  if (T <= 0 || rows % T != 0) return 0;
  // smallest divisor of T that is a multiple of the sweep and >= the one-wave size; else the largest usable divisor below it
  long best = 0;
  for (long d = sweep; d <= T; d += sweep) {
    if (T % d != 0) continue;
    best = d;
    if (d >= rpb) break;
  }
  if (best == 0) {  // samples shorter than a sweep (deep stages: 16 rows per sample): whole warp-rows, masked tail
    for (long d = unit; d <= T; d += unit)
      if (T % d == 0) best = d;
  }
  return (int)best;
}

template <int LPR, int V>
int launch_fwd(const ClnFwdArgs& a, cudaStream_t st) {
  constexpr int RPW = 32 / LPR;
  const int warps = 8;
  if constexpr (V <= 6) {
    constexpr int R = (V <= 3) ? 2 : 1;
    const int rpb = rows_kernel_rpb(a.rows, a.rows_per_sample, a.time != nullptr, 8L * RPW * R, 8L * RPW,
                                                     (long)num_sms() * (V <= 3 ? 2 : 1));
    if (rpb > 0) {
      const long blocks = (a.rows + rpb - 1) / rpb;
      if (a.C == 4 * LPR * V)
        SCOT_CHECK_CUDA(scot_launch_pdl(cln_fwd_rows_kernel<LPR, V, true>, dim3((unsigned)blocks), dim3(warps * 32), 0, st, a, rpb));
      else
        SCOT_CHECK_CUDA(scot_launch_pdl(cln_fwd_rows_kernel<LPR, V, false>, dim3((unsigned)blocks), dim3(warps * 32), 0, st, a, rpb));
      SCOT_LAUNCH_CHECK();
      return 0;
    }
  }
  const long blocks = (a.rows + (long)warps * RPW - 1) / ((long)warps * RPW);
  SCOT_CHECK_CUDA(scot_launch_pdl(cln_fwd_kernel<LPR, V>, dim3((unsigned)blocks), dim3(warps * 32), 0, st, a));
  SCOT_LAUNCH_CHECK();
  return 0;
}

template <int LPR, int V, bool SPLIT>
int launch_bwd_t(const ClnBwdArgs& a_in, cudaStream_t st) {
  ClnBwdArgs a = a_in;
  constexpr bool kRowsKernel = (4 * LPR * V <= 768) || V <= 6;  // [8][3][C] floats of smem: C <= 768 (wider rows: generic path)
  if constexpr (kRowsKernel) {
    if (a.rows_per_block <= 0 && a.C <= 768) {
      constexpr int R = (V <= 3 && !SPLIT) ? 2 : 1;
      const int rpb = rows_kernel_rpb(a.rows, a.rows_per_sample, a.time != nullptr, 8L * (32 / LPR) * R, 8L * (32 / LPR),
                                      (long)num_sms() * ((V <= 3 && !SPLIT) ? 2 : 1));
      if (rpb > 0) {
        a.rows_per_block = rpb;
        if (a.C == 4 * LPR * V) return launch_bwd_rows<LPR, V, SPLIT, true>(a, st);
        return launch_bwd_rows<LPR, V, SPLIT, false>(a, st);
      }
    }
  }
  return launch_bwd_span<LPR, V, SPLIT>(a, st);
}
template <int LPR, int V>
int launch_bwd(const ClnBwdArgs& a, cudaStream_t st) {
  return a.lo_off != 0 ? launch_bwd_t<LPR, V, true>(a, st) : launch_bwd_t<LPR, V, false>(a, st);
}

}  // namespace

#define CLN_DISPATCH(FN, ARGS)                                                     \
  do {                                                                             \
    const int nvec = (ARGS).C / 4;                                                 \
    /* narrow rows: several rows per warp (LPR lanes per row, 3 float4 per lane) */ \
    if (nvec == 12) return FN<4, 3>(ARGS, st);                                     \
    if (nvec == 24) return FN<8, 3>(ARGS, st);                                     \
    if (nvec == 48) return FN<16, 3>(ARGS, st);                                    \
    if (nvec <= 8) return FN<8, 1>(ARGS, st);                                      \
    if (nvec <= 16) return FN<16, 1>(ARGS, st);                                    \
    if (nvec <= 32) return FN<32, 1>(ARGS, st);                                    \
    if (nvec <= 64) return FN<32, 2>(ARGS, st);                                    \
    if (nvec <= 96) return FN<32, 3>(ARGS, st);                                    \
    if (nvec <= 192) return FN<32, 6>(ARGS, st);                                   \
    if (nvec <= 384) return FN<32, 12>(ARGS, st);                                  \
    SCOT_REQUIRE(false, "layer norm: C=%d not supported (max 1536)", (ARGS).C);    \
  } while (0)

int scot_cln_fwd_launch(const float* z, const float* residual, const float* time, const float* aw, const float* ab,
                        const float* cw, const float* cb, float* x_out, void* xb_out, void* zhat, float* rstd,
                        long rows, int C, int rows_per_sample, int perm_res, float eps, cudaStream_t st) {
  SCOT_REQUIRE(z && ab && cb, "cln_fwd: null pointer");
  SCOT_REQUIRE(C % 4 == 0 && rows > 0 && rows_per_sample > 0, "cln_fwd: bad shape rows=%ld C=%d", rows, C);
  SCOT_REQUIRE((aw == nullptr) == (cw == nullptr), "cln_fwd: aw/cw must both be given or both null");
  SCOT_REQUIRE(aw == nullptr || time != nullptr, "cln_fwd: conditioned norm needs time");
  SCOT_REQUIRE(perm_res == 0 || residual == nullptr, "cln_fwd: residual not supported with permutation");
  ClnFwdArgs a{z, residual, time, aw, ab, cw, cb, x_out, (bf16*)xb_out, (bf16*)zhat, rstd, rows, C, rows_per_sample,
               perm_res, eps, scot_split_off()};
  CLN_DISPATCH(launch_fwd, a);
}

int scot_cln_bwd_launch(const float* dy, const void* zhat, const float* rstd, const float* time, const float* aw,
                        const float* ab, void* dz, int dz_is_f32, float* g_aw, float* g_ab, float* g_cw, float* g_cb,
                        float* g_bias_prev, long rows, int C, int rows_per_sample, int perm_res, cudaStream_t st) {
  SCOT_REQUIRE(dy && zhat && rstd && ab && dz && g_ab && g_cb, "cln_bwd: null pointer");
  SCOT_REQUIRE(C % 4 == 0 && rows > 0, "cln_bwd: bad shape");
  SCOT_REQUIRE((aw == nullptr) == (g_aw == nullptr) && (g_aw == nullptr) == (g_cw == nullptr), "cln_bwd: aw/g_aw/g_cw mismatch");
  SCOT_REQUIRE(aw == nullptr || time != nullptr, "cln_bwd: conditioned norm needs time");
  const long rpb = 0;  // rows per block: sized by the launchers (one resident wave)
  ClnBwdArgs a{dy, (const bf16*)zhat, rstd, time, aw, ab, dz, dz_is_f32, g_aw, g_ab, g_cw, g_cb, g_bias_prev, rows, C,
               rows_per_sample, (int)rpb, perm_res, scot_split_off()};
  CLN_DISPATCH(launch_bwd, a);
}

// Shifted-window cosine attention with log-spaced continuous relative position bias (SwinV2), forward and
// backward, reading/writing token-major [tokens, 3C] / [tokens, C] buffers directly: cyclic shift
// (torch.roll, scOT/model.py:522-525,556-559), window_partition / window_reverse (HF modeling_swinv2.py:
// 146-166) and the head split/merge permutes exist only as index arithmetic in the loaders/stores.
//
// Arithmetic follows Swinv2SelfAttention.forward (HF:421-487):
//   S = (q/max(|q|,1e-12)) (k/max(|k|,1e-12))^T * exp(min(logit_scale, ln 100))
//       + 16*sigmoid(cpb_mlp(coords))[rel_index] + 2*mask          (mask {0,-100} is added twice in HF 5.5)
//   P = softmax(S);  O = P v
// Backward (derived in SURVEY.md appendix D) recomputes P flash-style from q,k and the saved row
// log-sum-exp; nothing N x N is stored.  Tensor work uses warp-level mma.sync m16n8k16 (bf16, fp32
// accumulate); a tcgen05 version of these kernels is the planned next step.
#include <cstdlib>

#include "common.cuh"
#include "internal.h"

namespace {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

struct WinGeom {
  int res;      // tokens per image side at this stage
  int shift;    // 0 or ws/2
  int nws;      // windows per side
  int heads;
  int C;        // channels (= heads * HD)
};

template <int WS>
__device__ __forceinline__ long token_row(const WinGeom& g, int bw, int n) {
  const int nw = g.nws * g.nws;
  const int b = bw / nw, w = bw - b * nw;
  const int wi = w / g.nws, wj = w - wi * g.nws;
  int i = wi * WS + n / WS + g.shift;
  int j = wj * WS + n % WS + g.shift;
  if (i >= g.res) i -= g.res;
  if (j >= g.res) j -= g.res;
  return ((long)b * g.res + i) * g.res + j;
}
// region code of token n inside a window of the shifted image (scOT/model.py:448-472). `flags` = bit0: window in
// the last window row, bit1: last window column (0 when there is no shift), from win_flags().
__device__ __forceinline__ int win_flags(const WinGeom& g, int bw) {
  if (g.shift == 0) return 0;
  const int nw = g.nws * g.nws;
  const int w = bw % nw;
  const int wi = w / g.nws, wj = w - wi * g.nws;
  return (wi == g.nws - 1 ? 1 : 0) | (wj == g.nws - 1 ? 2 : 0);
}
template <int WS>
__device__ __forceinline__ int mask_code(int flags, int shift, int n) {
  const int hm = (flags & 1) && (n / WS >= WS - shift);
  const int wm = (flags & 2) && (n % WS >= WS - shift);
  return hm | (wm << 1);
}
// WS == 16, shift == 8, 64-column chunks: an 8-column mma tile lies inside one half of one window row, so "masked or
// not" is uniform per (row half, column tile) and, within a chunk, depends only on the parity of the tile index:
// four additive terms (2 x -100, HF:465-468, in log2 units) per chunk replace a per-element region compare.
struct MaskTerms16 {
  float r0_even, r0_odd, r1_even, r1_odd;
};
__device__ __forceinline__ MaskTerms16 mask_terms16(int wf, int code0, int code1, int chunk) {
  const int col_h = (wf & 1) && (chunk >= 2);  // columns chunk*64.. : window row index = chunk*4 + tile/2 >= 8
  const int lastc = (wf >> 1) & 1;             // odd tiles are the right half (q >= 8) of a window row
  const float neg = -200.0f * kLog2e;
  MaskTerms16 t;
  t.r0_even = ((code0 & 1) != col_h || (code0 >> 1) != 0) ? neg : 0.f;
  t.r0_odd = ((code0 & 1) != col_h || (code0 >> 1) != lastc) ? neg : 0.f;
  t.r1_even = ((code1 & 1) != col_h || (code1 >> 1) != 0) ? neg : 0.f;
  t.r1_odd = ((code1 & 1) != col_h || (code1 >> 1) != lastc) ? neg : 0.f;
  return t;
}

template <int WS>
__device__ __forceinline__ int bias_rowbase(int m) {
  return (m / WS) * (2 * WS - 1) + (m % WS) + (WS - 1) * (2 * WS - 1) + (WS - 1);
}
template <int WS>
__host__ __device__ constexpr int bias_coloff(int n) {
  return (n / WS) * (2 * WS - 1) + (n % WS);
}
// bias_coloff is additive over the index decomposition used by the mma fragments,
//   n = chunk * KC + nt * 8 + 2 * cq + j   (KC a multiple of WS or a single chunk; 2*cq + j < 8 never carries),
// so a table lookup tab[rowbase(m) - coloff(n)] becomes ONE per-thread base pointer (rowbase(m0) - coloff(2*cq)),
// a per-chunk pointer bump (coloff(KC)) and compile-time immediates (coloff(nt*8) + j; + coloff(8) for row m0 + 8):
// no integer arithmetic per score element.

// ---- cooperative loads --------------------------------------------------------------------------------
// copy `nrows` token rows x HD bf16 from a [tokens, ld] buffer (column offset col0) into smem [nrows][HD+8]
template <int WS, int HD>
__device__ __forceinline__ void load_rows_async(bf16* dst, const bf16* src, long ld, int col0, const WinGeom& g, int bw,
                                                int row0, int nrows, int tid, int nthreads) {
  constexpr int CPR = HD / 8;  // 16-byte chunks per row
  for (int i = tid; i < nrows * CPR; i += nthreads) {
    const int r = i / CPR, ch = i - r * CPR;
    const long tr = token_row<WS>(g, bw, row0 + r);
    cp_async_16(dst + r * (HD + 8) + ch * 8, src + tr * ld + col0 + ch * 8);
  }
}
// L2-normalise rows in place (F.normalize eps 1e-12); optionally record 1/max(|x|,eps)
template <int HD>
__device__ __forceinline__ void normalize_rows(bf16* buf, int nrows, float* inv_norm, int tid, int nthreads) {
  for (int r = tid; r < nrows; r += nthreads) {
    uint4* p = reinterpret_cast<uint4*>(buf + r * (HD + 8));
    float v[HD];
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
      const uint4 u = p[c];
      const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), cc = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
      v[c * 8 + 0] = a.x; v[c * 8 + 1] = a.y; v[c * 8 + 2] = b.x; v[c * 8 + 3] = b.y;
      v[c * 8 + 4] = cc.x; v[c * 8 + 5] = cc.y; v[c * 8 + 6] = d.x; v[c * 8 + 7] = d.y;
    }
#pragma unroll
    for (int c = 0; c < HD; ++c) ss = fmaf(v[c], v[c], ss);
    const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    if (inv_norm != nullptr) inv_norm[r] = inv;
#pragma unroll
    for (int c = 0; c < HD / 8; ++c) {
      uint4 u;
      u.x = pack_bf16x2(v[c * 8 + 0] * inv, v[c * 8 + 1] * inv);
      u.y = pack_bf16x2(v[c * 8 + 2] * inv, v[c * 8 + 3] * inv);
      u.z = pack_bf16x2(v[c * 8 + 4] * inv, v[c * 8 + 5] * inv);
      u.w = pack_bf16x2(v[c * 8 + 6] * inv, v[c * 8 + 7] * inv);
      p[c] = u;
    }
  }
}

// A fragments (16 rows x HD) from smem rows [16][HD+8]
template <int HD>
__device__ __forceinline__ void load_a_frags(uint32_t (*a)[4], const bf16* rows, int lane) {
  const int i = lane >> 3;
#pragma unroll
  for (int kk = 0; kk < HD / 16; ++kk) {
    const bf16* p = rows + ((lane & 7) + (i & 1) * 8) * (HD + 8) + kk * 16 + (i >> 1) * 8;
    ldsm_x4(a[kk], smem_u32(p));
  }
}
// acc[NT][4] (16 x 8*NT) = A(16 x HD) * Bsm[n0 + (8*NT rows)][HD]^T, B rows are "n" (keys/queries)
template <int HD, int NT>
__device__ __forceinline__ void mma_a_bT(float (*acc)[4], const uint32_t (*a)[4], const bf16* bsm, int n0, int lane) {
  const int i = lane >> 3;
#pragma unroll
  for (int nt = 0; nt < NT; nt += 2) {
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) {
      uint32_t b[4];
      const bf16* p = bsm + (n0 + nt * 8 + (i >> 1) * 8 + (lane & 7)) * (HD + 8) + kk * 16 + (i & 1) * 8;
      ldsm_x4(b, smem_u32(p));
      mma_bf16_16816(acc[nt], a[kk], &b[0]);
      mma_bf16_16816(acc[nt + 1], a[kk], &b[2]);
    }
  }
}
// out[HD/8][4] (16 x HD) += P(16 x 16*KG, as packed A fragments pa[KG][4]) * Bsm[k0 + 16*KG rows][HD]
template <int HD, int KG>
__device__ __forceinline__ void mma_p_b(float (*out)[4], const uint32_t (*pa)[4], const bf16* bsm, int k0, int lane) {
  const int i = lane >> 3;
#pragma unroll
  for (int kg = 0; kg < KG; ++kg) {
#pragma unroll
    for (int dt = 0; dt < HD / 8; dt += 2) {
      uint32_t b[4];
      const bf16* p = bsm + (k0 + kg * 16 + (i & 1) * 8 + (lane & 7)) * (HD + 8) + dt * 8 + (i >> 1) * 8;
      ldsm_x4_trans(b, smem_u32(p));
      mma_bf16_16816(out[dt], pa[kg], &b[0]);
      mma_bf16_16816(out[dt + 1], pa[kg], &b[2]);
    }
  }
}

// =================================================================================================
// continuous position bias table (HF:450-460, 489-510):  tab2[r,h] = 16*sigmoid(mlp(coords[r]))[h]*log2(e)
// =================================================================================================
__device__ __forceinline__ float cpb_coord(int i, int ws) {
  // HF:491-506 in fp32: x = i/(ws-1) * 8 ; sign(x)*log2(|x|+1)/log2(8)
  float x = (float)i;
  if (ws > 1) x = x / (float)(ws - 1);
  x *= 8.0f;
  const float s = (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f);
  return s * log2f(fabsf(x) + 1.0f) / 3.0f;
}

// All attention layers of the model are handled by ONE launch each (blockIdx.y = layer): the per-layer work is a few
// thousand 512-term dot products, far too small to amortise a launch of its own.
// forward: one warp per table ROW r = (dy, dx): the 512 hidden activations are computed once (16 per lane) and reused
// by all heads. (Round-1 history: a warp per (row, head) entry that recomputed the hidden layer per head and launched
// ~46 k mostly empty blocks took 120 us; this one 50 us, bit-identical table.)
__global__ void __launch_bounds__(256)
cpb_fwd_kernel(ScotCpbTable tab, const float* __restrict__ params, uint8_t* __restrict__ arena) {
  const ScotCpbLayer L = tab.layer[blockIdx.y];
  const int ws = L.ws, heads = L.heads;
  const int side = 2 * ws - 1;
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const float* w1 = params + L.w1;
  const float* b1 = params + L.b1;
  const float* w2 = params + L.w2;
  const float* ls = params + L.ls;
  float* tab2 = reinterpret_cast<float*>(arena + (size_t)L.tab2 * 256);
  float* alpha = reinterpret_cast<float*>(arena + (size_t)L.alpha * 256);
  if (blockIdx.x == 0 && threadIdx.x < heads) alpha[threadIdx.x] = __expf(fminf(ls[threadIdx.x], 4.605170185988092f));  // ln 100, HF:448
  if (r >= side * side) return;
  const float c0 = cpb_coord(r / side - (ws - 1), ws), c1 = cpb_coord(r % side - (ws - 1), ws);
  float hid[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int j = lane + 32 * k;
    hid[k] = fmaxf(fmaf(w1[2 * j], c0, fmaf(w1[2 * j + 1], c1, b1[j])), 0.f);
  }
  for (int h = 0; h < heads; ++h) {
    const float* w2h = w2 + h * 512;
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) t = fmaf(w2h[lane + 32 * k], hid[k], t);
    t = warp_sum(t);
    if (lane == 0) tab2[r * heads + h] = 16.0f / (1.0f + __expf(-t)) * kLog2e;
  }
}
// dtab[r,h] = d loss / d (16*sigmoid(t)) ; produces dpre = dtab * 16 * s * (1-s) and d logit_scale. No MLP
// recomputation: the forward table holds tab2 = 16 sigmoid(t) log2(e), so sigmoid(t) = tab2 / (16 log2 e) and the
// pre-pass is element-wise (5 us instead of 114 us for the 64 layers of Poseidon-B).
__global__ void __launch_bounds__(256)
cpb_bwd_pre_kernel(ScotCpbTable tab, const float* __restrict__ params, float* __restrict__ grads,
                        uint8_t* __restrict__ arena) {
  const ScotCpbLayer L = tab.layer[blockIdx.y];
  const int ws = L.ws, heads = L.heads;
  const int side = 2 * ws - 1;
  const int total = side * side * heads;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const float* ls = params + L.ls;
  const float* tab2 = reinterpret_cast<const float*>(arena + (size_t)L.tab2 * 256);
  const float* dtab = reinterpret_cast<const float*>(arena + (size_t)L.dtab * 256);
  const float* dalpha = reinterpret_cast<const float*>(arena + (size_t)L.dalpha * 256);
  float* dpre = reinterpret_cast<float*>(arena + (size_t)L.dpre * 256);
  if (blockIdx.x == 0 && threadIdx.x < heads) {
    const float v = ls[threadIdx.x];
    if (v <= 4.605170185988092f) atomicAdd(grads + L.ls + threadIdx.x, dalpha[threadIdx.x] * __expf(v));
  }
  if (idx >= total) return;
  const float sg = tab2[idx] * (1.0f / (16.0f * kLog2e));
  dpre[idx] = dtab[idx] * 16.0f * sg * (1.0f - sg);
}

// one thread per hidden unit j, blockIdx.x splits the table rows; register accumulation, few atomics
__global__ void __launch_bounds__(512)
cpb_bwd_mlp_kernel(ScotCpbTable tab, const float* __restrict__ params, float* __restrict__ grads,
                   const uint8_t* __restrict__ arena) {
  constexpr int MAXH = 32;
  __shared__ float sc[31 * 2];  // coordinate values per table row/column index (2*ws-1 <= 31)
  const ScotCpbLayer L = tab.layer[blockIdx.y];
  const int ws = L.ws, heads = L.heads;
  const int j = threadIdx.x;
  const int side = 2 * ws - 1;
  const int rows = side * side;
  if (j < side) sc[j] = cpb_coord(j - (ws - 1), ws);
  __syncthreads();
  const int per = (rows + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * per, r1 = min(rows, r0 + per);
  if (r0 >= r1) return;
  const float* w1 = params + L.w1;
  const float* w2 = params + L.w2;
  const float* dpre = reinterpret_cast<const float*>(arena + (size_t)L.dpre * 256);
  const float wa = w1[2 * j], wb = w1[2 * j + 1], bb = params[L.b1 + j];
  float w2j[MAXH], acc2[MAXH];
#pragma unroll
  for (int h = 0; h < MAXH; ++h) {
    w2j[h] = h < heads ? w2[h * 512 + j] : 0.f;
    acc2[h] = 0.f;
  }
  float gwa = 0.f, gwb = 0.f, gb = 0.f;
  for (int r = r0; r < r1; ++r) {
    const float c0 = sc[r / side], c1 = sc[r % side];
    const float pre = fmaf(wa, c0, fmaf(wb, c1, bb));
    const float hid = fmaxf(pre, 0.f);
    float dh = 0.f;
#pragma unroll
    for (int h = 0; h < MAXH; ++h) {
      if (h < heads) {
        const float d = dpre[r * heads + h];
        acc2[h] = fmaf(d, hid, acc2[h]);
        dh = fmaf(d, w2j[h], dh);
      }
    }
    if (pre > 0.f) {
      gwa = fmaf(dh, c0, gwa);
      gwb = fmaf(dh, c1, gwb);
      gb += dh;
    }
  }
#pragma unroll
  for (int h = 0; h < MAXH; ++h)
    if (h < heads) atomicAdd(grads + L.w2 + h * 512 + j, acc2[h]);
  atomicAdd(grads + L.w1 + 2 * j, gwa);
  atomicAdd(grads + L.w1 + 2 * j + 1, gwb);
  atomicAdd(grads + L.b1 + j, gb);
}

// =================================================================================================
// forward
// =================================================================================================
template <int WS, int HD>
struct FwdCfg {
  static constexpr int N = WS * WS;
  static constexpr int MT = N / 16;
  static constexpr int NWARP = MT >= 8 ? 8 : 4;
  static constexpr int UPC = (NWARP / MT) > 1 ? (NWARP / MT) : 1;   // (window, head) units per CTA
  static constexpr int TPW = (MT / NWARP) > 1 ? (MT / NWARP) : 1;   // m-tiles per warp
  static constexpr int KC = N < 64 ? N : 64;                         // keys per chunk
  static constexpr int ROWB = (HD + 8);
  static constexpr int TABN = (2 * WS - 1) * (2 * WS - 1);
  static constexpr size_t smem = (size_t)UPC * 3 * N * ROWB * 2 + (size_t)UPC * TABN * 4;
};

template <int WS, int HD, bool SHIFT>
__global__ void __launch_bounds__(FwdCfg<WS, HD>::NWARP * 32)
attn_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, float* __restrict__ lse,
                const float* __restrict__ tab2, const float* __restrict__ alpha, WinGeom g, int total_units) {
  pdl_launch_dependents();
  pdl_wait();
  using Cfg = FwdCfg<WS, HD>;
  constexpr int N = Cfg::N, ROWB = Cfg::ROWB, KC = Cfg::KC, NT = KC / 8;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* sq = reinterpret_cast<bf16*>(smem_raw);
  bf16* sk = sq + Cfg::UPC * N * ROWB;
  bf16* sv = sk + Cfg::UPC * N * ROWB;
  float* stab = reinterpret_cast<float*>(sv + Cfg::UPC * N * ROWB);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nthreads = Cfg::NWARP * 32;
  const long ld = 3L * g.C;

  const int unit0 = blockIdx.x * Cfg::UPC;
#pragma unroll
  for (int u = 0; u < Cfg::UPC; ++u) {
    const int unit = unit0 + u;
    if (unit >= total_units) break;
    const int bw = unit / g.heads, h = unit - bw * g.heads;
    load_rows_async<WS, HD>(sq + u * N * ROWB, qkv, ld, h * HD, g, bw, 0, N, tid, nthreads);
    load_rows_async<WS, HD>(sk + u * N * ROWB, qkv, ld, g.C + h * HD, g, bw, 0, N, tid, nthreads);
    load_rows_async<WS, HD>(sv + u * N * ROWB, qkv, ld, 2 * g.C + h * HD, g, bw, 0, N, tid, nthreads);
    for (int i = tid; i < Cfg::TABN; i += nthreads) stab[u * Cfg::TABN + i] = tab2[i * g.heads + h];
  }
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();
  // q and k rows of all units are contiguous in smem: normalise both in one sweep
  normalize_rows<HD>(sq, 2 * Cfg::UPC * N, nullptr, tid, nthreads);
  __syncthreads();

  const int gq = lane >> 2, cq = lane & 3;
#pragma unroll 1
  for (int t = 0; t < Cfg::TPW; ++t) {
    const int wt = warp * Cfg::TPW + t;  // tile index within the CTA
    const int u = wt / Cfg::MT, mt = wt - u * Cfg::MT;
    const int unit = unit0 + u;
    if (unit >= total_units) break;
    const int bw = unit / g.heads, h = unit - bw * g.heads;
    bf16* uq = sq + u * N * ROWB;
    const bf16* uk = sk + u * N * ROWB;
    const bf16* uv = sv + u * N * ROWB;
    const float* utab = stab + u * Cfg::TABN;
    const float a2 = alpha[h] * kLog2e;
    const int m0 = mt * 16 + gq, m1 = m0 + 8;
    constexpr int RBD = bias_coloff<WS>(8);  // rowbase(m0 + 8) - rowbase(m0)
    const float* tb = utab + bias_rowbase<WS>(m0) - bias_coloff<WS>(2 * cq);
    const int wf = win_flags(g, bw);
    const int code0 = mask_code<WS>(wf, SHIFT ? WS / 2 : 0, m0), code1 = mask_code<WS>(wf, SHIFT ? WS / 2 : 0, m1);

    uint32_t qa[HD / 16][4];
    load_a_frags<HD>(qa, uq + mt * 16 * ROWB, lane);
    float o[HD / 8][4];
#pragma unroll
    for (int d = 0; d < HD / 8; ++d) o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f;
    float mx0 = -INFINITY, mx1 = -INFINITY, l0 = 0.f, l1 = 0.f;

#pragma unroll 1
    for (int kc = 0; kc < N / KC; ++kc) {
      float s[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      mma_a_bT<HD, NT>(s, qa, uk, kc * KC, lane);
      MaskTerms16 mterm = {0.f, 0.f, 0.f, 0.f};
      if constexpr (SHIFT && WS == 16) mterm = mask_terms16(wf, code0, code1, kc);
      float cm0 = -INFINITY, cm1 = -INFINITY;
      const float* tk = tb - kc * bias_coloff<WS>(KC);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int n = kc * KC + nt * 8 + 2 * cq + j;
          (void)n;
          float v0 = fmaf(s[nt][j], a2, tk[-(bias_coloff<WS>(nt * 8) + j)]);
          float v1 = fmaf(s[nt][2 + j], a2, tk[RBD - (bias_coloff<WS>(nt * 8) + j)]);
          if constexpr (SHIFT && WS == 16) {
            v0 += (nt & 1) ? mterm.r0_odd : mterm.r0_even;
            v1 += (nt & 1) ? mterm.r1_odd : mterm.r1_even;
          } else if constexpr (SHIFT) {
            const int cn = mask_code<WS>(wf, SHIFT ? WS / 2 : 0, n);
            if (cn != code0) v0 -= 200.0f * kLog2e;
            if (cn != code1) v1 -= 200.0f * kLog2e;
          }
          s[nt][j] = v0;
          s[nt][2 + j] = v1;
          cm0 = fmaxf(cm0, v0);
          cm1 = fmaxf(cm1, v1);
        }
      }
      cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1));
      cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
      cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1));
      cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
      const float nm0 = fmaxf(mx0, cm0), nm1 = fmaxf(mx1, cm1);
      const float sc0 = fast_exp2(mx0 - nm0), sc1 = fast_exp2(mx1 - nm1);
      mx0 = nm0; mx1 = nm1;
      float rs0 = 0.f, rs1 = 0.f;
      uint32_t pa[NT / 2][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const float p0 = fast_exp2(s[nt][0] - nm0), p1 = fast_exp2(s[nt][1] - nm0);
        const float p2 = fast_exp2(s[nt][2] - nm1), p3 = fast_exp2(s[nt][3] - nm1);
        rs0 += p0 + p1;
        rs1 += p2 + p3;
        pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(p0, p1);
        pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p2, p3);
      }
      l0 = l0 * sc0 + rs0;
      l1 = l1 * sc1 + rs1;
#pragma unroll
      for (int d = 0; d < HD / 8; ++d) {
        o[d][0] *= sc0; o[d][1] *= sc0; o[d][2] *= sc1; o[d][3] *= sc1;
      }
      mma_p_b<HD, NT / 2>(o, pa, uv, kc * KC, lane);
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float il0 = 1.0f / l0, il1 = 1.0f / l1;
    if (cq == 0) {
      lse[(long)unit * N + m0] = mx0 + log2f(l0);
      lse[(long)unit * N + m1] = mx1 + log2f(l1);
    }
    // stage the 16 x HD output tile through this warp's own (no longer needed) q rows, then 16-byte stores
    bf16* ot = uq + mt * 16 * ROWB;
    __syncwarp();
#pragma unroll
    for (int d = 0; d < HD / 8; ++d) {
      *reinterpret_cast<uint32_t*>(ot + gq * ROWB + d * 8 + 2 * cq) = pack_bf16x2(o[d][0] * il0, o[d][1] * il0);
      *reinterpret_cast<uint32_t*>(ot + (gq + 8) * ROWB + d * 8 + 2 * cq) = pack_bf16x2(o[d][2] * il1, o[d][3] * il1);
    }
    __syncwarp();
    constexpr int CPR = HD / 8;
    for (int i = lane; i < 16 * CPR; i += 32) {
      const int r = i / CPR, ch = i - r * CPR;
      const long tr = token_row<WS>(g, bw, mt * 16 + r);
      *reinterpret_cast<uint4*>(out + tr * g.C + h * HD + ch * 8) = *reinterpret_cast<const uint4*>(ot + r * ROWB + ch * 8);
    }
  }
}

// Bias-gradient fold executed at the end of the dq kernel. `sacc` holds, per warp, a 16 x N tile of summed dS in mma
// fragment order: element (m, n) of m-tile mt lives at  warp(mt) * 16N + (((n>>3)<<2) + reg) * 32 + lane  with
// lane = (m%8)*4 + (n%8)/2 and reg = ((m%16)/8)*2 + n%2. One thread per table entry (dp, dq) = (p_query - p_key,
// q_query - q_key) walks the (query, key) pairs with that displacement whose query row belongs to this CTA.
template <int WS, int NWARP>
__device__ __forceinline__ void fold_bias_to_table(const float* __restrict__ sacc, float* __restrict__ dtab, int h, int heads,
                                                   int rg, int tid, int nthreads) {
  constexpr int N = WS * WS, MT = N / 16, SIDE = 2 * WS - 1, TABN = SIDE * SIDE;
  constexpr int WPI = (NWARP / MT) > 1 ? (NWARP / MT) : 1;  // windows processed side by side (small windows)
  constexpr int ACC = 16 * N;
  const int mt_lo = (MT > NWARP) ? rg * NWARP : 0;            // m-tiles owned by this CTA (row groups exist for WS = 16 only)
  const int mt_hi = (MT > NWARP) ? mt_lo + NWARP : MT;
  for (int r = tid; r < TABN; r += nthreads) {
    const int dp = r / SIDE - (WS - 1), dq = r % SIDE - (WS - 1);
    int pm0 = dp > 0 ? dp : 0, pm1 = dp < 0 ? WS + dp : WS;
    const int qm0 = dq > 0 ? dq : 0, qm1 = dq < 0 ? WS + dq : WS;
    if (MT > NWARP) {  // WS == 16: m-tile index == window row
      pm0 = pm0 > mt_lo ? pm0 : mt_lo;
      pm1 = pm1 < mt_hi ? pm1 : mt_hi;
    }
    float acc = 0.f;
    for (int pm = pm0; pm < pm1; ++pm) {
      for (int qm = qm0; qm < qm1; ++qm) {
        const int m = pm * WS + qm, n = (pm - dp) * WS + (qm - dq);
        const int mt = m >> 4;
        const int lane = ((m & 7) << 2) | ((n & 7) >> 1);
        const int reg = (((m >> 3) & 1) << 1) | (n & 1);
        const int slot = (((n >> 3) << 2) + reg) * 32 + lane;
#pragma unroll
        for (int c = 0; c < WPI; ++c) acc += sacc[(c * MT + mt - mt_lo) * ACC + slot];
      }
    }
    if (pm1 > pm0) atomicAdd(dtab + r * heads + h, acc);  // dtab is zero at the start of the backward pass
  }
}

// Same fold for 16 x 16 windows with the address arithmetic taken out of the inner loop (validated on B200 in round 2:
// bit-compatible gradients, -0.19 ms per Poseidon-B step, profiles/r02_bringup_summary.txt): a slot address splits into f(m) + g(n) with n = m - (16 dp + dq), and stepping one window row down
// (m += 16, n += 16) adds the constant 16 N + 256. Per (query, key) pair the loop below is one LDS, one FADD and one
// pointer add instead of ~20 integer instructions — the fold is 14 % of the dq kernel's executed instructions
// (profiles/r01_ncu_full_summary.md). Summation order differs from fold_bias_to_table (fp32 reassociation only).
template <int NWARP>
__device__ __forceinline__ void fold_bias_to_table16(const float* __restrict__ sacc, float* __restrict__ dtab, int h, int heads,
                                                     int rg, int tid, int nthreads) {
  constexpr int WS = 16, N = 256, MT = 16, SIDE = 31, TABN = SIDE * SIDE, ACC = 16 * N;
  static_assert(MT >= NWARP, "one window per iteration");
  const int mt_lo = (MT > NWARP) ? rg * NWARP : 0;
  const int mt_hi = (MT > NWARP) ? mt_lo + NWARP : MT;
  for (int r = tid; r < TABN; r += nthreads) {
    const int dp = r / SIDE - (WS - 1), dq = r % SIDE - (WS - 1);
    int pm0 = dp > 0 ? dp : 0, pm1 = dp < 0 ? WS + dp : WS;
    const int qm0 = dq > 0 ? dq : 0, qm1 = dq < 0 ? WS + dq : WS;
    pm0 = pm0 > mt_lo ? pm0 : mt_lo;
    pm1 = pm1 < mt_hi ? pm1 : mt_hi;
    if (pm1 <= pm0) continue;
    const int d = dp * WS + dq;
    float acc = 0.f;
    for (int qm = qm0; qm < qm1; ++qm) {
      const int m = pm0 * WS + qm, n = m - d;
      const float* p = sacc + ((m >> 4) - mt_lo) * ACC + ((m >> 3) & 1) * 64 + (m & 7) * 4 + (n >> 3) * 128 + (n & 1) * 32 +
                       ((n & 7) >> 1);
#pragma unroll 4
      for (int pm = pm0; pm < pm1; ++pm, p += ACC + 256) acc += *p;
    }
    atomicAdd(dtab + r * heads + h, acc);
  }
}

// =================================================================================================
// backward, kernel 1: dq (+ relative-position-bias and logit-scale gradients)
//   CTA = (head, row group, window chunk); every warp owns one 16-query tile and loops over windows;
//   dS contributions to the bias are accumulated across windows in thread-private smem slots and
//   dumped once per CTA (deterministic two-stage reduction, no per-element atomics).
// =================================================================================================
template <int WS, int HD, int NWARP>
struct DqCfg {
  static constexpr int N = WS * WS;
  static constexpr int MT = N / 16;
  static constexpr int WPI = (NWARP / MT) > 1 ? (NWARP / MT) : 1;  // windows per iteration
  static constexpr int RG = (MT / NWARP) > 1 ? (MT / NWARP) : 1;   // row groups per (window, head)
  static constexpr int KC = N < 64 ? N : 64;
  static constexpr int ROWB = HD + 8;
  static constexpr int TABN = (2 * WS - 1) * (2 * WS - 1);
  static constexpr int ACC_PER_WARP = 16 * N;  // floats
  static constexpr size_t smem = (size_t)WPI * 2 * N * ROWB * 2      // k, v
                                 + (size_t)NWARP * 3 * 16 * ROWB * 2   // q, do, o rows per warp
                                 + (size_t)NWARP * ACC_PER_WARP * 4    // bias-gradient accumulators
                                 + (size_t)TABN * 4 + (size_t)NWARP * 16 * 4 /*inv norms*/;
};

template <int WS, int HD, int NWARP, bool SHIFT>
__global__ void __launch_bounds__(NWARP * 32)
attn_bwd_dq_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ o_buf, const bf16* __restrict__ do_buf,
                   const float* __restrict__ lse, const float* __restrict__ tab2, const float* __restrict__ alpha,
                   bf16* __restrict__ dqkv, float* __restrict__ dtab, float* __restrict__ dalpha,
                   float* __restrict__ g_qbias, WinGeom g, int total_windows, int windows_per_chunk) {
  pdl_launch_dependents();
  pdl_wait();
  using Cfg = DqCfg<WS, HD, NWARP>;
  constexpr int N = Cfg::N, ROWB = Cfg::ROWB, KC = Cfg::KC, NT = KC / 8, MT = Cfg::MT, WPI = Cfg::WPI;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* sk = reinterpret_cast<bf16*>(smem_raw);
  bf16* sv = sk + WPI * N * ROWB;
  bf16* srow = sv + WPI * N * ROWB;                     // [NWARP][3][16][ROWB]
  float* sacc = reinterpret_cast<float*>(srow + NWARP * 3 * 16 * ROWB);
  float* stab = sacc + NWARP * Cfg::ACC_PER_WARP;
  float* sinv = stab + Cfg::TABN;                        // [NWARP][16]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int nthreads = NWARP * 32;
  const long ld = 3L * g.C;
  const int h = blockIdx.x, rg = blockIdx.y, chunk = blockIdx.z;
  const int gq = lane >> 2, cq = lane & 3;

  for (int i = tid; i < Cfg::TABN; i += nthreads) stab[i] = tab2[i * g.heads + h];
  for (int i = tid; i < NWARP * Cfg::ACC_PER_WARP; i += nthreads) sacc[i] = 0.f;
  const float al = alpha[h];
  const float a2 = al * kLog2e;
  const int wsub = (WPI > 1) ? warp / MT : 0;           // which of the WPI windows this warp works on
  const int mt = (WPI > 1) ? warp % MT : rg * NWARP + warp;
  bf16* myq = srow + warp * 3 * 16 * ROWB;
  bf16* mydo = myq + 16 * ROWB;
  bf16* myo = mydo + 16 * ROWB;
  float* myacc = sacc + warp * Cfg::ACC_PER_WARP;
  const int m0 = mt * 16 + gq, m1 = m0 + 8;
  constexpr int RBD = bias_coloff<WS>(8);  // rowbase(m0 + 8) - rowbase(m0)
  const float* tb = stab + bias_rowbase<WS>(m0) - bias_coloff<WS>(2 * cq);
  float acc_alpha = 0.f;
  float qb_acc[HD / 8][2];
#pragma unroll
  for (int d = 0; d < HD / 8; ++d) qb_acc[d][0] = qb_acc[d][1] = 0.f;

  const int w_begin = chunk * windows_per_chunk;
  const int w_end = min(total_windows, w_begin + windows_per_chunk);
#pragma unroll 1
  for (int wbase = w_begin; wbase < w_end; wbase += WPI) {
    __syncthreads();  // previous iteration's readers of k/v are done
#pragma unroll
    for (int u = 0; u < WPI; ++u) {
      const int bw = wbase + u;
      if (bw < w_end) {
        load_rows_async<WS, HD>(sk + u * N * ROWB, qkv, ld, g.C + h * HD, g, bw, 0, N, tid, nthreads);
        load_rows_async<WS, HD>(sv + u * N * ROWB, qkv, ld, 2 * g.C + h * HD, g, bw, 0, N, tid, nthreads);
      }
    }
    const int bw = wbase + wsub;
    const bool active = bw < w_end;
    if (active) {
      load_rows_async<WS, HD>(myq, qkv, ld, h * HD, g, bw, mt * 16, 16, lane, 32);
      load_rows_async<WS, HD>(mydo, do_buf, g.C, h * HD, g, bw, mt * 16, 16, lane, 32);
      load_rows_async<WS, HD>(myo, o_buf, g.C, h * HD, g, bw, mt * 16, 16, lane, 32);
    }
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();
    normalize_rows<HD>(sk, WPI * N, nullptr, tid, nthreads);
    if (active) normalize_rows<HD>(myq, 16, sinv + warp * 16, lane, 32);
    __syncthreads();
    if (!active) continue;

    // D = rowsum(dO * O) for rows m0, m1 (each quad lane sums a quarter of the row)
    float D0 = 0.f, D1 = 0.f;
#pragma unroll
    for (int d = 0; d < HD / 8; ++d) {
      const float2 a0 = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(mydo + gq * ROWB + d * 8 + 2 * cq));
      const float2 b0 = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(myo + gq * ROWB + d * 8 + 2 * cq));
      const float2 a1 = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(mydo + (gq + 8) * ROWB + d * 8 + 2 * cq));
      const float2 b1 = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(myo + (gq + 8) * ROWB + d * 8 + 2 * cq));
      D0 += a0.x * b0.x + a0.y * b0.y;
      D1 += a1.x * b1.x + a1.y * b1.y;
    }
    D0 += __shfl_xor_sync(0xffffffffu, D0, 1); D0 += __shfl_xor_sync(0xffffffffu, D0, 2);
    D1 += __shfl_xor_sync(0xffffffffu, D1, 1); D1 += __shfl_xor_sync(0xffffffffu, D1, 2);
    const int unit = bw * g.heads + h;
    const float L0 = lse[(long)unit * N + m0], L1 = lse[(long)unit * N + m1];
    const int wf = win_flags(g, bw);
    const int code0 = mask_code<WS>(wf, SHIFT ? WS / 2 : 0, m0), code1 = mask_code<WS>(wf, SHIFT ? WS / 2 : 0, m1);
    const bf16* uk = sk + wsub * N * ROWB;
    const bf16* uv = sv + wsub * N * ROWB;

    uint32_t qa[HD / 16][4], da[HD / 16][4];
    load_a_frags<HD>(qa, myq, lane);
    load_a_frags<HD>(da, mydo, lane);
    float dq[HD / 8][4];
#pragma unroll
    for (int d = 0; d < HD / 8; ++d) dq[d][0] = dq[d][1] = dq[d][2] = dq[d][3] = 0.f;

#pragma unroll 1
    for (int kc = 0; kc < N / KC; ++kc) {
      float s[NT][4], dp[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
      }
      mma_a_bT<HD, NT>(s, qa, uk, kc * KC, lane);
      mma_a_bT<HD, NT>(dp, da, uv, kc * KC, lane);
      MaskTerms16 mterm = {0.f, 0.f, 0.f, 0.f};
      if constexpr (SHIFT && WS == 16) mterm = mask_terms16(wf, code0, code1, kc);
      uint32_t dsa[NT / 2][4];
      float* accp = myacc + (kc * NT) * 128 + lane;
      const float* tk = tb - kc * bias_coloff<WS>(KC);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        float ds[4];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int n = kc * KC + nt * 8 + 2 * cq + j;
          (void)n;
          float v0 = fmaf(s[nt][j], a2, tk[-(bias_coloff<WS>(nt * 8) + j)]);
          float v1 = fmaf(s[nt][2 + j], a2, tk[RBD - (bias_coloff<WS>(nt * 8) + j)]);
          if constexpr (SHIFT && WS == 16) {
            v0 += (nt & 1) ? mterm.r0_odd : mterm.r0_even;
            v1 += (nt & 1) ? mterm.r1_odd : mterm.r1_even;
          } else if constexpr (SHIFT) {
            const int cn = mask_code<WS>(wf, SHIFT ? WS / 2 : 0, n);
            if (cn != code0) v0 -= 200.0f * kLog2e;
            if (cn != code1) v1 -= 200.0f * kLog2e;
          }
          const float p0 = fast_exp2(v0 - L0), p1 = fast_exp2(v1 - L1);
          ds[j] = p0 * (dp[nt][j] - D0);
          ds[2 + j] = p1 * (dp[nt][2 + j] - D1);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) accp[(nt * 4 + r) * 32] += ds[r];
        dsa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(ds[0], ds[1]);
        dsa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(ds[2], ds[3]);
      }
      mma_p_b<HD, NT / 2>(dq, dsa, uk, kc * KC, lane);  // dq_raw += dS k_hat   (alpha is applied once, after the loop)
    }
    // dq = (dq_hat - q_hat (q_hat . dq_hat)) / max(|q|, eps), dq_hat = alpha dq_raw
    float dot0 = 0.f, dot1 = 0.f;
    float qh[HD / 8][4];
#pragma unroll
    for (int d = 0; d < HD / 8; ++d) {
      const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(myq + gq * ROWB + d * 8 + 2 * cq));
      const float2 b = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(myq + (gq + 8) * ROWB + d * 8 + 2 * cq));
      qh[d][0] = a.x; qh[d][1] = a.y; qh[d][2] = b.x; qh[d][3] = b.y;
      dot0 += a.x * dq[d][0] + a.y * dq[d][1];
      dot1 += b.x * dq[d][2] + b.y * dq[d][3];
      dq[d][0] *= al; dq[d][1] *= al; dq[d][2] *= al; dq[d][3] *= al;
    }
    dot0 += __shfl_xor_sync(0xffffffffu, dot0, 1); dot0 += __shfl_xor_sync(0xffffffffu, dot0, 2);
    dot1 += __shfl_xor_sync(0xffffffffu, dot1, 1); dot1 += __shfl_xor_sync(0xffffffffu, dot1, 2);
    // logit-scale gradient: sum_n dS[m, n] (q_hat[m] . k_hat[n]) = q_hat[m] . dq_raw[m]  (every quad lane holds the row total)
    if (cq == 0) acc_alpha += dot0 + dot1;
    dot0 *= al;
    dot1 *= al;
    const float in0 = sinv[warp * 16 + gq], in1 = sinv[warp * 16 + gq + 8];
    __syncwarp();
#pragma unroll
    for (int d = 0; d < HD / 8; ++d) {
      const uint32_t u0 = pack_bf16x2((dq[d][0] - qh[d][0] * dot0) * in0, (dq[d][1] - qh[d][1] * dot0) * in0);
      const uint32_t u1 = pack_bf16x2((dq[d][2] - qh[d][2] * dot1) * in1, (dq[d][3] - qh[d][3] * dot1) * in1);
      const float2 f0 = unpack_bf16x2(u0), f1 = unpack_bf16x2(u1);
      qb_acc[d][0] += f0.x + f1.x;
      qb_acc[d][1] += f0.y + f1.y;
      *reinterpret_cast<uint32_t*>(myq + gq * ROWB + d * 8 + 2 * cq) = u0;
      *reinterpret_cast<uint32_t*>(myq + (gq + 8) * ROWB + d * 8 + 2 * cq) = u1;
    }
    __syncwarp();
    constexpr int CPR = HD / 8;
    for (int i = lane; i < 16 * CPR; i += 32) {
      const int r = i / CPR, ch = i - r * CPR;
      const long tr = token_row<WS>(g, bw, mt * 16 + r);
      *reinterpret_cast<uint4*>(dqkv + tr * ld + h * HD + ch * 8) = *reinterpret_cast<const uint4*>(myq + r * ROWB + ch * 8);
    }
  }
  __syncthreads();
  // fold this CTA's bias-gradient accumulators (fragment order, summed over its windows) onto the (2ws-1)^2 table of
  // the head through the relative position index (HF:512-523) and add the <= 961 table entries to the layer's
  // gradient table: a gather per table entry, no N x N round trip through global memory, no second kernel
  if constexpr (WS == 16) fold_bias_to_table16<NWARP>(sacc, dtab, h, g.heads, rg, tid, nthreads);
  else fold_bias_to_table<WS, NWARP>(sacc, dtab, h, g.heads, rg, tid, nthreads);
  acc_alpha = warp_sum(acc_alpha);
  if (lane == 0) atomicAdd(dalpha + h, acc_alpha);
  // query-bias gradient: sum over the 8 row groups of the warp (lanes with equal cq), then one atomic per column
#pragma unroll
  for (int d = 0; d < HD / 8; ++d) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float v = qb_acc[d][j];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (gq == 0 && g_qbias != nullptr) atomicAdd(g_qbias + h * HD + d * 8 + 2 * cq + j, v);
    }
  }
}

// =================================================================================================
// backward, kernel 2: dk, dv  (every warp owns 16 keys of a window; transposed recomputation)
// =================================================================================================
template <int WS, int HD, int NWARP>
struct DkvCfg {
  static constexpr int N = WS * WS;
  static constexpr int MT = N / 16;
  static constexpr int WPI = (NWARP / MT) > 1 ? (NWARP / MT) : 1;
  static constexpr int KG = (MT / NWARP) > 1 ? (MT / NWARP) : 1;  // key groups per (window, head)
  static constexpr int QC = N < 64 ? N : 64;                        // queries per chunk
  static constexpr int ROWB = HD + 8;
  static constexpr int TABN = (2 * WS - 1) * (2 * WS - 1);
  static constexpr size_t smem = (size_t)WPI * 2 * N * ROWB * 2     // q, do (all rows)
                                 + (size_t)NWARP * 2 * 16 * ROWB * 2 // own k, v rows
                                 + (size_t)WPI * 2 * N * 4           // lse, D
                                 + (size_t)TABN * 4 + (size_t)NWARP * 16 * 4;
};

template <int WS, int HD, int NWARP, bool SHIFT>
__global__ void __launch_bounds__(NWARP * 32, (HD <= 32 ? 2 : 1))
attn_bwd_dkv_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ o_buf, const bf16* __restrict__ do_buf,
                    const float* __restrict__ lse, const float* __restrict__ tab2, const float* __restrict__ alpha,
                    bf16* __restrict__ dqkv, float* __restrict__ g_vbias, WinGeom g, int total_windows,
                    int windows_per_chunk) {
  pdl_launch_dependents();
  pdl_wait();
  using Cfg = DkvCfg<WS, HD, NWARP>;
  constexpr int N = Cfg::N, ROWB = Cfg::ROWB, QC = Cfg::QC, NT = QC / 8, MT = Cfg::MT, WPI = Cfg::WPI;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  bf16* sq = reinterpret_cast<bf16*>(smem_raw);
  bf16* sdo = sq + WPI * N * ROWB;
  bf16* srow = sdo + WPI * N * ROWB;  // [NWARP][2][16][ROWB]
  float* slse = reinterpret_cast<float*>(srow + NWARP * 2 * 16 * ROWB);
  float* sD = slse + WPI * N;
  float* stab = sD + WPI * N;
  float* sinv = stab + Cfg::TABN;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int nthreads = NWARP * 32;
  const long ld = 3L * g.C;
  const int h = blockIdx.x, kg = blockIdx.y, chunk = blockIdx.z;
  const int gq = lane >> 2, cq = lane & 3;
  for (int i = tid; i < Cfg::TABN; i += nthreads) stab[i] = tab2[i * g.heads + h];
  const float al = alpha[h];
  const float a2 = al * kLog2e;
  const int wsub = (WPI > 1) ? warp / MT : 0;
  const int kt = (WPI > 1) ? warp % MT : kg * NWARP + warp;  // key tile of this warp
  bf16* myk = srow + warp * 2 * 16 * ROWB;
  bf16* myv = myk + 16 * ROWB;
  const int n0 = kt * 16 + gq, n1 = n0 + 8;  // keys owned by this thread's fragment rows
  constexpr int RBD = bias_coloff<WS>(8);  // coloff(n0 + 8) - coloff(n0)
  // tab[rowbase(m) - coloff(n0)] with m = chunk * QC + nt * 8 + 2 * cq + j: base pointer + immediates (see bias_coloff)
  const float* tb = stab + bias_rowbase<WS>(0) - bias_coloff<WS>(n0) + bias_coloff<WS>(2 * cq);
  float vb_acc[HD / 8][2];
#pragma unroll
  for (int d = 0; d < HD / 8; ++d) vb_acc[d][0] = vb_acc[d][1] = 0.f;

  const int w_begin = chunk * windows_per_chunk;
  const int w_end = min(total_windows, w_begin + windows_per_chunk);
#pragma unroll 1
  for (int wbase = w_begin; wbase < w_end; wbase += WPI) {
    __syncthreads();
#pragma unroll
    for (int u = 0; u < WPI; ++u) {
      const int bw = wbase + u;
      if (bw < w_end) {
        load_rows_async<WS, HD>(sq + u * N * ROWB, qkv, ld, h * HD, g, bw, 0, N, tid, nthreads);
        load_rows_async<WS, HD>(sdo + u * N * ROWB, do_buf, g.C, h * HD, g, bw, 0, N, tid, nthreads);
      }
    }
    const int bw = wbase + wsub;
    const bool active = bw < w_end;
    if (active) {
      load_rows_async<WS, HD>(myk, qkv, ld, g.C + h * HD, g, bw, kt * 16, 16, lane, 32);
      load_rows_async<WS, HD>(myv, qkv, ld, 2 * g.C + h * HD, g, bw, kt * 16, 16, lane, 32);
    }
    cp_async_commit();
    // meanwhile: lse and D = rowsum(dO * O) straight from global memory (one thread per query row)
    for (int i = tid; i < WPI * N; i += nthreads) {
      const int u = i / N, m = i - u * N;
      const int bwu = wbase + u;
      if (bwu < w_end) {
        slse[i] = lse[((long)bwu * g.heads + h) * N + m];
        const long tr = token_row<WS>(g, bwu, m);
        const uint4* po = reinterpret_cast<const uint4*>(o_buf + tr * g.C + h * HD);
        const uint4* pd = reinterpret_cast<const uint4*>(do_buf + tr * g.C + h * HD);
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < HD / 8; ++c) {
          const uint4 a = po[c], b = pd[c];
          const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2_ = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
          const float2 b0 = unpack_bf16x2(b.x), b1 = unpack_bf16x2(b.y), b2 = unpack_bf16x2(b.z), b3 = unpack_bf16x2(b.w);
          acc += a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y + a2_.x * b2.x + a2_.y * b2.y + a3.x * b3.x + a3.y * b3.y;
        }
        sD[i] = acc;
      }
    }
    cp_async_wait_all();
    __syncthreads();
    normalize_rows<HD>(sq, WPI * N, nullptr, tid, nthreads);
    if (active) normalize_rows<HD>(myk, 16, sinv + warp * 16, lane, 32);
    __syncthreads();
    if (!active) continue;

    const int wf = win_flags(g, bw);
    const int code0 = mask_code<WS>(wf, SHIFT ? WS / 2 : 0, n0), code1 = mask_code<WS>(wf, SHIFT ? WS / 2 : 0, n1);
    const bf16* uq = sq + wsub * N * ROWB;
    const bf16* udo = sdo + wsub * N * ROWB;
    const float* ulse = slse + wsub * N;
    const float* uD = sD + wsub * N;
    uint32_t ka[HD / 16][4], va[HD / 16][4];
    load_a_frags<HD>(ka, myk, lane);
    load_a_frags<HD>(va, myv, lane);
    float dk[HD / 8][4], dv[HD / 8][4];
#pragma unroll
    for (int d = 0; d < HD / 8; ++d) {
      dk[d][0] = dk[d][1] = dk[d][2] = dk[d][3] = 0.f;
      dv[d][0] = dv[d][1] = dv[d][2] = dv[d][3] = 0.f;
    }
#pragma unroll 1
    for (int qc = 0; qc < N / QC; ++qc) {
      float st[NT][4], dpt[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        st[nt][0] = st[nt][1] = st[nt][2] = st[nt][3] = 0.f;
        dpt[nt][0] = dpt[nt][1] = dpt[nt][2] = dpt[nt][3] = 0.f;
      }
      mma_a_bT<HD, NT>(st, ka, uq, qc * QC, lane);    // S^T tile: rows = keys, cols = queries
      mma_a_bT<HD, NT>(dpt, va, udo, qc * QC, lane);  // dP^T tile
      MaskTerms16 mterm = {0.f, 0.f, 0.f, 0.f};
      if constexpr (SHIFT && WS == 16) mterm = mask_terms16(wf, code0, code1, qc);
      uint32_t pta[NT / 2][4], dsta[NT / 2][4];
      const float* tk = tb + qc * bias_coloff<WS>(QC);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        float p[4], ds[4];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int m = qc * QC + nt * 8 + 2 * cq + j;  // query index (column)
          const float Lm = ulse[m], Dm = uD[m];
          float v0 = fmaf(st[nt][j], a2, tk[bias_coloff<WS>(nt * 8) + j]);
          float v1 = fmaf(st[nt][2 + j], a2, tk[bias_coloff<WS>(nt * 8) + j - RBD]);
          if constexpr (SHIFT && WS == 16) {
            v0 += (nt & 1) ? mterm.r0_odd : mterm.r0_even;
            v1 += (nt & 1) ? mterm.r1_odd : mterm.r1_even;
          } else if constexpr (SHIFT) {
            const int cm = mask_code<WS>(wf, SHIFT ? WS / 2 : 0, m);
            if (cm != code0) v0 -= 200.0f * kLog2e;
            if (cm != code1) v1 -= 200.0f * kLog2e;
          }
          p[j] = fast_exp2(v0 - Lm);
          p[2 + j] = fast_exp2(v1 - Lm);
          ds[j] = p[j] * (dpt[nt][j] - Dm);
          ds[2 + j] = p[2 + j] * (dpt[nt][2 + j] - Dm);
        }
        pta[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(p[0], p[1]);
        pta[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p[2], p[3]);
        dsta[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(ds[0], ds[1]);
        dsta[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(ds[2], ds[3]);
      }
      mma_p_b<HD, NT / 2>(dv, pta, udo, qc * QC, lane);   // dV  += P^T dO
      mma_p_b<HD, NT / 2>(dk, dsta, uq, qc * QC, lane);   // dk_raw += dS^T q_hat   (alpha is applied once, below)
    }
#pragma unroll
    for (int d = 0; d < HD / 8; ++d) {
      dk[d][0] *= al; dk[d][1] *= al; dk[d][2] *= al; dk[d][3] *= al;
    }
    // dk = (dk_hat - k_hat (k_hat . dk_hat)) / max(|k|, eps); stage dk in my k rows, dv in my v rows
    float dot0 = 0.f, dot1 = 0.f;
    float kh[HD / 8][4];
#pragma unroll
    for (int d = 0; d < HD / 8; ++d) {
      const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(myk + gq * ROWB + d * 8 + 2 * cq));
      const float2 b = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(myk + (gq + 8) * ROWB + d * 8 + 2 * cq));
      kh[d][0] = a.x; kh[d][1] = a.y; kh[d][2] = b.x; kh[d][3] = b.y;
      dot0 += a.x * dk[d][0] + a.y * dk[d][1];
      dot1 += b.x * dk[d][2] + b.y * dk[d][3];
    }
    dot0 += __shfl_xor_sync(0xffffffffu, dot0, 1); dot0 += __shfl_xor_sync(0xffffffffu, dot0, 2);
    dot1 += __shfl_xor_sync(0xffffffffu, dot1, 1); dot1 += __shfl_xor_sync(0xffffffffu, dot1, 2);
    const float in0 = sinv[warp * 16 + gq], in1 = sinv[warp * 16 + gq + 8];
    __syncwarp();
#pragma unroll
    for (int d = 0; d < HD / 8; ++d) {
      *reinterpret_cast<uint32_t*>(myk + gq * ROWB + d * 8 + 2 * cq) =
          pack_bf16x2((dk[d][0] - kh[d][0] * dot0) * in0, (dk[d][1] - kh[d][1] * dot0) * in0);
      *reinterpret_cast<uint32_t*>(myk + (gq + 8) * ROWB + d * 8 + 2 * cq) =
          pack_bf16x2((dk[d][2] - kh[d][2] * dot1) * in1, (dk[d][3] - kh[d][3] * dot1) * in1);
      const uint32_t u0 = pack_bf16x2(dv[d][0], dv[d][1]), u1 = pack_bf16x2(dv[d][2], dv[d][3]);
      const float2 f0 = unpack_bf16x2(u0), f1 = unpack_bf16x2(u1);
      vb_acc[d][0] += f0.x + f1.x;
      vb_acc[d][1] += f0.y + f1.y;
      *reinterpret_cast<uint32_t*>(myv + gq * ROWB + d * 8 + 2 * cq) = u0;
      *reinterpret_cast<uint32_t*>(myv + (gq + 8) * ROWB + d * 8 + 2 * cq) = u1;
    }
    __syncwarp();
    constexpr int CPR = HD / 8;
    for (int i = lane; i < 16 * CPR; i += 32) {
      const int r = i / CPR, ch = i - r * CPR;
      const long tr = token_row<WS>(g, bw, kt * 16 + r);
      *reinterpret_cast<uint4*>(dqkv + tr * ld + g.C + h * HD + ch * 8) = *reinterpret_cast<const uint4*>(myk + r * ROWB + ch * 8);
      *reinterpret_cast<uint4*>(dqkv + tr * ld + 2 * g.C + h * HD + ch * 8) = *reinterpret_cast<const uint4*>(myv + r * ROWB + ch * 8);
    }
  }
#pragma unroll
  for (int d = 0; d < HD / 8; ++d) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float v = vb_acc[d][j];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (gq == 0 && g_vbias != nullptr) atomicAdd(g_vbias + h * HD + d * 8 + 2 * cq + j, v);
    }
  }
}

// =================================================================================================
// host launchers
// =================================================================================================
int g_sms = 0;
int num_sms() {
  if (g_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_sms <= 0) g_sms = 148;
  }
  return g_sms;
}

template <int WS, int HD>
int launch_fwd(const void* qkv, void* out, float* lse, const float* tab2, const float* alpha, WinGeom g, int total_windows,
               cudaStream_t st) {
  using Cfg = FwdCfg<WS, HD>;
  static bool done = false;
  if (!done) {
    SCOT_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<WS, HD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem));
    SCOT_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<WS, HD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem));
    done = true;
  }
  auto kern = g.shift > 0 ? attn_fwd_kernel<WS, HD, true> : attn_fwd_kernel<WS, HD, false>;
  const int units = total_windows * g.heads;
  const int grid = ceil_div(units, Cfg::UPC);
  SCOT_CHECK_CUDA(scot_launch_pdl(kern, dim3(grid), dim3(Cfg::NWARP * 32), Cfg::smem, st, (const bf16*)qkv, (bf16*)out, lse, tab2,
                                  alpha, g, units));
  SCOT_LAUNCH_CHECK();
  return 0;
}

template <int WS, int HD, int NWARP>
int launch_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, const float* tab2, const float* alpha,
               void* dqkv, float* partial, size_t partial_bytes, float* dtab, float* dalpha, float* g_qbias,
               float* g_vbias, WinGeom g, int total_windows, cudaStream_t st, const ScotAttnBwdFork* fk) {
  using C1 = DqCfg<WS, HD, NWARP>;
  using C2 = DkvCfg<WS, HD, NWARP>;
  static bool done = false;
  if (!done) {
    SCOT_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel<WS, HD, NWARP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C1::smem));
    SCOT_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel<WS, HD, NWARP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C1::smem));
    SCOT_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel<WS, HD, NWARP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C2::smem));
    SCOT_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel<WS, HD, NWARP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C2::smem));
    done = true;
  }
  void (*k1)(const bf16*, const bf16*, const bf16*, const float*, const float*, const float*, bf16*, float*, float*, float*,
             WinGeom, int, int) = g.shift > 0 ? attn_bwd_dq_kernel<WS, HD, NWARP, true> : attn_bwd_dq_kernel<WS, HD, NWARP, false>;
  auto k2 = g.shift > 0 ? attn_bwd_dkv_kernel<WS, HD, NWARP, true> : attn_bwd_dkv_kernel<WS, HD, NWARP, false>;
  // dq kernel: ~200 KB of smem -> one CTA per SM, so size the grid to a single wave; fewer chunks also means
  // fewer bias-gradient dumps for the second-stage reduction
  const int iters = ceil_div(total_windows, C1::WPI);
  int chunks = num_sms() / (g.heads * C1::RG);
  if (chunks > iters) chunks = iters;
  if (chunks < 1) chunks = 1;
  int wpc = ceil_div(iters, chunks) * C1::WPI;
  chunks = ceil_div(total_windows, wpc);
  (void)partial;
  (void)partial_bytes;
  // dk/dv kernel: size the grid to ONE wave of what is actually resident (two CTAs per SM at 128 registers x 256
  // threads; the occupancy API accounts for registers and shared memory) — a 1.3-wave grid costs a second, mostly
  // empty pass over the SMs
  static int occ2 = 0;
  if (occ2 == 0) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, attn_bwd_dkv_kernel<WS, HD, NWARP, false>, NWARP * 32, C2::smem) != cudaSuccess || nb < 1)
      nb = 1;
    int nb_s = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb_s, attn_bwd_dkv_kernel<WS, HD, NWARP, true>, NWARP * 32, C2::smem) != cudaSuccess || nb_s < 1)
      nb_s = 1;
    occ2 = nb < nb_s ? nb : nb_s;
  }
  int chunks2 = (occ2 * num_sms()) / (g.heads * C2::KG);
  if (chunks2 > iters) chunks2 = iters;
  if (chunks2 < 1) chunks2 = 1;
  const int wpc2 = ceil_div(iters, chunks2) * C2::WPI;
  chunks2 = ceil_div(total_windows, wpc2);
  dim3 grid1(g.heads, C1::RG, chunks);
  dim3 grid2(g.heads, C2::KG, chunks2);
  // The two kernels are independent (disjoint column ranges of dqkv, each recomputes P from the saved LSE). With a fork
  // descriptor the dk/dv kernel goes to a second stream and runs beside the dq kernel: at the small-window stages both are
  // latency-bound grids that leave most of the machine idle.
  if (fk != nullptr) {
    SCOT_CHECK_CUDA(cudaEventRecord(fk->fork, st));
    SCOT_CHECK_CUDA(cudaStreamWaitEvent(fk->stream, fk->fork, 0));
    SCOT_CHECK_CUDA(scot_launch_pdl(k2, grid2, dim3(NWARP * 32), C2::smem, fk->stream, (const bf16*)qkv, (const bf16*)o,
                                    (const bf16*)d_o, lse, tab2, alpha, (bf16*)dqkv, g_vbias, g, total_windows, wpc2));
    SCOT_LAUNCH_CHECK();
    SCOT_CHECK_CUDA(cudaEventRecord(fk->join, fk->stream));
  }
  SCOT_CHECK_CUDA(scot_launch_pdl(k1, grid1, dim3(NWARP * 32), C1::smem, st, (const bf16*)qkv, (const bf16*)o, (const bf16*)d_o, lse,
                                  tab2, alpha, (bf16*)dqkv, dtab, dalpha, g_qbias, g, total_windows, wpc));
  SCOT_LAUNCH_CHECK();
  if (fk == nullptr) {
    SCOT_CHECK_CUDA(scot_launch_pdl(k2, grid2, dim3(NWARP * 32), C2::smem, st, (const bf16*)qkv, (const bf16*)o, (const bf16*)d_o, lse,
                                    tab2, alpha, (bf16*)dqkv, g_vbias, g, total_windows, wpc2));
    SCOT_LAUNCH_CHECK();
  }
  if (fk != nullptr) SCOT_CHECK_CUDA(cudaStreamWaitEvent(st, fk->join, 0));
  return 0;
}

}  // namespace

size_t scot_attn_bwd_partial_bytes(int ws, int heads, int total_windows) {
  // The bias gradient is folded onto the table inside the dq kernel: no global scratch is needed any more. The entry
  // point (and the `partial` argument of scot_attn_bwd) stay for ABI stability; a small non-zero size keeps callers simple.
  (void)ws; (void)heads; (void)total_windows;
  return 256;
}

int scot_cpb_fwd_launch(const ScotCpbTable* tab, const float* params, void* arena, cudaStream_t st) {
  SCOT_REQUIRE(tab && params && arena && tab->n >= 1 && tab->n <= SCOT_CPB_MAX_LAYERS, "cpb_fwd: bad table");
  int max_total = 0, max_rows = 0;
  for (int i = 0; i < tab->n; ++i) {
    const int rows = (2 * tab->layer[i].ws - 1) * (2 * tab->layer[i].ws - 1);
    const int t = rows * tab->layer[i].heads;
    max_total = t > max_total ? t : max_total;
    max_rows = rows > max_rows ? rows : max_rows;
    SCOT_REQUIRE(tab->layer[i].heads <= 32 && tab->layer[i].ws <= 16, "cpb: at most 32 heads / window 16");
  }
  (void)max_total;
  cpb_fwd_kernel<<<dim3(ceil_div(max_rows, 8), tab->n), 256, 0, st>>>(*tab, params, (uint8_t*)arena);
  SCOT_LAUNCH_CHECK();
  return 0;
}

int scot_cpb_bwd_launch(const ScotCpbTable* tab, const float* params, float* grads, void* arena, cudaStream_t st) {
  SCOT_REQUIRE(tab && params && grads && arena && tab->n >= 1 && tab->n <= SCOT_CPB_MAX_LAYERS, "cpb_bwd: bad table");
  int max_total = 0;
  for (int i = 0; i < tab->n; ++i) {
    const int t = (2 * tab->layer[i].ws - 1) * (2 * tab->layer[i].ws - 1) * tab->layer[i].heads;
    max_total = t > max_total ? t : max_total;
    SCOT_REQUIRE(tab->layer[i].heads <= 32 && tab->layer[i].ws <= 16, "cpb: at most 32 heads / window 16");
  }
  cpb_bwd_pre_kernel<<<dim3(ceil_div(max_total, 256), tab->n), 256, 0, st>>>(*tab, params, grads, (uint8_t*)arena);
  SCOT_LAUNCH_CHECK();
  // 16 row splits per layer: 4 and 8 (fewer atomics, longer row loops) measured the same on B200
  cpb_bwd_mlp_kernel<<<dim3(16, tab->n), 512, 0, st>>>(*tab, params, grads, (const uint8_t*)arena);
  SCOT_LAUNCH_CHECK();
  return 0;
}

#define ATTN_DISPATCH(WS_, HD_, CALL)                                                   \
  if (ws == WS_ && hd == HD_) { return CALL; }

int scot_attn_fwd_launch(const void* qkv, void* out, float* lse, const float* tab2, const float* alpha, int batch,
                         int res, int ws, int shift, int heads, int hd, cudaStream_t st) {
  SCOT_REQUIRE(qkv && out && lse && tab2 && alpha, "attn_fwd: null pointer");
  SCOT_REQUIRE(res % ws == 0 && (shift == 0 || shift == ws / 2), "attn_fwd: bad geometry res=%d ws=%d shift=%d", res, ws, shift);
  if (const size_t lo = scot_split_off())  // "parity" precision: fp32 attention on the split-bf16 tensors
    return scot_attn32_fwd_launch(qkv, out, lse, tab2, alpha, batch, res, ws, shift, heads, hd, lo, st);
  // 16 x 16 windows (stages 0 and 1 of every shipped model): tcgen05 / TMEM / TMA kernels (attention_tc.cu)
  if (ws == 16) return scot_attn_tc_fwd_launch(qkv, out, lse, tab2, alpha, batch, res, shift, heads, hd, st);
  WinGeom g{res, shift, res / ws, heads, heads * hd};
  const int tw = batch * g.nws * g.nws;
  ATTN_DISPATCH(8, 16, (launch_fwd<8, 16>(qkv, out, lse, tab2, alpha, g, tw, st)))
  ATTN_DISPATCH(8, 32, (launch_fwd<8, 32>(qkv, out, lse, tab2, alpha, g, tw, st)))
  ATTN_DISPATCH(8, 64, (launch_fwd<8, 64>(qkv, out, lse, tab2, alpha, g, tw, st)))
  ATTN_DISPATCH(4, 16, (launch_fwd<4, 16>(qkv, out, lse, tab2, alpha, g, tw, st)))
  ATTN_DISPATCH(4, 32, (launch_fwd<4, 32>(qkv, out, lse, tab2, alpha, g, tw, st)))
  ATTN_DISPATCH(4, 64, (launch_fwd<4, 64>(qkv, out, lse, tab2, alpha, g, tw, st)))
  SCOT_REQUIRE(false, "attn_fwd: unsupported window %d / head_dim %d (windows 16/8/4, head_dim 16/32/64)", ws, hd);
}

int scot_attn_bwd_launch(const void* qkv, const void* o, const void* d_o, const float* lse, const float* tab2,
                         const float* alpha, void* dqkv, float* partial, size_t partial_bytes, float* dtab, float* dalpha,
                         float* g_qbias, float* g_vbias, int batch, int res, int ws, int shift, int heads, int hd,
                         cudaStream_t st) {
  return scot_attn_bwd_launch2(qkv, o, d_o, lse, tab2, alpha, dqkv, partial, partial_bytes, dtab, dalpha, g_qbias, g_vbias,
                               batch, res, ws, shift, heads, hd, st, nullptr);
}

int scot_attn_bwd_launch2(const void* qkv, const void* o, const void* d_o, const float* lse, const float* tab2,
                          const float* alpha, void* dqkv, float* partial, size_t partial_bytes, float* dtab, float* dalpha,
                          float* g_qbias, float* g_vbias, int batch, int res, int ws, int shift, int heads, int hd,
                          cudaStream_t st, const ScotAttnBwdFork* fk) {
  SCOT_REQUIRE(qkv && o && d_o && lse && tab2 && alpha && dqkv && dtab && dalpha, "attn_bwd: null pointer");
  SCOT_REQUIRE(fk == nullptr || (fk->stream != nullptr && fk->fork != nullptr && fk->join != nullptr), "attn_bwd: bad fork descriptor");
  if (const size_t lo = scot_split_off())
    return scot_attn32_bwd_launch(qkv, o, d_o, lse, tab2, alpha, dqkv, dtab, dalpha, g_qbias, g_vbias, batch, res, ws, shift,
                                  heads, hd, lo, st);
  if (ws == 16) {
    // tcgen05 backward; head_dim 64 (4 x 32 KB of operand tiles beside the staging tiles) stays on the mma.sync kernels
    const int rc = scot_attn_tc_bwd_launch(qkv, o, d_o, lse, tab2, alpha, dqkv, dtab, dalpha, g_qbias, g_vbias, batch, res, shift,
                                           heads, hd, st);
    if (rc != -1) return rc;
  }
  WinGeom g{res, shift, res / ws, heads, heads * hd};
  const int tw = batch * g.nws * g.nws;
#define BWD_ARGS qkv, o, d_o, lse, tab2, alpha, dqkv, partial, partial_bytes, dtab, dalpha, g_qbias, g_vbias, g, tw, st, fk
  ATTN_DISPATCH(16, 64, (launch_bwd<16, 64, 4>(BWD_ARGS)))
  ATTN_DISPATCH(8, 16, (launch_bwd<8, 16, 8>(BWD_ARGS)))
  ATTN_DISPATCH(8, 32, (launch_bwd<8, 32, 8>(BWD_ARGS)))
  ATTN_DISPATCH(8, 64, (launch_bwd<8, 64, 8>(BWD_ARGS)))
  ATTN_DISPATCH(4, 16, (launch_bwd<4, 16, 8>(BWD_ARGS)))
  ATTN_DISPATCH(4, 32, (launch_bwd<4, 32, 8>(BWD_ARGS)))
  ATTN_DISPATCH(4, 64, (launch_bwd<4, 64, 8>(BWD_ARGS)))
#undef BWD_ARGS
  SCOT_REQUIRE(false, "attn_bwd: unsupported window %d / head_dim %d", ws, hd);
}

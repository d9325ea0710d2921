// tcgen05 / TMEM / TMA window attention for 16 x 16 windows (N = 256 tokens: stages 0 and 1 of every shipped model, 85 % of
// the attention time of Poseidon-B), forward and backward. Same arithmetic as attention.cu (HF modeling_swinv2.py:421-487,
// scOT/model.py:522-559); what changes is where the work runs:
//
//   * q / k / v (/ dO) tiles of one (window, head) arrive by TMA from the token-major [B, res, res, 3C] tensor as four
//     8 x 8-token boxes ("quadrants"; an 8 x 8 quadrant of a 16 x 16 window never wraps under the cyclic shift of 8, so
//     torch.roll + window_partition are four box coordinates). Token order inside a tile is therefore quadrant-major:
//     row r = 64*(2a + b) + 8*i' + j'  <->  window pixel (p, q) = (8a + i', 8b + j').
//   * S = q_hat k_hat^T and O = P v (forward), S, dP, dQ, dK, dV (backward) are tcgen05.mma with fp32 accumulators in
//     TMEM; softmax / dS run on registers loaded with tcgen05.ld (one thread = one accumulator row), P / dS are staged
//     once in shared memory as bf16 (128-byte-swizzled K-major tiles) and feed the second-stage MMAs — as K-major A
//     operand for O / dQ and, through an MN-major descriptor over the SAME bytes, as the transposed A operand of dK / dV.
//   * the shift mask is uniform per (row quadrant, column quadrant) block: one additive scalar per 64-column chunk.
//   * relative-position bias: table lookups with one per-thread base + compile-time immediates (index is additive).
//
// Forward: one CTA = 128 threads = one query half (128 rows) of one (window, head); two CTAs per SM overlap each other's
// TMA / MMA latencies. Backward: one CTA = one (window, head), see attn_tc_bwd_kernel.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "internal.h"

namespace {

constexpr float kLog2e = 1.4426950408889634f;
constexpr int kN = 256;    // tokens per window
constexpr int kTabN = 961;  // (2*16-1)^2
// The bias table is kept in shared memory with a row pitch of 40 floats: the 32 lanes of a warp hold 4 x 8 window pixels
// (i', j'), their lookups for one key column land at 40 i' + j' (mod 32) = 8 i' + j' -> 32 distinct banks (pitch 31 gives
// up to 4-way conflicts, measured: the forward kernel spent more than half of its time in them).
constexpr int kTabPitch = 40;

struct TcGeom {
  int res, shift, nws, heads, C;
};

template <int HD>
struct TcCfg {
  static constexpr int ROWB = HD * 2;                                     // bytes per token row of a q/k/v/dO tile
  static constexpr uint64_t LAYOUT = ROWB == 128 ? 2 : (ROWB == 64 ? 4 : 6);  // UMMA layout_type: SW128 / SW64 / SW32
  static constexpr int SBO = 8 * ROWB;                                    // pitch of 8-row groups
  static constexpr int QUADB = 64 * ROWB;                                 // one 8 x 8-token TMA box
  static constexpr int TILEB = 4 * QUADB;                                 // 256 tokens
};

// one lane of a converged warp (the branch around it stays warp-uniform, so ptxas keeps descriptors in uniform registers)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// shared-memory matrix descriptor with an explicit swizzle mode (see umma_smem_desc in common.cuh for the field layout)
__device__ __forceinline__ uint64_t umma_desc_sw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint64_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// window pixel (p, q) of tile row r (quadrant-major order)
__device__ __forceinline__ void row_pq(int r, int& p, int& q) {
  const int quad = r >> 6;
  p = ((quad >> 1) << 3) + ((r >> 3) & 7);
  q = ((quad & 1) << 3) + (r & 7);
}
__device__ __forceinline__ long token_of(const TcGeom& g, int bw, int p, int q) {
  const int nw = g.nws * g.nws;
  const int b = bw / nw, w = bw - b * nw;
  const int wi = w / g.nws, wj = w - wi * g.nws;
  int i = wi * 16 + p + g.shift;
  int j = wj * 16 + q + g.shift;
  if (i >= g.res) i -= g.res;
  if (j >= g.res) j -= g.res;
  return ((long)b * g.res + i) * g.res + j;
}
// box origin (j0, i0, b) of quadrant (a, bq) of window bw
__device__ __forceinline__ void quad_origin(const TcGeom& g, int bw, int a, int bq, int& j0, int& i0, int& b) {
  const int nw = g.nws * g.nws;
  b = bw / nw;
  const int w = bw - b * nw;
  const int wi = w / g.nws, wj = w - wi * g.nws;
  i0 = wi * 16 + 8 * a + g.shift;
  j0 = wj * 16 + 8 * bq + g.shift;
  if (i0 >= g.res) i0 -= g.res;
  if (j0 >= g.res) j0 -= g.res;
}
// region code of a quadrant (scOT/model.py:448-472 with window 16, shift 8: the split falls on the quadrant boundary)
__device__ __forceinline__ int quad_code(const TcGeom& g, int bw, int quad) {
  if (g.shift == 0) return 0;
  const int w = bw % (g.nws * g.nws);
  const int wi = w / g.nws, wj = w - wi * g.nws;
  const int hm = (wi == g.nws - 1) && (quad >> 1);
  const int wm = (wj == g.nws - 1) && (quad & 1);
  return hm | (wm << 1);
}

// The two integer divisions of the helpers above, taken once per (window, head) unit: image index, window row / column and
// the shifted-window edge flags (bit 0: last window row, bit 1: last window column; 0 without shift). The per-step / per-row
// code then derives region codes and token indices with a few logic ops.
struct WinPos {
  int b, wi, wj, edge;
};
__device__ __forceinline__ WinPos win_pos(const TcGeom& g, int bw) {
  const int nw = g.nws * g.nws;
  WinPos w;
  w.b = bw / nw;
  const int r = bw - w.b * nw;
  w.wi = r / g.nws;
  w.wj = r - w.wi * g.nws;
  w.edge = g.shift == 0 ? 0 : ((w.wi == g.nws - 1) ? 1 : 0) | ((w.wj == g.nws - 1) ? 2 : 0);
  return w;
}
__device__ __forceinline__ int quad_code(const WinPos& w, int quad) { return w.edge & (((quad >> 1) & 1) | ((quad & 1) << 1)); }
__device__ __forceinline__ long token_of(const TcGeom& g, const WinPos& w, int p, int q) {
  int i = w.wi * 16 + p + g.shift;
  int j = w.wj * 16 + q + g.shift;
  if (i >= g.res) i -= g.res;
  if (j >= g.res) j -= g.res;
  return ((long)w.b * g.res + i) * g.res + j;
}

// L2-normalise one token row in place (F.normalize eps 1e-12, HF:445). The row occupies ROWB contiguous bytes; the swizzle
// only permutes its 16-byte chunks, which a sum of squares / a uniform scale do not care about. Returns 1 / max(|x|, eps).
template <int HD>
__device__ __forceinline__ float normalize_row_inplace(uint32_t row) {  // row = shared-space byte address
  uint4 u[HD / 8];
  float ss = 0.f;
#pragma unroll
  for (int c = 0; c < HD / 8; ++c) {
    u[c] = lds128(row + 16 * c);
    const uint32_t w[4] = {u[c].x, u[c].y, u[c].z, u[c].w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack_bf16x2(w[k]);
      ss = fmaf(f.x, f.x, fmaf(f.y, f.y, ss));
    }
  }
  const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
  for (int c = 0; c < HD / 8; ++c) {
    const uint32_t w[4] = {u[c].x, u[c].y, u[c].z, u[c].w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack_bf16x2(w[k]);
      o[k] = pack_bf16x2(f.x * inv, f.y * inv);
    }
    sts128(row + 16 * c, o[0], o[1], o[2], o[3]);
  }
  return inv;
}

// bias-table offset of the key column with in-chunk index j (32-column chunk = rows i' = 4*(cc&1) + (j>>3), j' = j&7 of a
// quadrant): coloff = pn * 31 + qn is additive -> chunk part + compile-time part
__host__ __device__ constexpr int col_imm(int j) { return (j >> 3) * kTabPitch + (j & 7); }
__device__ __forceinline__ int chunk_off(int cc) {  // cc = 32-column chunk 0..7: quadrant cc>>1, upper / lower 4 rows
  const int quad = cc >> 1;
  return (((quad >> 1) << 3) + ((cc & 1) << 2)) * kTabPitch + ((quad & 1) << 3);
}

// =================================================================================================
// forward
// =================================================================================================
struct TcFwdArgs {
  CUtensorMap tm_qkv;  // [B, res, res, 3C] bf16, box {HD, 8, 8, 1}
  bf16* out;           // [tokens, C]
  float* lse;          // [units, 256] (row order p-major, as in attention.cu)
  const float* tab2;   // [961, heads]
  const float* alpha;  // [heads]
  TcGeom g;
  int items;           // units * 2
};

template <int HD>
struct FwdSmem {
  using Cfg = TcCfg<HD>;
  static constexpr int kBar = 0;                        // 2 mbarriers + tmem pointer
  static constexpr int kQ = 1024;                       // 1024-aligned from here on
  static constexpr int kK = kQ + 2 * Cfg::QUADB;
  static constexpr int kV = kK + Cfg::TILEB;
  static constexpr int kP = kV + Cfg::TILEB;            // 4 slabs x [128 rows x 128 B]
  static constexpr int kTotal = kP + 4 * 16384;
};

// 32 lanes x 32 columns of 32-bit, registers -> TMEM (inverse of tmem_ld_32x32)
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void sts_f32(uint32_t saddr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr) : "memory");
  return v;
}

// One CTA = 256 threads = one query half (128 rows x 256 keys) of one (window, head); warp = (TMEM lane quarter, column
// half). Pass 1 adds bias + mask, takes the row maximum and parks the biased scores back in TMEM (tcgen05.st); pass 2
// exponentiates, sums and stages P. The two warps of a row exchange maximum / sum through the (then dead) q tile.
template <int HD>
__global__ void __launch_bounds__(256, (FwdSmem<HD>::kTotal + 1024 + 5120 <= 113 * 1024) ? 2 : 1)
attn_tc_fwd_kernel(const __grid_constant__ TcFwdArgs a) {
  using Cfg = TcCfg<HD>;
  using SM = FwdSmem<HD>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar_tma = reinterpret_cast<uint64_t*>(smem + SM::kBar);
  uint64_t* bar_mma = bar_tma + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bar_tma + 2);
  __shared__ float stab[31 * kTabPitch];  // static: the compiler knows the address space (LDS with immediate offsets)
  uint8_t* sQ = smem + SM::kQ;
  uint8_t* sK = smem + SM::kK;
  uint8_t* sV = smem + SM::kV;
  uint8_t* sP = smem + SM::kP;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lq = warp & 3, ch = warp >> 2;
  const int row = lq * 32 + lane;  // accumulator row (query within this half)
  const TcGeom g = a.g;
  const uint32_t xch = smem_u32(sQ);  // exchange area [2][2][128] floats: max, sum per (column half, row)

  pdl_launch_dependents();
  if (tid == 0) {
    tma_prefetch_desc(&a.tm_qkv);
    mbar_init(bar_tma, 1);
    mbar_init(bar_mma, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<256>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();

  uint32_t ph_tma = 0, ph_mma = 0;
  int cur_head = -1;
  for (int item = blockIdx.x; item < a.items; item += gridDim.x) {
    // head-major item order: a CTA keeps its bias table for many items (head changes at most heads - 1 times)
    const int half = item & 1, nwin = a.items / (2 * g.heads);
    const int h = (item >> 1) / nwin, bw = (item >> 1) - h * nwin;
    const int unit = bw * g.heads + h;
    // ---- loads: q rows of this half (2 quadrants), all keys / values (4 quadrants each) ----
    if (tid == 0) {
      mbar_expect_tx(bar_tma, 2 * Cfg::QUADB + 2 * Cfg::TILEB);
#pragma unroll
      for (int quad = 0; quad < 4; ++quad) {
        int j0, i0, b;
        quad_origin(g, bw, quad >> 1, quad & 1, j0, i0, b);
        if ((quad >> 1) == half) tma_load_4d(sQ + (quad & 1) * Cfg::QUADB, &a.tm_qkv, bar_tma, h * HD, j0, i0, b);
        tma_load_4d(sK + quad * Cfg::QUADB, &a.tm_qkv, bar_tma, g.C + h * HD, j0, i0, b);
        tma_load_4d(sV + quad * Cfg::QUADB, &a.tm_qkv, bar_tma, 2 * g.C + h * HD, j0, i0, b);
      }
    }
    if (h != cur_head) {
      for (int i = tid; i < kTabN; i += 256) stab[(i / 31) * kTabPitch + (i % 31)] = a.tab2[i * g.heads + h];
      cur_head = h;
    }
    mbar_wait(bar_tma, ph_tma);
    ph_tma ^= 1u;
    // ---- cosine attention: normalise q (128 rows) and k (256 rows) in place ----
    if (tid < 128) normalize_row_inplace<HD>(smem_u32(sQ) + tid * Cfg::ROWB);
    normalize_row_inplace<HD>(smem_u32(sK) + tid * Cfg::ROWB);
    fence_proxy_async_smem();
    __syncthreads();
    // ---- S[128 x 256] = q_hat k_hat^T ----
    if (warp == 0 && elect_one()) {
      tc_fence_after();
      constexpr uint32_t idesc = umma_idesc_bf16(128, 256, 0, 0);
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk) {
        const uint64_t da = umma_desc_sw(smem_u32(sQ) + kk * 32, 16, Cfg::SBO, Cfg::LAYOUT);
        const uint64_t db = umma_desc_sw(smem_u32(sK) + kk * 32, 16, Cfg::SBO, Cfg::LAYOUT);
        umma_bf16(tmem_base, da, db, idesc, kk > 0 ? 1u : 0u);
      }
      umma_commit(bar_mma);
    }
    // per-row constants while the MMA runs
    const int m = half * 128 + row;  // tile row of this thread
    int pm, qm;
    row_pq(m, pm, qm);
    const float a2 = a.alpha[h] * kLog2e;
    const float* tb = stab + (pm * kTabPitch + qm + 15 * kTabPitch + 15);
    const WinPos wp = win_pos(g, bw);
    const int code_m = quad_code(wp, m >> 6);
    // this thread's columns: quadrants 2 ch and 2 ch + 1
    const float mt0 = (quad_code(wp, 2 * ch) != code_m) ? -200.0f * kLog2e : 0.f;
    const float mt1 = (quad_code(wp, 2 * ch + 1) != code_m) ? -200.0f * kLog2e : 0.f;
    const uint32_t trow = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(ch * 128);
    mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1u;
    tc_fence_after();
    // ---- pass 1: biased scores (log2 units) back to TMEM, row maximum ----
    float mx = -INFINITY;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      float s[32];
      tmem_ld_32x32(trow + cc * 32, s);
      tmem_ld_wait();
      const float* tk = tb - ((((ch << 3) + ((cc & 1) << 2)) * kTabPitch) + ((cc >> 1) << 3));
      const float mt = (cc >> 1) ? mt1 : mt0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        s[j] = fmaf(s[j], a2, tk[-col_imm(j)]) + mt;
        mx = fmaxf(mx, s[j]);
      }
      tmem_st_32x32(trow + cc * 32, s);
    }
    sts_f32(xch + (uint32_t)(ch * 128 + row) * 4u, mx);
    tmem_st_wait();
    __syncthreads();  // q tile is dead (S is complete): its first 2 KB serve as the exchange area
    mx = fmaxf(mx, lds_f32(xch + (uint32_t)((ch ^ 1) * 128 + row) * 4u));
    // ---- pass 2: P = exp2(. - max) as bf16 into the staging tile, row sum ----
    float l = 0.f;
    const uint32_t prow = smem_u32(sP) + (uint32_t)row * 128u + (uint32_t)ch * 32768u;
    const uint32_t swz = (uint32_t)(row & 7);
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      float s[32];
      tmem_ld_32x32(trow + cc * 32, s);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        s[j] = fast_exp2(s[j] - mx);
        l += s[j];
      }
      const uint32_t slab = prow + (uint32_t)(cc >> 1) * 16384u;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const uint32_t chunk = (uint32_t)((cc & 1) * 4 + c4);
        sts128(slab + ((chunk ^ swz) << 4), pack_bf16x2(s[8 * c4], s[8 * c4 + 1]), pack_bf16x2(s[8 * c4 + 2], s[8 * c4 + 3]),
               pack_bf16x2(s[8 * c4 + 4], s[8 * c4 + 5]), pack_bf16x2(s[8 * c4 + 6], s[8 * c4 + 7]));
      }
    }
    sts_f32(xch + 1024u + (uint32_t)(ch * 128 + row) * 4u, l);
    tc_fence_before();
    fence_proxy_async_smem();
    __syncthreads();
    // ---- O[128 x HD] = P v (accumulator reuses the first HD columns of the dead score tile) ----
    if (warp == 0 && elect_one()) {
      tc_fence_after();
      constexpr uint32_t idesc = umma_idesc_bf16(128, HD, 0, 1);
#pragma unroll
      for (int ks = 0; ks < 16; ++ks) {  // 16 keys per step
        const uint64_t da = umma_smem_desc(smem_u32(sP) + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024);
        const uint64_t db = umma_desc_sw(smem_u32(sV) + ks * 16 * Cfg::ROWB, Cfg::TILEB, Cfg::SBO, Cfg::LAYOUT);
        umma_bf16(tmem_base, da, db, idesc, ks > 0 ? 1u : 0u);
      }
      umma_commit(bar_mma);
    }
    const long tr = token_of(g, wp, pm, qm);
    l += lds_f32(xch + 1024u + (uint32_t)((ch ^ 1) * 128 + row) * 4u);
    mbar_wait(bar_mma, ph_mma);
    ph_mma ^= 1u;
    tc_fence_after();
    {
      // the two warps of a row split the HD output columns (head_dim 16: the first one takes them all)
      constexpr int W = HD >= 32 ? HD / 2 : HD;
      if (HD >= 32 || ch == 0) {
        const float il = 1.0f / l;
        const int c0 = HD >= 32 ? ch * W : 0;
        bf16* dst = a.out + tr * g.C + h * HD + c0;
        float o[32];
        if constexpr (W == 32) tmem_ld_32x32(tmem_base + ((uint32_t)(lq * 32) << 16) + c0, o);
        else tmem_ld_32x16(tmem_base + ((uint32_t)(lq * 32) << 16) + c0, o);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < W; c += 8)
          *reinterpret_cast<uint4*>(dst + c) =
              make_uint4(pack_bf16x2(o[c] * il, o[c + 1] * il), pack_bf16x2(o[c + 2] * il, o[c + 3] * il),
                         pack_bf16x2(o[c + 4] * il, o[c + 5] * il), pack_bf16x2(o[c + 6] * il, o[c + 7] * il));
        if (ch == 0) a.lse[(long)unit * kN + pm * 16 + qm] = mx + log2f(l);
      }
    }
    tc_fence_before();
    __syncthreads();  // TMEM and the operand tiles are free for the next item
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

// =================================================================================================
// backward. One CTA = 512 threads = one (window, head) at a time, persistent over units.
//
//   per unit:   TMA -> q, k, v, dO tiles (16 boxes);  delta_m = dO_m . O_m, lse_m from global;  q, k normalised in place
//   4 steps (query half h, key half kb), each on a 128 x 128 tile of the score matrix:
//     S = q_h k_kb^T, dP = dO_h v_kb^T                    tcgen05.mma -> TMEM (2 x 128 columns)
//     P = exp2(S a2 + bias + mask - lse), dS = P (dP - delta)   16 warps: warp = (lane quarter, 32-column group)
//     P, dS -> bf16 staging tiles (128B-swizzled, [2 slabs][128 rows][128 B])
//     dQ_h += dS k_kb (K-major A) ; dK_kb += dS^T q_h, dV_kb += P^T dO_h (MN-major A over the same staging bytes)
//     W[dp][qm][qn] += dS folded over the window-row index (bias gradient, see below) while the MMAs run
//   epilogue: dq = alpha (dQ - q_hat (q_hat . dQ)) / |q|, same for k, dv; logit-scale gradient from q_hat . dQ.
//
// Bias gradient: dtab[(pm-pn+15)*31 + (qm-qn+15)] = sum dS[m, n]. Per step each thread owns a disjoint set of partially
// folded slots W[pm - pn][qm][qn .. qn+3] (sum over the 8 window rows i' of the tile that share a row displacement): 8
// eight-byte reads of the staged dS tile + two float4 read-modify-writes of W, no atomics, no N x N accumulator. W
// (31 x 16 x 16 floats) lives in shared memory across the units of a head and is folded over (qm - qn) onto the 961 table
// entries when the CTA moves to another head.
// =================================================================================================
struct TcBwdArgs {
  CUtensorMap tm_qkv;  // [B, res, res, 3C] bf16, box {HD, 8, 8, 1}
  CUtensorMap tm_do;   // [B, res, res, C]
  CUtensorMap tm_o;    // [B, res, res, C]
  const float* lse;    // [units, 256]
  const float* tab2;
  const float* alpha;
  bf16* dqkv;          // [tokens, 3C]
  float* dtab;         // [961, heads] +=
  float* dalpha;       // [heads] +=
  float* g_qbias;      // [C] += column sums of dq (nullable)
  float* g_vbias;      // [C] += column sums of dv (nullable)
  TcGeom g;
  int units, nwin;
};

template <int HD>
struct BwdSmem {
  using Cfg = TcCfg<HD>;
  static constexpr int kBar = 0;                       // 3 mbarriers + tmem pointer
  static constexpr int kQ = 1024;
  static constexpr int kK = kQ + Cfg::TILEB;
  static constexpr int kV = kK + Cfg::TILEB;
  static constexpr int kDO = kV + Cfg::TILEB;
  static constexpr int kO = kDO + Cfg::TILEB;          // forward output (only for delta = rowsum(dO * O))
  static constexpr int kP = kO + Cfg::TILEB;           // [2 slabs][128 rows][128 B]
  static constexpr int kDS = kP + 32768;               // two buffers (alternate by step) of the same shape
  static constexpr int kTotal = kDS + 2 * 32768;
  static constexpr int kStatic = (31 * kTabPitch + 4 * 256 + 2 * HD + 1 + 31 * 256) * 4;  // table, row arrays, sums, W
};

// 16-byte chunk c of tile row r sits at physical chunk c ^ swz_of(r) (Swizzle<B,4,3> on the byte address)
template <int HD>
__device__ __forceinline__ uint32_t swz_of(int r) {
  constexpr int ROWB = HD * 2;
  return (uint32_t)(((r * ROWB) >> 7) & (ROWB / 16 - 1));
}
template <int HD>
__device__ __forceinline__ void load_row_deswizzled(float* v, uint32_t tile, int r) {  // tile = shared-space address
  constexpr int ROWB = HD * 2;
  const uint32_t sw = swz_of<HD>(r);
#pragma unroll
  for (int c = 0; c < HD / 8; ++c) {
    const uint4 u = lds128(tile + (uint32_t)r * ROWB + (((uint32_t)c ^ sw) << 4));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack_bf16x2(w[k]);
      v[8 * c + 2 * k] = f.x;
      v[8 * c + 2 * k + 1] = f.y;
    }
  }
}

// Bias-gradient fold of one staged 128 x 128 dS tile (see the kernel header). Thread = (hc, j_m, b_m | G8, b_n): it sums, over
// the 8 window rows i_m of the tile, the four dS values of key pixels (i_n = (i_m - G8) mod 8, j_n = 4 hc .. 4 hc + 3) of
// key quadrant column b_n -> row displacement G8 (i_m >= G8) or G8 - 8 (i_m < G8); lanes of a warp read 32 distinct banks.
template <int G8>
__device__ __forceinline__ void fold_ds(uint32_t sds, float* sW, int lane, int b_n, int dquad /* h - kb */) {
  const int hc = lane & 1, j_m = (lane >> 1) & 7, b_m = lane >> 4;
  float acc_hi[4] = {0.f, 0.f, 0.f, 0.f}, acc_lo[4] = {0.f, 0.f, 0.f, 0.f};
  const uint32_t base = sds + (uint32_t)(b_n * 16384 + (64 * b_m + j_m) * 128 + hc * 8);
#pragma unroll
  for (int i_m = 0; i_m < 8; ++i_m) {
    const int i_n = (i_m - G8) & 7;
    uint2 v;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y)
                 : "r"(base + (uint32_t)(i_m * 8 * 128) + (uint32_t)((i_n ^ j_m) << 4)));
    const float2 f01 = unpack_bf16x2(v.x), f23 = unpack_bf16x2(v.y);
    if (i_m >= G8) {
      acc_hi[0] += f01.x; acc_hi[1] += f01.y; acc_hi[2] += f23.x; acc_hi[3] += f23.y;
    } else {
      acc_lo[0] += f01.x; acc_lo[1] += f01.y; acc_lo[2] += f23.x; acc_lo[3] += f23.y;
    }
  }
  const int qmm = 8 * b_m + j_m, qn0 = 8 * b_n + 4 * hc;
  const int dpi = 8 * dquad + G8 + 15;  // table row of displacement pm - pn = 8 (h - kb) + G8
  float4* w_hi = reinterpret_cast<float4*>(sW + dpi * 256 + qmm * 16 + qn0);
  float4 t = *w_hi;
  t.x += acc_hi[0]; t.y += acc_hi[1]; t.z += acc_hi[2]; t.w += acc_hi[3];
  *w_hi = t;
  if (G8 > 0) {
    float4* w_lo = reinterpret_cast<float4*>(sW + (dpi - 8) * 256 + qmm * 16 + qn0);
    float4 t2 = *w_lo;
    t2.x += acc_lo[0]; t2.y += acc_lo[1]; t2.z += acc_lo[2]; t2.w += acc_lo[3];
    *w_lo = t2;
  }
}

template <int HD>
__global__ void __launch_bounds__(512, 1)
attn_tc_bwd_kernel(const __grid_constant__ TcBwdArgs a) {
  using Cfg = TcCfg<HD>;
  using SM = BwdSmem<HD>;
  constexpr int kTmemS = 0, kTmemDP = 128, kTmemDQ = 256, kTmemDK = 256 + 2 * HD, kTmemDV = 256 + 4 * HD;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar_tma = reinterpret_cast<uint64_t*>(smem + SM::kBar);  // q, k tiles have landed
  uint64_t* bar_s = bar_tma + 1;   // S / dP of the current step are in TMEM
  uint64_t* bar_o = bar_tma + 2;   // the dQ / dK / dV MMAs of the previous step have read the staging tiles
  uint64_t* bar_tma2 = bar_tma + 3;  // v, dO, O tiles have landed (prefetched under the previous unit's epilogue)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bar_tma + 4);
  // small hot arrays are static: the compiler then knows their address space (LDS / STS, immediate offsets)
  __shared__ float stab[31 * kTabPitch];
  __shared__ float s_invq[256], s_invk[256], s_lse[256], s_delta[256];
  __shared__ float s_col[2 * HD + 1];  // [0,HD) dq sums, [HD,2HD) dv sums, [2HD] dalpha
  __shared__ __align__(16) float sW[31 * 256];
  uint8_t* sQ = smem + SM::kQ;
  uint8_t* sK = smem + SM::kK;
  uint8_t* sV = smem + SM::kV;
  uint8_t* sDO = smem + SM::kDO;
  uint8_t* sO = smem + SM::kO;
  uint8_t* sP = smem + SM::kP;
  uint8_t* sDS0 = smem + SM::kDS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lq = warp & 3, cg = warp >> 2;           // TMEM lane quarter, 32-column group of the 128-wide tile
  const int m_local = lq * 32 + lane;                // accumulator row of this thread
  const TcGeom g = a.g;

  pdl_launch_dependents();
  if (tid == 0) {
    tma_prefetch_desc(&a.tm_qkv);
    tma_prefetch_desc(&a.tm_do);
    tma_prefetch_desc(&a.tm_o);
    mbar_init(bar_tma, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    mbar_init(bar_tma2, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_ptr_smem);
  for (int i = tid; i < 31 * 256; i += 512) sW[i] = 0.f;
  for (int i = tid; i < 2 * HD + 1; i += 512) s_col[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();

  // flush of the per-head accumulators (bias-table fold, bias column sums, logit-scale gradient)
  auto flush_head = [&](int h) {
    for (int e = tid; e < kTabN; e += 512) {
      const int dpi = e / 31, dq = e % 31 - 15;
      float acc = 0.f;
      for (int qm = 0; qm < 16; ++qm) {
        const int qn = qm - dq;
        if (qn >= 0 && qn < 16) acc += sW[dpi * 256 + qm * 16 + qn];
      }
      atomicAdd(a.dtab + e * g.heads + h, acc);
    }
    if (tid < HD) {
      if (a.g_qbias != nullptr) atomicAdd(a.g_qbias + h * HD + tid, s_col[tid]);
      if (a.g_vbias != nullptr) atomicAdd(a.g_vbias + h * HD + tid, s_col[HD + tid]);
    }
    if (tid == 0) atomicAdd(a.dalpha + h, s_col[2 * HD]);
    __syncthreads();
    for (int i = tid; i < 31 * 256; i += 512) sW[i] = 0.f;
    for (int i = tid; i < 2 * HD + 1; i += 512) s_col[i] = 0.f;
    __syncthreads();
  };

  auto load_qk = [&](int bw, int h) {  // one thread
    mbar_expect_tx(bar_tma, 2 * Cfg::TILEB);
#pragma unroll
    for (int quad = 0; quad < 4; ++quad) {
      int j0, i0, b;
      quad_origin(g, bw, quad >> 1, quad & 1, j0, i0, b);
      tma_load_4d(sQ + quad * Cfg::QUADB, &a.tm_qkv, bar_tma, h * HD, j0, i0, b);
      tma_load_4d(sK + quad * Cfg::QUADB, &a.tm_qkv, bar_tma, g.C + h * HD, j0, i0, b);
    }
  };
  auto load_vdo = [&](int bw, int h) {  // one thread
    mbar_expect_tx(bar_tma2, 3 * Cfg::TILEB);
#pragma unroll
    for (int quad = 0; quad < 4; ++quad) {
      int j0, i0, b;
      quad_origin(g, bw, quad >> 1, quad & 1, j0, i0, b);
      tma_load_4d(sV + quad * Cfg::QUADB, &a.tm_qkv, bar_tma2, 2 * g.C + h * HD, j0, i0, b);
      tma_load_4d(sDO + quad * Cfg::QUADB, &a.tm_do, bar_tma2, h * HD, j0, i0, b);
      tma_load_4d(sO + quad * Cfg::QUADB, &a.tm_o, bar_tma2, h * HD, j0, i0, b);
    }
  };

  uint32_t ph_tma = 0, ph_s = 0, ph_o = 0;
  int cur_head = -1;
  // head-major unit order, strided over the CTAs: a CTA changes head at most (heads - 1) times
  for (int u = blockIdx.x; u < a.units; u += gridDim.x) {
    const int h = u / a.nwin, bw = u - h * a.nwin;
    const int unit = bw * g.heads + h;
    const WinPos wp = win_pos(g, bw);
    if (h != cur_head) {
      if (cur_head >= 0) flush_head(cur_head);
      for (int i = tid; i < kTabN; i += 512) stab[(i / 31) * kTabPitch + (i % 31)] = a.tab2[i * g.heads + h];
      cur_head = h;
    }
    // ---- loads: (v, dO, O) were prefetched under the previous unit's epilogue, (q, k) right after it ----
    if (u == (int)blockIdx.x && tid == 0) {
      load_vdo(bw, h);
      load_qk(bw, h);
    }
    if (tid < 256) {
      int p, q;
      row_pq(tid, p, q);
      s_lse[tid] = a.lse[(long)unit * kN + p * 16 + q];
    }
    mbar_wait(bar_tma2, ph_tma);
    // delta_r = dO_r . O_r: both tiles carry the same swizzle, so chunk c of one row pairs with chunk c of the other
    if (tid >= 256) {
      const int r = tid - 256;
      float d = 0.f;
#pragma unroll
      for (int c = 0; c < HD / 8; ++c) {
        const uint4 x = lds128(smem_u32(sO) + r * Cfg::ROWB + 16 * c), y = lds128(smem_u32(sDO) + r * Cfg::ROWB + 16 * c);
        const uint32_t xw[4] = {x.x, x.y, x.z, x.w}, yw[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = unpack_bf16x2(xw[k]), e = unpack_bf16x2(yw[k]);
          d = fmaf(f.x, e.x, fmaf(f.y, e.y, d));
        }
      }
      s_delta[r] = d;
    }
    mbar_wait(bar_tma, ph_tma);
    ph_tma ^= 1u;
    if (tid < 256) s_invq[tid] = normalize_row_inplace<HD>(smem_u32(sQ) + tid * Cfg::ROWB);
    else s_invk[tid - 256] = normalize_row_inplace<HD>(smem_u32(sK) + (tid - 256) * Cfg::ROWB);
    fence_proxy_async_smem();
    __syncthreads();

    const float al = a.alpha[h], a2 = al * kLog2e;
    // S / dP of step 0
    if (warp == 0 && elect_one()) {
      tc_fence_after();
      constexpr uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk)
        umma_bf16(tmem_base + kTmemS, umma_desc_sw(smem_u32(sQ) + kk * 32, 16, Cfg::SBO, Cfg::LAYOUT),
                  umma_desc_sw(smem_u32(sK) + kk * 32, 16, Cfg::SBO, Cfg::LAYOUT), idesc, kk > 0 ? 1u : 0u);
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk)
        umma_bf16(tmem_base + kTmemDP, umma_desc_sw(smem_u32(sDO) + kk * 32, 16, Cfg::SBO, Cfg::LAYOUT),
                  umma_desc_sw(smem_u32(sV) + kk * 32, 16, Cfg::SBO, Cfg::LAYOUT), idesc, kk > 0 ? 1u : 0u);
      umma_commit(bar_s);
    }

#pragma unroll 1  // (fully unrolled the kernel is 150 KB of SASS: measured 23 % of the stall samples in instruction fetch)
    for (int step = 0; step < 4; ++step) {
      const int hh = step >> 1, kb = step & 1;
      uint8_t* sDS = sDS0 + (step & 1) * 32768;  // alternating buffers: the fold of step s reads while step s+1 writes
      // per-row / per-column-group constants of this step
      const int m = hh * 128 + m_local;
      int pm, qm;
      row_pq(m, pm, qm);
      const int quad_n = 2 * kb + (cg >> 1);
      const float* tk = stab + (pm * kTabPitch + qm + 15 * kTabPitch + 15) -
                        ((8 * kb + 4 * (cg & 1)) * kTabPitch + 8 * (cg >> 1));
      const float c0 = ((quad_code(wp, quad_n) != quad_code(wp, m >> 6)) ? -200.0f * kLog2e : 0.f) - s_lse[m];
      const float delta = s_delta[m];
      const uint32_t trow = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(cg * 32);
      mbar_wait(bar_s, ph_s);
      ph_s ^= 1u;
      tc_fence_after();
      float s[32], dp[32];
      tmem_ld_32x32(trow + kTmemS, s);
      tmem_ld_32x32(trow + kTmemDP, dp);
      tmem_ld_wait();
      uint32_t pp[16], dd[16];
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float p0 = fast_exp2(fmaf(s[j], a2, tk[-col_imm(j)]) + c0);
        const float p1 = fast_exp2(fmaf(s[j + 1], a2, tk[-col_imm(j + 1)]) + c0);
        pp[j >> 1] = pack_bf16x2(p0, p1);
        dd[j >> 1] = pack_bf16x2(p0 * (dp[j] - delta), p1 * (dp[j + 1] - delta));
      }
      // the staging tiles are free once the output MMAs of the previous step have completed
      if (step > 0) {
        mbar_wait(bar_o, ph_o);
        ph_o ^= 1u;
      }
      {
        const uint32_t off = (uint32_t)(cg >> 1) * 16384u + (uint32_t)m_local * 128u;
        const uint32_t swz = (uint32_t)(m_local & 7);
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const uint32_t chunk = ((uint32_t)((cg & 1) * 4 + c4) ^ swz) << 4;
          sts128(smem_u32(sP) + off + chunk, pp[4 * c4], pp[4 * c4 + 1], pp[4 * c4 + 2], pp[4 * c4 + 3]);
          sts128(smem_u32(sDS) + off + chunk, dd[4 * c4], dd[4 * c4 + 1], dd[4 * c4 + 2], dd[4 * c4 + 3]);
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncthreads();
      if (warp == 0 && elect_one()) {
        tc_fence_after();
        // S / dP of the next step first (their TMEM tiles have been drained), then the three output products
        if (step < 3) {
          const int nh = (step + 1) >> 1, nkb = (step + 1) & 1;
          constexpr uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
          const uint32_t qa = smem_u32(sQ) + nh * 2 * Cfg::QUADB, ka = smem_u32(sK) + nkb * 2 * Cfg::QUADB;
          const uint32_t da_ = smem_u32(sDO) + nh * 2 * Cfg::QUADB, va = smem_u32(sV) + nkb * 2 * Cfg::QUADB;
#pragma unroll
          for (int kk = 0; kk < HD / 16; ++kk)
            umma_bf16(tmem_base + kTmemS, umma_desc_sw(qa + kk * 32, 16, Cfg::SBO, Cfg::LAYOUT),
                      umma_desc_sw(ka + kk * 32, 16, Cfg::SBO, Cfg::LAYOUT), idesc, kk > 0 ? 1u : 0u);
#pragma unroll
          for (int kk = 0; kk < HD / 16; ++kk)
            umma_bf16(tmem_base + kTmemDP, umma_desc_sw(da_ + kk * 32, 16, Cfg::SBO, Cfg::LAYOUT),
                      umma_desc_sw(va + kk * 32, 16, Cfg::SBO, Cfg::LAYOUT), idesc, kk > 0 ? 1u : 0u);
          umma_commit(bar_s);
        }
        constexpr uint32_t idq = umma_idesc_bf16(128, HD, 0, 1);
        constexpr uint32_t idk = umma_idesc_bf16(128, HD, 1, 1);
        const uint32_t kt = smem_u32(sK) + kb * 2 * Cfg::QUADB;    // 128 keys of this step
        const uint32_t qt = smem_u32(sQ) + hh * 2 * Cfg::QUADB;    // 128 queries of this step
        const uint32_t dot = smem_u32(sDO) + hh * 2 * Cfg::QUADB;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {  // dQ_h += dS[128 q x 128 keys] k_kb: 16 keys per instruction
          const uint64_t da = umma_smem_desc(smem_u32(sDS) + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024);
          const uint64_t db = umma_desc_sw(kt + ks * 16 * Cfg::ROWB, Cfg::TILEB, Cfg::SBO, Cfg::LAYOUT);
          umma_bf16(tmem_base + kTmemDQ + hh * HD, da, db, idq, (kb > 0 || ks > 0) ? 1u : 0u);
        }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {  // dK_kb += dS^T q_h, dV_kb += P^T dO_h: 16 queries per instruction
          const uint64_t da_s = umma_smem_desc(smem_u32(sDS) + ks * 2048, 16384, 1024);
          const uint64_t da_p = umma_smem_desc(smem_u32(sP) + ks * 2048, 16384, 1024);
          const uint64_t dbq = umma_desc_sw(qt + ks * 16 * Cfg::ROWB, Cfg::TILEB, Cfg::SBO, Cfg::LAYOUT);
          const uint64_t dbo = umma_desc_sw(dot + ks * 16 * Cfg::ROWB, Cfg::TILEB, Cfg::SBO, Cfg::LAYOUT);
          umma_bf16(tmem_base + kTmemDK + kb * HD, da_s, dbq, idk, (hh > 0 || ks > 0) ? 1u : 0u);
          umma_bf16(tmem_base + kTmemDV + kb * HD, da_p, dbo, idk, (hh > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(bar_o);
      }
      // ---- bias-gradient fold of the staged dS tile (generic-proxy reads, concurrent with the MMAs) ----
      switch (warp & 7) {
        case 0: fold_ds<0>(smem_u32(sDS), sW, lane, warp >> 3, hh - kb); break;
        case 1: fold_ds<1>(smem_u32(sDS), sW, lane, warp >> 3, hh - kb); break;
        case 2: fold_ds<2>(smem_u32(sDS), sW, lane, warp >> 3, hh - kb); break;
        case 3: fold_ds<3>(smem_u32(sDS), sW, lane, warp >> 3, hh - kb); break;
        case 4: fold_ds<4>(smem_u32(sDS), sW, lane, warp >> 3, hh - kb); break;
        case 5: fold_ds<5>(smem_u32(sDS), sW, lane, warp >> 3, hh - kb); break;
        case 6: fold_ds<6>(smem_u32(sDS), sW, lane, warp >> 3, hh - kb); break;
        default: fold_ds<7>(smem_u32(sDS), sW, lane, warp >> 3, hh - kb); break;
      }
    }
    // ---- epilogue: all products of this unit are complete when the last output commit arrives ----
    mbar_wait(bar_o, ph_o);
    ph_o ^= 1u;
    tc_fence_after();
    const int u_next = u + (int)gridDim.x;
    // every MMA of this unit is complete: v, dO (and O) are dead -> fetch the next unit's under the epilogue
    if (tid == 0 && u_next < a.units) load_vdo(u_next % a.nwin, u_next / a.nwin);
    // warps 0-7: dQ (normalisation backward + bias column sums); warps 8-15: dK (normalisation backward), then dV (sums)
#pragma unroll 1
    for (int pass = 0; pass < (warp < 8 ? 1 : 2); ++pass) {
      const int blk = warp < 8 ? (warp >> 2) : (pass == 0 ? 2 + ((warp - 8) >> 2) : 4 + ((warp - 8) >> 2));
      const int r = (blk & 1) * 128 + lq * 32 + lane;  // tile row (query or key)
      const uint32_t tcol = (blk < 2 ? kTmemDQ : (blk < 4 ? kTmemDK : kTmemDV)) + (uint32_t)((blk & 1) * HD);
      float acc[32];  // HD <= 32 columns of this row (zero padded: the column-sum butterfly works on 32)
#pragma unroll
      for (int c = HD; c < 32; ++c) acc[c] = 0.f;
      if constexpr (HD >= 32) tmem_ld_32x32(tmem_base + ((uint32_t)(lq * 32) << 16) + tcol, acc);
      else tmem_ld_32x16(tmem_base + ((uint32_t)(lq * 32) << 16) + tcol, acc);
      tmem_ld_wait();
      int p, q;
      row_pq(r, p, q);
      const long tr = token_of(g, wp, p, q);
      bf16* dst = a.dqkv + tr * (3L * g.C) + (blk >> 1) * g.C + h * HD;
      if (blk < 4) {
        // through the normalisation: d x = (alpha acc - x_hat (x_hat . alpha acc)) / max(|x|, eps)
        float xh[HD];
        load_row_deswizzled<HD>(xh, smem_u32(blk < 2 ? sQ : sK), r);
        float proj = 0.f;
#pragma unroll
        for (int c = 0; c < HD; ++c) proj = fmaf(xh[c], acc[c], proj);
        const float sc = al * (blk < 2 ? s_invq[r] : s_invk[r]);
#pragma unroll
        for (int c = 0; c < HD; ++c) acc[c] = (acc[c] - xh[c] * proj) * sc;
        if (blk < 2) {
          const float dal = warp_sum(proj);  // d alpha = sum_m q_hat_m . (sum_n dS[m,n] k_hat_n)
          if (lane == 0) atomicAdd(&s_col[2 * HD], dal);
        }
      }
      uint32_t pk[16];
#pragma unroll
      for (int c = 0; c < HD; c += 2) pk[c >> 1] = pack_bf16x2(acc[c], acc[c + 1]);
#pragma unroll
      for (int c = 0; c < HD; c += 8)
        *reinterpret_cast<uint4*>(dst + c) = make_uint4(pk[c >> 1], pk[(c >> 1) + 1], pk[(c >> 1) + 2], pk[(c >> 1) + 3]);
      if (blk < 2 || blk >= 4) {
        // bias gradients = column sums of dq / dv as stored. Butterfly: after the step with lane distance d every lane keeps
        // the half of its columns selected by bit d of its index -> 31 shuffles, lane L ends with the sum of column L.
#pragma unroll
        for (int c = 0; c < HD; c += 2) {
          const float2 f = unpack_bf16x2(pk[c >> 1]);
          acc[c] = f.x;
          acc[c + 1] = f.y;
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
          const bool up = (lane & d) != 0;
#pragma unroll
          for (int i = 0; i < d; ++i) {
            const float keep = up ? acc[i + d] : acc[i];
            const float send = up ? acc[i] : acc[i + d];
            acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, d);
          }
        }
        if (lane < HD) atomicAdd(&s_col[(blk >= 4 ? HD : 0) + lane], acc[0]);
      }
    }
    tc_fence_before();
    __syncthreads();  // accumulators, q / k tiles and row arrays are free for the next unit
    if (tid == 0 && u_next < a.units) load_qk(u_next % a.nwin, u_next / a.nwin);
  }
  if (cur_head >= 0) flush_head(cur_head);
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// =================================================================================================
// host side
// =================================================================================================
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled g_encode_tc = nullptr;
int get_encode_tc() {
  if (g_encode_tc != nullptr) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  SCOT_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  SCOT_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
  g_encode_tc = reinterpret_cast<PFN_encodeTiled>(fn);
  return 0;
}
// token-major [B, res, res, ch] bf16 tensor, box = 8 x 8 tokens x hd channels
int make_tmap_tokens(CUtensorMap* tm, const void* ptr, int B, int res, int ch, int hd) {
  cuuint64_t dims[4] = {(cuuint64_t)ch, (cuuint64_t)res, (cuuint64_t)res, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)ch * 2, (cuuint64_t)res * ch * 2, (cuuint64_t)res * res * ch * 2};
  cuuint32_t box[4] = {(cuuint32_t)hd, 8, 8, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUtensorMapSwizzle sw = hd == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (hd == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = g_encode_tc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SCOT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(tokens) failed (%d): ptr=%p B=%d res=%d ch=%d hd=%d", (int)r, ptr, B, res, ch, hd);
  return 0;
}

int g_sms_tc = 0;
int num_sms_tc() {
  if (g_sms_tc == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms_tc, cudaDevAttrMultiProcessorCount, dev);
    if (g_sms_tc <= 0) g_sms_tc = 148;
  }
  return g_sms_tc;
}

template <int HD>
int launch_tc_fwd(const void* qkv, void* out, float* lse, const float* tab2, const float* alpha, int batch, TcGeom g,
                  cudaStream_t st) {
  using SM = FwdSmem<HD>;
  TcFwdArgs a;
  memset(&a, 0, sizeof(a));
  int rc = get_encode_tc();
  if (rc) return rc;
  rc = make_tmap_tokens(&a.tm_qkv, qkv, batch, g.res, 3 * g.C, HD);
  if (rc) return rc;
  a.out = (bf16*)out;
  a.lse = lse;
  a.tab2 = tab2;
  a.alpha = alpha;
  a.g = g;
  a.items = batch * g.nws * g.nws * g.heads * 2;
  constexpr int smem = SM::kTotal + 1024;
  constexpr int per_sm = smem + 5120 <= 113 * 1024 ? 2 : 1;  // + the static bias table
  static bool done = false;
  if (!done) {
    SCOT_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    done = true;
  }
  int grid = num_sms_tc() * per_sm;
  if (grid > a.items) grid = a.items;
  SCOT_CHECK_CUDA(scot_launch_pdl(attn_tc_fwd_kernel<HD>, dim3(grid), dim3(256), (size_t)smem, st, a));
  SCOT_LAUNCH_CHECK();
  return 0;
}


template <int HD>
int launch_tc_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, const float* tab2, const float* alpha,
                  void* dqkv, float* dtab, float* dalpha, float* g_qbias, float* g_vbias, int batch, TcGeom g, cudaStream_t st) {
  using SM = BwdSmem<HD>;
  TcBwdArgs a;
  memset(&a, 0, sizeof(a));
  int rc = get_encode_tc();
  if (rc) return rc;
  rc = make_tmap_tokens(&a.tm_qkv, qkv, batch, g.res, 3 * g.C, HD);
  if (rc) return rc;
  rc = make_tmap_tokens(&a.tm_do, d_o, batch, g.res, g.C, HD);
  if (rc) return rc;
  rc = make_tmap_tokens(&a.tm_o, o, batch, g.res, g.C, HD);
  if (rc) return rc;
  a.lse = lse;
  a.tab2 = tab2;
  a.alpha = alpha;
  a.dqkv = (bf16*)dqkv;
  a.dtab = dtab;
  a.dalpha = dalpha;
  a.g_qbias = g_qbias;
  a.g_vbias = g_vbias;
  a.g = g;
  a.nwin = batch * g.nws * g.nws;
  a.units = a.nwin * g.heads;
  constexpr int smem = SM::kTotal + 1024;
  static_assert(smem + SM::kStatic <= 227 * 1024, "attn_tc_bwd: shared memory budget");
  static bool done = false;
  if (!done) {
    SCOT_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_bwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    done = true;
  }
  int grid = num_sms_tc();
  if (grid > a.units) grid = a.units;
  SCOT_CHECK_CUDA(scot_launch_pdl(attn_tc_bwd_kernel<HD>, dim3(grid), dim3(512), (size_t)smem, st, a));
  SCOT_LAUNCH_CHECK();
  return 0;
}

}  // namespace

int scot_attn_tc_fwd_launch(const void* qkv, void* out, float* lse, const float* tab2, const float* alpha, int batch, int res,
                            int shift, int heads, int hd, cudaStream_t st) {
  SCOT_REQUIRE(res % 16 == 0 && (shift == 0 || shift == 8), "attn_tc_fwd: 16 x 16 windows, shift 0 or 8 (res %d shift %d)", res, shift);
  SCOT_REQUIRE((((uintptr_t)qkv) & 15) == 0 && (((uintptr_t)out) & 15) == 0, "attn_tc_fwd: 16-byte aligned tensors");
  TcGeom g{res, shift, res / 16, heads, heads * hd};
  switch (hd) {
    case 16: return launch_tc_fwd<16>(qkv, out, lse, tab2, alpha, batch, g, st);
    case 32: return launch_tc_fwd<32>(qkv, out, lse, tab2, alpha, batch, g, st);
    case 64: return launch_tc_fwd<64>(qkv, out, lse, tab2, alpha, batch, g, st);
  }
  SCOT_REQUIRE(false, "attn_tc_fwd: head_dim %d must be 16/32/64", hd);
}

// head_dim 64 needs 4 x 32 KB of operand tiles and does not fit beside the staging tiles: returns -1 (caller falls back)
int scot_attn_tc_bwd_launch(const void* qkv, const void* o, const void* d_o, const float* lse, const float* tab2,
                            const float* alpha, void* dqkv, float* dtab, float* dalpha, float* g_qbias, float* g_vbias,
                            int batch, int res, int shift, int heads, int hd, cudaStream_t st) {
  SCOT_REQUIRE(res % 16 == 0 && (shift == 0 || shift == 8), "attn_tc_bwd: 16 x 16 windows, shift 0 or 8 (res %d shift %d)", res, shift);
  SCOT_REQUIRE((((uintptr_t)qkv) & 15) == 0 && (((uintptr_t)d_o) & 15) == 0 && (((uintptr_t)o) & 15) == 0 &&
               (((uintptr_t)dqkv) & 15) == 0, "attn_tc_bwd: 16-byte aligned tensors");
  TcGeom g{res, shift, res / 16, heads, heads * hd};
  switch (hd) {
    case 16: return launch_tc_bwd<16>(qkv, o, d_o, lse, tab2, alpha, dqkv, dtab, dalpha, g_qbias, g_vbias, batch, g, st);
    case 32: return launch_tc_bwd<32>(qkv, o, d_o, lse, tab2, alpha, dqkv, dtab, dalpha, g_qbias, g_vbias, batch, g, st);
  }
  return -1;
}

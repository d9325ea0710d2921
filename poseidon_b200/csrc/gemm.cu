// tcgen05 GEMM for every Linear / 1x1-conv on the scOT hot path (forward, dgrad, wgrad).
//
//   D[m, n] = sum_k A(m, k) * B(n, k)      bf16 operands, fp32 accumulation in TMEM
//
// Replaces the cuBLAS calls behind nn.Linear in the reference:
//   Swinv2SelfAttention.query/key/value (HF modeling_swinv2.py:416-418), Swinv2SelfOutput.dense (:531),
//   Swinv2Intermediate.dense (:571), Swinv2Output.dense (:586), ScOTPatchMerging.reduction
//   (scOT/model.py:669), ScOTPatchUnmerging.upsample/mixup (:725-726), ConvNeXtBlock.pwconv1/2 (:186-190)
// and their autograd backward (dgrad: B operand MN-major = the same weight read transposed;
// wgrad: both operands MN-major with the token dimension as the reduction, split over CTAs).
//
// Structure: persistent kernel, one CTA per SM looping over 128 x BN output tiles, 320 threads:
//   warp 0      : TMA producer  (cp.async.bulk.tensor.2d, 128B swizzle, mbarrier complete_tx); the smem ring
//                 runs ahead across tile boundaries, so operand loads of tile i+1 overlap the epilogue of tile i
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16, cta_group::1) into one
//                 of TWO accumulator buffers in TMEM (tmem_full / tmem_empty mbarriers)
//   warps 2..9  : epilogue: prefetch the auxiliary global operand (residual / saved pre-activation) while the MMAs
//                 run, tcgen05.ld (32x32b) -> fp32 staging smem -> coalesced 128-bit global stores with the fused
//                 op (bias / GELU / GELU' / RMW / red.add)
// These GEMMs have K = 96..3072 and are HBM/epilogue bound, so the design goal is to hide every fixed latency
// (TMA round trip, TMEM allocation, store drain) behind the epilogue rather than to saturate the tensor pipe.
// Tails in M, N and K are handled by TMA out-of-bounds zero fill + predicated stores.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "internal.h"
#include "scot_b200.h"

namespace {

constexpr int BM = 128;       // UMMA M (cta_group::1)
constexpr int BK = 64;        // bf16 elements per k-block = one 128B swizzle atom
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 320;
constexpr int EPI_THREADS = 256;

struct EpiArgs {
  const float* bias;
  void* out0;
  long ld0;
  void* out1;
  long ld1;
  const void* aux;
  long ldaux;
  float* colsum;
  size_t lo_off;  // split-bf16 ("parity") mode: byte distance of the lo twins of the bf16 outputs / aux (0 = plain bf16)
};

// ---- fused epilogue on a float4 of accumulators at (row, col..col+3); col is a multiple of 4 ----------
// auxiliary global operand of the epilogue (loaded ahead of the math so that several rows are in flight)
template <int MODE>
__device__ __forceinline__ float4 epi_load_aux(const EpiArgs& ep, long row, int col) {
  if constexpr (MODE == SCOT_EPI_GELU_BWD) {
    const bf16* ap = reinterpret_cast<const bf16*>(ep.aux) + row * ep.ldaux + col;
    const uint2 h = *reinterpret_cast<const uint2*>(ap);
    uint2 l = make_uint2(0u, 0u);  // bf16 zero pairs
    if (ep.lo_off) l = *reinterpret_cast<const uint2*>(reinterpret_cast<const char*>(ap) + ep.lo_off);
    return make_float4(__uint_as_float(h.x), __uint_as_float(h.y), __uint_as_float(l.x), __uint_as_float(l.y));
  } else if constexpr (MODE == SCOT_EPI_ADD_F32_BF16) {
    return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(ep.aux) + row * ep.ldaux + col);
  } else {
    return make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// (gelu'(x), gelu(x)) as two bf16 in one 32-bit word (low = derivative, high = activation)
__device__ __forceinline__ float gelu_pack(float x) {
  float cdf, pdf;
  gelu_parts(x, cdf, pdf);
  return __uint_as_float(pack_bf16x2(fmaf(x, pdf, cdf), x * cdf));
}

template <int MODE>
__device__ __forceinline__ void epi_store(const EpiArgs& ep, long row, int col, float4 v, float4 aux, float4 bias,
                                          float4& csum) {
  if constexpr (MODE != SCOT_EPI_GELU) {
    v.x += bias.x; v.y += bias.y; v.z += bias.z; v.w += bias.w;
  }
  if constexpr (MODE == SCOT_EPI_BF16) {
    st_bf16x4(reinterpret_cast<bf16*>(ep.out0) + row * ep.ld0 + col, ep.lo_off, v.x, v.y, v.z, v.w);
  } else if constexpr (MODE == SCOT_EPI_F32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out0) + row * ep.ld0 + col) = v;
  } else if constexpr (MODE == SCOT_EPI_GELU) {
    // out0 = gelu_erf'(h) (saved for backward, may be NULL), out1 = gelu_erf(h), h = acc + bias.
    // `v` arrives here already transformed by gelu_pack(): each 32-bit word holds (gelu', gelu) as two bf16.
    if (ep.lo_off) {
      // split-bf16 mode: `v` is the raw pre-activation h; both outputs are written as hi + lo pairs
      float c[4], p[4];
      const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) gelu_parts_precise(x[k], c[k], p[k]);
      if (ep.out0 != nullptr)
        st_bf16x4(reinterpret_cast<bf16*>(ep.out0) + row * ep.ld0 + col, ep.lo_off, fmaf(x[0], p[0], c[0]), fmaf(x[1], p[1], c[1]),
                  fmaf(x[2], p[2], c[2]), fmaf(x[3], p[3], c[3]));
      st_bf16x4(reinterpret_cast<bf16*>(ep.out1) + row * ep.ld1 + col, ep.lo_off, x[0] * c[0], x[1] * c[1], x[2] * c[2], x[3] * c[3]);
      return;
    }
    const uint32_t w0 = __float_as_uint(v.x), w1 = __float_as_uint(v.y), w2 = __float_as_uint(v.z), w3 = __float_as_uint(v.w);
    const uint2 o_grad = make_uint2(__byte_perm(w0, w1, 0x5410), __byte_perm(w2, w3, 0x5410));  // low halves
    const uint2 o_act = make_uint2(__byte_perm(w0, w1, 0x7632), __byte_perm(w2, w3, 0x7632));   // high halves
    if (ep.out0 != nullptr)
      *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(ep.out0) + row * ep.ld0 + col) = o_grad;
    *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(ep.out1) + row * ep.ld1 + col) = o_act;
  } else if constexpr (MODE == SCOT_EPI_GELU_BWD) {
    // aux = gelu'(h) saved by the forward epilogue: dh = (dy W) * gelu'(h)
    float2 g01 = unpack_bf16x2(__float_as_uint(aux.x)), g23 = unpack_bf16x2(__float_as_uint(aux.y));
    if (ep.lo_off) {
      const float2 l01 = unpack_bf16x2(__float_as_uint(aux.z)), l23 = unpack_bf16x2(__float_as_uint(aux.w));
      g01.x += l01.x; g01.y += l01.y; g23.x += l23.x; g23.y += l23.y;
    }
    v.x *= g01.x; v.y *= g01.y; v.z *= g23.x; v.w *= g23.y;
    // column sums of what was stored
    const float4 r = st_bf16x4(reinterpret_cast<bf16*>(ep.out0) + row * ep.ld0 + col, ep.lo_off, v.x, v.y, v.z, v.w);
    csum.x += r.x; csum.y += r.y; csum.z += r.z; csum.w += r.w;
  } else if constexpr (MODE == SCOT_EPI_RMW_F32 || MODE == SCOT_EPI_ATOMIC_F32) {
    // "+=" on a fp32 tensor: the add is performed by the L2 (fire-and-forget red.add), so the SM never waits for
    // the old value. RMW_F32 (one writer per element) is deterministic, ATOMIC_F32 (split reduction) is not.
    float* p = reinterpret_cast<float*>(ep.out0) + row * ep.ld0 + col;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
  } else if constexpr (MODE == SCOT_EPI_ADD_F32_BF16) {
    // out0 (fp32) = acc + aux (fp32 residual); out1 (bf16 copy) = same value rounded
    v.x += aux.x; v.y += aux.y; v.z += aux.z; v.w += aux.w;
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out0) + row * ep.ld0 + col) = v;
    if (ep.out1 != nullptr) st_bf16x4(reinterpret_cast<bf16*>(ep.out1) + row * ep.ld1 + col, ep.lo_off, v.x, v.y, v.z, v.w);
  }
}

// =================================================================================================
// tcgen05 kernel
// =================================================================================================
template <int BN>
struct TileCfg {
  static constexpr int kAccCols = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;  // one accumulator buffer
  static constexpr int kTmemCols = 2 * kAccCols;                                           // double buffered
  static constexpr int kABytes = BM * BK * 2;  // 16 KB
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagePitch = BN + 4;  // floats; +4 keeps 128-bit smem accesses conflict free
  static constexpr int kStagingBytes = BM * kStagePitch * 4;
};

// One launch can carry up to kMaxGroup independent problems of the same tile configuration (the four weight-gradient
// GEMMs of a transformer block are issued together): tiles are numbered across the problems.
constexpr int kMaxGroup = 4;
// Split-bf16 ("parity") mode: the reduction runs over three sections of `kblocks_real` k-blocks each,
//   A_hi B_hi + A_hi B_lo + A_lo B_hi   (the lo x lo term is below fp32 resolution of the sum),
// i.e. the virtual k-block v maps to section v / kblocks_real, real k-block v % kblocks_real; kblocks_total = 3 * real.
struct TileProblem {
  CUtensorMap tmA, tmB, tmAlo, tmBlo;
  int M, N, kblocks_total, kblocks_per_split, tiles_m, tiles_n, tile_begin, kblocks_real;
  EpiArgs ep;
};
struct GroupArgs {
  int n, total_tiles;
  TileProblem p[kMaxGroup];
};

struct TileCoord {
  int pi, m0, n0, kb_begin, kb_end, col_block;
};
__device__ __forceinline__ TileCoord decode_tile(const GroupArgs& ga, int t, int bn) {
  TileCoord c;
  c.pi = 0;
#pragma unroll
  for (int i = 1; i < kMaxGroup; ++i)
    if (i < ga.n && t >= ga.p[i].tile_begin) c.pi = i;
  const TileProblem& P = ga.p[c.pi];
  const int lt = t - P.tile_begin;
  const int tiles_mn = P.tiles_m * P.tiles_n;
  const int split = lt / tiles_mn, rem = lt - split * tiles_mn;
  c.col_block = rem / P.tiles_m;  // m fastest: consecutive tiles walk down one column block of the output
  c.m0 = (rem - c.col_block * P.tiles_m) * BM;
  c.n0 = c.col_block * bn;
  c.kb_begin = split * P.kblocks_per_split;
  c.kb_end = min(c.kb_begin + P.kblocks_per_split, P.kblocks_total);
  return c;
}

template <int BN, int AMN, int BMN, int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, (BN <= 64 ? 2 : 1))
gemm_tc_kernel(const __grid_constant__ GroupArgs ga, int num_stages) {
  using Cfg = TileCfg<BN>;
  static_assert(BN % 64 == 0 || BMN == 0, "MN-major B needs 64-wide atoms");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [0,1024) barriers + tmem ptr ; 1024-aligned operand stages ; fp32 staging tile for the epilogue
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* tmem_full_bar = empty_bar + 8;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;  // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  uint8_t* tiles = smem + 1024;
  float* stage = reinterpret_cast<float*>(tiles + (size_t)num_stages * Cfg::kStageBytes);

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // every CTA owns a contiguous range of tiles (m fastest): it streams down one column block of the output, so the
  // B tile (weights) stays hot and per-column epilogue state (bias-gradient sums) is flushed at most twice
  const int total_tiles = ga.total_tiles;
  const int tiles_per_cta = (total_tiles + gridDim.x - 1) / gridDim.x;
  const int t_begin = blockIdx.x * tiles_per_cta;
  const int t_end = min(total_tiles, t_begin + tiles_per_cta);

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < ga.n; ++i) {
      tma_prefetch_desc(&ga.p[i].tmA);
      tma_prefetch_desc(&ga.p[i].tmB);
      if (ga.p[i].kblocks_total != ga.p[i].kblocks_real) {
        tma_prefetch_desc(&ga.p[i].tmAlo);
        tma_prefetch_desc(&ga.p[i].tmBlo);
      }
    }
    for (int s = 0; s < num_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], EPI_THREADS);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();  // everything above (barriers, TMEM, descriptor prefetch) overlapped the previous kernel's tail

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int it = 0;  // running k-block counter over all my tiles (ring position)
      for (int t = t_begin; t < t_end; ++t) {
        const TileCoord tc = decode_tile(ga, t, BN);
        const TileProblem& PP = ga.p[tc.pi];
        for (int vkb = tc.kb_begin; vkb < tc.kb_end; ++vkb, ++it) {
          const int sec = vkb / PP.kblocks_real, kb = vkb - sec * PP.kblocks_real;
          const CUtensorMap* tmA = sec == 2 ? &PP.tmAlo : &PP.tmA;
          const CUtensorMap* tmB = sec == 1 ? &PP.tmBlo : &PP.tmB;
          const int s = it % num_stages;
          const uint32_t ph = (uint32_t)(it / num_stages) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          mbar_expect_tx(&full_bar[s], Cfg::kStageBytes);
          uint8_t* sa = tiles + (size_t)s * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          const int k0 = kb * BK;
          if constexpr (AMN == 0) {
            tma_load_2d(sa, tmA, &full_bar[s], k0, tc.m0);  // box {64 k, 128 rows}
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d(sa + j * (BK * 128), tmA, &full_bar[s], tc.m0 + 64 * j, k0);
          }
          if constexpr (BMN == 0) {
            tma_load_2d(sb, tmB, &full_bar[s], k0, tc.n0);  // box {64 k, BN rows}
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * (BK * 128), tmB, &full_bar[s], tc.n0 + 64 * j, k0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, AMN, BMN);
      int it = 0, lt = 0;
      for (int t = t_begin; t < t_end; ++t, ++lt) {
        const TileCoord tc = decode_tile(ga, t, BN);
        const int nkb = tc.kb_end - tc.kb_begin;
        const int buf = lt & 1;
        mbar_wait(&tmem_empty_bar[buf], (((uint32_t)lt >> 1) & 1u) ^ 1u);  // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(buf * Cfg::kAccCols);
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % num_stages;
          const uint32_t ph = (uint32_t)(it / num_stages) & 1u;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(tiles + (size_t)s * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // K-major: advance 16 elements = 32 B inside the swizzle atom. MN-major: advance 16 k-rows = 2048 B.
            const uint64_t da = (AMN == 0) ? umma_smem_desc(sa + k * 32, 16, 1024)
                                           : umma_smem_desc(sa + k * 2048, BK * 128, 1024);
            const uint64_t db = (BMN == 0) ? umma_smem_desc(sb + k * 32, 16, 1024)
                                           : umma_smem_desc(sb + k * 2048, BK * 128, 1024);
            umma_bf16(tacc, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);  // frees this smem stage once the MMAs above have read it
        }
        umma_commit(&tmem_full_bar[buf]);  // accumulator complete
      }
    }
  } else {
    // ------------------------------ epilogue (8 warps) ------------------------------
    const int ew = warp - 2;             // 0..7
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int half = ew >> 2;            // which 32-column TMEM chunks (even / odd) this warp moves in phase 1
    const int et = ew * 32 + lane;       // 0..255
    constexpr int VPR = BN / 4;              // float4 per tile row
    constexpr int RPP = EPI_THREADS / VPR;   // rows per pass
    constexpr int NPASS = (BM + RPP - 1) / RPP;
    constexpr bool kHasAux = (MODE == SCOT_EPI_GELU_BWD || MODE == SCOT_EPI_ADD_F32_BF16);
    constexpr int NCHUNK = BN / 32;          // 32-column TMEM chunks; warps 2-5 take the even ones, 6-9 the odd ones
    const int cv = et % VPR;
    const int r0 = et / VPR;
    const bool active = r0 < RPP;
    float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
    int lt = 0;
    for (int t = t_begin; t < t_end; ++t, ++lt) {
      const TileCoord tc = decode_tile(ga, t, BN);
      const TileProblem& P = ga.p[tc.pi];
      const EpiArgs& ep = P.ep;
      const int M = P.M, N = P.N;
      const int m0 = tc.m0, n0 = tc.n0;
      const int col = n0 + cv * 4;
      const bool col_ok = active && col < N;
      // prefetch the auxiliary operand of this tile (its addresses do not depend on the accumulators)
      float4 aux[kHasAux ? NPASS : 1];
      if constexpr (kHasAux) {
#pragma unroll
        for (int p = 0; p < NPASS; ++p) {
          const int r = r0 + p * RPP;
          if (col_ok && r < BM && (long)m0 + r < M) aux[p] = epi_load_aux<MODE>(ep, (long)m0 + r, col);
        }
      }
      const int buf = lt & 1;
      mbar_wait(&tmem_full_bar[buf], ((uint32_t)lt >> 1) & 1u);
      tc_fence_after();
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");  // previous tile's phase 2 is done with `stage`
      {
        const int r = q * 32 + lane;
        const uint32_t tacc = tmem_base + (uint32_t)(buf * Cfg::kAccCols) + ((uint32_t)(q * 32) << 16);
        constexpr int MYCH = (NCHUNK + 1) / 2;  // chunks per warp (even ones for warps 2-5, odd ones for 6-9)
        float v[MYCH][32];
#pragma unroll
        for (int k = 0; k < MYCH; ++k) {
          const int ci = 2 * k + half;
          if (ci < NCHUNK) tmem_ld_32x32(tacc + (uint32_t)(ci * 32), v[k]);
        }
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < MYCH; ++k) {
          const int ci = 2 * k + half;
          if (ci < NCHUNK) {
            if constexpr (MODE == SCOT_EPI_GELU) {
              // 32 independent activations per thread: plenty of ILP for the transcendental path
              const int cbase = n0 + ci * 32;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float b = (ep.bias != nullptr && cbase + j < N) ? __ldg(ep.bias + cbase + j) : 0.f;
                v[k][j] = ep.lo_off ? v[k][j] + b : gelu_pack(v[k][j] + b);  // split mode: activation in phase 2 (fp32)
              }
            }
            float4* dst = reinterpret_cast<float4*>(stage + (size_t)r * Cfg::kStagePitch + ci * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[k][4 * j], v[k][4 * j + 1], v[k][4 * j + 2], v[k][4 * j + 3]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[buf]);  // this thread no longer reads the accumulator buffer
      asm volatile("bar.sync 2, %0;" ::"n"(EPI_THREADS) : "memory");  // staging tile complete
      if (col_ok) {
        const float4 bias = ep.bias != nullptr ? *reinterpret_cast<const float4*>(ep.bias + col)
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int p = 0; p < NPASS; ++p) {
          const int r = r0 + p * RPP;
          if (r < BM && (long)m0 + r < M) {
            const float4 v = *reinterpret_cast<const float4*>(stage + (size_t)r * Cfg::kStagePitch + cv * 4);
            epi_store<MODE>(ep, (long)m0 + r, col, v, kHasAux ? aux[kHasAux ? p : 0] : make_float4(0.f, 0.f, 0.f, 0.f), bias, csum);
          }
        }
        if constexpr (MODE == SCOT_EPI_GELU_BWD) {
          // column sums (bias gradient): flush when the next tile of this CTA is in a different column block / problem
          bool last_of_col = (t + 1 >= t_end);
          if (!last_of_col) {
            const TileCoord nx = decode_tile(ga, t + 1, BN);
            last_of_col = nx.pi != tc.pi || nx.col_block != tc.col_block;
          }
          if (last_of_col && ep.colsum != nullptr) {
            atomicAdd(ep.colsum + col + 0, csum.x);
            atomicAdd(ep.colsum + col + 1, csum.y);
            atomicAdd(ep.colsum + col + 2, csum.z);
            atomicAdd(ep.colsum + col + 3, csum.w);
            csum = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// =================================================================================================
// tcgen05 kernel, "async epilogue" variant for the bf16-output modes (BF16 / GELU / GELU_BWD)
//
// Same producer / MMA structure as gemm_tc_kernel with 128 x 64 tiles and two CTAs per SM, but NOTHING in the epilogue
// waits on a global-memory round trip: the auxiliary operand (saved gelu') arrives by TMA through its own 2-stage
// smem ring, and the outputs leave by TMA stores (cp.async.bulk.tensor, 128B-swizzled bf16 tiles) that drain while the
// next tile is computed. In the epilogue one thread owns one accumulator row (tcgen05.ld 32x32b), so bias / GELU /
// gelu' scaling are register-only; the bias-gradient column sums are taken from the bf16 tile in smem.
// =================================================================================================
struct AsyncArgs {
  CUtensorMap tmA, tmB, tmAux, tmOut0, tmOut1;
  int M, N, kblocks, tiles_m, tiles_n, total_tiles;
  const float* bias;
  float* colsum;
  int has_out0;
};

constexpr int ABN = 64;                       // tile width of the async-epilogue kernel
constexpr int kOutTileBytes = BM * ABN * 2;   // one 128 x 64 bf16 tile = 16 KB (aux ring: 128 B rows, 128B swizzle)
constexpr int kSlabBytes = 32 * 32 * 2;       // output staging slab of one epilogue warp: 32 rows x 64 B (SWIZZLE_64B)
constexpr int kAStageBytes = BM * BK * 2 + ABN * BK * 2;

template <int BMN, int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_async_epi_kernel(const __grid_constant__ AsyncArgs ga, int num_stages) {
  constexpr bool kHasAux = (MODE == SCOT_EPI_GELU_BWD);
  constexpr int kNumOut = (MODE == SCOT_EPI_GELU) ? 2 : 1;
  constexpr int kAccCols = 64, kTmemCols = 128;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* tmem_full_bar = empty_bar + 8;        // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint64_t* aux_full_bar = tmem_empty_bar + 2;    // [2]
  uint64_t* aux_empty_bar = aux_full_bar + 2;     // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(aux_empty_bar + 2);
  uint8_t* tiles = smem + 1024;
  uint8_t* aux_s = tiles + (size_t)num_stages * kAStageBytes;             // [2][16 KB] (GELU_BWD only)
  uint8_t* out_s = aux_s + (kHasAux ? 2 * kOutTileBytes : 0);             // [kNumOut][16 KB]

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = ga.total_tiles;
  const int tiles_per_cta = (total_tiles + gridDim.x - 1) / gridDim.x;
  const int t_begin = blockIdx.x * tiles_per_cta;
  const int t_end = min(total_tiles, t_begin + tiles_per_cta);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&ga.tmA);
    tma_prefetch_desc(&ga.tmB);
    tma_prefetch_desc(&ga.tmOut1);
    if (kHasAux) tma_prefetch_desc(&ga.tmAux);
    if (kNumOut == 2 || MODE != SCOT_EPI_GELU) tma_prefetch_desc(&ga.tmOut0);
    for (int s = 0; s < num_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], EPI_THREADS / 32);  // one arrival per epilogue warp
      mbar_init(&aux_full_bar[b], 1);
      mbar_init(&aux_empty_bar[b], EPI_THREADS / 32);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int it = 0, lt = 0;
      for (int t = t_begin; t < t_end; ++t, ++lt) {
        const int cb = t / ga.tiles_m;  // m fastest
        const int m0 = (t - cb * ga.tiles_m) * BM, n0 = cb * ABN;
        if constexpr (kHasAux) {
          const int as = lt & 1;
          mbar_wait_backoff(&aux_empty_bar[as], (((uint32_t)lt >> 1) & 1u) ^ 1u);
          mbar_expect_tx(&aux_full_bar[as], kOutTileBytes);
          tma_load_2d(aux_s + as * kOutTileBytes, &ga.tmAux, &aux_full_bar[as], n0, m0);
        }
        for (int kb = 0; kb < ga.kblocks; ++kb, ++it) {
          const int s = it % num_stages;
          const uint32_t ph = (uint32_t)(it / num_stages) & 1u;
          mbar_wait_backoff(&empty_bar[s], ph ^ 1u);
          mbar_expect_tx(&full_bar[s], kAStageBytes);
          uint8_t* sa = tiles + (size_t)s * kAStageBytes;
          uint8_t* sb = sa + BM * BK * 2;
          const int k0 = kb * BK;
          tma_load_2d(sa, &ga.tmA, &full_bar[s], k0, m0);
          if constexpr (BMN == 0) tma_load_2d(sb, &ga.tmB, &full_bar[s], k0, n0);
          else tma_load_2d(sb, &ga.tmB, &full_bar[s], n0, k0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, ABN, 0, BMN);
      int it = 0, lt = 0;
      for (int t = t_begin; t < t_end; ++t, ++lt) {
        const int buf = lt & 1;
        mbar_wait_backoff(&tmem_empty_bar[buf], (((uint32_t)lt >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(buf * kAccCols);
        for (int i = 0; i < ga.kblocks; ++i, ++it) {
          const int s = it % num_stages;
          const uint32_t ph = (uint32_t)(it / num_stages) & 1u;
          mbar_wait_backoff(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(tiles + (size_t)s * kAStageBytes);
          const uint32_t sb = sa + BM * BK * 2;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = umma_smem_desc(sa + k * 32, 16, 1024);
            const uint64_t db = (BMN == 0) ? umma_smem_desc(sb + k * 32, 16, 1024)
                                           : umma_smem_desc(sb + k * 2048, BK * 128, 1024);
            umma_bf16(tacc, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full_bar[buf]);
      }
    }
  } else {
    // ------------------------------ epilogue (8 independent warps) ------------------------------
    // Warp (q, half) owns the 32 rows x 32 columns of the tile it can read from TMEM and a private 2 KB staging slab per
    // output: it converts, stages, and bulk-stores its slab on its own (64 B rows, SWIZZLE_64B box), so the eight warps
    // never meet at a CTA barrier and drift apart — one warp's SFU-heavy conversion overlaps another's TMEM load / store.
    const int ew = warp - 2;             // 0..7
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int half = ew >> 2;            // which 32-column half of the tile this warp converts
    const int row = q * 32 + lane;       // accumulator row (TMEM lane) of this thread
    const uint32_t slab0 = smem_u32(out_s) + (uint32_t)ew * kSlabBytes;             // output 0 (or the only one)
    const uint32_t slab1 = slab0 + 8u * kSlabBytes;                                  // output 1 (GELU)
    const uint32_t srow = (uint32_t)lane * 64u;                                      // this thread's 64 B row in the slab
    const uint32_t sswz = (uint32_t)((lane >> 1) & 3);                               // SWIZZLE_64B: chunk ^= (row >> 1) & 3
    const uint32_t aux_row = smem_u32(aux_s) + (uint32_t)row * 128u;
    const uint32_t aswz = (uint32_t)(row & 7);
    float cs0 = 0.f, cs1 = 0.f;          // GELU_BWD: running column sums (columns 2p, 2p+1 of this warp's half; rows of parity lane>>4)
    float bias[(MODE == SCOT_EPI_GELU_BWD) ? 1 : 32];  // bias of this thread's 32 columns, reloaded per column block
    int bias_cb = -1;
    int lt = 0;
    for (int t = t_begin; t < t_end; ++t, ++lt) {
      const int cb = t / ga.tiles_m;
      const int m0 = (t - cb * ga.tiles_m) * BM, n0 = cb * ABN;
      const int buf = lt & 1;
      if constexpr (MODE != SCOT_EPI_GELU_BWD) {
        if (cb != bias_cb) {
          bias_cb = cb;
          const int cbase = n0 + half * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) bias[j] = (ga.bias != nullptr && cbase + j < ga.N) ? __ldg(ga.bias + cbase + j) : 0.f;
        }
      }
      mbar_wait(&tmem_full_bar[buf], ((uint32_t)lt >> 1) & 1u);
      tc_fence_after();
      float v[32];
      tmem_ld_32x32(tmem_base + (uint32_t)(buf * kAccCols + half * 32) + ((uint32_t)(q * 32) << 16), v);
      // the previous tile's bulk store(s) of this warp must have finished reading the slab (they had a whole tile time)
      if (lane == 0) bulk_wait_read<0>();
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);  // the MMA warp may start the tile after next in this accumulator
      if constexpr (MODE == SCOT_EPI_GELU) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // 8 columns = one 16-byte chunk of each output row
          uint32_t g4[4], a4[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float g0, d0, g1, d1;
            gelu_and_grad(v[j * 8 + 2 * k] + bias[j * 8 + 2 * k], g0, d0);
            gelu_and_grad(v[j * 8 + 2 * k + 1] + bias[j * 8 + 2 * k + 1], g1, d1);
            g4[k] = pack_bf16x2(d0, d1);  // gelu'
            a4[k] = pack_bf16x2(g0, g1);  // gelu
          }
          const uint32_t off = srow + ((((uint32_t)j) ^ sswz) << 4);
          sts128(slab0 + off, g4[0], g4[1], g4[2], g4[3]);
          sts128(slab1 + off, a4[0], a4[1], a4[2], a4[3]);
        }
      } else if constexpr (MODE == SCOT_EPI_BF16) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t o4[4];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            o4[k] = pack_bf16x2(v[j * 8 + 2 * k] + bias[j * 8 + 2 * k], v[j * 8 + 2 * k + 1] + bias[j * 8 + 2 * k + 1]);
          sts128(slab0 + srow + ((((uint32_t)j) ^ sswz) << 4), o4[0], o4[1], o4[2], o4[3]);
        }
      } else {  // GELU_BWD: dh = acc * gelu'(h), gelu'(h) from the aux ring
        const int as = lt & 1;
        mbar_wait(&aux_full_bar[as], ((uint32_t)lt >> 1) & 1u);
        uint4 a[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) a[j] = lds128(aux_row + (uint32_t)(as * kOutTileBytes) + ((((uint32_t)(half * 4 + j)) ^ aswz) << 4));
        __syncwarp();
        if (lane == 0) mbar_arrive(&aux_empty_bar[as]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t w[4] = {a[j].x, a[j].y, a[j].z, a[j].w};
          uint32_t o4[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 g = unpack_bf16x2(w[k]);
            o4[k] = pack_bf16x2(v[j * 8 + 2 * k] * g.x, v[j * 8 + 2 * k + 1] * g.y);
          }
          sts128(slab0 + srow + ((((uint32_t)j) ^ sswz) << 4), o4[0], o4[1], o4[2], o4[3]);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      const int nc = n0 + half * 32, mr = m0 + q * 32;
      if (lane == 0 && nc < ga.N && mr < ga.M) {  // the box is clipped at the tensor edges by the TMA unit
        if constexpr (MODE == SCOT_EPI_GELU) {
          if (ga.has_out0) tma_store_2d(&ga.tmOut0, reinterpret_cast<const void*>(out_s + (size_t)ew * kSlabBytes), nc, mr);
          tma_store_2d(&ga.tmOut1, reinterpret_cast<const void*>(out_s + (size_t)(8 + ew) * kSlabBytes), nc, mr);
        } else {
          tma_store_2d(&ga.tmOut0, reinterpret_cast<const void*>(out_s + (size_t)ew * kSlabBytes), nc, mr);
        }
        bulk_commit();
      }
      if constexpr (MODE == SCOT_EPI_GELU_BWD) {
        // bias gradient: column sums of the bf16 values just staged (rows >= M hold zeros: their A rows were zero-filled).
        // Lane = (column pair p, row parity): even / odd rows sit in different bank halves, so the reads are conflict-free.
        const uint32_t p = (uint32_t)(lane & 15), par = (uint32_t)(lane >> 4);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint32_t r = 2u * (uint32_t)i + par;
          const float2 f = unpack_bf16x2(lds32(slab0 + r * 64u + ((((p >> 2)) ^ ((r >> 1) & 3u)) << 4) + (p & 3u) * 4u));
          cs0 += f.x;
          cs1 += f.y;
        }
        const bool last_of_col = (t + 1 >= t_end) || ((t + 1) / ga.tiles_m != cb);
        if (last_of_col) {
          cs0 += __shfl_xor_sync(0xffffffffu, cs0, 16);
          cs1 += __shfl_xor_sync(0xffffffffu, cs1, 16);
          if (ga.colsum != nullptr && lane < 16) {
            const int c = nc + 2 * (int)p;
            if (c < ga.N) atomicAdd(ga.colsum + c, cs0);
            if (c + 1 < ga.N) atomicAdd(ga.colsum + c + 1, cs1);
          }
          cs0 = cs1 = 0.f;
        }
      }
    }
    if (lane == 0) bulk_wait<0>();  // all stores of this warp are performed before the grid can complete
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// =================================================================================================
// SIMT reference kernel (bring-up / cross-check path; same epilogue semantics, fp32 FMA on CUDA cores)
// =================================================================================================
template <int MODE>
__global__ void gemm_simt_kernel(const bf16* __restrict__ A, long lda, int amn, const bf16* __restrict__ B, long ldb,
                                 int bmn, int M, int N, int K, EpiArgs ep, size_t op_lo) {
  const int col = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const long row = (long)blockIdx.y * blockDim.y + threadIdx.y;
  if (row >= M || col >= N) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k = 0; k < K; ++k) {
    const float a = ld_bf16(amn ? A + (long)k * lda + row : A + row * lda + k, op_lo);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float b = ld_bf16(bmn ? B + (long)k * ldb + col + j : B + (long)(col + j) * ldb + k, op_lo);
      acc[j] = fmaf(a, b, acc[j]);
    }
  }
  float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 bias = ep.bias != nullptr ? *reinterpret_cast<const float4*>(ep.bias + col) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 vv = make_float4(acc[0], acc[1], acc[2], acc[3]);
  if constexpr (MODE == SCOT_EPI_GELU) {
    if (ep.lo_off) vv = make_float4(vv.x + bias.x, vv.y + bias.y, vv.z + bias.z, vv.w + bias.w);
    else vv = make_float4(gelu_pack(vv.x + bias.x), gelu_pack(vv.y + bias.y), gelu_pack(vv.z + bias.z), gelu_pack(vv.w + bias.w));
  }
  epi_store<MODE>(ep, row, col, vv, epi_load_aux<MODE>(ep, row, col), bias, csum);
  if constexpr (MODE == SCOT_EPI_GELU_BWD) {
    if (ep.colsum != nullptr) {
      atomicAdd(ep.colsum + col + 0, csum.x);
      atomicAdd(ep.colsum + col + 1, csum.y);
      atomicAdd(ep.colsum + col + 2, csum.z);
      atomicAdd(ep.colsum + col + 3, csum.w);
    }
  }
}

// =================================================================================================
// host side
// =================================================================================================
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled g_encode = nullptr;

int get_encode_fn() {
  if (g_encode != nullptr) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  SCOT_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  SCOT_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
  g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
  return 0;
}

// inner = contiguous dimension (elements), outer = strided dimension, ld = stride of outer in elements
int make_tmap(CUtensorMap* tm, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
              uint32_t box_outer, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SCOT_REQUIRE(r == CUDA_SUCCESS,
               "cuTensorMapEncodeTiled failed (%d): ptr=%p inner=%llu outer=%llu ld=%llu box=%ux%u", (int)r, ptr,
               (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld, box_inner, box_outer);
  return 0;
}

int g_num_sms = 0;

struct HostProblem {
  const void* A;
  long lda;
  const void* B;
  long ldb;
  int M, N, K;
  EpiArgs ep;
};

template <int BN, int AMN, int BMN, int MODE>
int launch_tc_group(const HostProblem* hp, int n, cudaStream_t stream) {
  using Cfg = TileCfg<BN>;
  SCOT_REQUIRE(n >= 1 && n <= kMaxGroup, "gemm group: 1..%d problems", kMaxGroup);
  GroupArgs ga;
  memset(&ga, 0, sizeof(ga));
  ga.n = n;
  // tiles of all problems first (to size the K splits of the reduction-heavy wgrad mode for the whole group)
  int base_tiles = 0;
  for (int i = 0; i < n; ++i) base_tiles += ceil_div(hp[i].M, BM) * ceil_div(hp[i].N, BN);
  int total = 0;
  for (int i = 0; i < n; ++i) {
    TileProblem& P = ga.p[i];
    const HostProblem& h = hp[i];
    int rc;
    if (AMN == 0) rc = make_tmap(&P.tmA, h.A, (uint64_t)h.K, (uint64_t)h.M, (uint64_t)h.lda, BK, BM);
    else rc = make_tmap(&P.tmA, h.A, (uint64_t)h.M, (uint64_t)h.K, (uint64_t)h.lda, 64, BK);
    if (rc) return rc;
    if (BMN == 0) rc = make_tmap(&P.tmB, h.B, (uint64_t)h.K, (uint64_t)h.N, (uint64_t)h.ldb, BK, BN);
    else rc = make_tmap(&P.tmB, h.B, (uint64_t)h.N, (uint64_t)h.K, (uint64_t)h.ldb, 64, BK);
    if (rc) return rc;
    const size_t lo = h.ep.lo_off;
    if (lo != 0) {
      const char* Alo = reinterpret_cast<const char*>(h.A) + lo;
      const char* Blo = reinterpret_cast<const char*>(h.B) + lo;
      if (AMN == 0) rc = make_tmap(&P.tmAlo, Alo, (uint64_t)h.K, (uint64_t)h.M, (uint64_t)h.lda, BK, BM);
      else rc = make_tmap(&P.tmAlo, Alo, (uint64_t)h.M, (uint64_t)h.K, (uint64_t)h.lda, 64, BK);
      if (rc) return rc;
      if (BMN == 0) rc = make_tmap(&P.tmBlo, Blo, (uint64_t)h.K, (uint64_t)h.N, (uint64_t)h.ldb, BK, BN);
      else rc = make_tmap(&P.tmBlo, Blo, (uint64_t)h.N, (uint64_t)h.K, (uint64_t)h.ldb, 64, BK);
      if (rc) return rc;
    }
    P.M = h.M;
    P.N = h.N;
    P.tiles_m = ceil_div(h.M, BM);
    P.tiles_n = ceil_div(h.N, BN);
    P.kblocks_real = ceil_div(h.K, BK);
    P.kblocks_total = P.kblocks_real * (lo != 0 ? 3 : 1);
    int splits = 1;
    if (MODE == SCOT_EPI_ATOMIC_F32) {
      // split the (long) reduction so that the whole group has about two tiles of work per SM
      splits = (2 * g_num_sms) / base_tiles;
      if (splits < 1) splits = 1;
      if (splits > P.kblocks_total) splits = P.kblocks_total;
    } else if (MODE == SCOT_EPI_RMW_F32) {
      // "+=" epilogue (red.add): the reduction can be split as well. Deep stages have few output tiles but K up to
      // 3072, i.e. >1 MB streamed through a single SM per tile; keep at least 4 k-blocks per split.
      splits = (2 * g_num_sms) / base_tiles;
      if (splits > P.kblocks_total / 4) splits = P.kblocks_total / 4;
      if (splits < 1) splits = 1;
    }
    P.kblocks_per_split = ceil_div(P.kblocks_total, splits);
    splits = ceil_div(P.kblocks_total, P.kblocks_per_split);  // no empty split
    P.tile_begin = total;
    P.ep = h.ep;
    total += P.tiles_m * P.tiles_n * splits;
  }
  ga.total_tiles = total;

  // BN <= 64: two CTAs per SM (two independent epilogue pipelines hide each other's latencies)
  constexpr int kCtasPerSm = BN <= 64 ? 2 : 1;
  // smem: barriers + operand ring + dedicated fp32 staging tile (the ring keeps running during the epilogue)
  const size_t fixed = 1024 /*align slack*/ + 1024 /*barriers*/ + (size_t)Cfg::kStagingBytes;
  const size_t budget = (size_t)(227 * 1024) / kCtasPerSm - (kCtasPerSm > 1 ? 1024 : 0);
  int stages = (int)((budget - fixed) / Cfg::kStageBytes);
  if (stages > 6) stages = 6;
  SCOT_REQUIRE(stages >= 2, "gemm: tile too large for shared memory");
  const size_t smem = fixed + (size_t)stages * Cfg::kStageBytes;
  auto kern = gemm_tc_kernel<BN, AMN, BMN, MODE>;
  static bool attr_done = false;  // per instantiation
  if (!attr_done) {
    SCOT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_done = true;
  }
  const int max_ctas = g_num_sms * kCtasPerSm;
  const int grid = total < max_ctas ? total : max_ctas;
  SCOT_CHECK_CUDA(scot_launch_pdl(kern, dim3(grid), dim3(GEMM_THREADS), smem, stream, ga, stages));
  SCOT_LAUNCH_CHECK();
  return 0;
}

template <int BN, int AMN, int BMN, int MODE>
int launch_tc(const void* A, long lda, const void* B, long ldb, int M, int N, int K, const EpiArgs& ep,
              cudaStream_t stream) {
  HostProblem h{A, lda, B, ldb, M, N, K, ep};
  return launch_tc_group<BN, AMN, BMN, MODE>(&h, 1, stream);
}

// SCOT_GEMM_ASYNC_EPI=0 falls back to the register/staging epilogue of gemm_tc_kernel for the bf16-output modes (A/B runs)
bool async_epi_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SCOT_GEMM_ASYNC_EPI");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

template <int BMN, int MODE>
int launch_async(const void* A, long lda, const void* B, long ldb, int M, int N, int K, const EpiArgs& ep,
                 cudaStream_t stream) {
  AsyncArgs ga;
  memset(&ga, 0, sizeof(ga));
  int rc = make_tmap(&ga.tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM);
  if (rc) return rc;
  if (BMN == 0) rc = make_tmap(&ga.tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, ABN);
  else rc = make_tmap(&ga.tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, BK);
  if (rc) return rc;
  if (MODE == SCOT_EPI_GELU) {
    if (ep.out0 != nullptr) {
      rc = make_tmap(&ga.tmOut0, ep.out0, (uint64_t)N, (uint64_t)M, (uint64_t)ep.ld0, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc) return rc;
    }
    rc = make_tmap(&ga.tmOut1, ep.out1, (uint64_t)N, (uint64_t)M, (uint64_t)ep.ld1, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
  } else {
    rc = make_tmap(&ga.tmOut0, ep.out0, (uint64_t)N, (uint64_t)M, (uint64_t)ep.ld0, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    ga.tmOut1 = ga.tmOut0;
  }
  if (MODE == SCOT_EPI_GELU_BWD) {
    rc = make_tmap(&ga.tmAux, ep.aux, (uint64_t)N, (uint64_t)M, (uint64_t)ep.ldaux, 64, BM);
    if (rc) return rc;
  }
  ga.M = M;
  ga.N = N;
  ga.kblocks = ceil_div(K, BK);
  ga.tiles_m = ceil_div(M, BM);
  ga.tiles_n = ceil_div(N, ABN);
  ga.total_tiles = ga.tiles_m * ga.tiles_n;
  ga.bias = ep.bias;
  ga.colsum = ep.colsum;
  ga.has_out0 = ep.out0 != nullptr;
  // staging tiles: one per output (+ the aux ring for GELU_BWD)
  const size_t staging = (MODE == SCOT_EPI_GELU ? 2 : 1) * (size_t)kOutTileBytes;
  const size_t fixed = 1024 /*align slack*/ + 1024 /*barriers*/ + (MODE == SCOT_EPI_GELU_BWD ? 2 * kOutTileBytes : 0) + staging;
  const size_t budget = (size_t)(227 * 1024) / 2 - 1024;
  int stages = (int)((budget - fixed) / kAStageBytes);
  if (stages > 4) stages = 4;
  SCOT_REQUIRE(stages >= 2, "gemm(async epilogue): shared memory budget");
  const size_t smem = fixed + (size_t)stages * kAStageBytes;
  auto kern = gemm_async_epi_kernel<BMN, MODE>;
  static bool attr_done = false;  // per instantiation
  if (!attr_done) {
    SCOT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 / 2));
    attr_done = true;
  }
  const int max_ctas = g_num_sms * 2;
  const int grid = ga.total_tiles < max_ctas ? ga.total_tiles : max_ctas;
  SCOT_CHECK_CUDA(scot_launch_pdl(kern, dim3(grid), dim3(GEMM_THREADS), smem, stream, ga, stages));
  SCOT_LAUNCH_CHECK();
  return 0;
}

// TMA needs 16-byte aligned bases and row pitches for every tensor it touches
bool tma_ok(const void* p, long ld) { return p == nullptr || ((((uintptr_t)p) & 15) == 0 && ld % 8 == 0); }

template <int AMN, int BMN, int MODE>
int dispatch_bn(const void* A, long lda, const void* B, long ldb, int M, int N, int K, const EpiArgs& ep,
                cudaStream_t stream) {
  // bf16-output modes: async-epilogue kernel (warp-private TMA stores, TMA-fed auxiliary operand), 128 x 64 tiles, two CTAs
  // per SM. 128 x 128 tiles at one CTA per SM were measured slower for every stage shape (GELU: 31.6 vs 14.9 us at 4096 x 1536 x 384)
  if constexpr (AMN == 0 && (MODE == SCOT_EPI_GELU || MODE == SCOT_EPI_GELU_BWD || MODE == SCOT_EPI_BF16)) {
    if (async_epi_enabled() && ep.lo_off == 0 && tma_ok(ep.out0, ep.ld0) && tma_ok(ep.out1, ep.ld1) && tma_ok(ep.aux, ep.ldaux) &&
        (MODE != SCOT_EPI_GELU || ep.out1 != nullptr) && (MODE != SCOT_EPI_GELU_BWD || ep.aux != nullptr) &&
        (BMN == 0 || N % 8 == 0)) {
      return launch_async<BMN, MODE>(A, lda, B, ldb, M, N, K, ep, stream);
    }
  }
  // epilogue-bound modes (two bf16 streams / transcendental math): 128 x 64 tiles, two resident CTAs per SM
  if constexpr (MODE == SCOT_EPI_GELU || MODE == SCOT_EPI_GELU_BWD) {
    if (N % 64 == 0 || N > 64) return launch_tc<64, AMN, BMN, MODE>(A, lda, B, ldb, M, N, K, ep, stream);
  }
  // small problems (deep stages: a few thousand tokens): 128-wide tiles would leave most SMs idle and make every CTA
  // stream K/64 x 32 KB through one SM's L2 port; the 64-wide tile doubles the CTA count (two per SM)
  if constexpr (MODE != SCOT_EPI_ATOMIC_F32) {
    if (ceil_div(M, BM) * ceil_div(N, 128) < g_num_sms && N > 64)
      return launch_tc<64, AMN, BMN, MODE>(A, lda, B, ldb, M, N, K, ep, stream);
  }
  if constexpr (BMN == 0) {
    if (N % 128 == 0) return launch_tc<128, AMN, BMN, MODE>(A, lda, B, ldb, M, N, K, ep, stream);
    if (N % 96 == 0) return launch_tc<96, AMN, BMN, MODE>(A, lda, B, ldb, M, N, K, ep, stream);
    if (N > 64) return launch_tc<128, AMN, BMN, MODE>(A, lda, B, ldb, M, N, K, ep, stream);
    return launch_tc<64, AMN, BMN, MODE>(A, lda, B, ldb, M, N, K, ep, stream);
  } else {
    if (N > 64) return launch_tc<128, AMN, BMN, MODE>(A, lda, B, ldb, M, N, K, ep, stream);
    return launch_tc<64, AMN, BMN, MODE>(A, lda, B, ldb, M, N, K, ep, stream);
  }
}

template <int MODE>
int launch_simt(const void* A, long lda, int amn, const void* B, long ldb, int bmn, int M, int N, int K,
                const EpiArgs& ep, cudaStream_t stream) {
  dim3 block(32, 8);
  dim3 grid(ceil_div(N, 32 * 4), ceil_div(M, 8));
  gemm_simt_kernel<MODE><<<grid, block, 0, stream>>>(reinterpret_cast<const bf16*>(A), lda, amn,
                                                     reinterpret_cast<const bf16*>(B), ldb, bmn, M, N, K, ep, ep.lo_off);
  SCOT_LAUNCH_CHECK();
  return 0;
}

}  // namespace

// Weight gradients of up to four Linear layers in ONE launch: dW_i[N_i, K_i] += dY_i[tok, N_i]^T X_i[tok, K_i]
// (both operands MN-major, split reduction over the tokens, fp32 red.add into the gradient buffer).
int scot_gemm_wgrad_group_launch(const ScotWgradProblem* probs, int n, int impl, cudaStream_t stream) {
  SCOT_REQUIRE(probs && n >= 1 && n <= kMaxGroup, "wgrad group: 1..%d problems", kMaxGroup);
  if (g_num_sms == 0) {
    int dev = 0;
    SCOT_CHECK_CUDA(cudaGetDevice(&dev));
    SCOT_CHECK_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  bool narrow = false, wide = false;
  for (int i = 0; i < n; ++i) {
    const ScotWgradProblem& q = probs[i];
    SCOT_REQUIRE(q.dY && q.X && q.dW && q.tokens > 0 && q.n_out > 0 && q.n_in > 0, "wgrad group: bad problem %d", i);
    SCOT_REQUIRE(q.n_in % 4 == 0 && q.ld_dy % 8 == 0 && q.ld_x % 8 == 0, "wgrad group: alignment (problem %d)", i);
    (q.n_in > 64 ? wide : narrow) = true;
  }
  if (impl == SCOT_GEMM_SIMT || (narrow && wide)) {  // cross-check path / mixed tile widths: one launch per problem
    for (int i = 0; i < n; ++i) {
      const ScotWgradProblem& q = probs[i];
      ScotEpilogue e{SCOT_EPI_ATOMIC_F32, nullptr, q.dW, q.ld_dw, nullptr, 0, nullptr, 0, nullptr};
      int rc = scot_gemm_launch(q.dY, q.ld_dy, 1, q.X, q.ld_x, 1, q.n_out, q.n_in, (int)q.tokens, &e, impl, stream);
      if (rc) return rc;
    }
    return 0;
  }
  int rc = get_encode_fn();
  if (rc) return rc;
  HostProblem hp[kMaxGroup];
  for (int i = 0; i < n; ++i) {
    const ScotWgradProblem& q = probs[i];
    hp[i] = HostProblem{q.dY, q.ld_dy, q.X, q.ld_x, q.n_out, q.n_in, (int)q.tokens,
                        EpiArgs{nullptr, q.dW, q.ld_dw, nullptr, 0, nullptr, 0, nullptr, scot_split_off()}};
  }
  if (wide) return launch_tc_group<128, 1, 1, SCOT_EPI_ATOMIC_F32>(hp, n, stream);
  return launch_tc_group<64, 1, 1, SCOT_EPI_ATOMIC_F32>(hp, n, stream);
}

int scot_gemm_launch(const void* A, long lda, int a_mn_major, const void* B, long ldb, int b_mn_major, int M, int N,
                     int K, const ScotEpilogue* e, int impl, cudaStream_t stream) {
  SCOT_REQUIRE(A && B && e && e->out0, "gemm: null pointer");
  SCOT_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: bad shape %d %d %d", M, N, K);
  SCOT_REQUIRE(N % 4 == 0, "gemm: N=%d must be a multiple of 4", N);
  SCOT_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "gemm: leading dimensions must be multiples of 8 (16 B rows)");
  SCOT_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0, "gemm: operands must be 16 B aligned");
  EpiArgs ep{e->bias, e->out0, e->ld0, e->out1, e->ld1, e->aux, e->ldaux, e->colsum, scot_split_off()};
  if (g_num_sms == 0) {
    int dev = 0;
    SCOT_CHECK_CUDA(cudaGetDevice(&dev));
    SCOT_CHECK_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  if (impl == SCOT_GEMM_SIMT) {
    switch (e->mode) {
#define SIMT_CASE(MD) \
  case MD: return launch_simt<MD>(A, lda, a_mn_major, B, ldb, b_mn_major, M, N, K, ep, stream);
      SIMT_CASE(SCOT_EPI_BF16)
      SIMT_CASE(SCOT_EPI_F32)
      SIMT_CASE(SCOT_EPI_GELU)
      SIMT_CASE(SCOT_EPI_GELU_BWD)
      SIMT_CASE(SCOT_EPI_RMW_F32)
      SIMT_CASE(SCOT_EPI_ATOMIC_F32)
      SIMT_CASE(SCOT_EPI_ADD_F32_BF16)
#undef SIMT_CASE
      default: SCOT_REQUIRE(false, "gemm: unknown epilogue mode %d", e->mode);
    }
  }
  int rc = get_encode_fn();
  if (rc) return rc;
  const int majors = (a_mn_major ? 2 : 0) | (b_mn_major ? 1 : 0);
#define TC_CASE(AM, BM_, MD) \
  if (majors == ((AM ? 2 : 0) | (BM_ ? 1 : 0)) && e->mode == MD) \
    return dispatch_bn<AM, BM_, MD>(A, lda, B, ldb, M, N, K, ep, stream);
  // forward (x @ W^T): both K-major
  TC_CASE(0, 0, SCOT_EPI_BF16)
  TC_CASE(0, 0, SCOT_EPI_F32)
  TC_CASE(0, 0, SCOT_EPI_GELU)
  TC_CASE(0, 0, SCOT_EPI_ADD_F32_BF16)
  // dgrad (dy @ W): weight read MN-major
  TC_CASE(0, 1, SCOT_EPI_BF16)
  TC_CASE(0, 1, SCOT_EPI_F32)
  TC_CASE(0, 1, SCOT_EPI_GELU_BWD)
  TC_CASE(0, 1, SCOT_EPI_RMW_F32)
  // wgrad (dy^T @ x): both MN-major, split reduction, fp32 red.add into the gradient buffer
  TC_CASE(1, 1, SCOT_EPI_ATOMIC_F32)
  TC_CASE(1, 1, SCOT_EPI_F32)
#undef TC_CASE
  SCOT_REQUIRE(false, "gemm: unsupported combination a_mn=%d b_mn=%d mode=%d", a_mn_major, b_mn_major, e->mode);
}

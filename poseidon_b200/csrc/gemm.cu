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
};

// ---- fused epilogue on a float4 of accumulators at (row, col..col+3); col is a multiple of 4 ----------
// auxiliary global operand of the epilogue (loaded ahead of the math so that several rows are in flight)
template <int MODE>
__device__ __forceinline__ float4 epi_load_aux(const EpiArgs& ep, long row, int col) {
  if constexpr (MODE == SCOT_EPI_GELU_BWD) {
    const uint2 h = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(ep.aux) + row * ep.ldaux + col);
    return make_float4(__uint_as_float(h.x), __uint_as_float(h.y), 0.f, 0.f);
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
    uint2 o = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(ep.out0) + row * ep.ld0 + col) = o;
  } else if constexpr (MODE == SCOT_EPI_F32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out0) + row * ep.ld0 + col) = v;
  } else if constexpr (MODE == SCOT_EPI_GELU) {
    // out0 = gelu_erf'(h) (saved for backward, may be NULL), out1 = gelu_erf(h), h = acc + bias.
    // `v` arrives here already transformed by gelu_pack(): each 32-bit word holds (gelu', gelu) as two bf16.
    const uint32_t w0 = __float_as_uint(v.x), w1 = __float_as_uint(v.y), w2 = __float_as_uint(v.z), w3 = __float_as_uint(v.w);
    const uint2 o_grad = make_uint2(__byte_perm(w0, w1, 0x5410), __byte_perm(w2, w3, 0x5410));  // low halves
    const uint2 o_act = make_uint2(__byte_perm(w0, w1, 0x7632), __byte_perm(w2, w3, 0x7632));   // high halves
    if (ep.out0 != nullptr)
      *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(ep.out0) + row * ep.ld0 + col) = o_grad;
    *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(ep.out1) + row * ep.ld1 + col) = o_act;
  } else if constexpr (MODE == SCOT_EPI_GELU_BWD) {
    // aux = gelu'(h) saved by the forward epilogue: dh = (dy W) * gelu'(h)
    const float2 g01 = unpack_bf16x2(__float_as_uint(aux.x)), g23 = unpack_bf16x2(__float_as_uint(aux.y));
    v.x *= g01.x; v.y *= g01.y; v.z *= g23.x; v.w *= g23.y;
    uint2 o = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(ep.out0) + row * ep.ld0 + col) = o;
    const float2 r01 = unpack_bf16x2(o.x), r23 = unpack_bf16x2(o.y);  // column sums of what was stored
    csum.x += r01.x; csum.y += r01.y; csum.z += r23.x; csum.w += r23.y;
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
    if (ep.out1 != nullptr) {
      uint2 o = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
      *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(ep.out1) + row * ep.ld1 + col) = o;
    }
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
struct TileProblem {
  CUtensorMap tmA, tmB;
  int M, N, kblocks_total, kblocks_per_split, tiles_m, tiles_n, tile_begin, pad_;
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
        const CUtensorMap* tmA = &ga.p[tc.pi].tmA;
        const CUtensorMap* tmB = &ga.p[tc.pi].tmB;
        for (int kb = tc.kb_begin; kb < tc.kb_end; ++kb, ++it) {
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
                v[k][j] = gelu_pack(v[k][j] + b);
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
constexpr int kOutTileBytes = BM * ABN * 2;   // one 128 x 64 bf16 tile = 16 KB (128 B rows, 128B swizzle)
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
      mbar_init(&tmem_empty_bar[b], EPI_THREADS);
      mbar_init(&aux_full_bar[b], 1);
      mbar_init(&aux_empty_bar[b], EPI_THREADS);
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
    // ------------------------------ epilogue (8 warps) ------------------------------
    const int ew = warp - 2;             // 0..7
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int half = ew >> 2;            // which 32-column half of the tile this warp converts
    const int et = ew * 32 + lane;       // 0..255
    const int row = q * 32 + lane;       // accumulator row (TMEM lane) of this thread
    const bool issuer = (et == 0);       // issues / tracks the bulk stores
    const uint32_t swz = (uint32_t)(row & 7);
    const uint32_t out_row = smem_u32(out_s) + (uint32_t)row * 128u;
    const uint32_t aux_row = smem_u32(aux_s) + (uint32_t)row * 128u;
    float cs0 = 0.f, cs1 = 0.f;          // GELU_BWD: running sums of columns 2*(et&31), +1 over rows (et>>5)*16..+15
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
      // the previous tile's bulk store must have finished reading the output tile(s) in smem (it had a whole tile time)
      if (issuer) bulk_wait_read<0>();
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[buf]);  // the MMA warp may start the tile after next in this accumulator
      if constexpr (MODE == SCOT_EPI_GELU) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // 8 columns = one 16-byte chunk of each output row
          uint32_t g4[4], a4[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float x0 = v[j * 8 + 2 * k] + bias[j * 8 + 2 * k], x1 = v[j * 8 + 2 * k + 1] + bias[j * 8 + 2 * k + 1];
            float c0, p0, c1, p1;
            gelu_parts(x0, c0, p0);
            gelu_parts(x1, c1, p1);
            g4[k] = pack_bf16x2(fmaf(x0, p0, c0), fmaf(x1, p1, c1));  // gelu'
            a4[k] = pack_bf16x2(x0 * c0, x1 * c1);                    // gelu
          }
          const uint32_t off = (((uint32_t)(half * 4 + j)) ^ swz) << 4;
          sts128(out_row + off, g4[0], g4[1], g4[2], g4[3]);
          sts128(out_row + kOutTileBytes + off, a4[0], a4[1], a4[2], a4[3]);
        }
      } else if constexpr (MODE == SCOT_EPI_BF16) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t o4[4];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            o4[k] = pack_bf16x2(v[j * 8 + 2 * k] + bias[j * 8 + 2 * k], v[j * 8 + 2 * k + 1] + bias[j * 8 + 2 * k + 1]);
          sts128(out_row + ((((uint32_t)(half * 4 + j)) ^ swz) << 4), o4[0], o4[1], o4[2], o4[3]);
        }
      } else {  // GELU_BWD: dh = acc * gelu'(h), gelu'(h) from the aux ring
        const int as = lt & 1;
        mbar_wait(&aux_full_bar[as], ((uint32_t)lt >> 1) & 1u);
        uint4 a[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) a[j] = lds128(aux_row + (uint32_t)(as * kOutTileBytes) + ((((uint32_t)(half * 4 + j)) ^ swz) << 4));
        mbar_arrive(&aux_empty_bar[as]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t w[4] = {a[j].x, a[j].y, a[j].z, a[j].w};
          uint32_t o4[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 g = unpack_bf16x2(w[k]);
            o4[k] = pack_bf16x2(v[j * 8 + 2 * k] * g.x, v[j * 8 + 2 * k + 1] * g.y);
          }
          sts128(out_row + ((((uint32_t)(half * 4 + j)) ^ swz) << 4), o4[0], o4[1], o4[2], o4[3]);
        }
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync 2, %0;" ::"n"(EPI_THREADS) : "memory");
      if (issuer) {
        if constexpr (MODE == SCOT_EPI_GELU) {
          if (ga.has_out0) tma_store_2d(&ga.tmOut0, out_s, n0, m0);
          tma_store_2d(&ga.tmOut1, out_s + kOutTileBytes, n0, m0);
        } else {
          tma_store_2d(&ga.tmOut0, out_s, n0, m0);
        }
        bulk_commit();
      }
      if constexpr (MODE == SCOT_EPI_GELU_BWD) {
        // bias gradient: column sums of the bf16 values just staged (rows >= M hold zeros: their A rows were zero-filled)
        const int cp = et & 31, rg = et >> 5;
        const uint32_t chunk = (uint32_t)(cp >> 2), word = (uint32_t)(cp & 3) * 4;
        const uint32_t obase = smem_u32(out_s) + word;
#pragma unroll
        for (int rr = 0; rr < 16; ++rr) {
          const int r = rg * 16 + rr;
          const float2 f = unpack_bf16x2(lds32(obase + (uint32_t)r * 128u + ((chunk ^ (uint32_t)(r & 7)) << 4)));
          cs0 += f.x;
          cs1 += f.y;
        }
        const bool last_of_col = (t + 1 >= t_end) || ((t + 1) / ga.tiles_m != cb);
        if (last_of_col && ga.colsum != nullptr) {
          const int c = n0 + 2 * cp;
          if (c < ga.N) atomicAdd(ga.colsum + c, cs0);
          if (c + 1 < ga.N) atomicAdd(ga.colsum + c + 1, cs1);
          cs0 = cs1 = 0.f;
        }
      }
    }
    if (issuer) bulk_wait<0>();  // all stores of this CTA are performed before the grid can complete
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// =================================================================================================
// "v2" of the async-epilogue kernel (SCOT_GEMM_ASYNC_V2=1; NOT yet validated on hardware — round-2 candidate).
//
// Motivation (profiles/r01_ncu_full_summary.md, stall sampling of the kernel above): the epilogue is a serial chain per
// CTA. 28 % (bf16) / 16 % (GELU) of the samples wait behind `cp.async.bulk.wait_group.read 0`, i.e. for the previous
// tile's TMA store to read its staging tile (the store queues behind the operand loads in the TMA unit), and all the
// epilogue arithmetic sits AFTER that wait. Changes, same producer / MMA warps and barriers otherwise:
//   * every mode: tcgen05.ld + all arithmetic first (results packed as bf16 in registers, accumulator buffer released),
//     THEN the wait for the staging tile, then only the 128-bit shared-memory stores;
//   * BF16 mode: two staging tiles used alternately, `wait_group.read 1` (the previous store may still be reading);
//   * GELU_BWD mode: no staging tile at all — the product is written in place into the aux-ring stage that delivered the
//     saved gelu' tile (same 128 x 64 bf16 128B-swizzled layout) and stored from there; the stage goes back to the
//     producer when that store has read it (aux_empty: one arrival by the issuer, at the start of the next tile, instead
//     of 256 by the readers). This frees 16 KB: three operand stages instead of two (K = 384 / 768 dgrads have 6 / 12
//     k-blocks per tile, two stages starve the MMA: 31 % of the v1 samples wait on the aux / accumulator barriers);
//   * GELU mode: bias kept in shared memory (one broadcast LDS.128 per four columns) instead of 32 registers per thread,
//     so that both packed outputs (32 registers) fit next to the accumulator row without spilling at 2 CTAs / SM.
// =================================================================================================
// Epilogue of the v2 kernels (eight warps, one thread per accumulator row and 32-column half), shared by
// gemm_async_epi2_kernel and gemm_async_smallk_kernel.
template <int MODE>
__device__ __forceinline__ void async_epilogue_v2(const AsyncArgs& ga, int warp, int lane, int t_begin, int t_end,
                                                  uint32_t tmem_base, uint64_t* tmem_full_bar, uint64_t* tmem_empty_bar,
                                                  uint64_t* aux_full_bar, uint64_t* aux_empty_bar, float* bias_s,
                                                  uint8_t* aux_s, uint8_t* out_s) {
  constexpr int kAccCols = 64;
  // ------------------------------ epilogue (8 warps) ------------------------------
  const int ew = warp - 2;             // 0..7
  const int q = warp & 3;              // TMEM lane quarter this warp may access
  const int half = ew >> 2;            // which 32-column half of the tile this warp converts
  const int et = ew * 32 + lane;       // 0..255
  const int row = q * 32 + lane;       // accumulator row (TMEM lane) of this thread
  const bool issuer = (et == 0);       // issues / tracks the bulk stores
  const uint32_t swz = (uint32_t)(row & 7);
  const uint32_t row_off = (uint32_t)row * 128u;
  float cs0 = 0.f, cs1 = 0.f;          // GELU_BWD: running sums of columns 2*(et&31), +1 over rows (et>>5)*16..+15
  int bias_cb = -1;
  int lt = 0;
  for (int t = t_begin; t < t_end; ++t, ++lt) {
    const int cb = t / ga.tiles_m;
    const int m0 = (t - cb * ga.tiles_m) * BM, n0 = cb * ABN;
    const int buf = lt & 1;
    if constexpr (MODE != SCOT_EPI_GELU_BWD && MODE != SCOT_EPI_RMW_F32) {
      if (cb != bias_cb) {  // block-uniform: every epilogue thread sees the same tile sequence
        bias_cb = cb;
        asm volatile("bar.sync 3, %0;" ::"n"(EPI_THREADS) : "memory");  // readers of the previous column block's bias are done
        if (et < ABN) bias_s[et] = (ga.bias != nullptr && n0 + et < ga.N) ? __ldg(ga.bias + n0 + et) : 0.f;
        asm volatile("bar.sync 3, %0;" ::"n"(EPI_THREADS) : "memory");
      }
    }
    mbar_wait(&tmem_full_bar[buf], ((uint32_t)lt >> 1) & 1u);
    tc_fence_after();
    float v[32];
    tmem_ld_32x32(tmem_base + (uint32_t)(buf * kAccCols + half * 32) + ((uint32_t)(q * 32) << 16), v);
    tmem_ld_wait();
    tc_fence_before();
    mbar_arrive(&tmem_empty_bar[buf]);  // the MMA warp may start the tile after next in this accumulator
    if constexpr (MODE == SCOT_EPI_GELU_BWD) {
      // hand the OTHER aux stage back to the producer as early as possible (it must be re-loaded for tile lt + 1 while
      // this tile is processed): its in-place result was stored at the end of the previous tile; only the issuer's warp
      // waits here for that store to have read the stage, the other seven warps go on with the arithmetic
      if (issuer && lt > 0) {
        bulk_wait_read<0>();
        mbar_arrive(&aux_empty_bar[buf ^ 1]);
      }
    }
    // ---- arithmetic first: results as packed bf16 in registers ----
    uint32_t o0[16];
    uint32_t o1[(MODE == SCOT_EPI_GELU) ? 16 : 1];
    uint32_t stage_base;  // shared-memory address of this thread's row in the tile that will be stored
    if constexpr (MODE == SCOT_EPI_GELU) {
      const float4* b4 = reinterpret_cast<const float4*>(bias_s + half * 32);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 bb = b4[j];
        const float xs[4] = {v[4 * j] + bb.x, v[4 * j + 1] + bb.y, v[4 * j + 2] + bb.z, v[4 * j + 3] + bb.w};
        float c[4], p[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) gelu_parts(xs[k], c[k], p[k]);
        o0[2 * j] = pack_bf16x2(fmaf(xs[0], p[0], c[0]), fmaf(xs[1], p[1], c[1]));      // gelu'
        o0[2 * j + 1] = pack_bf16x2(fmaf(xs[2], p[2], c[2]), fmaf(xs[3], p[3], c[3]));
        o1[2 * j] = pack_bf16x2(xs[0] * c[0], xs[1] * c[1]);                             // gelu
        o1[2 * j + 1] = pack_bf16x2(xs[2] * c[2], xs[3] * c[3]);
      }
      stage_base = smem_u32(out_s) + row_off;
    } else if constexpr (MODE == SCOT_EPI_BF16) {
      const float4* b4 = reinterpret_cast<const float4*>(bias_s + half * 32);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 bb = b4[j];
        o0[2 * j] = pack_bf16x2(v[4 * j] + bb.x, v[4 * j + 1] + bb.y);
        o0[2 * j + 1] = pack_bf16x2(v[4 * j + 2] + bb.z, v[4 * j + 3] + bb.w);
      }
      stage_base = smem_u32(out_s) + (uint32_t)(buf * kOutTileBytes) + row_off;
    } else if constexpr (MODE == SCOT_EPI_F32 || MODE == SCOT_EPI_RMW_F32) {
      // fp32 output: this thread's 32 columns are one 128-byte row of its half's 128 x 32 fp32 box (128B swizzle)
      if constexpr (MODE == SCOT_EPI_F32) {
        const float4* b4 = reinterpret_cast<const float4*>(bias_s + half * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bb = b4[j];
          v[4 * j] += bb.x; v[4 * j + 1] += bb.y; v[4 * j + 2] += bb.z; v[4 * j + 3] += bb.w;
        }
      }
      stage_base = smem_u32(out_s) + (uint32_t)(half * kOutTileBytes) + row_off;
    } else {  // GELU_BWD: dh = acc * gelu'(h); gelu'(h) sits in aux stage `buf`, the product replaces it in place
      mbar_wait(&aux_full_bar[buf], ((uint32_t)lt >> 1) & 1u);
      stage_base = smem_u32(aux_s) + (uint32_t)(buf * kOutTileBytes) + row_off;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 a = lds128(stage_base + ((((uint32_t)(half * 4 + j)) ^ swz) << 4));
        const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 g = unpack_bf16x2(w[k]);
          o0[4 * j + k] = pack_bf16x2(v[j * 8 + 2 * k] * g.x, v[j * 8 + 2 * k + 1] * g.y);
        }
      }
    }
    // ---- now the tile that receives the results must be free ----
    if constexpr (MODE == SCOT_EPI_BF16) {
      if (issuer) bulk_wait_read<1>();   // the store issued two tiles ago used this buffer; the last one may still read
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
    } else if constexpr (MODE == SCOT_EPI_GELU || MODE == SCOT_EPI_F32 || MODE == SCOT_EPI_RMW_F32) {
      if (issuer) bulk_wait_read<0>();
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
    }
    // GELU_BWD: every thread rewrites exactly the 64 bytes it has just read; no other thread touches them before bar.sync 2
    if constexpr (MODE == SCOT_EPI_F32 || MODE == SCOT_EPI_RMW_F32) {
#pragma unroll
      for (int j = 0; j < 8; ++j)  // eight 16-byte chunks = the whole 128-byte row of this half
        sts128(stage_base + ((((uint32_t)j) ^ swz) << 4), __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
               __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t off = (((uint32_t)(half * 4 + j)) ^ swz) << 4;
        sts128(stage_base + off, o0[4 * j], o0[4 * j + 1], o0[4 * j + 2], o0[4 * j + 3]);
        if constexpr (MODE == SCOT_EPI_GELU)
          sts128(stage_base + kOutTileBytes + off, o1[4 * j], o1[4 * j + 1], o1[4 * j + 2], o1[4 * j + 3]);
      }
    }
    fence_proxy_async_smem();
    asm volatile("bar.sync 2, %0;" ::"n"(EPI_THREADS) : "memory");
    if (issuer) {
      if constexpr (MODE == SCOT_EPI_GELU) {
        if (ga.has_out0) tma_store_2d(&ga.tmOut0, out_s, n0, m0);
        tma_store_2d(&ga.tmOut1, out_s + kOutTileBytes, n0, m0);
        bulk_commit();
      } else if constexpr (MODE == SCOT_EPI_BF16) {
        tma_store_2d(&ga.tmOut0, out_s + buf * kOutTileBytes, n0, m0);
        bulk_commit();
      } else if constexpr (MODE == SCOT_EPI_F32) {
        tma_store_2d(&ga.tmOut0, out_s, n0, m0);  // fp32 tensor map, box {32, 128}: one store per column half
        if (n0 + 32 < ga.N) tma_store_2d(&ga.tmOut0, out_s + kOutTileBytes, n0 + 32, m0);
        bulk_commit();
      } else if constexpr (MODE == SCOT_EPI_RMW_F32) {
        tma_reduce_add_2d(&ga.tmOut0, out_s, n0, m0);  // out += tile, the fp32 add is performed by the L2
        if (n0 + 32 < ga.N) tma_reduce_add_2d(&ga.tmOut0, out_s + kOutTileBytes, n0 + 32, m0);
        bulk_commit();
      } else {
        tma_store_2d(&ga.tmOut0, aux_s + buf * kOutTileBytes, n0, m0);
        bulk_commit();
      }
    }
    if constexpr (MODE == SCOT_EPI_GELU_BWD) {
      // bias gradient: column sums of the bf16 values just staged (rows >= M hold zeros: their A rows were zero-filled)
      const int cp = et & 31, rg = et >> 5;
      const uint32_t chunk = (uint32_t)(cp >> 2), word = (uint32_t)(cp & 3) * 4;
      const uint32_t obase = smem_u32(aux_s) + (uint32_t)(buf * kOutTileBytes) + word;
#pragma unroll
      for (int rr = 0; rr < 16; ++rr) {
        const int r = rg * 16 + rr;
        const float2 f = unpack_bf16x2(lds32(obase + (uint32_t)r * 128u + ((chunk ^ (uint32_t)(r & 7)) << 4)));
        cs0 += f.x;
        cs1 += f.y;
      }
      const bool last_of_col = (t + 1 >= t_end) || ((t + 1) / ga.tiles_m != cb);
      if (last_of_col && ga.colsum != nullptr) {
        const int c = n0 + 2 * cp;
        if (c < ga.N) atomicAdd(ga.colsum + c, cs0);
        if (c + 1 < ga.N) atomicAdd(ga.colsum + c + 1, cs1);
        cs0 = cs1 = 0.f;
      }
    }
  }
  if (issuer) bulk_wait<0>();  // all stores of this CTA are performed before the grid can complete
}

template <int BMN, int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_async_epi2_kernel(const __grid_constant__ AsyncArgs ga, int num_stages) {
  constexpr bool kHasAux = (MODE == SCOT_EPI_GELU_BWD);
  constexpr int kNumOut = (MODE == SCOT_EPI_GELU) ? 2 : 1;
  constexpr int kOutBufs = (MODE == SCOT_EPI_BF16) ? 2 : (MODE == SCOT_EPI_GELU_BWD ? 0 : 1);  // staging tiles per output
  static_assert(kOutBufs >= 0, "");
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
  float* bias_s = reinterpret_cast<float*>(smem + 512);                    // [64] bias of the current column block
  uint8_t* tiles = smem + 1024;
  uint8_t* aux_s = tiles + (size_t)num_stages * kAStageBytes;             // [2][16 KB] (GELU_BWD only; doubles as staging)
  uint8_t* out_s = aux_s + (kHasAux ? 2 * kOutTileBytes : 0);             // [kOutBufs][kNumOut][16 KB]

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
      mbar_init(&tmem_empty_bar[b], EPI_THREADS);
      mbar_init(&aux_full_bar[b], 1);
      mbar_init(&aux_empty_bar[b], 1);  // v2: released by the store issuer once the in-place result has been stored
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
    // ------------------------------ TMA producer (as in v1) ------------------------------
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
    // ------------------------------ MMA issuer (as in v1) ------------------------------
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
    async_epilogue_v2<MODE>(ga, warp, lane, t_begin, t_end, tmem_base, tmem_full_bar, tmem_empty_bar, aux_full_bar,
                            aux_empty_bar, bias_s, aux_s, out_s);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// =================================================================================================
// "v3" GELU kernel (SCOT_GEMM_ASYNC_V2=2; NOT yet validated on hardware — round-2 candidate).
//
// Two epilogue GROUPS of four warps work on alternate tiles (group g owns accumulator buffer g and the tiles with
// lt % 2 == g); a thread owns one accumulator row and all 64 columns of its tile. The erf-GELU arithmetic of one group
// (17.9 instructions per element, profiles/r01_ncu_full_summary.md) then runs while the other group of the CTA — and the
// two groups of the co-resident CTA — sit in their TMEM / staging / store waits, instead of being appended to the
// per-tile skeleton of a single eight-warp group (37.3 us = 23.2 us skeleton + 14.1 us math in v1).
// Per group ONE 16 KB staging tile serves both outputs in turn: gelu' is staged while it is computed and stored, the packed
// gelu values wait in 32 registers until that store has read the tile. Producer / MMA warps as in v1.
// =================================================================================================
constexpr int GRP_THREADS = 128;
template <int BMN>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_async_gelu2g_kernel(const __grid_constant__ AsyncArgs ga, int num_stages) {
  constexpr int kAccCols = 64, kTmemCols = 128;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* tmem_full_bar = empty_bar + 8;        // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  float* bias_s = reinterpret_cast<float*>(smem + 512);   // [2 groups][64]
  uint8_t* tiles = smem + 1024;
  uint8_t* out_s = tiles + (size_t)num_stages * kAStageBytes;  // [2 groups][16 KB]

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
    if (ga.has_out0) tma_prefetch_desc(&ga.tmOut0);
    for (int s = 0; s < num_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], GRP_THREADS);  // only the group that owns the buffer reads it
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
    // ------------------------------ TMA producer (as in v1) ------------------------------
    if (lane == 0) {
      int it = 0;
      for (int t = t_begin; t < t_end; ++t) {
        const int cb = t / ga.tiles_m;  // m fastest
        const int m0 = (t - cb * ga.tiles_m) * BM, n0 = cb * ABN;
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
    // ------------------------------ MMA issuer (as in v1) ------------------------------
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
    // ------------------------------ epilogue: two groups of four warps ------------------------------
    const int ew = warp - 2;             // 0..7
    const int grp = ew >> 2;             // 0: warps 2-5, 1: warps 6-9  (warp % 4 covers the four TMEM lane quarters in both)
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int gt = (ew & 3) * 32 + lane; // thread index inside the group, 0..127
    const int row = q * 32 + lane;       // accumulator row (TMEM lane) of this thread
    const bool issuer = (gt == 0);       // issues / tracks this group's bulk stores
    const uint32_t swz = (uint32_t)(row & 7);
    const uint32_t stage_row = smem_u32(out_s) + (uint32_t)(grp * kOutTileBytes) + (uint32_t)row * 128u;
    const uint8_t* stage_tile = out_s + grp * kOutTileBytes;
    float* gbias = bias_s + grp * ABN;
    // named barriers of this group: (1, 2) for group 0, (3, 4) for group 1 — immediates, so that the kernel reserves 5 ids
#define GBAR_A()                                                                        \
  do {                                                                                  \
    if (grp == 0) asm volatile("bar.sync 1, %0;" ::"n"(GRP_THREADS) : "memory");        \
    else asm volatile("bar.sync 3, %0;" ::"n"(GRP_THREADS) : "memory");                 \
  } while (0)
#define GBAR_B()                                                                        \
  do {                                                                                  \
    if (grp == 0) asm volatile("bar.sync 2, %0;" ::"n"(GRP_THREADS) : "memory");        \
    else asm volatile("bar.sync 4, %0;" ::"n"(GRP_THREADS) : "memory");                 \
  } while (0)
    int bias_cb = -1;
    // this group's tiles: local index lt = grp, grp + 2, ...  (accumulator buffer lt & 1 == grp)
    for (int lt = grp; t_begin + lt < t_end; lt += 2) {
      const int t = t_begin + lt;
      const int cb = t / ga.tiles_m;
      const int m0 = (t - cb * ga.tiles_m) * BM, n0 = cb * ABN;
      const uint32_t use = (uint32_t)(lt >> 1);  // how many times this group has used its buffer before
      if (cb != bias_cb) {  // group-uniform
        bias_cb = cb;
        GBAR_A();  // previous bias no longer read
        if (gt < ABN) gbias[gt] = (ga.bias != nullptr && n0 + gt < ga.N) ? __ldg(ga.bias + n0 + gt) : 0.f;
        GBAR_A();
      }
      mbar_wait(&tmem_full_bar[grp], use & 1u);
      tc_fence_after();
      // the staging tile must be free before gelu' is written into it: this group's previous store (gelu of its last tile)
      if (issuer) bulk_wait_read<0>();
      GBAR_A();
      uint32_t act[32];  // packed gelu of the 64 columns, staged after the gelu' store has read the tile
#pragma unroll
      for (int qc = 0; qc < 4; ++qc) {  // 16 accumulator columns at a time (keeps the live fp32 set small)
        float v[16];
        tmem_ld_32x16(tmem_base + (uint32_t)(grp * kAccCols + qc * 16) + ((uint32_t)(q * 32) << 16), v);
        tmem_ld_wait();
        const float4* b4 = reinterpret_cast<const float4*>(gbias + qc * 16);
#pragma unroll
        for (int j = 0; j < 2; ++j) {  // 8 columns = one 16-byte chunk of the output row
          const float4 ba = b4[2 * j], bb = b4[2 * j + 1];
          const float xs[8] = {v[8 * j] + ba.x, v[8 * j + 1] + ba.y, v[8 * j + 2] + ba.z, v[8 * j + 3] + ba.w,
                               v[8 * j + 4] + bb.x, v[8 * j + 5] + bb.y, v[8 * j + 6] + bb.z, v[8 * j + 7] + bb.w};
          uint32_t g4[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float c0, p0, c1, p1;
            gelu_parts(xs[2 * k], c0, p0);
            gelu_parts(xs[2 * k + 1], c1, p1);
            g4[k] = pack_bf16x2(fmaf(xs[2 * k], p0, c0), fmaf(xs[2 * k + 1], p1, c1));  // gelu'
            act[qc * 8 + j * 4 + k] = pack_bf16x2(xs[2 * k] * c0, xs[2 * k + 1] * c1);  // gelu
          }
          sts128(stage_row + ((((uint32_t)(qc * 2 + j)) ^ swz) << 4), g4[0], g4[1], g4[2], g4[3]);
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[grp]);  // the MMA warp may refill this accumulator (tile lt + 2)
      fence_proxy_async_smem();
      GBAR_B();
      if (issuer) {
        if (ga.has_out0) {
          tma_store_2d(&ga.tmOut0, stage_tile, n0, m0);
          bulk_commit();
          bulk_wait_read<0>();  // the other groups keep the SM busy meanwhile
        }
      }
      GBAR_A();
#pragma unroll
      for (int c = 0; c < 8; ++c)
        sts128(stage_row + ((((uint32_t)c) ^ swz) << 4), act[4 * c], act[4 * c + 1], act[4 * c + 2], act[4 * c + 3]);
      fence_proxy_async_smem();
      GBAR_B();
      if (issuer) {
        tma_store_2d(&ga.tmOut1, stage_tile, n0, m0);
        bulk_commit();
      }
    }
    if (issuer) bulk_wait<0>();  // all stores of this group are performed before the grid can complete
#undef GBAR_A
#undef GBAR_B
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// =================================================================================================
// Small-K kernel (SCOT_GEMM_ASYNC_V2 bit 2 (=4); NOT yet validated on hardware — round-2 candidate), K <= 128.
//
// The stage-0 GEMMs have K = 96: per 128 x 64 tile the v1 / v2 producer moves 48 KB through the TMA unit, of which 16 KB
// are the SAME 64 x 96 weight tile for every row tile of a column block and 12 KB are out-of-bounds zero fill of the
// second, half-empty 64-wide k-block. Here (a) the weight tile of a column block is loaded ONCE into a resident
// shared-memory region (b_full / b_empty barriers; a CTA walks down one column block, so it reloads at most twice) and
// (b) a K tail of 32 elements uses its own 32-wide boxes (64-byte swizzle atom, UMMA SWIZZLE_64B descriptors) instead of
// a zero-filled 64-wide block: 24 KB of operand traffic per tile instead of 48 KB. The A ring holds one k-block per
// 16 KB stage. Epilogue = async_epilogue_v2.
// =================================================================================================
struct AsyncArgsS {
  AsyncArgs a;
  CUtensorMap tmA_tail, tmB_tail;
  int tail_k;  // 0: every k-block is full; 32: the last k-block has 32 elements
};
constexpr int kSmallAStage = BM * BK * 2;        // 16 KB
constexpr int kSmallBBlock = ABN * BK * 2;       // 8 KB per resident B k-block
constexpr int kSmallBBytes = 2 * kSmallBBlock;   // up to two k-blocks (K <= 128)

// K-major operand tile with 64-byte rows (32 bf16 along K), 64-byte swizzle: 8-row groups 512 B apart
__device__ __forceinline__ uint64_t umma_smem_desc_sw64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((16u >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((512u >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)4 << 61;  // SWIZZLE_64B
  return d;
}

template <int BMN, int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_async_smallk_kernel(const __grid_constant__ AsyncArgsS gs, int num_stages) {
  constexpr bool kHasAux = (MODE == SCOT_EPI_GELU_BWD);
  constexpr int kAccCols = 64, kTmemCols = 128;
  const AsyncArgs& ga = gs.a;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* tmem_full_bar = empty_bar + 8;        // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint64_t* aux_full_bar = tmem_empty_bar + 2;    // [2]
  uint64_t* aux_empty_bar = aux_full_bar + 2;     // [2]
  uint64_t* b_full_bar = aux_empty_bar + 2;       // [1]
  uint64_t* b_empty_bar = b_full_bar + 1;         // [1]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(b_empty_bar + 1);
  float* bias_s = reinterpret_cast<float*>(smem + 512);
  uint8_t* b_s = smem + 1024;                                              // resident weight tile, [2][8 KB]
  uint8_t* a_ring = b_s + kSmallBBytes;                                    // [num_stages][16 KB]
  uint8_t* aux_s = a_ring + (size_t)num_stages * kSmallAStage;             // [2][16 KB] (GELU_BWD only; doubles as staging)
  uint8_t* out_s = aux_s + (kHasAux ? 2 * kOutTileBytes : 0);

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = ga.total_tiles;
  const int tiles_per_cta = (total_tiles + gridDim.x - 1) / gridDim.x;
  const int t_begin = blockIdx.x * tiles_per_cta;
  const int t_end = min(total_tiles, t_begin + tiles_per_cta);
  const int kb_n = ga.kblocks;           // 1 or 2
  const int tail = gs.tail_k;            // 0 or 32

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&ga.tmA);
    tma_prefetch_desc(&ga.tmB);
    tma_prefetch_desc(&ga.tmOut0);
    if (tail) {
      tma_prefetch_desc(&gs.tmA_tail);
      tma_prefetch_desc(&gs.tmB_tail);
    }
    if (kHasAux) tma_prefetch_desc(&ga.tmAux);
    if (MODE == SCOT_EPI_GELU) tma_prefetch_desc(&ga.tmOut1);
    for (int s = 0; s < num_stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], EPI_THREADS);
      mbar_init(&aux_full_bar[b], 1);
      mbar_init(&aux_empty_bar[b], 1);
    }
    mbar_init(b_full_bar, 1);
    mbar_init(b_empty_bar, 1);
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
      int it = 0, lt = 0, cur_cb = -1;
      uint32_t b_use = 0;
      const uint32_t b_bytes = (uint32_t)((kb_n - (tail ? 1 : 0)) * kSmallBBlock + (tail ? ABN * tail * 2 : 0));
      for (int t = t_begin; t < t_end; ++t, ++lt) {
        const int cb = t / ga.tiles_m;  // m fastest
        const int m0 = (t - cb * ga.tiles_m) * BM, n0 = cb * ABN;
        if (cb != cur_cb) {
          // every MMA that read the previous weight tile has completed (committed by the MMA warp after the last tile of
          // the previous column block); the first use passes immediately
          mbar_wait_backoff(b_empty_bar, (b_use & 1u) ^ 1u);
          mbar_expect_tx(b_full_bar, b_bytes);
          for (int kb = 0; kb < kb_n; ++kb) {
            const bool is_tail = tail && kb == kb_n - 1;
            const CUtensorMap* tm = is_tail ? &gs.tmB_tail : &ga.tmB;
            if constexpr (BMN == 0) tma_load_2d(b_s + kb * kSmallBBlock, tm, b_full_bar, kb * BK, n0);
            else tma_load_2d(b_s + kb * kSmallBBlock, tm, b_full_bar, n0, kb * BK);
          }
          cur_cb = cb;
          ++b_use;
        }
        if constexpr (kHasAux) {
          const int as = lt & 1;
          mbar_wait_backoff(&aux_empty_bar[as], (((uint32_t)lt >> 1) & 1u) ^ 1u);
          mbar_expect_tx(&aux_full_bar[as], kOutTileBytes);
          tma_load_2d(aux_s + as * kOutTileBytes, &ga.tmAux, &aux_full_bar[as], n0, m0);
        }
        for (int kb = 0; kb < kb_n; ++kb, ++it) {
          const int s = it % num_stages;
          const uint32_t ph = (uint32_t)(it / num_stages) & 1u;
          const bool is_tail = tail && kb == kb_n - 1;
          mbar_wait_backoff(&empty_bar[s], ph ^ 1u);
          mbar_expect_tx(&full_bar[s], is_tail ? (uint32_t)(BM * tail * 2) : (uint32_t)kSmallAStage);
          tma_load_2d(a_ring + (size_t)s * kSmallAStage, is_tail ? &gs.tmA_tail : &ga.tmA, &full_bar[s], kb * BK, m0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, ABN, 0, BMN);
      int it = 0, lt = 0, cur_cb = -1;
      uint32_t b_use = 0;
      for (int t = t_begin; t < t_end; ++t, ++lt) {
        const int cb = t / ga.tiles_m;
        if (cb != cur_cb) {
          mbar_wait_backoff(b_full_bar, b_use & 1u);  // the weight tile of this column block has landed
          cur_cb = cb;
          ++b_use;
        }
        const int buf = lt & 1;
        mbar_wait_backoff(&tmem_empty_bar[buf], (((uint32_t)lt >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(buf * kAccCols);
        for (int i = 0; i < kb_n; ++i, ++it) {
          const int s = it % num_stages;
          const uint32_t ph = (uint32_t)(it / num_stages) & 1u;
          const bool is_tail = tail && i == kb_n - 1;
          mbar_wait_backoff(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(a_ring + (size_t)s * kSmallAStage);
          const uint32_t sb = smem_u32(b_s + i * kSmallBBlock);
          const int nk = is_tail ? tail / UMMA_K : BK / UMMA_K;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            if (k < nk) {
              const uint64_t da = is_tail ? umma_smem_desc_sw64(sa + k * 32) : umma_smem_desc(sa + k * 32, 16, 1024);
              const uint64_t db = (BMN == 0) ? (is_tail ? umma_smem_desc_sw64(sb + k * 32) : umma_smem_desc(sb + k * 32, 16, 1024))
                                             : umma_smem_desc(sb + k * 2048, BK * 128, 1024);
              umma_bf16(tacc, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full_bar[buf]);
        const bool last_of_cb = (t + 1 >= t_end) || ((t + 1) / ga.tiles_m != cb);
        if (last_of_cb) umma_commit(b_empty_bar);  // arrives once every MMA issued so far has read its operands
      }
    }
  } else {
    async_epilogue_v2<MODE>(ga, warp, lane, t_begin, t_end, tmem_base, tmem_full_bar, tmem_empty_bar, aux_full_bar,
                            aux_empty_bar, bias_s, aux_s, out_s);
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
                                 int bmn, int M, int N, int K, EpiArgs ep) {
  const int col = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const long row = (long)blockIdx.y * blockDim.y + threadIdx.y;
  if (row >= M || col >= N) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k = 0; k < K; ++k) {
    const float a = __bfloat162float(amn ? A[(long)k * lda + row] : A[row * lda + k]);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float b = __bfloat162float(bmn ? B[(long)k * ldb + col + j] : B[(long)(col + j) * ldb + k]);
      acc[j] = fmaf(a, b, acc[j]);
    }
  }
  float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 bias = ep.bias != nullptr ? *reinterpret_cast<const float4*>(ep.bias + col) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 vv = make_float4(acc[0], acc[1], acc[2], acc[3]);
  if constexpr (MODE == SCOT_EPI_GELU)
    vv = make_float4(gelu_pack(vv.x + bias.x), gelu_pack(vv.y + bias.y), gelu_pack(vv.z + bias.z), gelu_pack(vv.w + bias.w));
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

// fp32 row-major tensor [outer, inner], box {32, 128} (one 128-byte swizzle atom wide)
int make_tmap_f32(CUtensorMap* tm, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld) {
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)BM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SCOT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(fp32) failed (%d): ptr=%p inner=%llu outer=%llu ld=%llu", (int)r, ptr,
               (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld);
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
    P.M = h.M;
    P.N = h.N;
    P.tiles_m = ceil_div(h.M, BM);
    P.tiles_n = ceil_div(h.N, BN);
    P.kblocks_total = ceil_div(h.K, BK);
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

// SCOT_GEMM_ASYNC_V2=1 selects gemm_async_epi2_kernel (round-2 candidate, see its header)
// SCOT_GEMM_ASYNC_V2 is a bit mask: 1 = v2 kernels, 2 = + two-group GELU kernel (gemm_async_gelu2g_kernel), 4 = + small-K
// kernel with the resident weight tile (gemm_async_smallk_kernel, K <= 128). 0 / unset = the validated v1 kernels.
int async_v2_level() {
  const char* e = getenv("SCOT_GEMM_ASYNC_V2");
  const int v = e != nullptr ? atoi(e) : 0;
  return (v >= 1 && v <= 7) ? (v | 1) : 0;
}
bool async_v2_enabled() { return async_v2_level() != 0; }

template <int BMN, int MODE, bool V2 = false>
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
      rc = make_tmap(&ga.tmOut0, ep.out0, (uint64_t)N, (uint64_t)M, (uint64_t)ep.ld0, 64, BM);
      if (rc) return rc;
    }
    rc = make_tmap(&ga.tmOut1, ep.out1, (uint64_t)N, (uint64_t)M, (uint64_t)ep.ld1, 64, BM);
    if (rc) return rc;
  } else if (MODE == SCOT_EPI_F32 || MODE == SCOT_EPI_RMW_F32) {
    rc = make_tmap_f32(&ga.tmOut0, ep.out0, (uint64_t)N, (uint64_t)M, (uint64_t)ep.ld0);
    if (rc) return rc;
    ga.tmOut1 = ga.tmOut0;
  } else {
    rc = make_tmap(&ga.tmOut0, ep.out0, (uint64_t)N, (uint64_t)M, (uint64_t)ep.ld0, 64, BM);
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
  if constexpr (V2) {
    // small K: resident weight tile + exact K tail (see gemm_async_smallk_kernel)
    const int tail_k = K % BK;
    if ((async_v2_level() & 4) && K <= 2 * BK && (tail_k == 0 || tail_k == 32) &&
        !(MODE == SCOT_EPI_GELU && (async_v2_level() & 2))) {
      AsyncArgsS gs;
      memset(&gs, 0, sizeof(gs));
      gs.a = ga;
      gs.tail_k = tail_k;
      if (tail_k) {
        const int k0 = K - tail_k;  // the tail boxes are addressed with the same (k, row) coordinates as the full ones
        (void)k0;
        rc = make_tmap(&gs.tmA_tail, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 32, BM, CU_TENSOR_MAP_SWIZZLE_64B);
        if (rc) return rc;
        if (BMN == 0) rc = make_tmap(&gs.tmB_tail, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 32, ABN, CU_TENSOR_MAP_SWIZZLE_64B);
        else rc = make_tmap(&gs.tmB_tail, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, 32);
        if (rc) return rc;
      }
      const size_t staging_s = (MODE == SCOT_EPI_GELU_BWD ? 0 : 2) * (size_t)kOutTileBytes;
      const size_t fixed_s = 1024 + 1024 + kSmallBBytes + (MODE == SCOT_EPI_GELU_BWD ? 2 * kOutTileBytes : 0) + staging_s;
      const size_t budget_s = (size_t)(227 * 1024) / 2 - 1024;
      int stages_s = (int)((budget_s - fixed_s) / kSmallAStage);
      if (stages_s > 6) stages_s = 6;
      SCOT_REQUIRE(stages_s >= 2, "gemm(small K): shared memory budget");
      const size_t smem_s = fixed_s + (size_t)stages_s * kSmallAStage;
      auto kern_s = gemm_async_smallk_kernel<BMN, MODE>;
      static bool attr_s_done = false;
      if (!attr_s_done) {
        SCOT_CHECK_CUDA(cudaFuncSetAttribute(kern_s, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 / 2));
        attr_s_done = true;
      }
      const int max_ctas_s = g_num_sms * 2;
      const int grid_s = ga.total_tiles < max_ctas_s ? ga.total_tiles : max_ctas_s;
      SCOT_CHECK_CUDA(scot_launch_pdl(kern_s, dim3(grid_s), dim3(GEMM_THREADS), smem_s, stream, gs, stages_s));
      SCOT_LAUNCH_CHECK();
      return 0;
    }
  }
  // staging tiles: v1 one per output (+ the aux ring for GELU_BWD); v2: two for BF16, the aux ring alone for GELU_BWD
  const size_t staging = V2 ? (MODE == SCOT_EPI_GELU_BWD ? 0 : 2) * (size_t)kOutTileBytes  // GELU: two outputs; BF16: two
                            : (MODE == SCOT_EPI_GELU ? 2 : 1) * (size_t)kOutTileBytes;          // buffers; fp32: two halves
  const size_t fixed = 1024 /*align slack*/ + 1024 /*barriers*/ + (MODE == SCOT_EPI_GELU_BWD ? 2 * kOutTileBytes : 0) + staging;
  const size_t budget = (size_t)(227 * 1024) / 2 - 1024;
  int stages = (int)((budget - fixed) / kAStageBytes);
  if (stages > 4) stages = 4;
  SCOT_REQUIRE(stages >= 2, "gemm(async epilogue): shared memory budget");
  const size_t smem = fixed + (size_t)stages * kAStageBytes;
  void (*kern)(AsyncArgs, int);
  if constexpr (V2) kern = gemm_async_epi2_kernel<BMN, MODE>;
  else kern = gemm_async_epi_kernel<BMN, MODE>;
  if constexpr (V2 && MODE == SCOT_EPI_GELU) {
    // two-group variant: same shared-memory footprint (one 16 KB staging tile per group instead of one per output)
    if (async_v2_level() & 2) {
      kern = gemm_async_gelu2g_kernel<BMN>;
      static bool attr2_done = false;
      if (!attr2_done) {
        SCOT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 / 2));
        attr2_done = true;
      }
    }
  }
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
  // bf16-output modes: async-epilogue kernel (TMA stores, TMA-fed auxiliary operand), 128 x 64 tiles, two CTAs per SM
  if constexpr (AMN == 0 && (MODE == SCOT_EPI_GELU || MODE == SCOT_EPI_GELU_BWD || MODE == SCOT_EPI_BF16)) {
    if (async_epi_enabled() && tma_ok(ep.out0, ep.ld0) && tma_ok(ep.out1, ep.ld1) && tma_ok(ep.aux, ep.ldaux) &&
        (MODE != SCOT_EPI_GELU || ep.out1 != nullptr) && (MODE != SCOT_EPI_GELU_BWD || ep.aux != nullptr) &&
        (BMN == 0 || N % 8 == 0)) {
      if (async_v2_enabled()) return launch_async<BMN, MODE, true>(A, lda, B, ldb, M, N, K, ep, stream);
      return launch_async<BMN, MODE>(A, lda, B, ldb, M, N, K, ep, stream);
    }
  }
  // v2 only: fp32-output modes through the async epilogue (TMA store / TMA reduce-add). `+=` only where gemm_tc_kernel
  // would not split K (large token counts): the async kernel has no split reduction.
  if constexpr (AMN == 0 && (MODE == SCOT_EPI_F32 || MODE == SCOT_EPI_RMW_F32)) {
    if (async_v2_enabled() && (((uintptr_t)ep.out0) & 15) == 0 && ep.ld0 % 4 == 0 && (BMN == 0 || N % 8 == 0) &&
        (MODE == SCOT_EPI_F32 || ceil_div(M, BM) * ceil_div(N, 128) >= 2 * g_num_sms))
      return launch_async<BMN, MODE, true>(A, lda, B, ldb, M, N, K, ep, stream);
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
                                                     reinterpret_cast<const bf16*>(B), ldb, bmn, M, N, K, ep);
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
                        EpiArgs{nullptr, q.dW, q.ld_dw, nullptr, 0, nullptr, 0, nullptr}};
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
  EpiArgs ep{e->bias, e->out0, e->ld0, e->out1, e->ld1, e->aux, e->ldaux, e->colsum};
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

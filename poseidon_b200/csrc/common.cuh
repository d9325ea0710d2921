// Shared device helpers for the scOT sm_100a engine: PTX wrappers (mbarrier, TMA, tcgen05),
// small math helpers and the error plumbing used by every translation unit.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------------------------------------
// error plumbing (host)
// ------------------------------------------------------------------------------------------------
void scot_set_error(const char* fmt, ...);
void scot_count_launch();
// "parity" precision (split-bf16): every bf16 tensor T of the hot path has a twin holding bf16(x - float(bf16(x))) that
// lives `scot_split_off()` bytes after T (0 = plain bf16 mode). Thread-local; the engine sets it for the duration of one
// forward / backward call, the per-op C ABI through scot_set_split_offset().
size_t scot_split_off();
void scot_set_split_off(size_t bytes);
#define SCOT_CHECK_CUDA(expr)                                                                  \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      scot_set_error("%s:%d CUDA error %d (%s) in %s", __FILE__, __LINE__, (int)_e,            \
                     cudaGetErrorString(_e), #expr);                                           \
      return 2;                                                                                \
    }                                                                                          \
  } while (0)
#define SCOT_REQUIRE(cond, ...)                                                                \
  do {                                                                                         \
    if (!(cond)) {                                                                             \
      scot_set_error(__VA_ARGS__);                                                             \
      return 1;                                                                                \
    }                                                                                          \
  } while (0)
#define SCOT_LAUNCH_CHECK()                                                                    \
  do {                                                                                         \
    scot_count_launch();                                                                       \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess) {                                                                   \
      scot_set_error("%s:%d kernel launch failed: %s", __FILE__, __LINE__,                     \
                     cudaGetErrorString(_e));                                                  \
      return 2;                                                                                \
    }                                                                                          \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------------------------------------
// Programmatic dependent launch: the hot kernels are launched with programmaticStreamSerialization so
// that the launch + block scheduling (+ for the GEMM its barrier/TMEM setup) of kernel N+1 overlaps the
// tail of kernel N. Each such kernel calls pdl_launch_dependents() first and pdl_wait() before it touches
// global memory; ~1400 mostly 10-20 us kernels per training step make the launch gaps worth hiding.
// SCOT_PDL=0 in the environment disables the attribute (plain stream order).
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
bool scot_pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t scot_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                          Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = scot_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------
// generic device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// erf-GELU (nn.GELU default, reference ACT2FN["gelu"]) with the Abramowitz-Stegun 7.1.26 rational/exponential form of
// erf (|abs err| <= 1.5e-7, far below the bf16 the result is stored in): one ex2, one rcp and a handful of FMAs
// instead of erff's ~40-instruction path — the GEMM epilogues run on only four warps per CTA.
__device__ __forceinline__ void gelu_parts(float x, float& cdf, float& pdf) {
  const float ax = fabsf(x) * 0.70710678118654752440f;           // |x|/sqrt(2)
  float e2;  // exp(-x^2/2) = 2^(-x^2 * log2(e)/2): one MUFU.EX2 (flush-to-zero: no denormal range fix-up code)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"((x * x) * -0.72134752044448170368f));
  float t;  // 1/(1 + p|x|/sqrt2): MUFU.RCP (1 ulp) is plenty for a result that is rounded to bf16
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.0f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erfc_ax = poly * t * e2;                            // 1 - erf(|x|/sqrt 2)
  const float half_erfc = 0.5f * erfc_ax;
  cdf = x >= 0.f ? 1.0f - half_erfc : half_erfc;                  // Phi(x)
  pdf = 0.39894228040143267794f * e2;                             // phi(x)
}
// gelu(x) and gelu'(x) together, same erf form as gelu_parts with the constant factors folded (|x|/sqrt2 into the rcp
// argument, the 1/2 of erfc into the polynomial, phi's 1/sqrt(2 pi) into x): 6 FMUL + 6 FFMA + 2 MUFU + 3 select per
// element — the GELU GEMM epilogue is issue-bound on exactly this sequence.
__device__ __forceinline__ void gelu_and_grad(float x, float& g, float& dg) {
  float e2;  // exp(-x^2/2)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"((x * x) * -0.72134752044448170368f));
  float t;   // 1/(1 + p|x|/sqrt2)
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f * 0.70710678118654752440f, fabsf(x), 1.0f)));
  float poly = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  poly = fmaf(poly, t, 0.5f * 1.421413741f);
  poly = fmaf(poly, t, 0.5f * -0.284496736f);
  poly = fmaf(poly, t, 0.5f * 0.254829592f);
  const float half_erfc = (poly * t) * e2;                        // (1 - erf(|x|/sqrt 2)) / 2
  const float cdf = x >= 0.f ? 1.0f - half_erfc : half_erfc;      // Phi(x)
  g = x * cdf;
  dg = fmaf(x * 0.39894228040143267794f, e2, cdf);               // Phi(x) + x phi(x)
}
// full-precision variant for the split-bf16 ("parity") mode: libm erff / expf (a few ulp)
__device__ __forceinline__ void gelu_parts_precise(float x, float& cdf, float& pdf) {
  cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
}
__device__ __forceinline__ float gelu_erf(float x) {
  float cdf, pdf;
  gelu_parts(x, cdf, pdf);
  return x * cdf;
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float cdf, pdf;
  gelu_parts(x, cdf, pdf);
  return fmaf(x, pdf, cdf);
}
// 2^x on the SFU (MUFU.EX2, ~2 ulp, denormal results flush to zero): one instruction instead of exp2f()'s
// range-checked four. Inputs here are softmax exponents <= 0.
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(t);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// ---- split-bf16 ("parity" precision) loads / stores: value = hi + lo, lo tensor `lo_off` bytes after hi (0 = none) ----
__device__ __forceinline__ uint32_t split_lo_pack(float a, float b, uint32_t hi) {
  const float2 h = unpack_bf16x2(hi);
  return pack_bf16x2(a - h.x, b - h.y);
}
// stores 4 consecutive bf16 (8-byte aligned) and returns the value as stored (hi, or hi + lo)
__device__ __forceinline__ float4 st_bf16x4(bf16* p, size_t lo_off, float a, float b, float c, float d) {
  const uint2 h = make_uint2(pack_bf16x2(a, b), pack_bf16x2(c, d));
  *reinterpret_cast<uint2*>(p) = h;
  const float2 h01 = unpack_bf16x2(h.x), h23 = unpack_bf16x2(h.y);
  if (lo_off == 0) return make_float4(h01.x, h01.y, h23.x, h23.y);
  const uint2 l = make_uint2(pack_bf16x2(a - h01.x, b - h01.y), pack_bf16x2(c - h23.x, d - h23.y));
  *reinterpret_cast<uint2*>(reinterpret_cast<char*>(p) + lo_off) = l;
  const float2 l01 = unpack_bf16x2(l.x), l23 = unpack_bf16x2(l.y);
  return make_float4(h01.x + l01.x, h01.y + l01.y, h23.x + l23.x, h23.y + l23.y);
}
__device__ __forceinline__ float4 ld_bf16x4(const bf16* p, size_t lo_off) {
  const uint2 h = *reinterpret_cast<const uint2*>(p);
  const float2 h01 = unpack_bf16x2(h.x), h23 = unpack_bf16x2(h.y);
  if (lo_off == 0) return make_float4(h01.x, h01.y, h23.x, h23.y);
  const uint2 l = *reinterpret_cast<const uint2*>(reinterpret_cast<const char*>(p) + lo_off);
  const float2 l01 = unpack_bf16x2(l.x), l23 = unpack_bf16x2(l.y);
  return make_float4(h01.x + l01.x, h01.y + l01.y, h23.x + l23.x, h23.y + l23.y);
}
__device__ __forceinline__ void st_bf16(bf16* p, size_t lo_off, float v) {
  const bf16 h = __float2bfloat16_rn(v);
  *p = h;
  if (lo_off) *reinterpret_cast<bf16*>(reinterpret_cast<char*>(p) + lo_off) = __float2bfloat16_rn(v - __bfloat162float(h));
}
__device__ __forceinline__ float ld_bf16(const bf16* p, size_t lo_off) {
  float v = __bfloat162float(*p);
  if (lo_off) v += __bfloat162float(*reinterpret_cast<const bf16*>(reinterpret_cast<const char*>(p) + lo_off));
  return v;
}

// ------------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// for the single-lane producer / MMA-issuer loops whose waits are long (they run ahead of the epilogue): back off so
// that the polling does not take issue slots from the epilogue warps that share the scheduler
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(64);
}
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
  return v;
}

// ------------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) — 2D tile load global -> shared, completion on an mbarrier
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// 2D tile store shared -> global (bulk async group; rows / columns beyond the tensor bounds are clipped)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
// 2D tile reduction shared -> global: global[tile] += smem[tile] (fp32 add performed by the L2)
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the newest N bulk groups of this thread have finished READING their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// ... have completed entirely (global writes performed)
template <int N>
__device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// make generic-proxy shared-memory writes visible to the async proxy (TMA) before a bulk store reads them
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// tcgen05 (5th-gen tensor cores, TMEM)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate. One thread issues.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 columns of 32-bit: thread i of the warp receives TMEM lane (base_lane + i), columns [c, c+32)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// same, 16 columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (sm_100 "version 1"), 128-byte swizzle.
//   K-major tile : rows of 128 B (64 bf16 along K); 8-row groups 1024 B apart (SBO); LBO unused.
//   MN-major tile: rows of 128 B (64 bf16 along M/N), one row per k; 8-k groups 1024 B apart (SBO);
//                  64-wide MN atoms `lbo_bytes` apart (LBO).
// Field layout follows cute::UMMA::SmemDescriptor (cute/arch/mma_sm100_desc.hpp in the CUTLASS tree).
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::f16 with bf16 A/B and fp32 D (cute::UMMA::InstrDescriptor bit layout).
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
// legacy warp-level tensor-core helpers (mma.sync m16n8k16 bf16) used by the window-attention kernels
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_bf16_16816(float* d, const uint32_t* a, const uint32_t* b) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void ldsm_x4(uint32_t* r, uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t* r, uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t* r, uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(saddr));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t* r, uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];"
               : "=r"(r[0]), "=r"(r[1])
               : "r"(saddr));
}
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
#endif  // __CUDACC__

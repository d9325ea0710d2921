// Fused optimizer step on the flat parameter / gradient buffers.
//
// The reference trains with HF Trainer: accelerate clip_grad_norm_(max_grad_norm) followed by torch.optim.AdamW
// ("adamw_torch", scOT/train.py:286) over the 2-4 parameter groups built in scOT/trainer.py:295-400 (weight-decay
// exclusion for norms/biases, separate learning rates for embeddings/recovery and for the time-conditioned norms) —
// a few foreach kernels over ~1600 tensors. Here parameters and gradients already live in ONE flat fp32 buffer each, so
// the step is two HBM-bound launches: the squared gradient norm, and AdamW (decoupled weight decay, bias correction,
// the clip coefficient applied on the fly) with 128-bit loads/stores, optionally refreshing the bf16 GEMM copy of the
// weights in the same pass. Group hyper-parameters come from a small device table, so the launches are CUDA-graph
// capturable while learning-rate schedules keep changing the values between replays.
//
// Layout contract: the engine aligns every parameter to 64 elements (engine.cu Registry::alloc), so each 64-element
// chunk of the flat buffer belongs to one parameter (or padding): `chunk_group[i / 64]` = group id, 255 = skip.
#include "common.cuh"
#include "internal.h"
#include "scot_b200.h"

namespace {

// Deterministic: per-block partial sums land in a fixed scratch slot and the LAST block to finish adds them in block order,
// so every data-parallel rank computes bit-identical norms (an atomicAdd of the partials rounds differently from run to run
// and from rank to rank: the clip coefficient, and with it the replicas, would drift apart in the last bit).
constexpr int kNormMaxBlocks = 2048;
__device__ float g_norm_partials[kNormMaxBlocks];
__device__ unsigned int g_norm_count = 0;

__global__ void __launch_bounds__(256) grad_sq_norm_kernel(const float* __restrict__ g, long n4, float* __restrict__ out) {
  float acc = 0.f;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    acc += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  acc = warp_sum(acc);
  __shared__ float s[8];
  __shared__ bool last;
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int k = 0; k < 8; ++k) v += s[k];
    g_norm_partials[blockIdx.x] = v;
    __threadfence();
    last = (atomicAdd(&g_norm_count, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    float total = 0.f;
    for (unsigned k = 0; k < gridDim.x; ++k) total += reinterpret_cast<volatile float*>(g_norm_partials)[k];
    out[0] = total;
    g_norm_count = 0;  // ready for the next launch (launches of this kernel are stream-ordered)
  }
}

// hp[group][8] = {lr, weight_decay, beta1, beta2, eps, bias_correction1, sqrt(bias_correction2), unused}
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             bf16* __restrict__ p16, const uint8_t* __restrict__ chunk_group, long n4, const float* __restrict__ hp,
             const float* __restrict__ sq_norm, float max_norm, float grad_scale) {
  float clip = grad_scale;
  if (sq_norm != nullptr && max_norm > 0.f) {
    // torch.nn.utils.clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to 1
    const float total = sqrtf(*sq_norm) * grad_scale;
    clip *= fminf(max_norm / (total + 1e-6f), 1.0f);
  }
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    const int grp = chunk_group[i >> 4];  // 16 float4 per 64-element chunk
    if (grp == 255) continue;
    const float* h = hp + grp * 8;
    const float lr = h[0], wd = h[1], b1 = h[2], b2 = h[3], eps = h[4], bc1 = h[5], sbc2 = h[6];
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    const float decay = 1.0f - lr * wd;
    const float step = lr / bc1;
#define ADAMW_ONE(X)                                   \
  {                                                    \
    const float gr = gg.X * clip;                      \
    pp.X *= decay;                                     \
    mm.X = mm.X + (1.0f - b1) * (gr - mm.X);           \
    vv.X = b2 * vv.X + (1.0f - b2) * gr * gr;          \
    const float denom = sqrtf(vv.X) / sbc2 + eps;      \
    pp.X -= step * (mm.X / denom);                     \
  }
    ADAMW_ONE(x) ADAMW_ONE(y) ADAMW_ONE(z) ADAMW_ONE(w)
#undef ADAMW_ONE
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (p16 != nullptr) reinterpret_cast<uint2*>(p16)[i] = make_uint2(pack_bf16x2(pp.x, pp.y), pack_bf16x2(pp.z, pp.w));
  }
}

int grid_for(long n4) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long blocks = (n4 + 255) / 256;
  const long cap = (long)sms * 8;  // 8 resident 256-thread blocks per SM: one wave, grid-stride inside
  return (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

}  // namespace

extern "C" {

int scot_grad_sq_norm(const float* grads, long n_elems, float* out, void* stream) {
  SCOT_REQUIRE(grads && out && n_elems > 0 && n_elems % 4 == 0, "grad_sq_norm: bad arguments (n must be a multiple of 4)");
  cudaStream_t st = (cudaStream_t)stream;
  SCOT_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float), st));
  int grid = grid_for(n_elems / 4);
  if (grid > kNormMaxBlocks) grid = kNormMaxBlocks;
  grad_sq_norm_kernel<<<grid, 256, 0, st>>>(grads, n_elems / 4, out);
  SCOT_LAUNCH_CHECK();
  return 0;
}

int scot_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, void* params_bf16,
                    const uint8_t* chunk_group, long n_elems, const float* group_hp, int n_groups, const float* grad_sq_norm,
                    float max_norm, float grad_scale, void* stream) {
  SCOT_REQUIRE(params && grads && exp_avg && exp_avg_sq && chunk_group && group_hp, "adamw_step: null pointer");
  SCOT_REQUIRE(n_elems > 0 && n_elems % 64 == 0, "adamw_step: the flat buffer length must be a multiple of 64 (got %ld)", n_elems);
  SCOT_REQUIRE(n_groups >= 1 && n_groups <= 254, "adamw_step: 1..254 parameter groups");
  adamw_kernel<<<grid_for(n_elems / 4), 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, (bf16*)params_bf16,
                                                                       chunk_group, n_elems / 4, group_hp, grad_sq_norm,
                                                                       max_norm, grad_scale);
  SCOT_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// Evaluation metrics on the device: per-(sample, channel) sums of |pred - y|^p and |y|^p —
// the two reductions behind scOT/metrics.py:12-36 `relative_lp_error` — so that an evaluation pass moves
// 2 * B * C floats to the host instead of the full [N, C, 128, 128] predictions (scOT/train.py:344-398).
// One block per (sample, channel) plane, fixed summation order: deterministic.
// ---------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256)
lp_plane_sums_kernel(const float* __restrict__ pred, const float* __restrict__ labels, float* __restrict__ out, int p, long HW) {
  const long plane = blockIdx.x;
  const float* a = pred + plane * HW;
  const float* b = labels + plane * HW;
  float num = 0.f, den = 0.f;
  for (long k = threadIdx.x; k < HW; k += blockDim.x) {
    const float y = b[k], d = fabsf(a[k] - y), ay = fabsf(y);
    if (p == 1) {
      num += d;
      den += ay;
    } else if (p == 2) {
      num += d * d;
      den += ay * ay;
    } else {
      num += powf(d, (float)p);
      den += powf(ay, (float)p);
    }
  }
  num = warp_sum(num);
  den = warp_sum(den);
  __shared__ float s1[8], s2[8];
  if ((threadIdx.x & 31) == 0) {
    s1[threadIdx.x >> 5] = num;
    s2[threadIdx.x >> 5] = den;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float x = 0.f, y = 0.f;
    for (int k = 0; k < 8; ++k) {
      x += s1[k];
      y += s2[k];
    }
    out[plane * 2 + 0] = x;
    out[plane * 2 + 1] = y;
  }
}
}  // namespace

extern "C" int scot_lp_plane_sums(const float* pred, const float* labels, float* out, int p, long planes, long hw, void* stream) {
  SCOT_REQUIRE(pred && labels && out && planes > 0 && hw > 0 && p >= 1, "lp_plane_sums: bad arguments");
  lp_plane_sums_kernel<<<(unsigned)planes, 256, 0, (cudaStream_t)stream>>>(pred, labels, out, p, hw);
  SCOT_LAUNCH_CHECK();
  return 0;
}

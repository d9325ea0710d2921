// HBM-bound helper kernels around the GEMM / attention core: im2col for the 4x4 patch embedding, patch
// merging gather / scatter, depthwise 7x7 stencil (ConvNeXt), the 5x5 output mixing conv of
// ScOTPatchRecovery, layer-scale residual, casts and the (relative) Lp loss.
// Reference call sites: scOT/model.py:295-310 (embed), :694-704 (merge order (0,0),(1,0),(0,1),(1,1)),
// :198-217 (ConvNeXt), :639-647 (recovery), :1422-1484 (pixel_mask overwrite + loss).
#include "common.cuh"
#include "internal.h"

namespace {

constexpr int kThreads = 256;
inline unsigned blocks_for(long n, int per = kThreads) { return (unsigned)((n + per - 1) / per); }

// ---- casts / elementwise -----------------------------------------------------------------------------
__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, long n4) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = reinterpret_cast<const float4*>(in)[i];
  reinterpret_cast<uint2*>(out)[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}

__global__ void expand_bias_kernel(const float* __restrict__ bias, float* __restrict__ out, int n, int rep) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = bias[i / rep];
}

// ---- patch embedding im2col: [B,Cin,H,W] fp32 -> [B*(H/ps)*(W/ps), Cin*ps*ps] bf16, k = (c, di, dj) ------
__global__ void im2col_patch_kernel(const float* __restrict__ x, bf16* __restrict__ out, int B, int Cin, int H, int W,
                                    int ps) {
  const int gw = W / ps, gh = H / ps;
  const long total = (long)B * gh * gw * Cin * ps;  // one thread per (token, c, di): ps contiguous pixels
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int di = (int)(i % ps);
  long t = i / ps;
  const int c = (int)(t % Cin);
  t /= Cin;
  const int j = (int)(t % gw);
  t /= gw;
  const int ii = (int)(t % gh);
  const int b = (int)(t / gh);
  const float* src = x + (((long)b * Cin + c) * H + (ii * ps + di)) * W + j * ps;
  bf16* dst = out + (((long)b * gh + ii) * gw + j) * (Cin * ps * ps) + (c * ps + di) * ps;
  for (int dj = 0; dj < ps; ++dj) dst[dj] = __float2bfloat16_rn(src[dj]);
}

// ---- patch merging ------------------------------------------------------------------------------------
// out[b, (i,j), q*C + c] = x[b, 2i+(q&1), 2j+(q>>1), c] + inp[...]   q = 0..3 -> (0,0),(1,0),(0,1),(1,1)
__global__ void merge_gather_kernel(const float* __restrict__ x, const float* __restrict__ inp, bf16* __restrict__ out,
                                    int B, int res, int C) {
  const int c4n = C / 4, r2 = res / 2;
  const long total = (long)B * r2 * r2 * 4 * c4n;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c4 = (int)(i % c4n);
  long t = i / c4n;
  const int q = (int)(t & 3);
  t >>= 2;
  const int j = (int)(t % r2);
  t /= r2;
  const int ii = (int)(t % r2);
  const int b = (int)(t / r2);
  const long src = (((long)b * res + 2 * ii + (q & 1)) * res + 2 * j + (q >> 1)) * C + c4 * 4;
  float4 v = *reinterpret_cast<const float4*>(x + src);
  if (inp != nullptr) {
    const float4 w = *reinterpret_cast<const float4*>(inp + src);
    v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
  }
  *reinterpret_cast<uint2*>(out + (((long)b * r2 + ii) * r2 + j) * (4L * C) + q * C + c4 * 4) =
      make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}
// g_out[b,(y,x),c] = (g_in ? g_in : 0) + dG[b,(y/2,x/2), q*C + c]   (inverse of the gather)
__global__ void merge_scatter_kernel(const float* __restrict__ dG, const float* __restrict__ g_in, float* __restrict__ g_out,
                                     int B, int res, int C) {
  const int c4n = C / 4, r2 = res / 2;
  const long total = (long)B * res * res * c4n;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c4 = (int)(i % c4n);
  long t = i / c4n;
  const int x = (int)(t % res);
  t /= res;
  const int y = (int)(t % res);
  const int b = (int)(t / res);
  const int q = (y & 1) | ((x & 1) << 1);
  float4 v = *reinterpret_cast<const float4*>(dG + (((long)b * r2 + (y >> 1)) * r2 + (x >> 1)) * (4L * C) + q * C + c4 * 4);
  const long o = (((long)b * res + y) * res + x) * C + c4 * 4;
  if (g_in != nullptr) {
    const float4 w = *reinterpret_cast<const float4*>(g_in + o);
    v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
  }
  *reinterpret_cast<float4*>(g_out + o) = v;
}

// ---- ConvNeXt layer scale residual ---------------------------------------------------------------------
// out = in + gamma * z ; zb = bf16(z) saved for the gamma gradient
__global__ void scale_add_fwd_kernel(const float* __restrict__ in, const float* __restrict__ z, const float* __restrict__ gamma,
                                     float* __restrict__ out, bf16* __restrict__ zb, long rows, int C) {
  const int c4n = C / 4;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * c4n) return;
  const int c4 = (int)(i % c4n);
  const float4 a = reinterpret_cast<const float4*>(in)[i], zz = reinterpret_cast<const float4*>(z)[i];
  const float4 gm = *reinterpret_cast<const float4*>(gamma + c4 * 4);
  reinterpret_cast<float4*>(out)[i] =
      make_float4(fmaf(gm.x, zz.x, a.x), fmaf(gm.y, zz.y, a.y), fmaf(gm.z, zz.z, a.z), fmaf(gm.w, zz.w, a.w));
  reinterpret_cast<uint2*>(zb)[i] = make_uint2(pack_bf16x2(zz.x, zz.y), pack_bf16x2(zz.z, zz.w));
}
// dz = gamma * g (bf16); g_gamma[c] += sum_m g*z ; g_bias[c] += sum_m dz. One block = 64 rows.
__global__ void __launch_bounds__(256)
scale_add_bwd_kernel(const float* __restrict__ g, const bf16* __restrict__ zb, const float* __restrict__ gamma,
                     bf16* __restrict__ dz, float* __restrict__ g_gamma, float* __restrict__ g_bias, long rows, int C) {
  const int c4n = C / 4;
  const long row0 = (long)blockIdx.x * 64;
  for (int c4 = threadIdx.x % 64; c4 < c4n; c4 += 64) {
    const float4 gm = *reinterpret_cast<const float4*>(gamma + c4 * 4);
    float4 sg = make_float4(0.f, 0.f, 0.f, 0.f), sb = sg;
    for (int r = threadIdx.x / 64; r < 64; r += 4) {
      const long row = row0 + r;
      if (row >= rows) break;
      const float4 gv = *reinterpret_cast<const float4*>(g + row * C + c4 * 4);
      const uint2 zr = *reinterpret_cast<const uint2*>(zb + row * C + c4 * 4);
      const float2 z01 = unpack_bf16x2(zr.x), z23 = unpack_bf16x2(zr.y);
      sg.x += gv.x * z01.x; sg.y += gv.y * z01.y; sg.z += gv.z * z23.x; sg.w += gv.w * z23.y;
      const uint2 o = make_uint2(pack_bf16x2(gm.x * gv.x, gm.y * gv.y), pack_bf16x2(gm.z * gv.z, gm.w * gv.w));
      *reinterpret_cast<uint2*>(dz + row * C + c4 * 4) = o;
      const float2 d01 = unpack_bf16x2(o.x), d23 = unpack_bf16x2(o.y);
      sb.x += d01.x; sb.y += d01.y; sb.z += d23.x; sb.w += d23.y;
    }
    atomicAdd(g_gamma + c4 * 4 + 0, sg.x); atomicAdd(g_gamma + c4 * 4 + 1, sg.y);
    atomicAdd(g_gamma + c4 * 4 + 2, sg.z); atomicAdd(g_gamma + c4 * 4 + 3, sg.w);
    atomicAdd(g_bias + c4 * 4 + 0, sb.x); atomicAdd(g_bias + c4 * 4 + 1, sb.y);
    atomicAdd(g_bias + c4 * 4 + 2, sb.z); atomicAdd(g_bias + c4 * 4 + 3, sb.w);
  }
}

// ---- depthwise 7x7 (NHWC fp32, zero padding 3) -------------------------------------------------------
// FLIP=false: out = conv(x, w) + bias ; FLIP=true (data gradient): out = add + conv(x, flipped w)
template <bool FLIP>
__global__ void dwconv7_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                               const float* __restrict__ add, float* __restrict__ out, int B, int res, int C) {
  const int c4n = C / 4;
  const long total = (long)B * res * res * c4n;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c4 = (int)(i % c4n);
  long t = i / c4n;
  const int px = (int)(t % res);
  t /= res;
  const int py = (int)(t % res);
  const int b = (int)(t / res);
  const int c = c4 * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias != nullptr) acc = *reinterpret_cast<const float4*>(bias + c);
  for (int ky = 0; ky < 7; ++ky) {
    const int yy = py + ky - 3;
    if (yy < 0 || yy >= res) continue;
    for (int kx = 0; kx < 7; ++kx) {
      const int xx = px + kx - 3;
      if (xx < 0 || xx >= res) continue;
      const float4 v = *reinterpret_cast<const float4*>(x + (((long)b * res + yy) * res + xx) * C + c);
      const int tap = FLIP ? (6 - ky) * 7 + (6 - kx) : ky * 7 + kx;
      acc.x = fmaf(v.x, w[(c + 0) * 49 + tap], acc.x);
      acc.y = fmaf(v.y, w[(c + 1) * 49 + tap], acc.y);
      acc.z = fmaf(v.z, w[(c + 2) * 49 + tap], acc.z);
      acc.w = fmaf(v.w, w[(c + 3) * 49 + tap], acc.w);
    }
  }
  const long o = (((long)b * res + py) * res + px) * C + c;
  if (add != nullptr) {
    const float4 a = *reinterpret_cast<const float4*>(add + o);
    acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
  }
  *reinterpret_cast<float4*>(out + o) = acc;
}
// weight gradient: g_w[c,ky,kx] += sum_{b,y,x} dout[b,y,x,c] * x[b,y+ky-3,x+kx-3,c]
// block: 64 "tap" threads (49 used) x 4 channel quads; grid = (C/16, pixel chunks)
__global__ void __launch_bounds__(256)
dwconv7_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dout, float* __restrict__ g_w, int B, int res,
                     int C, int pix_per_block) {
  const int tap = threadIdx.x & 63, quad = threadIdx.x >> 6;
  const int c = (blockIdx.x * 4 + quad) * 4;
  if (tap >= 49 || c >= C) return;
  const int ky = tap / 7, kx = tap - ky * 7;
  const long npix = (long)B * res * res;
  const long p0 = (long)blockIdx.y * pix_per_block;
  const long p1 = p0 + pix_per_block < npix ? p0 + pix_per_block : npix;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long p = p0; p < p1; ++p) {
    const int px = (int)(p % res);
    const long t = p / res;
    const int py = (int)(t % res);
    const int yy = py + ky - 3, xx = px + kx - 3;
    if (yy < 0 || yy >= res || xx < 0 || xx >= res) continue;
    const float4 d = *reinterpret_cast<const float4*>(dout + p * C + c);
    const float4 v = *reinterpret_cast<const float4*>(x + (p + (long)(ky - 3) * res + (kx - 3)) * C + c);
    acc.x = fmaf(d.x, v.x, acc.x); acc.y = fmaf(d.y, v.y, acc.y);
    acc.z = fmaf(d.z, v.z, acc.z); acc.w = fmaf(d.w, v.w, acc.w);
  }
  atomicAdd(g_w + (c + 0) * 49 + tap, acc.x);
  atomicAdd(g_w + (c + 1) * 49 + tap, acc.y);
  atomicAdd(g_w + (c + 2) * 49 + tap, acc.z);
  atomicAdd(g_w + (c + 3) * 49 + tap, acc.w);
}

// ---- patch recovery tail: pixel shuffle of the transposed-conv GEMM output + 5x5 mixing conv -----------
// D: [B*(H/ps)*(W/ps), OC*ps*ps] fp32 token-major, column n = (oc, di, dj)  (ConvTranspose2d k=s=ps)
__device__ __forceinline__ float shuffled(const float* __restrict__ D, int b, int c, int y, int x, int H, int W, int OC,
                                          int ps) {
  const int gw = W / ps;
  return D[(((long)b * (H / ps) + y / ps) * gw + x / ps) * (OC * ps * ps) + (c * ps + (y % ps)) * ps + (x % ps)];
}
// pred[b,o,y,x] = sum_{i,dy,dx} P[b,i,y+dy-2,x+dx-2] * w[o,i,dy,dx] (+ residual input) ; masked -> labels
__global__ void conv5_fwd_kernel(const float* __restrict__ D, const float* __restrict__ w, const float* __restrict__ resid,
                                 int resid_channels, const float* __restrict__ labels, const uint8_t* __restrict__ mask,
                                 int mask_mode, float* __restrict__ pred, int B, int OC, int H, int W, int ps) {
  const long total = (long)B * OC * H * W;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = (int)(i % W);
  long t = i / W;
  const int y = (int)(t % H);
  t /= H;
  const int o = (int)(t % OC);
  const int b = (int)(t / OC);
  float acc = 0.f;
  for (int ic = 0; ic < OC; ++ic)
    for (int dy = 0; dy < 5; ++dy) {
      const int yy = y + dy - 2;
      if (yy < 0 || yy >= H) continue;
      for (int dx = 0; dx < 5; ++dx) {
        const int xx = x + dx - 2;
        if (xx < 0 || xx >= W) continue;
        acc = fmaf(shuffled(D, b, ic, yy, xx, H, W, OC, ps), w[((o * OC + ic) * 5 + dy) * 5 + dx], acc);
      }
    }
  if (resid != nullptr) acc += resid[(((long)b * resid_channels + o) * H + y) * W + x];
  if (mask_mode == 1 && mask[b * OC + o]) acc = labels[i];
  if (mask_mode == 2 && mask[i]) acc = labels[i];
  pred[i] = acc;
}
// dD[token, (ic,di,dj)] (bf16) = sum_{o,dy,dx} dpred[b,o,y-dy+2,x-dx+2] * w[o,ic,dy,dx]; also column sums per ic
__global__ void conv5_bwd_data_kernel(const float* __restrict__ dpred, const float* __restrict__ w, bf16* __restrict__ dD,
                                      float* __restrict__ g_bias, int B, int OC, int H, int W, int ps) {
  const long total = (long)B * OC * H * W;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  float val = 0.f;
  int ic = 0;
  if (i < total) {
    const int x = (int)(i % W);
    long t = i / W;
    const int y = (int)(t % H);
    t /= H;
    ic = (int)(t % OC);
    const int b = (int)(t / OC);
    float acc = 0.f;
    for (int o = 0; o < OC; ++o)
      for (int dy = 0; dy < 5; ++dy) {
        const int yy = y - dy + 2;
        if (yy < 0 || yy >= H) continue;
        for (int dx = 0; dx < 5; ++dx) {
          const int xx = x - dx + 2;
          if (xx < 0 || xx >= W) continue;
          acc = fmaf(dpred[(((long)b * OC + o) * H + yy) * W + xx], w[((o * OC + ic) * 5 + dy) * 5 + dx], acc);
        }
      }
    const int gw = W / ps;
    const bf16 r = __float2bfloat16_rn(acc);
    dD[(((long)b * (H / ps) + y / ps) * gw + x / ps) * (OC * ps * ps) + (ic * ps + (y % ps)) * ps + (x % ps)] = r;
    val = __bfloat162float(r);
  }
  // all threads of a block share (b, ic) when H*W is a multiple of the block size: block-reduce the bias grad
  __shared__ float sred[kThreads / 32];
  val = warp_sum(val);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = val;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int k = 0; k < kThreads / 32; ++k) s += sred[k];
    if (i < total) atomicAdd(g_bias + ic, s);
  }
}
// g_w[o,ic,dy,dx] += sum_{b,y,x} dpred[b,o,y,x] * P[b,ic,y+dy-2,x+dx-2]
// one thread per (o,ic,dy,dx); blockIdx.x = chunk of (b, y) rows
__global__ void conv5_wgrad_kernel(const float* __restrict__ D, const float* __restrict__ dpred, float* __restrict__ g_w,
                                   int B, int OC, int H, int W, int ps, int rows_per_block) {
  const int nw = OC * OC * 25;
  const int tix = threadIdx.x;
  if (tix >= nw) return;
  const int dx = tix % 5, dy = (tix / 5) % 5, ic = (tix / 25) % OC, o = tix / (25 * OC);
  const long r0 = (long)blockIdx.x * rows_per_block;
  float acc = 0.f;
  for (long r = r0; r < r0 + rows_per_block && r < (long)B * H; ++r) {
    const int b = (int)(r / H), y = (int)(r % H);
    const int yy = y + dy - 2;
    if (yy < 0 || yy >= H) continue;
    const float* dp = dpred + (((long)b * OC + o) * H + y) * W;
    for (int x = 0; x < W; ++x) {
      const int xx = x + dx - 2;
      if (xx < 0 || xx >= W) continue;
      acc = fmaf(dp[x], shuffled(D, b, ic, yy, xx, H, W, OC, ps), acc);
    }
  }
  atomicAdd(g_w + tix, acc);
}

// ---- loss (scOT/model.py:1425-1484) --------------------------------------------------------------------
// sums[g] = sum |pred-y|^p over group g ; sums[G+g] = sum |y|^p.  group of channel c from `slices`.
struct LossGroups {
  int n;          // number of groups (0 -> one group = all channels, un-normalised loss)
  int bound[10];  // slice boundaries
};
__device__ __forceinline__ int group_of(const LossGroups& lg, int c) {
  if (lg.n == 0) return 0;
  for (int k = 0; k < lg.n; ++k)
    if (c >= lg.bound[k] && c < lg.bound[k + 1]) return k;
  return -1;
}
__global__ void __launch_bounds__(256)
loss_sums_kernel(const float* __restrict__ pred, const float* __restrict__ labels, float* __restrict__ sums, LossGroups lg,
                 int p, int OC, long HW, long total) {
  // grid.x covers one (b,c) plane per blockIdx.y
  const long plane = blockIdx.y;
  const int c = (int)(plane % OC);
  const int gidx = group_of(lg, c);
  float num = 0.f, den = 0.f;
  if (gidx >= 0) {
    for (long k = (long)blockIdx.x * blockDim.x + threadIdx.x; k < HW; k += (long)gridDim.x * blockDim.x) {
      const float y = labels[plane * HW + k], d = pred[plane * HW + k] - y;
      if (p == 1) { num += fabsf(d); den += fabsf(y); } else { num += d * d; den += y * y; }
    }
  }
  __shared__ float s1[8], s2[8];
  num = warp_sum(num); den = warp_sum(den);
  if ((threadIdx.x & 31) == 0) { s1[threadIdx.x >> 5] = num; s2[threadIdx.x >> 5] = den; }
  __syncthreads();
  if (threadIdx.x == 0 && gidx >= 0) {
    float a = 0.f, b = 0.f;
    for (int k = 0; k < 8; ++k) { a += s1[k]; b += s2[k]; }
    const int G = lg.n == 0 ? 1 : lg.n;
    atomicAdd(sums + gidx, a);
    atomicAdd(sums + G + gidx, b);
  }
}
__global__ void loss_final_kernel(const float* __restrict__ sums, float* __restrict__ loss, LossGroups lg, int B, long HW) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (lg.n == 0) {
    loss[0] = sums[0] / (float)((double)B * lg.bound[0] * HW);  // bound[0] carries OC in this mode
    return;
  }
  float acc = 0.f;
  for (int k = 0; k < lg.n; ++k) {
    const float cnt = (float)((double)B * (lg.bound[k + 1] - lg.bound[k]) * HW);
    acc += (sums[k] / cnt) / (sums[lg.n + k] / cnt + 1e-10f);
  }
  loss[0] = acc / (float)lg.n;
}
// dpred = gscale * dloss/dpred (+ dpred_extra) ; zero where the prediction was overwritten by the mask
__global__ void loss_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ labels,
                                const float* __restrict__ sums, const float* __restrict__ gscale,
                                const float* __restrict__ extra, const uint8_t* __restrict__ mask, int mask_mode,
                                float* __restrict__ dpred, LossGroups lg, int p, int B, int OC, long HW, long total) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long plane = i / HW;
  const int c = (int)(plane % OC), b = (int)(plane / OC);
  float gval = 0.f;
  const float gs = gscale != nullptr ? gscale[0] : 0.f;
  if (labels != nullptr && gs != 0.f) {
    const int gidx = group_of(lg, c);
    if (gidx >= 0) {
      const float d = pred[i] - labels[i];
      const float base = (p == 1) ? ((d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f)) : 2.f * d;
      if (lg.n == 0) {
        gval = gs * base / (float)((double)B * OC * HW);
      } else {
        const float cnt = (float)((double)B * (lg.bound[gidx + 1] - lg.bound[gidx]) * HW);
        gval = gs * base / (cnt * (sums[lg.n + gidx] / cnt + 1e-10f) * (float)lg.n);
      }
    }
  }
  if (extra != nullptr) gval += extra[i];
  if (mask_mode == 1 && mask[b * OC + c]) gval = 0.f;
  if (mask_mode == 2 && mask[i]) gval = 0.f;
  dpred[i] = gval;
}

}  // namespace

int scot_cast_f32_bf16_launch(const float* in, void* out, long n, cudaStream_t st) {
  SCOT_REQUIRE(n % 4 == 0, "cast: n must be a multiple of 4");
  cast_f32_bf16_kernel<<<blocks_for(n / 4), kThreads, 0, st>>>(in, (bf16*)out, n / 4);
  SCOT_LAUNCH_CHECK();
  return 0;
}
int scot_expand_bias_launch(const float* bias, float* out, int n, int rep, cudaStream_t st) {
  expand_bias_kernel<<<blocks_for(n), kThreads, 0, st>>>(bias, out, n, rep);
  SCOT_LAUNCH_CHECK();
  return 0;
}
int scot_im2col_patch_launch(const float* x, void* out, int B, int Cin, int H, int W, int ps, cudaStream_t st) {
  SCOT_REQUIRE(H % ps == 0 && W % ps == 0, "im2col: image size must be a multiple of the patch size");
  const long total = (long)B * (H / ps) * (W / ps) * Cin * ps;
  im2col_patch_kernel<<<blocks_for(total), kThreads, 0, st>>>(x, (bf16*)out, B, Cin, H, W, ps);
  SCOT_LAUNCH_CHECK();
  return 0;
}
int scot_merge_gather_launch(const float* x, const float* inp, void* out, int B, int res, int C, cudaStream_t st) {
  SCOT_REQUIRE(res % 2 == 0 && C % 4 == 0, "merge_gather: res must be even");
  merge_gather_kernel<<<blocks_for((long)B * res * res * (C / 4)), kThreads, 0, st>>>(x, inp, (bf16*)out, B, res, C);
  SCOT_LAUNCH_CHECK();
  return 0;
}
int scot_merge_scatter_launch(const float* dG, const float* g_in, float* g_out, int B, int res, int C, cudaStream_t st) {
  merge_scatter_kernel<<<blocks_for((long)B * res * res * (C / 4)), kThreads, 0, st>>>(dG, g_in, g_out, B, res, C);
  SCOT_LAUNCH_CHECK();
  return 0;
}
int scot_scale_add_fwd_launch(const float* in, const float* z, const float* gamma, float* out, void* zb, long rows, int C,
                              cudaStream_t st) {
  scale_add_fwd_kernel<<<blocks_for(rows * (C / 4)), kThreads, 0, st>>>(in, z, gamma, out, (bf16*)zb, rows, C);
  SCOT_LAUNCH_CHECK();
  return 0;
}
int scot_scale_add_bwd_launch(const float* g, const void* zb, const float* gamma, void* dz, float* g_gamma, float* g_bias,
                              long rows, int C, cudaStream_t st) {
  scale_add_bwd_kernel<<<(unsigned)((rows + 63) / 64), 256, 0, st>>>(g, (const bf16*)zb, gamma, (bf16*)dz, g_gamma, g_bias,
                                                                   rows, C);
  SCOT_LAUNCH_CHECK();
  return 0;
}
int scot_dwconv7_fwd_launch(const float* x, const float* w, const float* bias, float* out, int B, int res, int C,
                            cudaStream_t st) {
  dwconv7_kernel<false><<<blocks_for((long)B * res * res * (C / 4)), kThreads, 0, st>>>(x, w, bias, nullptr, out, B, res, C);
  SCOT_LAUNCH_CHECK();
  return 0;
}
int scot_dwconv7_bwd_launch(const float* x, const float* w, const float* dout, const float* g_in, float* g_out, float* g_w,
                            int B, int res, int C, cudaStream_t st) {
  dwconv7_kernel<true><<<blocks_for((long)B * res * res * (C / 4)), kThreads, 0, st>>>(dout, w, nullptr, g_in, g_out, B, res, C);
  SCOT_LAUNCH_CHECK();
  const long npix = (long)B * res * res;
  const int ppb = 512;
  dwconv7_wgrad_kernel<<<dim3((C + 15) / 16, (unsigned)((npix + ppb - 1) / ppb)), 256, 0, st>>>(x, dout, g_w, B, res, C, ppb);
  SCOT_LAUNCH_CHECK();
  return 0;
}
int scot_conv5_fwd_launch(const float* D, const float* w, const float* resid, int resid_channels, const float* labels,
                          const uint8_t* mask, int mask_mode, float* pred, int B, int OC, int H, int W, int ps,
                          cudaStream_t st) {
  SCOT_REQUIRE(mask_mode == 0 || (mask != nullptr && labels != nullptr), "conv5_fwd: mask needs labels");
  conv5_fwd_kernel<<<blocks_for((long)B * OC * H * W), kThreads, 0, st>>>(D, w, resid, resid_channels, labels, mask, mask_mode,
                                                                        pred, B, OC, H, W, ps);
  SCOT_LAUNCH_CHECK();
  return 0;
}
int scot_conv5_bwd_launch(const float* D, const float* w, const float* dpred, void* dD, float* g_w, float* g_bias, int B,
                          int OC, int H, int W, int ps, cudaStream_t st) {
  SCOT_REQUIRE((H * W) % kThreads == 0, "conv5_bwd: H*W must be a multiple of %d", kThreads);
  SCOT_REQUIRE(OC * OC * 25 <= 1024, "conv5_bwd: at most 6 output channels supported");
  conv5_bwd_data_kernel<<<blocks_for((long)B * OC * H * W), kThreads, 0, st>>>(dpred, w, (bf16*)dD, g_bias, B, OC, H, W, ps);
  SCOT_LAUNCH_CHECK();
  const int rpb = 16;
  const int nthr = ((OC * OC * 25 + 31) / 32) * 32;
  conv5_wgrad_kernel<<<(unsigned)(((long)B * H + rpb - 1) / rpb), nthr, 0, st>>>(D, dpred, g_w, B, OC, H, W, ps, rpb);
  SCOT_LAUNCH_CHECK();
  return 0;
}

static int make_groups(LossGroups* lg, const int* slices, int n_slices, int OC) {
  lg->n = 0;
  for (int k = 0; k < 10; ++k) lg->bound[k] = 0;
  if (slices == nullptr || n_slices < 2) {
    lg->bound[0] = OC;
    return 0;
  }
  SCOT_REQUIRE(n_slices <= 10, "loss: at most 9 channel groups supported");
  lg->n = n_slices - 1;
  for (int k = 0; k < n_slices; ++k) lg->bound[k] = slices[k];
  return 0;
}
int scot_loss_fwd_launch(const float* pred, const float* labels, float* sums, float* loss, const int* slices_host,
                         int n_slices, int p, int B, int OC, long HW, cudaStream_t st) {
  LossGroups lg;
  if (int rc = make_groups(&lg, slices_host, n_slices, OC)) return rc;
  SCOT_REQUIRE(p == 1 || p == 2, "loss: p must be 1 or 2");
  SCOT_CHECK_CUDA(cudaMemsetAsync(sums, 0, 20 * sizeof(float), st));
  loss_sums_kernel<<<dim3(4, B * OC), 256, 0, st>>>(pred, labels, sums, lg, p, OC, HW, (long)B * OC * HW);
  SCOT_LAUNCH_CHECK();
  loss_final_kernel<<<1, 32, 0, st>>>(sums, loss, lg, B, HW);
  SCOT_LAUNCH_CHECK();
  return 0;
}
int scot_loss_bwd_launch(const float* pred, const float* labels, const float* sums, const float* gscale, const float* extra,
                         const uint8_t* mask, int mask_mode, float* dpred, const int* slices_host, int n_slices, int p, int B,
                         int OC, long HW, cudaStream_t st) {
  LossGroups lg;
  if (int rc = make_groups(&lg, slices_host, n_slices, OC)) return rc;
  const long total = (long)B * OC * HW;
  loss_bwd_kernel<<<blocks_for(total), kThreads, 0, st>>>(pred, labels, sums, gscale, extra, mask, mask_mode, dpred, lg, p, B,
                                                          OC, HW, total);
  SCOT_LAUNCH_CHECK();
  return 0;
}

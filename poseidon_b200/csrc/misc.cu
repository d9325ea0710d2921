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
__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, long n4, size_t lo_off) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = reinterpret_cast<const float4*>(in)[i];
  st_bf16x4(out + 4 * i, lo_off, v.x, v.y, v.z, v.w);
}

__global__ void expand_bias_kernel(const float* __restrict__ bias, float* __restrict__ out, int n, int rep) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = bias[i / rep];
}

// ---- patch embedding im2col: [B,Cin,H,W] fp32 -> [B*(H/ps)*(W/ps), Cin*ps*ps] bf16, k = (c, di, dj) ------
__global__ void im2col_patch_kernel(const float* __restrict__ x, bf16* __restrict__ out, int B, int Cin, int H, int W,
                                    int ps, size_t lo_off) {
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
  for (int dj = 0; dj < ps; ++dj) st_bf16(dst + dj, lo_off, src[dj]);
}

// ---- patch merging ------------------------------------------------------------------------------------
// out[b, (i,j), q*C + c] = x[b, 2i+(q&1), 2j+(q>>1), c] + inp[...]   q = 0..3 -> (0,0),(1,0),(0,1),(1,1)
__global__ void merge_gather_kernel(const float* __restrict__ x, const float* __restrict__ inp, bf16* __restrict__ out,
                                    int B, int res, int C, size_t lo_off) {
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
  st_bf16x4(out + (((long)b * r2 + ii) * r2 + j) * (4L * C) + q * C + c4 * 4, lo_off, v.x, v.y, v.z, v.w);
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
                                     float* __restrict__ out, bf16* __restrict__ zb, long rows, int C, size_t lo_off) {
  const int c4n = C / 4;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * c4n) return;
  const int c4 = (int)(i % c4n);
  const float4 a = reinterpret_cast<const float4*>(in)[i], zz = reinterpret_cast<const float4*>(z)[i];
  const float4 gm = *reinterpret_cast<const float4*>(gamma + c4 * 4);
  reinterpret_cast<float4*>(out)[i] =
      make_float4(fmaf(gm.x, zz.x, a.x), fmaf(gm.y, zz.y, a.y), fmaf(gm.z, zz.z, a.z), fmaf(gm.w, zz.w, a.w));
  st_bf16x4(zb + 4 * i, lo_off, zz.x, zz.y, zz.z, zz.w);
}
// dz = gamma * g (bf16); g_gamma[c] += sum_m g*z ; g_bias[c] += sum_m dz.
// A block sweeps `rows_per_block` rows; thread = (row phase, 4 channels) with every thread active for any C (256 / (C/4)
// rows per pass), column sums in registers, one smem fold + 8 atomics per column quad and block.
__global__ void __launch_bounds__(256)
scale_add_bwd_kernel(const float* __restrict__ g, const bf16* __restrict__ zb, const float* __restrict__ gamma,
                     bf16* __restrict__ dz, float* __restrict__ g_gamma, float* __restrict__ g_bias, long rows, int C,
                     int rows_per_block, size_t lo_off) {
  extern __shared__ float4 sab_red[];  // [rpp][c4n][2]
  const int c4n = C / 4;
  const int rpp = 256 / c4n;  // rows per pass
  const int c4 = threadIdx.x % c4n, rsub = threadIdx.x / c4n;
  const bool active = rsub < rpp;
  const long row0 = (long)blockIdx.x * rows_per_block;
  const long row1 = row0 + rows_per_block < rows ? row0 + rows_per_block : rows;
  float4 sg = make_float4(0.f, 0.f, 0.f, 0.f), sb = sg;
  if (active) {
    const float4 gm = *reinterpret_cast<const float4*>(gamma + c4 * 4);
#pragma unroll 4
    for (long row = row0 + rsub; row < row1; row += rpp) {
      const float4 gv = *reinterpret_cast<const float4*>(g + row * C + c4 * 4);
      const float4 zr = ld_bf16x4(zb + row * C + c4 * 4, lo_off);
      sg.x += gv.x * zr.x; sg.y += gv.y * zr.y; sg.z += gv.z * zr.z; sg.w += gv.w * zr.w;
      const float4 d = st_bf16x4(dz + row * C + c4 * 4, lo_off, gm.x * gv.x, gm.y * gv.y, gm.z * gv.z, gm.w * gv.w);
      sb.x += d.x; sb.y += d.y; sb.z += d.z; sb.w += d.w;
    }
    sab_red[(rsub * c4n + c4) * 2 + 0] = sg;
    sab_red[(rsub * c4n + c4) * 2 + 1] = sb;
  }
  __syncthreads();
  if (threadIdx.x < c4n) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), bsum = a;
    for (int r = 0; r < rpp; ++r) {
      const float4 u = sab_red[(r * c4n + threadIdx.x) * 2 + 0], v = sab_red[(r * c4n + threadIdx.x) * 2 + 1];
      a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
      bsum.x += v.x; bsum.y += v.y; bsum.z += v.z; bsum.w += v.w;
    }
    const int c = threadIdx.x * 4;
    atomicAdd(g_gamma + c + 0, a.x); atomicAdd(g_gamma + c + 1, a.y);
    atomicAdd(g_gamma + c + 2, a.z); atomicAdd(g_gamma + c + 3, a.w);
    atomicAdd(g_bias + c + 0, bsum.x); atomicAdd(g_bias + c + 1, bsum.y);
    atomicAdd(g_bias + c + 2, bsum.z); atomicAdd(g_bias + c + 3, bsum.w);
  }
}

// ---- depthwise 7x7 (NHWC fp32, zero padding 3) -------------------------------------------------------
// One thread = 8 consecutive x positions x 4 channels: per filter row it loads 14 input float4 for 224 FMAs (register
// sliding window). The filter is staged once per block in shared memory, tap-major ([49][C], already flipped for the data
// gradient), so the inner loop reads its four weights with one conflict-free LDS.128 instead of four strided global loads.
// FLIP=false: out = conv(x, w) + bias ; FLIP=true (data gradient): out = add + conv(x, flipped w)
template <bool FLIP>
__global__ void __launch_bounds__(128)
dwconv7_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
               const float* __restrict__ add, float* __restrict__ out, int B, int res, int C, int cqb) {
  // block = (chunk of `cqb` channel quads, 128 / cqb groups of 8 x positions); blockIdx.y = channel chunk. Blocks are
  // persistent (a few per SM) and stride over the pixel groups, so the filter is staged once per block: coalesced reads of
  // w[c][49], transposed into [49][CB + 4] (pitch + 4 floats: float4-aligned rows, 4-way instead of 32-way store conflicts)
  extern __shared__ __align__(16) float sw[];
  const int CB = 4 * cqb, WP = CB + 4, cbase = blockIdx.y * CB;
  for (int j = threadIdx.x; j < 49 * CB; j += blockDim.x) {
    const int cc = j / 49, tap = j - cc * 49;
    sw[(FLIP ? 48 - tap : tap) * WP + cc] = w[(long)cbase * 49 + j];
  }
  __syncthreads();
  const int xg = (res + 7) / 8, G = blockDim.x / cqb;
  const int cq = threadIdx.x % cqb, gl = threadIdx.x / cqb;
  const long items = (long)B * res * xg;
  if (gl >= G) return;
  const int cl = cq * 4, c = cbase + cl;
  const float4 bz = bias != nullptr ? *reinterpret_cast<const float4*>(bias + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (long it = (long)blockIdx.x * G + gl; it < items; it += (long)gridDim.x * G) {
  long t = it;
  const int gx = (int)(t % xg);
  t /= xg;
  const int py = (int)(t % res);
  const int b = (int)(t / res);
  const int x0 = gx * 8;
  float4 acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = bz;
#pragma unroll 1
  for (int ky = 0; ky < 7; ++ky) {
    const int yy = py + ky - 3;
    if (yy < 0 || yy >= res) continue;
    float4 in[14];
    const float* rowp = x + (((long)b * res + yy) * res) * C + c;
#pragma unroll
    for (int k = 0; k < 14; ++k) {
      const int xx = x0 + k - 3;
      in[k] = (xx >= 0 && xx < res) ? *reinterpret_cast<const float4*>(rowp + (long)xx * C) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float* wrow = sw + (ky * 7) * WP + cl;
#pragma unroll
    for (int kx = 0; kx < 7; ++kx) {
      const float4 wv = *reinterpret_cast<const float4*>(wrow + kx * WP);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        acc[k].x = fmaf(in[k + kx].x, wv.x, acc[k].x);
        acc[k].y = fmaf(in[k + kx].y, wv.y, acc[k].y);
        acc[k].z = fmaf(in[k + kx].z, wv.z, acc[k].z);
        acc[k].w = fmaf(in[k + kx].w, wv.w, acc[k].w);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int xx = x0 + k;
    if (xx >= res) break;
    const long o = (((long)b * res + py) * res + xx) * C + c;
    float4 v = acc[k];
    if (add != nullptr) {
      const float4 a = *reinterpret_cast<const float4*>(add + o);
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
    *reinterpret_cast<float4*>(out + o) = v;
  }
  }
}
// weight gradient: g_w[c,ky,kx] += sum_{b,y,x} dout[b,y,x,c] * x[b,y+ky-3,x+kx-3,c]
// block = (filter row ky, group of images); thread = (channel quad, row lane) sweeps image rows with a 7-wide
// register window, accumulating its 7x4 taps; block-level smem reduction, then one atomic per tap per block.
__global__ void __launch_bounds__(256, 2)
dwconv7_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dout, float* __restrict__ g_w, int B, int res,
                     int C, int imgs_per_block) {
  extern __shared__ float sred[];  // [28][C/4] accumulated with smem atomics
  const int c4n = C / 4;
  const int ky = blockIdx.x;
  (void)imgs_per_block;
  const int lanes = blockDim.x / c4n;            // row lanes
  const int c4 = threadIdx.x % c4n, rl = threadIdx.x / c4n;
  for (int k = threadIdx.x; k < 28 * c4n; k += blockDim.x) sred[k] = 0.f;
  __syncthreads();
  if (rl < lanes) {
    const int c = c4 * 4;
    float4 acc[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    // the B * res image rows are dealt round-robin to (block, row lane): every block gets the same share whatever B is
    const int nrows = B * res;
    for (int r = blockIdx.y * lanes + rl; r < nrows; r += gridDim.y * lanes) {
      const int b = r / res, y = r % res;
      const int yy = y + ky - 3;
      if (yy < 0 || yy >= res) continue;
      const float* inrow = x + (((long)b * res + yy) * res) * C + c;
      const float* drow = dout + (((long)b * res + y) * res) * C + c;
      // 8 output positions per pass: 14 input + 8 gradient float4 are loaded up front (all independent, so their
      // latencies overlap), then 8 x 7 x 4 FMAs run from registers
      for (int x0 = 0; x0 < res; x0 += 8) {
        float4 in[14], d[8];
#pragma unroll
        for (int k = 0; k < 14; ++k) {
          const int xx = x0 + k - 3;
          in[k] = (xx >= 0 && xx < res) ? *reinterpret_cast<const float4*>(inrow + (long)xx * C) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
          d[k] = (x0 + k < res) ? *reinterpret_cast<const float4*>(drow + (long)(x0 + k) * C) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            acc[kx].x = fmaf(d[k].x, in[k + kx].x, acc[kx].x); acc[kx].y = fmaf(d[k].y, in[k + kx].y, acc[kx].y);
            acc[kx].z = fmaf(d[k].z, in[k + kx].z, acc[kx].z); acc[kx].w = fmaf(d[k].w, in[k + kx].w, acc[kx].w);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      atomicAdd(&sred[(k * 4 + 0) * c4n + c4], acc[k].x);
      atomicAdd(&sred[(k * 4 + 1) * c4n + c4], acc[k].y);
      atomicAdd(&sred[(k * 4 + 2) * c4n + c4], acc[k].z);
      atomicAdd(&sred[(k * 4 + 3) * c4n + c4], acc[k].w);
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < 28 * c4n; k += blockDim.x) {
    const int q = k / c4n, cc4 = k - q * c4n;
    const int kx = q >> 2, cj = q & 3;
    atomicAdd(g_w + (cc4 * 4 + cj) * 49 + ky * 7 + kx, sred[k]);
  }
}

// ---- patch recovery tail: pixel shuffle of the transposed-conv GEMM output + 5x5 mixing conv -----------
// D: [B*(H/ps)*(W/ps), OC*ps*ps] fp32 token-major, column n = (oc, di, dj)  (ConvTranspose2d k=s=ps)
// P: planar [B, OC, H, W] fp32 (the ConvTranspose2d output of the reference, model.py:645)
__global__ void unshuffle_kernel(const float* __restrict__ D, float* __restrict__ P, int B, int OC, int H, int W, int ps) {
  const long total = (long)B * OC * H * W;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = (int)(i % W);
  long t = i / W;
  const int y = (int)(t % H);
  t /= H;
  const int c = (int)(t % OC);
  const int b = (int)(t / OC);
  P[i] = D[(((long)b * (H / ps) + y / ps) * (W / ps) + x / ps) * (OC * ps * ps) + (c * ps + (y % ps)) * ps + (x % ps)];
}
// planar gradient -> token-major bf16 [tokens, OC*ps*ps] + ConvTranspose2d bias gradient (sum per channel)
__global__ void shuffle_grad_kernel(const float* __restrict__ dP, bf16* __restrict__ dD, float* __restrict__ g_bias, int B,
                                    int OC, int H, int W, int ps, size_t lo_off) {
  const long total = (long)B * OC * H * W;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  float val = 0.f;
  int c = 0;
  if (i < total) {
    const int x = (int)(i % W);
    long t = i / W;
    const int y = (int)(t % H);
    t /= H;
    c = (int)(t % OC);
    const int b = (int)(t / OC);
    bf16* dst = dD + (((long)b * (H / ps) + y / ps) * (W / ps) + x / ps) * (OC * ps * ps) + (c * ps + (y % ps)) * ps + (x % ps);
    st_bf16(dst, lo_off, dP[i]);
    val = ld_bf16(dst, lo_off);
  }
  // H*W is a multiple of the block size, so a block never straddles two channels
  __shared__ float sred[kThreads / 32];
  val = warp_sum(val);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = val;
  __syncthreads();
  if (threadIdx.x == 0 && i < total) {
    float sacc = 0.f;
    for (int k = 0; k < kThreads / 32; ++k) sacc += sred[k];
    atomicAdd(g_bias + c, sacc);
  }
}

constexpr int C5_TW = 64, C5_TH = 16;  // output tile
// out[b,o,y,x] = sum_{i,dy,dx} in[b,i,y+dy-2,x+dx-2] * w[o,i,dy,dx]                    (TRANSPOSE=false, forward)
// out[b,i,y,x] = sum_{o,dy,dx} in[b,o,y-dy+2,x-dx+2] * w[o,i,dy,dx]                    (TRANSPOSE=true, data gradient)
// 256 threads = 16 x-groups (4 pixels) x 16 rows; halo tile of all channels in smem; per (channel, filter row)
// 8 smem loads feed 4*5*OC FMAs.
template <int OC, bool TRANSPOSE>
__global__ void __launch_bounds__(256)
conv5_tiled_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ resid,
                   int resid_channels, const float* __restrict__ labels, const uint8_t* __restrict__ mask, int mask_mode,
                   float* __restrict__ out, int B, int H, int W) {
  __shared__ float tile[OC][C5_TH + 4][C5_TW + 4];
  __shared__ float sw[OC * OC * 25];  // sw[(oo*OC + ii)*25 + dy*5 + dx], already transposed/flipped if TRANSPOSE
  const int b = blockIdx.z;
  const int ty0 = blockIdx.y * C5_TH, tx0 = blockIdx.x * C5_TW;
  for (int k = threadIdx.x; k < OC * OC * 25; k += 256) {
    if (!TRANSPOSE) {
      sw[k] = w[k];
    } else {
      const int tap = k % 25, ii = (k / 25) % OC, oo = k / (25 * OC);  // output channel oo of this pass = input ch of w
      sw[k] = w[(ii * OC + oo) * 25 + (24 - tap)];
    }
  }
  for (int k = threadIdx.x; k < OC * (C5_TH + 4) * (C5_TW + 4); k += 256) {
    const int xx = k % (C5_TW + 4), yy = (k / (C5_TW + 4)) % (C5_TH + 4), c = k / ((C5_TW + 4) * (C5_TH + 4));
    const int gy = ty0 + yy - 2, gx = tx0 + xx - 2;
    tile[c][yy][xx] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? in[(((long)b * OC + c) * H + gy) * W + gx] : 0.f;
  }
  __syncthreads();
  const int lx = (threadIdx.x & 15) * 4, ly = threadIdx.x >> 4;
  float acc[OC][4];
#pragma unroll
  for (int o = 0; o < OC; ++o) acc[o][0] = acc[o][1] = acc[o][2] = acc[o][3] = 0.f;
#pragma unroll 1
  for (int ic = 0; ic < OC; ++ic) {
#pragma unroll
    for (int dy = 0; dy < 5; ++dy) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = tile[ic][ly + dy][lx + k];
#pragma unroll
      for (int o = 0; o < OC; ++o) {
#pragma unroll
        for (int dx = 0; dx < 5; ++dx) {
          const float ww = sw[(o * OC + ic) * 25 + dy * 5 + dx];
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[o][k] = fmaf(v[k + dx], ww, acc[o][k]);
        }
      }
    }
  }
  const int gy = ty0 + ly;
  if (gy >= H) return;
#pragma unroll
  for (int o = 0; o < OC; ++o) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int gx = tx0 + lx + k;
      if (gx >= W) continue;
      const long idx = (((long)b * OC + o) * H + gy) * W + gx;
      float r = acc[o][k];
      if (!TRANSPOSE) {
        if (resid != nullptr) r += resid[(((long)b * resid_channels + o) * H + gy) * W + gx];
        if (mask_mode == 1 && mask[b * OC + o]) r = labels[idx];
        if (mask_mode == 2 && mask[idx]) r = labels[idx];
      }
      out[idx] = r;
    }
  }
}
// g_w[o,ic,dy,dx] += sum_{b,y,x} dpred[b,o,y,x] * P[b,ic,y+dy-2,x+dx-2].
// Thread = (o, ic, tile row y): it keeps all 25 taps of its (o, ic) pair in registers and slides a 5 x 5 register window
// of P along x, so one tile column costs 6 shared loads for 25 FMAs (the one-thread-per-weight form needs 50). The 16
// row-threads of a pair sit in one half-warp: a shuffle tree folds them, lane 0 issues the 25 atomics once per block.
constexpr int C5W_TW = 32, C5W_TH = 16;
template <int OC>
__global__ void __launch_bounds__((OC * OC * C5W_TH + 31) / 32 * 32)
conv5_wgrad_tiled_kernel(const float* __restrict__ P, const float* __restrict__ dpred, float* __restrict__ g_w, int B, int H,
                         int W) {
  constexpr int TPP = C5W_TW + 5, TDP = C5W_TW + 1;  // padded pitches: bank = 5y / y (conflict free over the 16 rows)
  __shared__ float tp[OC][C5W_TH + 4][TPP];
  __shared__ float td[OC][C5W_TH][TDP];
  constexpr int nthr = (OC * OC * C5W_TH + 31) / 32 * 32;  // whole warps: the shuffle fold below uses the full mask
  const int tix = threadIdx.x;
  const int y = tix % C5W_TH, pair = tix / C5W_TH;
  const bool valid = pair < OC * OC;  // padding threads of the last warp compute on pair 0 and write nothing
  const int ic = valid ? pair % OC : 0, o = valid ? pair / OC : 0;
  float acc[25];
#pragma unroll
  for (int k = 0; k < 25; ++k) acc[k] = 0.f;
  const int tiles_x = W / C5W_TW, tiles_y = H / C5W_TH;
  for (int tile_id = blockIdx.x; tile_id < B * tiles_x * tiles_y; tile_id += gridDim.x) {
    const int b = tile_id / (tiles_x * tiles_y);
    const int tr = tile_id % (tiles_x * tiles_y);
    const int ty0 = (tr / tiles_x) * C5W_TH, tx0 = (tr % tiles_x) * C5W_TW;
    __syncthreads();
    for (int k = tix; k < OC * (C5W_TH + 4) * (C5W_TW + 4); k += nthr) {
      const int xx = k % (C5W_TW + 4), yy = (k / (C5W_TW + 4)) % (C5W_TH + 4), c = k / ((C5W_TW + 4) * (C5W_TH + 4));
      const int gy = ty0 + yy - 2, gx = tx0 + xx - 2;
      tp[c][yy][xx] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? P[(((long)b * OC + c) * H + gy) * W + gx] : 0.f;
    }
    for (int k = tix; k < OC * C5W_TH * C5W_TW; k += nthr) {
      const int xx = k % C5W_TW, yy = (k / C5W_TW) % C5W_TH, c = k / (C5W_TW * C5W_TH);
      td[c][yy][xx] = dpred[(((long)b * OC + c) * H + ty0 + yy) * W + tx0 + xx];
    }
    __syncthreads();
    float win[5][5];  // win[dy][j] = tp[ic][y + dy][x + j]
#pragma unroll
    for (int dy = 0; dy < 5; ++dy)
#pragma unroll
      for (int j = 0; j < 4; ++j) win[dy][j + 1] = tp[ic][y + dy][j];
#pragma unroll 4
    for (int x = 0; x < C5W_TW; ++x) {
      const float d = td[o][y][x];
#pragma unroll
      for (int dy = 0; dy < 5; ++dy) {
#pragma unroll
        for (int j = 0; j < 4; ++j) win[dy][j] = win[dy][j + 1];
        win[dy][4] = tp[ic][y + dy][x + 4];
#pragma unroll
        for (int dx = 0; dx < 5; ++dx) acc[dy * 5 + dx] = fmaf(d, win[dy][dx], acc[dy * 5 + dx]);
      }
    }
  }
  // fold the 16 tile rows of each (o, ic) pair (one half-warp) and add the 25 taps to the gradient
#pragma unroll
  for (int k = 0; k < 25; ++k) {
    float v = acc[k];
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    if (y == 0 && valid) atomicAdd(g_w + pair * 25 + k, v);
  }
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
  cast_f32_bf16_kernel<<<blocks_for(n / 4), kThreads, 0, st>>>(in, (bf16*)out, n / 4, scot_split_off());
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
  im2col_patch_kernel<<<blocks_for(total), kThreads, 0, st>>>(x, (bf16*)out, B, Cin, H, W, ps, scot_split_off());
  SCOT_LAUNCH_CHECK();
  return 0;
}
int scot_merge_gather_launch(const float* x, const float* inp, void* out, int B, int res, int C, cudaStream_t st) {
  SCOT_REQUIRE(res % 2 == 0 && C % 4 == 0, "merge_gather: res must be even");
  merge_gather_kernel<<<blocks_for((long)B * res * res * (C / 4)), kThreads, 0, st>>>(x, inp, (bf16*)out, B, res, C, scot_split_off());
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
  scale_add_fwd_kernel<<<blocks_for(rows * (C / 4)), kThreads, 0, st>>>(in, z, gamma, out, (bf16*)zb, rows, C, scot_split_off());
  SCOT_LAUNCH_CHECK();
  return 0;
}
int scot_scale_add_bwd_launch(const float* g, const void* zb, const float* gamma, void* dz, float* g_gamma, float* g_bias,
                              long rows, int C, cudaStream_t st) {
  SCOT_REQUIRE(C % 4 == 0 && C / 4 <= 256, "scale_add_bwd: C must be a multiple of 4 and at most 1024");
  const int c4n = C / 4, rpp = 256 / c4n;
  // about two blocks per SM, each sweeping a multiple of the rows-per-pass (at least 8 passes)
  long rpb = (rows + 295) / 296;
  if (rpb < 8L * rpp) rpb = 8L * rpp;
  rpb = (rpb + rpp - 1) / rpp * rpp;
  const size_t smem = (size_t)rpp * c4n * 2 * sizeof(float4);
  scale_add_bwd_kernel<<<(unsigned)((rows + rpb - 1) / rpb), 256, smem, st>>>(g, (const bf16*)zb, gamma, (bf16*)dz, g_gamma,
                                                                              g_bias, rows, C, (int)rpb, scot_split_off());
  SCOT_LAUNCH_CHECK();
  return 0;
}
// channel quads per block of the depthwise kernels: the largest divisor of C/4 that is <= 32 (25 KB of staged filter at most)
static int dw_cqb(int C) {
  const int c4n = C / 4;
  int cqb = 1;
  for (int d = 1; d <= 32 && d <= c4n; ++d)
    if (c4n % d == 0) cqb = d;
  return cqb;
}
template <bool FLIP>
static int dwconv7_launch(const float* x, const float* w, const float* bias, const float* add, float* out, int B, int res, int C,
                          cudaStream_t st) {
  SCOT_REQUIRE(C % 4 == 0 && C > 0, "dwconv7: channels must be a multiple of 4");
  const int cqb = dw_cqb(C), G = 128 / cqb;
  const long items = (long)B * res * ((res + 7) / 8);
  const int chunks = (C / 4) / cqb;
  long gx = (items + G - 1) / G;
  const long cap = (4L * 148 + chunks - 1) / chunks;  // one resident wave: 4 blocks per SM (128 registers x 128 threads), all chunks
  if (gx > cap) gx = cap;
  dim3 grid((unsigned)gx, (unsigned)chunks);
  dwconv7_kernel<FLIP><<<grid, 128, (size_t)49 * (4 * cqb + 4) * sizeof(float), st>>>(x, w, bias, add, out, B, res, C, cqb);
  SCOT_LAUNCH_CHECK();
  return 0;
}
int scot_dwconv7_fwd_launch(const float* x, const float* w, const float* bias, float* out, int B, int res, int C,
                            cudaStream_t st) {
  return dwconv7_launch<false>(x, w, bias, nullptr, out, B, res, C, st);
}
int scot_dwconv7_bwd_launch(const float* x, const float* w, const float* dout, const float* g_in, float* g_out, float* g_w,
                            int B, int res, int C, cudaStream_t st) {
  if (int rc = dwconv7_launch<true>(dout, w, nullptr, g_in, g_out, B, res, C, st)) return rc;
  const int c4n = C / 4;
  SCOT_REQUIRE(c4n <= 256, "dwconv7: at most 1024 channels");
  // two resident blocks per SM in total over the 7 filter rows: 42 row groups per filter row
  int groups = (2 * 148) / 7;
  const int lanes = 256 / c4n > 0 ? 256 / c4n : 1;
  const int max_groups = (B * res + lanes - 1) / lanes;
  if (groups > max_groups) groups = max_groups;
  const size_t smem = (size_t)28 * c4n * sizeof(float);
  dwconv7_wgrad_kernel<<<dim3(7, groups), 256, smem, st>>>(x, dout, g_w, B, res, C, 0);
  SCOT_LAUNCH_CHECK();
  return 0;
}

#define C5_DISPATCH(OCV, CALL) \
  case OCV: { CALL; break; }

int scot_unshuffle_launch(const float* D, float* P, int B, int OC, int H, int W, int ps, cudaStream_t st) {
  unshuffle_kernel<<<blocks_for((long)B * OC * H * W), kThreads, 0, st>>>(D, P, B, OC, H, W, ps);
  SCOT_LAUNCH_CHECK();
  return 0;
}
int scot_conv5_fwd_launch(const float* P, const float* w, const float* resid, int resid_channels, const float* labels,
                          const uint8_t* mask, int mask_mode, float* pred, int B, int OC, int H, int W, cudaStream_t st) {
  SCOT_REQUIRE(mask_mode == 0 || (mask != nullptr && labels != nullptr), "conv5_fwd: mask needs labels");
  SCOT_REQUIRE(OC >= 1 && OC <= 6, "conv5: 1..6 output channels supported");
  dim3 grid((W + C5_TW - 1) / C5_TW, (H + C5_TH - 1) / C5_TH, B);
  switch (OC) {
#define C5F(N) C5_DISPATCH(N, (conv5_tiled_kernel<N, false><<<grid, 256, 0, st>>>(P, w, resid, resid_channels, labels, mask, mask_mode, pred, B, H, W)))
    C5F(1) C5F(2) C5F(3) C5F(4) C5F(5) C5F(6)
#undef C5F
  }
  SCOT_LAUNCH_CHECK();
  return 0;
}
// dP_scratch: planar fp32 [B,OC,H,W] scratch for the data gradient before it is shuffled to token-major bf16
int scot_conv5_bwd_launch(const float* P, const float* w, const float* dpred, float* dP_scratch, void* dD, float* g_w,
                          float* g_bias, int B, int OC, int H, int W, int ps, cudaStream_t st) {
  SCOT_REQUIRE((H * W) % kThreads == 0, "conv5_bwd: H*W must be a multiple of %d", kThreads);
  SCOT_REQUIRE(OC >= 1 && OC <= 6, "conv5: 1..6 output channels supported");
  SCOT_REQUIRE(H % C5W_TH == 0 && W % C5W_TW == 0, "conv5_bwd: image size must be a multiple of %dx%d", C5W_TW, C5W_TH);
  dim3 grid((W + C5_TW - 1) / C5_TW, (H + C5_TH - 1) / C5_TH, B);
  const int tiles = B * (W / C5W_TW) * (H / C5W_TH);
  const int wg_blocks = tiles < 592 ? tiles : 592;
  switch (OC) {
#define C5B(N) C5_DISPATCH(N, (conv5_tiled_kernel<N, true><<<grid, 256, 0, st>>>(dpred, w, nullptr, 0, nullptr, nullptr, 0, dP_scratch, B, H, W)); \
                             scot_count_launch(); \
                             (conv5_wgrad_tiled_kernel<N><<<wg_blocks, (N * N * 16 + 31) / 32 * 32, 0, st>>>(P, dpred, g_w, B, H, W)))
    C5B(1) C5B(2) C5B(3) C5B(4) C5B(5) C5B(6)
#undef C5B
  }
  SCOT_LAUNCH_CHECK();
  shuffle_grad_kernel<<<blocks_for((long)B * OC * H * W), kThreads, 0, st>>>(dP_scratch, (bf16*)dD, g_bias, B, OC, H, W, ps, scot_split_off());
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

// Whole-model forward / backward sequencing for scOT (the native "runtime" of this repo).
//
// Mirrors the structure of ScOT.forward (scOT/model.py:1318-1509): embeddings -> encoder stages with patch
// merging -> ConvNeXt blocks on the skips -> decoder stages with patch unmerging -> patch recovery -> loss,
// and the hand-derived reverse pass. All device memory (parameters, gradients,
// workspace arena) belongs to the caller; the engine object only holds host-side plans (offsets), so one
// forward+backward is a fixed sequence of kernel launches on the caller's stream and can be captured in
// a CUDA graph.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "internal.h"
#include "scot_b200.h"

namespace {

struct ParamInfo {
  std::string name;
  long offset;
  long numel;
  int ndim;
  long shape[4];
};

struct NormP {
  long ww = -1, wb = -1, bw = -1, bb = -1;  // conditioned: weight.weight, weight.bias, bias.weight, bias.bias
};                                           // plain LayerNorm: wb = weight, bb = bias, ww = bw = -1
struct BlockP {
  long wqkv, bqkv, ls, cw1, cb1, cw2, wo, bo, w1, b1, w2, b2;
  NormP ln1, ln2;
};
struct MergeP { long wred; NormP norm; };
struct UnmergeP { long wup, wmix; NormP norm; };
struct CnxP { long gamma, wdw, bdw, w1, b1, w2, b2; NormP norm; };

struct Geo {
  int res, C, heads, hd, ws, shift, depth;
  long M;  // batch * res * res
};

// offsets (bytes) into the caller's arena
struct BlockBuf {
  size_t qkv, o, lse, zhat1, rstd1, y1b, h, g, zhat2, rstd2, xout, xbout, tab2, alpha, dtab, dalpha, dpre;
};
struct NormBuf { size_t zhat, rstd; };
struct CnxBuf { size_t nb, h, g, z2b, out, zhat, rstd; };

struct Bump {
  size_t off = 0;
  size_t take(size_t bytes) {
    const size_t o = off;
    off += (bytes + 255) & ~size_t(255);
    return o;
  }
};

}  // namespace

struct ScotEngine {
  ScotModelDesc d;
  int batch;
  int ns;
  int hidden_mult_num;  // mlp hidden = int(mlp_ratio * C)
  std::vector<Geo> geo;
  std::vector<ParamInfo> params;
  long n_elems = 0;  // flat parameter buffer length (elements)
  // parameter offsets
  long emb_w, emb_b;
  NormP emb_norm;
  std::vector<std::vector<BlockP>> enc, dec;  // [stage][block] (dec indexed by decoder layer j)
  std::vector<MergeP> merge;                  // [stage s] for s < ns-1
  std::vector<UnmergeP> unmerge;              // [decoder layer j] valid when stage > 0
  std::vector<std::vector<CnxP>> cnx;         // [stage][k]
  long rec_w, rec_b, rec_mix;
  // arena plan
  size_t ws_bytes = 0;
  size_t split_off = 0;  // parity precision: byte distance hi -> lo twin of every bf16 arena tensor (a shadow arena)
  size_t wb16;  // bf16 copy of the flat parameters
  size_t p16, emb_zhat, emb_rstd, emb_x, emb_xb;
  std::vector<std::vector<BlockBuf>> ebuf, dbuf;
  std::vector<size_t> m_g16, m_x, m_xb;  // merge: gathered input (bf16), output fp32 / bf16
  std::vector<NormBuf> m_norm, u_norm;
  std::vector<size_t> u_nb, u_x, u_xb;   // unmerge: normalised (bf16), decoder stage input fp32 / bf16
  std::vector<std::vector<CnxBuf>> cbuf;
  std::vector<size_t> skip_xb;           // bf16 copy of the post-ConvNeXt skip of the last stage (if any)
  size_t rec_bias, rec_D, rec_P, pred_copy, loss_sums;
  size_t z32, y32;                       // fp32 scratch [M0*C0] each
  // backward scratch
  size_t dzb, dzb2, dqkv, dh, dob, partial, dpre, dpred, dD16, dgrads_zero_begin, dgrads_zero_bytes;
  size_t dzbB, dzb2B, dqkvB, dhB, dobB;  // second set of block-backward scratch (blocks alternate, see block_bwd)
  size_t partial_bytes;
  std::vector<size_t> gstage;            // fp32 [M_s, C_s]
  // relative-position-bias MLPs of all attention layers (<= 64 per table), split by WHEN their gradients are final in the
  // backward pass: "early" = decoder + deepest encoder stage, "late" = the remaining encoder stages (see
  // scot_engine_backward_part)
  std::vector<ScotCpbTable> cpb_early, cpb_late;
  long split_elem = 0;  // flat-buffer offset of the first parameter of the deepest encoder stage
  // forward state needed by backward
  const float* last_pixels = nullptr;
  const float* last_time = nullptr;
  const float* last_labels = nullptr;
  const uint8_t* last_mask = nullptr;
  int last_mask_mode = 0;
  float* last_pred = nullptr;
  bool have_forward = false;
  // side stream for the weight-gradient GEMMs of the block backward (fork/join with events; capturable)
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork[2] = {nullptr, nullptr}, ev_join[2] = {nullptr, nullptr};
  bool join_pending[2] = {false, false};
  int overlap = -1;  // SCOT_WGRAD_OVERLAP (default on)
  // second side stream: the ConvNeXt blocks on the skip connections are side branches of the U-Net (needed only when the
  // decoder / the encoder backward reaches their stage), so they run concurrently with the latency-bound deep stages
  cudaStream_t side2 = nullptr;
  cudaEvent_t ev_cfork[4] = {nullptr, nullptr, nullptr, nullptr}, ev_cjoin[4] = {nullptr, nullptr, nullptr, nullptr};
  bool cjoin_pending[4] = {false, false, false, false};
  int cnx_overlap = -1;  // SCOT_CNX_OVERLAP
  size_t z32C = 0, y32C = 0, dzbC = 0, dhC = 0;  // private scratch of the ConvNeXt side branch
  // third stream: dk/dv kernel of the window-attention backward beside the dq kernel (small-window stages)
  cudaStream_t side3 = nullptr;
  cudaEvent_t ev_afork = nullptr, ev_ajoin = nullptr;
  int attn_split_ws = -1;  // SCOT_ATTN_BWD_SPLIT = largest window size that forks (0 = never)
};

namespace {

long align_up(long v, long a) { return (v + a - 1) / a * a; }

struct Registry {
  ScotEngine* e;
  long cursor = 0;
  long alloc(long numel, long align = 64) {
    cursor = align_up(cursor, align);
    const long o = cursor;
    cursor += numel;
    return o;
  }
  void reg(const std::string& name, long off, std::initializer_list<long> shape) {
    ParamInfo p;
    p.name = name;
    p.offset = off;
    p.ndim = (int)shape.size();
    p.numel = 1;
    int i = 0;
    for (long s : shape) {
      p.shape[i++] = s;
      p.numel *= s;
    }
    for (; i < 4; ++i) p.shape[i] = 1;
    e->params.push_back(p);
  }
  long add(const std::string& name, std::initializer_list<long> shape) {
    long n = 1;
    for (long s : shape) n *= s;
    const long o = alloc(n);
    reg(name, o, shape);
    return o;
  }
  NormP norm(const std::string& pre, long C, bool cond) {
    NormP n;
    if (cond) {
      n.ww = add(pre + ".weight.weight", {C, 1});
      n.wb = add(pre + ".weight.bias", {C});
      n.bw = add(pre + ".bias.weight", {C, 1});
      n.bb = add(pre + ".bias.bias", {C});
    } else {
      n.wb = add(pre + ".weight", {C});
      n.bb = add(pre + ".bias", {C});
    }
    return n;
  }
  BlockP block(const std::string& pre, long C, long heads, long hidden, bool cond) {
    BlockP b;
    const std::string a = pre + ".attention.self";
    b.ls = add(a + ".logit_scale", {heads, 1, 1});
    b.cw1 = add(a + ".continuous_position_bias_mlp.0.weight", {512, 2});
    b.cb1 = add(a + ".continuous_position_bias_mlp.0.bias", {512});
    b.cw2 = add(a + ".continuous_position_bias_mlp.2.weight", {heads, 512});
    // q, k, v weights contiguous -> one [3C, C] GEMM operand
    b.wqkv = alloc(3 * C * C);
    reg(a + ".query.weight", b.wqkv, {C, C});
    reg(a + ".key.weight", b.wqkv + C * C, {C, C});
    reg(a + ".value.weight", b.wqkv + 2 * C * C, {C, C});
    // [bq | zeros (key has no bias, HF:417) | bv] contiguous -> one [3C] epilogue bias
    b.bqkv = alloc(3 * C);
    reg(a + ".query.bias", b.bqkv, {C});
    reg(a + ".value.bias", b.bqkv + 2 * C, {C});
    b.wo = add(pre + ".attention.output.dense.weight", {C, C});
    b.bo = add(pre + ".attention.output.dense.bias", {C});
    b.ln1 = norm(pre + ".layernorm_before", C, cond);
    b.w1 = add(pre + ".intermediate.dense.weight", {hidden, C});
    b.b1 = add(pre + ".intermediate.dense.bias", {hidden});
    b.w2 = add(pre + ".output.dense.weight", {C, hidden});
    b.b2 = add(pre + ".output.dense.bias", {C});
    b.ln2 = norm(pre + ".layernorm_after", C, cond);
    return b;
  }
};

long hidden_of(const ScotEngine* e, long C) { return (long)(e->d.mlp_ratio * (float)C); }

int build_params(ScotEngine* e) {
  const ScotModelDesc& d = e->d;
  const bool cond = d.use_conditioning != 0;
  Registry R{e};
  const long C0 = d.embed_dim, ps = d.patch_size;
  e->emb_w = R.add("embeddings.patch_embeddings.projection.weight", {C0, d.num_channels, ps, ps});
  e->emb_b = R.add("embeddings.patch_embeddings.projection.bias", {C0});
  e->emb_norm = R.norm("embeddings.norm", C0, cond);
  e->enc.resize(e->ns);
  e->merge.resize(e->ns);
  for (int s = 0; s < e->ns; ++s) {
    const Geo& g = e->geo[s];
    for (int i = 0; i < g.depth; ++i)
      e->enc[s].push_back(R.block("encoder.layers." + std::to_string(s) + ".blocks." + std::to_string(i), g.C, g.heads,
                                  hidden_of(e, g.C), cond));
    if (s < e->ns - 1) {
      const std::string pre = "encoder.layers." + std::to_string(s) + ".downsample";
      e->merge[s].wred = R.add(pre + ".reduction.weight", {2L * g.C, 4L * g.C});
      e->merge[s].norm = R.norm(pre + ".norm", 2L * g.C, cond);
    }
  }
  e->split_elem = e->enc[e->ns - 1].empty() ? R.cursor : e->enc[e->ns - 1][0].ls;  // first parameter of the deepest stage
  e->split_elem = e->split_elem / 64 * 64;
  e->dec.resize(e->ns);
  e->unmerge.resize(e->ns);
  for (int j = 0; j < e->ns; ++j) {
    const int s = e->ns - 1 - j;
    const Geo& g = e->geo[s];
    for (int i = 0; i < g.depth; ++i)
      e->dec[j].push_back(R.block("decoder.layers." + std::to_string(j) + ".blocks." + std::to_string(i), g.C, g.heads,
                                  hidden_of(e, g.C), cond));
    if (s > 0) {
      const std::string pre = "decoder.layers." + std::to_string(j) + ".upsample";
      e->unmerge[j].wup = R.add(pre + ".upsample.weight", {2L * g.C, (long)g.C});
      e->unmerge[j].wmix = R.add(pre + ".mixup.weight", {g.C / 2L, g.C / 2L});
      e->unmerge[j].norm = R.norm(pre + ".norm", g.C / 2L, cond);
    }
  }
  e->rec_w = R.add("patch_recovery.projection.weight", {C0, d.num_out_channels, ps, ps});
  e->rec_b = R.add("patch_recovery.projection.bias", {d.num_out_channels});
  e->rec_mix = R.add("patch_recovery.mixup.weight", {d.num_out_channels, d.num_out_channels, 5, 5});
  e->cnx.resize(e->ns);
  for (int s = 0; s < e->ns; ++s) {
    const Geo& g = e->geo[s];
    for (int k = 0; k < d.skip_blocks[s]; ++k) {
      const std::string pre = "residual_blocks." + std::to_string(s) + "." + std::to_string(k);
      CnxP c;
      c.gamma = R.add(pre + ".weight", {(long)g.C});
      c.wdw = R.add(pre + ".dwconv.weight", {(long)g.C, 1, 7, 7});
      c.bdw = R.add(pre + ".dwconv.bias", {(long)g.C});
      c.norm = R.norm(pre + ".norm", g.C, cond);
      c.w1 = R.add(pre + ".pwconv1.weight", {4L * g.C, (long)g.C});
      c.b1 = R.add(pre + ".pwconv1.bias", {4L * g.C});
      c.w2 = R.add(pre + ".pwconv2.weight", {(long)g.C, 4L * g.C});
      c.b2 = R.add(pre + ".pwconv2.bias", {(long)g.C});
      e->cnx[s].push_back(c);
    }
  }
  e->n_elems = align_up(R.cursor, 64);
  return 0;
}

void plan_block(ScotEngine* e, Bump& b, BlockBuf& bb, const Geo& g) {
  const size_t M = (size_t)g.M, C = (size_t)g.C, H = (size_t)hidden_of(e, g.C);
  const size_t units = (size_t)e->batch * (g.res / g.ws) * (g.res / g.ws) * g.heads;
  const size_t tabn = (size_t)(2 * g.ws - 1) * (2 * g.ws - 1) * g.heads;
  bb.qkv = b.take(M * 3 * C * 2);
  bb.o = b.take(M * C * 2);
  bb.lse = b.take(units * g.ws * g.ws * 4);
  bb.zhat1 = b.take(M * C * 2);
  bb.rstd1 = b.take(M * 4);
  bb.y1b = b.take(M * C * 2);
  bb.h = b.take(M * H * 2);
  bb.g = b.take(M * H * 2);
  bb.zhat2 = b.take(M * C * 2);
  bb.rstd2 = b.take(M * 4);
  bb.xout = b.take(M * C * 4);
  bb.xbout = b.take(M * C * 2);
  bb.tab2 = b.take(tabn * 4);
  bb.alpha = b.take((size_t)g.heads * 4);
}

int build_plan(ScotEngine* e) {
  const ScotModelDesc& d = e->d;
  Bump b;
  const Geo& g0 = e->geo[0];
  const size_t M0 = (size_t)g0.M, C0 = (size_t)g0.C;
  const size_t K0 = (size_t)d.num_channels * d.patch_size * d.patch_size;
  const size_t NR = (size_t)d.num_out_channels * d.patch_size * d.patch_size;
  e->wb16 = b.take((size_t)e->n_elems * 2);
  e->p16 = b.take(M0 * K0 * 2);
  e->emb_zhat = b.take(M0 * C0 * 2);
  e->emb_rstd = b.take(M0 * 4);
  e->emb_x = b.take(M0 * C0 * 4);
  e->emb_xb = b.take(M0 * C0 * 2);
  e->ebuf.resize(e->ns);
  e->dbuf.resize(e->ns);
  e->m_g16.assign(e->ns, 0); e->m_x.assign(e->ns, 0); e->m_xb.assign(e->ns, 0);
  e->m_norm.resize(e->ns); e->u_norm.resize(e->ns);
  e->u_nb.assign(e->ns, 0); e->u_x.assign(e->ns, 0); e->u_xb.assign(e->ns, 0);
  e->cbuf.resize(e->ns);
  e->skip_xb.assign(e->ns, 0);
  for (int s = 0; s < e->ns; ++s) {
    const Geo& g = e->geo[s];
    e->ebuf[s].resize(g.depth);
    for (int i = 0; i < g.depth; ++i) plan_block(e, b, e->ebuf[s][i], g);
    if (s < e->ns - 1) {
      const Geo& gn = e->geo[s + 1];
      e->m_g16[s] = b.take((size_t)gn.M * 4 * g.C * 2);
      e->m_x[s] = b.take((size_t)gn.M * gn.C * 4);
      e->m_xb[s] = b.take((size_t)gn.M * gn.C * 2);
      e->m_norm[s].zhat = b.take((size_t)gn.M * gn.C * 2);
      e->m_norm[s].rstd = b.take((size_t)gn.M * 4);
    }
    e->cbuf[s].resize(d.skip_blocks[s]);
    for (int k = 0; k < d.skip_blocks[s]; ++k) {
      CnxBuf& c = e->cbuf[s][k];
      const size_t M = (size_t)g.M, C = (size_t)g.C;
      c.nb = b.take(M * C * 2);
      c.h = b.take(M * 4 * C * 2);
      c.g = b.take(M * 4 * C * 2);
      c.z2b = b.take(M * C * 2);
      c.out = b.take(M * C * 4);
      c.zhat = b.take(M * C * 2);
      c.rstd = b.take(M * 4);
    }
    if (d.skip_blocks[s] > 0) e->skip_xb[s] = b.take((size_t)g.M * g.C * 2);
  }
  for (int j = 0; j < e->ns; ++j) {
    const int s = e->ns - 1 - j;
    const Geo& g = e->geo[s];
    e->dbuf[j].resize(g.depth);
    for (int i = 0; i < g.depth; ++i) plan_block(e, b, e->dbuf[j][i], g);
    if (s > 0) {
      const Geo& gf = e->geo[s - 1];  // finer stage
      e->u_nb[j] = b.take((size_t)gf.M * gf.C * 2);
      e->u_norm[j].zhat = b.take((size_t)gf.M * gf.C * 2);
      e->u_norm[j].rstd = b.take((size_t)gf.M * 4);
      e->u_x[j] = b.take((size_t)gf.M * gf.C * 4);
      e->u_xb[j] = b.take((size_t)gf.M * gf.C * 2);
    }
  }
  e->rec_bias = b.take(NR * 4);
  e->rec_D = b.take(M0 * NR * 4);
  e->rec_P = b.take(M0 * NR * 4);
  e->loss_sums = b.take(64 * 4);
  e->z32 = b.take(M0 * C0 * 4);
  e->y32 = b.take(M0 * C0 * 4);
  // ---- backward scratch ----
  size_t max_h = 0;
  for (int s = 0; s < e->ns; ++s) {
    const size_t h = (size_t)e->geo[s].M * (size_t)hidden_of(e, e->geo[s].C);
    const size_t h2 = (size_t)e->geo[s].M * 4 * e->geo[s].C;
    max_h = h > max_h ? h : max_h;
    max_h = h2 > max_h ? h2 : max_h;
  }
  e->dzb = b.take(M0 * C0 * 2);
  e->dzb2 = b.take(M0 * C0 * 2);
  e->dqkv = b.take(M0 * 3 * C0 * 2);
  e->dh = b.take(max_h * 2);
  e->dob = b.take(M0 * C0 * 2);
  e->dzbB = b.take(M0 * C0 * 2);
  e->dzb2B = b.take(M0 * C0 * 2);
  e->dqkvB = b.take(M0 * 3 * C0 * 2);
  e->dhB = b.take(max_h * 2);
  e->dobB = b.take(M0 * C0 * 2);
  size_t pb = 0;
  for (int s = 0; s < e->ns; ++s) {
    const Geo& g = e->geo[s];
    const size_t v = scot_attn_bwd_partial_bytes(g.ws, g.heads, e->batch * (g.res / g.ws) * (g.res / g.ws));
    pb = v > pb ? v : pb;
  }
  e->partial_bytes = pb;
  e->partial = b.take(pb);
  e->dpre = 0;
  e->dpred = b.take((size_t)e->batch * d.num_out_channels * d.image_size * d.image_size * 4);
  e->dD16 = b.take(M0 * NR * 2);
  // per-layer bias-table gradient accumulators (zeroed at the start of every backward)
  e->dgrads_zero_begin = b.off;
  auto plan_dt = [&](BlockBuf& bb, const Geo& g) {
    bb.dtab = b.take((size_t)(2 * g.ws - 1) * (2 * g.ws - 1) * g.heads * 4);
    bb.dalpha = b.take((size_t)g.heads * 4);
  };
  auto plan_dpre = [&](BlockBuf& bb, const Geo& g) { bb.dpre = b.take((size_t)(2 * g.ws - 1) * (2 * g.ws - 1) * g.heads * 4); };
  for (int s = 0; s < e->ns; ++s)
    for (auto& bb : e->ebuf[s]) plan_dt(bb, e->geo[s]);
  for (int j = 0; j < e->ns; ++j)
    for (auto& bb : e->dbuf[j]) plan_dt(bb, e->geo[e->ns - 1 - j]);
  e->dgrads_zero_bytes = b.off - e->dgrads_zero_begin;
  for (int s = 0; s < e->ns; ++s)
    for (auto& bb : e->ebuf[s]) plan_dpre(bb, e->geo[s]);
  for (int j = 0; j < e->ns; ++j)
    for (auto& bb : e->dbuf[j]) plan_dpre(bb, e->geo[e->ns - 1 - j]);
  // descriptor tables for the batched bias-MLP kernels
  {
    std::vector<ScotCpbLayer> late, early;
    auto add_layer = [&](std::vector<ScotCpbLayer>& all, const BlockP& p, const BlockBuf& bb, const Geo& g) {
      ScotCpbLayer L;
      L.w1 = (int)p.cw1; L.b1 = (int)p.cb1; L.w2 = (int)p.cw2; L.ls = (int)p.ls;
      L.tab2 = (int)(bb.tab2 / 256); L.alpha = (int)(bb.alpha / 256);
      L.dtab = (int)(bb.dtab / 256); L.dalpha = (int)(bb.dalpha / 256); L.dpre = (int)(bb.dpre / 256);
      L.ws = (short)g.ws; L.heads = (short)g.heads;
      all.push_back(L);
    };
    for (int s = 0; s < e->ns; ++s)
      for (int i = 0; i < e->geo[s].depth; ++i) add_layer(s == e->ns - 1 ? early : late, e->enc[s][i], e->ebuf[s][i], e->geo[s]);
    for (int j = 0; j < e->ns; ++j)
      for (int i = 0; i < e->geo[e->ns - 1 - j].depth; ++i) add_layer(early, e->dec[j][i], e->dbuf[j][i], e->geo[e->ns - 1 - j]);
    auto pack = [&](const std::vector<ScotCpbLayer>& all, std::vector<ScotCpbTable>& out) {
      for (size_t i = 0; i < all.size(); i += SCOT_CPB_MAX_LAYERS) {
        ScotCpbTable t;
        memset(&t, 0, sizeof(t));
        t.n = (int)std::min<size_t>(SCOT_CPB_MAX_LAYERS, all.size() - i);
        for (int k = 0; k < t.n; ++k) t.layer[k] = all[i + k];
        out.push_back(t);
      }
    };
    pack(early, e->cpb_early);
    pack(late, e->cpb_late);
  }
  e->gstage.resize(e->ns);
  for (int s = 0; s < e->ns; ++s) e->gstage[s] = b.take((size_t)e->geo[s].M * e->geo[s].C * 4);
  {
    // scratch of the ConvNeXt side branch (largest stage that has skip blocks and is not the deepest one)
    size_t mc = 0;
    for (int s = 0; s + 1 < e->ns; ++s)
      if (d.skip_blocks[s] > 0) mc = std::max(mc, (size_t)e->geo[s].M * e->geo[s].C);
    if (mc > 0) {
      e->z32C = b.take(mc * 4);
      e->y32C = b.take(mc * 4);
      e->dzbC = b.take(mc * 2);
      e->dhC = b.take(mc * 4 * 2);
    }
  }
  e->ws_bytes = b.off;
  if (e->d.precision == 1) {
    // shadow arena: the lo twin of the bf16 tensor at offset X lives at X + split_off (fp32 tensors leave theirs unused)
    e->split_off = (b.off + 255) & ~size_t(255);
    e->ws_bytes = 2 * e->split_off;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// execution context helpers
// ---------------------------------------------------------------------------------------------------
struct Ctx {
  ScotEngine* e;
  const float* P;   // fp32 parameters
  float* G;         // fp32 gradients (backward only)
  uint8_t* A;       // arena
  const float* time;
  cudaStream_t st;
  int impl;
  template <typename T>
  T* at(size_t off) const { return reinterpret_cast<T*>(A + off); }
  const bf16* w16(long off) const { return reinterpret_cast<const bf16*>(A + e->wb16) + off; }
  const float* p(long off) const { return off < 0 ? nullptr : P + off; }
  float* g(long off) const { return off < 0 ? nullptr : G + off; }
};

// sets the thread-local split offset for the duration of one engine call
struct SplitGuard {
  explicit SplitGuard(size_t off) { scot_set_split_off(off); }
  ~SplitGuard() { scot_set_split_off(0); }
};

#define RC(expr)                 \
  do {                           \
    int _rc = (expr);            \
    if (_rc != 0) return _rc;    \
  } while (0)

int gemm(const Ctx& c, const void* A, long lda, int amn, const void* B, long ldb, int bmn, long M, long N, long K, int mode,
         const float* bias, void* out0, long ld0, void* out1 = nullptr, long ld1 = 0, const void* aux = nullptr,
         long ldaux = 0, float* colsum = nullptr) {
  ScotEpilogue ep{mode, bias, out0, ld0, out1, ld1, aux, ldaux, colsum};
  return scot_gemm_launch(A, lda, amn, B, ldb, bmn, (int)M, (int)N, (int)K, &ep, c.impl, c.st);
}

int norm_fwd(const Ctx& c, const NormP& n, const float* z, const float* residual, float* x_out, void* xb_out, void* zhat,
             float* rstd, long rows, int C, int rows_per_sample, int perm_res, float eps) {
  return scot_cln_fwd_launch(z, residual, n.ww >= 0 ? c.time : nullptr, c.p(n.ww), c.p(n.wb), c.p(n.bw), c.p(n.bb), x_out,
                             xb_out, zhat, rstd, rows, C, rows_per_sample, perm_res, eps, c.st);
}
int norm_bwd(const Ctx& c, const NormP& n, const float* dy, const void* zhat, const float* rstd, void* dz, int dz_f32,
             float* g_bias_prev, long rows, int C, int rows_per_sample, int perm_res) {
  return scot_cln_bwd_launch(dy, zhat, rstd, n.ww >= 0 ? c.time : nullptr, c.p(n.ww), c.p(n.wb), dz, dz_f32, c.g(n.ww),
                             c.g(n.wb), c.g(n.bw), c.g(n.bb), g_bias_prev, rows, C, rows_per_sample, perm_res, c.st);
}

// ---------------------------------------------------------------------------------------------------
// ScOTLayer forward / backward (scOT/model.py:500-581)
// ---------------------------------------------------------------------------------------------------
// A block can run on a SAMPLE RANGE of the batch (token-major layouts: a range of samples is a contiguous range of rows of
// every per-token tensor): `BlkView` holds the block's buffers shifted to the first row of the range. (Round 2 measured two
// half-batch chains on two streams at the deep stages: 20.63 ms vs 20.15 ms per Poseidon-B step — slower, dropped; the
// engine and the stand-alone layer entry points run whole-batch views.)
struct BlkView {
  bf16 *qkv, *o;
  float* lse;
  bf16 *zhat1, *y1b, *h, *gl, *zhat2, *xbout;
  float *rstd1, *rstd2, *xout;
  float *tab2, *alpha, *dtab, *dalpha;
  float *z32, *y32;                         // fp32 scratch
  bf16 *dzb, *dzb2, *dh, *dob, *dqkv;       // backward scratch
  const float* time;                        // lead times of the range
  long M;                                   // rows of the range
  int batch;                                // samples of the range
};

BlkView view_of(const Ctx& c, const BlockBuf& b, const Geo& g, int sample0, int nsamples, int set) {
  const ScotEngine* e = c.e;
  const long r0 = (long)sample0 * g.res * g.res, C = g.C, H = hidden_of(e, g.C);
  BlkView v;
  v.qkv = c.at<bf16>(b.qkv) + r0 * 3 * C;
  v.o = c.at<bf16>(b.o) + r0 * C;
  v.lse = c.at<float>(b.lse) + r0 * g.heads;  // units * N = samples * res^2 * heads
  v.zhat1 = c.at<bf16>(b.zhat1) + r0 * C;
  v.rstd1 = c.at<float>(b.rstd1) + r0;
  v.y1b = c.at<bf16>(b.y1b) + r0 * C;
  v.h = c.at<bf16>(b.h) + r0 * H;
  v.gl = c.at<bf16>(b.g) + r0 * H;
  v.zhat2 = c.at<bf16>(b.zhat2) + r0 * C;
  v.rstd2 = c.at<float>(b.rstd2) + r0;
  v.xout = c.at<float>(b.xout) + r0 * C;
  v.xbout = c.at<bf16>(b.xbout) + r0 * C;
  v.tab2 = c.at<float>(b.tab2);
  v.alpha = c.at<float>(b.alpha);
  v.dtab = c.at<float>(b.dtab);
  v.dalpha = c.at<float>(b.dalpha);
  v.z32 = c.at<float>(e->z32) + r0 * C;
  v.y32 = c.at<float>(e->y32) + r0 * C;
  v.dzb = c.at<bf16>(set ? e->dzbB : e->dzb) + r0 * C;
  v.dzb2 = c.at<bf16>(set ? e->dzb2B : e->dzb2) + r0 * C;
  v.dh = c.at<bf16>(set ? e->dhB : e->dh) + r0 * H;
  v.dob = c.at<bf16>(set ? e->dobB : e->dob) + r0 * C;
  v.dqkv = c.at<bf16>(set ? e->dqkvB : e->dqkv) + r0 * 3 * C;
  v.time = c.time != nullptr ? c.time + sample0 : nullptr;
  v.M = (long)nsamples * g.res * g.res;
  v.batch = nsamples;
  return v;
}

int norm_fwd_t(const Ctx& c, const NormP& n, const float* time, const float* z, const float* residual, float* x_out, void* xb_out,
               void* zhat, float* rstd, long rows, int C, int rows_per_sample, float eps) {
  return scot_cln_fwd_launch(z, residual, n.ww >= 0 ? time : nullptr, c.p(n.ww), c.p(n.wb), c.p(n.bw), c.p(n.bb), x_out, xb_out,
                             zhat, rstd, rows, C, rows_per_sample, 0, eps, c.st);
}
int norm_bwd_t(const Ctx& c, const NormP& n, const float* time, const float* dy, const void* zhat, const float* rstd, void* dz,
               float* g_bias_prev, long rows, int C, int rows_per_sample) {
  return scot_cln_bwd_launch(dy, zhat, rstd, n.ww >= 0 ? time : nullptr, c.p(n.ww), c.p(n.wb), dz, 0, c.g(n.ww), c.g(n.wb),
                             c.g(n.bw), c.g(n.bb), g_bias_prev, rows, C, rows_per_sample, 0, c.st);
}

// ---------------------------------------------------------------------------------------------------
// ScOTLayer forward / backward (scOT/model.py:500-581) on the rows of `v`; everything is enqueued on c.st
// ---------------------------------------------------------------------------------------------------
int block_fwd_v(const Ctx& c, const BlockP& p, const BlkView& v, const Geo& g, int shift, const float* x_in, const bf16* xb_in) {
  const ScotEngine* e = c.e;
  const long M = v.M, C = g.C, H = hidden_of(e, g.C);
  const int T = g.res * g.res;
  RC(gemm(c, xb_in, C, 0, c.w16(p.wqkv), C, 0, M, 3 * C, C, SCOT_EPI_BF16, c.p(p.bqkv), v.qkv, 3 * C));
  RC(scot_attn_fwd_launch(v.qkv, v.o, v.lse, v.tab2, v.alpha, v.batch, g.res, g.ws, shift, g.heads, g.hd, c.st));
  RC(gemm(c, v.o, C, 0, c.w16(p.wo), C, 0, M, C, C, SCOT_EPI_F32, c.p(p.bo), v.z32, C));
  RC(norm_fwd_t(c, p.ln1, v.time, v.z32, x_in, v.y32, v.y1b, v.zhat1, v.rstd1, M, (int)C, T, e->d.layer_norm_eps));
  RC(gemm(c, v.y1b, C, 0, c.w16(p.w1), C, 0, M, H, C, SCOT_EPI_GELU, c.p(p.b1), v.h, H, v.gl, H));
  RC(gemm(c, v.gl, H, 0, c.w16(p.w2), H, 0, M, C, H, SCOT_EPI_F32, c.p(p.b2), v.z32, C));
  RC(norm_fwd_t(c, p.ln2, v.time, v.z32, v.y32, v.xout, v.xbout, v.zhat2, v.rstd2, M, (int)C, T, e->d.layer_norm_eps));
  return 0;
}
int block_fwd(const Ctx& c, const BlockP& p, const BlockBuf& b, const Geo& g, int shift, const float* x_in,
              const bf16* xb_in) {
  return block_fwd_v(c, p, view_of(c, b, g, 0, c.e->batch, 0), g, shift, x_in, xb_in);
}

// main stream waits until the side-stream work that used scratch set `set` has finished
int join_side(const Ctx& c, int set) {
  ScotEngine* e = c.e;
  if (e->join_pending[set]) {
    SCOT_CHECK_CUDA(cudaStreamWaitEvent(c.st, e->ev_join[set], 0));
    e->join_pending[set] = false;
  }
  return 0;
}
int join_all(const Ctx& c) {
  RC(join_side(c, 0));
  RC(join_side(c, 1));
  return 0;
}

// gr: gradient wrt the block output on entry, wrt the block input on exit (fp32, in place; rows of `v`).
// side_ok: the four weight-gradient GEMMs may go to the side stream (full-batch path): their operands stay alive until the
// end of the block (separate dz buffers for the two norms), consecutive blocks alternate between two sets of scratch buffers
// (`set`), so the weight gradients of block i run concurrently with the data-gradient chain of block i-1 (fork / join through
// events, captured into the CUDA graph as parallel branches). Sample-range calls (side_ok = false) run everything in line on
// c.st: their concurrency comes from the other half of the batch.
int block_bwd_v(const Ctx& c, const BlockP& p, const BlkView& v, const Geo& g, int shift, float* gr, const bf16* xb_in, int set,
                bool side_ok) {
  ScotEngine* e = c.e;
  const long M = v.M, C = g.C, H = hidden_of(e, g.C);
  const int T = g.res * g.res;
  if (side_ok) RC(join_side(c, set));  // the weight gradients issued two blocks ago read this scratch set
  // y = y1 + LN2(mlp(y1))
  RC(norm_bwd_t(c, p.ln2, v.time, gr, v.zhat2, v.rstd2, v.dzb2, c.g(p.b2), M, (int)C, T));
  RC(gemm(c, v.dzb2, C, 0, c.w16(p.w2), H, 1, M, H, C, SCOT_EPI_GELU_BWD, nullptr, v.dh, H, nullptr, 0, v.h, H, c.g(p.b1)));
  RC(gemm(c, v.dh, H, 0, c.w16(p.w1), C, 1, M, C, H, SCOT_EPI_RMW_F32, nullptr, gr, C));
  // y1 = x + LN1(attn(x))
  RC(norm_bwd_t(c, p.ln1, v.time, gr, v.zhat1, v.rstd1, v.dzb, c.g(p.bo), M, (int)C, T));
  RC(gemm(c, v.dzb, C, 0, c.w16(p.wo), C, 1, M, C, C, SCOT_EPI_BF16, nullptr, v.dob, C));
  ScotAttnBwdFork fk{e->side3, e->ev_afork, e->ev_ajoin};
  const bool fork_attn = side_ok && e->side3 != nullptr && g.ws <= e->attn_split_ws;
  RC(scot_attn_bwd_launch2(v.qkv, v.o, v.dob, v.lse, v.tab2, v.alpha, v.dqkv, c.at<float>(e->partial), e->partial_bytes, v.dtab,
                           v.dalpha, c.g(p.bqkv), c.g(p.bqkv + 2 * C), v.batch, g.res, g.ws, shift, g.heads, g.hd, c.st,
                           fork_attn ? &fk : nullptr));
  const ScotWgradProblem wg[4] = {
      {v.dzb2, C, v.gl, H, c.g(p.w2), H, M, (int)C, (int)H},              // output.dense
      {v.dh, H, v.y1b, C, c.g(p.w1), C, M, (int)H, (int)C},               // intermediate.dense
      {v.dzb, C, v.o, C, c.g(p.wo), C, M, (int)C, (int)C},                // attention.output.dense
      {v.dqkv, 3 * C, xb_in, C, c.g(p.wqkv), C, M, (int)(3 * C), (int)C},  // query | key | value
  };
  if (side_ok && e->overlap && e->side != nullptr) {
    SCOT_CHECK_CUDA(cudaEventRecord(e->ev_fork[set], c.st));  // all four dY operands are complete here
    SCOT_CHECK_CUDA(cudaStreamWaitEvent(e->side, e->ev_fork[set], 0));
    RC(scot_gemm_wgrad_group_launch(wg, 4, c.impl, e->side));
    SCOT_CHECK_CUDA(cudaEventRecord(e->ev_join[set], e->side));
    e->join_pending[set] = true;
    RC(gemm(c, v.dqkv, 3 * C, 0, c.w16(p.wqkv), C, 1, M, C, 3 * C, SCOT_EPI_RMW_F32, nullptr, gr, C));
  } else {
    RC(gemm(c, v.dqkv, 3 * C, 0, c.w16(p.wqkv), C, 1, M, C, 3 * C, SCOT_EPI_RMW_F32, nullptr, gr, C));
    RC(scot_gemm_wgrad_group_launch(wg, 4, c.impl, c.st));
  }
  return 0;
}
int block_bwd(const Ctx& c, const BlockP& p, const BlockBuf& b, const Geo& g, int shift, float* gr, const bf16* xb_in, int set) {
  return block_bwd_v(c, p, view_of(c, b, g, 0, c.e->batch, set), g, shift, gr, xb_in, set, true);
}

// ---------------------------------------------------------------------------------------------------
// ConvNeXt blocks on the skip connection of stage s (scOT/model.py:198-217, 1388-1393) and their backward.
// `z32` / `y32` / `dzb` / `dh` are scratch buffers: the main-stream instances when the blocks run in line, the private
// *C set when they run as a side branch on e->side2.
// ---------------------------------------------------------------------------------------------------
struct CnxScratch {
  float* z32;
  float* y32;
  bf16* dzb;
  bf16* dh;
};
CnxScratch cnx_scratch(const Ctx& c, bool side) {
  const ScotEngine* e = c.e;
  if (side) return CnxScratch{c.at<float>(e->z32C), c.at<float>(e->y32C), c.at<bf16>(e->dzbC), c.at<bf16>(e->dhC)};
  return CnxScratch{c.at<float>(e->z32), c.at<float>(e->y32), c.at<bf16>(e->dzb), c.at<bf16>(e->dh)};
}

// returns (through skip_io) the fp32 output of the last block
int cnx_fwd_stage(const Ctx& c, int s, const float** skip_io, const CnxScratch& sc) {
  const ScotEngine* e = c.e;
  const ScotModelDesc& d = e->d;
  const Geo& g = e->geo[s];
  const int B = e->batch;
  const float* skip = *skip_io;
  for (int k = 0; k < d.skip_blocks[s]; ++k) {
    const CnxP& p = e->cnx[s][k];
    const CnxBuf& b = e->cbuf[s][k];
    RC(scot_dwconv7_fwd_launch(skip, c.p(p.wdw), c.p(p.bdw), sc.z32, B, g.res, g.C, c.st));
    RC(norm_fwd(c, p.norm, sc.z32, nullptr, nullptr, c.at<bf16>(b.nb), c.at<bf16>(b.zhat), c.at<float>(b.rstd), g.M, g.C,
                g.res * g.res, 0, d.layer_norm_eps));
    RC(gemm(c, c.at<bf16>(b.nb), g.C, 0, c.w16(p.w1), g.C, 0, g.M, 4L * g.C, g.C, SCOT_EPI_GELU, c.p(p.b1), c.at<bf16>(b.h),
            4L * g.C, c.at<bf16>(b.g), 4L * g.C));
    RC(gemm(c, c.at<bf16>(b.g), 4L * g.C, 0, c.w16(p.w2), 4L * g.C, 0, g.M, g.C, 4L * g.C, SCOT_EPI_F32, c.p(p.b2), sc.z32,
            g.C));
    RC(scot_scale_add_fwd_launch(skip, sc.z32, c.p(p.gamma), c.at<float>(b.out), c.at<bf16>(b.z2b), g.M, g.C, c.st));
    skip = c.at<float>(b.out);
  }
  *skip_io = skip;
  return 0;
}

// gr: gradient wrt the post-ConvNeXt skip on entry, wrt the encoder stage output on exit (fp32, in place)
int cnx_bwd_stage(const Ctx& c, int s, float* gr, const CnxScratch& sc) {
  const ScotEngine* e = c.e;
  const ScotModelDesc& d = e->d;
  const Geo& g = e->geo[s];
  const int B = e->batch;
  for (int k = d.skip_blocks[s] - 1; k >= 0; --k) {
    const CnxP& p = e->cnx[s][k];
    const CnxBuf& b = e->cbuf[s][k];
    const float* blk_in = (k == 0) ? c.at<float>(e->ebuf[s].back().xout) : c.at<float>(e->cbuf[s][k - 1].out);
    RC(scot_scale_add_bwd_launch(gr, c.at<bf16>(b.z2b), c.p(p.gamma), sc.dzb, c.g(p.gamma), c.g(p.b2), g.M, g.C, c.st));
    RC(gemm(c, sc.dzb, g.C, 0, c.w16(p.w2), 4L * g.C, 1, g.M, 4L * g.C, g.C, SCOT_EPI_GELU_BWD, nullptr, sc.dh, 4L * g.C,
            nullptr, 0, c.at<bf16>(b.h), 4L * g.C, c.g(p.b1)));
    RC(gemm(c, sc.dh, 4L * g.C, 0, c.w16(p.w1), g.C, 1, g.M, g.C, 4L * g.C, SCOT_EPI_F32, nullptr, sc.z32, g.C));
    const ScotWgradProblem wg[2] = {
        {sc.dzb, g.C, c.at<bf16>(b.g), 4L * g.C, c.g(p.w2), 4L * g.C, g.M, g.C, 4 * g.C},  // pwconv2
        {sc.dh, 4L * g.C, c.at<bf16>(b.nb), g.C, c.g(p.w1), g.C, g.M, 4 * g.C, g.C},        // pwconv1
    };
    RC(scot_gemm_wgrad_group_launch(wg, 2, c.impl, c.st));
    RC(norm_bwd(c, p.norm, sc.z32, c.at<bf16>(b.zhat), c.at<float>(b.rstd), sc.y32, 1, c.g(p.bdw), g.M, g.C, g.res * g.res,
                0));
    RC(scot_dwconv7_bwd_launch(blk_in, c.p(p.wdw), sc.y32, gr, gr, c.g(p.wdw), B, g.res, g.C, c.st));
  }
  return 0;
}

// Side-branch plumbing: fork e->side2 from the main stream, run `fn` there, record the join event of stage s.
int cnx_side_init(ScotEngine* e) {
  if (e->cnx_overlap < 0) {
    const char* ev = getenv("SCOT_CNX_OVERLAP");  // default on; SCOT_CNX_OVERLAP=0 runs the blocks in line (A/B, tests)
    e->cnx_overlap = (ev != nullptr && ev[0] == '0') ? 0 : 1;
  }
  if (e->cnx_overlap && e->z32C == 0 && e->dhC == 0) e->cnx_overlap = 0;  // nothing to overlap
  if (e->cnx_overlap && e->side2 == nullptr) {
    SCOT_CHECK_CUDA(cudaStreamCreateWithFlags(&e->side2, cudaStreamNonBlocking));
    for (int k = 0; k < 4; ++k) {
      SCOT_CHECK_CUDA(cudaEventCreateWithFlags(&e->ev_cfork[k], cudaEventDisableTiming));
      SCOT_CHECK_CUDA(cudaEventCreateWithFlags(&e->ev_cjoin[k], cudaEventDisableTiming));
    }
  }
  for (int k = 0; k < 4; ++k) e->cjoin_pending[k] = false;
  return 0;
}
bool cnx_on_side(const ScotEngine* e, int s) { return e->cnx_overlap && s + 1 < e->ns && e->d.skip_blocks[s] > 0; }
int cnx_fork(const Ctx& c, int s) {
  ScotEngine* e = c.e;
  SCOT_CHECK_CUDA(cudaEventRecord(e->ev_cfork[s], c.st));
  SCOT_CHECK_CUDA(cudaStreamWaitEvent(e->side2, e->ev_cfork[s], 0));
  return 0;
}
int cnx_mark_join(const Ctx& c, int s) {
  ScotEngine* e = c.e;
  SCOT_CHECK_CUDA(cudaEventRecord(e->ev_cjoin[s], e->side2));
  e->cjoin_pending[s] = true;
  return 0;
}
int cnx_join(const Ctx& c, int s) {
  ScotEngine* e = c.e;
  if (e->cjoin_pending[s]) {
    SCOT_CHECK_CUDA(cudaStreamWaitEvent(c.st, e->ev_cjoin[s], 0));
    e->cjoin_pending[s] = false;
  }
  return 0;
}
int cnx_join_all(const Ctx& c) {
  for (int s = 0; s < c.e->ns; ++s) RC(cnx_join(c, s));
  return 0;
}

int shift_of(const Geo& g, int orig_index) { return (orig_index % 2 == 0) ? 0 : g.shift; }

}  // namespace


// ===================================================================================================
// Stand-alone ScOTLayer entry points: one transformer block through the same block_fwd /
// block_bwd sequencers the whole-model engine uses. The caller's workspace holds, in this order: a flat fp32 copy of the
// layer's parameters (engine layout), the flat gradient buffer, and the block's arena (bf16 weights, saved activations,
// scratch). Parameter order of `params` / `grads` = the reference's state_dict order of a layer:
//   attention.self.{logit_scale, continuous_position_bias_mlp.0.weight, .0.bias, .2.weight, query.weight, query.bias,
//   key.weight, value.weight, value.bias}, attention.output.dense.{weight, bias}, layernorm_before.*, intermediate.dense.
//   {weight, bias}, output.dense.{weight, bias}, layernorm_after.*   (norms: 4 tensors conditioned, 2 otherwise)
// ===================================================================================================
namespace {

struct LayerPlan {
  ScotEngine e;            // host-side plan only (ns = 1, one block)
  std::vector<long> poff;  // flat offsets of the caller's tensors, in `params` order
  std::vector<long> pnum;
  size_t off_params, off_grads, off_arena, off_xb, off_gr, total;
};

int make_layer_plan(const ScotLayerDesc& L, LayerPlan& P) {
  SCOT_REQUIRE(L.batch > 0 && L.res > 0 && L.C > 0 && L.heads > 0 && L.C % L.heads == 0, "layer: bad descriptor");
  SCOT_REQUIRE(L.res % L.window == 0 && (L.shift == 0 || L.shift == L.window / 2), "layer: window %d / shift %d do not fit res %d",
               L.window, L.shift, L.res);
  ScotEngine& e = P.e;
  memset(&e.d, 0, sizeof(e.d));
  e.d.num_stages = 1;
  e.d.mlp_ratio = L.mlp_ratio;
  e.d.use_conditioning = L.use_conditioning;
  e.d.layer_norm_eps = L.layer_norm_eps;
  e.d.window_size = L.window;
  e.d.precision = L.precision;
  e.batch = L.batch;
  e.ns = 1;
  Geo g;
  g.res = L.res; g.C = L.C; g.heads = L.heads; g.hd = L.C / L.heads; g.ws = L.window; g.shift = L.shift; g.depth = 1;
  g.M = (long)L.batch * L.res * L.res;
  SCOT_REQUIRE((g.ws == 16 || g.ws == 8 || g.ws == 4) && (g.hd == 16 || g.hd == 32 || g.hd == 64) && g.C % 16 == 0 &&
               (long)(L.mlp_ratio * g.C) % 16 == 0, "layer: unsupported window %d / head_dim %d", g.ws, g.hd);
  e.geo.assign(1, g);
  Registry R{&e};
  e.enc.assign(1, std::vector<BlockP>());
  e.enc[0].push_back(R.block("layer", g.C, g.heads, hidden_of(&e, g.C), L.use_conditioning != 0));
  e.n_elems = align_up(R.cursor, 64);
  // `params` order = registration order, except that the Registry registers q/k/v weights and q/v biases separately
  // already in state_dict order: logit_scale, cpb.0.weight, cpb.0.bias, cpb.2.weight, q.w, k.w, v.w, q.b, v.b, ...
  // -> map to the documented order (q.w, q.b, k.w, v.w, v.b)
  std::vector<int> order;
  for (size_t i = 0; i < e.params.size(); ++i) order.push_back((int)i);
  // registered: 0 ls, 1 cpb0.w, 2 cpb0.b, 3 cpb2.w, 4 q.w, 5 k.w, 6 v.w, 7 q.b, 8 v.b, 9.. rest in state_dict order
  const int fix[9] = {0, 1, 2, 3, 4, 7, 5, 6, 8};
  for (int i = 0; i < 9; ++i) order[i] = fix[i];
  for (int i : order) {
    P.poff.push_back(e.params[i].offset);
    P.pnum.push_back(e.params[i].numel);
  }
  Bump b;
  P.off_params = b.take((size_t)e.n_elems * 4);
  P.off_grads = b.take((size_t)e.n_elems * 4);
  P.off_arena = b.off;
  // arena of the block (offsets relative to off_arena; the split-bf16 shadow doubles it)
  Bump a;
  e.wb16 = a.take((size_t)e.n_elems * 2);
  e.ebuf.assign(1, std::vector<BlockBuf>(1));
  BlockBuf& bb = e.ebuf[0][0];
  plan_block(&e, a, bb, g);
  const size_t M = (size_t)g.M, C = (size_t)g.C, H = (size_t)hidden_of(&e, g.C);
  e.z32 = a.take(M * C * 4);
  e.y32 = a.take(M * C * 4);
  e.dzb = a.take(M * C * 2);
  e.dzb2 = a.take(M * C * 2);
  e.dqkv = a.take(M * 3 * C * 2);
  e.dh = a.take(M * H * 2);
  e.dob = a.take(M * C * 2);
  e.partial_bytes = 256;
  e.partial = a.take(256);
  const size_t tabn = (size_t)(2 * g.ws - 1) * (2 * g.ws - 1) * g.heads;
  bb.dtab = a.take(tabn * 4);
  bb.dalpha = a.take((size_t)g.heads * 4);
  bb.dpre = a.take(tabn * 4);
  P.off_xb = a.take(M * C * 2);  // bf16 copy of the layer input
  P.off_gr = a.take(M * C * 4);
  e.split_off = 0;
  size_t arena_bytes = a.off;
  if (L.precision == 1) {
    e.split_off = (a.off + 255) & ~size_t(255);
    arena_bytes = 2 * e.split_off;
  }
  ScotCpbTable t;
  memset(&t, 0, sizeof(t));
  t.n = 1;
  const BlockP& p = e.enc[0][0];
  ScotCpbLayer& cl = t.layer[0];
  cl.w1 = (int)p.cw1; cl.b1 = (int)p.cb1; cl.w2 = (int)p.cw2; cl.ls = (int)p.ls;
  cl.tab2 = (int)(bb.tab2 / 256); cl.alpha = (int)(bb.alpha / 256);
  cl.dtab = (int)(bb.dtab / 256); cl.dalpha = (int)(bb.dalpha / 256); cl.dpre = (int)(bb.dpre / 256);
  cl.ws = (short)g.ws; cl.heads = (short)g.heads;
  e.cpb_early.assign(1, t);
  P.total = P.off_arena + arena_bytes;
  e.overlap = 0;  // stand-alone layer: everything on the caller's stream
  e.attn_split_ws = 0;
  return 0;
}

}  // namespace

extern "C" {

int scot_layer_num_params(const ScotLayerDesc* L) { return L == nullptr ? 0 : (L->use_conditioning ? 23 : 19); }

size_t scot_layer_workspace_bytes(const ScotLayerDesc* L) {
  if (L == nullptr) return 0;
  LayerPlan P;
  if (make_layer_plan(*L, P) != 0) return 0;
  return P.total + 256;
}

int scot_layer_fwd(const ScotLayerDesc* L, const void* const* params, const float* x, const float* time, float* y,
                   void* workspace, size_t ws_bytes, void* stream) {
  SCOT_REQUIRE(L && params && x && y && workspace, "layer_fwd: null pointer");
  SCOT_REQUIRE(!L->use_conditioning || time != nullptr, "layer_fwd: time is required when use_conditioning=True");
  LayerPlan P;
  RC(make_layer_plan(*L, P));
  SCOT_REQUIRE(ws_bytes >= P.total && ((uintptr_t)workspace & 255) == 0, "layer_fwd: workspace too small (%zu < %zu) or not 256 B aligned",
               ws_bytes, P.total);
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* ws = (uint8_t*)workspace;
  float* flat = reinterpret_cast<float*>(ws + P.off_params);
  SCOT_CHECK_CUDA(cudaMemsetAsync(flat, 0, (size_t)P.e.n_elems * 4, st));  // the key-bias slot between q and v biases stays 0
  for (size_t i = 0; i < P.poff.size(); ++i)
    SCOT_CHECK_CUDA(cudaMemcpyAsync(flat + P.poff[i], params[i], (size_t)P.pnum[i] * 4, cudaMemcpyDeviceToDevice, st));
  Ctx c{&P.e, flat, nullptr, ws + P.off_arena, time, st, SCOT_GEMM_TCGEN05};
  SplitGuard split_guard(P.e.split_off);
  const Geo& g = P.e.geo[0];
  RC(scot_cast_f32_bf16_launch(flat, c.at<bf16>(P.e.wb16), P.e.n_elems, st));
  RC(scot_cpb_fwd_launch(&P.e.cpb_early[0], flat, ws + P.off_arena, st));
  RC(scot_cast_f32_bf16_launch(x, c.at<bf16>(P.off_xb), g.M * g.C, st));
  RC(block_fwd(c, P.e.enc[0][0], P.e.ebuf[0][0], g, g.shift, x, c.at<bf16>(P.off_xb)));
  SCOT_CHECK_CUDA(cudaMemcpyAsync(y, c.at<float>(P.e.ebuf[0][0].xout), (size_t)g.M * g.C * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}

int scot_layer_bwd(const ScotLayerDesc* L, void* const* grads, const float* time, const float* dy, float* dx, void* workspace,
                   size_t ws_bytes, void* stream) {
  SCOT_REQUIRE(L && grads && dy && dx && workspace, "layer_bwd: null pointer");
  LayerPlan P;
  RC(make_layer_plan(*L, P));
  SCOT_REQUIRE(ws_bytes >= P.total && ((uintptr_t)workspace & 255) == 0, "layer_bwd: workspace too small or not 256 B aligned");
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* ws = (uint8_t*)workspace;
  float* flat = reinterpret_cast<float*>(ws + P.off_params);   // parameters as left by scot_layer_fwd
  float* gflat = reinterpret_cast<float*>(ws + P.off_grads);
  Ctx c{&P.e, flat, gflat, ws + P.off_arena, time, st, SCOT_GEMM_TCGEN05};
  SplitGuard split_guard(P.e.split_off);
  const Geo& g = P.e.geo[0];
  const BlockBuf& bb = P.e.ebuf[0][0];
  SCOT_CHECK_CUDA(cudaMemsetAsync(gflat, 0, (size_t)P.e.n_elems * 4, st));
  SCOT_CHECK_CUDA(cudaMemsetAsync(c.at<uint8_t>(bb.dtab), 0, (bb.dpre - bb.dtab), st));
  float* gr = c.at<float>(P.off_gr);
  SCOT_CHECK_CUDA(cudaMemcpyAsync(gr, dy, (size_t)g.M * g.C * 4, cudaMemcpyDeviceToDevice, st));
  RC(block_bwd(c, P.e.enc[0][0], bb, g, g.shift, gr, c.at<bf16>(P.off_xb), 0));
  RC(scot_cpb_bwd_launch(&P.e.cpb_early[0], flat, gflat, ws + P.off_arena, st));
  SCOT_CHECK_CUDA(cudaMemcpyAsync(dx, gr, (size_t)g.M * g.C * 4, cudaMemcpyDeviceToDevice, st));
  for (size_t i = 0; i < P.poff.size(); ++i)
    if (grads[i] != nullptr)
      SCOT_CHECK_CUDA(cudaMemcpyAsync(grads[i], gflat + P.poff[i], (size_t)P.pnum[i] * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}

}  // extern "C"

// ===================================================================================================
// C ABI
// ===================================================================================================
extern "C" {

int scot_engine_create(const ScotModelDesc* desc, int batch, ScotEngine** out) {
  SCOT_REQUIRE(desc && out && batch > 0, "engine_create: bad arguments");
  const ScotModelDesc& d = *desc;
  SCOT_REQUIRE(d.num_stages >= 1 && d.num_stages <= 4, "engine_create: 1..4 stages supported (got %d)", d.num_stages);
  SCOT_REQUIRE(d.image_size % d.patch_size == 0, "engine_create: image_size must be a multiple of patch_size");
  SCOT_REQUIRE(d.mlp_ratio > 0.f, "engine_create: mlp_ratio");
  SCOT_REQUIRE(d.precision == 0 || d.precision == 1, "engine_create: precision must be 0 (bf16) or 1 (parity / split-bf16)");
  SCOT_REQUIRE(d.num_out_channels * d.num_out_channels * 25 <= 1024, "engine_create: at most 6 output channels");
  ScotEngine* e = new ScotEngine();
  e->d = d;
  e->batch = batch;
  e->ns = d.num_stages;
  const int grid = d.image_size / d.patch_size;
  for (int s = 0; s < e->ns; ++s) {
    Geo g;
    g.res = grid >> s;
    g.C = d.embed_dim << s;
    g.heads = d.num_heads[s];
    g.depth = d.depths[s];
    if (g.res < 1 || (grid % (1 << s)) != 0 || g.C % g.heads != 0) {
      delete e;
      SCOT_REQUIRE(false, "engine_create: stage %d geometry invalid (res %d, C %d, heads %d)", s, g.res, g.C, g.heads);
    }
    g.hd = g.C / g.heads;
    g.ws = g.res <= d.window_size ? g.res : d.window_size;        // scOT/model.py:428-430
    g.shift = g.res <= g.ws ? 0 : d.window_size / 2;               // scOT/model.py:431-440
    g.M = (long)batch * g.res * g.res;
    const bool ok = (g.ws == 16 || g.ws == 8 || g.ws == 4) && (g.hd == 16 || g.hd == 32 || g.hd == 64) &&
                    g.res % g.ws == 0 && g.C % 16 == 0 && (long)(d.mlp_ratio * g.C) % 16 == 0;
    if (!ok) {
      delete e;
      SCOT_REQUIRE(false,
                   "engine_create: stage %d unsupported (window %d must be 16/8/4, head_dim %d must be 16/32/64, res %d)", s,
                   g.ws, g.hd, g.res);
    }
    e->geo.push_back(g);
  }
  build_params(e);
  build_plan(e);
  *out = e;
  return 0;
}

void scot_engine_destroy(ScotEngine* e) {
  if (e == nullptr) return;
  if (e->side != nullptr) {
    cudaStreamDestroy(e->side);
    for (int k = 0; k < 2; ++k) {
      if (e->ev_fork[k]) cudaEventDestroy(e->ev_fork[k]);
      if (e->ev_join[k]) cudaEventDestroy(e->ev_join[k]);
    }
  }
  if (e->side3 != nullptr) {
    cudaStreamDestroy(e->side3);
    if (e->ev_afork) cudaEventDestroy(e->ev_afork);
    if (e->ev_ajoin) cudaEventDestroy(e->ev_ajoin);
  }
  if (e->side2 != nullptr) {
    cudaStreamDestroy(e->side2);
    for (int k = 0; k < 4; ++k) {
      if (e->ev_cfork[k]) cudaEventDestroy(e->ev_cfork[k]);
      if (e->ev_cjoin[k]) cudaEventDestroy(e->ev_cjoin[k]);
    }
  }
  delete e;
}

long scot_engine_num_params(const ScotEngine* e) { return (long)e->params.size(); }
long scot_engine_param_elems(const ScotEngine* e) { return e->n_elems; }
size_t scot_engine_workspace_bytes(const ScotEngine* e) { return e->ws_bytes; }

int scot_engine_param_info(const ScotEngine* e, long i, char* name, int name_cap, long* offset, long* numel, int* ndim,
                           long* shape4) {
  SCOT_REQUIRE(e && i >= 0 && i < (long)e->params.size(), "param_info: index out of range");
  const ParamInfo& p = e->params[i];
  if (name && name_cap > 0) {
    strncpy(name, p.name.c_str(), name_cap - 1);
    name[name_cap - 1] = 0;
  }
  if (offset) *offset = p.offset;
  if (numel) *numel = p.numel;
  if (ndim) *ndim = p.ndim;
  if (shape4)
    for (int k = 0; k < 4; ++k) shape4[k] = p.shape[k];
  return 0;
}

int scot_engine_forward(ScotEngine* e, const float* params, void* arena, const float* pixel_values, const float* time,
                        const float* labels, const uint8_t* mask, int mask_mode, float* pred_out, float* loss_out,
                        int gemm_impl, void* stream) {
  SCOT_REQUIRE(e && params && arena && pixel_values && pred_out, "engine_forward: null pointer");
  const ScotModelDesc& d = e->d;
  SCOT_REQUIRE(!d.use_conditioning || time != nullptr, "engine_forward: time is required when use_conditioning=True");
  SCOT_REQUIRE(labels == nullptr || loss_out != nullptr, "engine_forward: loss_out required with labels");
  SCOT_REQUIRE(((uintptr_t)arena & 255) == 0 && ((uintptr_t)params & 15) == 0, "engine_forward: arena must be 256 B aligned");
  Ctx c{e, params, nullptr, (uint8_t*)arena, time, (cudaStream_t)stream, gemm_impl};
  SplitGuard split_guard(e->split_off);
  const int B = e->batch, ns = e->ns;
  const Geo& g0 = e->geo[0];
  RC(cnx_side_init(e));
  // bf16 copy of all parameters (GEMM operands)
  RC(scot_cast_f32_bf16_launch(params, c.at<bf16>(e->wb16), e->n_elems, c.st));
  // relative-position-bias tables of every attention layer (batch independent), one launch
  for (const ScotCpbTable& t : e->cpb_early) RC(scot_cpb_fwd_launch(&t, params, arena, c.st));
  for (const ScotCpbTable& t : e->cpb_late) RC(scot_cpb_fwd_launch(&t, params, arena, c.st));
  // ---- embeddings (scOT/model.py:295-310, 345-366) ----
  const int K0 = d.num_channels * d.patch_size * d.patch_size;
  RC(scot_im2col_patch_launch(pixel_values, c.at<bf16>(e->p16), B, d.num_channels, d.image_size, d.image_size, d.patch_size,
                              c.st));
  RC(gemm(c, c.at<bf16>(e->p16), K0, 0, c.w16(e->emb_w), K0, 0, g0.M, g0.C, K0, SCOT_EPI_F32, c.p(e->emb_b),
          c.at<float>(e->z32), g0.C));
  RC(norm_fwd(c, e->emb_norm, c.at<float>(e->z32), nullptr, c.at<float>(e->emb_x), c.at<bf16>(e->emb_xb),
              c.at<bf16>(e->emb_zhat), c.at<float>(e->emb_rstd), g0.M, g0.C, g0.res * g0.res, 0, 1e-5f));
  // ---- encoder (model.py:816-861) ----
  const float* x = c.at<float>(e->emb_x);
  const bf16* xb = c.at<bf16>(e->emb_xb);
  std::vector<const float*> skip(ns);
  std::vector<const bf16*> skipb(ns);
  for (int s = 0; s < ns; ++s) {
    const Geo& g = e->geo[s];
    const float* stage_in = x;
    for (int i = 0; i < g.depth; ++i) {
      RC(block_fwd(c, e->enc[s][i], e->ebuf[s][i], g, shift_of(g, i), x, xb));
      x = c.at<float>(e->ebuf[s][i].xout);
      xb = c.at<bf16>(e->ebuf[s][i].xbout);
    }
    skip[s] = x;
    skipb[s] = xb;
    if (cnx_on_side(e, s)) {
      // side branch: needed again only when the decoder comes back up to this stage
      Ctx cs = c;
      cs.st = e->side2;
      RC(cnx_fork(c, s));
      RC(cnx_fwd_stage(cs, s, &skip[s], cnx_scratch(c, true)));
      RC(cnx_mark_join(c, s));
    }
    if (s < ns - 1) {
      const Geo& gn = e->geo[s + 1];
      RC(scot_merge_gather_launch(x, stage_in, c.at<bf16>(e->m_g16[s]), B, g.res, g.C, c.st));
      RC(gemm(c, c.at<bf16>(e->m_g16[s]), 4L * g.C, 0, c.w16(e->merge[s].wred), 4L * g.C, 0, gn.M, gn.C, 4L * g.C,
              SCOT_EPI_F32, nullptr, c.at<float>(e->z32), gn.C));
      RC(norm_fwd(c, e->merge[s].norm, c.at<float>(e->z32), nullptr, c.at<float>(e->m_x[s]), c.at<bf16>(e->m_xb[s]),
                  c.at<bf16>(e->m_norm[s].zhat), c.at<float>(e->m_norm[s].rstd), gn.M, gn.C, gn.res * gn.res, 0, 1e-5f));
      x = c.at<float>(e->m_x[s]);
      xb = c.at<bf16>(e->m_xb[s]);
    }
  }
  // ---- ConvNeXt blocks on the skips (model.py:198-217, 1388-1393) ----
  for (int s = 0; s < ns; ++s) {
    const Geo& g = e->geo[s];
    if (d.skip_blocks[s] > 0 && !cnx_on_side(e, s)) RC(cnx_fwd_stage(c, s, &skip[s], cnx_scratch(c, false)));
    if (d.skip_blocks[s] > 0 && s == ns - 1) {
      RC(scot_cast_f32_bf16_launch(skip[s], c.at<bf16>(e->skip_xb[s]), g.M * g.C, c.st));
      skipb[s] = c.at<bf16>(e->skip_xb[s]);
    }
  }
  // ---- decoder (model.py:916-961, 1145-1240) ----
  x = skip[ns - 1];
  xb = skipb[ns - 1];
  for (int j = 0; j < ns; ++j) {
    const int s = ns - 1 - j;
    const Geo& g = e->geo[s];
    for (int bi = 0; bi < g.depth; ++bi) {
      const int orig = g.depth - 1 - bi;  // blocks are stored in reversed construction order (model.py:900)
      RC(block_fwd(c, e->dec[j][bi], e->dbuf[j][bi], g, shift_of(g, orig), x, xb));
      x = c.at<float>(e->dbuf[j][bi].xout);
      xb = c.at<bf16>(e->dbuf[j][bi].xbout);
    }
    if (s > 0) {
      const Geo& gf = e->geo[s - 1];
      const UnmergeP& p = e->unmerge[j];
      RC(gemm(c, xb, g.C, 0, c.w16(p.wup), g.C, 0, g.M, 2L * g.C, g.C, SCOT_EPI_F32, nullptr, c.at<float>(e->z32),
              2L * g.C));
      RC(norm_fwd(c, p.norm, c.at<float>(e->z32), nullptr, nullptr, c.at<bf16>(e->u_nb[j]), c.at<bf16>(e->u_norm[j].zhat),
                  c.at<float>(e->u_norm[j].rstd), gf.M, gf.C, gf.res * gf.res, g.res, 1e-5f));
      RC(cnx_join(c, s - 1));  // the post-ConvNeXt skip of the finer stage is the residual of this GEMM
      RC(gemm(c, c.at<bf16>(e->u_nb[j]), gf.C, 0, c.w16(p.wmix), gf.C, 0, gf.M, gf.C, gf.C, SCOT_EPI_ADD_F32_BF16, nullptr,
              c.at<float>(e->u_x[j]), gf.C, c.at<bf16>(e->u_xb[j]), gf.C, skip[s - 1], gf.C));
      x = c.at<float>(e->u_x[j]);
      xb = c.at<bf16>(e->u_xb[j]);
    }
  }
  RC(cnx_join_all(c));
  // ---- patch recovery (model.py:639-647) ----
  const int NR = d.num_out_channels * d.patch_size * d.patch_size;
  RC(scot_expand_bias_launch(c.p(e->rec_b), c.at<float>(e->rec_bias), NR, d.patch_size * d.patch_size, c.st));
  RC(gemm(c, xb, g0.C, 0, c.w16(e->rec_w), NR, 1, g0.M, NR, g0.C, SCOT_EPI_F32, c.at<float>(e->rec_bias),
          c.at<float>(e->rec_D), NR));
  const float* resid = d.learn_residual ? pixel_values : nullptr;
  RC(scot_unshuffle_launch(c.at<float>(e->rec_D), c.at<float>(e->rec_P), B, d.num_out_channels, d.image_size, d.image_size,
                           d.patch_size, c.st));
  RC(scot_conv5_fwd_launch(c.at<float>(e->rec_P), c.p(e->rec_mix), resid, d.num_channels, labels, mask,
                           labels ? mask_mode : 0, pred_out, B, d.num_out_channels, d.image_size, d.image_size, c.st));
  if (labels != nullptr) {
    RC(scot_loss_fwd_launch(pred_out, labels, c.at<float>(e->loss_sums), loss_out, d.n_slices >= 2 ? d.slices : nullptr,
                            d.n_slices, d.loss_p, B, d.num_out_channels, (long)d.image_size * d.image_size, c.st));
  }
  e->last_pixels = pixel_values;
  e->last_time = time;
  e->last_labels = labels;
  e->last_mask = mask;
  e->last_mask_mode = labels ? mask_mode : 0;
  e->last_pred = pred_out;
  e->have_forward = true;
  return 0;
}

int scot_engine_bind_io(ScotEngine* e, const float* pixel_values, const float* time, const float* labels,
                        const uint8_t* mask, int mask_mode, float* pred) {
  SCOT_REQUIRE(e && pixel_values && pred, "engine_bind_io: null pointer");
  e->last_pixels = pixel_values;
  e->last_time = time;
  e->last_labels = labels;
  e->last_mask = mask;
  e->last_mask_mode = labels ? mask_mode : 0;
  e->last_pred = pred;
  e->have_forward = true;
  return 0;
}

int scot_engine_backward(ScotEngine* e, const float* params, float* grads, void* arena, const float* grad_loss,
                         const float* grad_pred, int gemm_impl, void* stream) {
  return scot_engine_backward_part(e, params, grads, arena, grad_loss, grad_pred, gemm_impl, 0, stream);
}

long scot_engine_grad_split(const ScotEngine* e) { return e ? e->split_elem : 0; }

// part 0: the whole backward pass. part 1: loss, recovery, decoder, ConvNeXt skips and the deepest encoder stage — when it
// returns (all side streams joined) the gradients of flat elements [scot_engine_grad_split(), end) are final; part 2: the
// remaining encoder stages and the embeddings, elements [0, split). Data parallel training all-reduces the first range on
// a communication stream while part 2 computes (runtime.GraphedTrainStep).
int scot_engine_backward_part(ScotEngine* e, const float* params, float* grads, void* arena, const float* grad_loss,
                              const float* grad_pred, int gemm_impl, int part, void* stream) {
  SCOT_REQUIRE(e && params && grads && arena, "engine_backward: null pointer");
  SCOT_REQUIRE(part >= 0 && part <= 2, "engine_backward: part must be 0, 1 or 2");
  const bool do1 = part != 2, do2 = part != 1;
  SCOT_REQUIRE(e->have_forward, "engine_backward: call scot_engine_forward first");
  SCOT_REQUIRE(!do1 || grad_loss != nullptr || grad_pred != nullptr, "engine_backward: need grad_loss and/or grad_pred");
  SCOT_REQUIRE(!do1 || grad_loss == nullptr || e->last_labels != nullptr, "engine_backward: grad_loss given but forward had no labels");
  const ScotModelDesc& d = e->d;
  Ctx c{e, params, grads, (uint8_t*)arena, e->last_time, (cudaStream_t)stream, gemm_impl};
  SplitGuard split_guard(e->split_off);
  const int B = e->batch, ns = e->ns;
  const Geo& g0 = e->geo[0];
  const int NR = d.num_out_channels * d.patch_size * d.patch_size;
  const long HW = (long)d.image_size * d.image_size;
  if (e->overlap < 0) {
    const char* ev = getenv("SCOT_WGRAD_OVERLAP");
    e->overlap = (ev != nullptr && ev[0] == '0') ? 0 : 1;
  }
  if (e->overlap && e->side == nullptr) {
    SCOT_CHECK_CUDA(cudaStreamCreateWithFlags(&e->side, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) {
      SCOT_CHECK_CUDA(cudaEventCreateWithFlags(&e->ev_fork[k], cudaEventDisableTiming));
      SCOT_CHECK_CUDA(cudaEventCreateWithFlags(&e->ev_join[k], cudaEventDisableTiming));
    }
  }
  e->join_pending[0] = e->join_pending[1] = false;
  RC(cnx_side_init(e));
  if (e->attn_split_ws < 0) {
    const char* ev = getenv("SCOT_ATTN_BWD_SPLIT");  // default: every window size; 0 = both kernels on the main stream
    e->attn_split_ws = ev != nullptr ? atoi(ev) : 16;
    if (e->attn_split_ws < 0) e->attn_split_ws = 0;
  }
  if (e->attn_split_ws > 0 && e->side3 == nullptr) {
    SCOT_CHECK_CUDA(cudaStreamCreateWithFlags(&e->side3, cudaStreamNonBlocking));
    SCOT_CHECK_CUDA(cudaEventCreateWithFlags(&e->ev_afork, cudaEventDisableTiming));
    SCOT_CHECK_CUDA(cudaEventCreateWithFlags(&e->ev_ajoin, cudaEventDisableTiming));
  }
  if (do1) {
  SCOT_CHECK_CUDA(cudaMemsetAsync(c.at<uint8_t>(e->dgrads_zero_begin), 0, e->dgrads_zero_bytes, c.st));
  // ---- loss + patch recovery backward ----
  float* dpred = c.at<float>(e->dpred);
  RC(scot_loss_bwd_launch(e->last_pred, e->last_labels, c.at<float>(e->loss_sums), grad_loss, grad_pred, e->last_mask,
                          e->last_mask_mode, dpred, d.n_slices >= 2 ? d.slices : nullptr, d.n_slices, d.loss_p, B,
                          d.num_out_channels, HW, c.st));
  // rec_D (token-major deconv output) is dead after the forward un-shuffle: reuse it as the planar dP scratch
  RC(scot_conv5_bwd_launch(c.at<float>(e->rec_P), c.p(e->rec_mix), dpred, c.at<float>(e->rec_D), c.at<bf16>(e->dD16),
                           c.g(e->rec_mix), c.g(e->rec_b), B, d.num_out_channels, d.image_size, d.image_size, d.patch_size,
                           c.st));
  // final decoder output (bf16) = input of the recovery GEMM
  const BlockBuf& lastb = e->dbuf[ns - 1].back();
  const bf16* xb_final = c.at<bf16>(lastb.xbout);
  RC(gemm(c, xb_final, g0.C, 1, c.at<bf16>(e->dD16), NR, 1, g0.C, NR, g0.M, SCOT_EPI_ATOMIC_F32, nullptr, c.g(e->rec_w), NR));
  RC(gemm(c, c.at<bf16>(e->dD16), NR, 0, c.w16(e->rec_w), NR, 0, g0.M, g0.C, NR, SCOT_EPI_F32, nullptr,
          c.at<float>(e->gstage[0]), g0.C));
  // ---- decoder backward (finest stage first) ----
  for (int j = ns - 1; j >= 0; --j) {
    const int s = ns - 1 - j;
    const Geo& g = e->geo[s];
    float* gr = c.at<float>(e->gstage[s]);
    for (int bi = g.depth - 1; bi >= 0; --bi) {
      const int orig = g.depth - 1 - bi;
      const bf16* xb_in;
      if (bi > 0) xb_in = c.at<bf16>(e->dbuf[j][bi - 1].xbout);
      else if (j == 0) xb_in = (d.skip_blocks[s] > 0) ? c.at<bf16>(e->skip_xb[s]) : c.at<bf16>(e->ebuf[s].back().xbout);
      else xb_in = c.at<bf16>(e->u_xb[j - 1]);
      RC(block_bwd(c, e->dec[j][bi], e->dbuf[j][bi], g, shift_of(g, orig), gr, xb_in, bi & 1));
    }
    RC(join_all(c));  // the code below reuses the scratch buffers on the main stream
    if (j > 0) {
      // gr = grad wrt (mixup(norm(shuffle(upsample(x_coarse)))) + skip[s]); the skip part stays in gstage[s]
      const int jc = j - 1;  // decoder layer that produced this stage's input
      const Geo& gc = e->geo[s + 1];
      const UnmergeP& p = e->unmerge[jc];
      bf16* gb = c.at<bf16>(e->dzb);
      RC(scot_cast_f32_bf16_launch(gr, gb, g.M * g.C, c.st));
      RC(gemm(c, gb, g.C, 1, c.at<bf16>(e->u_nb[jc]), g.C, 1, g.C, g.C, g.M, SCOT_EPI_ATOMIC_F32, nullptr, c.g(p.wmix), g.C));
      RC(gemm(c, gb, g.C, 0, c.w16(p.wmix), g.C, 1, g.M, g.C, g.C, SCOT_EPI_F32, nullptr, c.at<float>(e->z32), g.C));
      bf16* du = c.at<bf16>(e->dh);  // [M_coarse, 2*C_coarse]
      RC(norm_bwd(c, p.norm, c.at<float>(e->z32), c.at<bf16>(e->u_norm[jc].zhat), c.at<float>(e->u_norm[jc].rstd), du, 0,
                  nullptr, g.M, g.C, g.res * g.res, gc.res));
      const bf16* xb_coarse = c.at<bf16>(e->dbuf[jc].back().xbout);
      RC(gemm(c, du, 2L * gc.C, 1, xb_coarse, gc.C, 1, 2L * gc.C, gc.C, gc.M, SCOT_EPI_ATOMIC_F32, nullptr, c.g(p.wup), gc.C));
      RC(gemm(c, du, 2L * gc.C, 0, c.w16(p.wup), gc.C, 1, gc.M, gc.C, 2L * gc.C, SCOT_EPI_F32, nullptr,
              c.at<float>(e->gstage[s + 1]), gc.C));
    }
    if (cnx_on_side(e, s)) {
      // gstage[s] now holds the gradient of the post-ConvNeXt skip and is not touched by the main stream until the
      // encoder backward reaches stage s: the ConvNeXt backward runs beside the deeper stages
      Ctx cs = c;
      cs.st = e->side2;
      RC(cnx_fork(c, s));
      RC(cnx_bwd_stage(cs, s, gr, cnx_scratch(c, true)));
      RC(cnx_mark_join(c, s));
    }
  }
  // ---- ConvNeXt backward on every skip (those that did not run as a side branch above) ----
  for (int s = 0; s < ns; ++s)
    if (d.skip_blocks[s] > 0 && !cnx_on_side(e, s))
      RC(cnx_bwd_stage(c, s, c.at<float>(e->gstage[s]), cnx_scratch(c, false)));
  }  // do1
  // ---- encoder backward (coarsest stage first) ----
  for (int s = ns - 1; s >= 0; --s) {
    if (!(s == ns - 1 ? do1 : do2)) continue;
    const Geo& g = e->geo[s];
    float* gr = c.at<float>(e->gstage[s]);
    RC(cnx_join(c, s));  // ConvNeXt backward of this stage's skip (side branch) has updated gstage[s]
    if (s < ns - 1) {
      // merge backward: gstage[s+1] = grad wrt merge output
      const Geo& gn = e->geo[s + 1];
      const MergeP& p = e->merge[s];
      bf16* dzb = c.at<bf16>(e->dzb);
      RC(norm_bwd(c, p.norm, c.at<float>(e->gstage[s + 1]), c.at<bf16>(e->m_norm[s].zhat), c.at<float>(e->m_norm[s].rstd),
                  dzb, 0, nullptr, gn.M, gn.C, gn.res * gn.res, 0));
      RC(gemm(c, dzb, gn.C, 1, c.at<bf16>(e->m_g16[s]), 4L * g.C, 1, gn.C, 4L * g.C, gn.M, SCOT_EPI_ATOMIC_F32, nullptr,
              c.g(p.wred), 4L * g.C));
      RC(gemm(c, dzb, gn.C, 0, c.w16(p.wred), 4L * g.C, 1, gn.M, 4L * g.C, gn.C, SCOT_EPI_F32, nullptr, c.at<float>(e->z32),
              4L * g.C));
      RC(scot_merge_scatter_launch(c.at<float>(e->z32), gr, gr, B, g.res, g.C, c.st));
    }
    for (int i = g.depth - 1; i >= 0; --i) {
      const bf16* xb_in;
      if (i > 0) xb_in = c.at<bf16>(e->ebuf[s][i - 1].xbout);
      else if (s == 0) xb_in = c.at<bf16>(e->emb_xb);
      else xb_in = c.at<bf16>(e->m_xb[s - 1]);
      RC(block_bwd(c, e->enc[s][i], e->ebuf[s][i], g, shift_of(g, i), gr, xb_in, i & 1));
    }
    RC(join_all(c));
    if (s < ns - 1) {
      // the stage input also feeds the merge through `hidden + inputs` (model.py:847-849)
      RC(scot_merge_scatter_launch(c.at<float>(e->z32), gr, gr, B, g.res, g.C, c.st));
    }
    if (s == ns - 1) {
      // end of part 1: every side branch joins here, then the bias-MLP / logit-scale gradients of the layers done so far
      RC(cnx_join_all(c));
      for (const ScotCpbTable& t : e->cpb_early) RC(scot_cpb_bwd_launch(&t, params, grads, arena, c.st));
    }
  }
  if (!do2) return 0;
  RC(cnx_join_all(c));
  // ---- embeddings backward ----
  {
    const int K0 = d.num_channels * d.patch_size * d.patch_size;
    bf16* dzb = c.at<bf16>(e->dzb);
    RC(norm_bwd(c, e->emb_norm, c.at<float>(e->gstage[0]), c.at<bf16>(e->emb_zhat), c.at<float>(e->emb_rstd), dzb, 0,
                c.g(e->emb_b), g0.M, g0.C, g0.res * g0.res, 0));
    RC(gemm(c, dzb, g0.C, 1, c.at<bf16>(e->p16), K0, 1, g0.C, K0, g0.M, SCOT_EPI_ATOMIC_F32, nullptr, c.g(e->emb_w), K0));
  }
  // ---- relative-position-bias MLPs + logit scales of all attention layers (batched) ----
  for (const ScotCpbTable& t : e->cpb_late) RC(scot_cpb_bwd_launch(&t, params, grads, arena, c.st));
  return 0;
}

}  // extern "C"

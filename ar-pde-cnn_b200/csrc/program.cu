// program.cu -- DenseED "program" runner: executes the whole encoder-decoder as a short list of fused
// conv ops (host-side C++ loop, no Python between launches) and the auto-regressive rollout on top.
//
// Replaces DenseED.forward (2D-Burgers-SWAG/nn/denseEDcirc2d.py:313-314, 1D-*/nn/denseEDcirc.py:306-307
// resp. :294-295), i.e. nn.Sequential over ~30 modules / ~180 ATen ops, and the Python AR loops of
// main.py test()/testSample() (2D :112-135,:160-179; 1D :106-121,:148-163).
#include <vector>
#include "common.cuh"
#include "conv_common.cuh"
#include "bn.cuh"
#include "../../include/arpde.h"

static int check_program(const arpde_op_t* ops, int n_ops, const arpde_buf_t* bufs, int n_bufs, int n_params) {
  ARPDE_CHECK_ARG(ops && n_ops > 0 && (bufs || n_bufs == 0), "densed: empty program");
  for (int i = 0; i < n_ops; ++i) {
    const arpde_op_t& o = ops[i];
    ARPDE_CHECK_ARG(o.in_buf >= -1 && o.in_buf < n_bufs, "densed: op %d bad in_buf %d", i, o.in_buf);
    ARPDE_CHECK_ARG(o.out_buf == -2 || (o.out_buf >= 0 && o.out_buf < n_bufs), "densed: op %d bad out_buf %d", i, o.out_buf);
    ARPDE_CHECK_ARG(o.w_param >= 0 && o.w_param < n_params, "densed: op %d bad w_param", i);
    ARPDE_CHECK_ARG(o.bn_param == -1 || (o.bn_param >= 0 && o.bn_param + 4 < n_params), "densed: op %d bad bn_param", i);
  }
  return ARPDE_OK;
}

struct RunCtx {
  const arpde_op_t* ops; int n_ops;
  const arpde_buf_t* bufs; int n_bufs;
  void* const* params; const int64_t* gstride; int n_params;
  int N, n_per_group, groups, H, W;
  float* work; float* ss; int bn_total; double* stats; long long stats_len = 0; float* saved;
  float momentum, eps;
  cudaStream_t stream;
  float* tc_scratch = nullptr; long long tc_scratch_floats = 0;
  std::vector<const float*> wprep;          // per op: prepared tensor-core weights (nullptr = fp32 FFMA kernel)
  std::vector<long long> wprep_gs;
  bool fuse_dense = true;                   // eval mode: dense blocks as one launch (off for the layer-wise cross-check path)
  std::vector<char> zpair;                  // per op: 1 = this op and the next run as an operand-format pair (conv_z.cu; rollouts only)
  bool pdl = false;                         // rollouts: programmatic dependent launch between the kernels of the fused path
  std::vector<char> hconv;                  // per op: 1 = eval-mode stride-1 layer on the fp16 two-term kernel (conv_h.cu; rollouts only)
  const DpHook* dp = nullptr;               // data-parallel training: all-reduce of the BatchNorm sums (nullptr = single process)
  mutable std::vector<int> synced;          // per buffer: channels whose statistics already hold the global sums
};

static void op_desc(const RunCtx& c, const arpde_op_t& o, ConvDesc* d) {
  d->dilate = 0; d->pad_override = -1;
  d->N = c.N; d->cin = o.cin; d->cout = o.cout; d->kh = o.kh; d->kw = o.kw; d->stride = o.stride; d->up = o.up;
  d->relu_in = o.bn_param >= 0; d->act = o.act; d->n_per_group = c.n_per_group;
  if (o.in_buf < 0) { d->H = c.H; d->W = c.W; }
  else { d->H = c.bufs[o.in_buf].h; d->W = c.bufs[o.in_buf].w; }
  d->y_ctot = o.cout; d->y_coff = 0;
  d->pdl = c.pdl ? 1 : 0;
}

// Split/arrange the weights of every layer that runs on the tensor cores (once per forward / per rollout).
static int prep_tc(RunCtx& c) {
  c.wprep.assign(c.n_ops, nullptr); c.wprep_gs.assign(c.n_ops, 0);
  if (!c.tc_scratch || !arpde_tc_enabled()) return ARPDE_OK;
  long long used = 0;
  for (int i = 0; i < c.n_ops; ++i) {
    const arpde_op_t& o = c.ops[i];
    ConvDesc d; op_desc(c, o, &d);
    if (!arpde_conv_tc_wanted(d)) continue;
    const long long per_group = arpde_conv_tc_wprep_floats(o.cin, o.cout, o.kh, o.kw);
    const int wgroups = c.gstride[o.w_param] ? c.groups : 1;
    if (used + per_group * wgroups > c.tc_scratch_floats) continue;      // scratch too small: stay on the FFMA kernel
    float* dst = c.tc_scratch + used;
    int rc = arpde_conv_tc_prep_launch(d, (const float*)c.params[o.w_param], c.gstride[o.w_param], wgroups, dst, per_group, c.stream);
    if (rc) return rc;
    c.wprep[i] = dst; c.wprep_gs[i] = wgroups > 1 ? per_group : 0;
    used += per_group * wgroups;
  }
  return ARPDE_OK;
}

// Rollouts (no backward pass can follow): a tensor-core convolution whose output feeds exactly one BatchNorm + ReLU + 3x3 stride-2
// convolution hands its activation over in the consumer's fp16 operand format (conv_z.cu) instead of fp32 NCHW -- In_conv1 -> In_conv3 of
// the 2D network.  The pair shares the producer's work buffer, which must have room for the slightly larger operand form
// (arpde_conv_z_floats; Python's layout reserves it).  Re-prepares the consumer's weights in the fp16 layout.
static int mark_zpairs(RunCtx& c) {
  c.zpair.assign(c.n_ops, 0);
  if (c.wprep.empty()) return ARPDE_OK;
  for (int i = 0; i + 1 < c.n_ops; ++i) {
    const arpde_op_t& p = c.ops[i]; const arpde_op_t& q = c.ops[i + 1];
    if (!c.wprep[i] || !c.wprep[i + 1] || p.out_buf < 0 || q.in_buf != p.out_buf || p.out_coff != 0 || p.act != 0 || p.up || p.stride != 1) continue;
    const arpde_buf_t& b = c.bufs[p.out_buf];
    if (b.ctot != p.cout || q.cin != p.cout || q.bn_param < 0 || b.h < 2) continue;
    bool sole = true;
    for (int j = 0; j < c.n_ops; ++j) if ((j != i && c.ops[j].out_buf == p.out_buf) || (j != i + 1 && c.ops[j].in_buf == p.out_buf)) sole = false;
    if (!sole) continue;
    ConvDesc dq; op_desc(c, q, &dq);
    if (!arpde_conv_z_wanted(dq)) continue;
    long long next = -1;
    for (int k = 0; k < c.n_bufs; ++k) if (c.bufs[k].offset > b.offset && (next < 0 || c.bufs[k].offset < next)) next = c.bufs[k].offset;
    const long long need = (long long)c.N * 4 * arpde_conv_z_sample16(p.cout, b.h, b.w);
    if (next < 0 || next - b.offset < need || (b.offset & 3) || (((uintptr_t)c.work) & 15)) continue;
    const int wgroups = c.gstride[q.w_param] ? c.groups : 1;
    const long long per_group = arpde_conv_tc_wprep_floats(q.cin, q.cout, q.kh, q.kw);
    int rc = arpde_conv_z_prep_launch(dq, (const float*)c.params[q.w_param], c.gstride[q.w_param], wgroups, (float*)c.wprep[i + 1], per_group, c.stream);
    if (rc) return rc;
    c.zpair[i] = 1;
    ConvDesc dp; op_desc(c, p, &dp);
    if (p.bn_param < 0 && arpde_conv_stem_wanted(dp)) {        // the network's first layer: dedicated fp16 producer (conv_stem.cu)
      // the consumer's folded BatchNorm (fold_eval ran before) goes into the producer's prepared weights -- which needs as many weight
      // sets as there are BatchNorm sets (SWAG samples both per group; shared weights under per-group BatchNorm keep the generic producer)
      const int wsets = c.gstride[p.w_param] ? c.groups : 1;
      const int bnsets = (c.gstride[q.bn_param] || c.gstride[q.bn_param + 1]) ? c.groups : 1;
      if (bnsets > wsets) { ++i; continue; }
      ZOut zo; zo.z = nullptr; zo.scale = c.ss + q.bn_off; zo.shift = c.ss + (long long)c.groups * c.bn_total + q.bn_off; zo.ss_gs = c.bn_total;
      rc = arpde_conv_stem_prep_launch(dp, (const float*)c.params[p.w_param], c.gstride[p.w_param], wsets, (float*)c.wprep[i],
                                       arpde_conv_tc_wprep_floats(p.cin, p.cout, p.kh, p.kw), zo, c.stream);
      if (rc) return rc;
      c.zpair[i] = 2;
    }
    ++i;
  }
  return ARPDE_OK;
}

// Rollouts: the remaining stride-1 tensor-core layers (conv1x1 of a transition, the 3x3 convolution behind the nearest x2 upsampling) take
// the fp16 two-term kernel of conv_h.cu; their weights are re-prepared in its layout (in place of the TF32 image).
static int mark_hconvs(RunCtx& c) {
  c.hconv.assign(c.n_ops, 0);
  if (c.wprep.empty()) return ARPDE_OK;
  for (int i = 0; i < c.n_ops; ++i) {
    if (!c.wprep[i] || (!c.zpair.empty() && (c.zpair[i] || (i > 0 && c.zpair[i - 1])))) continue;
    const arpde_op_t& o = c.ops[i];
    ConvDesc d; op_desc(c, o, &d);
    if (!arpde_conv_h_wanted(d)) continue;
    const long long per_group = arpde_conv_tc_wprep_floats(o.cin, o.cout, o.kh, o.kw);
    if (arpde_conv_h_wprep_floats(o.cin, o.cout, o.kh, o.up) > per_group) continue;
    const int wgroups = c.gstride[o.w_param] ? c.groups : 1;
    int rc = arpde_conv_h_prep_launch(d, (const float*)c.params[o.w_param], c.gstride[o.w_param], wgroups, (float*)c.wprep[i], per_group, c.stream);
    if (rc) return rc;
    c.hconv[i] = 1;
  }
  return ARPDE_OK;
}

static int fold_eval(const RunCtx& c) {
  BnTable t; t.n = 0;
  for (int i = 0; i < c.n_ops; ++i) {
    const arpde_op_t& o = c.ops[i];
    if (o.bn_param < 0) continue;
    BnDesc& d = t.d[t.n++];
    d.gamma = (const float*)c.params[o.bn_param]; d.beta = (const float*)c.params[o.bn_param + 1];
    d.rmean = (const float*)c.params[o.bn_param + 2]; d.rvar = (const float*)c.params[o.bn_param + 3];
    d.g_gs = c.gstride[o.bn_param]; d.b_gs = c.gstride[o.bn_param + 1]; d.C = o.cin; d.off = o.bn_off;
    if (t.n == BN_TABLE_MAX) {
      int rc = arpde_bn_fold_eval_launch(t, c.groups, c.ss, c.ss + (long long)c.groups * c.bn_total, c.bn_total, c.eps, c.stream);
      if (rc) return rc; t.n = 0;
    }
  }
  return arpde_bn_fold_eval_launch(t, c.groups, c.ss, c.ss + (long long)c.groups * c.bn_total, c.bn_total, c.eps, c.stream);
}

// one pass over the ops.  x: network input (batch stride x_bs), y: output tensor with y_ctot channels, written at y_coff
static int run_ops(const RunCtx& c, const float* x, long long x_bs, float* y, int y_ctot, int y_coff, int training) {
  for (int i = 0; i < c.n_ops; ++i) {
    const arpde_op_t& o = c.ops[i];
    if (!training && c.fuse_dense) {
      // eval mode: a run of 4-channel dense layers on a 32-column map is ONE launch (dense_block.cu)
      const int nl = arpde_dense_block_match(c.ops, c.n_ops, i, c.bufs, c.N);
      if (nl > 0) {
        int rc = arpde_dense_block_launch(c.ops, i, nl, c.bufs, c.params, c.gstride, c.work, c.ss, c.ss + (long long)c.groups * c.bn_total, c.bn_total,
                                          c.N, c.n_per_group, c.stream, c.pdl);
        if (rc) return rc;
        i += nl - 1;
        continue;
      }
    }
    ConvDesc d;
    d.dilate = 0; d.pad_override = -1;
    d.N = c.N; d.cin = o.cin; d.cout = o.cout; d.kh = o.kh; d.kw = o.kw; d.stride = o.stride; d.up = o.up;
    d.relu_in = o.bn_param >= 0; d.act = o.act; d.n_per_group = c.n_per_group; d.pdl = (!training && c.pdl) ? 1 : 0;
    const float* in; long long in_bs;
    if (o.in_buf < 0) { in = x; in_bs = x_bs; d.H = c.H; d.W = c.W; }
    else { const arpde_buf_t& b = c.bufs[o.in_buf]; in = c.work + b.offset; in_bs = (long long)b.ctot * b.h * b.w; d.H = b.h; d.W = b.w; }
    float* out; double* st = nullptr;
    if (o.out_buf == -2) { out = y; d.y_ctot = y_ctot; d.y_coff = y_coff; }
    else {
      const arpde_buf_t& b = c.bufs[o.out_buf]; out = c.work + b.offset; d.y_ctot = b.ctot; d.y_coff = o.out_coff;
      if (training) st = c.stats + b.stats_off;
    }
    const float *scale = nullptr, *shift = nullptr;
    if (o.bn_param >= 0) {
      scale = c.ss + o.bn_off; shift = c.ss + (long long)c.groups * c.bn_total + o.bn_off;
      if (training) {
        // statistics of the BN input: accumulated by the producers of the buffer, or -- BatchNorm directly on the network input
        // (a stand-alone _Transition / LastTransUp, DenseED.forward_test) -- by a streaming pass into the tail slot of `stats`
        const double* st_in; double count;
        const bool sync = c.dp && c.dp->allreduce && c.dp->world > 1;
        if (o.in_buf >= 0) {
          const arpde_buf_t& b = c.bufs[o.in_buf]; st_in = c.stats + b.stats_off; count = (double)c.N * b.h * b.w;
          if (sync && c.synced[o.in_buf] < o.cin) {
            // global-batch statistics: sum the (sum x, sum x^2) pairs of the channels new to this BN layer over the ranks (a dense
            // block's buffer grows by `growth` channels per layer; the earlier ones were reduced for the previous layer)
            const int s0 = c.synced[o.in_buf];
            if (c.dp->allreduce(c.dp->user, c.stats + b.stats_off + 2 * s0, 2ll * (o.cin - s0))) {
              arpde_set_error("densed_forward: the all-reduce callback failed"); return ARPDE_ERR_CUDA;
            }
            c.synced[o.in_buf] = o.cin;
          }
        } else {
          double* slot = c.stats + c.stats_len - 2 * o.cin;
          int rc = arpde_bn_stats(in, in_bs, c.N, o.cin, (long long)d.H * d.W, slot, c.stream); if (rc) return rc;
          if (sync && c.dp->allreduce(c.dp->user, slot, 2ll * o.cin)) { arpde_set_error("densed_forward: the all-reduce callback failed"); return ARPDE_ERR_CUDA; }
          st_in = slot; count = (double)c.N * d.H * d.W;
        }
        if (sync) count *= c.dp->world;
        float* sm = c.saved ? c.saved + o.bn_off : nullptr;
        int rc = arpde_bn_finalize_train_launch(st_in, count, (const float*)c.params[o.bn_param],
                                                (const float*)c.params[o.bn_param + 1], (float*)c.params[o.bn_param + 2],
                                                (float*)c.params[o.bn_param + 3], (long long*)c.params[o.bn_param + 4], c.momentum,
                                                c.eps, (float*)scale, (float*)shift, sm, sm ? sm + c.bn_total : nullptr, o.cin, c.stream);
        if (rc) return rc;
      }
    }
    int rc;
    if (!training && !c.zpair.empty() && c.zpair[i]) {
      const arpde_op_t& q = c.ops[i + 1];
      ZOut zo; zo.z = out; zo.scale = c.ss + q.bn_off; zo.shift = c.ss + (long long)c.groups * c.bn_total + q.bn_off; zo.ss_gs = c.bn_total;
      rc = c.zpair[i] == 2 ? arpde_conv_stem_launch(d, in, in_bs, c.wprep[i], c.wprep_gs[i], zo, c.stream)
                           : arpde_conv_tc_launch(d, in, in_bs, c.wprep[i], c.wprep_gs[i], scale, shift, c.bn_total, nullptr, nullptr, c.stream, &zo);
      if (rc) return rc;
      ConvDesc dq; op_desc(c, q, &dq);
      float* out_q;
      if (q.out_buf == -2) { out_q = y; dq.y_ctot = y_ctot; dq.y_coff = y_coff; }
      else { const arpde_buf_t& bq = c.bufs[q.out_buf]; out_q = c.work + bq.offset; dq.y_ctot = bq.ctot; dq.y_coff = q.out_coff; }
      rc = arpde_conv_z_launch(dq, out, c.wprep[i + 1], c.wprep_gs[i + 1], out_q, c.stream);
      if (rc) return rc;
      ++i;
      continue;
    }
    if (!training && !c.hconv.empty() && c.hconv[i]) {
      rc = arpde_conv_h_launch(d, in, in_bs, c.wprep[i], c.wprep_gs[i], scale, shift, c.bn_total, out, c.stream);
    } else if (!c.wprep.empty() && c.wprep[i]) {
      // train mode: the batch statistics of the written channels come from a separate streaming pass over the (L2-resident) output
      // instead of the tensor-core epilogue's per-column warp reductions (10 shuffles + 2 double atomics per column and tile made
      // the epilogue warps the critical path: 2.2 ms vs 0.6 ms per AR step at B = 128)
      rc = arpde_conv_tc_launch(d, in, in_bs, c.wprep[i], c.wprep_gs[i], scale, shift, c.bn_total, out, nullptr, c.stream);
      if (!rc && st) {
        const int one_d = (d.H == 1 && d.kh == 1);
        const long long Ho = one_d ? 1 : ((long long)(d.H << (d.up ? 1 : 0)) / d.stride), Wo = ((long long)(d.W << (d.up ? 1 : 0)) / d.stride);
        rc = arpde_bn_stats(out + (long long)d.y_coff * Ho * Wo, (long long)d.y_ctot * Ho * Wo, d.N, d.cout, Ho * Wo, st + 2 * d.y_coff, c.stream);
      }
    } else if (st && arpde_conv_fwd_rt_shape_ok(d)) {
      // train mode, narrow layer: the row-tiled kernel has no statistics epilogue -- same separate streaming pass as above
      const long long Ho = d.H, Wo = d.W;
      rc = arpde_conv_fwd_launch(d, in, in_bs, (const float*)c.params[o.w_param], c.gstride[o.w_param], scale, shift, c.bn_total, out, nullptr, c.stream);
      if (!rc) rc = arpde_bn_stats(out + (long long)d.y_coff * Ho * Wo, (long long)d.y_ctot * Ho * Wo, d.N, d.cout, Ho * Wo, st + 2 * d.y_coff, c.stream);
    } else
      rc = arpde_conv_fwd_launch(d, in, in_bs, (const float*)c.params[o.w_param], c.gstride[o.w_param], scale, shift,
                                 c.bn_total, out, st, c.stream);
    if (rc) return rc;
  }
  return ARPDE_OK;
}

static int make_ctx(RunCtx* c, const arpde_op_t* ops, int n_ops, const arpde_buf_t* bufs, int n_bufs, void* const* params,
                    const int64_t* gstride, int n_params, int N, int n_per_group, int H, int W, float* work, float* ss,
                    int bn_total, double* stats, float* saved, double momentum, double eps, cudaStream_t stream) {
  int rc = check_program(ops, n_ops, bufs, n_bufs, n_params); if (rc) return rc;
  ARPDE_CHECK_ARG(params && gstride, "densed: null parameter table");
  ARPDE_CHECK_ARG(N > 0 && H > 0 && W > 0, "densed: bad dims N=%d H=%d W=%d", N, H, W);
  if (n_per_group <= 0) n_per_group = N;
  ARPDE_CHECK_ARG(N % n_per_group == 0, "densed: N=%d not a multiple of n_per_group=%d", N, n_per_group);
  c->ops = ops; c->n_ops = n_ops; c->bufs = bufs; c->n_bufs = n_bufs; c->params = params; c->gstride = gstride; c->n_params = n_params;
  c->N = N; c->n_per_group = n_per_group; c->groups = N / n_per_group; c->H = H; c->W = W;
  c->work = work; c->ss = ss; c->bn_total = bn_total; c->stats = stats; c->saved = saved;
  c->momentum = (float)momentum; c->eps = (float)eps; c->stream = stream;
  return ARPDE_OK;
}

// floats of tensor-core weight scratch arpde_densed_forward / arpde_rollout can use for this program (upper bound)
extern "C" int64_t arpde_densed_tc_scratch_floats(const arpde_op_t* ops, int n_ops, int groups) {
  if (!ops || n_ops <= 0) return 0;
  long long tot = 0;
  for (int i = 0; i < n_ops; ++i) tot += arpde_conv_tc_wprep_floats(ops[i].cin, ops[i].cout, ops[i].kh, ops[i].kw) * (groups > 0 ? groups : 1);
  return tot;
}

static int densed_forward_impl(const arpde_op_t* ops, int n_ops, const arpde_buf_t* bufs, int n_bufs, void* const* params_host,
                               const int64_t* param_gstride_host, int n_params, const float* x, int64_t x_bstride, int N,
                               int n_per_group, int H, int W, float* y, int y_ctot, int y_coff, float* workspace,
                               float* scale_shift, int bn_total, double* stats, int64_t stats_len, float* saved_mean_invstd,
                               int training, double momentum, double eps, float* tc_scratch, int64_t tc_scratch_floats,
                               arpde_stream_t stream_, const DpHook* dp) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RunCtx c;
  c.dp = dp; c.synced.assign(n_bufs > 0 ? n_bufs : 1, 0);
  int rc = make_ctx(&c, ops, n_ops, bufs, n_bufs, params_host, param_gstride_host, n_params, N, n_per_group, H, W, workspace,
                    scale_shift, bn_total, stats, saved_mean_invstd, momentum, eps, stream);
  if (rc) return rc;
  c.tc_scratch = tc_scratch; c.tc_scratch_floats = tc_scratch_floats;
  rc = prep_tc(c); if (rc) return rc;
  ARPDE_CHECK_ARG(x && y && (workspace || n_bufs == 0), "densed_forward: null pointer");
  ARPDE_CHECK_ARG(bn_total == 0 || scale_shift, "densed_forward: scale/shift scratch missing");
  if (training) {
    ARPDE_CHECK_ARG(c.groups == 1, "densed_forward: training with grouped weights is not defined");
    long long need = 0;
    for (int b = 0; b < n_bufs; ++b) need = need > bufs[b].stats_off + 2ll * bufs[b].ctot ? need : bufs[b].stats_off + 2ll * bufs[b].ctot;
    for (int i = 0; i < n_ops; ++i)
      if (ops[i].bn_param >= 0 && ops[i].in_buf < 0) { need += 2ll * ops[i].cin; break; }
    // a program without activation buffers and without BatchNorm (a lone convolution) needs no statistics
    ARPDE_CHECK_ARG(need == 0 || (stats && stats_len >= need), "densed_forward: training needs a stats scratch of %lld doubles (got %lld)", need,
                    (long long)(stats ? stats_len : 0));
    if (stats && stats_len > 0) ARPDE_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * stats_len, stream));
    c.stats_len = stats_len;
    // channels the caller pre-filled (the input copy of a stand-alone dense layer / block): no producer op accumulates their
    // batch statistics, so a streaming pass does
    for (int b = 0; b < n_bufs; ++b) {
      int first = bufs[b].ctot; bool read_bn = false;
      for (int i = 0; i < n_ops; ++i) {
        if (ops[i].out_buf == b && ops[i].out_coff < first) first = ops[i].out_coff;
        if (ops[i].in_buf == b && ops[i].bn_param >= 0) read_bn = true;
      }
      if (first > 0 && read_bn) {
        const long long hw = (long long)bufs[b].h * bufs[b].w;
        rc = arpde_bn_stats(workspace + bufs[b].offset, bufs[b].ctot * hw, N, first, hw, stats + bufs[b].stats_off, stream);
        if (rc) return rc;
      }
    }
  } else {
    rc = fold_eval(c); if (rc) return rc;
  }
  return run_ops(c, x, x_bstride, y, y_ctot, y_coff, training);
}

extern "C" int arpde_densed_forward(const arpde_op_t* ops, int n_ops, const arpde_buf_t* bufs, int n_bufs, void* const* params_host,
                                    const int64_t* param_gstride_host, int n_params, const float* x, int64_t x_bstride, int N,
                                    int n_per_group, int H, int W, float* y, int y_ctot, int y_coff, float* workspace,
                                    float* scale_shift, int bn_total, double* stats, int64_t stats_len, float* saved_mean_invstd,
                                    int training, double momentum, double eps, float* tc_scratch, int64_t tc_scratch_floats,
                                    arpde_stream_t stream) {
  return densed_forward_impl(ops, n_ops, bufs, n_bufs, params_host, param_gstride_host, n_params, x, x_bstride, N, n_per_group, H, W, y,
                             y_ctot, y_coff, workspace, scale_shift, bn_total, stats, stats_len, saved_mean_invstd, training, momentum,
                             eps, tc_scratch, tc_scratch_floats, stream, nullptr);
}

// Data-parallel training forward: train-mode BatchNorm over the GLOBAL mini-batch.  Between the kernels that accumulate a layer's
// (sum x, sum x^2) and the finalize kernel that turns them into scale / shift / running statistics, `allreduce` sums the pairs over
// the `dp_world` ranks (every rank holds N samples).  The callback is made from the calling host thread.
extern "C" int arpde_densed_forward_dp(const arpde_op_t* ops, int n_ops, const arpde_buf_t* bufs, int n_bufs, void* const* params_host,
                                       const int64_t* param_gstride_host, int n_params, const float* x, int64_t x_bstride, int N,
                                       int n_per_group, int H, int W, float* y, int y_ctot, int y_coff, float* workspace,
                                       float* scale_shift, int bn_total, double* stats, int64_t stats_len, float* saved_mean_invstd,
                                       int training, double momentum, double eps, float* tc_scratch, int64_t tc_scratch_floats,
                                       arpde_allreduce_fn allreduce, void* user, int dp_world, arpde_stream_t stream) {
  ARPDE_CHECK_ARG(dp_world >= 1 && (dp_world == 1 || allreduce), "densed_forward_dp: %d ranks need an all-reduce callback", dp_world);
  DpHook h; h.allreduce = (int (*)(void*, double*, long long))allreduce; h.user = user; h.world = dp_world;
  return densed_forward_impl(ops, n_ops, bufs, n_bufs, params_host, param_gstride_host, n_params, x, x_bstride, N, n_per_group, H, W, y,
                             y_ctot, y_coff, workspace, scale_shift, bn_total, stats, stats_len, saved_mean_invstd, training, momentum,
                             eps, tc_scratch, tc_scratch_floats, stream, &h);
}

// Auto-regressive rollout, eval mode.  traj: [N, traj_ctot, H, W] with traj_ctot = c_hist + T*c_out; the caller fills the first
// c_hist channels (the repeated initial condition); step t reads channels [t*c_out, t*c_out + c_in) and writes the prediction
// to channels [c_hist + t*c_out, +c_out): the sliding history window of the reference loops without any copy or cat.
static int rollout_impl(const arpde_op_t* ops, int n_ops, const arpde_buf_t* bufs, int n_bufs, void* const* params_host,
                        const int64_t* param_gstride_host, int n_params, float* traj, int traj_ctot, int c_in, int c_hist,
                        int c_out, int T, int N, int n_per_group, int H, int W, float* workspace, float* scale_shift,
                        int bn_total, double eps, float* tc_scratch, int64_t tc_scratch_floats, arpde_stream_t stream_, bool allow_persistent) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RunCtx c;
  int rc = make_ctx(&c, ops, n_ops, bufs, n_bufs, params_host, param_gstride_host, n_params, N, n_per_group, H, W, workspace,
                    scale_shift, bn_total, nullptr, nullptr, 0.1, eps, stream);
  if (rc) return rc;
  c.tc_scratch = tc_scratch; c.tc_scratch_floats = tc_scratch_floats;
  ARPDE_CHECK_ARG(traj && T >= 0 && c_in > 0 && c_out > 0 && c_hist >= c_in, "rollout: bad args");
  ARPDE_CHECK_ARG(traj_ctot >= c_hist + T * c_out, "rollout: trajectory has %d channels, needs %d", traj_ctot, c_hist + T * c_out);
  ARPDE_CHECK_ARG(ops[0].cin == c_in && ops[n_ops - 1].cout == c_out, "rollout: program does not match c_in/c_out");
  const long long HW = (long long)H * W, bs = (long long)traj_ctot * HW;
  // the reference's KS network on its 96-point grid: the whole rollout is ONE persistent launch (rollout1d.cu)
  if (allow_persistent && arpde_ks_program_matches(ops, n_ops, bufs, n_bufs, H, W, c_in, c_out))
    return arpde_ks_rollout_launch(ops, params_host, param_gstride_host, traj, bs, c_hist, T, N, c.n_per_group, (float)eps, stream);
  c.fuse_dense = allow_persistent; c.pdl = allow_persistent && !arpde_dev_env("ARPDE_NO_PDL");
  rc = prep_tc(c); if (rc) return rc;
  rc = fold_eval(c); if (rc) return rc;                     // before the pair preparation: the stem folds the BatchNorm affine into its weights
  if (allow_persistent) { rc = mark_zpairs(c); if (rc) return rc; rc = mark_hconvs(c); if (rc) return rc; }
  for (int t = 0; t < T; ++t) {
    const float* x = traj + (long long)(c_hist - c_in + t * c_out) * HW;
    rc = run_ops(c, x, bs, traj, traj_ctot, c_hist + t * c_out, 0);
    if (rc) return rc;
  }
  return ARPDE_OK;
}

extern "C" int arpde_rollout(const arpde_op_t* ops, int n_ops, const arpde_buf_t* bufs, int n_bufs, void* const* params_host,
                             const int64_t* param_gstride_host, int n_params, float* traj, int traj_ctot, int c_in, int c_hist,
                             int c_out, int T, int N, int n_per_group, int H, int W, float* workspace, float* scale_shift,
                             int bn_total, double eps, float* tc_scratch, int64_t tc_scratch_floats, arpde_stream_t stream) {
  return rollout_impl(ops, n_ops, bufs, n_bufs, params_host, param_gstride_host, n_params, traj, traj_ctot, c_in, c_hist, c_out, T, N,
                      n_per_group, H, W, workspace, scale_shift, bn_total, eps, tc_scratch, tc_scratch_floats, stream, true);
}

// same contract, always the layer-by-layer path (one launch per convolution and step): cross-check of the persistent kernels
extern "C" int arpde_rollout_layerwise(const arpde_op_t* ops, int n_ops, const arpde_buf_t* bufs, int n_bufs, void* const* params_host,
                                       const int64_t* param_gstride_host, int n_params, float* traj, int traj_ctot, int c_in, int c_hist,
                                       int c_out, int T, int N, int n_per_group, int H, int W, float* workspace, float* scale_shift,
                                       int bn_total, double eps, float* tc_scratch, int64_t tc_scratch_floats, arpde_stream_t stream) {
  return rollout_impl(ops, n_ops, bufs, n_bufs, params_host, param_gstride_host, n_params, traj, traj_ctot, c_in, c_hist, c_out, T, N,
                      n_per_group, H, W, workspace, scale_shift, bn_total, eps, tc_scratch, tc_scratch_floats, stream, false);
}

// Backward of arpde_densed_forward (train or eval mode): what loss.backward() does through DenseED in the reference
// (main.py:81).  Ops are visited in reverse; gradients of the activation buffers live in gwork (same layout as the
// workspace; the last consumer of a buffer overwrites, earlier consumers accumulate, so no memset is needed).
static int densed_backward_impl(const arpde_op_t* ops, int n_ops, const arpde_buf_t* bufs, int n_bufs, void* const* params_host,
                                int n_params, const float* x, int64_t x_bstride, int N, int H, int W, const float* gy,
                                const float* y, const float* workspace, const float* scale_shift, int bn_total,
                                const float* saved_mean_invstd, float* gwork, float* gx, void* const* grads_host, float* T1,
                                float* T2, float* wT, double* d64, int training, float* tc_scratch, int64_t tc_scratch_floats,
                                arpde_stream_t stream_, const DpHook* dp) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = check_program(ops, n_ops, bufs, n_bufs, n_params); if (rc) return rc;
  ARPDE_CHECK_ARG(params_host && grads_host && x && gy && T1 && T2 && wT && d64, "densed_backward: null pointer");
  ARPDE_CHECK_ARG((workspace && gwork) || n_bufs == 0, "densed_backward: workspace missing");
  ARPDE_CHECK_ARG(bn_total == 0 || (scale_shift && saved_mean_invstd), "densed_backward: BN scratch missing");
  std::vector<char> written(n_bufs > 0 ? n_bufs : 1, 0);
  for (int i = n_ops - 1; i >= 0; --i) {
    const arpde_op_t& o = ops[i];
    const float* in; long long in_bs; int Hin, Win;
    if (o.in_buf < 0) { in = x; in_bs = x_bstride; Hin = H; Win = W; }
    else { const arpde_buf_t& b = bufs[o.in_buf]; in = workspace + b.offset; in_bs = (long long)b.ctot * b.h * b.w; Hin = b.h; Win = b.w; }
    const int one_d = (Hin == 1 && o.kh == 1);
    const int up_w = o.up, up_h = one_d ? 0 : o.up;
    const int Hc = Hin << up_h, Wc = Win << up_w;
    const int Ho = one_d ? 1 : Hc / o.stride, Wo = Wc / o.stride;
    const float* dO; long long dO_bs;
    if (o.out_buf == -2) {
      dO = gy; dO_bs = (long long)o.cout * Ho * Wo;
      if (o.act != 0) {
        ARPDE_CHECK_ARG(y != nullptr, "densed_backward: output activation needs the forward output");
        rc = arpde_act_bwd_launch(gy, y, T2, (long long)N * o.cout * Ho * Wo, o.act, stream); if (rc) return rc;
        dO = T2;
      }
    } else {
      const arpde_buf_t& b = bufs[o.out_buf];
      dO = gwork + b.offset + (long long)o.out_coff * b.h * b.w; dO_bs = (long long)b.ctot * b.h * b.w;
    }
    const float *scale = nullptr, *shift = nullptr, *mean = nullptr, *invstd = nullptr;
    if (o.bn_param >= 0) {
      scale = scale_shift + o.bn_off; shift = scale_shift + bn_total + o.bn_off;
      mean = saved_mean_invstd + o.bn_off; invstd = saved_mean_invstd + bn_total + o.bn_off;
    }
    ConvDesc d;
    d.N = N; d.cin = o.cin; d.H = Hin; d.W = Win; d.cout = o.cout; d.kh = o.kh; d.kw = o.kw; d.stride = o.stride; d.up = o.up;
    d.relu_in = o.bn_param >= 0; d.act = 0; d.y_ctot = o.cout; d.y_coff = 0; d.n_per_group = N;
    float* gw = (float*)grads_host[o.w_param];
    if (gw) {
      // narrow layers: sliding-window kernel; wide layers: tensor-core kernel (conv_wgrad_tc.cu) with its partial-sum slabs in the
      // tensor-core scratch (free until the data gradient below prepares its weights there, later on the same stream); else generic
      rc = arpde_conv_wgrad_best_launch(d, in, in_bs, scale, shift, dO, dO_bs, d64, gw, tc_scratch, tc_scratch_floats, stream);
      if (rc) return rc;
    }
    const bool need_dx = o.in_buf >= 0 || gx != nullptr;
    if (!need_dx) continue;
    rc = arpde_weight_flip_transpose_launch((const float*)params_host[o.w_param], wT, o.cout, o.cin, o.kh, o.kw, stream); if (rc) return rc;
    ConvDesc g;
    g.N = N; g.cin = o.cout; g.H = Ho; g.W = Wo; g.cout = o.cin; g.kh = o.kh; g.kw = o.kw; g.stride = 1; g.up = 0;
    g.dilate = o.stride == 2; g.relu_in = 0; g.act = 0; g.n_per_group = N;
    const int p = o.kw == 4 ? 1 : (o.kw - 1) / 2;
    g.pad_override = o.kw - 1 - p;
    float* dest; long long dest_bs; bool acc = false;
    if (o.in_buf < 0) { dest = gx; dest_bs = (long long)o.cin * Hin * Win; }
    else { const arpde_buf_t& b = bufs[o.in_buf]; dest = gwork + b.offset; dest_bs = (long long)b.ctot * b.h * b.w; acc = written[o.in_buf]; }
    // the data-gradient conv of a stride-1 layer is an ordinary circular conv (pad k-1-p == p for odd k): tensor-core kernel
    if (g.pad_override == p) g.pad_override = -1;           // also for the zero-inserted input of a stride-2 layer
    const long long tc_need = arpde_conv_tc_wprep_floats(g.cin, g.cout, g.kh, g.kw);
    const bool use_tc = tc_scratch && tc_need <= tc_scratch_floats && arpde_conv_tc_wanted(g);
    if (use_tc) { rc = arpde_conv_tc_prep_launch(g, wT, 0, 1, tc_scratch, 0, stream); if (rc) return rc; }
    if (o.bn_param < 0 && !o.up && !acc) {
      g.y_ctot = (int)(dest_bs / ((long long)Hin * Win)); g.y_coff = 0;
      rc = use_tc ? arpde_conv_tc_launch(g, dO, dO_bs, tc_scratch, 0, nullptr, nullptr, 0, dest, nullptr, stream)
                  : arpde_conv_fwd_launch(g, dO, dO_bs, wT, 0, nullptr, nullptr, 0, dest, nullptr, stream);
      if (rc) return rc;
    } else {
      g.y_ctot = o.cin; g.y_coff = 0;
      rc = use_tc ? arpde_conv_tc_launch(g, dO, dO_bs, tc_scratch, 0, nullptr, nullptr, 0, T1, nullptr, stream)
                  : arpde_conv_fwd_launch(g, dO, dO_bs, wT, 0, nullptr, nullptr, 0, T1, nullptr, stream);
      if (rc) return rc;
      float* dgamma = o.bn_param >= 0 ? (float*)grads_host[o.bn_param] : nullptr;
      float* dbeta = o.bn_param >= 0 ? (float*)grads_host[o.bn_param + 1] : nullptr;
      rc = arpde_bn_relu_bwd_launch(T1, in, in_bs, scale, shift, mean, invstd, T2, d64, dest, dest_bs, dgamma, dbeta, N, o.cin, Hin, Win,
                                    up_h, up_w, training, acc ? 1 : 0, stream, dp);
      if (rc) return rc;
    }
    if (o.in_buf >= 0) written[o.in_buf] = 1;
  }
  return ARPDE_OK;
}

extern "C" int arpde_densed_backward(const arpde_op_t* ops, int n_ops, const arpde_buf_t* bufs, int n_bufs, void* const* params_host,
                                     int n_params, const float* x, int64_t x_bstride, int N, int H, int W, const float* gy,
                                     const float* y, const float* workspace, const float* scale_shift, int bn_total,
                                     const float* saved_mean_invstd, float* gwork, float* gx, void* const* grads_host, float* T1,
                                     float* T2, float* wT, double* d64, int training, float* tc_scratch, int64_t tc_scratch_floats,
                                     arpde_stream_t stream) {
  return densed_backward_impl(ops, n_ops, bufs, n_bufs, params_host, n_params, x, x_bstride, N, H, W, gy, y, workspace, scale_shift,
                              bn_total, saved_mean_invstd, gwork, gx, grads_host, T1, T2, wT, d64, training, tc_scratch,
                              tc_scratch_floats, stream, nullptr);
}

// Data-parallel backward: the (sum dz, sum dz*xhat) pairs of every BatchNorm are summed over the ranks between the reduction and the
// apply pass (global-batch BatchNorm backward); dgamma / dbeta and the weight gradients stay per-rank sums, to be averaged by the
// caller's gradient all-reduce.
extern "C" int arpde_densed_backward_dp(const arpde_op_t* ops, int n_ops, const arpde_buf_t* bufs, int n_bufs, void* const* params_host,
                                        int n_params, const float* x, int64_t x_bstride, int N, int H, int W, const float* gy,
                                        const float* y, const float* workspace, const float* scale_shift, int bn_total,
                                        const float* saved_mean_invstd, float* gwork, float* gx, void* const* grads_host, float* T1,
                                        float* T2, float* wT, double* d64, int training, float* tc_scratch, int64_t tc_scratch_floats,
                                        arpde_allreduce_fn allreduce, void* user, int dp_world, arpde_stream_t stream) {
  ARPDE_CHECK_ARG(dp_world >= 1 && (dp_world == 1 || allreduce), "densed_backward_dp: %d ranks need an all-reduce callback", dp_world);
  DpHook h; h.allreduce = (int (*)(void*, double*, long long))allreduce; h.user = user; h.world = dp_world;
  return densed_backward_impl(ops, n_ops, bufs, n_bufs, params_host, n_params, x, x_bstride, N, H, W, gy, y, workspace, scale_shift,
                              bn_total, saved_mean_invstd, gwork, gx, grads_host, T1, T2, wT, d64, training, tc_scratch,
                              tc_scratch_floats, stream, &h);
}

// RenderingNetwork (fields.py:114-175) and RefColor (fields.py:271-335, math_utils.py:12-22,138-144):
// ReLU MLPs whose first-layer input is [generated block | feature block].  FP32 path on the SIMT GEMM engine.
#include "gemm_tc.cuh"
#include "prof.cuh"
#include "sdf_chain.cuh"
#include <string.h>

namespace fneus {

int num_sms();
__global__ void extract_cols_kernel(const float*, int, int, int, float*, int, long long);
__global__ void sigmoid_bwd_kernel(const float*, const float*, int, float*, int, long long);
static inline int ew_blocks2(long long n) { return cdiv(n, 256); }
// a[m][j] = j < n ? in[m][j] : 0   (re-layout of a [M, n] gradient into a padded [M, lda] buffer)
__global__ void pad_rows_kernel(const float* in, int n, float* a, int lda, long long M) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * lda) return;
  long long m = idx / lda;
  int j = (int)(idx - m * lda);
  a[idx] = j < n ? in[m * n + j] : 0.f;
}

struct Lin { long long woff, boff; int in, out; };

// ReLU chain forward: layer 0 from `a0`, hidden activations to Hs[1..n] ([M, ldh]), last layer with `last_mode`
// into out ([M, ld_out]).
static size_t relu_chain_img_bytes(const Lin* lin, int n_lin, int gen0) {
  size_t b = 0;
  for (int l = 0; l < n_lin; l++) {
    b += wimg_bytes(lin[l].out, l == 0 ? gen0 : 0, l == 0 ? lin[l].in - gen0 : lin[l].in) + 1024;
    b += wimg_bytes(lin[l].in, 0, lin[l].out) + 1024;
  }
  return b + 8192;
}
static inline long long hid_floats(long long M, int hid, bool img) { return mat_floats(M, hid, img) + 256; }
static inline float* align1k(float* q) {
  return reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(q) + 1023) & ~(uintptr_t)1023);
}
static inline void zero_if_ragged(bool img, float* ptr, long long floats, long long M, cudaStream_t st) {
  if (img && (M & 127)) cudaMemsetAsync(ptr, 0, (size_t)floats * 4, st);
}
static ImgArena arena_at(float* after_floats, size_t cap_bytes) {
  ImgArena ar = arena_make(nullptr, 0);
  if (precision_mode() == 1 && after_floats) {
    ar.base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(after_floats) + 1023) & ~(uintptr_t)1023);
    ar.cap = cap_bytes - 1024;
  }
  return ar;
}

// First-operand image ([generated blocks | memory blocks], at most SC_OPB_RELU of 64 columns) kept for the
// layer-0 weight gradient, and the per-layer dY images of the fused backward chain.
static inline long long a0_img_floats(long long M) { return mat_floats(M, SC_OPB_RELU * 64, true) + 256; }

// Fused-chain eligibility (sdf_chain.cuh, <5, false> instantiation): 256-wide hidden layers (4-block activation
// images), first operand = at most one generated block + memory blocks, at most 5 blocks in all.
static bool chain_fusable(const Lin* lin, int n_lin, const ASeg& a0, int hid) {
  if (precision_mode() != 1 || tc_prepare() != 0 || sdf_chain_prepare() != 0) return false;
  if (tc_debug_flags() & 8) return false;                        // debug: layered execution
  if (n_lin < 2 || n_lin + 2 > SC_MAXS || n_lin > sc_bias_slots<FAM_RELU>() || a0.gen.deriv) return false;
  const int kb0 = cdiv(a0.gen.ncols, TC_BK) + cdiv(a0.kmem, TC_BK);
  if (kb0 < 1 || kb0 > SC_OPB_RELU || hid != 256 || a0.gen.ncols > TC_BK || a0.kmem > 256) return false;
  for (int l = 0; l < n_lin; l++) {
    if (lin[l].out > 256) return false;
    if (l > 0 && lin[l].in != hid) return false;
    if (l < n_lin - 1 && lin[l].out != hid) return false;
  }
  return true;
}

// defer: when non-null and the fused path applies, the chain is NOT launched; its arguments and FLOPs are returned for a
// paired launch (relu_chain_pair_launch) and `defer->used` is set.
struct ChainDefer { SdfChainArgs g; double flops; bool used; };
static bool relu_fwd_fused(const Lin* lin, int n_lin, const ASeg& a0, int ldh, int last_mode) {
  return ldh < 0 && (last_mode == EPI_SIGMOID || last_mode == EPI_LINEAR) && -ldh == cdiv(lin[0].out, TC_BK) &&
         chain_fusable(lin, n_lin, a0, lin[0].out);
}
static int relu_chain_fwd(const float* w, const Lin* lin, int n_lin, const ASeg& a0, float* const* Hs, int ldh,
                           int last_mode, float* out, int ld_out, long long M, cudaStream_t st, ImgArena& ar,
                           float* a0_img = nullptr, ChainDefer* defer = nullptr) {
  const uint8_t* img[16];
  bool all_img = true;
  // the fused forward chain computes on FP16 operands and leaves FP16 activation images (relu_chain_bwd recomputes this
  // predicate to know their format); the layered path stays BF16
  const bool fuse = relu_fwd_fused(lin, n_lin, a0, ldh, last_mode);
  for (int l = 0; l < n_lin; l++) {
    ASeg a = l == 0 ? a0 : aseg_mem(Hs[l], ldh, lin[l].in);
    img[l] = make_wimg(ar, false, w + lin[l].woff, lin[l].in, 0, lin[l].out, a.wred_gen, a.gen.ncols, a.wred_mem,
                       a.kmem, st, fuse ? 1 : 0);
    all_img = all_img && img[l] != nullptr;
  }
  ar.flush(st);
  if (fuse && !all_img) return FNEUS_ERR_WORKSPACE;
  if (fuse) {
    SdfChainArgs g;
    memset(&g, 0, sizeof(g));
    double flops = 0.0;
    g.nsteps = n_lin;
    for (int l = 0; l < n_lin; l++) {
      const bool last = l == n_lin - 1;
      const int KB = l == 0 ? cdiv(a0.gen.ncols, TC_BK) + cdiv(a0.kmem, TC_BK) : cdiv(lin[l].in, TC_BK);
      SdfStep S = sdf_step(last ? SC_OUT : SC_RELU, img[l], KB, lin[l].out, 0);
      S.src = l == 0 ? (a0.gen.ncols > 0 ? SRC_GENMEM : SRC_MEM) : SRC_CHAIN;
      S.bias = w + lin[l].boff; S.bias_slot = l;
      S.img_out = last ? nullptr : Hs[l + 1];
      S.out = last ? out : nullptr;
      S.ldo = ld_out; S.act = last_mode == EPI_SIGMOID ? 1 : 0; S.accumulate = 0;
      g.st[l] = S;
      flops += 2.0 * (double)M * lin[l].in * lin[l].out;
    }
    g.gen = a0.gen; g.gen_t = a0.gen; g.mem = a0.mem; g.ldm = a0.ldm; g.kmem = a0.kmem; g.a0_img = a0_img;
    g.f16 = 1;
    g.beta = 1.f; g.M = M; g.dbg = (tc_debug_flags() & 128) ? 1 : 0;
    g.xflags = (tc_debug_flags() >> 8) & 15;
    if (defer) { defer->g = g; defer->flops = flops; defer->used = true; return FNEUS_OK; }
    sdf_chain_launch(g, flops, st, FAM_RELU);
    return FNEUS_OK;
  }
  if (a0.ldm < 0) return FNEUS_ERR_UNSUPPORTED;              // an image first operand exists on the fused chain only
  for (int l = 0; l < n_lin; l++) {
    ASeg a = l == 0 ? a0 : aseg_mem(Hs[l], ldh, lin[l].in);
    Epi e = epi_default();
    e.bias = w + lin[l].boff;
    if (l < n_lin - 1) { e.mode = EPI_RELU; e.C = Hs[l + 1]; e.ldc = ldh; }
    else { e.mode = last_mode; e.C = out; e.ldc = ld_out; }
    launch_gemm_fwd(a, w + lin[l].woff, lin[l].in, 0, M, lin[l].out, e, st, img[l]);
  }
  return FNEUS_OK;
}

// ReLU chain backward. a_last = gradient wrt the last linear's pre-activation ([M, ld_last]).
// Layer-0 input gradient: generated block -> dsmall ([M, ld_small], may be null), feature block -> d_feats.
// dybuf: (n_lin - 1) hidden-sized buffers (the layered path ping-pongs between the first two).
// a0_img: image of the first operand written by the fused forward (null: regenerate it in the layer-0 wgrad).
static int relu_chain_bwd(const float* w, float* dw, const Lin* lin, int n_lin, const ASeg& a0, float* const* Hs,
                           int ldh, const float* a_last, int ld_last, float* dybuf, long long dy_stride, float* dsmall,
                           int ld_small, float* d_feats, int ld_feats, int accumulate_feats, long long M,
                           cudaStream_t st, ImgArena& ar, const float* a0_img = nullptr, ChainDefer* defer = nullptr,
                           WgradGroup* shared_wg = nullptr) {
  const int sms = num_sms();
  // the forward pass took the fused chain (FP16 activation images) exactly when relu_fwd_fused held; the fused
  // backward must then be possible too (the layered path reads BF16 images)
  const bool fwd_fused = relu_fwd_fused(lin, n_lin, a0, ldh, EPI_SIGMOID);
  const bool fused = fwd_fused && lin[n_lin - 1].out <= TC_BK * SC_OPB_RELU && n_lin >= 2 &&
                     (ld_small & 3) == 0 && (ld_feats & 3) == 0;
  if (fwd_fused && !fused) return FNEUS_ERR_UNSUPPORTED;
  if ((a0.ldm < 0 || ld_feats < 0) && (!fused || accumulate_feats)) return FNEUS_ERR_UNSUPPORTED;
  if (fused) {
    // ---- weight images: MN-major W_l for l >= 1; layer 0 split into its feature and generated column ranges ----
    const uint8_t* img[16];
    bool all_img = true;
    for (int l = n_lin - 1; l >= 1; l--) {
      img[l] = make_wimg(ar, true, w + lin[l].woff, lin[l].in, 0, lin[l].in, 0, 0, 0, lin[l].out, st);
      all_img = all_img && img[l] != nullptr;
    }
    const uint8_t* img_feat = nullptr;
    const uint8_t* img_gen = nullptr;
    if (d_feats && a0.kmem > 0) {
      img_feat = make_wimg(ar, true, w + lin[0].woff, lin[0].in, a0.wred_mem, a0.kmem, 0, 0, 0, lin[0].out, st);
      all_img = all_img && img_feat != nullptr;
    }
    if (dsmall && a0.gen.ncols > 0) {
      img_gen = make_wimg(ar, true, w + lin[0].woff, lin[0].in, a0.wred_gen, a0.gen.ncols, 0, 0, 0, lin[0].out, st);
      all_img = all_img && img_gen != nullptr;
    }
    if (!all_img) return FNEUS_ERR_WORKSPACE;
    {
      ar.flush(st);
      SdfChainArgs g;
      memset(&g, 0, sizeof(g));
      double flops = 0.0;
      int ns = 0;
      for (int l = n_lin - 1; l >= 1; l--) {
        SdfStep S = sdf_step(SC_MASK, img[l], cdiv(lin[l].out, TC_BK), lin[l].in, 1);
        S.src = l == n_lin - 1 ? SRC_MEM : SRC_CHAIN;
        S.h = Hs[l];
        S.img_out = dybuf + (long long)(l - 1) * dy_stride;          // dz_{l-1} = (dz_l W_l) * [h_l > 0]
        g.st[ns++] = S;
        flops += 2.0 * (double)M * lin[l].in * lin[l].out;
      }
      if (img_feat) {
        SdfStep S = sdf_step(SC_OUT, img_feat, cdiv(lin[0].out, TC_BK), a0.kmem, 1);
        if (ld_feats < 0) { S.e_out = d_feats; }                    // BF16 operand image (feat_image): see SC_OUT
        else { S.out = d_feats; S.ldo = ld_feats; S.accumulate = accumulate_feats; }
        g.st[ns++] = S;
        flops += 2.0 * (double)M * a0.kmem * lin[0].out;
      }
      if (img_gen) {
        SdfStep S = sdf_step(SC_OUT, img_gen, cdiv(lin[0].out, TC_BK), a0.gen.ncols, 1);
        S.out = dsmall; S.ldo = ld_small;
        g.st[ns++] = S;
        flops += 2.0 * (double)M * a0.gen.ncols * lin[0].out;
      }
      g.nsteps = ns;
      g.gen = gen_none(); g.gen_t = g.gen; g.mem = a_last; g.ldm = ld_last; g.kmem = lin[n_lin - 1].out;
      g.beta = 1.f; g.M = M; g.dbg = (tc_debug_flags() & 128) ? 1 : 0;
      g.xflags = (tc_debug_flags() >> 8) & 15;
      if (defer) { defer->g = g; defer->flops = flops; defer->used = true; }
      else sdf_chain_launch(g, flops, st, FAM_RELU);
      // ---- all weight gradients in one grouped launch (the caller's group when chains are paired) ----
      WgradGroup own_wg;
      WgradGroup& wg = shared_wg ? *shared_wg : own_wg;
      if (!shared_wg) wg.reset(M, sms);
      for (int l = n_lin - 1; l >= 0; l--) {
        const float* dy = l == n_lin - 1 ? a_last : dybuf + (long long)l * dy_stride;    // dz_l
        const int ldy = l == n_lin - 1 ? ld_last : ldh;
        float* dW = dw + lin[l].woff;
        float* db = dw + lin[l].boff;
        // dz_l: BF16 images of this pass; h_l and the first operand: FP16 images of the forward pass
        if (l > 0) wg.add(dy, ldy, aseg_mem(Hs[l], ldh, lin[l].in), dW, lin[l].in, 0, db, lin[l].out, st, 0, 1);
        else if (a0_img == nullptr) wg.add(dy, ldy, a0, dW, lin[l].in, 0, db, lin[l].out, st);
        else {
          const int kbg = cdiv(a0.gen.ncols, TC_BK), kb0 = kbg + cdiv(a0.kmem, TC_BK);
          bool bias_done = false;
          if (a0.kmem > 0) {
            wg.add(dy, ldy, aseg_mem(a0_img + (size_t)kbg * (TC_A_BYTES / 4), -kb0, a0.kmem, a0.wred_mem), dW, lin[l].in, 0,
                   db, lin[l].out, st, 0, 1);
            bias_done = true;
          }
          if (a0.gen.ncols > 0)
            wg.add(dy, ldy, aseg_mem(a0_img, -kb0, a0.gen.ncols, a0.wred_gen), dW, lin[l].in, 0, bias_done ? nullptr : db,
                   lin[l].out, st, 0, 1);
        }
      }
      if (!shared_wg) wg.flush(st);
      return FNEUS_OK;
    }
  }
  const float* al = a_last;
  int ld_al = ld_last;
  float* ab[2] = {dybuf, dybuf + dy_stride};
  const uint8_t* img[16];
  for (int l = n_lin - 1; l >= 0; l--)
    img[l] = (l > 0 || dsmall || d_feats)
                 ? make_wimg(ar, true, w + lin[l].woff, lin[l].in, 0, lin[l].in, 0, 0, 0, lin[l].out, st) : nullptr;
  ar.flush(st);
  for (int l = n_lin - 1; l >= 0; l--) {
    ASeg h = l == 0 ? a0 : aseg_mem(Hs[l], ldh, lin[l].in);
    launch_gemm_wgrad(al, ld_al, h, dw + lin[l].woff, lin[l].in, 0, dw + lin[l].boff, M, lin[l].out, sms, st);
    ASeg a = aseg_mem(al, ld_al, lin[l].out);
    Epi e = epi_default();
    e.mode = EPI_RELUMASK;
    if (l > 0) {
      e.H = Hs[l]; e.ldh = ldh;
      e.C = ab[l & 1]; e.ldc = ldh;
      launch_gemm_bwd_data(a, w + lin[l].woff, lin[l].in, 0, M, lin[l].in, e, st, img[l]);
      al = ab[l & 1]; ld_al = ldh;
    } else if (dsmall || d_feats) {
      e.H = nullptr;
      e.C = dsmall; e.ldc = ld_small;
      e.csplit = a0.gen.ncols;
      e.C2 = d_feats; e.ldc2 = ld_feats; e.accumulate2 = accumulate_feats;
      launch_gemm_bwd_data(a, w + lin[l].woff, lin[l].in, 0, M, lin[l].in, e, st, img[l]);
    }
  }
  return FNEUS_OK;
}

// ------------------------------------------------------------------------------------------------
// RenderingNetwork
// ------------------------------------------------------------------------------------------------
struct ColorPlan { int n_lin; Lin lin[12]; long long pack; int hid; int ldh; bool img; int gen_cols; bool ok; };
static ColorPlan color_plan(const fneus_color_cfg* c) {
  ColorPlan p;
  p.ok = c && c->n_layers >= 1 && c->n_layers <= 10 && c->d_feature > 0 && c->d_feature % 4 == 0 &&
         c->d_hidden > 0 && c->d_hidden % 4 == 0 && c->d_out >= 1 && c->d_out <= 4 && c->multires_view >= 0 &&
         c->multires_view <= 8;
  if (!p.ok) return p;
  p.n_lin = c->n_layers + 1;
  p.gen_cols = 3 + pe_dim(3, c->multires_view) + 3;
  p.hid = c->d_hidden; p.img = precision_mode() == 1; p.ldh = mat_ld(p.hid, p.img);
  long long off = 0;
  for (int l = 0; l < p.n_lin; l++) {
    p.lin[l].in = l == 0 ? p.gen_cols + c->d_feature : c->d_hidden;
    p.lin[l].out = l == p.n_lin - 1 ? c->d_out : c->d_hidden;
    p.lin[l].woff = off; off += (long long)p.lin[l].in * p.lin[l].out;
    p.lin[l].boff = off; off += p.lin[l].out;
  }
  p.pack = off;
  return p;
}
static long long color_scratch_main(const fneus_color_cfg* c, const ColorPlan& p, long long n) {
  long long per = (long long)c->n_layers * hid_floats(n, p.hid, p.img);
  long long bwd = (long long)(c->n_layers > 2 ? c->n_layers : 2) * hid_floats(n, p.hid, p.img) +
                  n * (round_up(p.gen_cols, 4) + 4);
  return (per > bwd ? per : bwd) + 2048;
}
static ASeg color_a0(const fneus_color_cfg* c, const ColorPlan& p, const float* pts, const float* nrm,
                     const float* view, const float* feats) {
  GenSpec g = gen_none();
  gen_add(g, pts, 3, 0);
  gen_add(g, view, 3, c->multires_view);
  gen_add(g, nrm, 3, 0);
  // feat_image: the features arrive as the SDF chain's FP16 operand image (one tile = 4 blocks of 128 x 64)
  return aseg_gen_mem(g, 0, feats, c->feat_image ? -4 : c->d_feature, c->d_feature, p.gen_cols);
}

// ------------------------------------------------------------------------------------------------
// RefColor
// ------------------------------------------------------------------------------------------------
struct RefPlan { Lin cd[5]; Lin vd[4]; Lin cs; long long pack; int hid; int ldh; bool img; bool ok; };
static RefPlan ref_plan(const fneus_ref_cfg* c) {
  RefPlan p;
  p.ok = c && c->d_feature > 0 && c->d_feature % 4 == 0 && c->d_hidden > 0 && c->d_hidden % 4 == 0;
  if (!p.ok) return p;
  p.hid = c->d_hidden; p.img = precision_mode() == 1; p.ldh = mat_ld(p.hid, p.img);
  long long off = 0;
  auto put = [&](Lin& l, int in, int out) {
    l.in = in; l.out = out; l.woff = off; off += (long long)in * out; l.boff = off; off += out;
  };
  put(p.cd[0], 30 + c->d_feature, c->d_hidden);
  for (int i = 1; i < 4; i++) put(p.cd[i], c->d_hidden, c->d_hidden);
  put(p.cd[4], c->d_hidden, 3);
  put(p.vd[0], 33 + c->d_feature, c->d_hidden);
  for (int i = 1; i < 4; i++) put(p.vd[i], c->d_hidden, c->d_hidden);
  put(p.cs, c->d_hidden, 1);
  p.pack = off;
  return p;
}

// 2 x 4 hidden-sized dz buffers (one set per chain: the two chains run concurrently), dsmall_cd [n,32], dsmall_cs [n,36],
// a_cd / a_cs [n,4], the specular chain's feature gradient [n, d_feature <= 256]
static long long ref_scratch_main(const RefPlan& p, long long n) {
  return 8LL * hid_floats(n, p.hid, p.img) + (32 + 36 + 8 + 256) * n + 4096;
}
__global__ void add_into_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}
__device__ __forceinline__ float srgb_f(float c) {
  const float eps = 1.1920928955078125e-07f;
  return c <= 0.0031308f ? (323.0f / 25.0f) * c : (211.0f * powf(fmaxf(eps, c), 5.0f / 12.0f) - 11.0f) / 200.0f;
}
__device__ __forceinline__ float srgb_df(float c) {
  const float eps = 1.1920928955078125e-07f;
  if (c <= 0.0031308f) return 323.0f / 25.0f;
  return c > eps ? (211.0f / 200.0f) * (5.0f / 12.0f) * powf(c, -7.0f / 12.0f) : 0.f;
}
__device__ __forceinline__ float clip01_mask(float v) { return (v >= 0.f && v <= 1.f) ? 1.f : 0.f; }

// refl = reflect(-d, normalize(n))   (math_utils.py:12-22, fields.py:305-307)
__global__ void ref_prep_kernel(const float* dirs, const float* nrm, float* refl, long long M) {
  long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float eps = 1.1920928955078125e-07f;
  float nx = nrm[m * 3], ny = nrm[m * 3 + 1], nz = nrm[m * 3 + 2];
  float inv = 1.f / sqrtf(fmaxf(nx * nx + ny * ny + nz * nz, eps));
  nx *= inv; ny *= inv; nz *= inv;
  float wx = -dirs[m * 3], wy = -dirs[m * 3 + 1], wz = -dirs[m * 3 + 2];
  float dn = 2.f * (wx * nx + wy * ny + wz * nz);
  refl[m * 3] = dn * nx - wx; refl[m * 3 + 1] = dn * ny - wy; refl[m * 3 + 2] = dn * nz - wz;
}

// yd [M,3] diffuse (sigmoid), ys [M] specular (sigmoid) -> three sRGB outputs (fields.py:320-328)
__global__ void ref_final_kernel(const float* yd, const float* ys, float* rgb, float* spec, float* diff, long long M) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * 3) return;
  long long m = idx / 3;
  float s = ys[m], d = yd[idx];
  rgb[idx] = fminf(fmaxf(srgb_f(s + d), 0.f), 1.f);
  spec[idx] = fminf(fmaxf(srgb_f(s), 0.f), 1.f);
  diff[idx] = fminf(fmaxf(srgb_f(d), 0.f), 1.f);
}
// backward of ref_final + the two sigmoids: a_cd [M,4], a_cs [M,4] (col 0)
__global__ void ref_final_bwd_kernel(const float* yd, const float* ys, const float* d_rgb, const float* d_spec,
                                     const float* d_diff, float* a_cd, float* a_cs, long long M) {
  long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float s = ys[m];
  float gs = 0.f;
  for (int c = 0; c < 3; c++) {
    float d = yd[m * 3 + c];
    float gd = 0.f;
    if (d_rgb) {
      float v = srgb_f(s + d);
      float g = d_rgb[m * 3 + c] * clip01_mask(v) * srgb_df(s + d);
      gd += g; gs += g;
    }
    if (d_spec) gs += d_spec[m * 3 + c] * clip01_mask(srgb_f(s)) * srgb_df(s);
    if (d_diff) gd += d_diff[m * 3 + c] * clip01_mask(srgb_f(d)) * srgb_df(d);
    a_cd[m * 4 + c] = gd * d * (1.f - d);
  }
  a_cd[m * 4 + 3] = 0.f;
  a_cs[m * 4 + 0] = gs * s * (1.f - s);
  a_cs[m * 4 + 1] = 0.f; a_cs[m * 4 + 2] = 0.f; a_cs[m * 4 + 3] = 0.f;
}

// d_n from the generated blocks: cd block = [pts(3), PE4(n)(27)], cs block = [n(3), pts(3), PE4(refl)(27)]
__global__ void ref_dn_kernel(const float* dirs, const float* nrm, const float* refl, const float* ds_cd, int ld_cd,
                              const float* ds_cs, int ld_cs, float* d_n, long long M) {
  long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float eps = 1.1920928955078125e-07f;
  float n[3] = {nrm[m * 3], nrm[m * 3 + 1], nrm[m * 3 + 2]};
  float r[3] = {refl[m * 3], refl[m * 3 + 1], refl[m * 3 + 2]};
  float dn[3], dr[3];
  for (int c = 0; c < 3; c++) {
    // PE4(n) at cd cols 3..29 ; raw n at cs cols 0..2
    float g = ds_cd[m * ld_cd + 3 + c] + ds_cs[m * ld_cs + c];
    float gr = ds_cs[m * ld_cs + 6 + c];
    for (int k = 0; k < 4; k++) {
      float f = (float)(1u << k);
      float sn, cn, sr, cr;
      sincosf(n[c] * f, &sn, &cn);
      sincosf(r[c] * f, &sr, &cr);
      g += f * cn * ds_cd[m * ld_cd + 3 + 3 * (1 + 2 * k) + c] - f * sn * ds_cd[m * ld_cd + 3 + 3 * (2 + 2 * k) + c];
      gr += f * cr * ds_cs[m * ld_cs + 6 + 3 * (1 + 2 * k) + c] - f * sr * ds_cs[m * ld_cs + 6 + 3 * (2 + 2 * k) + c];
    }
    dn[c] = g; dr[c] = gr;
  }
  // refl = 2 (wo.nh) nh - wo ; nh = n / sqrt(max(|n|^2, eps))
  float n2 = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
  float inv = 1.f / sqrtf(fmaxf(n2, eps));
  float nh[3] = {n[0] * inv, n[1] * inv, n[2] * inv};
  float wo[3] = {-dirs[m * 3], -dirs[m * 3 + 1], -dirs[m * 3 + 2]};
  float won = wo[0] * nh[0] + wo[1] * nh[1] + wo[2] * nh[2];
  float drn = dr[0] * nh[0] + dr[1] * nh[1] + dr[2] * nh[2];
  float dnh[3];
  for (int c = 0; c < 3; c++) dnh[c] = 2.f * (drn * wo[c] + won * dr[c]);
  if (n2 > eps) {
    float dot = dnh[0] * nh[0] + dnh[1] * nh[1] + dnh[2] * nh[2];
    for (int c = 0; c < 3; c++) dn[c] += (dnh[c] - nh[c] * dot) * inv;
  } else {
    for (int c = 0; c < 3; c++) dn[c] += dnh[c] * inv;
  }
  for (int c = 0; c < 3; c++) d_n[m * 3 + c] = dn[c];
}

}  // namespace fneus

using namespace fneus;

extern "C" {

long long fneus_color_pack_floats(const fneus_color_cfg* cfg) {
  PrecScope prec_scope_(cfg ? cfg->precision : 0);
  ColorPlan p = color_plan(cfg);
  return p.ok ? p.pack : -1;
}
// saved: H_1..H_n ; scratch: 2 abufs + dsmall + a_last
long long fneus_color_saved_floats(const fneus_color_cfg* cfg, long long n) {
  PrecScope prec_scope_(cfg ? cfg->precision : 0);
  ColorPlan p = color_plan(cfg);
  return p.ok ? (long long)cfg->n_layers * hid_floats(n, p.hid, p.img) + (p.img ? a0_img_floats(n) : 0) + 1024 : -1;
}
long long fneus_color_scratch_floats(const fneus_color_cfg* cfg, long long n) {
  PrecScope prec_scope_(cfg ? cfg->precision : 0);
  ColorPlan p = color_plan(cfg);
  if (!p.ok) return -1;
  return color_scratch_main(cfg, p, n) + (long long)(relu_chain_img_bytes(p.lin, p.n_lin, p.gen_cols) / 4) + 256;
}

int fneus_color_feat_image_ok(const fneus_color_cfg* cfg) {
  if (!cfg) return 0;
  PrecScope prec_scope_(cfg->precision);
  ColorPlan p = color_plan(cfg);
  if (!p.ok || cfg->d_feature != 256) return 0;
  fneus_color_cfg c = *cfg;
  c.feat_image = 1;
  const ASeg a0 = color_a0(&c, p, nullptr, nullptr, nullptr, nullptr);
  return (relu_fwd_fused(p.lin, p.n_lin, a0, p.ldh, EPI_SIGMOID) && p.lin[p.n_lin - 1].out <= TC_BK * SC_OPB_RELU) ? 1 : 0;
}

int fneus_color_fwd(const fneus_color_cfg* cfg, const float* wpack, const float* points, const float* normals,
                    const float* view_dirs, const float* feats, long long M, float* rgb_out, float* saved,
                    float* scratch, void* stream) {
  PrecScope prec_scope_(cfg ? cfg->precision : 0);
  ColorPlan p = color_plan(cfg);
  if (!p.ok) return FNEUS_ERR_UNSUPPORTED;
  if (M == 0) return FNEUS_OK;
  if (!wpack || !points || !normals || !view_dirs || !feats || !rgb_out || (!saved && !scratch)) return FNEUS_ERR_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  float* base = align1k(saved ? saved : scratch);
  float* Hs[12];
  const long long hf = hid_floats(M, p.hid, p.img);
  for (int l = 1; l <= cfg->n_layers; l++) Hs[l] = base + (long long)(l - 1) * hf;
  zero_if_ragged(p.img, base, (long long)cfg->n_layers * hf, M, st);
  ImgArena ar = arena_at(scratch ? scratch + color_scratch_main(cfg, p, M) : nullptr,
                         relu_chain_img_bytes(p.lin, p.n_lin, p.gen_cols));
  { const int rc_ = relu_chain_fwd(wpack, p.lin, p.n_lin, color_a0(cfg, p, points, normals, view_dirs, feats), Hs, p.ldh, EPI_SIGMOID,
                 rgb_out, cfg->d_out, M, st, ar, (saved && p.img) ? base + (long long)cfg->n_layers * hf : nullptr); if (rc_) return rc_; }
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

int fneus_color_bwd(const fneus_color_cfg* cfg, const float* wpack, const float* points, const float* normals,
                    const float* view_dirs, const float* feats, long long M, const float* rgb, const float* d_rgb,
                    float* d_normals, float* d_feats, float* saved, float* scratch, float* d_wpack, void* stream) {
  PrecScope prec_scope_(cfg ? cfg->precision : 0);
  ColorPlan p = color_plan(cfg);
  if (!p.ok) return FNEUS_ERR_UNSUPPORTED;
  if (M == 0) return FNEUS_OK;
  if (!wpack || !points || !normals || !view_dirs || !feats || !rgb || !d_rgb || !saved || !scratch || !d_wpack)
    return FNEUS_ERR_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  float* Hs[12];
  const long long hf = hid_floats(M, p.hid, p.img);
  for (int l = 1; l <= cfg->n_layers; l++) Hs[l] = align1k(saved) + (long long)(l - 1) * hf;
  const int lds = round_up(p.gen_cols, 4);
  const int nbuf = cfg->n_layers > 2 ? cfg->n_layers : 2;
  float* ab0 = align1k(scratch);
  float* dsmall = ab0 + nbuf * hf;
  zero_if_ragged(p.img, ab0, nbuf * hf, M, st);
  float* alast = dsmall + M * lds;
  prof_begin(PC_ELEMENTWISE, 0.0, 0.0, st);
  sigmoid_bwd_kernel<<<ew_blocks2(M * 4), 256, 0, st>>>(d_rgb, rgb, cfg->d_out, alast, 4, M);
  prof_end(st);
  ImgArena ar = arena_at(scratch + color_scratch_main(cfg, p, M), relu_chain_img_bytes(p.lin, p.n_lin, p.gen_cols));
  { const int rc_ = relu_chain_bwd(wpack, d_wpack, p.lin, p.n_lin, color_a0(cfg, p, points, normals, view_dirs, feats), Hs, p.ldh,
                 alast, 4, ab0, hf, d_normals ? dsmall : nullptr, lds, d_feats, cfg->feat_image ? -4 : cfg->d_feature, 0, M, st, ar,
                 p.img ? align1k(saved) + (long long)cfg->n_layers * hf : nullptr); if (rc_) return rc_; }
  if (d_normals) {
    prof_begin(PC_ELEMENTWISE, 0.0, 0.0, st);
    extract_cols_kernel<<<ew_blocks2(M * 3), 256, 0, st>>>(dsmall, lds, p.gen_cols - 3, 3, d_normals, 3, M);
    prof_end(st);
  }
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

long long fneus_ref_pack_floats(const fneus_ref_cfg* cfg) {
  PrecScope prec_scope_(cfg ? cfg->precision : 0);
  RefPlan p = ref_plan(cfg);
  return p.ok ? p.pack : -1;
}
// saved: cd H1..H4, cs G1..G4, yd [M,4], ys [M,4], refl [M,4]
long long fneus_ref_saved_floats(const fneus_ref_cfg* cfg, long long n) {
  PrecScope prec_scope_(cfg ? cfg->precision : 0);
  RefPlan p = ref_plan(cfg);
  return p.ok ? 8LL * hid_floats(n, p.hid, p.img) + 12 * n + 2048 + (p.img ? 2 * a0_img_floats(n) + 512 : 0) : -1;
}
// scratch: 2 abufs, dsmall_cd [M,32], dsmall_cs [M,36], a_cd [M,4], a_cs [M,4]
long long fneus_ref_scratch_floats(const fneus_ref_cfg* cfg, long long n) {
  PrecScope prec_scope_(cfg ? cfg->precision : 0);
  RefPlan p = ref_plan(cfg);
  if (!p.ok) return -1;
  Lin chain[5] = {p.vd[0], p.vd[1], p.vd[2], p.vd[3], p.cs};
  return ref_scratch_main(p, n) +
         (long long)((relu_chain_img_bytes(p.cd, 5, 30) + relu_chain_img_bytes(chain, 5, 33)) / 4) + 256;
}

namespace {
struct RefBufs { float* H[5]; float* G[5]; float* yd; float* ys; float* refl; float* a0_cd; float* a0_cs; };
RefBufs ref_carve(const RefPlan& p, float* saved, long long M) {
  RefBufs b;
  float* ptr = align1k(saved);
  for (int i = 1; i <= 4; i++) { b.H[i] = ptr; ptr += hid_floats(M, p.hid, p.img); }
  for (int i = 1; i <= 4; i++) { b.G[i] = ptr; ptr += hid_floats(M, p.hid, p.img); }
  b.yd = ptr; ptr += M * 4; b.ys = ptr; ptr += M * 4; b.refl = ptr; ptr += M * 4;
  b.a0_cd = b.a0_cs = nullptr;
  if (p.img) { b.a0_cd = align1k(ptr); b.a0_cs = b.a0_cd + a0_img_floats(M); }
  return b;
}
ASeg ref_cd_a0(const fneus_ref_cfg* c, const float* pts, const float* nrm, const float* feats) {
  GenSpec g = gen_none();
  gen_add(g, pts, 3, 0);
  gen_add(g, nrm, 3, 4);
  return aseg_gen_mem(g, 0, feats, c->d_feature, c->d_feature, 30);
}
ASeg ref_cs_a0(const fneus_ref_cfg* c, const float* pts, const float* nrm, const float* refl, const float* feats) {
  GenSpec g = gen_none();
  gen_add(g, nrm, 3, 0);
  gen_add(g, pts, 3, 0);
  gen_add(g, refl, 3, 4);
  return aseg_gen_mem(g, 0, feats, c->d_feature, c->d_feature, 33);
}
}  // namespace

int fneus_ref_fwd(const fneus_ref_cfg* cfg, const float* wpack, const float* points, const float* feats,
                  const float* dirs, const float* normals, long long M, float* rgb_out, float* spec_out,
                  float* diff_out, float* saved, float* scratch, void* stream) {
  PrecScope prec_scope_(cfg ? cfg->precision : 0);
  RefPlan p = ref_plan(cfg);
  if (!p.ok) return FNEUS_ERR_UNSUPPORTED;
  if (M == 0) return FNEUS_OK;
  if (!wpack || !points || !feats || !dirs || !normals || !rgb_out || !spec_out || !diff_out || !saved)
    return FNEUS_ERR_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  RefBufs b = ref_carve(p, saved, M);
  Lin chain0[5] = {p.vd[0], p.vd[1], p.vd[2], p.vd[3], p.cs};
  zero_if_ragged(p.img, align1k(saved), 8LL * hid_floats(M, p.hid, p.img), M, st);
  ImgArena ar = arena_at(scratch ? scratch + ref_scratch_main(p, M) : nullptr,
                         relu_chain_img_bytes(p.cd, 5, 30) + relu_chain_img_bytes(chain0, 5, 33));
  prof_begin(PC_ELEMENTWISE, 0.0, 0.0, st);
  ref_prep_kernel<<<ew_blocks2(M), 256, 0, st>>>(dirs, normals, b.refl, M);
  prof_end(st);
  // The diffuse and the specular network are independent: on the fused path both chains go out in ONE launch.
  ChainDefer d_cd, d_cs;
  d_cd.used = d_cs.used = false;
  { const int rc_ = relu_chain_fwd(wpack, p.cd, 5, ref_cd_a0(cfg, points, normals, feats), b.H, p.ldh, EPI_SIGMOID, b.yd, 3, M, st, ar,
                 b.a0_cd, &d_cd); if (rc_) return rc_; }
  // viewdir_mlp: 4 x (Linear+ReLU); then net_cs Linear+Sigmoid
  {
    Lin chain[5] = {p.vd[0], p.vd[1], p.vd[2], p.vd[3], p.cs};
    // hidden layer outputs G1..G4 are all ReLU'd; the chain helper applies ReLU to all but the last linear.
    { const int rc_ = relu_chain_fwd(wpack, chain, 5, ref_cs_a0(cfg, points, normals, b.refl, feats), b.G, p.ldh, EPI_SIGMOID, b.ys,
                   1, M, st, ar, b.a0_cs, &d_cs); if (rc_) return rc_; }
  }
  if (d_cd.used && d_cs.used) relu_chain_pair_launch(d_cd.g, d_cs.g, d_cd.flops + d_cs.flops, st);
  else {
    if (d_cd.used) sdf_chain_launch(d_cd.g, d_cd.flops, st, FAM_RELU);
    if (d_cs.used) sdf_chain_launch(d_cs.g, d_cs.flops, st, FAM_RELU);
  }
  prof_begin(PC_ELEMENTWISE, 0.0, 0.0, st);
  ref_final_kernel<<<ew_blocks2(M * 3), 256, 0, st>>>(b.yd, b.ys, rgb_out, spec_out, diff_out, M);
  prof_end(st);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

int fneus_ref_bwd(const fneus_ref_cfg* cfg, const float* wpack, const float* points, const float* feats,
                  const float* dirs, const float* normals, long long M, const float* d_rgb, const float* d_spec,
                  const float* d_diff, float* d_feats, float* d_normals, float* saved, float* scratch,
                  float* d_wpack, void* stream) {
  PrecScope prec_scope_(cfg ? cfg->precision : 0);
  RefPlan p = ref_plan(cfg);
  if (!p.ok) return FNEUS_ERR_UNSUPPORTED;
  if (M == 0) return FNEUS_OK;
  if (!wpack || !points || !feats || !dirs || !normals || !d_feats || !d_normals || !saved || !scratch || !d_wpack)
    return FNEUS_ERR_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  RefBufs b = ref_carve(p, saved, M);
  const long long hf = hid_floats(M, p.hid, p.img);
  float* ab0 = align1k(scratch);
  float* ab1 = ab0 + 4 * hf;
  float* ds_cd = ab1 + 4 * hf;
  zero_if_ragged(p.img, ab0, 8 * hf, M, st);
  float* ds_cs = ds_cd + M * 32;
  float* a_cd = ds_cs + M * 36;
  float* a_cs = a_cd + M * 4;
  float* df_cs = align1k(a_cs + M * 4);                 // [M, d_feature]: the specular chain's feature gradient
  prof_begin(PC_ELEMENTWISE, 0.0, 0.0, st);
  ref_final_bwd_kernel<<<ew_blocks2(M), 256, 0, st>>>(b.yd, b.ys, d_rgb, d_spec, d_diff, a_cd, a_cs, M);
  prof_end(st);
  Lin chain[5] = {p.vd[0], p.vd[1], p.vd[2], p.vd[3], p.cs};
  ImgArena ar = arena_at(scratch + ref_scratch_main(p, M),
                         relu_chain_img_bytes(p.cd, 5, 30) + relu_chain_img_bytes(chain, 5, 33));
  // The two chains are independent: on the fused path both backward-data chains go out in ONE launch (own dz buffers,
  // own feature-gradient buffer, summed below) and all ten weight-gradient GEMMs in ONE grouped launch.
  ChainDefer d_cd, d_cs;
  d_cd.used = d_cs.used = false;
  const bool pair = cfg->d_feature <= 256 && (cfg->d_feature & 3) == 0;
  // 2 chains x (4 hidden layers + layer 0 split into its feature and generated parts) = 12 jobs: the shared group must
  // not flush by itself before the deferred chains have been launched
  static_assert(WG_MAX_JOBS >= 12, "the paired RefColor backward needs all weight-gradient jobs in one group");
  WgradGroup wg;
  wg.reset(M, num_sms());
  { const int rc_ = relu_chain_bwd(wpack, d_wpack, p.cd, 5, ref_cd_a0(cfg, points, normals, feats), b.H, p.ldh, a_cd, 4, ab0, hf,
                 ds_cd, 32, d_feats, cfg->d_feature, 0, M, st, ar, b.a0_cd, pair ? &d_cd : nullptr, pair ? &wg : nullptr); if (rc_) return rc_; }
  const bool paired = pair && d_cd.used;              // the diffuse chain took the fused path
  { const int rc_ = relu_chain_bwd(wpack, d_wpack, chain, 5, ref_cs_a0(cfg, points, normals, b.refl, feats), b.G, p.ldh, a_cs, 4,
                 paired ? ab1 : ab0, hf, ds_cs, 36, paired ? df_cs : d_feats, cfg->d_feature, paired ? 0 : 1, M, st, ar,
                 b.a0_cs, paired ? &d_cs : nullptr, paired ? &wg : nullptr); if (rc_) return rc_; }
  if (paired && d_cs.used) relu_chain_pair_launch(d_cd.g, d_cs.g, d_cd.flops + d_cs.flops, st);
  else if (d_cd.used) sdf_chain_launch(d_cd.g, d_cd.flops, st, FAM_RELU);
  wg.flush(st);
  prof_begin(PC_ELEMENTWISE, 0.0, 0.0, st);
  if (paired && d_cs.used)
    add_into_kernel<<<ew_blocks2(M * cfg->d_feature), 256, 0, st>>>(d_feats, df_cs, M * cfg->d_feature);
  ref_dn_kernel<<<ew_blocks2(M), 256, 0, st>>>(dirs, normals, b.refl, ds_cd, 32, ds_cs, 36, d_normals, M);
  prof_end(st);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// Generic positional-encoded ReLU MLP: Lvis (fields.py:338-369: [PE10(pts), PE4(view)] -> 256x4 -> 1, sigmoid) and
// IndirectLight's trunk (fields.py:372-399: PE10(pts) -> 512x4 -> 144, linear).  No input gradients.
// ------------------------------------------------------------------------------------------------
namespace fneus {
struct MlpPlan { int n_lin; Lin lin[12]; long long pack; int hid; int ldh; bool img; int gen_cols; bool ok; };
static MlpPlan mlp_plan(const fneus_mlp_cfg* c) {
  MlpPlan p;
  p.ok = c && c->n_inputs >= 1 && c->n_inputs <= 2 && c->n_layers >= 1 && c->n_layers <= 10 && c->d_hidden > 0 &&
         c->d_hidden % 4 == 0 && c->d_out >= 1 && c->d_out <= 1024 && c->d_hidden <= 1024;
  if (!p.ok) return p;
  p.gen_cols = 0;
  for (int i = 0; i < c->n_inputs; i++) {
    if (c->in_dim[i] < 1 || c->in_dim[i] > 4 || c->in_multires[i] < 0 || c->in_multires[i] > 12) { p.ok = false; return p; }
    p.gen_cols += pe_dim(c->in_dim[i], c->in_multires[i]);
  }
  p.n_lin = c->n_layers + 1;
  p.hid = c->d_hidden; p.img = precision_mode() == 1; p.ldh = mat_ld(p.hid, p.img);
  long long off = 0;
  for (int l = 0; l < p.n_lin; l++) {
    p.lin[l].in = l == 0 ? p.gen_cols : c->d_hidden;
    p.lin[l].out = l == p.n_lin - 1 ? c->d_out : c->d_hidden;
    p.lin[l].woff = off; off += (long long)p.lin[l].in * p.lin[l].out;
    p.lin[l].boff = off; off += p.lin[l].out;
  }
  p.pack = off;
  return p;
}
static ASeg mlp_a0(const fneus_mlp_cfg* c, const float* in0, const float* in1) {
  GenSpec g = gen_none();
  gen_add(g, in0, c->in_dim[0], c->in_multires[0]);
  if (c->n_inputs > 1) gen_add(g, in1, c->in_dim[1], c->in_multires[1]);
  return aseg_gen(g);
}
static long long mlp_scratch_main(const fneus_mlp_cfg* c, const MlpPlan& p, long long n) {
  return (long long)(c->n_layers > 2 ? c->n_layers : 2) * hid_floats(n, p.hid, p.img) + n * (round_up(c->d_out, 4) + 4) +
         2048;
}
}  // namespace fneus

extern "C" {

long long fneus_mlp_pack_floats(const fneus_mlp_cfg* cfg) { MlpPlan p = mlp_plan(cfg); return p.ok ? p.pack : -1; }
long long fneus_mlp_saved_floats(const fneus_mlp_cfg* cfg, long long n) {
  PrecScope prec_scope_(cfg ? cfg->precision : 0);
  MlpPlan p = mlp_plan(cfg);
  return p.ok ? (long long)cfg->n_layers * hid_floats(n, p.hid, p.img) + (p.img ? a0_img_floats(n) : 0) + 1024 : -1;
}
long long fneus_mlp_scratch_floats(const fneus_mlp_cfg* cfg, long long n) {
  PrecScope prec_scope_(cfg ? cfg->precision : 0);
  MlpPlan p = mlp_plan(cfg);
  return p.ok ? mlp_scratch_main(cfg, p, n) + (long long)(relu_chain_img_bytes(p.lin, p.n_lin, p.gen_cols) / 4) + 256 : -1;
}

int fneus_mlp_fwd(const fneus_mlp_cfg* cfg, const float* wpack, const float* in0, const float* in1, long long M,
                  float* out, float* saved, float* scratch, void* stream) {
  PrecScope prec_scope_(cfg ? cfg->precision : 0);
  MlpPlan p = mlp_plan(cfg);
  if (!p.ok) return FNEUS_ERR_UNSUPPORTED;
  if (M == 0) return FNEUS_OK;
  if (!wpack || !in0 || (cfg->n_inputs > 1 && !in1) || !out || !saved || !scratch) return FNEUS_ERR_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  float* base = align1k(saved);
  float* Hs[12];
  const long long hf = hid_floats(M, p.hid, p.img);
  for (int l = 1; l <= cfg->n_layers; l++) Hs[l] = base + (long long)(l - 1) * hf;
  zero_if_ragged(p.img, base, (long long)cfg->n_layers * hf, M, st);
  ImgArena ar = arena_at(scratch + mlp_scratch_main(cfg, p, M), relu_chain_img_bytes(p.lin, p.n_lin, p.gen_cols));
  { const int rc_ = relu_chain_fwd(wpack, p.lin, p.n_lin, mlp_a0(cfg, in0, in1), Hs, p.ldh, cfg->last_act == 1 ? EPI_SIGMOID : EPI_LINEAR,
                 out, cfg->d_out, M, st, ar, p.img ? base + (long long)cfg->n_layers * hf : nullptr); if (rc_) return rc_; }
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

int fneus_mlp_bwd(const fneus_mlp_cfg* cfg, const float* wpack, const float* in0, const float* in1, long long M,
                  const float* out, const float* d_out, float* saved, float* scratch, float* d_wpack, void* stream) {
  PrecScope prec_scope_(cfg ? cfg->precision : 0);
  MlpPlan p = mlp_plan(cfg);
  if (!p.ok) return FNEUS_ERR_UNSUPPORTED;
  if (M == 0) return FNEUS_OK;
  if (!wpack || !in0 || (cfg->n_inputs > 1 && !in1) || !out || !d_out || !saved || !scratch || !d_wpack)
    return FNEUS_ERR_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  const long long hf = hid_floats(M, p.hid, p.img);
  float* Hs[12];
  for (int l = 1; l <= cfg->n_layers; l++) Hs[l] = align1k(saved) + (long long)(l - 1) * hf;
  const int nbuf = cfg->n_layers > 2 ? cfg->n_layers : 2;
  float* ab0 = align1k(scratch);
  float* alast = ab0 + nbuf * hf;
  const int ldl = round_up(cfg->d_out, 4);
  zero_if_ragged(p.img, ab0, nbuf * hf, M, st);
  prof_begin(PC_ELEMENTWISE, 0.0, 0.0, st);
  if (cfg->last_act == 1) sigmoid_bwd_kernel<<<ew_blocks2(M * ldl), 256, 0, st>>>(d_out, out, cfg->d_out, alast, ldl, M);
  else pad_rows_kernel<<<ew_blocks2(M * ldl), 256, 0, st>>>(d_out, cfg->d_out, alast, ldl, M);
  prof_end(st);
  ImgArena ar = arena_at(scratch + mlp_scratch_main(cfg, p, M), relu_chain_img_bytes(p.lin, p.n_lin, p.gen_cols));
  { const int rc_ = relu_chain_bwd(wpack, d_wpack, p.lin, p.n_lin, mlp_a0(cfg, in0, in1), Hs, p.ldh, alast, ldl, ab0, hf, nullptr, 4,
                 nullptr, 4, 0, M, st, ar, p.img ? align1k(saved) + (long long)cfg->n_layers * hf : nullptr); if (rc_) return rc_; }
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

}  // extern "C"

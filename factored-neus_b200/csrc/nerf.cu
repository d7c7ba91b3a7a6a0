// Outside NeRF (fields.py:178-259) and render_core_outside (renderer.py:112-149) for the womask configuration.
// FP32 path on the SIMT GEMM engine.  Pack order (effective = plain weights, NeRF has no weight-norm):
//   pts_linears.0..D-1, head = [alpha_linear ; feature_linear] (W [1+W, W] then b [1+W]), views_linears.0,
//   rgb_linear.
#include "gemm_tc.cuh"
#include "prof.cuh"
#include "sdf_chain.cuh"
#include <string.h>

namespace fneus {

int num_sms();
__global__ void colsum_kernel(const float*, int, int, const float*, float, float*, float*, long long, int);
void launch_colsum(const float* X, int ldx, int K, const float* w, float wscale, float* out, float* osum, long long M,
                   cudaStream_t st, int f16);

struct NerfPlan {
  int D, W, e_p, e_v, skip;   // skip: index i after whose output the embedded input is concatenated (4)
  Lin pts[16]; Lin head, views, rgb;
  long long pack; bool ok;
};
static NerfPlan nerf_plan(const fneus_nerf_cfg* c) {
  NerfPlan p;
  p.ok = c && c->D >= 2 && c->D <= 16 && c->W >= 8 && c->W % 8 == 0 && c->d_in >= 1 && c->d_in <= 4 &&
         c->d_in_view == 3 && c->multires >= 0 && c->multires <= 12 && c->multires_view >= 0 &&
         c->multires_view <= 8 && c->skip >= -1 && c->skip < c->D - 1;
  if (!p.ok) return p;
  p.D = c->D; p.W = c->W; p.skip = c->skip;
  p.e_p = pe_dim(c->d_in, c->multires);
  p.e_v = pe_dim(3, c->multires_view);
  if (p.e_p > 96) { p.ok = false; return p; }
  long long off = 0;
  auto put = [&](Lin& l, int in, int out) {
    l.in = in; l.out = out; l.woff = off; off += (long long)in * out; l.boff = off; off += out;
  };
  put(p.pts[0], p.e_p, p.W);
  for (int i = 1; i < p.D; i++) put(p.pts[i], (i - 1 == p.skip) ? p.W + p.e_p : p.W, p.W);
  put(p.head, p.W, 1 + p.W);
  put(p.views, p.W + p.e_v, p.W / 2);
  put(p.rgb, p.W / 2, 3);
  p.pack = off;
  return p;
}

// saved: H_1..H_D [M,W], feature [M,W], V [M,W/2]
static long long nerf_saved_per_point(const NerfPlan& p) { return (long long)(p.D + 1) * p.W + p.W / 2; }
// scratch: 2 abufs [M,W], d_feature [M,W], a_V [M,W/2], a_rgb [M,4]
static long long nerf_scratch_per_point(const NerfPlan& p) { return 3LL * p.W + p.W / 2 + 4; }

// inverted-sphere reparametrisation + section geometry (renderer.py:116-128)
__global__ void outside_geometry_kernel(const float* __restrict__ o, const float* __restrict__ d,
                                        const float* __restrict__ z, long long total, int n, float sample_dist,
                                        float* __restrict__ dists, float* __restrict__ pts4, float* __restrict__ dirs) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  long long b = idx / n;
  int j = (int)(idx - b * n);
  float zj = z[idx];
  float dist = j + 1 < n ? z[idx + 1] - zj : sample_dist;
  float mid = zj + dist * 0.5f;
  dists[idx] = dist;
  float p[3];
#pragma unroll
  for (int c = 0; c < 3; c++) p[c] = o[b * 3 + c] + d[b * 3 + c] * mid;
  float r = sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
  r = fminf(fmaxf(r, 1.0f), 1e10f);
#pragma unroll
  for (int c = 0; c < 3; c++) { pts4[idx * 4 + c] = p[c] / r; dirs[idx * 3 + c] = d[b * 3 + c]; }
  pts4[idx * 4 + 3] = 1.0f / r;
}

// alpha = 1 - exp(-softplus(density) * dists) ; color = sigmoid(rgb)   (renderer.py:131-134)
__global__ void outside_alpha_fwd_kernel(const float* __restrict__ dens, const float* __restrict__ rgb,
                                         const float* __restrict__ dists, long long total, float* __restrict__ alpha,
                                         float* __restrict__ color) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  float x = dens[idx];
  float sp = x > 20.f ? x : log1pf(expf(x));
  alpha[idx] = 1.0f - expf(-sp * dists[idx]);
#pragma unroll
  for (int c = 0; c < 3; c++) color[idx * 3 + c] = sigmoidf_(rgb[idx * 3 + c]);
}
__global__ void outside_alpha_bwd_kernel(const float* __restrict__ dens, const float* __restrict__ color,
                                         const float* __restrict__ dists, const float* __restrict__ d_alpha,
                                         const float* __restrict__ d_color, long long total, float* __restrict__ d_dens,
                                         float* __restrict__ d_rgb) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  float x = dens[idx], dl = dists[idx];
  float sp = x > 20.f ? x : log1pf(expf(x));
  float dsp = x > 20.f ? 1.f : sigmoidf_(x);
  d_dens[idx] = (d_alpha ? d_alpha[idx] : 0.f) * expf(-sp * dl) * dl * dsp;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    float y = color[idx * 3 + c];
    d_rgb[idx * 3 + c] = (d_color ? d_color[idx * 3 + c] : 0.f) * y * (1.f - y);
  }
}
// pad [M,3] -> [M,4]
__global__ void pad3to4_kernel(const float* __restrict__ in, float* __restrict__ out, long long M) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * 4) return;
  long long m = idx / 4; int j = (int)(idx - m * 4);
  out[idx] = j < 3 ? in[m * 3 + j] : 0.f;
}

// ------------------------------------------------------------------------------------------------
// Tensor-core path: the whole NeRF as ONE fused chain per pass (sdf_chain.cuh, ReLU family).
// PE(pts) (<= 2 blocks) and PE(views) (1 block) are generated once per tile into the auxiliary area and stay there:
// layer 0 reads PE(pts) as its operand, the skip layer reads [hidden (4 blocks) | PE(pts)], views_linears reads
// [feature (4 blocks) | PE(views)] -- the weight images put the matching column ranges of W in that block order.
// alpha_linear is a 1-wide output step off the last hidden state, which then stays in place for feature_linear.
// Forward operands / images FP16, backward BF16 (fneus_common.cuh).
// ------------------------------------------------------------------------------------------------
static bool nerf_chain_ok(const NerfPlan& p) {
  if (precision_mode() != 1 || tc_prepare() != 0 || sdf_chain_prepare() != 0) return false;
  if (tc_debug_flags() & 32) return false;                        // debug: layered execution
  return p.W == 256 && p.D >= 2 && p.D + 4 <= SC_MAXS && p.D + 4 <= sc_bias_slots<FAM_RELU>() && p.e_p <= 128 &&
         p.e_v <= 64 && cdiv(p.e_p, TC_BK) + 1 <= SC_AUXGEN_MAX;
}
static inline long long nerf_hf(long long M) { return mat_floats(M, 256, true) + 256; }
static inline float* nerf_align(float* q) {
  return reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(q) + 1023) & ~(uintptr_t)1023);
}
// weight images of one pass: every layer in both forms is < 64 tiles of 32 KB
static constexpr long long NERF_IMG_FLOATS = (64LL * TC_B_BYTES + 65536) / 4;
struct NerfImgs { float* H[20]; float* feat; float* V; float* pe; float* arena; };
static NerfImgs nerf_carve(const NerfPlan& p, float* saved, long long M) {
  NerfImgs b;
  float* ptr = nerf_align(saved);
  const long long hf = nerf_hf(M);
  for (int i = 1; i <= p.D; i++) { b.H[i] = ptr; ptr += hf; }
  b.feat = ptr; ptr += hf;
  b.V = ptr; ptr += hf;
  b.pe = ptr; ptr += mat_floats(M, SC_AUXGEN_MAX * TC_BK, true) + 256;
  b.arena = ptr;
  return b;
}
static long long nerf_saved_img_floats(const NerfPlan& p, long long M) {
  return (p.D + 2) * nerf_hf(M) + mat_floats(M, SC_AUXGEN_MAX * TC_BK, true) + 256 + NERF_IMG_FLOATS + 1024;
}
static long long nerf_scratch_img_floats(const NerfPlan& p, long long M) {
  return (p.D + 2) * nerf_hf(M) + 4 * M + 256 + NERF_IMG_FLOATS + 1024;
}
// [W[:, c0 : c0 + k0] | W[:, c1 : c1 + k1]] as consecutive K-major tiles (the chain's weight loader walks one image)
static const uint8_t* nerf_wimg2(ImgArena& ar, const float* W, int ldw, int N, int c0, int k0, int c1, int k1,
                                 cudaStream_t st) {
  const uint8_t* a = make_wimg(ar, false, W, ldw, 0, N, 0, 0, c0, k0, st, 1);
  const uint8_t* b = make_wimg(ar, false, W, ldw, 0, N, 0, 0, c1, k1, st, 1);
  if (!a || !b || b != a + wimg_bytes(N, 0, k0)) return nullptr;
  return a;
}
static GenSpec nerf_gen_aux(const fneus_nerf_cfg* c, const NerfPlan& p, const float* pts, const float* views) {
  GenSpec g = gen_none();
  gen_add(g, pts, c->d_in, c->multires);
  gen_add(g, views, 3, c->multires_view);
  g.it[1].col0 = cdiv(p.e_p, TC_BK) * TC_BK;                      // PE(views) starts its own block
  g.ncols = g.it[1].col0 + p.e_v;
  return g;
}

static int nerf_fwd_chain(const fneus_nerf_cfg* cfg, const NerfPlan& p, const float* w, const float* pts,
                          const float* views, long long M, float* density_out, float* rgb_out, float* saved,
                          cudaStream_t st) {
  NerfImgs b = nerf_carve(p, saved, M);
  ImgArena ar = arena_make(reinterpret_cast<uint8_t*>(nerf_align(b.arena)), (size_t)(NERF_IMG_FLOATS - 512) * 4);
  const int W = p.W, kbp = cdiv(p.e_p, TC_BK);
  SdfChainArgs g;
  memset(&g, 0, sizeof(g));
  double flops = 0.0;
  int ns = 0, slot = 0;
  for (int i = 0; i < p.D; i++) {
    const float* Wi = w + p.pts[i].woff;
    const bool skip = i > 0 && i - 1 == p.skip;
    const uint8_t* img = i == 0 ? make_wimg(ar, false, Wi, p.e_p, 0, W, 0, 0, 0, p.e_p, st, 1)
                       : skip   ? nerf_wimg2(ar, Wi, W + p.e_p, W, p.e_p, W, 0, p.e_p, st)
                                : make_wimg(ar, false, Wi, W, 0, W, 0, 0, 0, W, st, 1);
    if (!img) return FNEUS_ERR_WORKSPACE;
    SdfStep S = sdf_step(SC_RELU, img, i == 0 ? kbp : (skip ? 4 + kbp : 4), W, 0);
    if (i == 0) { S.src = SRC_AUXGEN; S.kb_op = 0; S.aux_blk0 = 0; }
    else if (skip) { S.kb_op = 4; S.aux_blk0 = 0; }
    S.bias = w + p.pts[i].boff; S.bias_slot = slot++;
    S.img_out = b.H[i + 1];
    g.st[ns++] = S;
    flops += 2.0 * (double)M * p.pts[i].in * W;
  }
  {
    const uint8_t* img = make_wimg(ar, false, w + p.head.woff, W, 0, 1, 0, 0, 0, W, st, 1);
    if (!img) return FNEUS_ERR_WORKSPACE;
    SdfStep S = sdf_step(SC_OUT, img, 4, 1, 0);                    // alpha_linear: raw density, the operand stays
    S.bias = w + p.head.boff; S.bias_slot = slot++;
    S.out = density_out; S.ldo = 1;
    g.st[ns++] = S;
    flops += 2.0 * (double)M * W;
  }
  {
    const uint8_t* img = make_wimg(ar, false, w + p.head.woff, W, 1, W, 0, 0, 0, W, st, 1);
    if (!img) return FNEUS_ERR_WORKSPACE;
    SdfStep S = sdf_step(SC_RELU, img, 4, W, 0);                   // feature_linear: no activation
    S.act = 2;
    S.bias = w + p.head.boff + 1; S.bias_slot = slot++;
    S.img_out = b.feat;
    g.st[ns++] = S;
    flops += 2.0 * (double)M * W * W;
  }
  {
    const uint8_t* img = nerf_wimg2(ar, w + p.views.woff, W + p.e_v, W / 2, 0, W, W, p.e_v, st);
    if (!img) return FNEUS_ERR_WORKSPACE;
    SdfStep S = sdf_step(SC_RELU, img, 5, W / 2, 0);               // views_linears.0 on [feature | PE(views)]
    S.kb_op = 4; S.aux_blk0 = kbp;
    S.bias = w + p.views.boff; S.bias_slot = slot++;
    S.img_out = b.V;
    g.st[ns++] = S;
    flops += 2.0 * (double)M * (W + p.e_v) * (W / 2);
  }
  {
    const uint8_t* img = make_wimg(ar, false, w + p.rgb.woff, W / 2, 0, 3, 0, 0, 0, W / 2, st, 1);
    if (!img) return FNEUS_ERR_WORKSPACE;
    SdfStep S = sdf_step(SC_OUT, img, cdiv(W / 2, TC_BK), 3, 0);   // rgb_linear: raw colour
    S.bias = w + p.rgb.boff; S.bias_slot = slot++;
    S.out = rgb_out; S.ldo = 3;
    g.st[ns++] = S;
    flops += 2.0 * (double)M * (W / 2) * 3;
  }
  ar.flush(st);
  g.nsteps = ns;
  g.gen = nerf_gen_aux(cfg, p, pts, views); g.gen_t = g.gen;
  g.aux_gen_blocks = kbp + 1;
  g.a0_img = b.pe;
  g.f16 = 1;
  g.beta = 1.f; g.M = M; g.dbg = 0;
  g.xflags = (tc_debug_flags() >> 8) & 15;
  sdf_chain_launch(g, flops, st, FAM_RELU);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

static int nerf_bwd_chain(const fneus_nerf_cfg* cfg, const NerfPlan& p, const float* w, long long M,
                          const float* d_density, const float* d_rgb, float* saved, float* scratch, float* dw,
                          cudaStream_t st) {
  NerfImgs b = nerf_carve(p, saved, M);
  const int W = p.W, kbp = cdiv(p.e_p, TC_BK), D = p.D;
  const long long hf = nerf_hf(M);
  float* base = nerf_align(scratch);
  float* dz[20];
  for (int i = 0; i < D; i++) dz[i] = base + (long long)i * hf;   // gradient wrt layer i's pre-activation
  float* dfeat = base + (long long)D * hf;
  float* aV = dfeat + hf;
  float* argb = aV + hf;
  ImgArena ar = arena_make(reinterpret_cast<uint8_t*>(nerf_align(argb + 4 * M + 256)), (size_t)(NERF_IMG_FLOATS - 512) * 4);
  prof_begin(PC_ELEMENTWISE, 0.0, 0.0, st);
  pad3to4_kernel<<<cdiv(M * 4, 256), 256, 0, st>>>(d_rgb, argb, M);
  prof_end(st);
  SdfChainArgs g;
  memset(&g, 0, sizeof(g));
  double flops = 0.0;
  int ns = 0;
  {
    const uint8_t* img = make_wimg(ar, true, w + p.rgb.woff, W / 2, 0, W / 2, 0, 0, 0, 3, st);
    if (!img) return FNEUS_ERR_WORKSPACE;
    SdfStep S = sdf_step(SC_MASK, img, 1, W / 2, 1);               // a_V = (d_rgb W_rgb) * [V > 0]
    S.src = SRC_MEM; S.h = b.V; S.img_out = aV;
    g.st[ns++] = S;
    flops += 2.0 * (double)M * 3 * (W / 2);
  }
  {
    const uint8_t* img = make_wimg(ar, true, w + p.views.woff, W + p.e_v, 0, W, 0, 0, 0, W / 2, st);
    if (!img) return FNEUS_ERR_WORKSPACE;
    SdfStep S = sdf_step(SC_MASK, img, cdiv(W / 2, TC_BK), W, 1);  // d_feature = a_V W_views[:, :W] (no activation)
    S.img_out = dfeat;
    g.st[ns++] = S;
    flops += 2.0 * (double)M * (W / 2) * W;
  }
  {
    const uint8_t* img = make_wimg(ar, true, w + p.head.woff, W, 0, W, 0, 0, 1, W, st);
    if (!img) return FNEUS_ERR_WORKSPACE;
    SdfStep S = sdf_step(SC_MASK, img, 4, W, 1);                   // dz_{D-1} = (d_feature W_feat + d_density w_alpha) * [h_D > 0]
    S.h = b.H[D]; S.use_rs = 1; S.img_out = dz[D - 1];
    g.st[ns++] = S;
    flops += 2.0 * (double)M * W * (W + 1);
  }
  for (int i = D - 1; i >= 1; i--) {
    const bool skip = i - 1 == p.skip;
    const uint8_t* img = make_wimg(ar, true, w + p.pts[i].woff, p.pts[i].in, skip ? p.e_p : 0, W, 0, 0, 0, W, st);
    if (!img) return FNEUS_ERR_WORKSPACE;
    SdfStep S = sdf_step(SC_MASK, img, 4, W, 1);                   // dz_{i-1} = (dz_i W_i[:, hidden part]) * [h_i > 0]
    S.h = b.H[i]; S.img_out = dz[i - 1];
    g.st[ns++] = S;
    flops += 2.0 * (double)M * W * W;
  }
  ar.flush(st);
  g.nsteps = ns;
  g.gen = gen_none(); g.gen_t = g.gen;
  g.mem = argb; g.ldm = 4; g.kmem = 3;
  g.rvec = w + p.head.woff; g.rs = d_density; g.rscale = 1.f;
  g.f16 = 0;
  g.beta = 1.f; g.M = M; g.dbg = 0;
  g.xflags = (tc_debug_flags() >> 8) & 15;
  sdf_chain_launch(g, flops, st, FAM_RELU);
  // ---- weight gradients, one grouped launch: dz / d_feature / a_V are BF16 images of this pass, the activations FP16 ----
  WgradGroup wg;
  wg.reset(M, num_sms());
  const int pe_ld = -(kbp + 1);
  wg.add(argb, 4, aseg_mem(b.V, -4, W / 2), dw + p.rgb.woff, W / 2, 0, dw + p.rgb.boff, 3, st, 0, 1);
  wg.add(aV, -4, aseg_mem(b.feat, -4, W, 0), dw + p.views.woff, W + p.e_v, 0, dw + p.views.boff, W / 2, st, 0, 1);
  wg.add(aV, -4, aseg_mem(b.pe + (size_t)kbp * (TC_A_BYTES / 4), pe_ld, p.e_v, W), dw + p.views.woff, W + p.e_v, 0, nullptr,
         W / 2, st, 0, 1);
  wg.add(dfeat, -4, aseg_mem(b.H[D], -4, W), dw + p.head.woff, W, 1, dw + p.head.boff, W, st, 0, 1);
  for (int i = D - 1; i >= 0; i--) {
    float* dW = dw + p.pts[i].woff;
    float* db = dw + p.pts[i].boff;
    if (i == 0) wg.add(dz[0], -4, aseg_mem(b.pe, pe_ld, p.e_p, 0), dW, p.e_p, 0, db, W, st, 0, 1);
    else if (i - 1 == p.skip) {
      wg.add(dz[i], -4, aseg_mem(b.H[i], -4, W, p.e_p), dW, W + p.e_p, 0, db, W, st, 0, 1);
      wg.add(dz[i], -4, aseg_mem(b.pe, pe_ld, p.e_p, 0), dW, W + p.e_p, 0, nullptr, W, st, 0, 1);
    } else wg.add(dz[i], -4, aseg_mem(b.H[i], -4, W), dW, W, 0, db, W, st, 0, 1);
  }
  wg.flush(st);
  // alpha_linear: dW[0, :] += sum_m d_density[m] h_D[m, :], db[0] += sum_m d_density[m]
  launch_colsum(b.H[D], -4, W, d_density, 1.f, dw + p.head.woff, dw + p.head.boff, M, st, 1);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

static GenSpec nerf_gen_pts(const fneus_nerf_cfg* c, const float* pts) {
  GenSpec g = gen_none(); gen_add(g, pts, c->d_in, c->multires); return g;
}
static GenSpec nerf_gen_views(const fneus_nerf_cfg* c, const float* views) {
  GenSpec g = gen_none(); gen_add(g, views, 3, c->multires_view); return g;
}

}  // namespace fneus

using namespace fneus;

extern "C" {

long long fneus_nerf_pack_floats(const fneus_nerf_cfg* cfg) { NerfPlan p = nerf_plan(cfg); return p.ok ? p.pack : -1; }
long long fneus_nerf_saved_floats(const fneus_nerf_cfg* cfg, long long n) {
  PrecScope prec_scope_(cfg ? cfg->precision : 0);
  NerfPlan p = nerf_plan(cfg);
  if (!p.ok) return -1;
  return nerf_chain_ok(p) ? nerf_saved_img_floats(p, n) : nerf_saved_per_point(p) * n;
}
long long fneus_nerf_scratch_floats(const fneus_nerf_cfg* cfg, long long n) {
  PrecScope prec_scope_(cfg ? cfg->precision : 0);
  NerfPlan p = nerf_plan(cfg);
  if (!p.ok) return -1;
  return nerf_chain_ok(p) ? nerf_scratch_img_floats(p, n) : nerf_scratch_per_point(p) * n;
}

int fneus_nerf_fwd(const fneus_nerf_cfg* cfg, const float* wpack, const float* pts, const float* views, long long M,
                   float* density_out, float* rgb_out, float* saved, void* stream) {
  PrecScope prec_scope_(cfg ? cfg->precision : 0);
  NerfPlan p = nerf_plan(cfg);
  if (!p.ok) return FNEUS_ERR_UNSUPPORTED;
  if (M == 0) return FNEUS_OK;
  if (!wpack || !pts || !views || !density_out || !rgb_out || !saved) return FNEUS_ERR_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  if (nerf_chain_ok(p)) return nerf_fwd_chain(cfg, p, wpack, pts, views, M, density_out, rgb_out, saved, st);
  const int W = p.W;
  float* H[20];
  for (int i = 1; i <= p.D; i++) H[i] = saved + (long long)(i - 1) * M * W;
  float* feat = saved + (long long)p.D * M * W;
  float* V = feat + M * W;
  for (int i = 0; i < p.D; i++) {
    ASeg a;
    if (i == 0) a = aseg_gen(nerf_gen_pts(cfg, pts));
    else if (i - 1 == p.skip) a = aseg_gen_mem(nerf_gen_pts(cfg, pts), 0, H[i], W, W, p.e_p);
    else a = aseg_mem(H[i], W, W);
    Epi e = epi_default();
    e.mode = EPI_RELU; e.bias = wpack + p.pts[i].boff; e.C = H[i + 1]; e.ldc = W;
    launch_gemm_fwd(a, wpack + p.pts[i].woff, p.pts[i].in, 0, M, W, e, st);
  }
  {
    Epi e = epi_default();
    e.mode = EPI_SDF_OUT; e.bias = wpack + p.head.boff; e.out0 = density_out; e.out0_scale = 1.f;
    e.C = feat; e.ldc = W;
    launch_gemm_fwd(aseg_mem(H[p.D], W, W), wpack + p.head.woff, W, 0, M, 1 + W, e, st);
  }
  {
    Epi e = epi_default();
    e.mode = EPI_RELU; e.bias = wpack + p.views.boff; e.C = V; e.ldc = W / 2;
    launch_gemm_fwd(aseg_gen_mem(nerf_gen_views(cfg, views), W, feat, W, W, 0), wpack + p.views.woff, p.views.in, 0, M,
                    W / 2, e, st);
  }
  {
    Epi e = epi_default();
    e.mode = EPI_LINEAR; e.bias = wpack + p.rgb.boff; e.C = rgb_out; e.ldc = 3;
    launch_gemm_fwd(aseg_mem(V, W / 2, W / 2), wpack + p.rgb.woff, W / 2, 0, M, 3, e, st);
  }
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

int fneus_nerf_bwd(const fneus_nerf_cfg* cfg, const float* wpack, const float* pts, const float* views, long long M,
                   const float* d_density, const float* d_rgb, float* saved, float* scratch, float* d_wpack,
                   void* stream) {
  PrecScope prec_scope_(cfg ? cfg->precision : 0);
  NerfPlan p = nerf_plan(cfg);
  if (!p.ok) return FNEUS_ERR_UNSUPPORTED;
  if (M == 0) return FNEUS_OK;
  if (!wpack || !pts || !views || !d_density || !d_rgb || !saved || !scratch || !d_wpack) return FNEUS_ERR_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  if (nerf_chain_ok(p)) return nerf_bwd_chain(cfg, p, wpack, M, d_density, d_rgb, saved, scratch, d_wpack, st);
  const int sms = num_sms();
  const int W = p.W;
  float* H[20];
  for (int i = 1; i <= p.D; i++) H[i] = saved + (long long)(i - 1) * M * W;
  float* feat = saved + (long long)p.D * M * W;
  float* V = feat + M * W;
  float* ab[2] = {scratch, scratch + M * W};
  float* dfeat = scratch + 2LL * M * W;
  float* aV = dfeat + M * W;
  float* argb = aV + M * (W / 2);
  prof_begin(PC_ELEMENTWISE, 0.0, 0.0, st);
  pad3to4_kernel<<<cdiv(M * 4, 256), 256, 0, st>>>(d_rgb, argb, M);
  prof_end(st);
  // rgb_linear
  launch_gemm_wgrad(argb, 4, aseg_mem(V, W / 2, W / 2), d_wpack + p.rgb.woff, W / 2, 0, d_wpack + p.rgb.boff, M, 3, sms, st);
  {
    Epi e = epi_default();
    e.mode = EPI_RELUMASK; e.H = V; e.ldh = W / 2; e.C = aV; e.ldc = W / 2;
    launch_gemm_bwd_data(aseg_mem(argb, 4, 3), wpack + p.rgb.woff, W / 2, 0, M, W / 2, e, st);
  }
  // views_linears.0 : input [feature | PE(views)]
  launch_gemm_wgrad(aV, W / 2, aseg_gen_mem(nerf_gen_views(cfg, views), W, feat, W, W, 0), d_wpack + p.views.woff,
                    p.views.in, 0, d_wpack + p.views.boff, M, W / 2, sms, st);
  {
    Epi e = epi_default();
    e.mode = EPI_RELUMASK; e.H = nullptr; e.C = dfeat; e.ldc = W;     // no activation on feature_linear's output
    launch_gemm_bwd_data(aseg_mem(aV, W / 2, W / 2), wpack + p.views.woff, p.views.in, 0, M, W, e, st);
  }
  // head = [alpha_linear ; feature_linear]
  {
    int mpb = 256;
    dim3 grid(cdiv(W, 256), cdiv(M, mpb));
    prof_begin(PC_ELEMENTWISE, 0.0, 0.0, st);
    colsum_kernel<<<grid, 256, 0, st>>>(H[p.D], W, W, d_density, 1.f, d_wpack + p.head.woff, d_wpack + p.head.boff, M, mpb);
    prof_end(st);
  }
  launch_gemm_wgrad(dfeat, W, aseg_mem(H[p.D], W, W), d_wpack + p.head.woff, W, 1, d_wpack + p.head.boff, M, W, sms, st);
  {
    Epi e = epi_default();
    e.mode = EPI_RELUMASK; e.H = H[p.D]; e.ldh = W; e.C = ab[p.D & 1]; e.ldc = W;
    e.rs = d_density; e.rvec = wpack + p.head.woff; e.rscale = 1.f;
    launch_gemm_bwd_data(aseg_mem(dfeat, W, W, /*wred=*/1), wpack + p.head.woff, W, 0, M, W, e, st);
  }
  // pts_linears D-1 .. 0 ; ab[(i+1)&1] holds the gradient wrt layer i's pre-activation
  for (int i = p.D - 1; i >= 0; i--) {
    const float* al = ab[(i + 1) & 1];
    ASeg h;
    if (i == 0) h = aseg_gen(nerf_gen_pts(cfg, pts));
    else if (i - 1 == p.skip) h = aseg_gen_mem(nerf_gen_pts(cfg, pts), 0, H[i], W, W, p.e_p);
    else h = aseg_mem(H[i], W, W);
    launch_gemm_wgrad(al, W, h, d_wpack + p.pts[i].woff, p.pts[i].in, 0, d_wpack + p.pts[i].boff, M, W, sms, st);
    if (i == 0) break;
    Epi e = epi_default();
    e.mode = EPI_RELUMASK;
    if (i - 1 == p.skip) {   // input = [PE(84) | h(256)]: only the h block carries gradient
      e.csplit = p.e_p; e.C = nullptr; e.C2 = ab[i & 1]; e.ldc2 = W; e.H2 = H[i]; e.ldh2 = W;
    } else {
      e.H = H[i]; e.ldh = W; e.C = ab[i & 1]; e.ldc = W;
    }
    launch_gemm_bwd_data(aseg_mem(al, W, W), wpack + p.pts[i].woff, p.pts[i].in, 0, M, p.pts[i].in, e, st);
  }
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

int fneus_outside_geometry(const float* rays_o, const float* rays_d, const float* z, long long B, int n,
                           float sample_dist, float* dists, float* pts4, float* dirs, void* stream) {
  if (B == 0 || n == 0) return FNEUS_OK;
  if (!rays_o || !rays_d || !z || !dists || !pts4 || !dirs) return FNEUS_ERR_NULL;
  long long total = B * n;
  prof_begin(PC_SAMPLING, 0.0, (double)total * 36.0, (cudaStream_t)stream);
  outside_geometry_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, z, total, n, sample_dist,
                                                                             dists, pts4, dirs);
  prof_end((cudaStream_t)stream);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

int fneus_outside_alpha_fwd(const float* density, const float* rgb_raw, const float* dists, long long total,
                            float* alpha, float* color, void* stream) {
  if (total == 0) return FNEUS_OK;
  if (!density || !rgb_raw || !dists || !alpha || !color) return FNEUS_ERR_NULL;
  prof_begin(PC_COMPOSITE, 0.0, (double)total * 36.0, (cudaStream_t)stream);
  outside_alpha_fwd_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(density, rgb_raw, dists, total, alpha, color);
  prof_end((cudaStream_t)stream);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

int fneus_outside_alpha_bwd(const float* density, const float* color, const float* dists, const float* d_alpha,
                            const float* d_color, long long total, float* d_density, float* d_rgb_raw, void* stream) {
  if (total == 0) return FNEUS_OK;
  if (!density || !color || !dists || !d_density || !d_rgb_raw) return FNEUS_ERR_NULL;
  prof_begin(PC_COMPOSITE, 0.0, (double)total * 52.0, (cudaStream_t)stream);
  outside_alpha_bwd_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(density, color, dists, d_alpha, d_color,
                                                                              total, d_density, d_rgb_raw);
  prof_end((cudaStream_t)stream);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

}  // extern "C"

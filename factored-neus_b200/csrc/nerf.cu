// Outside NeRF (fields.py:178-259) and render_core_outside (renderer.py:112-149) for the womask configuration.
// FP32 path on the SIMT GEMM engine.  Pack order (effective = plain weights, NeRF has no weight-norm):
//   pts_linears.0..D-1, head = [alpha_linear ; feature_linear] (W [1+W, W] then b [1+W]), views_linears.0,
//   rgb_linear.
#include "gemm_tc.cuh"
#include "prof.cuh"

namespace fneus {

int num_sms();
__global__ void colsum_kernel(const float*, int, int, const float*, float, float*, float*, long long, int);

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
  NerfPlan p = nerf_plan(cfg); return p.ok ? nerf_saved_per_point(p) * n : -1;
}
long long fneus_nerf_scratch_floats(const fneus_nerf_cfg* cfg, long long n) {
  NerfPlan p = nerf_plan(cfg); return p.ok ? nerf_scratch_per_point(p) * n : -1;
}

int fneus_nerf_fwd(const fneus_nerf_cfg* cfg, const float* wpack, const float* pts, const float* views, long long M,
                   float* density_out, float* rgb_out, float* saved, void* stream) {
  NerfPlan p = nerf_plan(cfg);
  if (!p.ok) return FNEUS_ERR_UNSUPPORTED;
  if (M == 0) return FNEUS_OK;
  if (!wpack || !pts || !views || !density_out || !rgb_out || !saved) return FNEUS_ERR_NULL;
  cudaStream_t st = (cudaStream_t)stream;
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
  NerfPlan p = nerf_plan(cfg);
  if (!p.ok) return FNEUS_ERR_UNSUPPORTED;
  if (M == 0) return FNEUS_OK;
  if (!wpack || !pts || !views || !d_density || !d_rgb || !saved || !scratch || !d_wpack) return FNEUS_ERR_NULL;
  cudaStream_t st = (cudaStream_t)stream;
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

// Stage-2 light-visibility trace (reference: models/calLvis.py:339-397, the ground-truth half of cal_indiLgt) as ONE
// C-ABI call: per chunk of secondary rays a fixed sequence of persistent kernels, no host synchronisation, no
// allocation (the caller's workspace is carved up here), CUDA-graph capturable.
//
//   coarse points on the shared depth table -> sdf-only chain (512 evaluations per ray: the hot spot)
//   -> inverse-CDF importance depths at the LEARNED inv_s (device scalar) -> section geometry of the 32 depths
//   -> value + normal chain -> alpha / weights (compute_weight, cos_anneal_ratio = 0)
//   -> visibility = 1 - sum of inside-sphere weights, first sign change + secant root
//   -> value + feature + normal chain at the root -> colour chain -> radiance masked by the hit
//
// A single monolithic kernel would not move fewer bytes that matter: the per-chunk intermediates (sdf of the coarse
// samples, 2 KB per ray) stay in L2, and every phase is already one persistent tile kernel.
#include "gemm_tc.cuh"
#include "prof.cuh"

namespace fneus {
__global__ void lvis_coarse_points_kernel(const float*, int, const float*, const float*, long long, int, float*, float*);
__global__ void mask_rows3_kernel(const float*, const int*, long long, float*);
int upsample_step_launch(const float* rays_o, const float* rays_d, const float* z, const float* sdf, long long B, int n,
                         int k, float inv_s, const float* inv_s_dev, const float* u_table, float* new_z, float* cdf_out,
                         long long* inds_out, void* stream, int z_shared);

struct LvisWs {
  long long o, pts_c, sdf_c, z_f, dists, mid, pts_f, dirs_f, sdf_f, feat, nrm, zero_rgb, color, weights, wsum, wmax, cdf,
      inside, eik, wpair, hit_tmp, p_s, sdf_s, n_s, rgb_s, saved, scratch, scratch_floats, total;
};
static inline long long up256(long long x) { return (x + 255) / 256 * 256; }
static LvisWs lvis_layout(const fneus_sdf_cfg* sc, const fneus_color_cfg* cc, long long R, int n_coarse, int n_imp) {
  LvisWs w;
  long long off = 256;
  auto take = [&](long long n) { long long o = off; off += up256(n); return o; };
  const long long Mf = R * n_imp, Mc = R * n_coarse;
  const int F = sc->d_out - 1;
  w.o = take(R * 3); w.pts_c = take(Mc * 3); w.sdf_c = take(Mc);
  w.z_f = take(Mf); w.dists = take(Mf); w.mid = take(Mf); w.pts_f = take(Mf * 3); w.dirs_f = take(Mf * 3);
  w.sdf_f = take(Mf); w.feat = take(Mf * F); w.nrm = take(Mf * 3); w.zero_rgb = take(Mf * 3);
  w.color = take(R * 3); w.weights = take(Mf); w.wsum = take(R); w.wmax = take(R); w.cdf = take(Mf); w.inside = take(Mf);
  w.eik = take(R * 2); w.wpair = take(R * 2); w.hit_tmp = take(R);
  w.p_s = take(R * 3); w.sdf_s = take(R); w.n_s = take(R * 3); w.rgb_s = take(R * 3);
  w.saved = take(fneus_sdf_saved_floats(sc, Mf));
  long long s1 = fneus_sdf_scratch_floats(sc, Mf), s2 = fneus_color_scratch_floats(cc, R);
  // the coarse sdf-only pass takes whatever scratch it is given (it chunks by itself); the value + normal chain needs s1
  w.scratch_floats = (s1 > s2 ? s1 : s2) + (1 << 20);
  w.scratch = take(w.scratch_floats);
  w.total = off + 256;
  return w;
}

}  // namespace fneus

using namespace fneus;

extern "C" {

long long fneus_lvis_trace_workspace_floats(const fneus_sdf_cfg* sdf_cfg, const fneus_color_cfg* color_cfg,
                                            long long rays_per_chunk, int n_coarse, int n_imp) {
  PrecScope prec_scope_(sdf_cfg ? sdf_cfg->precision : 0);
  if (!sdf_cfg || !color_cfg || rays_per_chunk < 1 || n_coarse < 2 || n_imp < 2) return -1;
  if (fneus_sdf_saved_floats(sdf_cfg, 1) < 0 || fneus_color_scratch_floats(color_cfg, 1) < 0) return -1;
  return lvis_layout(sdf_cfg, color_cfg, rays_per_chunk, n_coarse, n_imp).total;
}

int fneus_lvis_trace(const fneus_sdf_cfg* sdf_cfg, const float* sdf_wpack, const fneus_color_cfg* color_cfg,
                     const float* color_wpack, const float* surf, const float* dirs, long long m, int n_dirs,
                     int n_coarse, int n_imp, const float* inv_s, const float* z_table, const float* u_table,
                     float* lvis_out, float* rgb_out, int* hit_out, float* ws, long long ws_floats,
                     long long rays_per_chunk, void* stream) {
  PrecScope prec_scope_(sdf_cfg ? sdf_cfg->precision : 0);
  if (m == 0) return FNEUS_OK;
  if (!sdf_cfg || !sdf_wpack || !color_cfg || !color_wpack || !surf || !dirs || !inv_s || !z_table || !u_table ||
      !lvis_out || !rgb_out || !hit_out || !ws)
    return FNEUS_ERR_NULL;
  if (m < 0 || n_dirs < 1 || n_coarse < 2 || n_imp < 2 || n_imp > 2048 || rays_per_chunk < n_dirs) return FNEUS_ERR_BAD_SHAPE;
  if (sdf_cfg->d_in != 3 || color_cfg->d_feature != sdf_cfg->d_out - 1) return FNEUS_ERR_UNSUPPORTED;
  rays_per_chunk = rays_per_chunk / n_dirs * n_dirs;                 // whole surface points per chunk
  const LvisWs w = lvis_layout(sdf_cfg, color_cfg, rays_per_chunk, n_coarse, n_imp);
  if (ws_floats < w.total) return FNEUS_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(ws) & 15) return FNEUS_ERR_MISALIGNED;
  cudaStream_t st = (cudaStream_t)stream;
  const float sample_dist = (1.0f - 0.1f) / 32.0f;                   // calLvis.py:95,155 (a constant there, not (far-near)/n)
  const long long total_rays = m * n_dirs;
  cudaError_t ce = cudaMemsetAsync(ws + w.zero_rgb, 0, (size_t)rays_per_chunk * n_imp * 3 * sizeof(float), st);
  if (ce != cudaSuccess) return fneus_cuda_error((int)ce);
  for (long long r0 = 0; r0 < total_rays; r0 += rays_per_chunk) {
    const long long R = total_rays - r0 < rays_per_chunk ? total_rays - r0 : rays_per_chunk;
    const long long Mc = R * n_coarse, Mf = R * n_imp;
    const float* d = dirs + r0 * 3;
    float* o = ws + w.o;
    int rc;
    // coarse samples: points on the shared depth table, sdf-only chain (no_grad in the reference, calLvis.py:363-368)
    prof_begin(PC_SAMPLING, 0.0, (double)Mc * 12.0, st);
    lvis_coarse_points_kernel<<<cdiv(Mc, 256), 256, 0, st>>>(surf + (r0 / n_dirs) * 3, n_dirs, d, z_table, Mc, n_coarse,
                                                            ws + w.pts_c, o);
    prof_end(st);
    FNEUS_CHECK_LAUNCH();
    if ((rc = fneus_sdf_fwd(sdf_cfg, sdf_wpack, ws + w.pts_c, Mc, ws + w.sdf_c, nullptr, ws + w.scratch, w.scratch_floats, stream)))
      return rc;
    // 32 importance depths from the inverse CDF at the learned inv_s; ONLY these are kept (calLvis.py:374-379)
    if ((rc = upsample_step_launch(o, d, z_table, ws + w.sdf_c, R, n_coarse, n_imp, 0.f, inv_s, u_table, ws + w.z_f, nullptr,
                                   nullptr, stream, 1)))
      return rc;
    if ((rc = fneus_core_geometry(o, d, ws + w.z_f, R, n_imp, sample_dist, ws + w.dists, ws + w.mid, ws + w.pts_f,
                                  ws + w.dirs_f, stream)))
      return rc;
    // value + normal at the section mid-points (compute_weight / cal_firHit_rgb evaluate the same points)
    if ((rc = fneus_sdf_fwd_grad(sdf_cfg, sdf_wpack, ws + w.pts_f, Mf, ws + w.sdf_f, ws + w.feat, ws + w.nrm, ws + w.saved,
                                 ws + w.scratch, stream)))
      return rc;
    // alpha and weights of compute_weight (calLvis.py:93-150: cos_anneal_ratio = 0, zero colours)
    if ((rc = fneus_composite_fwd(ws + w.sdf_f, ws + w.nrm, ws + w.zero_rgb, ws + w.dists, ws + w.pts_f, d, nullptr, nullptr,
                                  nullptr, R, n_imp, 0, inv_s, 0.f, nullptr, ws + w.color, ws + w.weights, ws + w.wsum,
                                  ws + w.wmax, ws + w.cdf, ws + w.inside, ws + w.eik, reinterpret_cast<int*>(ws + w.hit_tmp),
                                  ws + w.wpair, stream)))
      return rc;
    // visibility = 1 - sum w * inside (calLvis.py:387-392); first sign change + secant root (calLvis.py:180-196)
    if ((rc = fneus_first_hit_secant(ws + w.sdf_f, ws + w.mid, ws + w.pts_f, o, d, ws + w.weights, n_imp, R, n_imp,
                                     hit_out + r0, nullptr, ws + w.p_s, lvis_out + r0, nullptr, stream)))
      return rc;
    // radiance at the first hit: normal + feature at the root, colour network (calLvis.py:197-203)
    if ((rc = fneus_sdf_fwd_grad(sdf_cfg, sdf_wpack, ws + w.p_s, R, ws + w.sdf_s, ws + w.feat, ws + w.n_s, ws + w.saved,
                                 ws + w.scratch, stream)))
      return rc;
    if ((rc = fneus_color_fwd(color_cfg, color_wpack, ws + w.p_s, ws + w.n_s, d, ws + w.feat, R, ws + w.rgb_s, nullptr,
                              ws + w.scratch, stream)))
      return rc;
    prof_begin(PC_ELEMENTWISE, 0.0, (double)R * 28.0, st);
    mask_rows3_kernel<<<cdiv(R * 3, 256), 256, 0, st>>>(ws + w.rgb_s, hit_out + r0, R, rgb_out + r0 * 3);
    prof_end(st);
    FNEUS_CHECK_LAUNCH();
  }
  return FNEUS_OK;
}

}  // extern "C"

// render_core's per-ray arithmetic (renderer.py:245-274 alpha from the sigmoid CDF, :284-292 first sign change,
// :328-332 inside-sphere weights, :350-372 background mix, transmittance product, compositing, eikonal term)
// and its closed-form backward.  One warp per ray, lane-strided samples, shuffle scans.
#include "fneus_common.cuh"
#include "prof.cuh"

namespace fneus {

constexpr int COMP_WARPS = 4;

__device__ __forceinline__ float wscan_mul(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v *= t;
  }
  return v;
}
// inclusive suffix sum: result(lane) = sum_{l >= lane} v(l)
__device__ __forceinline__ float wscan_add_rev(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_down_sync(0xffffffffu, v, o);
    if (lane + o < 32) v += t;
  }
  return v;
}
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int wmin_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

struct SampleFwd {
  float tc, ic, em, ep, P, N, araw, a;   // a = clipped sdf-alpha
  float in, relax, gn;                   // inside flag, relaxed flag, |g|
};

__device__ __forceinline__ SampleFwd sample_forward(float f, float gx, float gy, float gz, float delta, float px,
                                                    float py, float pz, float dx, float dy, float dz, float s,
                                                    float r) {
  SampleFwd o;
  o.tc = dx * gx + dy * gy + dz * gz;
  float u = -o.tc * 0.5f + 0.5f, v = -o.tc;
  o.ic = -(fmaxf(u, 0.f) * (1.0f - r) + fmaxf(v, 0.f) * r);
  float half = o.ic * delta * 0.5f;
  o.em = f - half;
  o.ep = f + half;
  o.P = sigmoid_bw(o.em * s);
  o.N = sigmoid_bw(o.ep * s);
  o.araw = (o.P - o.N + 1e-5f) / (o.P + 1e-5f);
  o.a = fminf(fmaxf(o.araw, 0.f), 1.f);
  const float rad2 = px * px + py * py + pz * pz;        // |p| < 1, |p| < 1.2 as tests on the squared radius
  o.in = radius_lt_1(rad2) ? 1.f : 0.f;
  o.relax = radius_lt_1p2(rad2) ? 1.f : 0.f;
  o.gn = sqrtf(gx * gx + gy * gy + gz * gz);
  return o;
}

__global__ void __launch_bounds__(COMP_WARPS * 32)
composite_fwd_kernel(const float* __restrict__ sdf, const float* __restrict__ nrm, const float* __restrict__ rgb,
                     const float* __restrict__ dists, const float* __restrict__ pts, const float* __restrict__ rays_d,
                     const float* __restrict__ bg_alpha, const float* __restrict__ bg_color,
                     const float* __restrict__ bg_rgb, long long B, int n_in, int n_out,
                     const float* __restrict__ inv_s_p, float car_host, const float* __restrict__ car_dev,
                     float* __restrict__ color,
                     float* __restrict__ weights, float* __restrict__ weight_sum, float* __restrict__ weight_max,
                     float* __restrict__ cdf, float* __restrict__ inside, float* __restrict__ eik_part,
                     int* __restrict__ hit_idx, float* __restrict__ w_pair) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const long long ray = (long long)blockIdx.x * COMP_WARPS + warp;
  if (ray >= B) return;
  const int n_tot = n_in + n_out;
  float* s_win = smem + (size_t)warp * n_in;   // inside-weights per sample (for the surface blend)
  const float s = __ldg(inv_s_p);
  const float car = car_dev != nullptr ? __ldg(car_dev) : car_host;   // device scalar: the annealing schedule inside a CUDA graph
  const float dx = rays_d[ray * 3], dy = rays_d[ray * 3 + 1], dz = rays_d[ray * 3 + 2];
  const bool has_bg = bg_alpha != nullptr;

  float carry = 1.f, carry_in = 1.f;
  float cr = 0.f, cg = 0.f, cb = 0.f, ws = 0.f, wm = -1e30f, e_num = 0.f, e_den = 0.f, any_in = 0.f;
  int first_neg = 1 << 30;
  for (int c0 = 0; c0 < n_tot; c0 += 32) {
    int i = c0 + lane;
    float a_mix = 0.f, a_in = 0.f, c0r = 0.f, c0g = 0.f, c0b = 0.f;
    if (i < n_in) {
      long long q = ray * n_in + i;
      float f = __ldg(sdf + q);
      SampleFwd o = sample_forward(f, __ldg(nrm + q * 3), __ldg(nrm + q * 3 + 1), __ldg(nrm + q * 3 + 2),
                                   __ldg(dists + q), __ldg(pts + q * 3), __ldg(pts + q * 3 + 1),
                                   __ldg(pts + q * 3 + 2), dx, dy, dz, s, car);
      cdf[q] = o.P;
      inside[q] = o.in;
      any_in += o.in;
      if (f < 0.f) first_neg = min(first_neg, i);
      float ge = o.gn - 1.0f;
      e_num += o.relax * (ge * ge);
      e_den += o.relax;
      a_in = o.a * o.in;
      c0r = __ldg(rgb + q * 3); c0g = __ldg(rgb + q * 3 + 1); c0b = __ldg(rgb + q * 3 + 2);
      if (has_bg) {
        long long qb = ray * n_tot + i;
        float ba = __ldg(bg_alpha + qb), om = 1.0f - o.in;
        a_mix = o.a * o.in + ba * om;
        c0r = c0r * o.in + __ldg(bg_color + qb * 3) * om;
        c0g = c0g * o.in + __ldg(bg_color + qb * 3 + 1) * om;
        c0b = c0b * o.in + __ldg(bg_color + qb * 3 + 2) * om;
      } else {
        a_mix = o.a;
      }
    } else if (i < n_tot) {
      long long qb = ray * n_tot + i;
      a_mix = __ldg(bg_alpha + qb);
      c0r = __ldg(bg_color + qb * 3); c0g = __ldg(bg_color + qb * 3 + 1); c0b = __ldg(bg_color + qb * 3 + 2);
    }
    float fac = i < n_tot ? (1.0f - a_mix + 1e-7f) : 1.f;
    float incl = wscan_mul(fac, lane);
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 1.f;
    float w = a_mix * (carry * excl);
    carry *= __shfl_sync(0xffffffffu, incl, 31);
    if (i < n_tot) {
      weights[ray * n_tot + i] = w;
      cr += w * c0r; cg += w * c0g; cb += w * c0b;
      ws += w;
      wm = fmaxf(wm, w);
    }
    if (c0 < n_in) {   // warp-uniform: inside-weights scan only over the inside block
      float fac2 = i < n_in ? (1.0f - a_in + 1e-7f) : 1.f;
      float incl2 = wscan_mul(fac2, lane);
      float excl2 = __shfl_up_sync(0xffffffffu, incl2, 1);
      if (lane == 0) excl2 = 1.f;
      if (i < n_in) s_win[i] = a_in * (carry_in * excl2);
      carry_in *= __shfl_sync(0xffffffffu, incl2, 31);
    }
  }
  cr = wsum(cr); cg = wsum(cg); cb = wsum(cb); ws = wsum(ws); wm = wmax(wm);
  e_num = wsum(e_num); e_den = wsum(e_den); any_in = wsum(any_in);
  first_neg = wmin_i(first_neg);
  __syncwarp();
  if (lane == 0) {
    if (bg_rgb) {
      float rest = 1.0f - ws;
      cr += bg_rgb[0] * rest; cg += bg_rgb[1] * rest; cb += bg_rgb[2] * rest;
    }
    color[ray * 3] = cr; color[ray * 3 + 1] = cg; color[ray * 3 + 2] = cb;
    weight_sum[ray] = ws;
    weight_max[ray] = wm;
    eik_part[ray * 2] = e_num; eik_part[ray * 2 + 1] = e_den;
    bool hit = first_neg < n_in && first_neg >= 1 && any_in > 0.f;
    hit_idx[ray] = hit ? first_neg : -1;
    w_pair[ray * 2] = hit ? s_win[first_neg - 1] + 1e-5f : 1.f;
    w_pair[ray * 2 + 1] = hit ? s_win[first_neg] + 1e-5f : 1.f;
  }
}

__global__ void __launch_bounds__(COMP_WARPS * 32)
composite_bwd_kernel(const float* __restrict__ sdf, const float* __restrict__ nrm, const float* __restrict__ rgb,
                     const float* __restrict__ dists, const float* __restrict__ pts, const float* __restrict__ rays_d,
                     const float* __restrict__ bg_alpha, const float* __restrict__ bg_color,
                     const float* __restrict__ bg_rgb, long long B, int n_in, int n_out,
                     const float* __restrict__ inv_s_p, float car_host, const float* __restrict__ car_dev,
                     const int* __restrict__ hit_idx,
                     const float* __restrict__ d_color, const float* __restrict__ d_weights,
                     const float* __restrict__ d_weight_sum, const float* __restrict__ d_w_pair,
                     const float* __restrict__ d_eik, const float* __restrict__ eik_denom, float* __restrict__ d_sdf,
                     float* __restrict__ d_nrm, float* __restrict__ d_rgb, float* __restrict__ d_inv_s,
                     float* __restrict__ d_bg_alpha, float* __restrict__ d_bg_color) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const long long ray = (long long)blockIdx.x * COMP_WARPS + warp;
  if (ray >= B) return;
  const int n_tot = n_in + n_out;
  // per-warp arrays over n_tot: mixed alpha, transmittance, w*wbar suffix sums, inside-transmittance
  float* s_a = smem + (size_t)warp * 4 * n_tot;
  float* s_T = s_a + n_tot;
  float* s_S = s_T + n_tot;
  float* s_Tin = s_S + n_tot;
  const float s = __ldg(inv_s_p);
  const float car = car_dev != nullptr ? __ldg(car_dev) : car_host;   // device scalar: the annealing schedule inside a CUDA graph
  const float dx = rays_d[ray * 3], dy = rays_d[ray * 3 + 1], dz = rays_d[ray * 3 + 2];
  const bool has_bg = bg_alpha != nullptr;
  const float dcr = d_color ? d_color[ray * 3] : 0.f, dcg = d_color ? d_color[ray * 3 + 1] : 0.f,
              dcb = d_color ? d_color[ray * 3 + 2] : 0.f;
  const float dws = d_weight_sum ? d_weight_sum[ray] : 0.f;
  float bgdot = 0.f;
  if (bg_rgb) bgdot = dcr * bg_rgb[0] + dcg * bg_rgb[1] + dcb * bg_rgb[2];
  const int hidx = hit_idx[ray];
  const float dp0 = (d_w_pair && hidx >= 1) ? d_w_pair[ray * 2] : 0.f;
  const float dp1 = (d_w_pair && hidx >= 1) ? d_w_pair[ray * 2 + 1] : 0.f;
  const bool pair_path = (dp0 != 0.f) || (dp1 != 0.f);
  const float eik_scale = d_eik ? __ldg(d_eik) / __ldg(eik_denom) : 0.f;

  // pass 1: recompute alpha (mixed), transmittance T, inside transmittance
  float carry = 1.f, carry_in = 1.f;
  for (int c0 = 0; c0 < n_tot; c0 += 32) {
    int i = c0 + lane;
    float a_mix = 0.f, a_in = 0.f;
    if (i < n_in) {
      long long q = ray * n_in + i;
      SampleFwd o = sample_forward(__ldg(sdf + q), __ldg(nrm + q * 3), __ldg(nrm + q * 3 + 1), __ldg(nrm + q * 3 + 2),
                                   __ldg(dists + q), __ldg(pts + q * 3), __ldg(pts + q * 3 + 1),
                                   __ldg(pts + q * 3 + 2), dx, dy, dz, s, car);
      a_in = o.a * o.in;
      a_mix = has_bg ? o.a * o.in + __ldg(bg_alpha + ray * n_tot + i) * (1.0f - o.in) : o.a;
    } else if (i < n_tot) {
      a_mix = __ldg(bg_alpha + ray * n_tot + i);
    }
    float fac = i < n_tot ? (1.0f - a_mix + 1e-7f) : 1.f;
    float incl = wscan_mul(fac, lane);
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 1.f;
    if (i < n_tot) { s_a[i] = a_mix; s_T[i] = carry * excl; }
    carry *= __shfl_sync(0xffffffffu, incl, 31);
    float fac2 = i < n_in ? (1.0f - a_in + 1e-7f) : 1.f;
    float incl2 = wscan_mul(fac2, lane);
    float excl2 = __shfl_up_sync(0xffffffffu, incl2, 1);
    if (lane == 0) excl2 = 1.f;
    if (i < n_tot) s_Tin[i] = carry_in * excl2;
    carry_in *= __shfl_sync(0xffffffffu, incl2, 31);
  }
  __syncwarp();
  // pass 2 (reverse): S_i = sum_{k>i} w_k * wbar_k
  float tail = 0.f;
  const int nchunks = (n_tot + 31) / 32;
  for (int c = nchunks - 1; c >= 0; c--) {
    int i = c * 32 + lane;
    float v = 0.f;
    if (i < n_tot) {
      float cr_, cg_, cb_;
      if (i < n_in) {
        long long q = ray * n_in + i;
        cr_ = __ldg(rgb + q * 3); cg_ = __ldg(rgb + q * 3 + 1); cb_ = __ldg(rgb + q * 3 + 2);
        if (has_bg) {
          float px = __ldg(pts + q * 3), py = __ldg(pts + q * 3 + 1), pz = __ldg(pts + q * 3 + 2);
          float in = radius_lt_1(px * px + py * py + pz * pz) ? 1.f : 0.f, om = 1.0f - in;
          long long qb = ray * n_tot + i;
          cr_ = cr_ * in + __ldg(bg_color + qb * 3) * om;
          cg_ = cg_ * in + __ldg(bg_color + qb * 3 + 1) * om;
          cb_ = cb_ * in + __ldg(bg_color + qb * 3 + 2) * om;
        }
      } else {
        long long qb = ray * n_tot + i;
        cr_ = __ldg(bg_color + qb * 3); cg_ = __ldg(bg_color + qb * 3 + 1); cb_ = __ldg(bg_color + qb * 3 + 2);
      }
      float wbar = dcr * cr_ + dcg * cg_ + dcb * cb_ + dws - bgdot + (d_weights ? __ldg(d_weights + ray * n_tot + i) : 0.f);
      v = s_a[i] * s_T[i] * wbar;
    }
    float incl = wscan_add_rev(v, lane);
    if (i < n_tot) s_S[i] = tail + incl - v;
    tail += __shfl_sync(0xffffffffu, incl, 0);
  }
  __syncwarp();
  // inside-weight pair (surface blend weights)
  float win0 = 0.f, win1 = 0.f;
  if (pair_path) {
    // a_in_j = w_in_j / Tin_j is recomputed below per sample; w_in at the two gathered indices:
    // need a_in at hidx-1, hidx: recompute from inputs by the owning lanes and broadcast through smem T arrays.
    // (cheap: two samples)
    for (int t = 0; t < 2; t++) {
      int j = hidx - 1 + t;
      long long q = ray * n_in + j;
      SampleFwd o = sample_forward(__ldg(sdf + q), __ldg(nrm + q * 3), __ldg(nrm + q * 3 + 1), __ldg(nrm + q * 3 + 2),
                                   __ldg(dists + q), __ldg(pts + q * 3), __ldg(pts + q * 3 + 1),
                                   __ldg(pts + q * 3 + 2), dx, dy, dz, s, car);
      float wj = o.a * o.in * s_Tin[j];
      if (t == 0) win0 = wj; else win1 = wj;
    }
  }
  // pass 3: per-sample gradients
  float dinv = 0.f;
  for (int c0 = 0; c0 < n_tot; c0 += 32) {
    int i = c0 + lane;
    if (i >= n_tot) continue;
    float a_mix = s_a[i], T = s_T[i];
    float w = a_mix * T;
    if (i < n_in) {
      long long q = ray * n_in + i;
      float f = __ldg(sdf + q), gx = __ldg(nrm + q * 3), gy = __ldg(nrm + q * 3 + 1), gz = __ldg(nrm + q * 3 + 2);
      float delta = __ldg(dists + q);
      SampleFwd o = sample_forward(f, gx, gy, gz, delta, __ldg(pts + q * 3), __ldg(pts + q * 3 + 1),
                                   __ldg(pts + q * 3 + 2), dx, dy, dz, s, car);
      float cr_ = __ldg(rgb + q * 3), cg_ = __ldg(rgb + q * 3 + 1), cb_ = __ldg(rgb + q * 3 + 2);
      float mr = cr_, mg = cg_, mb = cb_;
      float om = 1.0f - o.in;
      long long qb = ray * n_tot + i;
      if (has_bg) {
        mr = cr_ * o.in + __ldg(bg_color + qb * 3) * om;
        mg = cg_ * o.in + __ldg(bg_color + qb * 3 + 1) * om;
        mb = cb_ * o.in + __ldg(bg_color + qb * 3 + 2) * om;
      }
      float wbar = dcr * mr + dcg * mg + dcb * mb + dws - bgdot + (d_weights ? __ldg(d_weights + qb) : 0.f);
      float abar_mix = T * wbar - s_S[i] / (1.0f - a_mix + 1e-7f);
      float cfac = has_bg ? o.in : 1.f;
      d_rgb[q * 3] = w * dcr * cfac; d_rgb[q * 3 + 1] = w * dcg * cfac; d_rgb[q * 3 + 2] = w * dcb * cfac;
      if (has_bg) {
        if (d_bg_alpha) d_bg_alpha[qb] = abar_mix * om;
        if (d_bg_color) {
          d_bg_color[qb * 3] = w * dcr * om; d_bg_color[qb * 3 + 1] = w * dcg * om; d_bg_color[qb * 3 + 2] = w * dcb * om;
        }
      }
      float abar = abar_mix * cfac;
      if (pair_path) {
        float a_in = o.a * o.in;
        float ain_bar = 0.f;
        float den = 1.0f - a_in + 1e-7f;
        if (i == hidx - 1) ain_bar += s_Tin[i] * dp0;
        if (i == hidx) ain_bar += s_Tin[i] * dp1;
        if (i < hidx - 1) ain_bar -= win0 * dp0 / den;
        if (i < hidx) ain_bar -= win1 * dp1 / den;
        abar += ain_bar * o.in;
      }
      if (!(o.araw >= 0.f && o.araw <= 1.f)) abar = 0.f;
      float pe = o.P + 1e-5f;
      float Pbar = abar * (1.0f / pe - (o.P - o.N + 1e-5f) / (pe * pe));
      float Nbar = -abar / pe;
      float tP = o.P * (1.0f - o.P) * Pbar, tN = o.N * (1.0f - o.N) * Nbar;
      float em_bar = s * tP, ep_bar = s * tN;
      dinv += o.em * tP + o.ep * tN;
      d_sdf[q] = em_bar + ep_bar;
      float ic_bar = (delta * 0.5f) * (ep_bar - em_bar);
      float u = -o.tc * 0.5f + 0.5f, v = -o.tc;
      float tc_bar = ic_bar * ((u > 0.f ? (1.0f - car) * 0.5f : 0.f) + (v > 0.f ? car : 0.f));
      float gxb = tc_bar * dx, gyb = tc_bar * dy, gzb = tc_bar * dz;
      if (eik_scale != 0.f && o.relax > 0.f && o.gn > 0.f) {
        float k = eik_scale * 2.0f * (o.gn - 1.0f) / o.gn;
        gxb += k * gx; gyb += k * gy; gzb += k * gz;
      }
      d_nrm[q * 3] = gxb; d_nrm[q * 3 + 1] = gyb; d_nrm[q * 3 + 2] = gzb;
    } else {
      long long qb = ray * n_tot + i;
      float mr = __ldg(bg_color + qb * 3), mg = __ldg(bg_color + qb * 3 + 1), mb = __ldg(bg_color + qb * 3 + 2);
      float wbar = dcr * mr + dcg * mg + dcb * mb + dws - bgdot + (d_weights ? __ldg(d_weights + qb) : 0.f);
      float abar_mix = T * wbar - s_S[i] / (1.0f - a_mix + 1e-7f);
      if (d_bg_alpha) d_bg_alpha[qb] = abar_mix;
      if (d_bg_color) {
        d_bg_color[qb * 3] = w * dcr; d_bg_color[qb * 3 + 1] = w * dcg; d_bg_color[qb * 3 + 2] = w * dcb;
      }
    }
  }
  dinv = wsum(dinv);
  if (lane == 0 && d_inv_s) d_inv_s[ray] = dinv;
}

}  // namespace fneus

using namespace fneus;

extern "C" {

int fneus_composite_fwd(const float* sdf, const float* normals, const float* rgb, const float* dists,
                        const float* pts, const float* rays_d, const float* bg_alpha, const float* bg_color,
                        const float* bg_rgb, long long B, int n_in, int n_out, const float* inv_s,
                        float cos_anneal_ratio, const float* cos_anneal_dev, float* color, float* weights,
                        float* weight_sum, float* weight_max,
                        float* cdf, float* inside, float* eik_part, int* hit_idx, float* w_pair, void* stream) {
  if (B == 0) return FNEUS_OK;
  if (!sdf || !normals || !rgb || !dists || !pts || !rays_d || !inv_s || !color || !weights || !weight_sum ||
      !weight_max || !cdf || !inside || !eik_part || !hit_idx || !w_pair)
    return FNEUS_ERR_NULL;
  // same bound as the backward entry: a shape accepted here can always be back-propagated
  if (B < 0 || n_in < 1 || n_out < 0 || n_in + n_out > 2048) return FNEUS_ERR_BAD_SHAPE;
  if ((n_out > 0) && (!bg_alpha || !bg_color)) return FNEUS_ERR_NULL;
  if ((bg_alpha == nullptr) != (bg_color == nullptr)) return FNEUS_ERR_NULL;
  size_t smem = (size_t)COMP_WARPS * n_in * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(composite_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fneus_cuda_error((int)e);
  }
  prof_begin(PC_COMPOSITE, 0.0, (double)B * ((n_in + n_out) * 44.0 + 100.0), (cudaStream_t)stream);
  composite_fwd_kernel<<<cdiv(B, COMP_WARPS), COMP_WARPS * 32, smem, (cudaStream_t)stream>>>(
      sdf, normals, rgb, dists, pts, rays_d, bg_alpha, bg_color, bg_rgb, B, n_in, n_out, inv_s, cos_anneal_ratio,
      cos_anneal_dev, color, weights, weight_sum, weight_max, cdf, inside, eik_part, hit_idx, w_pair);
  prof_end((cudaStream_t)stream);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

int fneus_composite_bwd(const float* sdf, const float* normals, const float* rgb, const float* dists,
                        const float* pts, const float* rays_d, const float* bg_alpha, const float* bg_color,
                        const float* bg_rgb, long long B, int n_in, int n_out, const float* inv_s,
                        float cos_anneal_ratio, const float* cos_anneal_dev, const int* hit_idx, const float* d_color,
                        const float* d_weights,
                        const float* d_weight_sum, const float* d_w_pair, const float* d_eik,
                        const float* eik_denom, float* d_sdf, float* d_normals, float* d_rgb, float* d_inv_s,
                        float* d_bg_alpha, float* d_bg_color, void* stream) {
  if (B == 0) return FNEUS_OK;
  if (!sdf || !normals || !rgb || !dists || !pts || !rays_d || !inv_s || !hit_idx || !d_sdf || !d_normals || !d_rgb)
    return FNEUS_ERR_NULL;
  if (d_eik && !eik_denom) return FNEUS_ERR_NULL;
  if (B < 0 || n_in < 1 || n_out < 0 || n_in + n_out > 2048) return FNEUS_ERR_BAD_SHAPE;
  if ((n_out > 0) && (!bg_alpha || !bg_color)) return FNEUS_ERR_NULL;
  size_t smem = (size_t)COMP_WARPS * 4 * (n_in + n_out) * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(composite_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fneus_cuda_error((int)e);
  }
  prof_begin(PC_COMPOSITE, 0.0, (double)B * ((n_in + n_out) * 64.0 + 100.0), (cudaStream_t)stream);
  composite_bwd_kernel<<<cdiv(B, COMP_WARPS), COMP_WARPS * 32, smem, (cudaStream_t)stream>>>(
      sdf, normals, rgb, dists, pts, rays_d, bg_alpha, bg_color, bg_rgb, B, n_in, n_out, inv_s, cos_anneal_ratio,
      cos_anneal_dev, hit_idx, d_color, d_weights, d_weight_sum, d_w_pair, d_eik, eik_denom, d_sdf, d_normals, d_rgb, d_inv_s,
      d_bg_alpha, d_bg_color);
  prof_end((cudaStream_t)stream);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

}  // extern "C"

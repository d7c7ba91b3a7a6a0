// Hierarchical sampling kernels (renderer.py:43-77 sample_pdf, :152-189 up_sample, :191-205 cat_z_vals,
// :224-236 section geometry; duplicates calLvis.py:25-90).  One warp per ray: coalesced loads into a
// per-warp shared-memory row, shuffle scans for the transmittance product and the CDF sum, binary
// search per importance sample.  No gradients flow through any of this (renderer.py:426 no_grad).
#include "fneus_common.cuh"
#include "prof.cuh"

namespace fneus {

constexpr int SAMP_WARPS = 4;

__device__ __forceinline__ float warp_incl_scan_mul(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v *= t;
  }
  return v;
}
__device__ __forceinline__ float warp_incl_scan_add(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// first index i in [0,n) with c[i] > u  (torch.searchsorted(..., right=True)), n if none
__device__ __forceinline__ int upper_bound(const float* c, int n, float u) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (c[mid] > u) hi = mid; else lo = mid + 1;
  }
  return lo;
}
// number of entries < v  (lower bound) in sorted c[0..n)
__device__ __forceinline__ int count_less(const float* c, int n, float v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (c[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ int count_leq(const float* c, int n, float v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (c[mid] <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// pts (optional, [k,3]) with o3 / d3 = the ray: the new samples' positions, rounded exactly like ray_points_kernel
__device__ __forceinline__ void invert_cdf_warp(const float* sz, const float* scdf, int n, int k,
                                                const float* u_table, float* out, long long* inds, int lane,
                                                float* pts = nullptr, const float* o3 = nullptr, const float* d3 = nullptr) {
  for (int m = lane; m < k; m += 32) {
    float u = __ldg(u_table + m);
    int idx = upper_bound(scdf, n, u);
    int lo = idx - 1 > 0 ? idx - 1 : 0;
    int hi = idx < n - 1 ? idx : n - 1;
    float den = scdf[hi] - scdf[lo];
    if (den < 1e-5f) den = 1.f;
    float t = (u - scdf[lo]) / den;
    const float zn = __fadd_rn(sz[lo], __fmul_rn(t, sz[hi] - sz[lo]));
    out[m] = zn;
    if (inds) inds[m] = idx;
    if (pts) {
#pragma unroll
      for (int c = 0; c < 3; c++) pts[m * 3 + c] = __fadd_rn(o3[c], __fmul_rn(d3[c], zn));
    }
  }
}

// One up-sampling step on a ray whose depths / sdf values sit in shared memory (sz, sf; sc: scratch for alpha, then the CDF):
// up_sample + sample_pdf (renderer.py:152-189, 43-77).  Shared by the single-step kernel and the fused iteration kernel.
__device__ __forceinline__ void upsample_row(float* sz, float* sf, float* sc, int n, const float* __restrict__ rays_o,
                                             const float* __restrict__ rays_d, long long ray, float inv_s, int k,
                                             const float* __restrict__ u_table, float* new_z_row, float* cdf_row,
                                             long long* inds_row, float* pts_row, int lane) {
  const float ox = rays_o[ray * 3], oy = rays_o[ray * 3 + 1], oz = rays_o[ray * 3 + 2];
  const float dx = rays_d[ray * 3], dy = rays_d[ray * 3 + 1], dz = rays_d[ray * 3 + 2];
  // |o + d t| < 1 as a test on the squared radius (radius_lt_1: identical to comparing the rounded square root)
  auto in_sphere = [&](float t) {
    float px = __fadd_rn(ox, __fmul_rn(dx, t)), py = __fadd_rn(oy, __fmul_rn(dy, t)),
          pz = __fadd_rn(oz, __fmul_rn(dz, t));
    return radius_lt_1(__fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), __fmul_rn(pz, pz)));
  };
  auto slope = [&](int j) { return (sf[j + 1] - sf[j]) / (sz[j + 1] - sz[j] + 1e-5f); };
  const int ni = n - 1;  // intervals
  // pass 1: alpha_j, transmittance scan, weights; accumulate sum(w + 1e-5).  Each lane evaluates ITS interval's slope and
  // its left end point's sphere test once; the neighbour's values come by shuffle (the previous chunk's last ones are carried).
  float carry = 1.f, wsum = 0.f, slope_carry = 0.f;
  for (int c0 = 0; c0 < ni; c0 += 32) {
    int j = c0 + lane;
    float alpha = 0.f;
    const bool act = j < ni;
    const float cosv = act ? slope(j) : 0.f;
    const bool in_l = j < n ? in_sphere(sz[j]) : false;
    float prev = __shfl_up_sync(0xffffffffu, cosv, 1);
    if (lane == 0) prev = slope_carry;                                  // 0 for the first interval of the ray
    int in_r = __shfl_down_sync(0xffffffffu, (int)in_l, 1);
    if (lane == 31) in_r = (j + 1 < n) ? (int)in_sphere(sz[j + 1]) : 0;
    slope_carry = __shfl_sync(0xffffffffu, cosv, 31);
    if (act) {
      bool inside = in_l || (in_r != 0);
      float cm = fminf(prev, cosv);
      cm = fminf(fmaxf(cm, -1e3f), 0.f) * (inside ? 1.f : 0.f);
      float dist = sz[j + 1] - sz[j];
      float mid = (sf[j] + sf[j + 1]) * 0.5f;
      float half = __fmul_rn(__fmul_rn(cm, dist), 0.5f);
      // IEEE sigmoids here: (pc - nc + 1e-5) / (pc + 1e-5) amplifies their last bits where pc is tiny, and the CDF built
      // from it is compared to the reference at 2e-6 (an approximate reciprocal moved 7 of 1536 entries by up to 7e-6)
      float pc = sigmoidf_(__fmul_rn(mid - half, inv_s));
      float nc = sigmoidf_(__fmul_rn(mid + half, inv_s));
      alpha = (pc - nc + 1e-5f) / (pc + 1e-5f);
    }
    float fac = j < ni ? (1.f - alpha + 1e-7f) : 1.f;
    float incl = warp_incl_scan_mul(fac, lane);
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 1.f;
    float T = carry * excl;
    float w = alpha * T + 1e-5f;
    if (j < ni) { sc[j + 1] = w; wsum += w; }
    carry *= __shfl_sync(0xffffffffu, incl, 31);
  }
  wsum = warp_sum(wsum);
  __syncwarp();
  // pass 2: cdf = [0, cumsum(w / wsum)]
  float run = 0.f;
  for (int c0 = 0; c0 < ni; c0 += 32) {
    int j = c0 + lane;
    float p = j < ni ? sc[j + 1] / wsum : 0.f;
    float incl = warp_incl_scan_add(p, lane);
    if (j < ni) sc[j + 1] = run + incl;
    run += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) sc[0] = 0.f;
  __syncwarp();
  if (cdf_row)
    for (int j = lane; j < n; j += 32) cdf_row[j] = sc[j];
  const float o3[3] = {ox, oy, oz}, d3[3] = {dx, dy, dz};
  invert_cdf_warp(sz, sc, n, k, u_table, new_z_row, inds_row, lane, pts_row, o3, d3);
}

__global__ void __launch_bounds__(SAMP_WARPS * 32)
upsample_step_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                     const float* __restrict__ z, long long z_stride, const float* __restrict__ sdf, long long B, int n,
                     int k, float inv_s_host, const float* __restrict__ inv_s_dev, const float* __restrict__ u_table,
                     float* __restrict__ new_z, float* __restrict__ cdf_out, long long* __restrict__ inds_out) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const long long ray = (long long)blockIdx.x * SAMP_WARPS + warp;
  if (ray >= B) return;
  const float inv_s = inv_s_dev != nullptr ? __ldg(inv_s_dev) : inv_s_host;   // device scalar: the learned inv_s of stage 2
  float* sz = smem + (size_t)warp * 3 * n;
  float* sf = sz + n;
  float* sc = sf + n;   // alpha, then cdf
  for (int j = lane; j < n; j += 32) {
    sz[j] = __ldg(z + ray * z_stride + j);             // z_stride 0: one depth table shared by all rays (stage 2)
    sf[j] = __ldg(sdf + ray * n + j);
  }
  __syncwarp();
  upsample_row(sz, sf, sc, n, rays_o, rays_d, ray, inv_s, k, u_table, new_z + ray * k, cdf_out ? cdf_out + ray * n : nullptr,
               inds_out ? inds_out + ray * k : nullptr, nullptr, lane);
}

__global__ void __launch_bounds__(SAMP_WARPS * 32)
inverse_cdf_kernel(const float* __restrict__ bins, const float* __restrict__ cdf, const float* __restrict__ u_table,
                   long long B, int n, int k, float* __restrict__ out, long long* __restrict__ inds) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const long long ray = (long long)blockIdx.x * SAMP_WARPS + warp;
  if (ray >= B) return;
  float* sz = smem + (size_t)warp * 2 * n;
  float* sc = sz + n;
  for (int j = lane; j < n; j += 32) { sz[j] = bins[ray * n + j]; sc[j] = cdf[ray * n + j]; }
  __syncwarp();
  invert_cdf_warp(sz, sc, n, k, u_table, out + ray * k, inds ? inds + ray * k : nullptr, lane);
}

// cat_z_vals (renderer.py:191-205) as a rank merge.  Only the k new samples are ranked (binary search into the old row);
// their output slots are marked in a bitmask, and every output position then finds its source with two popcounts
// (new element: its ordinal among the set bits; old element: position minus the set bits below it), so all stores are
// coalesced in output order.  Ties: old samples first, each list in its own order (what a stable sort of the
// concatenation [old, new] gives).
// Rank merge of two sorted rows in shared memory (cat_z_vals, renderer.py:191-205): every new sample goes to position
// j + #(old <= new_j), the old ones fill the gaps in order; the riding sdf values follow.  Results to global rows zo / so
// (so may be null) and, when dz is given, also to the shared rows dz / df (the next step's input).
__device__ __forceinline__ void merge_row(const float* sa, const float* sb, const float* sfa, const float* sfb, int n, int k,
                                          unsigned* mask, unsigned* pre, bool carry, float* zo, float* so, float* dz,
                                          float* df, int lane) {
  const int tot = n + k, W = (tot + 31) >> 5;
  for (int w = lane; w < W; w += 32) mask[w] = 0u;
  __syncwarp();
  for (int j = lane; j < k; j += 32) {
    const int pos = j + count_leq(sa, n, sb[j]);
    atomicOr(&mask[pos >> 5], 1u << (pos & 31));
  }
  __syncwarp();
  unsigned run = 0u;
  for (int w0 = 0; w0 < W; w0 += 32) {
    const int w = w0 + lane;
    const unsigned c = w < W ? (unsigned)__popc(mask[w]) : 0u;
    unsigned incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (w < W) pre[w] = run + incl - c;
    run += __shfl_sync(0xffffffffu, incl, 31);
  }
  __syncwarp();
  for (int q = lane; q < tot; q += 32) {
    const unsigned m = mask[q >> 5];
    const int bit = q & 31;
    const int below = (int)pre[q >> 5] + __popc(m & ((1u << bit) - 1u));
    const bool from_b = (m >> bit) & 1u;
    const int i = from_b ? below : q - below;
    const float zv = from_b ? sb[i] : sa[i];
    zo[q] = zv;
    if (dz) dz[q] = zv;
    if (carry) {
      const float fv = from_b ? sfb[i] : sfa[i];
      so[q] = fv;
      if (df) df[q] = fv;
    }
  }
}

__global__ void __launch_bounds__(SAMP_WARPS * 32)
merge_sorted_kernel(const float* __restrict__ z, const float* __restrict__ new_z, const float* __restrict__ sdf,
                    const float* __restrict__ new_sdf, long long B, int n, int k, float* __restrict__ z_out,
                    float* __restrict__ sdf_out) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const long long ray = (long long)blockIdx.x * SAMP_WARPS + warp;
  if (ray >= B) return;
  const int tot = n + k, W = (tot + 31) >> 5;
  float* sa = smem + (size_t)warp * (2 * tot + 2 * W);
  float* sb = sa + n;
  float* sfa = sb + k;                                   // the riding sdf values: every global load of the ray is issued
  float* sfb = sfa + n;                                  // up front (memory-level parallelism), nothing is loaded later
  unsigned* mask = reinterpret_cast<unsigned*>(sfb + k);
  unsigned* pre = mask + W;                              // set bits in the words below
  const bool carry = sdf && new_sdf && sdf_out;
  for (int j = lane; j < n; j += 32) {
    sa[j] = z[ray * n + j];
    if (carry) sfa[j] = sdf[ray * n + j];
  }
  for (int j = lane; j < k; j += 32) {
    sb[j] = new_z[ray * k + j];
    if (carry) sfb[j] = new_sdf[ray * k + j];
  }
  merge_row(sa, sb, sfa, sfb, n, k, mask, pre, carry, z_out + ray * tot, carry ? sdf_out + ray * tot : nullptr, nullptr,
            nullptr, lane);
}

// One iteration of the hierarchical sampling loop (renderer.py:166-176) in one launch: merge the previous iteration's kp new
// samples (with their sdf values) into the ray's sorted row, up-sample k new depths from the merged row, and emit their
// positions for the next SDF pass.  kp = 0: no merge (first iteration).  Same arithmetic as the three separate kernels.
__global__ void __launch_bounds__(SAMP_WARPS * 32)
upsample_iter_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ z,
                     const float* __restrict__ sdf, long long B, int n, const float* __restrict__ prev_z,
                     const float* __restrict__ prev_sdf, int kp, int k, float inv_s, const float* __restrict__ u_table,
                     float* __restrict__ z_out, float* __restrict__ sdf_out, float* __restrict__ new_z,
                     float* __restrict__ pts_out) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const long long ray = (long long)blockIdx.x * SAMP_WARPS + warp;
  if (ray >= B) return;
  const int tot = n + kp, W = (tot + 31) >> 5;
  float* sz = smem + (size_t)warp * (5 * tot + 2 * W);   // merged row: z, sdf, scratch; then the two inputs of the merge
  float* sf = sz + tot;
  float* sc = sf + tot;
  float* sa = sc + tot;
  float* sb = sa + n;
  float* sfa = sb + kp;
  float* sfb = sfa + n;
  unsigned* mask = reinterpret_cast<unsigned*>(sfb + kp);
  unsigned* pre = mask + W;
  if (kp > 0) {
    for (int j = lane; j < n; j += 32) { sa[j] = z[ray * n + j]; sfa[j] = sdf[ray * n + j]; }
    for (int j = lane; j < kp; j += 32) { sb[j] = prev_z[ray * kp + j]; sfb[j] = prev_sdf[ray * kp + j]; }
    merge_row(sa, sb, sfa, sfb, n, kp, mask, pre, true, z_out + ray * tot, sdf_out + ray * tot, sz, sf, lane);
  } else {
    for (int j = lane; j < n; j += 32) { sz[j] = __ldg(z + ray * n + j); sf[j] = __ldg(sdf + ray * n + j); }
  }
  __syncwarp();
  upsample_row(sz, sf, sc, tot, rays_o, rays_d, ray, inv_s, k, u_table, new_z + ray * k, nullptr, nullptr,
               pts_out ? pts_out + ray * k * 3 : nullptr, lane);
}

__global__ void ray_points_kernel(const float* __restrict__ o, const float* __restrict__ d, const float* __restrict__ z,
                                  long long total, int n, float* __restrict__ pts) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  long long b = idx / n;
  float t = z[idx];
#pragma unroll
  for (int c = 0; c < 3; c++) pts[idx * 3 + c] = __fadd_rn(o[b * 3 + c], __fmul_rn(d[b * 3 + c], t));
}

// stage 2: points of R = m k secondary rays on ONE shared depth table, origins = surface points repeated k times
// (calLvis.py:357-366); also writes the expanded origins [R,3] once (first depth) for the later per-ray kernels
__global__ void lvis_coarse_points_kernel(const float* __restrict__ surf, int k, const float* __restrict__ d,
                                          const float* __restrict__ z_table, long long total, int n,
                                          float* __restrict__ pts, float* __restrict__ o_out) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long r = idx / n;
  const int i = (int)(idx - r * n);
  const float t = __ldg(z_table + i);
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const float oc = __ldg(surf + (r / k) * 3 + c);
    pts[idx * 3 + c] = __fadd_rn(oc, __fmul_rn(__ldg(d + r * 3 + c), t));
    if (i == 0) o_out[r * 3 + c] = oc;
  }
}
// rgb_out = hit ? rgb : 0   (calLvis.py:200-203: rays without a first hit keep the zero radiance)
__global__ void mask_rows3_kernel(const float* __restrict__ rgb, const int* __restrict__ hit, long long R,
                                  float* __restrict__ out) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * 3) return;
  out[idx] = hit[idx / 3] >= 0 ? rgb[idx] : 0.f;
}

__global__ void core_geometry_kernel(const float* __restrict__ o, const float* __restrict__ d,
                                     const float* __restrict__ z, long long total, int n, float sample_dist,
                                     float* __restrict__ dists, float* __restrict__ mid_z, float* __restrict__ pts,
                                     float* __restrict__ dirs) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  long long b = idx / n;
  int j = (int)(idx - b * n);
  float zj = z[idx];
  float dist = j + 1 < n ? z[idx + 1] - zj : sample_dist;
  float mid = __fadd_rn(zj, __fmul_rn(dist, 0.5f));
  if (dists) dists[idx] = dist;
  if (mid_z) mid_z[idx] = mid;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    float dc = d[b * 3 + c];
    pts[idx * 3 + c] = __fadd_rn(o[b * 3 + c], __fmul_rn(dc, mid));
    if (dirs) dirs[idx * 3 + c] = dc;
  }
}


// dataset.py:186-192 near_far_from_sphere: mid = 0.5 * (-b) / a with a = sum d^2, b = 2 sum o.d ; near/far = mid -/+ 1
__global__ void near_far_kernel(const float* __restrict__ o, const float* __restrict__ d, long long B,
                                float* __restrict__ near, float* __restrict__ far) {
  long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= B) return;
  const float ox = o[r * 3], oy = o[r * 3 + 1], oz = o[r * 3 + 2], dx = d[r * 3], dy = d[r * 3 + 1], dz = d[r * 3 + 2];
  const float a = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  const float b = __fmul_rn(2.0f, __fadd_rn(__fadd_rn(__fmul_rn(ox, dx), __fmul_rn(oy, dy)), __fmul_rn(oz, dz)));
  const float mid = __fdiv_rn(__fmul_rn(0.5f, -b), a);
  near[r] = __fsub_rn(mid, 1.0f);
  far[r] = __fadd_rn(mid, 1.0f);
}
// renderer.py:395-408: z = near + (far - near) * linspace(0,1,n) (+ (rand - 0.5) * 2.0 / n_samples when perturbed)
__global__ void coarse_z_kernel(const float* __restrict__ near, const float* __restrict__ far,
                                const float* __restrict__ lin, const float* __restrict__ rnd, long long total, int n,
                                float inv_n_samples, float* __restrict__ z) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long r = idx / n;
  const int j = (int)(idx - r * n);
  float v = __fadd_rn(near[r], __fmul_rn(__fsub_rn(far[r], near[r]), lin[j]));
  // torch divides a tensor by a host scalar as a multiplication by its FP32 reciprocal
  if (rnd != nullptr) v = __fadd_rn(v, __fmul_rn(__fmul_rn(__fsub_rn(rnd[r], 0.5f), 2.0f), inv_n_samples));
  z[idx] = v;
}
// First sign change along a ray and the secant root between the two bracketing section mid-points
// (renderer.py:588-602, calLvis.py:180-196): idx = argmin_i sign(sdf_i) * (n - i) -- the first sample with sdf < 0;
// sign(0) = 0, so an exact zero is NOT a hit -- valid iff that minimum is negative, idx >= 1 and at least one sample lies
// inside the unit sphere;  z* = (s_lo z_hi - s_hi z_lo) / (s_lo - s_hi + 1e-10),  p* = o + d z*.
// Fixed shapes: rays without a hit get hit_idx = -1 and the root of the clamped index (finite, ignored by the caller).
// Optionally also the light visibility of calLvis.py:387-392: lvis = 1 - sum_i w_i * inside_i.
// One warp per ray.
__global__ void __launch_bounds__(SAMP_WARPS * 32)
first_hit_secant_kernel(const float* __restrict__ sdf, const float* __restrict__ mid_z, const float* __restrict__ pts,
                        const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                        const float* __restrict__ weights, int ldw, long long B, int n, int* __restrict__ hit_idx,
                        float* __restrict__ z_surf, float* __restrict__ pts_surf, float* __restrict__ lvis,
                        int* __restrict__ any_inside) {
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const long long ray = (long long)blockIdx.x * SAMP_WARPS + warp;
  if (ray >= B) return;
  int first = n;                     // first index with sdf < 0
  bool any_in = false;
  float occ = 0.f;
  for (int i = lane; i < n; i += 32) {
    const long long q = ray * n + i;
    if (__ldg(sdf + q) < 0.f && i < first) first = i;
    const float px = __ldg(pts + q * 3), py = __ldg(pts + q * 3 + 1), pz = __ldg(pts + q * 3 + 2);
    const bool in = radius_lt_1(__fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), __fmul_rn(pz, pz)));
    any_in = any_in || in;
    if (weights != nullptr && in) occ += __ldg(weights + ray * ldw + i);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
  any_in = __any_sync(0xffffffffu, any_in);
  if (weights != nullptr) occ = warp_sum(occ);
  if (lane == 0) {
    const bool hit = first < n && first >= 1 && any_in;
    const int i = first < 1 ? 1 : (first > n - 1 ? n - 1 : first);
    const float s_lo = sdf[ray * n + i - 1], s_hi = sdf[ray * n + i];
    const float z_lo = mid_z[ray * n + i - 1], z_hi = mid_z[ray * n + i];
    const float zs = __fdiv_rn(__fsub_rn(__fmul_rn(s_lo, z_hi), __fmul_rn(s_hi, z_lo)),
                               __fadd_rn(__fsub_rn(s_lo, s_hi), 1e-10f));
    hit_idx[ray] = hit ? first : -1;
    if (z_surf) z_surf[ray] = zs;
    if (pts_surf) {
#pragma unroll
      for (int c = 0; c < 3; c++) pts_surf[ray * 3 + c] = __fadd_rn(rays_o[ray * 3 + c], __fmul_rn(rays_d[ray * 3 + c], zs));
    }
    if (lvis) lvis[ray] = 1.0f - occ;
    if (any_inside) any_inside[ray] = any_in ? 1 : 0;
  }
}
// Rays of a pinhole camera from pixel coordinates (dataset.py:115-151, gen_rays_at / gen_random_rays_at) on the device:
//   p = Kinv[:3,:3] (px, py, 1);  v = pose[:3,:3] (p / |p|);  o = pose[:3,3];  colour / mask looked up at (py, px)
// out [B,10] = (o, v, rgb, mask) like Dataset.gen_random_rays_at; near / far = near_far_from_sphere (dataset.py:186-192).
// The reference assembles these rows on the CPU and uploads them every iteration.
__global__ void gen_rays_kernel(const float* __restrict__ px, const float* __restrict__ py,
                                const float* __restrict__ kinv, const float* __restrict__ pose,
                                const float* __restrict__ image, const float* __restrict__ mask, int H, int W,
                                long long B, float* __restrict__ out10, float* __restrict__ near,
                                float* __restrict__ far) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= B) return;
  const float x = px[r], y = py[r];
  float p[3], v[3];
#pragma unroll
  for (int i = 0; i < 3; i++)
    p[i] = __fadd_rn(__fadd_rn(__fmul_rn(kinv[i * 4], x), __fmul_rn(kinv[i * 4 + 1], y)), kinv[i * 4 + 2]);
  const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(p[0], p[0]), __fmul_rn(p[1], p[1])), __fmul_rn(p[2], p[2])));
#pragma unroll
  for (int i = 0; i < 3; i++) p[i] = __fdiv_rn(p[i], nrm);
#pragma unroll
  for (int i = 0; i < 3; i++)
    v[i] = __fadd_rn(__fadd_rn(__fmul_rn(pose[i * 4], p[0]), __fmul_rn(pose[i * 4 + 1], p[1])), __fmul_rn(pose[i * 4 + 2], p[2]));
  const float o[3] = {pose[3], pose[7], pose[11]};
  float* row = out10 + r * 10;
#pragma unroll
  for (int i = 0; i < 3; i++) { row[i] = o[i]; row[3 + i] = v[i]; }
  int ix = (int)x, iy = (int)y;
  ix = ix < 0 ? 0 : (ix >= W ? W - 1 : ix);
  iy = iy < 0 ? 0 : (iy >= H ? H - 1 : iy);
  const long long pix = ((long long)iy * W + ix) * 3;
#pragma unroll
  for (int i = 0; i < 3; i++) row[6 + i] = image ? image[pix + i] : 0.f;
  row[9] = mask ? mask[pix] : 1.f;
  if (near && far) {
    const float a = __fadd_rn(__fadd_rn(__fmul_rn(v[0], v[0]), __fmul_rn(v[1], v[1])), __fmul_rn(v[2], v[2]));
    const float b = __fmul_rn(2.0f, __fadd_rn(__fadd_rn(__fmul_rn(o[0], v[0]), __fmul_rn(o[1], v[1])), __fmul_rn(o[2], v[2])));
    const float mid = __fdiv_rn(__fmul_rn(0.5f, -b), a);
    near[r] = __fsub_rn(mid, 1.0f);
    far[r] = __fadd_rn(mid, 1.0f);
  }
}
// renderer.py:296-303: flat sample rows (idx - 1, idx) of the two samples bracketing the first sign change
__global__ void hit_rows_kernel(const int* __restrict__ hit_idx, long long B, int n, long long* __restrict__ rows) {
  long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= B) return;
  int i = hit_idx[r];
  i = i < 1 ? 1 : i;
  rows[2 * r] = r * n + i - 1;
  rows[2 * r + 1] = r * n + i;
}

}  // namespace fneus

using namespace fneus;

extern "C" {

int fneus_ray_points(const float* rays_o, const float* rays_d, const float* z, long long B, int n, float* pts,
                     void* stream) {
  if (B == 0 || n == 0) return FNEUS_OK;
  if (!rays_o || !rays_d || !z || !pts) return FNEUS_ERR_NULL;
  if (B < 0 || n < 0) return FNEUS_ERR_BAD_SHAPE;
  long long total = B * n;
  prof_begin(PC_SAMPLING, 0.0, (double)total * 16.0, (cudaStream_t)stream);
  ray_points_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, z, total, n, pts);
  prof_end((cudaStream_t)stream);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

int fneus_first_hit_secant(const float* sdf, const float* mid_z, const float* pts, const float* rays_o,
                           const float* rays_d, const float* weights, int ldw, long long B, int n, int* hit_idx,
                           float* z_surf, float* pts_surf, float* lvis, int* any_inside, void* stream) {
  if (B == 0) return FNEUS_OK;
  if (!sdf || !mid_z || !pts || !hit_idx) return FNEUS_ERR_NULL;
  if (pts_surf && (!rays_o || !rays_d)) return FNEUS_ERR_NULL;
  if (lvis && !weights) return FNEUS_ERR_NULL;
  if (B < 0 || n < 2 || (weights && ldw < n)) return FNEUS_ERR_BAD_SHAPE;
  prof_begin(PC_SAMPLING, 0.0, (double)B * n * (weights ? 24.0 : 20.0), (cudaStream_t)stream);
  first_hit_secant_kernel<<<cdiv(B, SAMP_WARPS), SAMP_WARPS * 32, 0, (cudaStream_t)stream>>>(
      sdf, mid_z, pts, rays_o, rays_d, weights, ldw, B, n, hit_idx, z_surf, pts_surf, lvis, any_inside);
  prof_end((cudaStream_t)stream);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

}  // extern "C"

namespace fneus {
int upsample_step_launch(const float* rays_o, const float* rays_d, const float* z, const float* sdf, long long B,
                         int n, int k, float inv_s, const float* inv_s_dev, const float* u_table, float* new_z,
                         float* cdf_out, long long* inds_out, void* stream, int z_shared = 0) {
  if (B == 0 || k == 0) return FNEUS_OK;
  if (!rays_o || !rays_d || !z || !sdf || !u_table || !new_z) return FNEUS_ERR_NULL;
  if (B < 0 || n < 2 || k < 0 || n > 4096) return FNEUS_ERR_BAD_SHAPE;
  size_t smem = (size_t)SAMP_WARPS * 3 * n * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(upsample_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fneus_cuda_error((int)e);
  }
  prof_begin(PC_SAMPLING, 0.0, (double)B * (n * 8.0 + k * 4.0 + 24.0), (cudaStream_t)stream);
  upsample_step_kernel<<<cdiv(B, SAMP_WARPS), SAMP_WARPS * 32, smem, (cudaStream_t)stream>>>(
      rays_o, rays_d, z, z_shared ? 0 : (long long)n, sdf, B, n, k, inv_s, inv_s_dev, u_table, new_z, cdf_out, inds_out);
  prof_end((cudaStream_t)stream);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}
}  // namespace fneus

extern "C" {

int fneus_upsample_step(const float* rays_o, const float* rays_d, const float* z, const float* sdf, long long B,
                        int n, int k, float inv_s, const float* u_table, float* new_z, float* cdf_out,
                        long long* inds_out, void* stream) {
  return upsample_step_launch(rays_o, rays_d, z, sdf, B, n, k, inv_s, nullptr, u_table, new_z, cdf_out, inds_out, stream);
}
// the same step with inv_s read from device memory (stage 2 uses the LEARNED inv_s, calLvis.py:371-379: no host sync)
int fneus_upsample_step_dev(const float* rays_o, const float* rays_d, const float* z, const float* sdf, long long B,
                            int n, int k, const float* inv_s_dev, const float* u_table, float* new_z, void* stream) {
  if (!inv_s_dev) return FNEUS_ERR_NULL;
  return upsample_step_launch(rays_o, rays_d, z, sdf, B, n, k, 0.f, inv_s_dev, u_table, new_z, nullptr, nullptr, stream);
}

int fneus_upsample_iter(const float* rays_o, const float* rays_d, const float* z, const float* sdf, long long B, int n,
                        const float* prev_z, const float* prev_sdf, int kp, int k, float inv_s, const float* u_table,
                        float* z_out, float* sdf_out, float* new_z, float* pts_out, void* stream) {
  if (B == 0 || k == 0) return FNEUS_OK;
  if (!rays_o || !rays_d || !z || !sdf || !u_table || !new_z) return FNEUS_ERR_NULL;
  if (kp > 0 && (!prev_z || !prev_sdf || !z_out || !sdf_out)) return FNEUS_ERR_NULL;
  if (B < 0 || n < 2 || k < 0 || kp < 0 || n + kp > 4096) return FNEUS_ERR_BAD_SHAPE;
  const int tot = n + kp;
  size_t smem = (size_t)SAMP_WARPS * (5 * tot + 2 * ((tot + 31) / 32)) * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(upsample_iter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fneus_cuda_error((int)e);
  }
  prof_begin(PC_SAMPLING, 0.0, (double)B * (tot * 16.0 + kp * 8.0 + k * 16.0 + 24.0), (cudaStream_t)stream);
  upsample_iter_kernel<<<cdiv(B, SAMP_WARPS), SAMP_WARPS * 32, smem, (cudaStream_t)stream>>>(
      rays_o, rays_d, z, sdf, B, n, prev_z, prev_sdf, kp, k, inv_s, u_table, z_out, sdf_out, new_z, pts_out);
  prof_end((cudaStream_t)stream);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

int fneus_inverse_cdf(const float* bins, const float* cdf, const float* u_table, long long B, int n, int k,
                      float* samples_out, long long* inds_out, void* stream) {
  if (B == 0 || k == 0) return FNEUS_OK;
  if (!bins || !cdf || !u_table || !samples_out) return FNEUS_ERR_NULL;
  if (B < 0 || n < 1 || k < 0 || n > 4096) return FNEUS_ERR_BAD_SHAPE;
  size_t smem = (size_t)SAMP_WARPS * 2 * n * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(inverse_cdf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fneus_cuda_error((int)e);
  }
  prof_begin(PC_SAMPLING, 0.0, (double)B * (n * 8.0 + k * 12.0), (cudaStream_t)stream);
  inverse_cdf_kernel<<<cdiv(B, SAMP_WARPS), SAMP_WARPS * 32, smem, (cudaStream_t)stream>>>(bins, cdf, u_table, B, n,
                                                                                         k, samples_out, inds_out);
  prof_end((cudaStream_t)stream);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

int fneus_merge_sorted(const float* z, const float* new_z, const float* sdf, const float* new_sdf, long long B, int n,
                       int k, float* z_out, float* sdf_out, void* stream) {
  if (B == 0) return FNEUS_OK;
  if (!z || !new_z || !z_out) return FNEUS_ERR_NULL;
  if ((sdf == nullptr) != (new_sdf == nullptr)) return FNEUS_ERR_NULL;
  if (sdf && !sdf_out) return FNEUS_ERR_NULL;
  if (B < 0 || n < 0 || k < 0 || n + k > 8192) return FNEUS_ERR_BAD_SHAPE;
  size_t smem = (size_t)SAMP_WARPS * (2 * (n + k) + 2 * ((n + k + 31) / 32)) * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(merge_sorted_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fneus_cuda_error((int)e);
  }
  prof_begin(PC_SAMPLING, 0.0, (double)B * (n + k) * (sdf ? 16.0 : 8.0), (cudaStream_t)stream);
  merge_sorted_kernel<<<cdiv(B, SAMP_WARPS), SAMP_WARPS * 32, smem, (cudaStream_t)stream>>>(z, new_z, sdf, new_sdf, B,
                                                                                        n, k, z_out, sdf_out);
  prof_end((cudaStream_t)stream);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

int fneus_core_geometry(const float* rays_o, const float* rays_d, const float* z, long long B, int n,
                        float sample_dist, float* dists, float* mid_z, float* pts, float* dirs, void* stream) {
  if (B == 0 || n == 0) return FNEUS_OK;
  if (!rays_o || !rays_d || !z || !pts) return FNEUS_ERR_NULL;
  if (B < 0 || n < 0) return FNEUS_ERR_BAD_SHAPE;
  long long total = B * n;
  prof_begin(PC_SAMPLING, 0.0, (double)total * 36.0, (cudaStream_t)stream);
  core_geometry_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, z, total, n, sample_dist,
                                                                          dists, mid_z, pts, dirs);
  prof_end((cudaStream_t)stream);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

int fneus_near_far(const float* rays_o, const float* rays_d, long long B, float* near, float* far, void* stream) {
  if (B == 0) return FNEUS_OK;
  if (!rays_o || !rays_d || !near || !far) return FNEUS_ERR_NULL;
  if (B < 0) return FNEUS_ERR_BAD_SHAPE;
  prof_begin(PC_SAMPLING, 0.0, (double)B * 32.0, (cudaStream_t)stream);
  near_far_kernel<<<cdiv(B, 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, B, near, far);
  prof_end((cudaStream_t)stream);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

int fneus_coarse_z(const float* near, const float* far, const float* lin, const float* rnd, long long B, int n,
                   float inv_n_samples, float* z, void* stream) {
  if (B == 0 || n == 0) return FNEUS_OK;
  if (!near || !far || !lin || !z) return FNEUS_ERR_NULL;
  if (B < 0 || n < 0) return FNEUS_ERR_BAD_SHAPE;
  prof_begin(PC_SAMPLING, 0.0, (double)B * n * 4.0, (cudaStream_t)stream);
  coarse_z_kernel<<<cdiv(B * n, 256), 256, 0, (cudaStream_t)stream>>>(near, far, lin, rnd, B * n, n, inv_n_samples, z);
  prof_end((cudaStream_t)stream);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

int fneus_hit_rows(const int* hit_idx, long long B, int n, long long* rows, void* stream) {
  if (B == 0) return FNEUS_OK;
  if (!hit_idx || !rows) return FNEUS_ERR_NULL;
  if (B < 0 || n < 2) return FNEUS_ERR_BAD_SHAPE;
  prof_begin(PC_SAMPLING, 0.0, (double)B * 20.0, (cudaStream_t)stream);
  hit_rows_kernel<<<cdiv(B, 256), 256, 0, (cudaStream_t)stream>>>(hit_idx, B, n, rows);
  prof_end((cudaStream_t)stream);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

int fneus_gen_rays(const float* px, const float* py, const float* intrinsics_inv, const float* pose,
                   const float* image, const float* mask, int H, int W, long long B, float* out10, float* near,
                   float* far, void* stream) {
  if (B == 0) return FNEUS_OK;
  if (!px || !py || !intrinsics_inv || !pose || !out10) return FNEUS_ERR_NULL;
  if ((near == nullptr) != (far == nullptr)) return FNEUS_ERR_NULL;
  if (B < 0 || ((image || mask) && (H < 1 || W < 1))) return FNEUS_ERR_BAD_SHAPE;
  prof_begin(PC_SAMPLING, 0.0, (double)B * 56.0, (cudaStream_t)stream);
  gen_rays_kernel<<<cdiv(B, 256), 256, 0, (cudaStream_t)stream>>>(px, py, intrinsics_inv, pose, image, mask, H, W, B,
                                                                   out10, near, far);
  prof_end((cudaStream_t)stream);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

}  // extern "C"

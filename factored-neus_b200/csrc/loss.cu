// Per-ray tail of the training step: surface-colour blend of the two bracketing RefColor evaluations
// (renderer.py:328-343) and the stage-1 loss with its gradients (exp_runner.py:134-177).  These are [B]-sized
// element-wise chains and reductions; each is ONE launch here instead of dozens of framework kernels.
#include "fneus_common.cuh"
#include "prof.cuh"

namespace fneus {
int num_sms();

// ---- surface blend: out_k = hit ? (c0_k w0 + c1_k w1) / (w0 + w1) : 1, for k over the three colour sets --------------
__global__ void surface_blend_fwd_kernel(const float* __restrict__ c_rgb, const float* __restrict__ c_spec,
                                         const float* __restrict__ c_diff, const float* __restrict__ w_pair,
                                         const int* __restrict__ hit_idx, int B, float* __restrict__ o_rgb,
                                         float* __restrict__ o_spec, float* __restrict__ o_diff) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 3) return;
  const int b = i / 3, k = i - b * 3;
  const bool hit = hit_idx[b] >= 0;
  const float w0 = w_pair[2 * b], w1 = w_pair[2 * b + 1], den = w0 + w1;
  const float* cs[3] = {c_rgb, c_spec, c_diff};
  float* os[3] = {o_rgb, o_spec, o_diff};
#pragma unroll
  for (int s = 0; s < 3; s++) {
    const float c0 = cs[s][(2 * b) * 3 + k], c1 = cs[s][(2 * b + 1) * 3 + k];
    os[s][i] = hit ? (c0 * w0 + c1 * w1) / den : 1.f;
  }
}
// one thread per ray: d_c (both rows, three sets) and d_w_pair
__global__ void surface_blend_bwd_kernel(const float* __restrict__ c_rgb, const float* __restrict__ c_spec,
                                         const float* __restrict__ c_diff, const float* __restrict__ w_pair,
                                         const int* __restrict__ hit_idx, int B, const float* __restrict__ g_rgb,
                                         const float* __restrict__ g_spec, const float* __restrict__ g_diff,
                                         float* __restrict__ d_rgb, float* __restrict__ d_spec,
                                         float* __restrict__ d_diff, float* __restrict__ d_w) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const bool hit = hit_idx[b] >= 0;
  const float w0 = w_pair[2 * b], w1 = w_pair[2 * b + 1], den = w0 + w1;
  const float* cs[3] = {c_rgb, c_spec, c_diff};
  const float* gs[3] = {g_rgb, g_spec, g_diff};
  float* ds[3] = {d_rgb, d_spec, d_diff};
  float dw0 = 0.f, dw1 = 0.f;
#pragma unroll
  for (int s = 0; s < 3; s++) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const float g = (hit && gs[s] != nullptr) ? gs[s][b * 3 + k] : 0.f;
      const float c0 = cs[s][(2 * b) * 3 + k], c1 = cs[s][(2 * b + 1) * 3 + k];
      float a0 = 0.f, a1 = 0.f;
      if (hit) {
        // out = (c0 w0 + c1 w1) / den : d/dc0 = w0/den, d/dw0 = c0/den - num/den^2
        const float num = c0 * w0 + c1 * w1;
        const float gd = g / den;
        a0 = gd * w0; a1 = gd * w1;
        const float t = gd * (num / den);
        dw0 += gd * c0 - t;
        dw1 += gd * c1 - t;
      }
      ds[s][(2 * b) * 3 + k] = a0;
      ds[s][(2 * b + 1) * 3 + k] = a1;
    }
  }
  d_w[2 * b] = dw0;
  d_w[2 * b + 1] = dw1;
}

// ---- loss normalisers: den = [sum mask, sum mask*hit, eik_den, B] -----------------------------------------------------
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += red[i];
  return t;
}
__device__ __forceinline__ float mask_of(const float* mask, int b, int use_mask) {
  return use_mask ? (mask[b] > 0.5f ? 1.f : 0.f) : 1.f;
}
__global__ void loss_norms_kernel(const float* __restrict__ mask, const int* __restrict__ hit_idx,
                                  const float* __restrict__ eik_den, int B, int use_mask, float* __restrict__ den) {
  __shared__ float red[32];
  float s0 = 0.f, s1 = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float m = mask_of(mask, b, use_mask);
    s0 += m;
    s1 += hit_idx[b] >= 0 ? m : 0.f;
  }
  s0 = block_sum(s0, red);
  s1 = block_sum(s1, red);
  if (threadIdx.x == 0) { den[0] = s0; den[1] = s1; den[2] = *eik_den; den[3] = (float)B; }
}

// ---- loss + gradients (single CTA; B is a ray batch) ------------------------------------------------------------------
// parts = [loss, color_loss, surface_loss, eikonal_loss, mask_loss]
__global__ void stage1_loss_kernel(const float* __restrict__ color, const float* __restrict__ surf,
                                   const float* __restrict__ wsum, const float* __restrict__ true_rgb,
                                   const float* __restrict__ mask, const int* __restrict__ hit_idx,
                                   const float* __restrict__ eik_num, const float* __restrict__ den, int B, int use_mask,
                                   float surface_w, float igr_w, float mask_w, float* __restrict__ parts,
                                   float* __restrict__ d_color, float* __restrict__ d_surf, float* __restrict__ d_wsum,
                                   float* __restrict__ d_eik_num) {
  __shared__ float red[32];
  const float mask_sum = den[0] + 1e-5f, mask_sdf_sum = den[1] + 1e-5f, relax_sum = den[2] + 1e-5f, n_global = den[3];
  float lc = 0.f, ls = 0.f, lm = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float m = mask_of(mask, b, use_mask);
    const float h = hit_idx[b] >= 0 ? 1.f : 0.f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const float t = true_rgb[b * 3 + k];
      const float ec = (color[b * 3 + k] - t) * m;
      lc += fabsf(ec);
      d_color[b * 3 + k] = (ec > 0.f ? 1.f : (ec < 0.f ? -1.f : 0.f)) * m / mask_sum;
      const float es = surface_w * (surf[b * 3 + k] - t) * m * h;
      ls += fabsf(es);
      d_surf[b * 3 + k] = (es > 0.f ? 1.f : (es < 0.f ? -1.f : 0.f)) * surface_w * m * h / mask_sdf_sum;
    }
    // binary cross entropy on the clipped opacity (torch clamps each log at -100)
    const float w = wsum[b];
    const float wc = fminf(fmaxf(w, 1e-3f), 1.f - 1e-3f);
    lm += -(m * fmaxf(logf(wc), -100.f) + (1.f - m) * fmaxf(logf(1.f - wc), -100.f));
    const float pass = (w >= 1e-3f && w <= 1.f - 1e-3f) ? 1.f : 0.f;
    d_wsum[b] = mask_w * pass * (-(m / wc) + (1.f - m) / (1.f - wc)) / n_global;
  }
  lc = block_sum(lc, red);
  ls = block_sum(ls, red);
  lm = block_sum(lm, red);
  if (threadIdx.x == 0) {
    const float color_loss = lc / mask_sum, surf_loss = ls / mask_sdf_sum, eik_loss = *eik_num / relax_sum,
                mask_loss = lm / n_global;
    parts[0] = color_loss + surf_loss + eik_loss * igr_w + mask_loss * mask_w;
    parts[1] = color_loss; parts[2] = surf_loss; parts[3] = eik_loss; parts[4] = mask_loss;
    *d_eik_num = igr_w / relax_sum;
  }
}


// ---- optimiser of the stage-1 step: torch.optim.Adam (exp_runner.py:118, default betas / eps, no weight decay) over ONE
// flat FP32 parameter buffer, with the reference's warm-up / cosine learning-rate factor (exp_runner.py:229-238)
// evaluated on the device from a device-resident iteration counter, so the step is CUDA-graph capturable without any
// host-written scalar.  state4 = [iterations done, lr used by the last step, 1 - beta1^t, 1 - beta2^t].
__global__ void adam_tick_kernel(float* __restrict__ state4, float base_lr, float lr_alpha, float warm_up_end,
                                 float end_iter, float beta1, float beta2) {
  const float it = state4[0];
  float f;
  if (it < warm_up_end) f = it / warm_up_end;
  else {
    const float prog = (it - warm_up_end) / fmaxf(1.f, end_iter - warm_up_end);
    f = (cospif(prog) + 1.f) * 0.5f * (1.f - lr_alpha) + lr_alpha;
  }
  const float t = it + 1.f;
  state4[0] = t;
  state4[1] = base_lr * f;
  state4[2] = 1.f - powf(beta1, t);
  state4[3] = 1.f - powf(beta2, t);
}
__global__ void adam_step_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, long long n4, long long n, const float* __restrict__ state4,
                                 float beta1, float beta2, float eps, float gscale, int zero_grad) {
  const float lr = state4[1], bc1 = state4[2], bc2 = state4[3];
  const float step_size = lr / bc1, rsq_bc2 = 1.f / sqrtf(bc2);
  auto upd = [&](float& pp, float& gg, float& mm, float& vv) {
    const float gr = gg * gscale;
    mm = beta1 * mm + (1.f - beta1) * gr;
    vv = beta2 * vv + (1.f - beta2) * gr * gr;
    pp -= step_size * (mm / (sqrtf(vv) * rsq_bc2 + eps));
    if (zero_grad) gg = 0.f;
  };
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 P = reinterpret_cast<float4*>(p)[i], G = reinterpret_cast<float4*>(g)[i];
    float4 Mv = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
    upd(P.x, G.x, Mv.x, V.x); upd(P.y, G.y, Mv.y, V.y); upd(P.z, G.z, Mv.z, V.z); upd(P.w, G.w, Mv.w, V.w);
    reinterpret_cast<float4*>(p)[i] = P;
    reinterpret_cast<float4*>(m)[i] = Mv;
    reinterpret_cast<float4*>(v)[i] = V;
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = G;
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    upd(p[i], g[i], m[i], v[i]);
}

// ---- stage-2 loss (lvis.py:163-170), one block: ------------------------------------------------------------------
//   lvis_loss = sum |gt_lvis - pre_lvis| / (k n_hit + 1e-6)                 (difference NOT masked: rows without a hit are
//                                                                            ones on both sides in the reference)
//   rad_loss  = sum |(gt_rad - pre_rad) * hit| / (3 k n_hit + 1e-6)
// parts3 = [loss, lvis_loss, rad_loss]; d_pre_lvis / d_pre_rad = d loss / d prediction (sign / denominator).
// den2: optional [k n_hit + 1e-6, 3 k n_hit + 1e-6] over the WHOLE batch (ray-sharded data parallelism), else local.
__global__ void stage2_loss_kernel(const float* __restrict__ gt_lvis, const float* __restrict__ pre_lvis,
                                   const float* __restrict__ gt_rad, const float* __restrict__ pre_rad,
                                   const int* __restrict__ hit_idx, const float* __restrict__ den2, int B, int k,
                                   float* __restrict__ parts3, float* __restrict__ d_pre_lvis,
                                   float* __restrict__ d_pre_rad) {
  __shared__ float red[3][32];
  __shared__ float tot[3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float nh = 0.f, sl = 0.f, sr = 0.f;
  for (int b = tid; b < B; b += blockDim.x) {
    const bool hit = hit_idx[b] >= 0;
    nh += hit ? 1.f : 0.f;
    for (int j = 0; j < k; j++) {
      sl += fabsf(gt_lvis[b * k + j] - pre_lvis[b * k + j]);
      if (hit)
        for (int c = 0; c < 3; c++) sr += fabsf(gt_rad[(b * k + j) * 3 + c] - pre_rad[(b * k + j) * 3 + c]);
    }
  }
  float v[3] = {nh, sl, sr};
#pragma unroll
  for (int q = 0; q < 3; q++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
    if (lane == 0) red[q][warp] = v[q];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int q = 0; q < 3; q++) {
      float x = lane < (int)(blockDim.x >> 5) ? red[q][lane] : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      if (lane == 0) tot[q] = x;
    }
  }
  __syncthreads();
  const float dl = den2 ? den2[0] : (float)k * tot[0] + 1e-6f;
  const float dr = den2 ? den2[1] : 3.f * (float)k * tot[0] + 1e-6f;
  if (tid == 0) {
    const float a = tot[1] / dl, r = tot[2] / dr;
    parts3[0] = a + r; parts3[1] = a; parts3[2] = r;
  }
  auto sgn = [](float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); };
  for (int i = tid; i < B * k; i += blockDim.x) {
    const bool hit = hit_idx[i / k] >= 0;
    if (d_pre_lvis) d_pre_lvis[i] = sgn(pre_lvis[i] - gt_lvis[i]) / dl;
    if (d_pre_rad)
      for (int c = 0; c < 3; c++) d_pre_rad[i * 3 + c] = hit ? sgn(pre_rad[i * 3 + c] - gt_rad[i * 3 + c]) / dr : 0.f;
  }
}

// ---- step glue: the small per-ray / per-scalar pieces around the render path that the reference leaves to ATen ----------
// (each of them was 1-7 launches of a few microseconds inside the 2 ms step)
// batch [B,10] = (rays_o, rays_d, rgb, mask) of Dataset.gen_random_rays_at (dataset.py:133-151) -> four dense tensors
__global__ void split_batch_kernel(const float* __restrict__ batch, long long B, float* __restrict__ ro,
                                   float* __restrict__ rd, float* __restrict__ rgb, float* __restrict__ mask) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 10) return;
  const long long b = i / 10;
  const int c = (int)(i % 10);
  const float v = batch[i];
  if (c < 3) ro[b * 3 + c] = v;
  else if (c < 6) rd[b * 3 + c - 3] = v;
  else if (c < 9) rgb[b * 3 + c - 6] = v;
  else mask[b] = v;
}
// inv_s = clip(exp(10 variance), 1e-6, 1e6) (fields.py:267-268, renderer.py:238) and its derivative
__global__ void inv_s_kernel(const float* __restrict__ variance, const float* __restrict__ d_inv_s, float* __restrict__ out) {
  const float e = expf(variance[0] * 10.f);
  if (d_inv_s == nullptr) out[0] = fminf(fmaxf(e, 1e-6f), 1e6f);
  else out[0] = (e >= 1e-6f && e <= 1e6f) ? d_inv_s[0] * e * 10.f : 0.f;
}
// after the compositing kernel: eikonal totals over the rays (fixed order), their quotient (renderer.py:282) and the
// sign-change mask (renderer.py:286)
__global__ void composite_post_kernel(const float* __restrict__ eik, const int* __restrict__ hit_idx, int B,
                                      float* __restrict__ tot3, unsigned char* __restrict__ hit_mask) {
  __shared__ float red[32];
  float s0 = 0.f, s1 = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    s0 += eik[2 * b];
    s1 += eik[2 * b + 1];
    hit_mask[b] = hit_idx[b] >= 0 ? 1 : 0;
  }
  s0 = block_sum(s0, red);
  s1 = block_sum(s1, red);
  if (threadIdx.x == 0) { tot3[0] = s0; tot3[1] = s1; tot3[2] = s0 / (s1 + 1e-5f); }
}
// rows of three [N,3] tensors (the surface samples RefColor sees, renderer.py:296-327) and the scatter of a gradient
__global__ void gather_rows3_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                                    const long long* __restrict__ rows, long long n, float* __restrict__ oa,
                                    float* __restrict__ ob, float* __restrict__ oc) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 3) return;
  const long long src = rows[i / 3] * 3 + i % 3;
  oa[i] = a[src]; ob[i] = b[src]; oc[i] = c[src];
}
__global__ void scatter_rows3_kernel(const float* __restrict__ vals, const long long* __restrict__ rows, long long n,
                                     float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 3) return;
  out[rows[i / 3] * 3 + i % 3] += vals[i];
}

}  // namespace fneus

using namespace fneus;

extern "C" {

int fneus_surface_blend_fwd(const float* c_rgb, const float* c_spec, const float* c_diff, const float* w_pair,
                            const int* hit_idx, long long B, float* o_rgb, float* o_spec, float* o_diff, void* stream) {
  if (B == 0) return FNEUS_OK;
  if (!c_rgb || !c_spec || !c_diff || !w_pair || !hit_idx || !o_rgb || !o_spec || !o_diff) return FNEUS_ERR_NULL;
  if (B < 0 || B > (1 << 28)) return FNEUS_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  prof_begin(PC_COMPOSITE, 0.0, 0.0, st);
  surface_blend_fwd_kernel<<<cdiv(B * 3, 256), 256, 0, st>>>(c_rgb, c_spec, c_diff, w_pair, hit_idx, (int)B, o_rgb, o_spec,
                                                             o_diff);
  prof_end(st);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

int fneus_surface_blend_bwd(const float* c_rgb, const float* c_spec, const float* c_diff, const float* w_pair,
                            const int* hit_idx, long long B, const float* g_rgb, const float* g_spec,
                            const float* g_diff, float* d_rgb, float* d_spec, float* d_diff, float* d_w_pair,
                            void* stream) {
  if (B == 0) return FNEUS_OK;
  if (!c_rgb || !c_spec || !c_diff || !w_pair || !hit_idx || !d_rgb || !d_spec || !d_diff || !d_w_pair)
    return FNEUS_ERR_NULL;
  if (B < 0 || B > (1 << 28)) return FNEUS_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  prof_begin(PC_COMPOSITE, 0.0, 0.0, st);
  surface_blend_bwd_kernel<<<cdiv(B, 128), 128, 0, st>>>(c_rgb, c_spec, c_diff, w_pair, hit_idx, (int)B, g_rgb, g_spec,
                                                         g_diff, d_rgb, d_spec, d_diff, d_w_pair);
  prof_end(st);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

int fneus_loss_norms(const float* mask, const int* hit_idx, const float* eik_den, long long B, int use_mask,
                     float* den4, void* stream) {
  if (!mask || !hit_idx || !eik_den || !den4) return FNEUS_ERR_NULL;
  if (B < 0 || B > (1 << 28)) return FNEUS_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  prof_begin(PC_COMPOSITE, 0.0, 0.0, st);
  loss_norms_kernel<<<1, 1024, 0, st>>>(mask, hit_idx, eik_den, (int)B, use_mask, den4);
  prof_end(st);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

int fneus_stage1_loss(const float* color, const float* surface_color, const float* weight_sum, const float* true_rgb,
                      const float* mask, const int* hit_idx, const float* eik_num, const float* den4, long long B,
                      int use_mask, float surface_weight, float igr_weight, float mask_weight, float* parts5,
                      float* d_color, float* d_surface_color, float* d_weight_sum, float* d_eik_num, void* stream) {
  if (!color || !surface_color || !weight_sum || !true_rgb || !mask || !hit_idx || !eik_num || !den4 || !parts5 ||
      !d_color || !d_surface_color || !d_weight_sum || !d_eik_num)
    return FNEUS_ERR_NULL;
  if (B < 0 || B > (1 << 28)) return FNEUS_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  prof_begin(PC_COMPOSITE, 0.0, 0.0, st);
  stage1_loss_kernel<<<1, 1024, 0, st>>>(color, surface_color, weight_sum, true_rgb, mask, hit_idx, eik_num, den4, (int)B,
                                         use_mask, surface_weight, igr_weight, mask_weight, parts5, d_color,
                                         d_surface_color, d_weight_sum, d_eik_num);
  prof_end(st);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

int fneus_adam_step(float* p, float* g, float* m, float* v, long long n, float* state4, float base_lr, float lr_alpha,
                    float warm_up_end, float end_iter, float beta1, float beta2, float eps, float grad_scale,
                    int zero_grad, void* stream) {
  if (n == 0) return FNEUS_OK;
  if (!p || !g || !m || !v || !state4) return FNEUS_ERR_NULL;
  if (n < 0) return FNEUS_ERR_BAD_SHAPE;
  if ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
       reinterpret_cast<uintptr_t>(v)) & 15)
    return FNEUS_ERR_MISALIGNED;
  cudaStream_t st = (cudaStream_t)stream;
  prof_begin(PC_ELEMENTWISE, 0.0, 0.0, st);
  adam_tick_kernel<<<1, 1, 0, st>>>(state4, base_lr, lr_alpha, warm_up_end, end_iter, beta1, beta2);
  prof_end(st);
  const long long n4 = n / 4;
  long long blocks = cdiv(n4 > 0 ? n4 : 1, 256);
  if (blocks > 4LL * num_sms()) blocks = 4LL * num_sms();
  prof_begin(PC_ELEMENTWISE, 0.0, 32.0 * (double)n, st);
  adam_step_kernel<<<(int)blocks, 256, 0, st>>>(p, g, m, v, n4, n, state4, beta1, beta2, eps, grad_scale, zero_grad);
  prof_end(st);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

int fneus_stage2_loss(const float* gt_lvis, const float* pre_lvis, const float* gt_rad, const float* pre_rad,
                      const int* hit_idx, const float* den2, long long B, int k, float* parts3, float* d_pre_lvis,
                      float* d_pre_rad, void* stream) {
  if (!gt_lvis || !pre_lvis || !gt_rad || !pre_rad || !hit_idx || !parts3) return FNEUS_ERR_NULL;
  if (B < 0 || B > (1 << 26) || k < 1 || k > 64) return FNEUS_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  prof_begin(PC_COMPOSITE, 0.0, 0.0, st);
  stage2_loss_kernel<<<1, 1024, 0, st>>>(gt_lvis, pre_lvis, gt_rad, pre_rad, hit_idx, den2, (int)B, k, parts3, d_pre_lvis,
                                         d_pre_rad);
  prof_end(st);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

int fneus_split_batch(const float* batch, long long B, float* rays_o, float* rays_d, float* rgb, float* mask, void* stream) {
  if (B == 0) return FNEUS_OK;
  if (!batch || !rays_o || !rays_d || !rgb || !mask) return FNEUS_ERR_NULL;
  if (B < 0) return FNEUS_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  prof_begin(PC_ELEMENTWISE, 0.0, 0.0, st);
  split_batch_kernel<<<(int)cdiv(B * 10, 256), 256, 0, st>>>(batch, B, rays_o, rays_d, rgb, mask);
  prof_end(st);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

int fneus_inv_s(const float* variance, const float* d_inv_s, float* out, void* stream) {
  if (!variance || !out) return FNEUS_ERR_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  prof_begin(PC_ELEMENTWISE, 0.0, 0.0, st);
  inv_s_kernel<<<1, 1, 0, st>>>(variance, d_inv_s, out);
  prof_end(st);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

int fneus_composite_post(const float* eik, const int* hit_idx, long long B, float* tot3, unsigned char* hit_mask,
                         void* stream) {
  if (!eik || !hit_idx || !tot3 || !hit_mask) return FNEUS_ERR_NULL;
  if (B < 0 || B > (1 << 28)) return FNEUS_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  prof_begin(PC_COMPOSITE, 0.0, 0.0, st);
  composite_post_kernel<<<1, 1024, 0, st>>>(eik, hit_idx, (int)B, tot3, hit_mask);
  prof_end(st);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

int fneus_gather_rows3(const float* a, const float* b, const float* c, const long long* rows, long long n, float* oa,
                       float* ob, float* oc, void* stream) {
  if (n == 0) return FNEUS_OK;
  if (!a || !b || !c || !rows || !oa || !ob || !oc) return FNEUS_ERR_NULL;
  if (n < 0) return FNEUS_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  prof_begin(PC_ELEMENTWISE, 0.0, 0.0, st);
  gather_rows3_kernel<<<(int)cdiv(n * 3, 256), 256, 0, st>>>(a, b, c, rows, n, oa, ob, oc);
  prof_end(st);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

int fneus_scatter_rows3(const float* vals, const long long* rows, long long n, float* out, void* stream) {
  if (n == 0) return FNEUS_OK;
  if (!vals || !rows || !out) return FNEUS_ERR_NULL;
  if (n < 0) return FNEUS_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  prof_begin(PC_ELEMENTWISE, 0.0, 0.0, st);
  scatter_rows3_kernel<<<(int)cdiv(n * 3, 256), 256, 0, st>>>(vals, rows, n, out);
  prof_end(st);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

}  // extern "C"

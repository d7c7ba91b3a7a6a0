/* fneus.h -- C ABI of libfneus_b200.so: the B200-native (sm_100a) kernels behind the Factored-NeuS
 * per-ray volume-rendering hot path.
 *
 * The reference has no FFI layer: its seam is the Python class NeuSRenderer (models/renderer.py:80-110) and
 * the nn.Module fields (models/fields.py).  Each entry point below replaces the ATen op chain of one
 * reference function (cited per declaration) and is what the reference-side binding (INTEGRATION.md:
 * a ctypes stub inside models/renderer.py / models/fields.py) would call.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer to FP32 unless stated; tensors are row-major and contiguous;
 *  - the callee never allocates, frees or synchronises; workspaces are sized by the *_floats queries and
 *    owned by the caller; every call takes the CUDA stream to launch on (void* = cudaStream_t);
 *  - return value: 0 = ok, otherwise an fneus_status (never throws, never exits);
 *    fneus_status_string() decodes it;
 *  - weight packs are flat FP32 buffers of EFFECTIVE weights (weight-norm g*v/|v| already applied by the
 *    host), layer after layer: W_l [out_l, in_l] row-major then b_l [out_l]; gradient packs have the same
 *    layout and are ACCUMULATED into (the caller zero-fills them).
 */
#ifndef FNEUS_H_
#define FNEUS_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  FNEUS_OK = 0,
  FNEUS_ERR_BAD_SHAPE = 1,
  FNEUS_ERR_MISALIGNED = 2,
  FNEUS_ERR_UNSUPPORTED = 3,
  FNEUS_ERR_NULL = 4,
  FNEUS_ERR_WORKSPACE = 5,
  FNEUS_ERR_CUDA_BASE = 1000 /* 1000 + cudaError_t */
} fneus_status;

const char* fneus_status_string(int status);
int fneus_abi_version(void);
int fneus_num_sms(void);

/* Compute precision of the dense layers.  It is a field of every network configuration struct (no process-global
 * state: two renderers of one process may run different precisions, and a call is fully described by its arguments):
 *   FNEUS_PREC_FP32: FP32 on CUDA cores -- the exactness anchor, <= 1e-4 vs the reference;
 *   FNEUS_PREC_TC  : tcgen05 tensor cores, FP32 accumulation in TMEM; forward-path operands FP16 (11-bit significand),
 *                    backward-path operands BF16 -- <= 2e-2 and within 0.1 dB PSNR of the FP32 path after 1k iterations. */
typedef enum { FNEUS_PREC_FP32 = 0, FNEUS_PREC_TC = 1 } fneus_precision;

/* ---- development hooks (NOT part of the production surface: process-wide switches for bisecting and profiling) --------
 * Bisect switches of the tensor-core kernels (results may be wrong when non-zero). */
int fneus_debug_hang_record(unsigned long long* out64); /* wait watchdog: see csrc/api.cu (development hook) */
int fneus_debug_flags(int flags);
/* debug: copy the first n entries (n <= 8192) of the fused-chain timeline buffer (clock stamps, flag bit 6) */
int fneus_debug_timeline(unsigned long long* host_dst, int n);
/* Test hook: one raw dense-layer contraction in the given precision.  kind 0: C[M,N] = A[M,K] W[N,K]^T
 * + bias; kind 1: C[M,N] = A[M,K] W[K,N]; kind 2: C[N,K] += Y[M,N]^T A[M,K], bias[N] += colsum(Y) (W := Y). */
int fneus_debug_gemm(int precision, int kind, const float* A, int lda, const float* W, int ldw, float* bias, long long M,
                     int N, int K, float* C, int ldc, void* stream);

/* Profiling hooks used by bench.py: when enabled every kernel launch is bracketed by CUDA events on its own
 * stream.  fneus_prof_collect synchronises those events and ADDS per-class milliseconds, launch counts and
 * algorithmic flops/bytes into arrays of length fneus_prof_classes() (classes: 0 gemm fwd, 1 gemm bwd-data,
 * 2 gemm bwd-weight, 3 sampling, 4 compositing, 5 elementwise, 6 single-layer tensor-core GEMMs, 7 SDF forward chain,
 * 8 SDF backward chain, 9 ReLU chains, 10 grouped weight gradients), then resets.  Development hook like the above. */
int fneus_prof_classes(void);
int fneus_prof_enable(int on);
int fneus_prof_collect(double* ms, long long* launches, double* flops, double* bytes);

/* ---- flat weight packs (host glue of fields.py:67-68,143-144 weight_norm + the per-layer parameter tensors) -----
 * One launch builds a network's flat pack from its parameter tensors: segment i is either a weight-normalised
 * matrix (g[i] [rows], v[i] [rows,cols] -> W = g*v/|v|_row) or a plain copy (g[i] == NULL) of v[i], written at
 * flat + off[i].  The arrays are HOST arrays of DEVICE pointers.  fneus_pack_bwd turns the flat gradient pack
 * into dg[i] / dv[i] (NULL = skip), overwriting or (accumulate != 0) adding to them. */
int fneus_pack_max_segments(void);
int fneus_pack_fwd(int nseg, const float* const* g, const float* const* v, const long long* off, const int* rows,
                   const int* cols, float* flat, void* stream);
int fneus_pack_bwd(int nseg, const float* const* g, const float* const* v, float* const* dg, float* const* dv,
                   const long long* off, const int* rows, const int* cols, const float* dflat, int accumulate,
                   void* stream);

/* ---- SDF network (fields.py:9-111) ------------------------------------------------------------ */
typedef struct {
  int d_in;       /* 3 */
  int d_hidden;   /* 256 */
  int n_layers;   /* 8 hidden layers -> n_layers+1 linears */
  int d_out;      /* 257: col 0 = sdf, 1.. = feature */
  int multires;   /* 6 */
  int skip_layer; /* 4: input of this linear is cat([h, PE(x)])/sqrt(2); -1 = none */
  float scale;    /* 1.0 */
  float beta;     /* Softplus beta = 100 */
  int precision; /* FNEUS_PREC_FP32 (CUDA-core exactness anchor) or FNEUS_PREC_TC (tcgen05 tensor cores) */
  int feat_image; /* 0: features / feature gradients are FP32 rows [n, d_out-1].  1 (FNEUS_PREC_TC, d_out-1 == 256, value +
                   * normal graph): `feat_out` of fneus_sdf_fwd / fneus_sdf_fwd_grad is an FP16 operand image and `d_feat` of
                   * fneus_sdf_bwd a BF16 one (fneus_image_bytes(n, 256) bytes each) -- the layout the colour chain's first
                   * MMA operand and the weight-gradient kernel read as they are; see fneus_image_* below */
} fneus_sdf_cfg;

long long fneus_sdf_pack_floats(const fneus_sdf_cfg* cfg);
long long fneus_sdf_saved_floats(const fneus_sdf_cfg* cfg, long long n_points);
long long fneus_sdf_scratch_floats(const fneus_sdf_cfg* cfg, long long n_points);

/* SDFNetwork.forward / .sdf (fields.py:74-95), no graph.  sdf_out [n]; feat_out [n, d_out-1] or NULL
 * (sdf only).  Points are processed in chunks that fit scratch_floats. */
int fneus_sdf_fwd(const fneus_sdf_cfg* cfg, const float* wpack, const float* x, long long n_points,
                  float* sdf_out, float* feat_out, float* scratch, long long scratch_floats, void* stream);

/* extract_fields (renderer.py:14-29): u[ix,iy,iz] = -sdf(ax[ix], ay[iy], az[iz]) for ix in [ix0, ix1) (an x-slab,
 * the unit of multi-GPU sharding); ax/ay/az are the per-axis torch.linspace tables on the device;
 * u_out [(ix1-ix0)*ny*nz]. */
int fneus_sdf_grid(const fneus_sdf_cfg* cfg, const float* wpack, const float* ax, const float* ay, const float* az,
                   int nx, int ny, int nz, int ix0, int ix1, float* u_out, float* scratch,
                   long long scratch_floats, void* stream);

/* SDFNetwork.forward + .gradient (fields.py:74-111) in one pass: value, feature and the analytic
 * normal d sdf/d x [n,3] (normal_out NULL = value only); `saved` keeps the activations the backward needs. */
int fneus_sdf_fwd_grad(const fneus_sdf_cfg* cfg, const float* wpack, const float* x, long long n_points,
                       float* sdf_out, float* feat_out, float* normal_out, float* saved, float* scratch,
                       void* stream);

/* Backward of fneus_sdf_fwd_grad incl. the double-backward through the normal (autograd of fields.py:100-111
 * with create_graph=True).  Any of d_sdf [n], d_feat [n,d_out-1], d_normal [n,3] may be NULL.
 * Destroys `saved`. */
int fneus_sdf_bwd(const fneus_sdf_cfg* cfg, const float* wpack, const float* x, long long n_points,
                  const float* d_sdf, const float* d_feat, const float* d_normal, float* saved, float* scratch,
                  float* d_wpack, void* stream);

/* ---- ReLU MLPs: RenderingNetwork (fields.py:114-175), RefColor (fields.py:271-335), NeRF (fields.py:178-259)
 * share one description: first-layer input = [generated block | feature block]. ------------------- */
typedef struct {
  int d_feature;     /* 256 */
  int d_hidden;      /* 256 */
  int n_layers;      /* 4 hidden -> n_layers+1 linears */
  int d_out;         /* 3 */
  int multires_view; /* 4 */
  int precision; /* FNEUS_PREC_FP32 (CUDA-core exactness anchor) or FNEUS_PREC_TC (tcgen05 tensor cores) */
  int feat_image; /* 1: `feats` is the FP16 operand image written by the SDF chain (fneus_sdf_cfg.feat_image) and `d_feats`
                   * of fneus_color_bwd leaves as a BF16 operand image; 0: FP32 rows [n, d_feature] */
} fneus_color_cfg;

long long fneus_color_pack_floats(const fneus_color_cfg* cfg);
long long fneus_color_saved_floats(const fneus_color_cfg* cfg, long long n_points);
long long fneus_color_scratch_floats(const fneus_color_cfg* cfg, long long n_points);

/* RenderingNetwork.forward, mode 'idr', squeeze_out (fields.py:150-175): rgb_out [n,3]. saved may be NULL
 * (inference; scratch then also holds the activations). */
int fneus_color_fwd(const fneus_color_cfg* cfg, const float* wpack, const float* points, const float* normals,
                    const float* view_dirs, const float* feats, long long n_points, float* rgb_out,
                    float* saved, float* scratch, void* stream);
/* Backward: d_rgb [n,3] -> d_normals [n,3], d_feats [n,d_feature] (overwritten), d_wpack (accumulated). */
int fneus_color_bwd(const fneus_color_cfg* cfg, const float* wpack, const float* points, const float* normals,
                    const float* view_dirs, const float* feats, long long n_points, const float* rgb,
                    const float* d_rgb, float* d_normals, float* d_feats, float* saved, float* scratch,
                    float* d_wpack, void* stream);

/* ---- step glue: the small per-ray / per-scalar pieces around the render path that the reference leaves to ATen --------
 * split_batch: [B,10] rows of Dataset.gen_random_rays_at (dataset.py:133-151) -> rays_o, rays_d, rgb [B,3], mask [B].
 * inv_s: out = clip(exp(10 variance), 1e-6, 1e6) (fields.py:267-268, renderer.py:238) when d_inv_s is NULL, else the
 *        gradient d_variance = d_inv_s * 10 exp(10 variance) inside the clip range.
 * composite_post: eik [B,2] (per-ray numerator / denominator of renderer.py:282) -> tot3 = (sum num, sum den,
 *        num / (den + 1e-5)) in a fixed summation order; hit_mask[b] = hit_idx[b] >= 0 (renderer.py:286).
 * gather_rows3 / scatter_rows3: rows of three [N,3] tensors (points, directions, normals of the surface samples,
 *        renderer.py:296-327) in one launch; scatter adds vals [n,3] into out [N,3] at distinct rows. */
int fneus_split_batch(const float* batch, long long B, float* rays_o, float* rays_d, float* rgb, float* mask, void* stream);
int fneus_inv_s(const float* variance, const float* d_inv_s, float* out, void* stream);
int fneus_composite_post(const float* eik, const int* hit_idx, long long B, float* tot3, unsigned char* hit_mask,
                         void* stream);
int fneus_gather_rows3(const float* a, const float* b, const float* c, const long long* rows, long long n, float* oa,
                       float* ob, float* oc, void* stream);
int fneus_scatter_rows3(const float* vals, const long long* rows, long long n, float* out, void* stream);

/* ---- operand images: the hand-over format of `feature_vector` between the SDF and the colour network on the tensor-core
 * path (renderer.py:225-232 passes it as an FP32 [n,256] tensor; here it never takes that form in HBM).  An image holds
 * ceil(n/128) row tiles x (n_cols/64) blocks of 128 rows x 64 columns of 16-bit elements in the 128-byte-swizzled K-major
 * shared-memory layout of tcgen05 operands.  gather: FP32 rows out of an image (RefColor reads 2 rows per ray,
 * renderer.py:296-327); scatter_add: FP32 rows added into a BF16 image (their gradient); rows must be distinct. */
/* 1 when a configuration can take / produce the image form (tensor-core precision, chain-kernel shapes), else 0 */
int fneus_sdf_feat_image_ok(const fneus_sdf_cfg* cfg);
int fneus_color_feat_image_ok(const fneus_color_cfg* cfg);
long long fneus_image_bytes(long long n_rows, int n_cols);
int fneus_image_gather_rows(const void* image, int is_fp16, int n_cols, const long long* rows, long long n_sel,
                            float* out /* [n_sel, n_cols] */, void* stream);
int fneus_image_scatter_add_rows(void* image_bf16, int n_cols, const long long* rows, long long n_sel,
                                 const float* vals /* [n_sel, n_cols] */, void* stream);

/* RefColor.forward (fields.py:303-335).  Pack order: net_cd.{0,2,4,6,8}, viewdir_mlp.{0..3}, net_cs.0.
 * Outputs rgb/specular_rgb/diffuse_rgb [n,3]. */
typedef struct {
  int d_feature; /* 256 */
  int d_hidden;  /* 256 */
  int precision; /* FNEUS_PREC_FP32 (CUDA-core exactness anchor) or FNEUS_PREC_TC (tcgen05 tensor cores) */
} fneus_ref_cfg;
long long fneus_ref_pack_floats(const fneus_ref_cfg* cfg);
long long fneus_ref_saved_floats(const fneus_ref_cfg* cfg, long long n_points);
long long fneus_ref_scratch_floats(const fneus_ref_cfg* cfg, long long n_points);
int fneus_ref_fwd(const fneus_ref_cfg* cfg, const float* wpack, const float* points, const float* feats,
                  const float* dirs, const float* normals, long long n_points, float* rgb_out, float* spec_out,
                  float* diff_out, float* saved, float* scratch, void* stream);
int fneus_ref_bwd(const fneus_ref_cfg* cfg, const float* wpack, const float* points, const float* feats,
                  const float* dirs, const float* normals, long long n_points, const float* d_rgb,
                  const float* d_spec, const float* d_diff, float* d_feats, float* d_normals, float* saved,
                  float* scratch, float* d_wpack, void* stream);

/* ---- generic positional-encoded ReLU MLP without input gradients: Lvis (fields.py:338-369) and the trunk of
 * IndirectLight (fields.py:372-399).  Input = [PE(in0) | PE(in1)] (n_inputs 1 or 2, include_input encodings), n_layers
 * hidden ReLU layers of d_hidden, output d_out with last_act 0 = linear, 1 = sigmoid. */
typedef struct {
  int n_inputs;
  int in_dim[2];
  int in_multires[2];
  int d_hidden;
  int n_layers;
  int d_out;
  int last_act;
  int precision; /* FNEUS_PREC_FP32 (CUDA-core exactness anchor) or FNEUS_PREC_TC (tcgen05 tensor cores) */
} fneus_mlp_cfg;
long long fneus_mlp_pack_floats(const fneus_mlp_cfg* cfg);
long long fneus_mlp_saved_floats(const fneus_mlp_cfg* cfg, long long n_points);
long long fneus_mlp_scratch_floats(const fneus_mlp_cfg* cfg, long long n_points);
int fneus_mlp_fwd(const fneus_mlp_cfg* cfg, const float* wpack, const float* in0, const float* in1,
                  long long n_points, float* out, float* saved, float* scratch, void* stream);
int fneus_mlp_bwd(const fneus_mlp_cfg* cfg, const float* wpack, const float* in0, const float* in1,
                  long long n_points, const float* out, const float* d_out, float* saved, float* scratch,
                  float* d_wpack, void* stream);

/* ---- outside NeRF (fields.py:178-259) + render_core_outside (renderer.py:112-149), womask configuration.
 * Pack order: pts_linears.0..D-1, head = [alpha_linear ; feature_linear] (W [1+W, W] then b [1+W]),
 * views_linears.0, rgb_linear.  No input gradients (sample positions carry none, renderer.py:426). */
typedef struct {
  int D;             /* 8 */
  int W;             /* 256 */
  int d_in;          /* 4 (inverted-sphere point) */
  int d_in_view;     /* 3 */
  int multires;      /* 10 */
  int multires_view; /* 4 */
  int skip;          /* 4: embedded input is concatenated after pts_linears[skip]; -1 = none */
  int precision; /* FNEUS_PREC_FP32 (CUDA-core exactness anchor) or FNEUS_PREC_TC (tcgen05 tensor cores) */
} fneus_nerf_cfg;
long long fneus_nerf_pack_floats(const fneus_nerf_cfg* cfg);
long long fneus_nerf_saved_floats(const fneus_nerf_cfg* cfg, long long n_points);
long long fneus_nerf_scratch_floats(const fneus_nerf_cfg* cfg, long long n_points);
/* NeRF.forward (use_viewdirs): pts [n,d_in], views [n,3] -> raw density [n], raw rgb [n,3]. */
int fneus_nerf_fwd(const fneus_nerf_cfg* cfg, const float* wpack, const float* pts, const float* views,
                   long long n_points, float* density_out, float* rgb_out, float* saved, void* stream);
int fneus_nerf_bwd(const fneus_nerf_cfg* cfg, const float* wpack, const float* pts, const float* views,
                   long long n_points, const float* d_density, const float* d_rgb, float* saved, float* scratch,
                   float* d_wpack, void* stream);
/* renderer.py:116-128: dists [B,n], pts4 = [p/|p|, 1/|p|] with |p| clipped to [1,1e10] ([B*n,4]), dirs [B*n,3] */
int fneus_outside_geometry(const float* rays_o, const float* rays_d, const float* z, long long n_rays, int n,
                           float sample_dist, float* dists, float* pts4, float* dirs, void* stream);
/* renderer.py:131-134: alpha = 1-exp(-softplus(density)*dists), color = sigmoid(rgb); and the backward. */
int fneus_outside_alpha_fwd(const float* density, const float* rgb_raw, const float* dists, long long total,
                            float* alpha, float* color, void* stream);
int fneus_outside_alpha_bwd(const float* density, const float* color, const float* dists, const float* d_alpha,
                            const float* d_color, long long total, float* d_density, float* d_rgb_raw,
                            void* stream);

/* ---- sampling (renderer.py:43-77,152-205, 391-447) -------------------------------------------- */
/* pts[b*n+j] = o[b] + d[b]*z[b,j]   (renderer.py:159,194,428) */
int fneus_ray_points(const float* rays_o, const float* rays_d, const float* z, long long n_rays, int n,
                     float* pts_out, void* stream);
/* One hierarchical up-sampling step (NeuSRenderer.up_sample + sample_pdf(det=True)): z, sdf [B,n] ->
 * new_z [B,k].  u_table [k] = torch.linspace(0.5/k, 1-0.5/k, k) supplied by the host (bit pattern of the
 * device's own linspace).  Optional debug outputs cdf_out [B,n], inds_out [B,k] (int64). */
int fneus_upsample_step(const float* rays_o, const float* rays_d, const float* z, const float* sdf,
                        long long n_rays, int n, int k, float inv_s, const float* u_table, float* new_z,
                        float* cdf_out, long long* inds_out, void* stream);
/* The same step with inv_s read from device memory: stage 2 samples at the LEARNED inv_s (calLvis.py:371-379), which
 * lives on the device; no host read-back, CUDA-graph capturable. */
int fneus_upsample_step_dev(const float* rays_o, const float* rays_d, const float* z, const float* sdf,
                            long long n_rays, int n, int k, const float* inv_s_dev, const float* u_table,
                            float* new_z, void* stream);
/* First sign change + secant root (renderer.py:588-602 in lvis_render, calLvis.py:180-196 in cal_firHit_rgb):
 * idx = first sample with sdf < 0 (the reference's argmin of sign(sdf) * (n - i); sign(0) = 0 is not a hit), valid iff
 * idx >= 1 and some sample of the ray lies inside the unit sphere (|pts| < 1).  hit_idx [B] int32 (-1 = no hit),
 * z_surf [B] = (s_lo z_hi - s_hi z_lo) / (s_lo - s_hi + 1e-10) over the bracketing mid-points, pts_surf [B,3] = o + d z
 * (fixed shapes: rays without a hit get the root of the clamped index).  With weights [B, ldw] also
 * lvis [B] = 1 - sum_i weights_i [|pts_i| < 1] (calLvis.py:387-392); any_inside [B] int32 = the reference's
 * inside_sphere_mask (renderer.py:556).  z_surf, pts_surf, weights, lvis, any_inside may be NULL. */
int fneus_first_hit_secant(const float* sdf, const float* mid_z, const float* pts, const float* rays_o,
                           const float* rays_d, const float* weights, int ldw, long long n_rays, int n,
                           int* hit_idx, float* z_surf, float* pts_surf, float* lvis, int* any_inside, void* stream);
/* One iteration of the hierarchical sampling loop (renderer.py:166-176) in one launch: merge the previous iteration's kp new
 * depths prev_z [B,kp] with their sdf values prev_sdf into the sorted rows z / sdf [B,n] (cat_z_vals, :191-205; results
 * z_out / sdf_out [B,n+kp]; kp = 0: no merge), up-sample k depths new_z [B,k] from the merged row (up_sample + sample_pdf,
 * :152-189, 43-77) and emit their positions pts_out [B,k,3] = o + d z (may be NULL).  Bit-identical to fneus_merge_sorted +
 * fneus_upsample_step + fneus_ray_points. */
int fneus_upsample_iter(const float* rays_o, const float* rays_d, const float* z, const float* sdf, long long B, int n,
                        const float* prev_z, const float* prev_sdf, int kp, int k, float inv_s, const float* u_table,
                        float* z_out, float* sdf_out, float* new_z, float* pts_out, void* stream);
/* Inverse-CDF alone (renderer.py:64-77): searchsorted(right=True) + interpolation on a SUPPLIED cdf. */
int fneus_inverse_cdf(const float* bins, const float* cdf, const float* u_table, long long n_rays, int n, int k,
                      float* samples_out, long long* inds_out, void* stream);
/* cat_z_vals (renderer.py:191-205) as a merge of two sorted lists; carries sdf along when both sdf and
 * new_sdf are given (NULL,NULL = last step / depth-only merge). */
int fneus_merge_sorted(const float* z, const float* new_z, const float* sdf, const float* new_sdf,
                       long long n_rays, int n, int k, float* z_out, float* sdf_out, void* stream);
/* Section geometry of render_core (renderer.py:224-236): dists, mid_z, mid points [B*n,3], dirs [B*n,3]. */
int fneus_core_geometry(const float* rays_o, const float* rays_d, const float* z, long long n_rays, int n,
                        float sample_dist, float* dists, float* mid_z, float* pts, float* dirs, void* stream);

/* ---- alpha / transmittance / compositing (renderer.py:245-274,350-372) -------------------------- */
/* n_in = inside samples (128), n_out = extra outside samples (0 or 32).  inv_s: DEVICE scalar
 * clip(exp(10*variance),1e-6,1e6) (renderer.py:245; no host sync).   bg_alpha [B,n_in+n_out],
 * bg_color [B,n_in+n_out,3] (NULL when n_out==0 and no background model), bg_rgb [3] or NULL.
 * Outputs: color [B,3], weights [B,n_in+n_out], weight_sum [B], weight_max [B], cdf [B,n_in],
 * inside [B,n_in], eik_part [B,2] (sum relax*(|g|-1)^2, sum relax), hit_idx [B] int32 (-1 = no surface hit),
 * w_pair [B,2] (inside-weights at hit_idx-1, hit_idx, +1e-5; 1 when no hit).
 * cos_anneal_dev: optional DEVICE scalar that overrides cos_anneal_ratio (exp_runner.py:223-227 moves it every iteration
 * of womask; read from the device it can change between replays of a captured step). */
int fneus_composite_fwd(const float* sdf, const float* normals, const float* rgb, const float* dists,
                        const float* pts, const float* rays_d, const float* bg_alpha, const float* bg_color,
                        const float* bg_rgb, long long n_rays, int n_in, int n_out, const float* inv_s,
                        float cos_anneal_ratio, const float* cos_anneal_dev, float* color, float* weights, float* weight_sum,
                        float* weight_max, float* cdf, float* inside, float* eik_part, int* hit_idx,
                        float* w_pair, void* stream);
/* Backward.  Upstream: d_color [B,3], d_weights [B,n_in+n_out] (NULL ok), d_weight_sum [B] (NULL ok),
 * d_w_pair [B,2] (NULL ok), d_eik scalar pointer (device, d loss / d gradient_error; NULL ok) with
 * eik_denom = sum relax + 1e-5 over the WHOLE batch (device scalar).
 * Outputs: d_sdf [B*n_in], d_normals [B*n_in,3], d_rgb [B*n_in,3], d_inv_s [B] (per-ray partials),
 * d_bg_alpha / d_bg_color (NULL ok). */
int fneus_composite_bwd(const float* sdf, const float* normals, const float* rgb, const float* dists,
                        const float* pts, const float* rays_d, const float* bg_alpha, const float* bg_color,
                        const float* bg_rgb, long long n_rays, int n_in, int n_out, const float* inv_s,
                        float cos_anneal_ratio, const float* cos_anneal_dev, const int* hit_idx, const float* d_color,
                        const float* d_weights, const float* d_weight_sum, const float* d_w_pair,
                        const float* d_eik, const float* eik_denom, float* d_sdf, float* d_normals,
                        float* d_rgb, float* d_inv_s, float* d_bg_alpha, float* d_bg_color, void* stream);

/* Rays of a pinhole camera from pixel coordinates, on the device (dataset.py:115-151: gen_rays_at and
 * gen_random_rays_at assemble them on the CPU and upload every iteration).  px, py [B] pixel coordinates (float);
 * intrinsics_inv, pose: 4x4 row-major; image, mask: [H,W,3] or NULL.  out10 [B,10] = (rays_o, rays_v, rgb, mask[...,0]);
 * near / far [B] = near_far_from_sphere (dataset.py:186-192), both or neither NULL. */
int fneus_gen_rays(const float* px, const float* py, const float* intrinsics_inv, const float* pose,
                   const float* image, const float* mask, int H, int W, long long n_rays, float* out10, float* near,
                   float* far, void* stream);

/* ---- stage-2 light visibility (calLvis.py:339-397: the ground-truth half of cal_indiLgt) as ONE call ----------------------
 * surf [m,3] surface points, dirs [m*n_dirs,3] secondary-ray directions (sample_dirs, calLvis.py:302-320; ray r starts
 * at surf[r / n_dirs]).  Per ray: n_coarse sdf evaluations on the shared depth table z_table [n_coarse]
 * (torch.linspace(0,1,n_coarse) of the device under test), n_imp importance depths by the inverse CDF at inv_s (DEVICE
 * scalar: the learned, clipped inv_s; u_table [n_imp] = linspace(0.5/n_imp, 1-0.5/n_imp, n_imp)), then on those n_imp
 * sections only (the reference discards the coarse ones): lvis_out [m*n_dirs] = 1 - sum_i w_i [|p_i| < 1]
 * (compute_weight, cos_anneal_ratio 0), hit_out [m*n_dirs] int32 = index of the first sign change (-1: none),
 * rgb_out [m*n_dirs,3] = colour network at the secant root of the hit (0 without a hit).
 * The rays are processed in chunks of rays_per_chunk (rounded down to whole surface points); ws must hold
 * fneus_lvis_trace_workspace_floats(.., rays_per_chunk, ..) floats (16-byte aligned).  No host synchronisation, no
 * allocation: CUDA-graph capturable.  Runs in sdf_cfg->precision. */
long long fneus_lvis_trace_workspace_floats(const fneus_sdf_cfg* sdf_cfg, const fneus_color_cfg* color_cfg,
                                            long long rays_per_chunk, int n_coarse, int n_imp);
int fneus_lvis_trace(const fneus_sdf_cfg* sdf_cfg, const float* sdf_wpack, const fneus_color_cfg* color_cfg,
                     const float* color_wpack, const float* surf, const float* dirs, long long m, int n_dirs,
                     int n_coarse, int n_imp, const float* inv_s, const float* z_table, const float* u_table,
                     float* lvis_out, float* rgb_out, int* hit_out, float* ws, long long ws_floats,
                     long long rays_per_chunk, void* stream);

/* ---- marching cubes on the device (renderer.py:32-40 hands the grid to PyMCubes on the CPU) -------------------------------
 * u [nx,ny,nz] (row-major, as extract_fields writes it); inside: u > isovalue.  Each grid point owns its +x/+y/+z edges.
 * fneus_mc_classify: vmask [npts] (bit a: the owned edge along axis a is crossed), vcount [npts] = popcount(vmask),
 * tcount [(nx-1)(ny-1)(nz-1)] = triangles of the cell's case (tri_count [256]).  The caller prefix-sums vcount / tcount
 * (inclusive, int64) and calls fneus_mc_emit: verts [V,3] in grid-index coordinates (linear interpolation on the edge, like
 * mcubes.marching_cubes), tris [T,3] int64 indices into verts (shared vertices); tri_edges [256, 3 * max_tris] lists the
 * cube edges of each case's triangles (-1 padded).  The case table is derived in factored-neus_b200/mcubes.py. */
int fneus_mc_classify(const float* u, int nx, int ny, int nz, float isovalue, const int* tri_count,
                      unsigned char* vmask, int* vcount, int* tcount, void* stream);
int fneus_mc_emit(const float* u, int nx, int ny, int nz, float isovalue, const int* tri_count, const int* tri_edges,
                  int max_tris, const unsigned char* vmask, const int* vcount, const long long* voff_incl,
                  const int* tcount, const long long* toff_incl, float* verts, long long* tris, void* stream);

/* ---- per-ray tail of the training step ----------------------------------------------------------------------------
 * Surface-colour blend of the two bracketing RefColor evaluations per ray (renderer.py:328-343): c_* are [2B,3]
 * (rows 2b, 2b+1), w_pair [B,2], hit_idx [B] (< 0: no sign change -> ones).  Backward: g_* may be NULL (no gradient). */
int fneus_surface_blend_fwd(const float* c_rgb, const float* c_spec, const float* c_diff, const float* w_pair,
                            const int* hit_idx, long long n_rays, float* o_rgb, float* o_spec, float* o_diff,
                            void* stream);
int fneus_surface_blend_bwd(const float* c_rgb, const float* c_spec, const float* c_diff, const float* w_pair,
                            const int* hit_idx, long long n_rays, const float* g_rgb, const float* g_spec,
                            const float* g_diff, float* d_rgb, float* d_spec, float* d_diff, float* d_w_pair,
                            void* stream);
/* Stage-1 loss (exp_runner.py:134-177) on a ray shard.  fneus_loss_norms writes den4 = [sum mask, sum mask*hit,
 * eik_den, n_rays] (all-reduce it across ray shards before the next call); fneus_stage1_loss writes
 * parts5 = [loss, color, surface, eikonal, mask] and d loss / d (color_fine, surface_color, weight_sum, eik_num). */
int fneus_loss_norms(const float* mask, const int* hit_idx, const float* eik_den, long long n_rays, int use_mask,
                     float* den4, void* stream);
int fneus_stage1_loss(const float* color, const float* surface_color, const float* weight_sum, const float* true_rgb,
                      const float* mask, const int* hit_idx, const float* eik_num, const float* den4, long long n_rays,
                      int use_mask, float surface_weight, float igr_weight, float mask_weight, float* parts5,
                      float* d_color, float* d_surface_color, float* d_weight_sum, float* d_eik_num, void* stream);

/* Ray set-up glue of the caller and of render (one launch each instead of ~10 framework kernels):
 *  fneus_near_far   dataset.py:186-192 near_far_from_sphere -> near/far [n_rays]
 *  fneus_coarse_z   renderer.py:395-408: z [n_rays, n] = near + (far - near) * lin[n] (+ (rnd[n_rays] - 0.5) * 2 * inv_n_samples
 *                   when rnd != NULL: the caller draws rnd with its own generator)
 *  fneus_hit_rows   renderer.py:296-303: rows [2 n_rays] (int64) = flat indices of samples idx-1, idx around the first
 *                   sign change (idx clamped to >= 1; rays without one are masked later by hit_idx < 0) */
int fneus_near_far(const float* rays_o, const float* rays_d, long long n_rays, float* near, float* far, void* stream);
int fneus_coarse_z(const float* near, const float* far, const float* lin, const float* rnd, long long n_rays, int n,
                   float inv_n_samples, float* z, void* stream);
int fneus_hit_rows(const int* hit_idx, long long n_rays, int n, long long* rows, void* stream);

/* Stage-2 loss (lvis.py:163-170): parts3 = [loss, lvis_loss, radiance_loss] with
 *   lvis_loss = sum |gt_lvis - pre_lvis| / (k n_hit + 1e-6)   (not masked: rows without a hit are ones on both sides)
 *   radiance_loss = sum |(gt_rad - pre_rad) hit| / (3 k n_hit + 1e-6)
 * gt_lvis, pre_lvis [B,k]; gt_rad, pre_rad [B,k,3]; hit_idx [B] int32 (< 0: no hit); den2 (device, optional) = the two
 * denominators over the whole batch when rays are sharded; d_pre_lvis / d_pre_rad = d loss / d prediction (NULL ok). */
int fneus_stage2_loss(const float* gt_lvis, const float* pre_lvis, const float* gt_rad, const float* pre_rad,
                      const int* hit_idx, const float* den2, long long n_rays, int k, float* parts3,
                      float* d_pre_lvis, float* d_pre_rad, void* stream);

/* Optimiser of the stage-1 step (exp_runner.py:118 torch.optim.Adam, :229-238 update_learning_rate): one fused Adam
 * update over flat FP32 buffers p/g/m/v [n] (16-byte aligned).  state4 (device) = [iterations done, lr of the last
 * step, 1-beta1^t, 1-beta2^t]; the call first advances it (lr = base_lr * warm-up/cosine factor of the iteration
 * count BEFORE the step, as the reference updates the rate after each step), then updates p.  g is multiplied by
 * grad_scale first and cleared afterwards when zero_grad != 0.  No host-written scalars: graph-capturable. */
int fneus_adam_step(float* p, float* g, float* m, float* v, long long n, float* state4, float base_lr, float lr_alpha,
                    float warm_up_end, float end_iter, float beta1, float beta2, float eps, float grad_scale,
                    int zero_grad, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FNEUS_H_ */

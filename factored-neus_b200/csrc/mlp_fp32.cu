// Network-level host orchestration + small elementwise kernels for the FP32 path:
// SDFNetwork (fields.py:9-111), RenderingNetwork (fields.py:114-175), RefColor (fields.py:271-335).
// Every dense layer is one launch of the SIMT GEMM engine (gemm_simt.cuh) with a fused epilogue;
// positional encodings are generated in the GEMM tile loaders.
#include "gemm_tc.cuh"
#include <string.h>
#include "sdf_fused.cuh"
#include "sdf_chain.cuh"
#include "prof.cuh"

namespace fneus {

static int g_num_sms = 0;
int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    g_num_sms = n > 0 ? n : 148;
  }
  return g_num_sms;
}

// ------------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------------
// C[m][col0 + j] = gen(m, j) * scale
__global__ void append_gen_cols_kernel(GenSpec g, float* C, int ldc, int col0, float scale, long long M) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = M * g.ncols;
  if (idx >= total) return;
  long long m = idx / g.ncols;
  int j = (int)(idx - m * g.ncols);
  mat_put(C, ldc, m, col0 + j, gen_eval(g, m, j) * scale);
}

// n[m][c] = sum_j [comp_j == c] dPE_j/dx_c (x) * g0[m][j]      (T0^T g0)
__global__ void normal_from_g0_kernel(const float* x, int d, int multires, float scale, const float* g0, int ldg,
                                      float* n_out, long long M) {
  long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  for (int c = 0; c < d; c++) {
    float v = x[m * d + c] * scale;
    float acc = g0[m * ldg + c];
    for (int k = 0; k < multires; k++) {
      float f = (float)(1u << k);
      float s, co;
      sincosf(v * f, &s, &co);
      acc += f * co * g0[m * ldg + d * (1 + 2 * k) + c] - f * s * g0[m * ldg + d * (2 + 2 * k) + c];
    }
    n_out[m * d + c] = acc;
  }
}

// out[k] += sum_m w[m]*wscale * X[m][k] (w == nullptr -> 1); osum += sum_m w[m]*wscale
__global__ void colsum_kernel(const float* X, int ldx, int K, const float* w, float wscale, float* out, float* osum,
                              long long M, int m_per_block) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  long long mbeg = (long long)blockIdx.y * m_per_block;
  long long mend = mbeg + m_per_block < M ? mbeg + m_per_block : M;
  float acc = 0.f, ws = 0.f;
  for (long long m = mbeg; m < mend; m++) {
    float wm = w ? __ldg(w + m) * wscale : 1.f;
    if (k < K) acc += wm * mat_get(X, ldx, m, k);
    ws += wm;
  }
  if (k < K) atomicAdd(out + k, acc);
  if (osum && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(osum, ws);
}

// Column sums of a BF16 activation image: block = (column block of 64, range of rows); thread t owns the logical
// 16-byte column group t%8 (8 columns) and rows t/8 + 32 i, so every load is one 16-byte chunk and a warp reads
// four full 128-byte image rows per instruction.
__global__ void colsum_img_kernel(const uint8_t* __restrict__ img, int kbs, int K, const float* __restrict__ w,
                                  float wscale, float* __restrict__ out, float* __restrict__ osum, long long M,
                                  int rows_per_block, int f16) {
  const int cb = blockIdx.x;                       // column block
  const long long mbeg = (long long)blockIdx.y * rows_per_block;
  const long long mend = mbeg + rows_per_block < M ? mbeg + rows_per_block : M;
  const int g = threadIdx.x & 7, rsub = threadIdx.x >> 3;     // 256 threads: 32 row lanes x 8 column groups
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float ws = 0.f;
  // four rows per round: the loads of a round are issued before any of them is used (the kernel is pure streaming)
  for (long long m0 = mbeg + rsub; m0 < mend; m0 += 128) {
    uint4 u[4];
    float wm[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const long long m = m0 + 32 * i;
      const bool ok = m < mend;
      const long long mc = ok ? m : m0;
      const uint8_t* p = img + ((size_t)(mc >> 7) * kbs + cb) * 16384 + (((mc & 127) >> 3) * 1024 + (mc & 7) * 128) +
                         (((g ^ (int)(mc & 7)) & 7) << 4);
      u[i] = __ldg(reinterpret_cast<const uint4*>(p));
      wm[i] = ok ? (w ? __ldg(w + mc) * wscale : 1.f) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      float v[8];
      u16x8_to_f32(u[i], f16 != 0, v);
#pragma unroll
      for (int t = 0; t < 8; t++) acc[t] += wm[i] * v[t];
      if (g == 0) ws += wm[i];
    }
  }
  __shared__ float red[64];
  __shared__ float redw;
  if (threadIdx.x < 64) red[threadIdx.x] = 0.f;
  if (threadIdx.x == 0) redw = 0.f;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; i++) atomicAdd(&red[g * 8 + i], acc[i]);
  if (g == 0) atomicAdd(&redw, ws);
  __syncthreads();
  if (threadIdx.x < 64) {
    int k = cb * 64 + threadIdx.x;
    if (k < K) atomicAdd(out + k, red[threadIdx.x]);
  }
  if (osum && cb == 0 && threadIdx.x == 0) atomicAdd(osum, redw);
}

// out0[m] = (dot(H[m,:K], w) + b) * scale     one warp per row (H: FP32 row-major or BF16 activation image)
__global__ void rowdot_kernel(const float* H, int ldh, int K, const float* w, const float* b, float scale,
                              float* out0, long long M) {
  long long m = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / 32;
  int lane = threadIdx.x % 32;
  if (m >= M) return;
  float acc = 0.f;
  if (ldh >= 0) {
    for (int k = lane * 4; k < K; k += 128) {
      float4 h = __ldg(reinterpret_cast<const float4*>(H + m * ldh + k));
      if (k + 0 < K) acc = fmaf(h.x, __ldg(w + k + 0), acc);
      if (k + 1 < K) acc = fmaf(h.y, __ldg(w + k + 1), acc);
      if (k + 2 < K) acc = fmaf(h.z, __ldg(w + k + 2), acc);
      if (k + 3 < K) acc = fmaf(h.w, __ldg(w + k + 3), acc);
    }
  } else {
    for (int k = lane; k < K; k += 32) acc = fmaf(mat_get(H, ldh, m, k), __ldg(w + k), acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out0[m] = (acc + __ldg(b)) * scale;
}

// out[m][j] = in[m][col0 + j], j < ncols
__global__ void extract_cols_kernel(const float* in, int ldi, int col0, int ncols, float* out, int ldo, long long M) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * ncols) return;
  long long m = idx / ncols;
  int j = (int)(idx - m * ncols);
  out[m * ldo + j] = in[m * ldi + col0 + j];
}

// a[m][j] = dy[m][j] * y[m][j] * (1 - y[m][j])   (sigmoid backward) into a padded buffer
__global__ void sigmoid_bwd_kernel(const float* dy, const float* y, int n, float* a, int lda, long long M) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * lda) return;
  long long m = idx / lda;
  int j = (int)(idx - m * lda);
  float v = 0.f;
  if (j < n) { float yy = y[m * n + j]; v = dy[m * n + j] * yy * (1.f - yy); }
  a[idx] = v;
}

// pts[i] = (X[ix], Y[iy], Z[iz]) for the flat ij-meshgrid index i0 + i (renderer.py:16-26)
__global__ void grid_points_kernel(const float* __restrict__ ax, const float* __restrict__ ay,
                                   const float* __restrict__ az, int ny, int nz, long long i0, long long count,
                                   float* __restrict__ pts) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  long long g = i0 + i;
  int iz = (int)(g % nz);
  long long t = g / nz;
  int iy = (int)(t % ny);
  long long ix = t / ny;
  pts[i * 3] = ax[ix]; pts[i * 3 + 1] = ay[iy]; pts[i * 3 + 2] = az[iz];
}

static inline int ew_blocks(long long n) { return cdiv(n, 256); }

void launch_colsum(const float* X, int ldx, int K, const float* w, float wscale, float* out, float* osum, long long M,
                   cudaStream_t st, int f16 = 0) {
  if (M <= 0 || K <= 0) return;
  prof_begin(PC_ELEMENTWISE, 0.0, 0.0, st);
  if (ldx < 0) {
    int rpb = 512;
    dim3 grid(cdiv(K, 64), cdiv(M, rpb));
    colsum_img_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint8_t*>(X), -ldx, K, w, wscale, out, osum, M, rpb, f16);
  } else {
    int mpb = 256;
    dim3 grid(cdiv(K, 256), cdiv(M, mpb));
    colsum_kernel<<<grid, 256, 0, st>>>(X, ldx, K, w, wscale, out, osum, M, mpb);
  }
  prof_end(st);
}

// ------------------------------------------------------------------------------------------------
// SDF network
// ------------------------------------------------------------------------------------------------
struct SdfPlan {
  int L;          // number of hidden layers; linears 0..L
  int e;          // embedded input width
  int in[20], out[20], ldin[20], ldout[20];
  long long woff[20], boff[20];
  long long pack;
  int ldmax;      // FP32 row-major leading dimension of the widest activation
  int wmax;       // widest activation (columns)
  int skip;
  bool img;       // activations are BF16 images (tensor-core mode); ldin/ldout are then -kbs
  bool ok;
};

static SdfPlan sdf_plan(const fneus_sdf_cfg* c) {
  SdfPlan p;
  p.ok = c && c->n_layers >= 1 && c->n_layers < 18 && c->d_in >= 1 && c->d_in <= 4 && c->d_hidden >= 4 &&
         c->d_out >= 1 && c->multires >= 0 && c->multires <= 12;
  if (!p.ok) return p;
  p.L = c->n_layers;
  p.e = pe_dim(c->d_in, c->multires);
  p.skip = c->skip_layer;
  if (p.skip == 0 || p.skip >= p.L) p.ok = false;   // skip in 1..L-1 (or <0 = none)
  if (p.skip > 0 && c->d_hidden - p.e < 1) p.ok = false;
  long long off = 0;
  p.ldmax = 4; p.wmax = 4;
  p.img = precision_mode() == 1;
  for (int l = 0; l <= p.L; l++) {
    p.in[l] = l == 0 ? p.e : c->d_hidden;
    p.out[l] = l == p.L ? c->d_out : ((l + 1 == p.skip) ? c->d_hidden - p.e : c->d_hidden);
    p.ldin[l] = mat_ld(p.in[l], p.img && l > 0);
    p.ldout[l] = mat_ld(p.out[l], p.img && l < p.L);
    p.woff[l] = off; off += (long long)p.out[l] * p.in[l];
    p.boff[l] = off; off += p.out[l];
    if (round_up(p.in[l], 4) > p.ldmax) p.ldmax = round_up(p.in[l], 4);
    if (l < p.L && round_up(p.out[l], 4) > p.ldmax) p.ldmax = round_up(p.out[l], 4);
    if (l > 0 && p.in[l] > p.wmax) p.wmax = p.in[l];
    if (l < p.L && p.out[l] > p.wmax) p.wmax = p.out[l];
  }
  p.pack = off;
  return p;
}

// bytes of BF16 weight images one SDF pass may build (forward + input-gradient form of every layer)
static size_t sdf_img_bytes(const SdfPlan& p) {
  size_t b = 0;
  for (int l = 0; l <= p.L; l++) {
    b += wimg_bytes(p.out[l], l == 0 ? p.in[l] : 0, l == 0 ? 0 : p.in[l]) + 1024;
    b += wimg_bytes(p.in[l], 0, p.out[l]) + 1024;
  }
  return b + 4096;
}

// saved layout: H_1..H_L ([M, ldin[l]]), Q_0..Q_{L-1} ([M, ldout[l]])
static long long sdf_saved_floats(const SdfPlan& p, long long M) {
  long long s = 0;
  for (int l = 1; l <= p.L; l++) s += mat_floats(M, p.in[l], p.img);
  for (int l = 0; l < p.L; l++) s += mat_floats(M, p.out[l], p.img);
  if (p.img) s += mat_floats(M, TC_BK, true) + 256;      // image of PE(x) for the layer-0 weight gradient
  return s + 1024;
}
static long long sdf_buf_floats(const SdfPlan& p, long long M) { return mat_floats(M, p.wmax, p.img) + 256; }
// activation buffers (4 ping-pong ones layer by layer; gbar_0..L, e_0..L-1, abar_0..L-1 for the fused backward) + two
// [M, e] FP32 side buffers; the weight-image arena follows
static int sdf_nbuf(const SdfPlan& p) { return p.img && 3 * p.L + 1 > 4 ? 3 * p.L + 1 : 4; }
static long long sdf_scratch_main(const SdfPlan& p, long long M) {
  return (long long)sdf_nbuf(p) * sdf_buf_floats(p, M) + 2LL * M * round_up(p.e, 4) + 256;
}
static long long sdf_scratch_floats(const SdfPlan& p, long long n) {
  return sdf_scratch_main(p, n) + (long long)(sdf_img_bytes(p) / 4) + 256;
}
// activation images need zero padding rows (they are summed over by the weight-gradient GEMM)
static void sdf_zero_images(const SdfPlan& p, float* ptr, long long floats, long long M, cudaStream_t st) {
  if (p.img && (M & 127)) cudaMemsetAsync(ptr, 0, (size_t)floats * 4, st);
}

struct SdfBufs {
  float* H[20];
  float* Q[20];
};
static SdfBufs sdf_carve(const SdfPlan& p, float* saved, long long M) {
  SdfBufs b;
  float* ptr = saved;
  auto align = [](float* q) { return reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(q) + 1023) & ~(uintptr_t)1023); };
  ptr = align(ptr);
  for (int l = 1; l <= p.L; l++) { b.H[l] = ptr; ptr += mat_floats(M, p.in[l], p.img); }
  for (int l = 0; l < p.L; l++) { b.Q[l] = ptr; ptr += mat_floats(M, p.out[l], p.img); }
  b.H[0] = p.img ? align(ptr) : nullptr;                  // PE(x) image (fused forward only)
  return b;
}

static GenSpec sdf_gen(const fneus_sdf_cfg* c, const float* x, const float* tan) {
  GenSpec g = gen_none();
  gen_add(g, x, c->d_in, c->multires, c->scale, tan);
  g.deriv = tan ? 1 : 0;
  return g;
}

// All weight images one pass needs, declared up front and packed by ONE launch (ImgArena::flush).
struct SdfImgs { const uint8_t* F[20]; const uint8_t* B[20]; };
static SdfImgs sdf_make_images(const fneus_sdf_cfg* c, const SdfPlan& p, const float* w, bool fwd, bool fwd_split_last,
                               bool bwd_hidden, bool bwd_last, ImgArena& ar, cudaStream_t st, int f16 = 0) {
  SdfImgs im;
  for (int l = 0; l <= p.L; l++) { im.F[l] = nullptr; im.B[l] = nullptr; }
  if (precision_mode() != 1) return im;
  for (int l = 0; l <= p.L; l++) {
    const float* W = w + p.woff[l];
    if (fwd && l < p.L) im.F[l] = make_wimg(ar, false, W, p.in[l], 0, p.out[l], 0, l == 0 ? p.in[l] : 0, 0, l == 0 ? 0 : p.in[l], st, f16);
    if (fwd && l == p.L && fwd_split_last) im.F[l] = make_wimg(ar, false, W, p.in[l], 1, p.out[l] - 1, 0, 0, 0, p.in[l], st, f16);
    if (bwd_hidden && l < p.L) im.B[l] = make_wimg(ar, true, W, p.in[l], 0, p.in[l], 0, 0, 0, p.out[l], st, f16);
    if (bwd_last && l == p.L) im.B[l] = make_wimg(ar, true, W, p.in[l], 0, p.in[l], 0, 0, 1, c->d_out - 1, st, f16);
  }
  ar.flush(st);
  return im;
}

static bool sdf_fused_ok(const SdfPlan& p) {
  if (p.L < 2 || p.L > FZ_MAXL || p.e > TC_BK) return false;
  for (int l = 0; l < p.L; l++)
    if (p.in[l] > 256 || p.out[l] > 256) return false;
  if (p.skip > 0 && p.out[p.skip - 1] + p.e > 256) return false;
  return true;
}
static bool sdf_chain_ok(const fneus_sdf_cfg* c, const SdfPlan& p);
static void sdf_chain_value_launch(const fneus_sdf_cfg* c, const SdfPlan& p, const float* w, const float* x, long long M,
                                   float* sdf_out, float* feat_out, float out_sign, const SdfImgs& im, cudaStream_t st);
static int sdf_fused_launch(const fneus_sdf_cfg* c, const SdfPlan& p, const float* w, const float* x, long long M,
                            float* sdf_out, float out_sign, const SdfImgs& im, cudaStream_t st) {
  if (sdf_chain_ok(c, p)) {
    sdf_chain_value_launch(c, p, w, x, M, sdf_out, nullptr, out_sign, im, st);
    return FNEUS_OK;
  }
  static int prepared = 0;
  if (!prepared) {
    cudaError_t e = cudaFuncSetAttribute(sdf_fused_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FZ_SMEM_BYTES);
    if (e != cudaSuccess) return fneus_cuda_error((int)e);
    prepared = 1;
  }
  FusedSdfArgs g;
  g.L = p.L;
  for (int l = 0; l < p.L; l++) {
    g.in[l] = p.in[l]; g.out[l] = p.out[l]; g.img[l] = im.F[l]; g.bias[l] = w + p.boff[l];
  }
  g.w_last = w + p.woff[p.L]; g.b_last = w + p.boff[p.L];
  g.skip = p.skip; g.beta = c->beta; g.scale = c->scale; g.out_sign = out_sign;
  g.gen = sdf_gen(c, x, nullptr);
  g.x = x; g.sdf_out = sdf_out; g.M = M;
  long long npairs = ((M + 127) / 128 + 1) / 2;
  int grid = (int)(npairs < tc_num_sms() ? npairs : tc_num_sms());
  double flops = 0;
  for (int l = 0; l < p.L; l++) flops += 2.0 * M * p.in[l] * p.out[l];
  prof_begin(PC_TC_MLP, flops + 2.0 * M * p.in[p.L], 0.0, st);
  sdf_fused_fwd_kernel<<<grid, FZ_THREADS, FZ_SMEM_BYTES, st>>>(g);
  prof_end(st);
  return FNEUS_OK;
}

// Fused-chain eligibility (sdf_chain.cuh): 256-wide hidden layers, one 64-column PE block, a feature block as wide
// as the last hidden layer (FEATQ rewrites the operand column for column).
static bool sdf_chain_ok(const fneus_sdf_cfg* c, const SdfPlan& p) {
  if (!p.img || precision_mode() != 1 || tc_prepare() != 0 || sdf_chain_prepare() != 0) return false;
  if (tc_debug_flags() & 16) return false;                       // debug: layered execution
  if (!sdf_fused_ok(p) || 2 * p.L + 1 > SC_MAXS || p.L + 1 > sc_bias_slots<FAM_SDF_FWD>()) return false;
  if (c->d_out - 1 != p.in[p.L] || p.in[p.L] > 256 || (c->d_out - 1) % 4 != 0) return false;
  for (int l = 0; l < p.L; l++)                                  // every activation image is 4 blocks wide
    if (cdiv(p.out[l], TC_BK) != 4 || cdiv(p.in[l + 1], TC_BK) != 4) return false;
  return true;
}

// No-grad value chain through the fused kernel: sdf (and optionally the features), nothing saved.
static void sdf_chain_value_launch(const fneus_sdf_cfg* c, const SdfPlan& p, const float* w, const float* x, long long M,
                                   float* sdf_out, float* feat_out, float out_sign, const SdfImgs& im, cudaStream_t st) {
  const int L = p.L;
  const float rsqrt2 = 0.70710678118654752440f;
  SdfChainArgs g;
  memset(&g, 0, sizeof(g));
  double flops = 2.0 * (double)M * p.in[L];
  int ns = 0;
  for (int l = 0; l < L; l++) {
    SdfStep S = sdf_step(SC_SOFTPLUS, im.F[l], l == 0 ? 1 : cdiv(p.in[l], TC_BK), p.out[l], 0);
    S.bias = w + p.boff[l]; S.bias_slot = l;
    S.src = l == 0 ? SRC_PE : SRC_CHAIN;
    if (l + 1 == p.skip) { S.oscale = rsqrt2; S.append = 1; }
    if (l == L - 1) S.dot = 1;
    g.st[ns++] = S;
    flops += 2.0 * (double)M * p.in[l] * p.out[l];
  }
  if (feat_out) {
    SdfStep S = sdf_step(SC_FEATQ, im.F[L], cdiv(p.in[L], TC_BK), c->d_out - 1, 0);
    S.bias = w + p.boff[L] + 1; S.bias_slot = L;
    if (c->feat_image) S.e_out = feat_out;                      // FP16 operand image instead of FP32 rows
    else { S.out = feat_out; S.ldo = c->d_out - 1; }
    g.st[ns++] = S;
    flops += 2.0 * (double)M * p.in[L] * (c->d_out - 1);
  }
  g.nsteps = ns;
  g.gen = sdf_gen(c, x, nullptr); g.gen_t = g.gen;
  g.rvec = w + p.woff[L]; g.b_last = w + p.boff[L];
  g.sdf_out = sdf_out; g.sdf_scale = out_sign / c->scale;
  g.beta = c->beta; g.M = M; g.dbg = 0;
  g.xflags = (tc_debug_flags() >> 8) & 15;
    sdf_chain_launch(g, flops, st);
}

// value chain. If bufs != nullptr activations go to bufs->H (saved) else ping-pong in scratch.
static int sdf_value_chain(const fneus_sdf_cfg* c, const SdfPlan& p, const float* w, const float* x, long long M,
                           float* sdf_out, float* feat_out, SdfBufs* bufs, float* scratch, cudaStream_t st,
                           float out_sign, const SdfImgs& im) {
  const float rsqrt2 = 0.70710678118654752440f;
  float* pp[2] = {scratch, scratch + sdf_buf_floats(p, M)};
  const float* Hin = nullptr;
  for (int l = 0; l <= p.L; l++) {
    ASeg a = l == 0 ? aseg_gen(sdf_gen(c, x, nullptr)) : aseg_mem(Hin, p.ldin[l], p.in[l]);
    const float* W = w + p.woff[l];
    const float* b = w + p.boff[l];
    Epi e = epi_default();
    e.beta = c->beta;
    e.bias = b;
    const bool tc_split_last = (l == p.L) && feat_out && precision_mode() == 1;
    const uint8_t* img = im.F[l];
    if (l < p.L) {
      float* Hout = bufs ? bufs->H[l + 1] : pp[(l + 1) & 1];
      e.mode = EPI_SOFTPLUS;
      e.C = Hout; e.ldc = p.ldin[l + 1];
      e.oscale = (l + 1 == p.skip) ? rsqrt2 : 1.f;
      if (bufs && l == p.L - 1) {
        e.mode = EPI_SOFTPLUS_Q;
        e.Q = bufs->Q[l]; e.ldq = p.ldout[l];
        e.rvec = w + p.woff[p.L];   // row 0 of the last linear
      }
      launch_gemm_fwd(a, W, p.in[l], 0, M, p.out[l], e, st, img);
      if (l + 1 == p.skip) {
        GenSpec g = sdf_gen(c, x, nullptr);
        prof_begin(PC_ELEMENTWISE, 0.0, 0.0, st);
        append_gen_cols_kernel<<<ew_blocks(M * g.ncols), 256, 0, st>>>(g, Hout, p.ldin[l + 1], p.out[l], rsqrt2, M);
        prof_end(st);
      }
      Hin = Hout;
    } else if (tc_split_last) {
      // tensor-core path: sdf row by a row-dot, the feature rows as one <=256-wide GEMM (no 1-column N chunk)
      prof_begin(PC_ELEMENTWISE, 0.0, 0.0, st);
      rowdot_kernel<<<ew_blocks(M * 32), 256, 0, st>>>(Hin, p.ldin[l], p.in[l], W, b, out_sign / c->scale, sdf_out, M);
      prof_end(st);
      e.mode = EPI_LINEAR;
      e.bias = b + 1;
      e.C = feat_out; e.ldc = c->d_out - 1;
      launch_gemm_fwd(a, W, p.in[l], 1, M, p.out[l] - 1, e, st, img);
    } else if (feat_out) {
      e.mode = EPI_SDF_OUT;
      e.out0 = sdf_out; e.out0_scale = out_sign / c->scale;
      e.C = feat_out; e.ldc = c->d_out - 1;
      launch_gemm_fwd(a, W, p.in[l], 0, M, p.out[l], e, st);
    } else {
      prof_begin(PC_ELEMENTWISE, 0.0, 0.0, st);
      rowdot_kernel<<<ew_blocks(M * 32), 256, 0, st>>>(Hin, p.ldin[l], p.in[l], W, b, out_sign / c->scale, sdf_out, M);
      prof_end(st);
    }
  }
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

}  // namespace fneus

using namespace fneus;

extern "C" {

long long fneus_sdf_pack_floats(const fneus_sdf_cfg* cfg) {
  PrecScope prec_scope_(cfg ? cfg->precision : 0);
  SdfPlan p = sdf_plan(cfg);
  return p.ok ? p.pack : -1;
}
long long fneus_sdf_saved_floats(const fneus_sdf_cfg* cfg, long long n) {
  PrecScope prec_scope_(cfg ? cfg->precision : 0);
  SdfPlan p = sdf_plan(cfg);
  return p.ok ? sdf_saved_floats(p, n) : -1;
}
long long fneus_sdf_scratch_floats(const fneus_sdf_cfg* cfg, long long n) {
  PrecScope prec_scope_(cfg ? cfg->precision : 0);
  SdfPlan p = sdf_plan(cfg);
  return p.ok ? sdf_scratch_floats(p, n) : -1;
}

int fneus_sdf_fwd(const fneus_sdf_cfg* cfg, const float* wpack, const float* x, long long n, float* sdf_out,
                  float* feat_out, float* scratch, long long scratch_floats, void* stream) {
  PrecScope prec_scope_(cfg ? cfg->precision : 0);
  SdfPlan p = sdf_plan(cfg);
  if (!p.ok) return FNEUS_ERR_UNSUPPORTED;
  if (n == 0) return FNEUS_OK;
  if (!wpack || !x || !sdf_out || !scratch) return FNEUS_ERR_NULL;
  if (n < 0) return FNEUS_ERR_BAD_SHAPE;
  long long per128 = 2LL * sdf_buf_floats(p, 128);           // scratch floats per 128 points (upper bound)
  long long img_floats = (long long)(sdf_img_bytes(p) / 4) + 256;
  ImgArena ar = arena_make(nullptr, 0);
  long long avail = scratch_floats - 1024;
  if (p.img) {
    if (scratch_floats < img_floats + per128 + 2048) return FNEUS_ERR_WORKSPACE;
    avail = scratch_floats - img_floats - 1024;
    ar.base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(scratch + avail) + 1023) & ~(uintptr_t)1023);
    ar.cap = (size_t)(img_floats - 256) * 4;
  }
  long long chunk = avail / per128 * 128;
  if (chunk < 128) return FNEUS_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  scratch = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(scratch) + 1023) & ~(uintptr_t)1023);
  // the forward chain kernel computes on FP16 operands; the older pair-tile fallback kernel and the layered path on BF16
  const int f16 = sdf_chain_ok(cfg, p) ? 1 : 0;
  SdfImgs im = sdf_make_images(cfg, p, wpack, true, feat_out != nullptr, false, false, ar, st, f16);
  if (p.img && !feat_out && sdf_fused_ok(p) && im.F[0]) {
    int rc = sdf_fused_launch(cfg, p, wpack, x, n, sdf_out, 1.f, im, st);
    if (rc) return rc;
    FNEUS_CHECK_LAUNCH();
    return FNEUS_OK;
  }
  if (f16 && feat_out && (!im.F[0] || !im.F[p.L])) return FNEUS_ERR_WORKSPACE;   // FP16 images only fit the chain kernel
  if (feat_out && im.F[0] && im.F[p.L] && sdf_chain_ok(cfg, p)) {
    sdf_chain_value_launch(cfg, p, wpack, x, n, sdf_out, feat_out, 1.f, im, st);
    FNEUS_CHECK_LAUNCH();
    return FNEUS_OK;
  }
  if (cfg->feat_image && feat_out) return FNEUS_ERR_UNSUPPORTED;   // the image hand-over exists on the chain kernel only
  for (long long m0 = 0; m0 < n; m0 += chunk) {
    long long M = n - m0 < chunk ? n - m0 : chunk;
    sdf_zero_images(p, scratch, 2LL * sdf_buf_floats(p, M), M, st);
    int rc = sdf_value_chain(cfg, p, wpack, x + m0 * cfg->d_in, M, sdf_out + m0,
                             feat_out ? feat_out + m0 * (cfg->d_out - 1) : nullptr, nullptr, scratch, st, 1.f, im);
    if (rc) return rc;
  }
  return FNEUS_OK;
}

int fneus_sdf_grid(const fneus_sdf_cfg* cfg, const float* wpack, const float* ax, const float* ay, const float* az,
                   int nx, int ny, int nz, int ix0, int ix1, float* u_out, float* scratch, long long scratch_floats,
                   void* stream) {
  PrecScope prec_scope_(cfg ? cfg->precision : 0);
  SdfPlan p = sdf_plan(cfg);
  if (!p.ok || cfg->d_in != 3) return FNEUS_ERR_UNSUPPORTED;
  if (!wpack || !ax || !ay || !az || !u_out || !scratch) return FNEUS_ERR_NULL;
  if (nx < 1 || ny < 1 || nz < 1 || ix0 < 0 || ix1 > nx || ix0 > ix1) return FNEUS_ERR_BAD_SHAPE;
  long long per128 = 2LL * sdf_buf_floats(p, 128) + 3 * 128;
  long long img_floats = (long long)(sdf_img_bytes(p) / 4) + 256;
  ImgArena ar = arena_make(nullptr, 0);
  long long avail = scratch_floats - 1024;
  if (p.img) {
    if (scratch_floats < img_floats + per128 + 2048) return FNEUS_ERR_WORKSPACE;
    avail = scratch_floats - img_floats - 1024;
    ar.base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(scratch + avail) + 1023) & ~(uintptr_t)1023);
    ar.cap = (size_t)(img_floats - 256) * 4;
  }
  long long chunk = avail / per128 * 128;
  if (chunk < 128) return FNEUS_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  scratch = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(scratch) + 1023) & ~(uintptr_t)1023);
  long long i0 = (long long)ix0 * ny * nz, i1 = (long long)ix1 * ny * nz;
  SdfImgs im = sdf_make_images(cfg, p, wpack, true, false, false, false, ar, st, sdf_chain_ok(cfg, p) ? 1 : 0);
  const bool fused = p.img && sdf_fused_ok(p) && im.F[0];
  for (long long b = i0; b < i1; b += chunk) {
    long long M = i1 - b < chunk ? i1 - b : chunk;
    float* pts = scratch + 2LL * sdf_buf_floats(p, chunk);
    sdf_zero_images(p, scratch, 2LL * sdf_buf_floats(p, M), M, st);
    prof_begin(PC_ELEMENTWISE, 0.0, 0.0, st);
    grid_points_kernel<<<ew_blocks(M), 256, 0, st>>>(ax, ay, az, ny, nz, b, M, pts);
    prof_end(st);
    int rc = fused ? sdf_fused_launch(cfg, p, wpack, pts, M, u_out + (b - i0), -1.f, im, st)
                   : sdf_value_chain(cfg, p, wpack, pts, M, u_out + (b - i0), nullptr, nullptr, scratch, st, -1.f, im);
    if (rc) return rc;
  }
  return FNEUS_OK;
}

int fneus_sdf_fwd_grad(const fneus_sdf_cfg* cfg, const float* wpack, const float* x, long long M, float* sdf_out,
                       float* feat_out, float* normal_out, float* saved, float* scratch, void* stream) {
  PrecScope prec_scope_(cfg ? cfg->precision : 0);
  SdfPlan p = sdf_plan(cfg);
  if (!p.ok) return FNEUS_ERR_UNSUPPORTED;
  if (M == 0) return FNEUS_OK;
  if (!wpack || !x || !sdf_out || !feat_out || !saved || !scratch) return FNEUS_ERR_NULL;
  if (M < 0) return FNEUS_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  const float rsqrt2 = 0.70710678118654752440f, sqrt2 = 1.41421356237309504880f;
  SdfBufs b = sdf_carve(p, saved, M);
  ImgArena ar = arena_make(nullptr, 0);
  if (p.img) {
    ar.base = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(scratch + sdf_scratch_main(p, M)) + 1023) & ~(uintptr_t)1023);
    ar.cap = sdf_img_bytes(p) - 1024;
  }
  // The fused chain (FP16 forward images) is taken exactly when fneus_sdf_bwd will take its fused chain too -- value +
  // normal graphs on an eligible network; a value-only graph runs layer by layer on BF16 images, like its backward.
  const bool use_chain = sdf_chain_ok(cfg, p) && normal_out != nullptr;
  SdfImgs im = sdf_make_images(cfg, p, wpack, true, true, normal_out != nullptr, false, ar, st, use_chain ? 1 : 0);
  if (use_chain && (im.F[p.L] == nullptr || im.B[0] == nullptr)) return FNEUS_ERR_WORKSPACE;
  if (cfg->feat_image && (!use_chain || cfg->d_out - 1 != 256)) return FNEUS_ERR_UNSUPPORTED;
  if (use_chain) {
    const int L = p.L;
    const float rsqrt2 = 0.70710678118654752440f, sqrt2 = 1.41421356237309504880f;
    SdfChainArgs g;
    memset(&g, 0, sizeof(g));
    double flops = 0.0;
    int ns = 0;
    for (int l = 0; l < L; l++) {
      SdfStep S = sdf_step(SC_SOFTPLUS, im.F[l], l == 0 ? 1 : cdiv(p.in[l], TC_BK), p.out[l], 0);
      S.bias = wpack + p.boff[l]; S.bias_slot = l;
      S.img_out = b.H[l + 1];
      S.src = l == 0 ? SRC_PE : SRC_CHAIN;
      if (l + 1 == p.skip) { S.oscale = rsqrt2; S.append = 1; }
      if (l == L - 1) S.dot = 1;
      g.st[ns++] = S;
      flops += 2.0 * (double)M * p.in[l] * p.out[l];
    }
    {
      SdfStep S = sdf_step(SC_FEATQ, im.F[L], cdiv(p.in[L], TC_BK), cfg->d_out - 1, 0);
      S.bias = wpack + p.boff[L] + 1; S.bias_slot = L;
      if (cfg->feat_image) S.e_out = feat_out;                  // FP16 operand image instead of FP32 rows
      else { S.out = feat_out; S.ldo = cfg->d_out - 1; }
      S.img_out = b.Q[L - 1];
      S.sync_stores = normal_out ? 1 : 0;             // the reverse chain reads the h images back
      g.st[ns++] = S;
      flops += 2.0 * (double)M * p.in[L] * p.out[L];
    }
    if (normal_out) {
      for (int l = L - 1; l >= 1; l--) {
        SdfStep S = sdf_step(SC_SPMUL, im.B[l], cdiv(p.out[l], TC_BK), p.in[l], 1);
        S.h = b.H[l]; S.img_out = b.Q[l - 1];
        S.wait_sync = l == L - 1 ? 1 : 0;
        if (l == p.skip) { S.hscale = sqrt2; S.oscale = rsqrt2; S.csplit = p.out[l - 1]; }
        g.st[ns++] = S;
        flops += 2.0 * (double)M * p.in[l] * p.out[l];
      }
      SdfStep S = sdf_step(SC_G0, im.B[0], cdiv(p.out[0], TC_BK), p.in[0], 1);
      S.out = normal_out;
      S.csplit = p.skip > 0 ? p.out[p.skip - 1] : 0;   // where the parked skip part starts
      g.st[ns++] = S;
      flops += 2.0 * (double)M * p.in[0] * p.out[0];
    }
    g.nsteps = ns;
    g.gen = sdf_gen(cfg, x, nullptr); g.gen_t = g.gen;
    g.pe_img = b.H[0];
    g.rvec = wpack + p.woff[L]; g.b_last = wpack + p.boff[L];
    g.sdf_out = sdf_out; g.sdf_scale = 1.f / cfg->scale;
    g.beta = cfg->beta; g.M = M; g.dbg = (tc_debug_flags() & 64) ? 1 : 0;
    g.xflags = (tc_debug_flags() >> 8) & 15;
    sdf_chain_launch(g, flops, st);
    FNEUS_CHECK_LAUNCH();
    return FNEUS_OK;
  }
  sdf_zero_images(p, saved, sdf_saved_floats(p, M), M, st);
  int rc = sdf_value_chain(cfg, p, wpack, x, M, sdf_out, feat_out, &b, scratch, st, 1.f, im);
  if (rc) return rc;
  if (!normal_out) return FNEUS_OK;   // value-only graph (SDFNetwork.forward under autograd)
  // reverse chain for the normal: g_l = q_l W_l, q_{l-1} = s_{l-1} * g_l
  const int e4 = round_up(p.e, 4);
  float* g0e = scratch + (long long)sdf_nbuf(p) * sdf_buf_floats(p, M);
  float* g0 = g0e + M * e4;
  for (int l = p.L - 1; l >= 1; l--) {
    ASeg a = aseg_mem(b.Q[l], p.ldout[l], p.out[l]);
    Epi e = epi_default();
    e.beta = cfg->beta;
    e.mode = EPI_SPMUL;
    e.H = b.H[l]; e.ldh = p.ldin[l];
    e.C = b.Q[l - 1]; e.ldc = p.ldout[l - 1];
    if (l == p.skip) {
      e.hscale = sqrt2; e.oscale = rsqrt2;
      e.csplit = p.out[l - 1];
      e.C2 = g0e; e.ldc2 = e4;
    }
    launch_gemm_bwd_data(a, wpack + p.woff[l], p.in[l], 0, M, p.in[l], e, st, im.B[l]);
  }
  {
    ASeg a = aseg_mem(b.Q[0], p.ldout[0], p.out[0]);
    Epi e = epi_default();
    e.mode = EPI_LINEAR_ADD;
    e.Q = p.skip > 0 ? g0e : nullptr; e.ldq = e4;
    e.C = g0; e.ldc = e4;
    launch_gemm_bwd_data(a, wpack + p.woff[0], p.in[0], 0, M, p.in[0], e, st, im.B[0]);
  }
  prof_begin(PC_ELEMENTWISE, 0.0, 0.0, st);
  normal_from_g0_kernel<<<ew_blocks(M), 256, 0, st>>>(x, cfg->d_in, cfg->multires, cfg->scale, g0, e4, normal_out, M);
  prof_end(st);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

int fneus_sdf_feat_image_ok(const fneus_sdf_cfg* cfg) {
  if (!cfg) return 0;
  PrecScope prec_scope_(cfg->precision);
  SdfPlan p = sdf_plan(cfg);
  return (p.ok && sdf_chain_ok(cfg, p) && cfg->d_out - 1 == 256) ? 1 : 0;
}

int fneus_sdf_bwd(const fneus_sdf_cfg* cfg, const float* wpack, const float* x, long long M, const float* d_sdf,
                  const float* d_feat, const float* d_normal, float* saved, float* scratch, float* d_wpack,
                  void* stream) {
  PrecScope prec_scope_(cfg ? cfg->precision : 0);
  SdfPlan p = sdf_plan(cfg);
  if (!p.ok) return FNEUS_ERR_UNSUPPORTED;
  if (M == 0) return FNEUS_OK;
  if (!wpack || !x || !saved || !scratch || !d_wpack) return FNEUS_ERR_NULL;
  if (M < 0) return FNEUS_ERR_BAD_SHAPE;
  if (d_feat && ((cfg->d_out - 1) % 4 != 0)) return FNEUS_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int sms = num_sms();
  const float rsqrt2 = 0.70710678118654752440f, sqrt2 = 1.41421356237309504880f;
  SdfBufs b = sdf_carve(p, saved, M);
  scratch = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(scratch) + 1023) & ~(uintptr_t)1023);
  const long long bf = sdf_buf_floats(p, M);
  float* gbuf[2] = {scratch, scratch + bf};
  float* abuf[2] = {scratch + 2 * bf, scratch + 3 * bf};
  sdf_zero_images(p, scratch, 4 * bf, M, st);
  const int L = p.L;
  ImgArena ar = arena_make(nullptr, 0);
  if (p.img) {
    ar.base = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(scratch + sdf_scratch_main(p, M)) + 1023) & ~(uintptr_t)1023);
    ar.cap = sdf_img_bytes(p) - 1024;
  }

  SdfImgs im = sdf_make_images(cfg, p, wpack, d_normal != nullptr, false, true, d_feat != nullptr, ar, st);
  // same predicate as fneus_sdf_fwd_grad: the forward pass of a value + normal graph left FP16 images
  const bool fwd_was_chain = sdf_chain_ok(cfg, p) && d_normal != nullptr;
  if (fwd_was_chain && (!d_feat || b.H[0] == nullptr || im.F[0] == nullptr || im.B[L] == nullptr)) return FNEUS_ERR_UNSUPPORTED;
  if (cfg->feat_image && !fwd_was_chain) return FNEUS_ERR_UNSUPPORTED;
  if (fwd_was_chain) {
    float* G[20]; float* E[20]; float* A[20];
    for (int l = 0; l <= L; l++) G[l] = scratch + (long long)l * bf;
    for (int l = 0; l < L; l++) { E[l] = scratch + (long long)(L + 1 + l) * bf; A[l] = scratch + (long long)(2 * L + 1 + l) * bf; }
    SdfChainArgs g;
    memset(&g, 0, sizeof(g));
    double flops = 0.0;
    int ns = 0;
    for (int l = 0; l < L; l++) {
      SdfStep S = sdf_step(SC_SWEEP, im.F[l], l == 0 ? 1 : cdiv(p.in[l], TC_BK), p.out[l], 0);
      S.src = l == 0 ? SRC_TAN : SRC_CHAIN;
      S.h = b.H[l + 1]; S.q = b.Q[l]; S.e_out = E[l]; S.img_out = G[l + 1];
      if (l + 1 == p.skip) { S.hscale = sqrt2; S.oscale = rsqrt2; S.append = 1; }
      if (l == L - 1) S.sync_stores = 1;               // the value-path backward reads the e images back
      g.st[ns++] = S;
      flops += 2.0 * (double)M * p.in[l] * p.out[l];
    }
    {
      SdfStep S = sdf_step(SC_SDFBWD, im.B[L], cdiv(cfg->d_out - 1, TC_BK), p.in[L], 1);
      S.src = SRC_MEM; S.wait_sync = 1;
      S.h = b.H[L]; S.q = E[L - 1]; S.img_out = A[L - 1];
      S.use_rs = d_sdf != nullptr ? 1 : 0;
      g.st[ns++] = S;
      flops += 2.0 * (double)M * p.in[L] * (cfg->d_out - 1);
    }
    for (int l = L - 1; l >= 1; l--) {
      SdfStep S = sdf_step(SC_SDFBWD, im.B[l], cdiv(p.out[l], TC_BK), p.in[l], 1);
      S.h = b.H[l]; S.q = E[l - 1]; S.img_out = A[l - 1];
      if (l == p.skip) { S.hscale = sqrt2; S.oscale = rsqrt2; S.csplit = p.out[l - 1]; }
      g.st[ns++] = S;
      flops += 2.0 * (double)M * p.in[l] * p.out[l];
    }
    g.nsteps = ns;
    g.gen = sdf_gen(cfg, x, nullptr); g.gen_t = sdf_gen(cfg, x, d_normal);
    g.pe_img = G[0];
    g.mem = d_feat; g.ldm = cfg->feat_image ? -4 : cfg->d_out - 1; g.kmem = cfg->d_out - 1;
    g.rvec = wpack + p.woff[L]; g.b_last = wpack + p.boff[L];
    g.rs = d_sdf; g.rscale = 1.f / cfg->scale;
    g.beta = cfg->beta; g.M = M; g.dbg = (tc_debug_flags() & 64) ? 1 : 0;
    g.xflags = (tc_debug_flags() >> 8) & 15;
    sdf_chain_launch(g, flops, st, FAM_SDF_BWD);
    // weight gradients: dW_l += q_l^T gbar_l + abar_l^T h_l, db_l += colsum abar_l ; last linear: features and row 0
    WgradGroup wg;
    wg.reset(M, sms);
    for (int l = 0; l < L; l++)
      wg.add(b.Q[l], p.ldout[l], aseg_mem(G[l], l == 0 ? -1 : p.ldin[l], p.in[l]), d_wpack + p.woff[l], p.in[l], 0, nullptr,
             p.out[l], st, /*q_l: forward image*/ 1, 0);
    for (int l = L - 1; l >= 0; l--)
      wg.add(A[l], p.ldout[l], aseg_mem(l == 0 ? b.H[0] : b.H[l], l == 0 ? -1 : p.ldin[l], p.in[l]), d_wpack + p.woff[l],
             p.in[l], 0, d_wpack + p.boff[l], p.out[l], st, 0, /*h_l: forward image*/ 1);
    wg.add(d_feat, cfg->feat_image ? -4 : cfg->d_out - 1, aseg_mem(b.H[L], p.ldin[L], p.in[L]), d_wpack + p.woff[L], p.in[L], 1,
           d_wpack + p.boff[L], cfg->d_out - 1, st, 0, 1);
    wg.flush(st);
    launch_colsum(G[L], p.ldin[L], p.in[L], nullptr, 1.f, d_wpack + p.woff[L], nullptr, M, st);
    if (d_sdf)
      launch_colsum(b.H[L], p.ldin[L], p.in[L], d_sdf, 1.f / cfg->scale, d_wpack + p.woff[L], d_wpack + p.boff[L], M, st, 1);
    FNEUS_CHECK_LAUNCH();
    return FNEUS_OK;
  }
  if (d_normal) {
    // double-backward sweep: gbar_0 = T0 nbar ; qbar_l = gbar_l W_l^T ; dW_l += q_l^T gbar_l ;
    // gbar_{l+1} = s_l qbar_l ; e_l = beta (1-s_l) q_l qbar_l (written over q_l)
    for (int l = 0; l < L; l++) {
      ASeg a = l == 0 ? aseg_gen(sdf_gen(cfg, x, d_normal)) : aseg_mem(gbuf[l & 1], p.ldin[l], p.in[l]);
      launch_gemm_wgrad(b.Q[l], p.ldout[l], a, d_wpack + p.woff[l], p.in[l], 0, nullptr, M, p.out[l], sms, st);
      Epi e = epi_default();
      e.beta = cfg->beta;
      e.mode = EPI_SWEEP;
      e.H = b.H[l + 1]; e.ldh = p.ldin[l + 1];
      e.Q = b.Q[l]; e.ldq = p.ldout[l];
      float* gout = gbuf[(l + 1) & 1];
      e.C = gout; e.ldc = p.ldin[l + 1];
      if (l + 1 == p.skip) { e.hscale = sqrt2; e.oscale = rsqrt2; }
      launch_gemm_fwd(a, wpack + p.woff[l], p.in[l], 0, M, p.out[l], e, st, im.F[l]);
      if (l + 1 == p.skip) {
        GenSpec g = sdf_gen(cfg, x, d_normal);
        prof_begin(PC_ELEMENTWISE, 0.0, 0.0, st);
        append_gen_cols_kernel<<<ew_blocks(M * g.ncols), 256, 0, st>>>(g, gout, p.ldin[l + 1], p.out[l], rsqrt2, M);
        prof_end(st);
      }
    }
    // q_L = e_0 (row 0 of the last linear): dW_L[0,:] += sum_m gbar_L
    launch_colsum(gbuf[L & 1], p.ldin[L], p.in[L], nullptr, 1.f, d_wpack + p.woff[L], nullptr, M, st);
  }
  // value-path backward with the augmented abar_l
  float* dWL = d_wpack + p.woff[L];
  float* dbL = d_wpack + p.boff[L];
  if (d_sdf) launch_colsum(b.H[L], p.ldin[L], p.in[L], d_sdf, 1.f / cfg->scale, dWL, dbL, M, st);
  if (d_feat) {
    launch_gemm_wgrad(d_feat, cfg->d_out - 1, aseg_mem(b.H[L], p.ldin[L], p.in[L]), dWL, p.in[L], 1, dbL, M,
                      cfg->d_out - 1, sms, st);
  }
  {
    ASeg a = aseg_mem(d_feat, cfg->d_out - 1, d_feat ? cfg->d_out - 1 : 0, /*wred=*/1);
    if (!d_feat) { a.mem = wpack; a.ldm = 4; }
    Epi e = epi_default();
    e.beta = cfg->beta;
    e.mode = EPI_SDF_BWD;
    e.H = b.H[L]; e.ldh = p.ldin[L];
    if (L == p.skip) { e.hscale = sqrt2; e.oscale = rsqrt2; e.csplit = p.out[L - 1]; }
    e.rs = d_sdf; e.rvec = wpack + p.woff[L]; e.rscale = 1.f / cfg->scale;
    e.Q = d_normal ? b.Q[L - 1] : nullptr; e.ldq = p.ldout[L - 1];
    e.C = abuf[(L - 1) & 1]; e.ldc = p.ldout[L - 1];
    launch_gemm_bwd_data(a, wpack + p.woff[L], p.in[L], 0, M, p.in[L], e, st, im.B[L]);
  }
  for (int l = L - 1; l >= 0; l--) {
    float* al = abuf[l & 1];
    ASeg h = l == 0 ? aseg_gen(sdf_gen(cfg, x, nullptr)) : aseg_mem(b.H[l], p.ldin[l], p.in[l]);
    launch_gemm_wgrad(al, p.ldout[l], h, d_wpack + p.woff[l], p.in[l], 0, d_wpack + p.boff[l], M, p.out[l], sms, st);
    if (l == 0) break;
    ASeg a = aseg_mem(al, p.ldout[l], p.out[l]);
    Epi e = epi_default();
    e.beta = cfg->beta;
    e.mode = EPI_SDF_BWD;
    e.H = b.H[l]; e.ldh = p.ldin[l];
    if (l == p.skip) { e.hscale = sqrt2; e.oscale = rsqrt2; e.csplit = p.out[l - 1]; }
    e.Q = d_normal ? b.Q[l - 1] : nullptr; e.ldq = p.ldout[l - 1];
    e.C = abuf[(l - 1) & 1]; e.ldc = p.ldout[l - 1];
    launch_gemm_bwd_data(a, wpack + p.woff[l], p.in[l], 0, M, p.in[l], e, st, im.B[l]);
  }
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

}  // extern "C"

// Shared device/host helpers for the fneus kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/fneus.h"

namespace fneus {

#define FNEUS_CHECK_LAUNCH()                                   \
  do {                                                         \
    cudaError_t e__ = cudaGetLastError();                      \
    if (e__ != cudaSuccess) return fneus_cuda_error((int)e__); \
  } while (0)

inline int fneus_cuda_error(int e) { return FNEUS_ERR_CUDA_BASE + e; }

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// Generated A-operand columns: concatenation of up to 4 items, each a raw d-vector per row optionally
// followed by its positional encoding [sin(2^k v), cos(2^k v)]_k (reference: models/embedder.py:11-36).
// The encoding is evaluated in registers by whichever kernel consumes it and never stored as a tensor.
// deriv=1 evaluates (dPE/dv)(v) * tan instead (the T0 * t product of SURVEY.md A.1).
// ---------------------------------------------------------------------------------------------
struct GenItem {
  const float* src;  // [M, dim]
  const float* tan;  // [M, dim] tangent (deriv mode) or nullptr
  int dim;
  int multires;      // 0 -> raw vector only
  int col0;          // first generated column of this item
  float scale;       // raw value multiplier (SDFNetwork.scale)
};
struct GenSpec {
  int nitems;
  int ncols;
  int deriv;
  GenItem it[4];
};
__device__ __forceinline__ void sincos_any(float a, bool fast, float* s, float* c) {
  if (fast) { *s = __sinf(a); *c = __cosf(a); } else { sincosf(a, s, c); }
}

__host__ __device__ inline int pe_dim(int d, int multires) { return d * (1 + 2 * multires); }

inline GenSpec gen_none() {
  GenSpec g;
  g.nitems = 0; g.ncols = 0; g.deriv = 0;
  for (int i = 0; i < 4; i++) g.it[i] = GenItem{nullptr, nullptr, 1, 0, 0, 1.f};
  return g;
}
inline void gen_add(GenSpec& g, const float* src, int dim, int multires, float scale = 1.f,
                    const float* tan = nullptr) {
  GenItem& it = g.it[g.nitems++];
  it.src = src; it.tan = tan; it.dim = dim; it.multires = multires; it.col0 = g.ncols; it.scale = scale;
  g.ncols += pe_dim(dim, multires);
}

// decode column j of an item into (component, kind, freq); kind 0 = identity, 1 = sin, 2 = cos
__device__ __forceinline__ void pe_decode(int jj, int d, int& comp, int& kind, float& freq) {
  if (jj < d) { comp = jj; kind = 0; freq = 1.f; return; }
  jj -= d;
  int k = jj / (2 * d);
  int r = jj - k * 2 * d;
  kind = r < d ? 1 : 2;
  comp = r < d ? r : r - d;
  freq = (float)(1u << k);
}

__device__ __forceinline__ float gen_eval(const GenSpec& g, long long m, int j) {
  int sel = 0;
#pragma unroll
  for (int i = 1; i < 4; i++)
    if (i < g.nitems && j >= g.it[i].col0) sel = i;
  const float* src = g.it[0].src; const float* tan = g.it[0].tan;
  int d = g.it[0].dim, c0 = g.it[0].col0; float sc = g.it[0].scale;
#pragma unroll
  for (int i = 1; i < 4; i++)
    if (sel == i) { src = g.it[i].src; tan = g.it[i].tan; d = g.it[i].dim; c0 = g.it[i].col0; sc = g.it[i].scale; }
  int comp, kind; float freq;
  pe_decode(j - c0, d, comp, kind, freq);
  float v = __ldg(src + m * d + comp) * sc;
  if (!g.deriv) {
    if (kind == 0) return v;
    float a = v * freq;
    return kind == 1 ? sinf(a) : cosf(a);
  }
  float t = __ldg(tan + m * d + comp);
  if (kind == 0) return t;
  float a = v * freq;
  return kind == 1 ? freq * cosf(a) * t : -freq * sinf(a) * t;
}

// ---------------------------------------------------------------------------------------------
// Activation images (BF16 tensor-core mode): a matrix X[M, K] stored as BF16 tiles of 128 rows x 64 columns
// (16 KB) in exactly the 128B-swizzled K-major shared-memory layout tcgen05.mma reads, tiles ordered
// [row block][column block].  One tile = one bulk async copy; the same bytes are a valid MN-major operand of
// the weight-gradient GEMM (rows = reduction index).  In the host orchestration an image is passed as a
// (float*) base pointer with a NEGATIVE leading dimension -kbs (kbs = column blocks per row block).
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline size_t img_off(long long m, int k, int kbs) {
  return ((size_t)(m >> 7) * kbs + (k >> 6)) * 16384 + (size_t)(((m & 127) >> 3) * 1024 + (m & 7) * 128 +
         ((((k & 63) >> 3) ^ (int)(m & 7)) << 4) + ((k & 7) << 1));
}
__host__ __device__ inline long long mat_floats(long long M, int width, bool img) {
  if (!img) return M * (long long)round_up(width, 4);
  return ((M + 127) / 128) * (long long)((width + 63) / 64) * 4096;
}
inline int mat_ld(int width, bool img) { return img ? -((width + 63) / 64) : round_up(width, 4); }
__device__ __forceinline__ float mat_get(const float* p, int ld, long long m, int k) {
  if (ld >= 0) return p[m * ld + k];
  unsigned short b = *reinterpret_cast<const unsigned short*>(reinterpret_cast<const uint8_t*>(p) + img_off(m, k, -ld));
  return __uint_as_float((unsigned)b << 16);
}
__device__ __forceinline__ void mat_put(float* p, int ld, long long m, int k, float v) {
  if (ld >= 0) { p[m * ld + k] = v; return; }
  unsigned u = __float_as_uint(v);
  u += 0x7FFFu + ((u >> 16) & 1u);                    // round to nearest even (finite inputs)
  *reinterpret_cast<unsigned short*>(reinterpret_cast<uint8_t*>(p) + img_off(m, k, -ld)) = (unsigned short)(u >> 16);
}

// Row-wise generation of all generated columns of row m: emit(column, value).  One accurate sincosf per
// component and per third octave, the octaves in between by angle doubling (error <= ~4 ulp growth, far below
// BF16 resolution) -- used by the tensor-core producers, where per-column sinf/cosf dominated layer 0.
template <class Emit>
__device__ __forceinline__ void gen_row(const GenSpec& g, long long m, Emit emit) {
#pragma unroll 1
  for (int i = 0; i < g.nitems; i++) {
    const GenItem it = g.it[i];
#pragma unroll 1
    for (int c = 0; c < it.dim; c++) {
      const float v = __ldg(it.src + m * it.dim + c) * it.scale;
      const float t = g.deriv ? __ldg(it.tan + m * it.dim + c) : 0.f;
      emit(it.col0 + c, g.deriv ? t : v);
      float sn = 0.f, cs = 1.f;
#pragma unroll 1
      for (int k = 0; k < it.multires; k++) {
        const float f = (float)(1u << k);
        if (k % 3 == 0) sincosf(v * f, &sn, &cs);
        else { float s2 = 2.f * sn * cs; cs = 1.f - 2.f * sn * sn; sn = s2; }
        emit(it.col0 + it.dim * (1 + 2 * k) + c, g.deriv ? f * cs * t : sn);
        emit(it.col0 + it.dim * (2 + 2 * k) + c, g.deriv ? -f * sn * t : cs);
      }
    }
  }
}
// The same, restricted to the components (counted across items) congruent to `part` modulo `nparts`: the threads
// that share a row split its sincosf work.
template <class Emit>
__device__ __forceinline__ void gen_row_part(const GenSpec& g, long long m, int part, int nparts, Emit emit) {
  int q = 0;
#pragma unroll 1
  for (int i = 0; i < g.nitems; i++) {
    const GenItem it = g.it[i];
#pragma unroll 1
    for (int c = 0; c < it.dim; c++, q++) {
      if (q % nparts != part) continue;
      const float v = __ldg(it.src + m * it.dim + c) * it.scale;
      const float t = g.deriv ? __ldg(it.tan + m * it.dim + c) : 0.f;
      emit(it.col0 + c, g.deriv ? t : v);
      float sn = 0.f, cs = 1.f;
#pragma unroll 1
      for (int k = 0; k < it.multires; k++) {
        const float f = (float)(1u << k);
        if (k % 3 == 0) sincosf(v * f, &sn, &cs);
        else { float s2 = 2.f * sn * cs; cs = 1.f - 2.f * sn * sn; sn = s2; }
        emit(it.col0 + it.dim * (1 + 2 * k) + c, g.deriv ? f * cs * t : sn);
        emit(it.col0 + it.dim * (2 + 2 * k) + c, g.deriv ? -f * sn * t : cs);
      }
    }
  }
}
__device__ __forceinline__ unsigned short f32_to_bf16_bits(float v) {
  unsigned u = __float_as_uint(v);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (unsigned short)(u >> 16);
}

// ---------------------------------------------------------------------------------------------
// 16-bit operand formats of the tensor-core path.  FORWARD-path operands (positional encodings, activations h_l,
// the normal chain's q_l, ReLU activations, their weight images) are FP16: 11 significant bits, i.e. 8x finer
// than BF16, which is what the SDF value needs (inv_s amplifies its error inside the sigmoids; with BF16 operands
// the rendered PSNR sat 0.10-0.18 dB below FP32, with FP16 it is within 0.01 dB) -- their dynamic range is benign
// (activations O(1), weights O(0.1); conversions saturate).  BACKWARD-path operands (upstream gradients: 1e-7 and
// below at 65 536 points per step) stay BF16 for its FP32-sized exponent.  tcgen05.mma kind::f16 takes the two
// formats per operand (instruction descriptor bits 7-9 / 10-12), so weight-gradient GEMMs mix them freely.
// ---------------------------------------------------------------------------------------------
enum Fmt16 { FMT_BF16 = 0, FMT_F16 = 1 };
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));   // first source -> upper half
  return r;
}
__device__ __forceinline__ unsigned short f32_to_f16_bits(float v) { return (unsigned short)(pack_f16x2(v, 0.f) & 0xFFFFu); }
__device__ __forceinline__ float2 unpack_f16x2(uint32_t w) {
  float2 r;
  asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tcvt.f32.f16 %0, lo;\n\tcvt.f32.f16 %1, hi;\n\t}"
      : "=f"(r.x), "=f"(r.y) : "r"(w));
  return r;
}

// Softplus(beta) with torch's threshold 20 (fields.py:72), and its derivative recovered from the
// stored activation h = softplus(a):  sigma'(a) = 1 - exp(-beta h)   (SURVEY.md A.1).
__device__ __forceinline__ float softplus_beta(float a, float beta) {
  float z = a * beta;
  return z > 20.f ? a : log1pf(expf(z)) / beta;
}
__device__ __forceinline__ float softplus_grad_from_pre(float a, float beta) {
  float z = a * beta;
  if (z > 20.f) return 1.f;
  float e = expf(z);
  return e / (e + 1.f);
}
__device__ __forceinline__ float softplus_grad_from_act(float h, float beta) {
  float z = h * beta;
  return z > 20.f ? 1.f : -expm1f(-z);
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
// Bandwidth kernels (sampling, compositing): accurate expf, reciprocal by MUFU.RCP + one multiply (<= 2 ulp) instead of
// the ~12-instruction IEEE division -- the result feeds a 1e-4 (FP32 gate) / 2e-6 (CDF) comparison, not a bit-exact one.
__device__ __forceinline__ float sigmoid_bw(float x) { return __fdividef(1.f, 1.f + expf(-x)); }
// sqrtf(x2) < 1.0f and sqrtf(x2) < 1.2f without the square root: sqrtf is correctly rounded and monotone, so each test is
// a threshold on x2 -- 1.0f itself, and 0x3FB851EC = 1.44000006f, the smallest float whose root reaches 1.2f (both checked
// exhaustively around the thresholds, tests/test_cpu_host.py).
__device__ __forceinline__ bool radius_lt_1(float x2) { return x2 < 1.0f; }
__device__ __forceinline__ bool radius_lt_1p2(float x2) { return x2 < __uint_as_float(0x3FB851ECu); }

// MUFU-based, branch-free variants for the BF16 tensor-core path (results are consumed at BF16 precision; the
// FP32 anchor path keeps the accurate library versions above).  One ex2/lg2/rcp.approx.ftz each, no range fix-ups,
// so that the epilogue's rows interleave freely (the branchy versions serialised on their dependent chains).
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float softplus_beta_fast(float a, float beta, float inv_beta) {
  float z = a * beta;
  float sp = lg2_approx(1.f + ex2_approx(fminf(z, 30.f) * 1.4426950408889634f)) * (0.6931471805599453f * inv_beta);
  return z > 20.f ? a : sp;
}
__device__ __forceinline__ float softplus_grad_from_pre_fast(float a, float beta) {
  float z = a * beta;
  float sg = rcp_approx(1.f + ex2_approx(fmaxf(-z, -30.f) * 1.4426950408889634f));
  return z > 20.f ? 1.f : sg;
}
__device__ __forceinline__ float softplus_grad_from_act_fast(float h, float beta) {
  float z = h * beta;
  float sg = 1.f - ex2_approx(-z * 1.4426950408889634f);
  return z > 20.f ? 1.f : sg;
}
__device__ __forceinline__ float sigmoid_fast(float x) {
  return rcp_approx(1.f + ex2_approx(fminf(-x, 80.f) * 1.4426950408889634f));
}

}  // namespace fneus

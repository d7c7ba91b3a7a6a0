// Flat weight packs in one launch: effective weights of weight-normalised layers (W = g * v / ||v||_row,
// torch.nn.utils.weight_norm as used at fields.py:67-68,143-144) and plain copies of biases / un-normalised weights,
// plus the backward (dg, dv, db from the flat gradient pack), written or accumulated straight into the parameters'
// gradient buffers.  Replaces ~400 tiny torch launches per training step.
#include "fneus_common.cuh"
#include "prof.cuh"

namespace fneus {

constexpr int PACK_MAX_SEGS = 48;
struct PackSeg {
  const float* g;      // [rows] or nullptr (plain copy)
  const float* v;      // [rows, cols] (weight-norm direction, or the tensor to copy)
  float* dg;           // backward destinations (may be nullptr)
  float* dv;
  long long off;       // offset in the flat pack
  int rows, cols;
};
struct PackArgs { int nseg; int accumulate; PackSeg seg[PACK_MAX_SEGS]; };

// one warp per row of a segment; blockIdx.y = segment
__global__ void pack_fwd_kernel(PackArgs a, float* __restrict__ flat) {
  const PackSeg& s = a.seg[blockIdx.y];
  const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
  for (int r = blockIdx.x * warps + (threadIdx.x >> 5); r < s.rows; r += gridDim.x * warps) {
    const float* vr = s.v + (long long)r * s.cols;
    float* out = flat + s.off + (long long)r * s.cols;
    float scale = 1.f;
    if (s.g) {
      float ss = 0.f;
      for (int c = lane; c < s.cols; c += 32) { float x = vr[c]; ss += x * x; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      scale = s.g[r] / sqrtf(ss);
    }
    for (int c = lane; c < s.cols; c += 32) out[c] = s.g ? vr[c] * scale : vr[c];
  }
}

__global__ void pack_bwd_kernel(PackArgs a, const float* __restrict__ dflat) {
  const PackSeg& s = a.seg[blockIdx.y];
  const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
  for (int r = blockIdx.x * warps + (threadIdx.x >> 5); r < s.rows; r += gridDim.x * warps) {
    const float* dw = dflat + s.off + (long long)r * s.cols;
    float* dv = s.dv ? s.dv + (long long)r * s.cols : nullptr;
    if (!s.g) {
      if (dv)
        for (int c = lane; c < s.cols; c += 32) dv[c] = a.accumulate ? dv[c] + dw[c] : dw[c];
      continue;
    }
    const float* vr = s.v + (long long)r * s.cols;
    float ss = 0.f, dot = 0.f;
    for (int c = lane; c < s.cols; c += 32) { float x = vr[c]; ss += x * x; dot += x * dw[c]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ss += __shfl_xor_sync(0xffffffffu, ss, o);
      dot += __shfl_xor_sync(0xffffffffu, dot, o);
    }
    const float norm = sqrtf(ss), gr = s.g[r];
    if (s.dg && lane == 0) {
      float d = dot / norm;
      s.dg[r] = a.accumulate ? s.dg[r] + d : d;
    }
    if (dv) {
      const float k1 = gr / norm, k2 = dot / ss;
      for (int c = lane; c < s.cols; c += 32) {
        float d = k1 * (dw[c] - vr[c] * k2);
        dv[c] = a.accumulate ? dv[c] + d : d;
      }
    }
  }
}

// ---- rows of a 16-bit operand image <-> FP32 rows (include/fneus.h: fneus_image_*) ----
// one thread per 8 consecutive columns (one 16-byte chunk of the swizzled row)
__global__ void image_gather_rows_kernel(const uint8_t* __restrict__ img, int is_fp16, int kbs, int n_cols,
                                         const long long* __restrict__ rows, long long n_sel, float* __restrict__ out) {
  const int cpr = n_cols >> 3;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_sel * cpr) return;
  const long long sel = i / cpr, m = rows[sel];
  const int k = (int)(i % cpr) * 8;
  const uint4 u = *reinterpret_cast<const uint4*>(img + img_off(m, k, kbs));
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
  float v[8];
#pragma unroll
  for (int t = 0; t < 4; t++) {
    if (is_fp16) {
      const float2 f = unpack_f16x2(w[t]);
      v[2 * t] = f.x; v[2 * t + 1] = f.y;
    } else {
      v[2 * t] = __uint_as_float(w[t] << 16); v[2 * t + 1] = __uint_as_float(w[t] & 0xFFFF0000u);
    }
  }
  float* o = out + sel * n_cols + k;
  *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__global__ void image_scatter_add_rows_kernel(uint8_t* __restrict__ img, int kbs, int n_cols,
                                              const long long* __restrict__ rows, long long n_sel,
                                              const float* __restrict__ vals) {
  const int cpr = n_cols >> 3;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_sel * cpr) return;
  const long long sel = i / cpr, m = rows[sel];
  const int k = (int)(i % cpr) * 8;
  uint4* p = reinterpret_cast<uint4*>(img + img_off(m, k, kbs));
  const uint4 u = *p;
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
  const float* a = vals + sel * n_cols + k;
  uint32_t r[4];
#pragma unroll
  for (int t = 0; t < 4; t++) {
    const float lo = __uint_as_float(w[t] << 16) + a[2 * t], hi = __uint_as_float(w[t] & 0xFFFF0000u) + a[2 * t + 1];
    uint32_t ul = __float_as_uint(lo), uh = __float_as_uint(hi);
    ul += 0x7FFFu + ((ul >> 16) & 1u);                  // round to nearest even (finite inputs)
    uh += 0x7FFFu + ((uh >> 16) & 1u);
    r[t] = (ul >> 16) | (uh & 0xFFFF0000u);
  }
  *p = make_uint4(r[0], r[1], r[2], r[3]);
}

}  // namespace fneus

using namespace fneus;

extern "C" {

int fneus_pack_max_segments(void) { return PACK_MAX_SEGS; }

int fneus_pack_fwd(int nseg, const float* const* g, const float* const* v, const long long* off, const int* rows,
                   const int* cols, float* flat, void* stream) {
  if (nseg <= 0) return FNEUS_OK;
  if (nseg > PACK_MAX_SEGS) return FNEUS_ERR_UNSUPPORTED;
  if (!g || !v || !off || !rows || !cols || !flat) return FNEUS_ERR_NULL;
  PackArgs a;
  a.nseg = nseg; a.accumulate = 0;
  int maxrows = 1;
  for (int i = 0; i < nseg; i++) {
    a.seg[i] = PackSeg{g[i], v[i], nullptr, nullptr, off[i], rows[i], cols[i]};
    if (rows[i] > maxrows) maxrows = rows[i];
  }
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(cdiv(maxrows, 8), nseg);
  prof_begin(PC_ELEMENTWISE, 0.0, 0.0, st);
  pack_fwd_kernel<<<grid, 256, 0, st>>>(a, flat);
  prof_end(st);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

int fneus_pack_bwd(int nseg, const float* const* g, const float* const* v, float* const* dg, float* const* dv,
                   const long long* off, const int* rows, const int* cols, const float* dflat, int accumulate,
                   void* stream) {
  if (nseg <= 0) return FNEUS_OK;
  if (nseg > PACK_MAX_SEGS) return FNEUS_ERR_UNSUPPORTED;
  if (!g || !v || !dg || !dv || !off || !rows || !cols || !dflat) return FNEUS_ERR_NULL;
  PackArgs a;
  a.nseg = nseg; a.accumulate = accumulate;
  int maxrows = 1;
  for (int i = 0; i < nseg; i++) {
    a.seg[i] = PackSeg{g[i], v[i], dg[i], dv[i], off[i], rows[i], cols[i]};
    if (rows[i] > maxrows) maxrows = rows[i];
  }
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(cdiv(maxrows, 8), nseg);
  prof_begin(PC_ELEMENTWISE, 0.0, 0.0, st);
  pack_bwd_kernel<<<grid, 256, 0, st>>>(a, dflat);
  prof_end(st);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

long long fneus_image_bytes(long long n_rows, int n_cols) {
  if (n_rows < 0 || n_cols <= 0) return 0;
  return ((n_rows + 127) / 128) * (long long)((n_cols + 63) / 64) * 16384;
}

int fneus_image_gather_rows(const void* image, int is_fp16, int n_cols, const long long* rows, long long n_sel, float* out,
                            void* stream) {
  if (n_sel == 0) return FNEUS_OK;
  if (!image || !rows || !out) return FNEUS_ERR_NULL;
  if (n_sel < 0 || n_cols <= 0 || (n_cols & 63)) return FNEUS_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  prof_begin(PC_ELEMENTWISE, 0.0, 0.0, st);
  image_gather_rows_kernel<<<(int)cdiv(n_sel * (n_cols >> 3), 256), 256, 0, st>>>(
      reinterpret_cast<const uint8_t*>(image), is_fp16, n_cols >> 6, n_cols, rows, n_sel, out);
  prof_end(st);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

int fneus_image_scatter_add_rows(void* image_bf16, int n_cols, const long long* rows, long long n_sel, const float* vals,
                                 void* stream) {
  if (n_sel == 0) return FNEUS_OK;
  if (!image_bf16 || !rows || !vals) return FNEUS_ERR_NULL;
  if (n_sel < 0 || n_cols <= 0 || (n_cols & 63)) return FNEUS_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  prof_begin(PC_ELEMENTWISE, 0.0, 0.0, st);
  image_scatter_add_rows_kernel<<<(int)cdiv(n_sel * (n_cols >> 3), 256), 256, 0, st>>>(
      reinterpret_cast<uint8_t*>(image_bf16), n_cols >> 6, n_cols, rows, n_sel, vals);
  prof_end(st);
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

}  // extern "C"

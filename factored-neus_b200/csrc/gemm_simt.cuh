// FP32 SIMT GEMM engine: the exactness anchor (<=1e-4 vs the reference's FP32 cuBLAS path) for every
// dense contraction on the hot path (fields.py:86,168,242 forward; their autograd backward).
//
//   gemm_mk   : C[M, N]  = epilogue( A[M, K] * B + bias )      B[k][n] = W[n][k]  (forward, W^T)
//                                                          or  B[k][n] = W[k][n]  (backward-data)
//   gemm_wgrad: dW[n][k] += sum_m dY[m][n] * A[m][k]            (backward-weight, split over m, RED.ADD)
//
// The A operand is [generated columns | memory columns]; generated columns (positional encodings of raw
// per-row vectors) are evaluated in registers by the tile loader.  Tiles 128x128x16, 256 threads, 8x8
// register micro-tile, register-prefetch double buffering.
#pragma once
#include "fneus_common.cuh"
#include "prof.cuh"

namespace fneus {

struct ASeg {
  GenSpec gen;       // gen.ncols == 0 -> none
  int wred_gen;      // offset of the generated block along W's reduction axis
  const float* mem;  // [M, ldm] or nullptr
  int ldm;           // multiple of 4, base 16B aligned
  int kmem;
  int wred_mem;      // offset of the memory block along W's reduction axis
};

enum EpiMode {
  EPI_LINEAR = 0,
  EPI_RELU,
  EPI_SIGMOID,
  EPI_SOFTPLUS,     // C = softplus(acc+b) * oscale
  EPI_SOFTPLUS_Q,   // + Q[m][n] = softplus'(acc+b) * rvec[n]
  EPI_SDF_OUT,      // col 0 -> out0[m] = v*out0_scale ; col n>=1 -> C[m][n-1]
  EPI_SPMUL,        // primary: C = s(H) * acc * oscale ; secondary (n>=csplit): C2 = acc * oscale
  EPI_SWEEP,        // s = s(H); C = s*acc*oscale ; Q = beta*(1-s)*Q*acc
  EPI_SDF_BWD,      // primary: C = s(H) * (acc + rs[m]*rvec[n]*rscale) * oscale + Q[m][n]
  EPI_RELUMASK,     // primary: C = H>0 ? acc (+ rs[m]*rvec[n]*rscale) : 0 ; secondary: C2 (+)= (H2 ? (H2>0 ? acc : 0) : acc)
  EPI_LINEAR_ADD,   // C = acc + Q[m][n]   (Q read-only addend, may be null)
};

struct Epi {
  int mode;
  const float* bias;
  float oscale;
  float* C; int ldc;
  float* C2; int ldc2; int csplit; int accumulate2;
  const float* H; int ldh; float hscale;
  const float* H2; int ldh2;
  float* Q; int ldq;
  const float* rvec;
  const float* rs; float rscale;
  float beta;
  float* out0; float out0_scale;
  int dbg;
};

inline Epi epi_default() {
  Epi e;
  e.mode = EPI_LINEAR; e.bias = nullptr; e.oscale = 1.f; e.C = nullptr; e.ldc = 0; e.C2 = nullptr; e.ldc2 = 0;
  e.csplit = 1 << 30; e.accumulate2 = 0; e.H = nullptr; e.ldh = 0; e.hscale = 1.f; e.H2 = nullptr; e.ldh2 = 0;
  e.Q = nullptr; e.ldq = 0; e.rvec = nullptr; e.rs = nullptr; e.rscale = 1.f; e.beta = 100.f; e.out0 = nullptr;
  e.out0_scale = 1.f; e.dbg = 0;
  return e;
}

constexpr int GB_M = 128, GB_N = 128, GB_K = 16, GB_PAD = 4, GB_THREADS = 256;

// Addresses of the auxiliary operands the epilogue of element (m, n) reads (nullptr = none); split from the
// arithmetic so that tile epilogues can issue all their loads before consuming them.
__device__ __forceinline__ void epilogue_aux(const Epi& e, long long m, int n, const float*& hp, const float*& qp) {
  hp = nullptr; qp = nullptr;
  switch (e.mode) {
    case EPI_SPMUL:
      if (n < e.csplit) hp = e.H + m * e.ldh + n;
      break;
    case EPI_SWEEP:
      hp = e.H + m * e.ldh + n; qp = e.Q + m * e.ldq + n;
      break;
    case EPI_SDF_BWD:
      if (n < e.csplit) { hp = e.H + m * e.ldh + n; if (e.Q) qp = e.Q + m * e.ldq + n; }
      break;
    case EPI_RELUMASK:
      if (n < e.csplit) { if (e.H && e.C) hp = e.H + m * e.ldh + n; }
      else if (e.C2) {
        if (e.H2) hp = e.H2 + m * e.ldh2 + (n - e.csplit);
        if (e.accumulate2) qp = e.C2 + m * e.ldc2 + (n - e.csplit);
      }
      break;
    case EPI_LINEAR_ADD:
      if (e.Q) qp = e.Q + m * e.ldq + n;
      break;
    default: break;
  }
}

// h, q: values at the addresses reported by epilogue_aux (ignored when that address was nullptr)
template <bool FAST = false>
__device__ __forceinline__ void epilogue_apply(const Epi& e, long long m, int n, float acc, float h, float q) {
  auto sp = [&](float v) { return FAST ? softplus_beta_fast(v, e.beta, 1.f / e.beta) : softplus_beta(v, e.beta); };
  auto sp_pre = [&](float v) { return FAST ? softplus_grad_from_pre_fast(v, e.beta) : softplus_grad_from_pre(v, e.beta); };
  auto sp_act = [&](float v) { return FAST ? softplus_grad_from_act_fast(v, e.beta) : softplus_grad_from_act(v, e.beta); };
  auto sg = [&](float v) { return FAST ? sigmoid_fast(v) : sigmoidf_(v); };
  switch (e.mode) {
    case EPI_LINEAR: {
      float v = acc + (e.bias ? __ldg(e.bias + n) : 0.f);
      e.C[m * e.ldc + n] = v * e.oscale;
    } break;
    case EPI_RELU: {
      float v = acc + (e.bias ? __ldg(e.bias + n) : 0.f);
      e.C[m * e.ldc + n] = fmaxf(v, 0.f);
    } break;
    case EPI_SIGMOID: {
      float v = acc + (e.bias ? __ldg(e.bias + n) : 0.f);
      e.C[m * e.ldc + n] = sg(v);
    } break;
    case EPI_SOFTPLUS: {
      float v = acc + __ldg(e.bias + n);
      e.C[m * e.ldc + n] = sp(v) * e.oscale;
    } break;
    case EPI_SOFTPLUS_Q: {
      float v = acc + __ldg(e.bias + n);
      e.C[m * e.ldc + n] = sp(v) * e.oscale;
      e.Q[m * e.ldq + n] = sp_pre(v) * __ldg(e.rvec + n);
    } break;
    case EPI_SDF_OUT: {
      float v = acc + __ldg(e.bias + n);
      if (n == 0) e.out0[m] = v * e.out0_scale;
      else if (e.C) e.C[m * e.ldc + (n - 1)] = v;
    } break;
    case EPI_SPMUL: {
      if (n < e.csplit) {
        float s = sp_act(h * e.hscale);
        e.C[m * e.ldc + n] = s * acc * e.oscale;
      } else if (e.C2) {
        e.C2[m * e.ldc2 + (n - e.csplit)] = acc * e.oscale;
      }
    } break;
    case EPI_SWEEP: {
      float s = sp_act(h * e.hscale);
      e.C[m * e.ldc + n] = s * acc * e.oscale;
      e.Q[m * e.ldq + n] = e.beta * (1.f - s) * q * acc;
    } break;
    case EPI_SDF_BWD: {
      if (n < e.csplit) {
        float s = sp_act(h * e.hscale);
        float v = acc;
        if (e.rs) v += __ldg(e.rs + m) * __ldg(e.rvec + n) * e.rscale;
        float add = e.Q ? q : 0.f;
        e.C[m * e.ldc + n] = s * v * e.oscale + add;
      }
    } break;
    case EPI_RELUMASK: {
      if (n < e.csplit) {
        if (e.C) {
          float v = acc;
          if (e.rs) v += __ldg(e.rs + m) * __ldg(e.rvec + n) * e.rscale;
          e.C[m * e.ldc + n] = e.H ? (h > 0.f ? v : 0.f) : v;
        }
      } else if (e.C2) {
        int c = n - e.csplit;
        float v = e.H2 ? (h > 0.f ? acc : 0.f) : acc;
        if (e.accumulate2) v += q;
        e.C2[m * e.ldc2 + c] = v;
      }
    } break;
    case EPI_LINEAR_ADD: {
      float v = acc + (e.Q ? q : 0.f);
      e.C[m * e.ldc + n] = v * e.oscale;
    } break;
  }
}

__device__ __forceinline__ void epilogue_store(const Epi& e, long long m, int n, float acc) {
  const float *hp, *qp;
  epilogue_aux(e, m, n, hp, qp);
  float h = hp ? __ldg(hp) : 0.f;
  float q = qp ? *qp : 0.f;
  epilogue_apply<false>(e, m, n, acc, h, q);
}

// ------------------------------------------------------------------------------------------------
// A-tile loader: rows m0..m0+127, 16 reduction columns starting at k0 of the current phase.
// thread t: row = t/4 (+64), kq = (t%4)*4 -> one float4 per half.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 load_a4(const ASeg& a, int phase, long long m, int M, int k) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (m >= M) return v;
  if (phase == 0) {
    int nc = a.gen.ncols;
    if (k + 0 < nc) v.x = gen_eval(a.gen, m, k + 0);
    if (k + 1 < nc) v.y = gen_eval(a.gen, m, k + 1);
    if (k + 2 < nc) v.z = gen_eval(a.gen, m, k + 2);
    if (k + 3 < nc) v.w = gen_eval(a.gen, m, k + 3);
  } else if (k < a.kmem) {
    v = __ldg(reinterpret_cast<const float4*>(a.mem + m * a.ldm + k));
    if (k + 1 >= a.kmem) v.y = 0.f;
    if (k + 2 >= a.kmem) v.z = 0.f;
    if (k + 3 >= a.kmem) v.w = 0.f;
  }
  return v;
}

template <bool WT>
__global__ void __launch_bounds__(GB_THREADS, 2)
gemm_mk_kernel(ASeg a, const float* __restrict__ W, int ldw, int wout0, int M, int N, Epi e) {
  __shared__ float As[2][GB_K][GB_M + GB_PAD];
  __shared__ float Bs[2][GB_K][GB_N + GB_PAD];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const long long m0 = (long long)blockIdx.x * GB_M;
  const int n0 = blockIdx.y * GB_N;

  const int tiles_gen = (a.gen.ncols + GB_K - 1) / GB_K;
  const int tiles_mem = (a.kmem + GB_K - 1) / GB_K;
  const int T = tiles_gen + tiles_mem;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j] = 0.f;

  float4 ra[2];
  float rb[8];
  const int a_row = tid / 4, a_kq = (tid % 4) * 4;
  // W tile loader mapping. WT: n = tid/2, 8 consecutive k at (tid%2)*8.  !WT: kk = tid/16, 8 n at (tid%16)*8.
  auto load_tile = [&](int t) {
    int phase = t < tiles_gen ? 0 : 1;
    int k0 = (phase == 0 ? t : t - tiles_gen) * GB_K;
    int kmax = phase == 0 ? a.gen.ncols : a.kmem;
    int wred0 = phase == 0 ? a.wred_gen : a.wred_mem;
    ra[0] = load_a4(a, phase, m0 + a_row, M, k0 + a_kq);
    ra[1] = load_a4(a, phase, m0 + a_row + 64, M, k0 + a_kq);
    if (WT) {
      int n = n0 + tid / 2;
      int kb = k0 + (tid % 2) * 8;
      const float* wp = W + (long long)(wout0 + n) * ldw + wred0 + kb;
#pragma unroll
      for (int i = 0; i < 8; i++) rb[i] = (n < N && kb + i < kmax) ? __ldg(wp + i) : 0.f;
    } else {
      int kk = k0 + tid / 16;
      int nb = n0 + (tid % 16) * 8;
      const float* wp = W + (long long)(wred0 + kk) * ldw + wout0 + nb;
#pragma unroll
      for (int i = 0; i < 8; i++) rb[i] = (kk < kmax && nb + i < N) ? __ldg(wp + i) : 0.f;
    }
  };
  auto store_tile = [&](int buf) {
    As[buf][a_kq + 0][a_row] = ra[0].x; As[buf][a_kq + 1][a_row] = ra[0].y;
    As[buf][a_kq + 2][a_row] = ra[0].z; As[buf][a_kq + 3][a_row] = ra[0].w;
    As[buf][a_kq + 0][a_row + 64] = ra[1].x; As[buf][a_kq + 1][a_row + 64] = ra[1].y;
    As[buf][a_kq + 2][a_row + 64] = ra[1].z; As[buf][a_kq + 3][a_row + 64] = ra[1].w;
    if (WT) {
      int n = tid / 2, kb = (tid % 2) * 8;
#pragma unroll
      for (int i = 0; i < 8; i++) Bs[buf][kb + i][n] = rb[i];
    } else {
      int kk = tid / 16, nb = (tid % 16) * 8;
#pragma unroll
      for (int i = 0; i < 8; i++) Bs[buf][kk][nb + i] = rb[i];
    }
  };

  if (T > 0) {
    load_tile(0);
    store_tile(0);
  }
  __syncthreads();
  for (int t = 0; t < T; t++) {
    int cur = t & 1;
    if (t + 1 < T) load_tile(t + 1);
#pragma unroll
    for (int kk = 0; kk < GB_K; kk++) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[cur][kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[cur][kk][ty * 4 + 64]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * 4 + 64]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (t + 1 < T) store_tile(cur ^ 1);
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; i++) {
    long long m = m0 + ty * 4 + (i % 4) + (i / 4) * 64;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      int n = n0 + tx * 4 + (j % 4) + (j / 4) * 64;
      if (n < N) epilogue_store(e, m, n, acc[i][j]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// dW[(wout0+n)*ldw + wred + k] += sum_m dY[m][n] * A[m][k] ; db[wout0+n] += sum_m dY[m][n]
// grid.x = n_tiles * k_tiles (k tiles enumerate the generated block then the memory block),
// grid.y = splits over m.  dY: [M, ldy] (ldy % 4 == 0).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GB_THREADS, 2)
gemm_wgrad_kernel(const float* __restrict__ dY, int ldy, ASeg a, float* __restrict__ dW, int ldw, int wout0,
                  float* __restrict__ db, int M, int N, int m_per_split) {
  __shared__ float As[2][GB_K][GB_M + GB_PAD];   // dY tile   [mm][n]
  __shared__ float Bs[2][GB_K][GB_N + GB_PAD];   // A tile    [mm][k]
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int ktiles_gen = (a.gen.ncols + GB_N - 1) / GB_N;
  const int ktiles_mem = (a.kmem + GB_N - 1) / GB_N;
  const int ktiles = ktiles_gen + ktiles_mem;
  const int nt = blockIdx.x / ktiles, kt = blockIdx.x % ktiles;
  const int n0 = nt * GB_M;
  const int phase = kt < ktiles_gen ? 0 : 1;
  const int k0 = (phase == 0 ? kt : kt - ktiles_gen) * GB_N;
  const int kmax = phase == 0 ? a.gen.ncols : a.kmem;
  const int wred0 = phase == 0 ? a.wred_gen : a.wred_mem;
  const long long mbeg = (long long)blockIdx.y * m_per_split;
  long long mend = mbeg + m_per_split;
  if (mend > M) mend = M;
  const int T = mbeg < mend ? (int)((mend - mbeg + GB_K - 1) / GB_K) : 0;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j] = 0.f;
  float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool do_bias = (db != nullptr) && kt == 0;

  const int l_mm = tid / 32, l_c = (tid % 32) * 4;
  float4 ry[2], rx[2];
  auto load_tile = [&](int t) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
      long long m = mbeg + (long long)t * GB_K + l_mm + h * 8;
      float4 y = make_float4(0.f, 0.f, 0.f, 0.f), x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < mend) {
        int n = n0 + l_c;
        if (n < N) {
          y = __ldg(reinterpret_cast<const float4*>(dY + m * ldy + n));
          if (n + 1 >= N) y.y = 0.f;
          if (n + 2 >= N) y.z = 0.f;
          if (n + 3 >= N) y.w = 0.f;
        }
        x = load_a4(a, phase, m, M, k0 + l_c);
      }
      ry[h] = y; rx[h] = x;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
      *reinterpret_cast<float4*>(&As[buf][l_mm + h * 8][l_c]) = ry[h];
      *reinterpret_cast<float4*>(&Bs[buf][l_mm + h * 8][l_c]) = rx[h];
      if (do_bias) { bsum.x += ry[h].x; bsum.y += ry[h].y; bsum.z += ry[h].z; bsum.w += ry[h].w; }
    }
  };
  if (T > 0) { load_tile(0); store_tile(0); }
  __syncthreads();
  for (int t = 0; t < T; t++) {
    int cur = t & 1;
    if (t + 1 < T) load_tile(t + 1);
#pragma unroll
    for (int kk = 0; kk < GB_K; kk++) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[cur][kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[cur][kk][ty * 4 + 64]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * 4 + 64]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (t + 1 < T) store_tile(cur ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    int n = n0 + ty * 4 + (i % 4) + (i / 4) * 64;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      int k = k0 + tx * 4 + (j % 4) + (j / 4) * 64;
      if (k < kmax) atomicAdd(dW + (long long)(wout0 + n) * ldw + wred0 + k, acc[i][j]);
    }
  }
  if (do_bias && T > 0) {
    int n = n0 + l_c;
    if (n + 0 < N) atomicAdd(db + wout0 + n + 0, bsum.x);
    if (n + 1 < N) atomicAdd(db + wout0 + n + 1, bsum.y);
    if (n + 2 < N) atomicAdd(db + wout0 + n + 2, bsum.z);
    if (n + 3 < N) atomicAdd(db + wout0 + n + 3, bsum.w);
  }
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
inline ASeg aseg_mem(const float* mem, int ldm, int kmem, int wred = 0) {
  ASeg a; a.gen = gen_none(); a.wred_gen = 0; a.mem = mem; a.ldm = ldm; a.kmem = kmem; a.wred_mem = wred;
  return a;
}
inline ASeg aseg_gen(const GenSpec& g, int wred = 0) {
  ASeg a; a.gen = g; a.wred_gen = wred; a.mem = nullptr; a.ldm = 4; a.kmem = 0; a.wred_mem = 0;
  return a;
}
inline ASeg aseg_gen_mem(const GenSpec& g, int wred_gen, const float* mem, int ldm, int kmem, int wred_mem) {
  ASeg a; a.gen = g; a.wred_gen = wred_gen; a.mem = mem; a.ldm = ldm; a.kmem = kmem; a.wred_mem = wred_mem;
  return a;
}

// forward: C[M,N] = epi(A * W[wout0:wout0+N, :]^T)
inline void launch_simt_fwd(const ASeg& a, const float* W, int ldw, int wout0, long long M, int N, const Epi& e,
                            cudaStream_t st) {
  if (M <= 0 || N <= 0) return;
  dim3 grid(cdiv(M, GB_M), cdiv(N, GB_N));
  prof_begin(PC_GEMM_FWD, 2.0 * (double)M * N * (a.gen.ncols + a.kmem), 0.0, st);
  gemm_mk_kernel<true><<<grid, GB_THREADS, 0, st>>>(a, W, ldw, wout0, (int)M, N, e);
  prof_end(st);
}
// backward-data: C[M,N] = epi(A[M, Kred] * W[wred.., wout0:wout0+N])
inline void launch_simt_bwd_data(const ASeg& a, const float* W, int ldw, int wout0, long long M, int N,
                                 const Epi& e, cudaStream_t st) {
  if (M <= 0 || N <= 0) return;
  dim3 grid(cdiv(M, GB_M), cdiv(N, GB_N));
  prof_begin(PC_GEMM_BWD_DATA, 2.0 * (double)M * N * (a.gen.ncols + a.kmem), 0.0, st);
  gemm_mk_kernel<false><<<grid, GB_THREADS, 0, st>>>(a, W, ldw, wout0, (int)M, N, e);
  prof_end(st);
}
inline void launch_simt_wgrad(const float* dY, int ldy, const ASeg& a, float* dW, int ldw, int wout0, float* db,
                              long long M, int N, int num_sms, cudaStream_t st) {
  if (M <= 0 || N <= 0) return;
  int ktiles = cdiv(a.gen.ncols, GB_N) + cdiv(a.kmem, GB_N);
  if (ktiles == 0) return;
  int tiles = cdiv(N, GB_M) * ktiles;
  int splits = (4 * num_sms + tiles - 1) / tiles;
  int max_splits = cdiv(M, 4 * GB_K);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int mps = round_up(cdiv(M, splits), GB_K);
  splits = cdiv(M, mps);
  dim3 grid(tiles, splits);
  prof_begin(PC_GEMM_WGRAD, 2.0 * (double)M * N * (a.gen.ncols + a.kmem), 0.0, st);
  gemm_wgrad_kernel<<<grid, GB_THREADS, 0, st>>>(dY, ldy, a, dW, ldw, wout0, db, (int)M, N, mps);
  prof_end(st);
}

}  // namespace fneus

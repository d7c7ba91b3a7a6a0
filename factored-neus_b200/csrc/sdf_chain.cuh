// Fused SDF chains (BF16 tensor-core mode): the three layer sequences of the SDF network that a training step runs
// on the render_core points, each as ONE persistent kernel over pairs of 128-point tiles, operand on chip between
// layers (shared memory BF16 <-> TMEM FP32), weights streamed as pre-packed images:
//
//   forward  : value chain h_{l+1} = softplus(W_l h_l + b_l) (fields.py:74-95), sdf = h_L . W_L[0] and the feature
//              GEMM, then the reverse chain of the analytic gradient q_{l-1} = s_l * (q_l W_l) down to g_0
//              (fields.py:101-111 as reverse-mode) -- 2L+1 GEMM steps
//   backward : the double-backward sweep gbar_{l+1} = s_l * (gbar_l W_l^T), e_l = beta (1 - s_l) q_l (gbar_l W_l^T),
//              then the value-path backward abar_{l-1} = s_l * (abar_l W_l) + e_{l-1} -- 2L GEMM steps
//
// Every activation a later pass or the weight-gradient GEMMs need is written once as a BF16 image, by the thread
// that owns the row; a thread only ever reads back image bytes it wrote itself (same row, same column chunks), so
// no cross-thread visibility is involved.  The weight gradients run afterwards as grouped tensor-core launches.
//
//   warp 0      : MMA issuer (+ TMEM alloc: one 256-column accumulator per tile of the pair)
//   warp 1      : weight-image loader, ring of 4 half-tiles (128 output columns x 64 reduction, 16 KB)
//   warps 2-17  : 8 per tile, thread = one row (TMEM lane), 16 columns at a time, the next chunk's auxiliary
//                 image rows prefetched into registers while the current chunk is computed
#pragma once
#include "chain_fused.cuh"

namespace fneus {

constexpr int SC_THREADS = 576, SC_WSTAGES = 4, SC_MAXS = 20, SC_SLOT_BLOCKS = 4, SC_BIAS_SLOTS = 12;
enum SdfStepMode { SC_SOFTPLUS = 0, SC_FEATQ, SC_SPMUL, SC_G0, SC_SWEEP, SC_SDFBWD };
enum SdfStepSrc { SRC_CHAIN = 0, SRC_PE, SRC_TAN, SRC_MEM };

struct SdfStep {
  const uint8_t* wimg;  // weight image ([1 n-chunk][KB] tiles of 32 KB)
  const float* bias;    // SOFTPLUS / FEATQ
  float* img_out;       // image of the step's result (4 blocks per tile) or null
  const float* h;       // SPMUL / SWEEP / SDFBWD: image of the forward activation the softplus derivative comes from
  const float* q;       // SWEEP / FEATQ: image q_l ; SDFBWD: image e_{l-1} ; G0: FP32 [M, ldo] addend (or null)
  float* e_out;         // SWEEP: image e_l ; SOFTPLUS with dot: image q_{L-1}
  float* out;           // FEATQ: features FP32 [M, ldo] ; G0: g_0 FP32 [M, ldo] ; SPMUL at the skip: PE part FP32 [M, ldo]
  int KB, N, mode, bmn, src;
  int ldo, csplit, append, dot, use_rs, bias_slot;
  float hscale, oscale;
};
struct SdfChainArgs {
  int nsteps;
  SdfStep st[SC_MAXS];
  GenSpec gen;          // PE(x * scale)
  GenSpec gen_t;        // its tangent form (deriv = 1) for the double-backward sweep
  const float* mem;     // SRC_MEM: FP32 [M, ldm], kmem columns
  int ldm, kmem;
  float* pe_img;        // optional 1-block image copy of the SRC_PE / SRC_TAN operand
  const float* rvec;    // row 0 of the last linear [<= 256]
  const float* b_last;  // its bias
  float* sdf_out;       // [M]
  float sdf_scale;      // out_sign / scale
  const float* rs;      // SDFBWD with use_rs: d_sdf [M]
  float rscale, beta;
  long long M;
};
struct SCSmem {
  uint64_t wfull[SC_WSTAGES], wempty[SC_WSTAGES];
  uint64_t a_ready[2], acc_full[2];
  uint32_t tmem_base;
};
constexpr int SC_SLOT_BYTES = SC_SLOT_BLOCKS * TC_A_BYTES;
constexpr int SC_SMEM_BYTES = 2 * SC_SLOT_BYTES + SC_WSTAGES * CH_WBYTES + (SC_BIAS_SLOTS * 256 + 256 + 256) * 4 + 1024 + 256;

__device__ __forceinline__ void bf16x8_to_f32(const uint4 u, float* out) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int t = 0; t < 4; t++) {
    out[2 * t] = __uint_as_float(w[t] << 16);
    out[2 * t + 1] = __uint_as_float(w[t] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ uint4 f32x8_to_bf16(const float* y) {
  const uint2 lo = pack_bf16x4(make_float4(y[0], y[1], y[2], y[3]));
  const uint2 hi = pack_bf16x4(make_float4(y[4], y[5], y[6], y[7]));
  return make_uint4(lo.x, lo.y, hi.x, hi.y);
}

__global__ void __launch_bounds__(SC_THREADS, 1) sdf_chain_kernel(const __grid_constant__ SdfChainArgs g) {
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sAt[2] = {base, base + SC_SLOT_BYTES};
  uint8_t* sW0 = base + 2 * SC_SLOT_BYTES;
  float* sbias = reinterpret_cast<float*>(sW0 + SC_WSTAGES * CH_WBYTES);             // [SC_BIAS_SLOTS][256]
  float* srvec = sbias + SC_BIAS_SLOTS * 256;                                        // [256]
  float* sdot = srvec + 256;                                                         // [2][128]
  SCSmem* ctl = reinterpret_cast<SCSmem*>(sdot + 256);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long ntiles = (g.M + 127) / 128;
  const long long npairs = (ntiles + 1) / 2;
  const float rsqrt2 = 0.70710678118654752440f;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < SC_WSTAGES; s++) { mbar_init(&ctl->wfull[s], 1); mbar_init(&ctl->wempty[s], 1); }
#pragma unroll
    for (int t = 0; t < 2; t++) { mbar_init(&ctl->a_ready[t], 256); mbar_init(&ctl->acc_full[t], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int s = 0; s < g.nsteps; s++) {
    const SdfStep& S = g.st[s];
    if (S.bias != nullptr && S.bias_slot >= 0)
      for (int c = tid; c < 256; c += SC_THREADS) sbias[S.bias_slot * 256 + c] = c < S.N ? __ldg(S.bias + c) : 0.f;
  }
  for (int c = tid; c < 256; c += SC_THREADS) srvec[c] = g.rvec != nullptr ? __ldg(g.rvec + c) : 0.f;
  if (warp == 0) tmem_alloc(&ctl->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 1) {
    // ------------------------------ weight loader ------------------------------
    if (lane == 0) {
      int kbg = 0;
      for (long long p = blockIdx.x; p < npairs; p += gridDim.x) {
        for (int s = 0; s < g.nsteps; s++) {
          const SdfStep& S = g.st[s];
          const int Nc = (S.N + 15) & ~15;
          for (int kb = 0; kb < S.KB; kb++) {
            for (int h = 0; h * 128 < Nc; h++, kbg++) {
              const int rows = min(128, Nc - h * 128);
              const uint32_t bytes = S.bmn ? (uint32_t)((rows + 63) >> 6) * 8192u : (uint32_t)rows * 128u;
              const int stg = kbg % SC_WSTAGES;
              if (kbg >= SC_WSTAGES) mbar_wait(&ctl->wempty[stg], ((kbg / SC_WSTAGES) - 1) & 1);
              mbar_arrive_expect_tx(&ctl->wfull[stg], bytes);
              bulk_g2s(sW0 + stg * CH_WBYTES, S.wimg + (size_t)kb * TC_B_BYTES + (size_t)h * CH_WBYTES, bytes,
                       &ctl->wfull[stg]);
            }
          }
        }
      }
    }
  } else if (warp == 0) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      int kbg = 0, lg = 0;
      for (long long p = blockIdx.x; p < npairs; p += gridDim.x) {
        for (int s = 0; s < g.nsteps; s++, lg++) {
          const SdfStep& S = g.st[s];
          const int Nc = (S.N + 15) & ~15;
          mbar_wait(&ctl->a_ready[0], lg & 1);
          mbar_wait(&ctl->a_ready[1], lg & 1);
          tc_fence_after();
          for (int kb = 0; kb < S.KB; kb++) {
            for (int h = 0; h * 128 < Nc; h++, kbg++) {
              const int rows = min(128, Nc - h * 128);
              const uint32_t idesc = make_idesc(rows, 0, S.bmn);
              const int stg = kbg % SC_WSTAGES;
              mbar_wait(&ctl->wfull[stg], (kbg / SC_WSTAGES) & 1);
              tc_fence_after();
              const uint32_t b_addr = smem_u32(sW0 + stg * CH_WBYTES);
#pragma unroll
              for (int t = 0; t < 2; t++) {
                const uint32_t a_addr = smem_u32(sAt[t]) + kb * TC_A_BYTES;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                  const uint64_t bd = S.bmn ? make_desc(b_addr + k * 2048, 8192, 1024) : make_desc(b_addr + k * 32, 16, 1024);
                  umma_bf16(tmem_base + t * 256 + h * 128, make_desc(a_addr + k * 32, 16, 1024), bd, idesc,
                            (kb > 0 || k > 0) ? 1 : 0);
                }
              }
              umma_commit(&ctl->wempty[stg]);
            }
          }
          umma_commit(&ctl->acc_full[0]);
          umma_commit(&ctl->acc_full[1]);
        }
      }
      tc_fence_before();
    }
  } else {
    // ------------------------------ operand builders / epilogue ------------------------------
    const int t = (warp - 2) >> 3;                 // tile of the pair
    const int wslot = (warp - 2) & 7;
    const int grp = wslot >> 2;                    // column interleave group (0/1)
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;             // row within the tile == TMEM lane
    const int r7 = r & 7;
    const int rowoff = (r >> 3) * 1024 + r7 * 128; // byte offset of the row inside a 128 x 64 BF16 block
    uint8_t* rowA = sAt[t] + rowoff;
    const uint32_t taddr = tmem_base + t * 256 + ((uint32_t)(quarter * 32) << 16);
    const float beta = g.beta, inv_beta = 1.f / g.beta;
    int lg = 0;
    for (long long p = blockIdx.x; p < npairs; p += gridDim.x) {
      const long long tile = 2 * p + t;
      const bool tile_ok = tile < ntiles;
      const long long m = tile * 128 + r;
      const bool valid = tile_ok && m < g.M;
      const size_t tile_img = (size_t)(tile_ok ? tile : 0) * SC_SLOT_BLOCKS * TC_A_BYTES + rowoff;   // 4-block images

      for (int s = 0; s < g.nsteps; s++, lg++) {
        const SdfStep& S = g.st[s];
        // ---- operand for this step, when it does not come from the previous step's epilogue ----
        if (S.src == SRC_PE || S.src == SRC_TAN) {
          if (grp == 0) {
#pragma unroll
            for (int c = 0; c < 8; c++) *reinterpret_cast<uint4*>(rowA + c * 16) = make_uint4(0u, 0u, 0u, 0u);
            if (valid)
              gen_row(S.src == SRC_PE ? g.gen : g.gen_t, m, [&](int j, float val) {
                if (j < TC_BK)
                  *reinterpret_cast<unsigned short*>(rowA + ((((j >> 3) ^ r7) & 7) << 4) + ((j & 7) << 1)) = f32_to_bf16_bits(val);
              });
            if (g.pe_img != nullptr && tile_ok) {
              uint8_t* dst = reinterpret_cast<uint8_t*>(g.pe_img) + (size_t)tile * TC_A_BYTES + rowoff;
#pragma unroll
              for (int c = 0; c < 8; c++) *reinterpret_cast<uint4*>(dst + c * 16) = *reinterpret_cast<const uint4*>(rowA + c * 16);
            }
          }
        } else if (S.src == SRC_MEM) {
          const bool vec_ok = (g.ldm & 3) == 0 && (reinterpret_cast<uintptr_t>(g.mem) & 15) == 0;
          const int nch = ((g.kmem + TC_BK - 1) / TC_BK) * 8;
          for (int rr = wslot * 16; rr < wslot * 16 + 16; rr++) {
            const long long mm = tile * 128 + rr;
            const bool rv = tile_ok && mm < g.M;
            const float* src = g.mem + mm * g.ldm;
            uint8_t* drow = sAt[t] + (rr >> 3) * 1024 + (rr & 7) * 128;
            for (int ch = lane; ch < nch; ch += 32) {
              const int c = ch * 8;
              float v[8];
              if (rv && vec_ok && c + 8 <= g.kmem) {
                const float4 lo = __ldg(reinterpret_cast<const float4*>(src + c));
                const float4 hi = __ldg(reinterpret_cast<const float4*>(src + c + 4));
                v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
              } else {
#pragma unroll
                for (int j = 0; j < 8; j++) v[j] = (rv && c + j < g.kmem) ? __ldg(src + c + j) : 0.f;
              }
              *reinterpret_cast<uint4*>(drow + (ch >> 3) * TC_A_BYTES + ((((ch & 7) ^ (rr & 7)) & 7) << 4)) = f32x8_to_bf16(v);
            }
          }
        }
        if (s == 0 || S.src != SRC_CHAIN) {
          tc_fence_before();
          fence_proxy_async();
          mbar_arrive(&ctl->a_ready[t]);
        }

        const int N = S.N, Nc = (N + 15) & ~15, mode = S.mode;
        const bool writes_operand = mode != SC_G0;
        const int cover = writes_operand ? 256 : ((Nc + 31) & ~31);
        const float* sb = sbias + (S.bias_slot >= 0 ? S.bias_slot : 0) * 256;
        const bool need_h = (mode == SC_SPMUL || mode == SC_SWEEP || mode == SC_SDFBWD) && tile_ok;
        const bool need_q = (mode == SC_SWEEP || mode == SC_SDFBWD || mode == SC_FEATQ) && S.q != nullptr && tile_ok;
        const uint8_t* hp = reinterpret_cast<const uint8_t*>(S.h) + tile_img;
        const uint8_t* qp = reinterpret_cast<const uint8_t*>(S.q) + tile_img;
        uint8_t* op = reinterpret_cast<uint8_t*>(S.img_out) + tile_img;
        uint8_t* ep = reinterpret_cast<uint8_t*>(S.e_out) + tile_img;
        const bool st_img = S.img_out != nullptr && tile_ok;
        const float hscale = S.hscale, oscale = S.oscale;
        const float rsv = (mode == SC_SDFBWD && S.use_rs && valid) ? __ldg(g.rs + m) * g.rscale : 0.f;
        float dot = 0.f;

        // auxiliary rows of the first chunk are requested before the accumulator is waited for
        uint4 ah0 = make_uint4(0u, 0u, 0u, 0u), ah1 = ah0, aq0 = ah0, aq1 = ah0;
        {
          const int n = grp * 32;
          const int o0 = (n >> 6) * TC_A_BYTES + (((((n & 63) >> 3)) ^ r7) << 4);
          const int o1 = (n >> 6) * TC_A_BYTES + (((((n & 63) >> 3) + 1) ^ r7) << 4);
          if (need_h && n < N) { ah0 = __ldcg(reinterpret_cast<const uint4*>(hp + o0)); ah1 = __ldcg(reinterpret_cast<const uint4*>(hp + o1)); }
          if (need_q && n < N) { aq0 = __ldcg(reinterpret_cast<const uint4*>(qp + o0)); aq1 = __ldcg(reinterpret_cast<const uint4*>(qp + o1)); }
        }
        if (mode == SC_G0 && S.q != nullptr) slot_bar(t);   // the addend's columns were written by both column groups
        mbar_wait(&ctl->acc_full[t], lg & 1);
        tc_fence_after();

#pragma unroll 1
        for (int n = grp * 32; n < cover; n = ((n & 16) ? n + 48 : n + 16)) {
          const int o0 = (n >> 6) * TC_A_BYTES + (((((n & 63) >> 3)) ^ r7) << 4);
          const int o1 = (n >> 6) * TC_A_BYTES + (((((n & 63) >> 3) + 1) ^ r7) << 4);
          float a[16], hv[16], qv[16], y[16];
          bf16x8_to_f32(ah0, hv); bf16x8_to_f32(ah1, hv + 8);
          bf16x8_to_f32(aq0, qv); bf16x8_to_f32(aq1, qv + 8);
          {
            // prefetch the next chunk's auxiliary rows
            const int nn = (n & 16) ? n + 48 : n + 16;
            const int p0 = (nn >> 6) * TC_A_BYTES + (((((nn & 63) >> 3)) ^ r7) << 4);
            const int p1 = (nn >> 6) * TC_A_BYTES + (((((nn & 63) >> 3) + 1) ^ r7) << 4);
            const bool more = nn < cover && nn < N;
            if (need_h && more) { ah0 = __ldcg(reinterpret_cast<const uint4*>(hp + p0)); ah1 = __ldcg(reinterpret_cast<const uint4*>(hp + p1)); }
            if (need_q && more) { aq0 = __ldcg(reinterpret_cast<const uint4*>(qp + p0)); aq1 = __ldcg(reinterpret_cast<const uint4*>(qp + p1)); }
          }
          if (n < Nc) tmem_ld16(taddr + n, a);
          else {
#pragma unroll
            for (int j = 0; j < 16; j++) a[j] = 0.f;
          }
          if (mode == SC_SOFTPLUS) {
#pragma unroll
            for (int j = 0; j < 16; j++) {
              float v = softplus_beta_fast(a[j] + sb[n + j], beta, inv_beta);
              v = (valid && n + j < N) ? v : 0.f;
              if (S.dot) {
                dot += v * srvec[n + j];
                qv[j] = softplus_grad_from_act_fast(v, beta) * srvec[n + j];   // q_{L-1} from the unrounded activation
              }
              y[j] = v * oscale;
            }
            if (S.dot && S.e_out != nullptr && tile_ok) {
              *reinterpret_cast<uint4*>(ep + o0) = f32x8_to_bf16(qv);
              *reinterpret_cast<uint4*>(ep + o1) = f32x8_to_bf16(qv + 8);
            }
          } else if (mode == SC_FEATQ) {
            if (valid && n < N) {
              float f[16];
#pragma unroll
              for (int j = 0; j < 16; j++) f[j] = a[j] + sb[n + j];
              row_store16(S.out, S.ldo, m, n, N - n, f);
            }
            // the reverse chain starts from q_{L-1}, written to its image by the previous step (same thread, same chunk)
#pragma unroll
            for (int j = 0; j < 16; j++) y[j] = (valid && n + j < N) ? qv[j] : 0.f;
          } else if (mode == SC_SPMUL) {
#pragma unroll
            for (int j = 0; j < 16; j++) {
              const float sg = softplus_grad_from_act_fast(hv[j] * hscale, beta);
              y[j] = (valid && n + j < S.csplit) ? sg * a[j] * oscale : 0.f;
            }
            if (S.out != nullptr && valid && n + 16 > S.csplit) {
#pragma unroll 1
              for (int j = 0; j < 16; j++) {
                const int nn = n + j;
                if (nn >= S.csplit && nn < N) S.out[m * S.ldo + nn - S.csplit] = a[j] * oscale;
              }
            }
          } else if (mode == SC_G0) {
            if (valid) {
#pragma unroll 1
              for (int j = 0; j < 16; j++) {
                const int nn = n + j;
                if (nn < N) S.out[m * S.ldo + nn] = a[j] + (S.q != nullptr ? __ldcg(S.q + m * S.ldo + nn) : 0.f);
              }
            }
          } else if (mode == SC_SWEEP) {
            float e[16];
#pragma unroll
            for (int j = 0; j < 16; j++) {
              const float sg = softplus_grad_from_act_fast(hv[j] * hscale, beta);
              const bool ok = valid && n + j < N;
              y[j] = ok ? sg * a[j] * oscale : 0.f;
              e[j] = ok ? beta * (1.f - sg) * qv[j] * a[j] : 0.f;
            }
            if (S.e_out != nullptr && tile_ok) {
              *reinterpret_cast<uint4*>(ep + o0) = f32x8_to_bf16(e);
              *reinterpret_cast<uint4*>(ep + o1) = f32x8_to_bf16(e + 8);
            }
          } else {  // SC_SDFBWD
#pragma unroll
            for (int j = 0; j < 16; j++) {
              const float sg = softplus_grad_from_act_fast(hv[j] * hscale, beta);
              y[j] = (valid && n + j < S.csplit) ? sg * (a[j] + rsv * srvec[n + j]) * oscale + qv[j] : 0.f;
            }
          }
          if (writes_operand) {
            const uint4 c0 = f32x8_to_bf16(y), c1 = f32x8_to_bf16(y + 8);
            *reinterpret_cast<uint4*>(rowA + o0) = c0;
            *reinterpret_cast<uint4*>(rowA + o1) = c1;
            if (st_img) {
              *reinterpret_cast<uint4*>(op + o0) = c0;
              *reinterpret_cast<uint4*>(op + o1) = c1;
            }
          }
        }
        if (S.append) {
          // skip connection (fields.py:83-84): PE columns (value or tangent form) follow the N hidden columns, once
          // BOTH column groups of the tile have written their chunks; the image gets the touched 16-byte chunks again
          slot_bar(t);
          if (grp == 1) {
            if (valid)
              gen_row(mode == SC_SWEEP ? g.gen_t : g.gen, m, [&](int j, float val) {
                const int col = N + j;
                if (col < 256)
                  *reinterpret_cast<unsigned short*>(rowA + (col >> 6) * TC_A_BYTES + (((((col & 63) >> 3) ^ r7) & 7) << 4) +
                                                     ((col & 7) << 1)) = f32_to_bf16_bits(val * rsqrt2);
              });
            if (st_img) {
              for (int ch = N >> 3; ch < 32; ch++) {
                const int o = (ch >> 3) * TC_A_BYTES + ((((ch & 7) ^ r7) & 7) << 4);
                *reinterpret_cast<uint4*>(op + o) = *reinterpret_cast<const uint4*>(rowA + o);
              }
            }
          }
        }
        if (S.dot) {
          if (grp == 0) sdot[t * 128 + r] = dot;
          slot_bar(t);
          if (grp == 1 && valid)
            g.sdf_out[m] = (sdot[t * 128 + r] + dot + __ldg(g.b_last)) * g.sdf_scale;
        }
        tc_fence_before();
        if (s + 1 < g.nsteps && g.st[s + 1].src == SRC_CHAIN) {
          fence_proxy_async();
          mbar_arrive(&ctl->a_ready[t]);
        }
      }
    }
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

inline int sdf_chain_prepare() {
  static int done = 0;
  if (done) return 0;
  if (cudaFuncSetAttribute(sdf_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SC_SMEM_BYTES) != cudaSuccess)
    return 1;
  done = 1;
  return 0;
}
inline void sdf_chain_launch(const SdfChainArgs& g, double flops, cudaStream_t st) {
  const long long npairs = ((g.M + 127) / 128 + 1) / 2;
  const int sms = tc_num_sms();
  const int grid = (int)(npairs < sms ? npairs : sms);
  prof_begin(PC_TC_MLP, flops, 0.0, st);
  sdf_chain_kernel<<<grid, SC_THREADS, SC_SMEM_BYTES, st>>>(g);
  prof_end(st);
}

}  // namespace fneus

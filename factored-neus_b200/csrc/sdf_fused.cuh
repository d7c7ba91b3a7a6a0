// Fused SDF forward (sdf value only): the whole layer chain of SDFNetwork.sdf (fields.py:74-95) for a pair of
// 128-point tiles in ONE persistent kernel -- positional encoding generated into the layer-0 operand, hidden
// activations kept on chip (BF16, swizzled K-major in shared memory <-> FP32 accumulators in TMEM), weights
// streamed from L2 as pre-packed BF16 images by bulk async copies, the final sdf row folded into the last
// hidden layer's epilogue as an FP32 dot product.  Used by hierarchical up-sampling (112 of the 240 SDF
// evaluations per training ray), the grid query and the stage-2 coarse marching.
//
//   warp 0      : MMA issuer (+ TMEM alloc: 2 accumulators x 256 columns, one per tile of the pair)
//   warp 1      : weight-image loader (2-stage ring of [<=256 x 64] tiles, shared by both tiles of the pair)
//   warps 2-17  : epilogue / operand writers, 8 per tile: thread = one point (TMEM lane), 16 columns at a time:
//                 bias + softplus (MUFU) -> BF16 -> 16-byte stores into the tile's next-layer A operand
#pragma once
#include "gemm_tc.cuh"

namespace fneus {

constexpr int FZ_THREADS = 576, FZ_WSTAGES = 2, FZ_MAXL = 12;
struct FusedSdfArgs {
  int L;                         // hidden layers; linears 0..L (linear L is folded in as a dot with row 0)
  int in[FZ_MAXL], out[FZ_MAXL]; // dims of linears 0..L-1
  const uint8_t* img[FZ_MAXL];   // K-major weight images of linears 0..L-1
  const float* bias[FZ_MAXL];
  const float* w_last;           // row 0 of linear L  [in_L]
  const float* b_last;           // its bias
  int skip;                      // linear index whose input is cat([h, PE])/sqrt(2), -1 none
  float beta, scale, out_sign;
  GenSpec gen;                   // PE(x * scale)
  const float* x; float* sdf_out; long long M;
};
struct FZSmem {
  uint64_t wfull[FZ_WSTAGES], wempty[FZ_WSTAGES];
  uint64_t a_ready[2], acc_full[2];
  uint32_t tmem_base;
};
constexpr int FZ_A_BYTES = 4 * TC_A_BYTES;                 // one tile's operand: 128 rows x 256 columns BF16
constexpr int FZ_SMEM_BYTES = 2 * FZ_A_BYTES + FZ_WSTAGES * TC_B_BYTES + (FZ_MAXL * 256 + 256 + 2 * 128) * 4 + 1024 + 256;

__global__ void __launch_bounds__(FZ_THREADS, 1) sdf_fused_fwd_kernel(FusedSdfArgs g) {
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* base = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared
                                                 // address space (an integer round trip turned every access into a generic LD.E / ST.E)
  uint8_t* sAt[2] = {base, base + FZ_A_BYTES};
  uint8_t* sW[FZ_WSTAGES];
#pragma unroll
  for (int s = 0; s < FZ_WSTAGES; s++) sW[s] = base + 2 * FZ_A_BYTES + s * TC_B_BYTES;
  float* sbias = reinterpret_cast<float*>(base + 2 * FZ_A_BYTES + FZ_WSTAGES * TC_B_BYTES);   // [L][256]
  float* swl = sbias + FZ_MAXL * 256;                                                             // [256]
  float* sdot = swl + 256;                                                                        // [2][128]
  FZSmem* ctl = reinterpret_cast<FZSmem*>(sdot + 256);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long ntiles = (g.M + 127) / 128;
  const long long npairs = (ntiles + 1) / 2;
  const float rsqrt2 = 0.70710678118654752440f;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < FZ_WSTAGES; s++) { mbar_init(&ctl->wfull[s], 1); mbar_init(&ctl->wempty[s], 1); }
#pragma unroll
    for (int t = 0; t < 2; t++) { mbar_init(&ctl->a_ready[t], 256); mbar_init(&ctl->acc_full[t], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < g.L * 256; i += FZ_THREADS) {
    int l = i >> 8, c = i & 255;
    sbias[i] = c < g.out[l] ? __ldg(g.bias[l] + c) : 0.f;
  }
  for (int i = tid; i < 256; i += FZ_THREADS) swl[i] = i < g.out[g.L - 1] ? __ldg(g.w_last + i) : 0.f;
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
        for (int l = 0; l < g.L; l++) {
          const int Nc = (g.out[l] + 15) & ~15;
          const int KB = (g.in[l] + TC_BK - 1) / TC_BK;
          const uint32_t bytes = (uint32_t)Nc * 128u;
          for (int kb = 0; kb < KB; kb++, kbg++) {
            const int s = kbg % FZ_WSTAGES;
            if (kbg >= FZ_WSTAGES) mbar_wait(&ctl->wempty[s], ((kbg / FZ_WSTAGES) - 1) & 1);
            mbar_arrive_expect_tx(&ctl->wfull[s], bytes);
            bulk_g2s(sW[s], g.img[l] + (size_t)kb * TC_B_BYTES, bytes, &ctl->wfull[s]);
          }
        }
      }
    }
  } else if (warp == 0) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      int kbg = 0, lg = 0;     // lg: global layer counter (phase of a_ready / acc_full)
      for (long long p = blockIdx.x; p < npairs; p += gridDim.x) {
        for (int l = 0; l < g.L; l++, lg++) {
          const int Nc = (g.out[l] + 15) & ~15;
          const int KB = (g.in[l] + TC_BK - 1) / TC_BK;
          const uint32_t idesc = make_idesc(Nc, 0, 0);
          mbar_wait(&ctl->a_ready[0], lg & 1);
          mbar_wait(&ctl->a_ready[1], lg & 1);
          tc_fence_after();
          for (int kb = 0; kb < KB; kb++, kbg++) {
            const int s = kbg % FZ_WSTAGES;
            mbar_wait(&ctl->wfull[s], (kbg / FZ_WSTAGES) & 1);
            tc_fence_after();
            const uint32_t b_addr = smem_u32(sW[s]);
#pragma unroll
            for (int t = 0; t < 2; t++) {
              const uint32_t a_addr = smem_u32(sAt[t]) + kb * TC_A_BYTES;
#pragma unroll
              for (int k = 0; k < 4; k++)
                umma_bf16(tmem_base + t * 256, make_desc(a_addr + k * 32, 16, 1024), make_desc(b_addr + k * 32, 16, 1024),
                          idesc, (kb > 0 || k > 0) ? 1 : 0);
            }
            umma_commit(&ctl->wempty[s]);
          }
          umma_commit(&ctl->acc_full[0]);
          umma_commit(&ctl->acc_full[1]);
        }
      }
      tc_fence_before();
    }
  } else {
    // ------------------------------ epilogue / operand writers ------------------------------
    const int t = (warp - 2) >> 3;                 // tile of the pair
    const int grp = ((warp - 2) & 7) >> 2;         // column half-interleave group (0/1)
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;             // row within the tile == TMEM lane
    const int r7 = r & 7;
    uint8_t* rowA = sAt[t] + (r >> 3) * 1024 + r7 * 128;     // + kb*16384 + swizzled chunk
    const uint32_t taddr = tmem_base + t * 256 + ((uint32_t)(quarter * 32) << 16);
    const float beta = g.beta, inv_beta = 1.f / g.beta;
    int lg = 0;
    for (long long p = blockIdx.x; p < npairs; p += gridDim.x) {
      const long long m = (2 * p + t) * 128 + r;
      const bool valid = m < g.M;
      // layer-0 operand: PE(x) into block 0 (only group 0 writes; both groups arrive)
      if (grp == 0) {
#pragma unroll
        for (int c = 0; c < 8; c++) *reinterpret_cast<uint4*>(rowA + c * 16) = make_uint4(0u, 0u, 0u, 0u);
        if (valid)
          gen_row(g.gen, m, [&](int j, float val) {
            if (j < TC_BK)
              *reinterpret_cast<unsigned short*>(rowA + ((((j >> 3) ^ r7) & 7) << 4) + ((j & 7) << 1)) = f32_to_bf16_bits(val);
          });
      }
      if (grp == 0) sdot[t * 128 + r] = 0.f;
      fence_proxy_async();
      mbar_arrive(&ctl->a_ready[t]);
      for (int l = 0; l < g.L; l++, lg++) {
        const int N = g.out[l];
        const int Nc = (N + 15) & ~15;
        const bool last = l == g.L - 1;
        const bool pre_skip = (l + 1 == g.skip);
        const float oscale = pre_skip ? rsqrt2 : 1.f;
        mbar_wait(&ctl->acc_full[t], lg & 1);
        tc_fence_after();
        float dot = 0.f;
        // this group's 32-column chunks; cover all 256 operand columns so stale columns never reach the next MMA
#pragma unroll 1
        for (int c0 = grp * 32; c0 < 256; c0 += 64) {
#pragma unroll 1
          for (int hs = 0; hs < 32; hs += 16) {
            const int n = c0 + hs;
            float a[16], y[16];
            if (n < Nc) tmem_ld16(taddr + n, a);
            else {
#pragma unroll
              for (int j = 0; j < 16; j++) a[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 16; j++) {
              float v = softplus_beta_fast(a[j] + sbias[l * 256 + n + j], beta, inv_beta);
              v = (n + j < N) ? v : 0.f;
              if (last) dot += v * swl[n + j];
              y[j] = v * oscale;
            }
            if (!last) {
              // next layer's operand: block (n>>6), 16-byte chunks ((n&63)>>3) and +1
              uint8_t* dst = rowA + (n >> 6) * TC_A_BYTES;
              const int ch = (n & 63) >> 3;
#pragma unroll
              for (int i = 0; i < 2; i++) {
                uint2 lo = pack_bf16x4(make_float4(y[i * 8], y[i * 8 + 1], y[i * 8 + 2], y[i * 8 + 3]));
                uint2 hi = pack_bf16x4(make_float4(y[i * 8 + 4], y[i * 8 + 5], y[i * 8 + 6], y[i * 8 + 7]));
                *reinterpret_cast<uint4*>(dst + ((((ch + i) ^ r7) & 7) << 4)) = make_uint4(lo.x, lo.y, hi.x, hi.y);
              }
            }
          }
        }
        if (!last) {
          if (pre_skip) {
            // append PE(x)/sqrt(2) after the N hidden columns (fields.py:83-84) once BOTH column groups of the
            // tile have written their (zero-padded) chunks
            asm volatile("bar.sync %0, 256;" ::"r"(1 + t) : "memory");
          }
          if (pre_skip && valid && grp == 1) {
            gen_row(g.gen, m, [&](int j, float val) {
              const int col = N + j;
              if (col < 256)
                *reinterpret_cast<unsigned short*>(rowA + (col >> 6) * TC_A_BYTES + (((((col & 63) >> 3) ^ r7) & 7) << 4) +
                                                   ((col & 7) << 1)) = f32_to_bf16_bits(val * rsqrt2);
            });
          }
          tc_fence_before();
          fence_proxy_async();
          mbar_arrive(&ctl->a_ready[t]);
        } else {
          atomicAdd(&sdot[t * 128 + r], dot);
          tc_fence_before();
          // both groups of this tile must have added their halves: named barrier over the tile's 8 warps
          asm volatile("bar.sync %0, 256;" ::"r"(1 + t) : "memory");
          if (grp == 0 && valid) g.sdf_out[m] = (sdot[t * 128 + r] + __ldg(g.b_last)) * g.out_sign / g.scale;
          asm volatile("bar.sync %0, 256;" ::"r"(1 + t) : "memory");
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

}  // namespace fneus

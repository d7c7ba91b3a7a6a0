// Fused MLP chains (BF16 tensor-core mode): a whole sequence of dense layers for a pair of 128-row tiles in ONE
// persistent kernel.  The running operand stays on chip -- BF16, 128B-swizzled K-major in shared memory, FP32
// accumulators in TMEM -- and every layer's weights arrive as pre-packed images by bulk async copies (L2 hits
// after the first CTA).  What a layered execution would write to and re-read from HBM between two launches is
// instead written ONCE, as an activation image, by bulk shared->global copies of the operand itself, and only
// when the backward pass needs it.
//
// Used for the ReLU networks of the render path: RenderingNetwork (fields.py:114-175), RefColor (fields.py:271-335),
// Lvis (fields.py:338-369) -- forward (steps RELU.. + OUT) and backward-data (steps MASK.. + OUT).
//
//   warp 0      : MMA issuer (+ TMEM alloc: one 256-column accumulator per tile of the pair)
//   warp 1      : weight-image loader, ring of 3 half-tiles (128 output columns x 64 reduction, 16 KB)
//   warps 2-17  : 8 per tile: build the first operand (generated PE columns, FP32 -> BF16 conversion of the
//                 memory segment), then per step: thread = one row (TMEM lane), 16 columns at a time,
//                 bias / ReLU / mask -> BF16 -> next operand (+ bulk image store), or FP32 results to HBM
#pragma once
#include "gemm_tc.cuh"

namespace fneus {

constexpr int CH_THREADS = 576, CH_WSTAGES = 3, CH_MAXS = 14, CH_SLOT_BLOCKS = 5, CH_WBYTES = 16384;
enum ChainMode { CH_RELU = 0, CH_MASK = 1, CH_OUT = 2 };

struct ChainStep {
  const uint8_t* wimg;  // weight image of this step ([1 n-chunk][KB] tiles of 32 KB)
  const float* bias;    // [N] or null
  float* img_out;       // RELU / MASK: activation image of the result ([M, ceil(N/64) blocks]) or null
  const float* mask;    // MASK: image whose sign gates the result (the forward activation)
  float* out;           // OUT: FP32 [M, ldo]
  int KB;               // reduction blocks of the operand this step reads
  int N;                // output columns (<= 256)
  int mode;             // ChainMode
  int bmn;              // weight image is MN-major (backward-data) instead of K-major
  int mask_kbs;
  int ldo, act, accumulate;   // OUT: act 0 linear / 1 sigmoid ; accumulate: out += result
};
struct ChainArgs {
  int nsteps;
  ChainStep st[CH_MAXS];
  GenSpec gen;          // generated columns of the first operand (blocks 0 .. kb_gen-1)
  const float* mem;     // memory segment of the first operand: FP32 row-major (ldm > 0) or image (ldm < 0)
  int ldm, kmem;
  float* a0_img;        // optional image copy of the first operand ([M, kb_gen + kb_mem blocks])
  long long M;
};
struct CHSmem {
  uint64_t wfull[CH_WSTAGES], wempty[CH_WSTAGES];
  uint64_t a_ready[2], acc_full[2];
  uint32_t tmem_base;
};
constexpr int CH_SLOT_BYTES = CH_SLOT_BLOCKS * TC_A_BYTES;
constexpr int CH_SMEM_BYTES = 2 * CH_SLOT_BYTES + CH_WSTAGES * CH_WBYTES + CH_MAXS * 256 * 4 + 1024 + 256;

__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void slot_bar(int t) { asm volatile("bar.sync %0, 256;" ::"r"(1 + t) : "memory"); }

__global__ void __launch_bounds__(CH_THREADS, 1) mlp_chain_kernel(const __grid_constant__ ChainArgs g) {
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sAt[2] = {base, base + CH_SLOT_BYTES};
  uint8_t* sW0 = base + 2 * CH_SLOT_BYTES;
  float* sbias = reinterpret_cast<float*>(sW0 + CH_WSTAGES * CH_WBYTES);             // [nsteps][256]
  CHSmem* ctl = reinterpret_cast<CHSmem*>(sbias + CH_MAXS * 256);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long ntiles = (g.M + 127) / 128;
  const long long npairs = (ntiles + 1) / 2;
  const int kb_gen = (g.gen.ncols + TC_BK - 1) / TC_BK;
  const int kb_mem = (g.kmem + TC_BK - 1) / TC_BK;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < CH_WSTAGES; s++) { mbar_init(&ctl->wfull[s], 1); mbar_init(&ctl->wempty[s], 1); }
#pragma unroll
    for (int t = 0; t < 2; t++) { mbar_init(&ctl->a_ready[t], 256); mbar_init(&ctl->acc_full[t], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < g.nsteps * 256; i += CH_THREADS) {
    const int s = i >> 8, c = i & 255;
    sbias[i] = (g.st[s].bias != nullptr && c < g.st[s].N) ? __ldg(g.st[s].bias + c) : 0.f;
  }
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
          const ChainStep& S = g.st[s];
          const int Nc = (S.N + 15) & ~15;
          for (int kb = 0; kb < S.KB; kb++) {
            for (int h = 0; h * 128 < Nc; h++, kbg++) {
              const int rows = min(128, Nc - h * 128);
              const uint32_t bytes = S.bmn ? (uint32_t)((rows + 63) >> 6) * 8192u : (uint32_t)rows * 128u;
              const int stg = kbg % CH_WSTAGES;
              if (kbg >= CH_WSTAGES) mbar_wait(&ctl->wempty[stg], ((kbg / CH_WSTAGES) - 1) & 1);
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
          const ChainStep& S = g.st[s];
          const int Nc = (S.N + 15) & ~15;
          mbar_wait(&ctl->a_ready[0], lg & 1);
          mbar_wait(&ctl->a_ready[1], lg & 1);
          tc_fence_after();
          for (int kb = 0; kb < S.KB; kb++) {
            for (int h = 0; h * 128 < Nc; h++, kbg++) {
              const int rows = min(128, Nc - h * 128);
              const uint32_t idesc = make_idesc(rows, 0, S.bmn);
              const int stg = kbg % CH_WSTAGES;
              mbar_wait(&ctl->wfull[stg], (kbg / CH_WSTAGES) & 1);
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
    const int wslot = (warp - 2) & 7;              // warp within the tile's group
    const int grp = wslot >> 2;                    // column interleave group (0/1)
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;             // row within the tile == TMEM lane
    const int r7 = r & 7;
    const bool elected = wslot == 0 && lane == 0;
    uint8_t* rowA = sAt[t] + (r >> 3) * 1024 + r7 * 128;     // + kb*16384 + swizzled chunk
    const uint32_t taddr = tmem_base + t * 256 + ((uint32_t)(quarter * 32) << 16);
    bool pending = false;                          // bulk stores of this tile's operand buffer in flight
    int lg = 0;
    for (long long p = blockIdx.x; p < npairs; p += gridDim.x) {
      const long long tile = 2 * p + t;
      const bool tile_ok = tile < ntiles;
      const long long m = tile * 128 + r;
      const bool valid = tile_ok && m < g.M;
      if (pending) {
        if (elected) bulk_wait_read0();
        slot_bar(t);
        pending = false;
      }
      // ---- first operand: [generated blocks | memory blocks] ----
      if (kb_gen > 0 && grp == 0) {
        for (int b = 0; b < kb_gen; b++)
#pragma unroll
          for (int c = 0; c < 8; c++) *reinterpret_cast<uint4*>(rowA + b * TC_A_BYTES + c * 16) = make_uint4(0u, 0u, 0u, 0u);
        if (valid)
          gen_row(g.gen, m, [&](int j, float val) {
            *reinterpret_cast<unsigned short*>(rowA + (j >> 6) * TC_A_BYTES + (((((j & 63) >> 3) ^ r7) & 7) << 4) +
                                               ((j & 7) << 1)) = f32_to_bf16_bits(val);
          });
      }
      if (kb_mem > 0) {
        uint8_t* dst0 = sAt[t] + kb_gen * TC_A_BYTES;
        if (g.ldm < 0) {
          // image source: the tile bytes are the operand
          const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(g.mem) +
                                                            (size_t)(tile_ok ? tile : 0) * (size_t)(-g.ldm) * TC_A_BYTES);
          const int nvec = kb_mem * (TC_A_BYTES / 16);
          for (int i = wslot * 32 + lane; i < nvec; i += 256)
            reinterpret_cast<uint4*>(dst0)[i] = tile_ok ? __ldg(src + i) : make_uint4(0u, 0u, 0u, 0u);
        } else {
          // FP32 row-major source: each warp converts 16 rows, a lane 8 consecutive columns at a time
          const bool vec_ok = (g.ldm & 3) == 0 && (reinterpret_cast<uintptr_t>(g.mem) & 15) == 0;
          const int nch = kb_mem * 8;              // 16-byte chunks per row
          for (int rr = wslot * 16; rr < wslot * 16 + 16; rr++) {
            const long long mm = tile * 128 + rr;
            const bool rv = tile_ok && mm < g.M;
            const float* src = g.mem + mm * g.ldm;
            uint8_t* drow = dst0 + (rr >> 3) * 1024 + (rr & 7) * 128;
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
              const uint2 lo2 = pack_bf16x4(make_float4(v[0], v[1], v[2], v[3]));
              const uint2 hi2 = pack_bf16x4(make_float4(v[4], v[5], v[6], v[7]));
              *reinterpret_cast<uint4*>(drow + (ch >> 3) * TC_A_BYTES + ((((ch & 7) ^ (rr & 7)) & 7) << 4)) =
                  make_uint4(lo2.x, lo2.y, hi2.x, hi2.y);
            }
          }
        }
      }
      tc_fence_before();
      fence_proxy_async();
      if (g.a0_img != nullptr && tile_ok) {
        slot_bar(t);
        if (elected) {
          const int kb0 = kb_gen + kb_mem;
          bulk_s2g(reinterpret_cast<uint8_t*>(g.a0_img) + (size_t)tile * kb0 * TC_A_BYTES, sAt[t], (uint32_t)kb0 * TC_A_BYTES);
          bulk_commit();
        }
        pending = true;
      }
      mbar_arrive(&ctl->a_ready[t]);

      for (int s = 0; s < g.nsteps; s++, lg++) {
        const ChainStep& S = g.st[s];
        const int N = S.N, Nc = (N + 15) & ~15, mode = S.mode;
        const bool writes_operand = mode != CH_OUT;
        const int cover = writes_operand ? ((N + 63) & ~63) : ((Nc + 31) & ~31);
        const float* sb = sbias + s * 256;
        mbar_wait(&ctl->acc_full[t], lg & 1);
        tc_fence_after();
        if (writes_operand && pending) {
          if (elected) bulk_wait_read0();
          slot_bar(t);
          pending = false;
        }
#pragma unroll 1
        for (int c0 = grp * 32; c0 < cover; c0 += 64) {
#pragma unroll 1
          for (int hs = 0; hs < 32; hs += 16) {
            const int n = c0 + hs;
            float a[16], y[16];
            if (n < Nc) tmem_ld16(taddr + n, a);
            else {
#pragma unroll
              for (int j = 0; j < 16; j++) a[j] = 0.f;
            }
            if (mode == CH_OUT) {
              if (valid && n < N) {
                const int nv = N - n;
                if (S.accumulate) row_load16(S.out, S.ldo, m, n, nv, y);
#pragma unroll
                for (int j = 0; j < 16; j++) {
                  float v = a[j] + sb[n + j];
                  if (S.act == 1) v = sigmoid_fast(v);
                  y[j] = S.accumulate ? y[j] + v : v;
                }
                row_store16(S.out, S.ldo, m, n, nv, y);
              }
            } else {
              if (mode == CH_RELU) {
#pragma unroll
                for (int j = 0; j < 16; j++) y[j] = (valid && n + j < N) ? fmaxf(a[j] + sb[n + j], 0.f) : 0.f;
              } else {
                float h[16];
                if (valid && n < N) row_load16(S.mask, -S.mask_kbs, m, n, 16, h);
                else {
#pragma unroll
                  for (int j = 0; j < 16; j++) h[j] = 0.f;
                }
#pragma unroll
                for (int j = 0; j < 16; j++) y[j] = (h[j] > 0.f && n + j < N) ? a[j] : 0.f;
              }
              uint8_t* dst = rowA + (n >> 6) * TC_A_BYTES;
              const int ch = (n & 63) >> 3;
#pragma unroll
              for (int i = 0; i < 2; i++) {
                const uint2 lo = pack_bf16x4(make_float4(y[i * 8], y[i * 8 + 1], y[i * 8 + 2], y[i * 8 + 3]));
                const uint2 hi = pack_bf16x4(make_float4(y[i * 8 + 4], y[i * 8 + 5], y[i * 8 + 6], y[i * 8 + 7]));
                *reinterpret_cast<uint4*>(dst + ((((ch + i) ^ r7) & 7) << 4)) = make_uint4(lo.x, lo.y, hi.x, hi.y);
              }
            }
          }
        }
        tc_fence_before();
        if (writes_operand) {
          fence_proxy_async();
          if (S.img_out != nullptr && tile_ok) {
            slot_bar(t);
            if (elected) {
              const int kbn = (N + 63) >> 6;
              bulk_s2g(reinterpret_cast<uint8_t*>(S.img_out) + (size_t)tile * kbn * TC_A_BYTES, sAt[t], (uint32_t)kbn * TC_A_BYTES);
              bulk_commit();
            }
            pending = true;
          }
        }
        if (s + 1 < g.nsteps) mbar_arrive(&ctl->a_ready[t]);
      }
    }
    if (elected) bulk_wait0();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

inline int chain_prepare() {
  static int done = 0;
  if (done) return 0;
  if (cudaFuncSetAttribute(mlp_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_BYTES) != cudaSuccess)
    return 1;
  done = 1;
  return 0;
}
inline void chain_launch(const ChainArgs& g, double flops, cudaStream_t st) {
  const long long npairs = ((g.M + 127) / 128 + 1) / 2;
  const int sms = tc_num_sms();
  const int grid = (int)(npairs < sms ? npairs : sms);
  prof_begin(PC_TC_MLP, flops, 0.0, st);
  mlp_chain_kernel<<<grid, CH_THREADS, CH_SMEM_BYTES, st>>>(g);
  prof_end(st);
}

}  // namespace fneus

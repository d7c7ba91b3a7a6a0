// BF16 tensor-core GEMM engine (tcgen05.mma, accumulators in TMEM) behind the same operand/epilogue description
// as the FP32 SIMT engine (gemm_simt.cuh).  One CTA = one 128-row output tile x up to 256 columns.
//
//   warps 0-3 : producers (global FP32 / generated PE columns -> BF16 -> 128B-swizzled shared memory, the canonical
//               UMMA layouts) and, once the accumulation is committed, the epilogue (tcgen05.ld -> fused epilogue).
//   warp 4    : TMEM allocation + single-thread tcgen05.mma issue, tcgen05.commit onto mbarriers.
//
//   tc_gemm_mk   <WT=true > : C = epi(A W^T)   A, W K-major
//   tc_gemm_mk   <WT=false> : C = epi(A W)     A K-major, W as MN-major B operand (rows of W are N-contiguous)
//   tc_gemm_wgrad           : dW += dY^T A     both operands MN-major (reduction index = row index), split over rows,
//                                              FP32 RED.ADD epilogue
#pragma once
#include <cuda_bf16.h>

#include "gemm_simt.cuh"

namespace fneus {

constexpr int TC_BM = 128, TC_BK = 64, TC_STAGES = 2, TC_THREADS = 160;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;          // 16 KB
constexpr int TC_B_BYTES = 256 * TC_BK * 2;            // 32 KB (max)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, int cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, int cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, int accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// 32 consecutive FP32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor, SWIZZLE_128B, version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor: BF16 x BF16 -> FP32, M = 128
__device__ __forceinline__ uint32_t make_idesc(int n, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;                       // c_format = F32
  d |= 1u << 7;                       // a_format = BF16
  d |= 1u << 10;                      // b_format = BF16
  d |= (uint32_t)(a_mn_major & 1) << 15;
  d |= (uint32_t)(b_mn_major & 1) << 16;
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(TC_BM >> 4) << 24;
  return d;
}

__device__ __forceinline__ uint2 pack_bf16x4(float4 v) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&lo);
  r.y = *reinterpret_cast<uint32_t*>(&hi);
  return r;
}
// K-major tile [rows][64]: row r, element k (multiple of 4)
__device__ __forceinline__ void sts_kmajor(uint8_t* tile, int r, int k, float4 v) {
  int off = (r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 3) ^ (r & 7)) & 7) << 4) + ((k & 7) << 1);
  *reinterpret_cast<uint2*>(tile + off) = pack_bf16x4(v);
}
// MN-major tile: 64-column blocks of [64 k-rows][64 mn] (8 KB each): k-row kk, column n (multiple of 4)
__device__ __forceinline__ void sts_mnmajor(uint8_t* tile, int kk, int n, float4 v) {
  int off = (n >> 6) * 8192 + (kk >> 3) * 1024 + (kk & 7) * 128 + (((((n & 63) >> 3) ^ (kk & 7)) & 7) << 4) +
            ((n & 7) << 1);
  *reinterpret_cast<uint2*>(tile + off) = pack_bf16x4(v);
}

struct TcSmem {
  uint64_t full[TC_STAGES];
  uint64_t empty[TC_STAGES];
  uint64_t accum;
  uint32_t tmem_base;
};

__device__ __forceinline__ float4 ldw4(const float* p, bool v0, bool v1, bool v2, bool v3) {
  float4 r;
  r.x = v0 ? __ldg(p + 0) : 0.f;
  r.y = v1 ? __ldg(p + 1) : 0.f;
  r.z = v2 ? __ldg(p + 2) : 0.f;
  r.w = v3 ? __ldg(p + 3) : 0.f;
  return r;
}

// ---------------------------------------------------------------------------------------------
// C[M, N] = epi( A[M, K] * B ),  grid = (ceil(M/128), ceil(N/256))
// ---------------------------------------------------------------------------------------------
template <bool WT>
__global__ void __launch_bounds__(TC_THREADS, 2)
tc_gemm_mk_kernel(ASeg a, const float* __restrict__ W, int ldw, int wout0, int M, int N, Epi e) {
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA[TC_STAGES];
  uint8_t* sB[TC_STAGES];
#pragma unroll
  for (int s = 0; s < TC_STAGES; s++) {
    sA[s] = base + s * (TC_A_BYTES + TC_B_BYTES);
    sB[s] = sA[s] + TC_A_BYTES;
  }
  TcSmem* ctl = reinterpret_cast<TcSmem*>(base + TC_STAGES * (TC_A_BYTES + TC_B_BYTES));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long m0 = (long long)blockIdx.x * TC_BM;
  const int n0 = blockIdx.y * 256;
  const int nvalid = min(256, N - n0);                 // valid output columns of this CTA
  const int Nc = (nvalid + 15) & ~15;                  // UMMA N
  const int kb_gen = (a.gen.ncols + TC_BK - 1) / TC_BK;
  const int kb_mem = (a.kmem + TC_BK - 1) / TC_BK;
  const int KB = kb_gen + kb_mem;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < TC_STAGES; s++) { mbar_init(&ctl->full[s], 128); mbar_init(&ctl->empty[s], 1); }
    mbar_init(&ctl->accum, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) tmem_alloc(&ctl->tmem_base, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = ctl->tmem_base;

  if (warp < 4) {
    // ------------------------------ producers ------------------------------
    for (int kb = 0; kb < KB; kb++) {
      const int s = kb % TC_STAGES;
      if (kb >= TC_STAGES) mbar_wait(&ctl->empty[s], ((kb / TC_STAGES) - 1) & 1);
      const int phase = kb < kb_gen ? 0 : 1;
      const int k0 = (phase == 0 ? kb : kb - kb_gen) * TC_BK;
      const int kmax = phase == 0 ? a.gen.ncols : a.kmem;
      const int wred0 = phase == 0 ? a.wred_gen : a.wred_mem;
      // A tile: 128 rows x 64 k
#pragma unroll 4
      for (int it = 0; it < 16; it++) {
        int idx = it * 128 + tid;
        int r = idx >> 4, k = (idx & 15) << 2;
        float4 v = load_a4(a, phase, m0 + r, M, k0 + k);
        sts_kmajor(sA[s], r, k, v);
      }
      if (WT) {
        // B tile K-major: rows n (Nc), 64 k ; W[(wout0+n)*ldw + wred0 + k]
        for (int idx = tid; idx < Nc * 16; idx += 128) {
          int r = idx >> 4, k = (idx & 15) << 2;
          int kk = k0 + k;
          bool rv = r < nvalid;
          const float* wp = W + (long long)(wout0 + n0 + r) * ldw + wred0 + kk;
          float4 v = ldw4(wp, rv && kk < kmax, rv && kk + 1 < kmax, rv && kk + 2 < kmax, rv && kk + 3 < kmax);
          sts_kmajor(sB[s], r, k, v);
        }
      } else {
        // B tile MN-major: 64 k-rows x Nc n ; W[(wred0+k)*ldw + wout0 + n]
        const int nq = Nc >> 2;
        for (int idx = tid; idx < 64 * nq; idx += 128) {
          int kr = idx / nq, n = (idx - kr * nq) << 2;
          int kk = k0 + kr;
          bool kv = kk < kmax;
          const float* wp = W + (long long)(wred0 + kk) * ldw + wout0 + n0 + n;
          float4 v = ldw4(wp, kv && n < nvalid, kv && n + 1 < nvalid, kv && n + 2 < nvalid, kv && n + 3 < nvalid);
          sts_mnmajor(sB[s], kr, n, v);
        }
      }
      fence_proxy_async();
      mbar_arrive(&ctl->full[s]);
    }
    // ------------------------------ epilogue ------------------------------
    mbar_wait(&ctl->accum, 0);
    tc_fence_after();
    const long long m = m0 + tid;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int c0 = 0; c0 < Nc; c0 += 32) {
      float v[32];
      tmem_ld32(tmem_d + lane_base + c0, v);
      if (m < M) {
#pragma unroll
        for (int j = 0; j < 32; j++) {
          int n = n0 + c0 + j;
          if (c0 + j < nvalid) epilogue_store(e, m, n, KB > 0 ? v[j] : 0.f);
        }
      }
    }
    tc_fence_before();
  } else if (lane == 0) {
    // ------------------------------ MMA issuer ------------------------------
    const uint32_t idesc = make_idesc(Nc, 0, WT ? 0 : 1);
    for (int kb = 0; kb < KB; kb++) {
      const int s = kb % TC_STAGES;
      mbar_wait(&ctl->full[s], (kb / TC_STAGES) & 1);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(sA[s]), b_addr = smem_u32(sB[s]);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        uint64_t ad = make_desc(a_addr + k * 32, 16, 1024);
        uint64_t bd = WT ? make_desc(b_addr + k * 32, 16, 1024) : make_desc(b_addr + k * 2048, 8192, 1024);
        umma_bf16(tmem_d, ad, bd, idesc, (kb > 0 || k > 0) ? 1 : 0);
      }
      umma_commit(&ctl->empty[s]);
    }
    umma_commit(&ctl->accum);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_d, 256);
  }
}

// ---------------------------------------------------------------------------------------------
// dW[(wout0+n)*ldw + wred + k] += sum_m dY[m][n] * A[m][k] ; db += sum_m dY[m][n]
// grid.x = n_tiles(128) * k_chunks(256 per phase), grid.y = splits over m
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 2)
tc_gemm_wgrad_kernel(const float* __restrict__ dY, int ldy, ASeg a, float* __restrict__ dW, int ldw, int wout0,
                     float* __restrict__ db, int M, int N, int m_per_split) {
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA[TC_STAGES];
  uint8_t* sB[TC_STAGES];
#pragma unroll
  for (int s = 0; s < TC_STAGES; s++) {
    sA[s] = base + s * (TC_A_BYTES + TC_B_BYTES);
    sB[s] = sA[s] + TC_A_BYTES;
  }
  TcSmem* ctl = reinterpret_cast<TcSmem*>(base + TC_STAGES * (TC_A_BYTES + TC_B_BYTES));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const int kc_gen = (a.gen.ncols + 255) / 256, kc_mem = (a.kmem + 255) / 256;
  const int kchunks = kc_gen + kc_mem;
  const int nt = blockIdx.x / kchunks, kc = blockIdx.x % kchunks;
  const int n0 = nt * 128;
  const int phase = kc < kc_gen ? 0 : 1;
  const int k0 = (phase == 0 ? kc : kc - kc_gen) * 256;
  const int kmax = phase == 0 ? a.gen.ncols : a.kmem;
  const int wred0 = phase == 0 ? a.wred_gen : a.wred_mem;
  const int kvalid = min(256, kmax - k0);
  const int Nc = (kvalid + 15) & ~15;
  const int nvalid = min(128, N - n0);
  const long long mbeg = (long long)blockIdx.y * m_per_split;
  long long mend = mbeg + m_per_split;
  if (mend > M) mend = M;
  const int KB = mbeg < mend ? (int)((mend - mbeg + TC_BK - 1) / TC_BK) : 0;
  const bool do_bias = db != nullptr && kc == 0;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < TC_STAGES; s++) { mbar_init(&ctl->full[s], 128); mbar_init(&ctl->empty[s], 1); }
    mbar_init(&ctl->accum, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) tmem_alloc(&ctl->tmem_base, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = ctl->tmem_base;

  if (warp < 4) {
    float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
    const int bn = (tid & 31) << 2;            // this thread always loads dY columns n0+bn..+3
    for (int kb = 0; kb < KB; kb++) {
      const int s = kb % TC_STAGES;
      if (kb >= TC_STAGES) mbar_wait(&ctl->empty[s], ((kb / TC_STAGES) - 1) & 1);
      const long long mb = mbeg + (long long)kb * TC_BK;
      // A operand (dY^T), MN-major: 64 m-rows x 128 n
#pragma unroll 4
      for (int it = 0; it < 16; it++) {
        int mr = it * 4 + (tid >> 5);
        long long m = mb + mr;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < mend && bn < nvalid) {
          v = __ldg(reinterpret_cast<const float4*>(dY + m * ldy + n0 + bn));
          if (bn + 1 >= nvalid) v.y = 0.f;
          if (bn + 2 >= nvalid) v.z = 0.f;
          if (bn + 3 >= nvalid) v.w = 0.f;
        }
        if (do_bias) { bsum.x += v.x; bsum.y += v.y; bsum.z += v.z; bsum.w += v.w; }
        sts_mnmajor(sA[s], mr, bn, v);
      }
      // B operand (A rows), MN-major: 64 m-rows x Nc k
      const int nq = Nc >> 2;
      for (int idx = tid; idx < 64 * nq; idx += 128) {
        int mr = idx / nq, k = (idx - mr * nq) << 2;
        long long m = mb + mr;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < mend) v = load_a4(a, phase, m, M, k0 + k);
        sts_mnmajor(sB[s], mr, k, v);
      }
      fence_proxy_async();
      mbar_arrive(&ctl->full[s]);
    }
    mbar_wait(&ctl->accum, 0);
    tc_fence_after();
    const int n = n0 + tid;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    if (KB > 0) {
      for (int c0 = 0; c0 < Nc; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_d + lane_base + c0, v);
        if (tid < nvalid) {
          float* dst = dW + (long long)(wout0 + n) * ldw + wred0 + k0 + c0;
#pragma unroll
          for (int j = 0; j < 32; j++)
            if (c0 + j < kvalid) atomicAdd(dst + j, v[j]);
        }
      }
      if (do_bias) {
        if (bn + 0 < nvalid) atomicAdd(db + wout0 + n0 + bn + 0, bsum.x);
        if (bn + 1 < nvalid) atomicAdd(db + wout0 + n0 + bn + 1, bsum.y);
        if (bn + 2 < nvalid) atomicAdd(db + wout0 + n0 + bn + 2, bsum.z);
        if (bn + 3 < nvalid) atomicAdd(db + wout0 + n0 + bn + 3, bsum.w);
      }
    }
    tc_fence_before();
  } else if (lane == 0) {
    const uint32_t idesc = make_idesc(Nc, 1, 1);
    for (int kb = 0; kb < KB; kb++) {
      const int s = kb % TC_STAGES;
      mbar_wait(&ctl->full[s], (kb / TC_STAGES) & 1);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(sA[s]), b_addr = smem_u32(sB[s]);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        uint64_t ad = make_desc(a_addr + k * 2048, 8192, 1024);
        uint64_t bd = make_desc(b_addr + k * 2048, 8192, 1024);
        umma_bf16(tmem_d, ad, bd, idesc, (kb > 0 || k > 0) ? 1 : 0);
      }
      umma_commit(&ctl->empty[s]);
    }
    umma_commit(&ctl->accum);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_d, 256);
  }
}

constexpr int TC_SMEM_BYTES = TC_STAGES * (TC_A_BYTES + TC_B_BYTES) + 1024 + 256;

inline int tc_prepare() {
  static int done = 0;
  if (done) return 0;
  cudaError_t e1 = cudaFuncSetAttribute(tc_gemm_mk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  cudaError_t e2 = cudaFuncSetAttribute(tc_gemm_mk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  cudaError_t e3 = cudaFuncSetAttribute(tc_gemm_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) return 1;
  done = 1;
  return 0;
}

inline void launch_tc_fwd(const ASeg& a, const float* W, int ldw, int wout0, long long M, int N, const Epi& e,
                          cudaStream_t st) {
  dim3 grid(cdiv(M, TC_BM), cdiv(N, 256));
  prof_begin(PC_TC_MLP, 2.0 * (double)M * N * (a.gen.ncols + a.kmem), 0.0, st);
  tc_gemm_mk_kernel<true><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(a, W, ldw, wout0, (int)M, N, e);
  prof_end(st);
}
inline void launch_tc_bwd_data(const ASeg& a, const float* W, int ldw, int wout0, long long M, int N, const Epi& e,
                               cudaStream_t st) {
  dim3 grid(cdiv(M, TC_BM), cdiv(N, 256));
  prof_begin(PC_TC_MLP, 2.0 * (double)M * N * (a.gen.ncols + a.kmem), 0.0, st);
  tc_gemm_mk_kernel<false><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(a, W, ldw, wout0, (int)M, N, e);
  prof_end(st);
}
inline void launch_tc_wgrad(const float* dY, int ldy, const ASeg& a, float* dW, int ldw, int wout0, float* db,
                            long long M, int N, int num_sms, cudaStream_t st) {
  int kchunks = cdiv(a.gen.ncols, 256) + cdiv(a.kmem, 256);
  if (kchunks == 0) return;
  int tiles = cdiv(N, 128) * kchunks;
  int splits = (2 * num_sms + tiles - 1) / tiles;
  int max_splits = cdiv(M, 4 * TC_BK);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int mps = round_up(cdiv(M, splits), TC_BK);
  splits = cdiv(M, mps);
  dim3 grid(tiles, splits);
  prof_begin(PC_TC_MLP, 2.0 * (double)M * N * (a.gen.ncols + a.kmem), 0.0, st);
  tc_gemm_wgrad_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(dY, ldy, a, dW, ldw, wout0, db, (int)M, N, mps);
  prof_end(st);
}

// ---------------------------------------------------------------------------------------------
// precision dispatch: 0 = FP32 SIMT (exactness anchor), 1 = BF16 tcgen05 (FP32 accumulate)
// ---------------------------------------------------------------------------------------------
inline int& precision_mode() { static int m = 0; return m; }

inline void launch_gemm_fwd(const ASeg& a, const float* W, int ldw, int wout0, long long M, int N, const Epi& e,
                            cudaStream_t st) {
  if (M <= 0 || N <= 0) return;
  if (precision_mode() == 1 && tc_prepare() == 0) launch_tc_fwd(a, W, ldw, wout0, M, N, e, st);
  else launch_simt_fwd(a, W, ldw, wout0, M, N, e, st);
}
inline void launch_gemm_bwd_data(const ASeg& a, const float* W, int ldw, int wout0, long long M, int N,
                                 const Epi& e, cudaStream_t st) {
  if (M <= 0 || N <= 0) return;
  if (precision_mode() == 1 && tc_prepare() == 0) launch_tc_bwd_data(a, W, ldw, wout0, M, N, e, st);
  else launch_simt_bwd_data(a, W, ldw, wout0, M, N, e, st);
}
inline void launch_gemm_wgrad(const float* dY, int ldy, const ASeg& a, float* dW, int ldw, int wout0, float* db,
                              long long M, int N, int num_sms, cudaStream_t st) {
  if (M <= 0 || N <= 0) return;
  if (precision_mode() == 1 && tc_prepare() == 0) launch_tc_wgrad(dY, ldy, a, dW, ldw, wout0, db, M, N, num_sms, st);
  else launch_simt_wgrad(dY, ldy, a, dW, ldw, wout0, db, M, N, num_sms, st);
}

}  // namespace fneus

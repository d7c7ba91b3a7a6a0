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
#include <stdlib.h>

#include "gemm_simt.cuh"

namespace fneus {

constexpr int TC_BM = 128, TC_BK = 64, TC_STAGES = 2, TC_THREADS = 192;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;          // 16 KB
constexpr int TC_B_BYTES = 256 * TC_BK * 2;            // 32 KB (max)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
// Every mbarrier wait is bounded: a wait that has not completed after FNEUS_WAIT_LIMIT_NS (far beyond the run time of any
// kernel of this library) records who waits on what in a host-mapped record (fneus_debug_hang_record) and traps -- a
// protocol error surfaces as a CUDA launch failure with a diagnosis instead of a hung GPU.
#ifndef FNEUS_WAIT_LIMIT_NS
#define FNEUS_WAIT_LIMIT_NS 4000000000ull
#endif
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ unsigned long long* g_hang_rec = nullptr;     // 64 words of host-mapped memory (tc_prepare), or null
__device__ __noinline__ void mbar_hang(uint32_t bar_addr, uint32_t parity) {
  volatile unsigned long long* r = g_hang_rec;
  if (r != nullptr && atomicCAS(g_hang_rec, 0ull, 1ull) == 0ull) {
    r[1] = ((unsigned long long)blockIdx.x << 32) | threadIdx.x;
    r[2] = ((unsigned long long)gridDim.x << 32) | blockDim.x;
    r[3] = ((unsigned long long)bar_addr << 32) | parity;
    r[4] = ((unsigned long long)blockIdx.y << 32) | blockIdx.z;
    // the 64-bit state words of the shared-memory neighbourhood [bar - 128, bar + 256): the control block of the kernel
    // (which barriers are incomplete, and by how many arrivals)
    for (int i = 0; i < 48; i++) {
      unsigned long long w;
      const uint32_t a = bar_addr - 128u + 8u * i;
      asm volatile("ld.shared.b64 %0, [%1];" : "=l"(w) : "r"(a));
      r[8 + i] = w;
    }
    __threadfence_system();
    r[5] = 0x600DD06ull;                                  // record complete
    __threadfence_system();
    const unsigned long long t0 = global_ns();
    while (global_ns() - t0 < 2000000ull) { }             // let the writes reach the host before the context dies
    __trap();
  }
  // every other stuck thread leaves the trap to the recorder (a trap kills the kernel at once, record or not)
  const unsigned long long t1 = global_ns();
  while (global_ns() - t1 < 50000000ull) { }
  __trap();
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar_addr, uint32_t parity, uint32_t hint) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar_addr), "r"(parity), "r"(hint) : "memory");
  return ok != 0;
}
// suspend-time hint: the hardware parks the warp until the phase flips instead of re-polling (spin loops were 45% of the
// issued instructions)
__device__ __forceinline__ void mbar_wait_hint(uint64_t* bar, uint32_t parity, uint32_t hint) {
  const uint32_t a = smem_u32(bar);
  if (mbar_try_wait(a, parity, hint)) return;
  unsigned long long t0 = 0;
  uint32_t spins = 0;
  while (!mbar_try_wait(a, parity, hint)) {
    if ((++spins & 1023u) == 0u) {
      const unsigned long long t = global_ns();
      if (t0 == 0) t0 = t;
      else if (t - t0 > FNEUS_WAIT_LIMIT_NS) mbar_hang(a, parity);
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { mbar_wait_hint(bar, parity, 0x989680u); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
// 1-D bulk async copy global -> shared (TMA engine, UBLKCP), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
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
// instruction descriptor: {BF16 | FP16} x {BF16 | FP16} -> FP32, M = 128 (kind::f16 takes the format per operand)
__device__ __forceinline__ uint32_t make_idesc(int n, int a_mn_major, int b_mn_major, bool a_f16 = false,
                                               bool b_f16 = false) {
  uint32_t d = 0;
  d |= 1u << 4;                       // c_format = F32
  d |= (a_f16 ? 0u : 1u) << 7;        // a_format: 0 = F16, 1 = BF16
  d |= (b_f16 ? 0u : 1u) << 10;       // b_format
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
__device__ __forceinline__ uint2 pack_f16x4(float4 v) { return make_uint2(pack_f16x2(v.x, v.y), pack_f16x2(v.z, v.w)); }
// 8 consecutive 16-bit elements (one 16-byte chunk) <-> FP32
__device__ __forceinline__ void u16x8_to_f32(const uint4 u, bool f16, float* out) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
  if (f16) {
#pragma unroll
    for (int t = 0; t < 4; t++) { const float2 f = unpack_f16x2(w[t]); out[2 * t] = f.x; out[2 * t + 1] = f.y; }
  } else {
#pragma unroll
    for (int t = 0; t < 4; t++) { out[2 * t] = __uint_as_float(w[t] << 16); out[2 * t + 1] = __uint_as_float(w[t] & 0xFFFF0000u); }
  }
}
__device__ __forceinline__ uint4 f32x8_to_u16(const float* y, bool f16) {
  if (f16) return make_uint4(pack_f16x2(y[0], y[1]), pack_f16x2(y[2], y[3]), pack_f16x2(y[4], y[5]), pack_f16x2(y[6], y[7]));
  const uint2 lo = pack_bf16x4(make_float4(y[0], y[1], y[2], y[3]));
  const uint2 hi = pack_bf16x4(make_float4(y[4], y[5], y[6], y[7]));
  return make_uint4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ unsigned short f32_to_u16_bits(float v, bool f16) { return f16 ? f32_to_f16_bits(v) : f32_to_bf16_bits(v); }
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

// Warp-local transpose of a 32(rows) x 32(cols) accumulator chunk through shared memory so that the fused epilogue
// touches global memory with lanes along the contiguous (column) axis.
__device__ __forceinline__ float lds_f32(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t saddr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory");
}

// Fused epilogue of one 32-row x 32-column accumulator chunk held (transposed) in this warp's shared-memory
// buffer: lane = column, so every global access is a coalesced 128-byte row segment.  MODE is a template
// parameter (one specialised loop per epilogue kind), per-column constants (bias, rvec, split side, pointers)
// are hoisted, auxiliary operands of 8 rows are fetched before use.  MUFU-based softplus/sigmoid.
template <int MODE>
__device__ __forceinline__ void epilogue_rows(const Epi& e, uint32_t wb, long long mrow0, int rows, int n, int lane) {
  const bool prim = n < e.csplit;
  const float beta = e.beta, inv_beta = 1.f / e.beta, oscale = e.oscale, hscale = e.hscale;
  float bias_n = 0.f, rvec_n = 0.f;
  if (MODE == EPI_LINEAR || MODE == EPI_RELU || MODE == EPI_SIGMOID || MODE == EPI_SOFTPLUS ||
      MODE == EPI_SOFTPLUS_Q || MODE == EPI_SDF_OUT)
    bias_n = e.bias ? __ldg(e.bias + n) : 0.f;
  if (MODE == EPI_SOFTPLUS_Q) rvec_n = __ldg(e.rvec + n);
  const bool rank1 = (MODE == EPI_SDF_BWD || MODE == EPI_RELUMASK) && e.rs != nullptr && prim;
  if (rank1) rvec_n = __ldg(e.rvec + n) * e.rscale;

  // primary / secondary destinations and auxiliary sources for THIS column
  float* dst = nullptr; long long ldd = 0;
  const float* hsrc = nullptr; long long ldh = 0;
  float* qptr = nullptr; long long ldq = 0;       // read (and for SWEEP written) auxiliary
  bool q_is_acc = false;                           // RELUMASK secondary accumulate: q aliases the destination
  if (MODE == EPI_SDF_OUT) {
    if (n == 0) { dst = e.out0 + mrow0; ldd = 1; }
    else if (e.C) { dst = e.C + mrow0 * e.ldc + (n - 1); ldd = e.ldc; }
  } else if (prim) {
    if (e.C) { dst = e.C + mrow0 * e.ldc + n; ldd = e.ldc; }
    if (MODE == EPI_SPMUL || MODE == EPI_SWEEP || MODE == EPI_SDF_BWD || (MODE == EPI_RELUMASK && e.H && e.C)) {
      hsrc = e.H + mrow0 * e.ldh + n; ldh = e.ldh;
    }
    if ((MODE == EPI_SWEEP) || ((MODE == EPI_SDF_BWD || MODE == EPI_LINEAR_ADD) && e.Q)) {
      qptr = e.Q + mrow0 * e.ldq + n; ldq = e.ldq;
    }
    if (MODE == EPI_SOFTPLUS_Q) { qptr = e.Q + mrow0 * e.ldq + n; ldq = e.ldq; }
  } else {
    const int c = n - e.csplit;
    if ((MODE == EPI_SPMUL || MODE == EPI_RELUMASK) && e.C2) { dst = e.C2 + mrow0 * e.ldc2 + c; ldd = e.ldc2; }
    if (MODE == EPI_RELUMASK && e.C2) {
      if (e.H2) { hsrc = e.H2 + mrow0 * e.ldh2 + c; ldh = e.ldh2; }
      if (e.accumulate2) { qptr = dst; ldq = ldd; q_is_acc = true; }
    }
  }
  const bool need_q_load = (MODE == EPI_SWEEP || MODE == EPI_SDF_BWD || MODE == EPI_LINEAR_ADD || q_is_acc) && qptr;
  const float* rs = rank1 ? e.rs + mrow0 : nullptr;

#pragma unroll 1
  for (int r0 = 0; r0 < rows; r0 += 8) {
    float acc[8], h[8], q[8], rsv[8];
#pragma unroll
    for (int r = 0; r < 8; r++) {
      const bool ok = r0 + r < rows;
      acc[r] = lds_f32(wb + (uint32_t)((r0 + r) * 33 + lane) * 4u);
      h[r] = (hsrc && ok) ? __ldg(hsrc + (long long)(r0 + r) * ldh) : 0.f;
      q[r] = (need_q_load && ok) ? qptr[(long long)(r0 + r) * ldq] : 0.f;
      rsv[r] = (rs && ok) ? __ldg(rs + r0 + r) : 0.f;
    }
    // branch-free arithmetic for all 8 rows first (independent chains interleave), predicated stores after
    float y[8], y2[8];
#pragma unroll
    for (int r = 0; r < 8; r++) {
      float a = acc[r];
      y2[r] = 0.f;
      if (MODE == EPI_LINEAR) y[r] = (a + bias_n) * oscale;
      else if (MODE == EPI_RELU) y[r] = fmaxf(a + bias_n, 0.f);
      else if (MODE == EPI_SIGMOID) y[r] = sigmoid_fast(a + bias_n);
      else if (MODE == EPI_SOFTPLUS) y[r] = softplus_beta_fast(a + bias_n, beta, inv_beta) * oscale;
      else if (MODE == EPI_SOFTPLUS_Q) {
        float v = a + bias_n;
        y[r] = softplus_beta_fast(v, beta, inv_beta) * oscale;
        y2[r] = softplus_grad_from_pre_fast(v, beta) * rvec_n;
      } else if (MODE == EPI_SDF_OUT) y[r] = (n == 0) ? (a + bias_n) * e.out0_scale : (a + bias_n);
      else if (MODE == EPI_SPMUL) y[r] = prim ? softplus_grad_from_act_fast(h[r] * hscale, beta) * a * oscale : a * oscale;
      else if (MODE == EPI_SWEEP) {
        float sg = softplus_grad_from_act_fast(h[r] * hscale, beta);
        y[r] = sg * a * oscale;
        y2[r] = beta * (1.f - sg) * q[r] * a;
      } else if (MODE == EPI_SDF_BWD) {
        float sg = softplus_grad_from_act_fast(h[r] * hscale, beta);
        y[r] = sg * (a + rsv[r] * rvec_n) * oscale + q[r];
      } else if (MODE == EPI_RELUMASK) {
        float v = a + rsv[r] * rvec_n;
        if (hsrc) v = h[r] > 0.f ? v : 0.f;
        y[r] = v + q[r];
      } else y[r] = (a + q[r]) * oscale;   // EPI_LINEAR_ADD
    }
    const bool st_ok = dst != nullptr && !(MODE == EPI_SDF_BWD && !prim);
    const bool st2_ok = (MODE == EPI_SOFTPLUS_Q || MODE == EPI_SWEEP) && qptr != nullptr;
#pragma unroll
    for (int r = 0; r < 8; r++) {
      const long long ro = r0 + r;
      if (st_ok && r0 + r < rows) dst[ro * ldd] = y[r];
      if (st2_ok && r0 + r < rows) qptr[ro * ldq] = y2[r];
    }
  }
}

__device__ __forceinline__ void epilogue_chunk(const Epi& e, float* wbuf, const float* v, long long mrow0, int M,
                                               int ncol0, int nvalid_cols, int lane) {
  const uint32_t wb = smem_u32(wbuf);
#pragma unroll
  for (int j = 0; j < 32; j++) sts_f32(wb + (uint32_t)(lane * 33 + j) * 4u, v[j]);
  __syncwarp();
  long long left = (long long)M - mrow0;
  const int rows = left >= 32 ? 32 : (left > 0 ? (int)left : 0);
  if (lane < nvalid_cols && rows > 0 && !(e.dbg & 8)) {
    const int n = ncol0 + lane;
    switch (e.mode) {
      case EPI_LINEAR: epilogue_rows<EPI_LINEAR>(e, wb, mrow0, rows, n, lane); break;
      case EPI_RELU: epilogue_rows<EPI_RELU>(e, wb, mrow0, rows, n, lane); break;
      case EPI_SIGMOID: epilogue_rows<EPI_SIGMOID>(e, wb, mrow0, rows, n, lane); break;
      case EPI_SOFTPLUS: epilogue_rows<EPI_SOFTPLUS>(e, wb, mrow0, rows, n, lane); break;
      case EPI_SOFTPLUS_Q: epilogue_rows<EPI_SOFTPLUS_Q>(e, wb, mrow0, rows, n, lane); break;
      case EPI_SDF_OUT: epilogue_rows<EPI_SDF_OUT>(e, wb, mrow0, rows, n, lane); break;
      case EPI_SPMUL: epilogue_rows<EPI_SPMUL>(e, wb, mrow0, rows, n, lane); break;
      case EPI_SWEEP: epilogue_rows<EPI_SWEEP>(e, wb, mrow0, rows, n, lane); break;
      case EPI_SDF_BWD: epilogue_rows<EPI_SDF_BWD>(e, wb, mrow0, rows, n, lane); break;
      case EPI_RELUMASK: epilogue_rows<EPI_RELUMASK>(e, wb, mrow0, rows, n, lane); break;
      case EPI_LINEAR_ADD: epilogue_rows<EPI_LINEAR_ADD>(e, wb, mrow0, rows, n, lane); break;
    }
  }
  __syncwarp();
}

// The weight-gradient kernel runs ONE CTA per SM with a 4-deep ring: per stage the chain is bulk load (~2 us under load) ->
// format conversion (mixed FP16 / BF16 operands) -> 4 MMAs, and with only two stages per CTA that latency chain, not
// HBM, set the pace (measured +33% when the conversion joined it).
constexpr int WG_STAGES = 4;
struct TcSmem {
  uint64_t full[WG_STAGES];
  uint64_t empty[WG_STAGES];
  uint64_t conv[WG_STAGES];           // weight-gradient kernel: stage converted to a common operand format
  uint64_t fullc[WG_STAGES];          // ... and the landing of the operand that needs the conversion (it is fetched first,
                                      // so that it is converted while the other operand is still in flight)
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
tc_gemm_mk_kernel(ASeg a, const float* __restrict__ W, int ldw, int wout0, int M, int N, Epi e,
                  const uint8_t* __restrict__ wimg) {
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* base = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared
                                                 // address space (an integer round trip turned every access into a generic LD.E / ST.E)
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
    for (int s = 0; s < TC_STAGES; s++) { mbar_init(&ctl->full[s], wimg ? 129 : 128); mbar_init(&ctl->empty[s], 1); }
    mbar_init(&ctl->accum, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) tmem_alloc(&ctl->tmem_base, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = ctl->tmem_base;

  if (warp == 5) {
    // ------------------------------ weight-image loader (bulk async copies) ------------------------------
    if (wimg && lane == 0) {
      const uint32_t bytes = WT ? (uint32_t)Nc * 128u : (uint32_t)((Nc + 63) / 64) * 8192u;
      const uint8_t* src = wimg + (size_t)blockIdx.y * KB * TC_B_BYTES;
      for (int kb = 0; kb < KB; kb++) {
        const int s = kb % TC_STAGES;
        if (kb >= TC_STAGES) mbar_wait(&ctl->empty[s], ((kb / TC_STAGES) - 1) & 1);
        mbar_arrive_expect_tx(&ctl->full[s], bytes);
        bulk_g2s(sB[s], src + (size_t)kb * TC_B_BYTES, bytes, &ctl->full[s]);
      }
    }
  } else if (warp < 4) {
    // ------------------------------ producers ------------------------------
    for (int kb = 0; kb < KB; kb++) {
      const int s = kb % TC_STAGES;
      if (kb >= TC_STAGES) mbar_wait(&ctl->empty[s], ((kb / TC_STAGES) - 1) & 1);
      const int phase = kb < kb_gen ? 0 : 1;
      const int k0 = (phase == 0 ? kb : kb - kb_gen) * TC_BK;
      const int kmax = phase == 0 ? a.gen.ncols : a.kmem;
      const int wred0 = phase == 0 ? a.wred_gen : a.wred_mem;
      // A tile: 128 rows x 64 k
      if (phase == 0) {
#pragma unroll 1
        for (int it = 0; it < 16; it++) {
          int idx = it * 128 + tid;
          sts_kmajor(sA[s], idx >> 4, (idx & 15) << 2, load_a4(a, 0, m0 + (idx >> 4), M, k0 + ((idx & 15) << 2)));
        }
      } else {
        float4 av[16];
#pragma unroll
        for (int it = 0; it < 16; it++) {
          int idx = it * 128 + tid;
          av[it] = load_a4(a, 1, m0 + (idx >> 4), M, k0 + ((idx & 15) << 2));
        }
#pragma unroll
        for (int it = 0; it < 16; it++) {
          int idx = it * 128 + tid;
          sts_kmajor(sA[s], idx >> 4, (idx & 15) << 2, av[it]);
        }
      }
      if (wimg) {
        // B tile arrives by bulk copy (warp 5)
      } else if (WT) {
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
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    float* wbuf = reinterpret_cast<float*>(sB[0]) + warp * (32 * 33);     // stage buffers are free now
#pragma unroll 1
    for (int c0 = 0; c0 < Nc; c0 += 32) {
      float v[32];
      tmem_ld32(tmem_d + lane_base + c0, v);
      if (KB == 0) {
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = 0.f;
      }
      epilogue_chunk(e, wbuf, v, m0 + warp * 32, M, n0 + c0, nvalid - c0, lane);
    }
    tc_fence_before();
  } else if (warp == 4 && lane == 0) {
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

// sum over the 64 reduction rows of column n (0..127) of a BF16 MN-major dY^T tile in shared memory
__device__ __forceinline__ float wgrad_col_sum(const uint8_t* sA, int n, bool f16) {
  const uint8_t* tile = sA + (n >> 6) * 8192 + ((n & 7) << 1);
  const int ch = (n & 63) >> 3;
  float acc = 0.f;
#pragma unroll 8
  for (int kk = 0; kk < 64; kk++) {
    unsigned short b16 = *reinterpret_cast<const unsigned short*>(tile + (kk >> 3) * 1024 + (kk & 7) * 128 +
                                                                 (((ch ^ (kk & 7)) & 7) << 4));
    acc += f16 ? unpack_f16x2((uint32_t)b16).x : __uint_as_float((unsigned)b16 << 16);
  }
  return acc;
}

// FP16 -> BF16 in place over `bytes` of a landed stage (threads 0..127, 16-byte chunks: the layout is unchanged)
__device__ __forceinline__ void wgrad_to_bf16(uint8_t* tile, int bytes, int tid) {
#pragma unroll 4
  for (int i = tid * 16; i < bytes; i += 128 * 16) {
    float v[8];
    u16x8_to_f32(*reinterpret_cast<const uint4*>(tile + i), true, v);
    *reinterpret_cast<uint4*>(tile + i) = f32x8_to_u16(v, false);
  }
}

// Warps 0-3, one stage: convert the FP16 image operand once IT has landed (`fullc`; the other operand may still be in
// flight), then the bias column sums off the dY tile (an image: BF16 after the conversion, or never FP16 to begin with).
__device__ __forceinline__ void wgrad_convert_and_bias(TcSmem* ctl, uint8_t* sAs, uint8_t* sBs, int s, uint32_t parity,
                                                       bool conv_y, bool conv_x, int yblocks, int xblocks,
                                                       bool bias_smem, bool yf, int tid, float& bcol) {
  const bool any_conv = conv_y || conv_x;
  if (any_conv) {
    mbar_wait(&ctl->fullc[s], parity);
    wgrad_to_bf16(conv_y ? sAs : sBs, (conv_y ? yblocks : xblocks) * 8192, tid);
    fence_proxy_async();
    mbar_arrive(&ctl->conv[s]);
  }
  if (bias_smem) {
    // the dY tile must be complete (and, if it was converted, by ALL threads): conv when it was the converted operand
    mbar_wait(conv_y ? &ctl->conv[s] : &ctl->full[s], parity);
    bcol += wgrad_col_sum(sAs, tid, yf && !conv_y);
    mbar_arrive(&ctl->empty[s]);
  }
}

// ---------------------------------------------------------------------------------------------
// dW[(wout0+n)*ldw + wred + k] += sum_m dY[m][n] * A[m][k] ; db += sum_m dY[m][n]
// grid.x = n_tiles(128) * k_chunks(256 per phase), grid.y = splits over m (multiples of 64 rows).
// Operands that are BF16 activation images (negative leading dimension) arrive by bulk async copies issued by
// warp 5 (their tile bytes ARE the MN-major shared-memory operand); FP32 row-major / generated operands are
// converted by the four producer warps.
// ---------------------------------------------------------------------------------------------
// y_f16 / x_f16: element format of an IMAGE operand (forward-path images are FP16, backward-path ones BF16); operands
// converted from FP32 by the producer warps are always BF16.  tcgen05.mma kind::f16 wants ONE format for both operands
// (a mixed descriptor raises an illegal-instruction fault on sm_100a), so when exactly one operand is an FP16 image the
// otherwise idle warps 0-3 rewrite that stage FP16 -> BF16 in place (exact for |v| >= 2^-14 up to BF16's 8 bits; the
// gradient operand keeps its BF16 exponent range) and hand the stage on through `conv`.
template <int WG_ST>
__device__ __forceinline__ void wgrad_body(const float* __restrict__ dY, int ldy, const ASeg& a, float* __restrict__ dW,
                                           int ldw, int wout0, float* __restrict__ db, int M, int N, int m_per_split,
                                           const int bx, const int by, const bool y_f16 = false,
                                           const bool x_f16 = false, const bool no_conv = false) {
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* base = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared
                                                 // address space (an integer round trip turned every access into a generic LD.E / ST.E)
  uint8_t* sA[WG_ST];
  uint8_t* sB[WG_ST];
#pragma unroll
  for (int s = 0; s < WG_ST; s++) {
    sA[s] = base + s * (TC_A_BYTES + TC_B_BYTES);
    sB[s] = sA[s] + TC_A_BYTES;
  }
  TcSmem* ctl = reinterpret_cast<TcSmem*>(base + WG_ST * (TC_A_BYTES + TC_B_BYTES));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const int kc_gen = (a.gen.ncols + 255) / 256, kc_mem = (a.kmem + 255) / 256;
  const int kchunks = kc_gen + kc_mem;
  const int nt = bx / kchunks, kc = bx % kchunks;
  const int n0 = nt * 128;
  const int phase = kc < kc_gen ? 0 : 1;
  const int k0 = (phase == 0 ? kc : kc - kc_gen) * 256;
  const int kmax = phase == 0 ? a.gen.ncols : a.kmem;
  const int wred0 = phase == 0 ? a.wred_gen : a.wred_mem;
  const int kvalid = min(256, kmax - k0);
  const int Nc = (kvalid + 15) & ~15;
  const int nvalid = min(128, N - n0);
  const long long mbeg = (long long)by * m_per_split;
  long long mend = mbeg + m_per_split;
  if (mend > M) mend = M;
  const int KB = mbeg < mend ? (int)((mend - mbeg + TC_BK - 1) / TC_BK) : 0;
  const bool y_img = ldy < 0;
  const bool x_img = phase == 1 && a.ldm < 0;
  const bool need_prod = !(y_img && x_img);
  const bool do_bias = db != nullptr && kc == 0 && !y_img;
  // image dY: the bias gradient is summed from the shared-memory tile by warps 0-3 (they hold the stage open)
  const bool bias_smem = db != nullptr && kc == 0 && y_img;
  const bool yf = y_img && y_f16, xf = x_img && x_f16;
  const bool conv_y = yf && !xf && !no_conv, conv_x = xf && !yf && !no_conv, any_conv = conv_y || conv_x;
  const int yblocks = y_img ? min(2, -ldy - nt * 2) : 0;
  const int xblocks = x_img ? min((Nc + 63) / 64, -a.ldm - (k0 >> 6)) : 0;

  // `full` covers the produced operand and the image operand that needs no conversion; `fullc` the image operand that does
  const bool ld_y_full = y_img && !conv_y, ld_x_full = x_img && !conv_x;
  const bool use_full = need_prod || ld_y_full || ld_x_full;
  if (tid == 0) {
    const int cnt = (need_prod ? 128 : 0) + ((ld_y_full || ld_x_full) ? 1 : 0);
#pragma unroll
    for (int s = 0; s < WG_ST; s++) {
      mbar_init(&ctl->full[s], cnt > 0 ? cnt : 1); mbar_init(&ctl->empty[s], bias_smem ? 129 : 1);
      mbar_init(&ctl->conv[s], 128); mbar_init(&ctl->fullc[s], 1);
    }
    mbar_init(&ctl->accum, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) tmem_alloc(&ctl->tmem_base, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = ctl->tmem_base;

  if (warp == 5) {
    if (lane == 0 && (y_img || x_img)) {
      const int ykbs = -ldy, xkbs = -a.ldm;
      for (int kb = 0; kb < KB; kb++) {
        const int s = kb % WG_ST;
        if (kb >= WG_ST) mbar_wait(&ctl->empty[s], ((kb / WG_ST) - 1) & 1);
        const long long mblk = (mbeg >> 6) + kb;               // 64-row block index
        const size_t half = (size_t)(mblk & 1) * 8192;
        uint64_t* ybar = conv_y ? &ctl->fullc[s] : &ctl->full[s];
        uint64_t* xbar = conv_x ? &ctl->fullc[s] : &ctl->full[s];
        auto load_y = [&]() {
          for (int b = 0; b < yblocks; b++)
            bulk_g2s(sA[s] + b * 8192, reinterpret_cast<const uint8_t*>(dY) + ((size_t)(mblk >> 1) * ykbs + nt * 2 + b) * 16384 + half,
                     8192, ybar);
        };
        auto load_x = [&]() {
          for (int b = 0; b < xblocks; b++)
            bulk_g2s(sB[s] + b * 8192, reinterpret_cast<const uint8_t*>(a.mem) + ((size_t)(mblk >> 1) * xkbs + (k0 >> 6) + b) * 16384 + half,
                     8192, xbar);
        };
        if (any_conv) {                                          // the operand to convert goes first, on its own barrier
          mbar_arrive_expect_tx(&ctl->fullc[s], (uint32_t)(conv_y ? yblocks : xblocks) * 8192u);
          if (conv_y) load_y(); else load_x();
          const int rest = conv_y ? xblocks : yblocks;
          if (rest > 0) {
            mbar_arrive_expect_tx(&ctl->full[s], (uint32_t)rest * 8192u);
            if (conv_y) load_x(); else load_y();
          }
        } else {
          mbar_arrive_expect_tx(&ctl->full[s], (uint32_t)(yblocks + xblocks) * 8192u);
          load_y();
          load_x();
        }
      }
    }
  } else if (warp < 4) {
    float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
    float bcol = 0.f;
    const int bn = (tid & 31) << 2;            // this thread always loads dY columns n0+bn..+3
    if (need_prod) {
      for (int kb = 0; kb < KB; kb++) {
        const int s = kb % WG_ST;
        if (kb >= WG_ST) mbar_wait(&ctl->empty[s], ((kb / WG_ST) - 1) & 1);
        const long long mb = mbeg + (long long)kb * TC_BK;
        if (!y_img) {
          // A operand (dY^T), MN-major: 64 m-rows x 128 n
          float4 yv[16];
#pragma unroll
          for (int it = 0; it < 16; it++) {
            long long m = mb + it * 4 + (tid >> 5);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < mend && bn < nvalid) v = __ldg(reinterpret_cast<const float4*>(dY + m * ldy + n0 + bn));
            yv[it] = v;
          }
#pragma unroll
          for (int it = 0; it < 16; it++) {
            float4 v = yv[it];
            if (bn + 1 >= nvalid) v.y = 0.f;
            if (bn + 2 >= nvalid) v.z = 0.f;
            if (bn + 3 >= nvalid) v.w = 0.f;
            if (do_bias) { bsum.x += v.x; bsum.y += v.y; bsum.z += v.z; bsum.w += v.w; }
            sts_mnmajor(sA[s], it * 4 + (tid >> 5), bn, v);
          }
        }
        // B operand (A rows), MN-major: 64 m-rows x Nc k ; thread -> column group (tid % 64)*4, rows (tid/64) + 2*it
        if (phase == 1 && !x_img) {
          const int kq = (tid & 63) << 2;
          if (kq < Nc) {
#pragma unroll
            for (int half = 0; half < 2; half++) {
              float4 xv[16];
#pragma unroll
              for (int it = 0; it < 16; it++) {
                long long m = mb + (tid >> 6) + 2 * (half * 16 + it);
                xv[it] = (m < mend) ? load_a4(a, 1, m, M, k0 + kq) : make_float4(0.f, 0.f, 0.f, 0.f);
              }
#pragma unroll
              for (int it = 0; it < 16; it++) sts_mnmajor(sB[s], (tid >> 6) + 2 * (half * 16 + it), kq, xv[it]);
            }
          }
        } else if (phase == 0) {
          // generated columns, MN-major: one reduction row per thread (64 rows), zero then fill
          if (tid < 64) {
            const int r7 = tid & 7;
            uint8_t* rowp = sB[s] + (tid >> 3) * 1024 + r7 * 128;
            const int nblk = (Nc + 63) >> 6;
            for (int b = 0; b < nblk; b++)
#pragma unroll
              for (int c = 0; c < 8; c++) *reinterpret_cast<uint4*>(rowp + b * 8192 + c * 16) = make_uint4(0u, 0u, 0u, 0u);
            const long long m = mb + tid;
            if (m < mend) {
              gen_row(a.gen, m, [&](int j, float val) {
                const int jj = j - k0;
                if (jj >= 0 && jj < Nc)
                  *reinterpret_cast<unsigned short*>(rowp + (jj >> 6) * 8192 + (((((jj & 63) >> 3) ^ r7) & 7) << 4) +
                                                     ((jj & 7) << 1)) = f32_to_bf16_bits(val);
              });
            }
          }
        }
        fence_proxy_async();
        mbar_arrive(&ctl->full[s]);
        wgrad_convert_and_bias(ctl, sA[s], sB[s], s, (kb / WG_ST) & 1, conv_y, conv_x, yblocks, xblocks, bias_smem, yf, tid, bcol);
      }
    } else if (any_conv || bias_smem) {
      for (int kb = 0; kb < KB; kb++) {
        const int s = kb % WG_ST;
        wgrad_convert_and_bias(ctl, sA[s], sB[s], s, (kb / WG_ST) & 1, conv_y, conv_x, yblocks, xblocks, bias_smem, yf, tid, bcol);
      }
    }
    mbar_wait(&ctl->accum, 0);
    tc_fence_after();
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    float* wbuf = reinterpret_cast<float*>(sB[0]) + warp * (32 * 33);
    if (bias_smem && KB > 0 && tid < nvalid) atomicAdd(db + wout0 + n0 + tid, bcol);
    if (KB > 0) {
      for (int c0 = 0; c0 < Nc; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_d + lane_base + c0, v);
#pragma unroll
        for (int j = 0; j < 32; j++) wbuf[lane * 33 + j] = v[j];
        __syncwarp();
        if (c0 + lane < kvalid) {
#pragma unroll 4
          for (int r = 0; r < 32; r++) {
            int nn = warp * 32 + r;
            if (nn < nvalid)
              atomicAdd(dW + (long long)(wout0 + n0 + nn) * ldw + wred0 + k0 + c0 + lane, wbuf[r * 33 + lane]);
          }
        }
        __syncwarp();
      }
      if (do_bias) {
        if (bn + 0 < nvalid) atomicAdd(db + wout0 + n0 + bn + 0, bsum.x);
        if (bn + 1 < nvalid) atomicAdd(db + wout0 + n0 + bn + 1, bsum.y);
        if (bn + 2 < nvalid) atomicAdd(db + wout0 + n0 + bn + 2, bsum.z);
        if (bn + 3 < nvalid) atomicAdd(db + wout0 + n0 + bn + 3, bsum.w);
      }
    }
    tc_fence_before();
  } else if (warp == 4 && lane == 0) {
    const uint32_t idesc = make_idesc(Nc, 1, 1, yf && !conv_y && !no_conv, xf && !conv_x && !no_conv);
    for (int kb = 0; kb < KB; kb++) {
      const int s = kb % WG_ST;
      if (use_full) mbar_wait(&ctl->full[s], (kb / WG_ST) & 1);
      if (any_conv) mbar_wait(&ctl->conv[s], (kb / WG_ST) & 1);
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

__global__ void __launch_bounds__(TC_THREADS, 2)
tc_gemm_wgrad_kernel(const float* __restrict__ dY, int ldy, const __grid_constant__ ASeg a, float* __restrict__ dW,
                     int ldw, int wout0, float* __restrict__ db, int M, int N, int m_per_split) {
  wgrad_body<TC_STAGES>(dY, ldy, a, dW, ldw, wout0, db, M, N, m_per_split, blockIdx.x, blockIdx.y);
}

// Grouped variant: the weight gradients of all layers of one chain in ONE launch (blockIdx.z = job).  Every job
// shares M; a job's unused (tile, split) slots of the common grid exit immediately.
constexpr int WG_MAX_JOBS = 20;
struct WgradJob {
  const float* dY; int ldy;
  ASeg a;
  float* dW; int ldw, wout0;
  float* db;
  int N, m_per_split, tiles, splits;
  int y_f16, x_f16;                  // element format of the image operands (0 = BF16, 1 = FP16)
};
struct WgradJobs { int n; int M; int no_conv; WgradJob job[WG_MAX_JOBS]; };
template <int WG_ST>
__global__ void __launch_bounds__(TC_THREADS, WG_ST == 2 ? 2 : 1) tc_gemm_wgrad_group_kernel(const __grid_constant__ WgradJobs jobs) {
  const WgradJob& j = jobs.job[blockIdx.z];
  if ((int)blockIdx.x >= j.tiles || (int)blockIdx.y >= j.splits) return;
  wgrad_body<WG_ST>(j.dY, j.ldy, j.a, j.dW, j.ldw, j.wout0, j.db, jobs.M, j.N, j.m_per_split, blockIdx.x, blockIdx.y,
                    j.y_f16 != 0, j.x_f16 != 0, jobs.no_conv != 0);
}

// ---------------------------------------------------------------------------------------------
// Persistent, warp-specialised variant of tc_gemm_mk (needs a weight image): one CTA per SM loops over
// output tiles; the epilogue of tile i (8 warps) overlaps the operand loads + MMAs of tile i+1 through two
// TMEM accumulator stages (2 x 256 columns).
//   warps 0-7  : A producers for generated / FP32 row-major operands, two groups of 4 on alternate reduction blocks
//   warp  8    : MMA issuer (+ TMEM alloc)
//   warp  9    : bulk-copy loader: weight-image tile and, when A is a BF16 activation image, the A tile too
//   warps 10-17: epilogue, one thread per accumulator row (TMEM lane), 16 columns at a time, vector I/O on
//                FP32 row-major tensors and on BF16 activation images; two groups on alternate 32-column chunks
// ---------------------------------------------------------------------------------------------
constexpr int P_STAGES = 4, P_THREADS = 576;
inline int& tc_debug_flags() { static int f = 0; return f; }   // bit0: no A loads, bit1: no epilogue, bit2: no MMA
struct PSmem {
  uint64_t full[P_STAGES];
  uint64_t empty[P_STAGES];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint32_t tmem_base;
};
constexpr int P_VEC_FLOATS = 2 * 1024;                 // staged bias / rvec (N <= 1024)
constexpr int P_SMEM_BYTES = P_STAGES * (TC_A_BYTES + TC_B_BYTES) + P_VEC_FLOATS * 4 + 1024 + 256;

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// 16 consecutive columns [c, c+16) (c % 16 == 0) of row m.  Image: two 16-byte chunks of BF16.  FP32 row-major:
// four float4 when aligned, else scalar.  Only columns j < nv are guaranteed meaningful for FP32 sources.
__device__ __forceinline__ void row_load16(const float* p, int ld, long long m, int c, int nv, float* out) {
  if (ld < 0) {
    const int kbs = -ld, r7 = (int)(m & 7);
    const uint8_t* base = reinterpret_cast<const uint8_t*>(p) + ((size_t)(m >> 7) * kbs + (c >> 6)) * 16384 +
                          (((m & 127) >> 3) * 1024 + r7 * 128);
    const int ch = (c & 63) >> 3;
#pragma unroll
    for (int i = 0; i < 2; i++) {
      uint4 u = __ldg(reinterpret_cast<const uint4*>(base + (((ch + i) ^ r7) << 4)));
      uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int t = 0; t < 4; t++) {
        out[i * 8 + 2 * t] = __uint_as_float(w[t] << 16);
        out[i * 8 + 2 * t + 1] = __uint_as_float(w[t] & 0xFFFF0000u);
      }
    }
  } else {
    const float* q = p + m * ld + c;
    if (nv >= 16 && (reinterpret_cast<uintptr_t>(q) & 15) == 0) {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        float4 f = __ldg(reinterpret_cast<const float4*>(q) + i);
        out[4 * i] = f.x; out[4 * i + 1] = f.y; out[4 * i + 2] = f.z; out[4 * i + 3] = f.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; j++) out[j] = j < nv ? __ldg(q + j) : 0.f;
    }
  }
}
// Store columns j < nv of the 16-column group; image destinations always get the full group (the caller zeroes the
// tail) so that the padding columns of an image stay finite.
__device__ __forceinline__ void row_store16(float* p, int ld, long long m, int c, int nv, const float* v) {
  if (ld < 0) {
    const int kbs = -ld, r7 = (int)(m & 7);
    uint8_t* base = reinterpret_cast<uint8_t*>(p) + ((size_t)(m >> 7) * kbs + (c >> 6)) * 16384 +
                    (((m & 127) >> 3) * 1024 + r7 * 128);
    const int ch = (c & 63) >> 3;
#pragma unroll
    for (int i = 0; i < 2; i++) {
      uint2 lo = pack_bf16x4(make_float4(v[i * 8], v[i * 8 + 1], v[i * 8 + 2], v[i * 8 + 3]));
      uint2 hi = pack_bf16x4(make_float4(v[i * 8 + 4], v[i * 8 + 5], v[i * 8 + 6], v[i * 8 + 7]));
      *reinterpret_cast<uint4*>(base + (((ch + i) ^ r7) << 4)) = make_uint4(lo.x, lo.y, hi.x, hi.y);
    }
  } else {
    float* q = p + m * ld + c;
    if (nv >= 16 && (reinterpret_cast<uintptr_t>(q) & 15) == 0) {
#pragma unroll
      for (int i = 0; i < 4; i++)
        reinterpret_cast<float4*>(q)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 16; j++)
        if (j < nv) q[j] = v[j];
    }
  }
}

// One accumulator row (global row m), 16 columns starting at global column n.  sb / sr: staged bias and rvec.
template <int MODE>
__device__ __forceinline__ void epilogue_row16(const Epi& e, long long m, int n, int N, const float* a, const float* sb,
                                               const float* sr) {
  const float beta = e.beta, inv_beta = 1.f / e.beta, oscale = e.oscale, hscale = e.hscale;
  const int lim = (N < e.csplit ? N : e.csplit) - n;        // primary columns of this group: j < lim
  constexpr bool kNeedH = MODE == EPI_SPMUL || MODE == EPI_SWEEP || MODE == EPI_SDF_BWD || MODE == EPI_RELUMASK;
  constexpr bool kNeedQ = MODE == EPI_SWEEP || MODE == EPI_SDF_BWD || MODE == EPI_LINEAR_ADD;
  float h[16], q[16], y[16];
  const bool has_h = kNeedH && e.H != nullptr && lim > 0;
  const bool has_q = kNeedQ && e.Q != nullptr && lim > 0;
  if (has_h) row_load16(e.H, e.ldh, m, n, lim, h);
  if (has_q) row_load16(e.Q, e.ldq, m, n, lim, q);
  const float rs = ((MODE == EPI_SDF_BWD || MODE == EPI_RELUMASK) && e.rs) ? __ldg(e.rs + m) * e.rscale : 0.f;
#pragma unroll
  for (int j = 0; j < 16; j++) {
    float acc = a[j], hv = has_h ? h[j] : 0.f, qv = has_q ? q[j] : 0.f, out = 0.f, out2 = 0.f;
    if (MODE == EPI_LINEAR) out = (acc + sb[n + j]) * oscale;
    else if (MODE == EPI_RELU) out = fmaxf(acc + sb[n + j], 0.f);
    else if (MODE == EPI_SIGMOID) out = sigmoid_fast(acc + sb[n + j]);
    else if (MODE == EPI_SOFTPLUS) out = softplus_beta_fast(acc + sb[n + j], beta, inv_beta) * oscale;
    else if (MODE == EPI_SOFTPLUS_Q) {
      float v = acc + sb[n + j];
      out = softplus_beta_fast(v, beta, inv_beta) * oscale;
      out2 = softplus_grad_from_pre_fast(v, beta) * sr[n + j];
    } else if (MODE == EPI_SPMUL) out = softplus_grad_from_act_fast(hv * hscale, beta) * acc * oscale;
    else if (MODE == EPI_SWEEP) {
      float sg = softplus_grad_from_act_fast(hv * hscale, beta);
      out = sg * acc * oscale;
      out2 = beta * (1.f - sg) * qv * acc;
    } else if (MODE == EPI_SDF_BWD) {
      float sg = softplus_grad_from_act_fast(hv * hscale, beta);
      out = sg * (acc + rs * sr[n + j]) * oscale + qv;
    } else if (MODE == EPI_RELUMASK) {
      float v = acc + rs * sr[n + j];
      out = has_h ? (hv > 0.f ? v : 0.f) : v;
    } else if (MODE == EPI_LINEAR_ADD) out = (acc + qv) * oscale;
    y[j] = j < lim ? out : 0.f;
    q[j] = j < lim ? out2 : 0.f;
  }
  if (e.C != nullptr && (lim > 0 || e.ldc < 0)) row_store16(e.C, e.ldc, m, n, lim, y);
  if ((MODE == EPI_SOFTPLUS_Q || MODE == EPI_SWEEP) && (lim > 0 || e.ldq < 0)) row_store16(e.Q, e.ldq, m, n, lim, q);
  // secondary part (columns >= csplit): rare, element-wise
  if ((MODE == EPI_SPMUL || MODE == EPI_RELUMASK) && e.C2 != nullptr && n + 16 > e.csplit) {
#pragma unroll 1
    for (int j = 0; j < 16; j++) {
      const int nn = n + j;
      if (nn < e.csplit || nn >= N) continue;
      const int c = nn - e.csplit;
      float v = a[j];
      if (MODE == EPI_SPMUL) v *= oscale;
      else {
        if (e.H2) v = mat_get(e.H2, e.ldh2, m, c) > 0.f ? v : 0.f;
        if (e.accumulate2) v += mat_get(e.C2, e.ldc2, m, c);
      }
      mat_put(e.C2, e.ldc2, m, c, v);
    }
  }
}

template <int MODE>
__device__ __forceinline__ void epilogue_tile(const Epi& e, uint32_t taddr, long long m, int M, int n0, int Nc, int N,
                                              int grp, int ngroups, int cover, const float* sb, const float* sr) {
  // cover: columns (relative to n0) that must be written even beyond Nc (image padding), multiple of 32
#pragma unroll 1
  for (int c0 = grp * 32; c0 < cover; c0 += 32 * ngroups) {
#pragma unroll 1
    for (int hseg = 0; hseg < 32; hseg += 16) {
      float a[16];
      if (c0 + hseg < Nc) tmem_ld16(taddr + c0 + hseg, a);
      else {
#pragma unroll
        for (int j = 0; j < 16; j++) a[j] = 0.f;
      }
      if (m < M) epilogue_row16<MODE>(e, m, n0 + c0 + hseg, N, a, sb, sr);
    }
  }
}

template <bool WT>
__global__ void __launch_bounds__(P_THREADS, 1)
tc_gemm_mk_persistent_kernel(ASeg a, int M, int N, Epi e, const uint8_t* __restrict__ wimg, int dbg) {
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* base = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared
                                                 // address space (an integer round trip turned every access into a generic LD.E / ST.E)
  uint8_t* sA[P_STAGES];
  uint8_t* sB[P_STAGES];
#pragma unroll
  for (int s = 0; s < P_STAGES; s++) {
    sA[s] = base + s * (TC_A_BYTES + TC_B_BYTES);
    sB[s] = sA[s] + TC_A_BYTES;
  }
  float* svec = reinterpret_cast<float*>(base + P_STAGES * (TC_A_BYTES + TC_B_BYTES));
  PSmem* ctl = reinterpret_cast<PSmem*>(base + P_STAGES * (TC_A_BYTES + TC_B_BYTES) + P_VEC_FLOATS * 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles_m = (M + TC_BM - 1) / TC_BM;
  const int tiles_n = (N + 255) / 256;
  const int ntiles = tiles_m * tiles_n;
  const int kb_gen = (a.gen.ncols + TC_BK - 1) / TC_BK;
  const int kb_mem = (a.kmem + TC_BK - 1) / TC_BK;
  const int KB = kb_gen + kb_mem;
  const bool a_img = a.ldm < 0;
  // when every reduction block of A arrives by bulk copy the 8 producer warps have nothing to produce: they join
  // the epilogue (4 column groups instead of 2), halving its per-tile latency chain
  const bool extra_epi = a_img && kb_gen == 0;
  const int ngroups = extra_epi ? 4 : 2;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < P_STAGES; s++) { mbar_init(&ctl->full[s], extra_epi ? 1 : 129); mbar_init(&ctl->empty[s], 1); }
#pragma unroll
    for (int s = 0; s < 2; s++) { mbar_init(&ctl->acc_full[s], 1); mbar_init(&ctl->acc_empty[s], 128 * ngroups); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // stage the per-column vectors of the epilogue (bias, rvec) once
  for (int i = tid; i < 1024; i += P_THREADS) {
    svec[i] = (e.bias && i < N) ? __ldg(e.bias + i) : 0.f;
    svec[1024 + i] = (e.rvec && i < N) ? __ldg(e.rvec + i) : 0.f;
  }
  if (warp == 8) tmem_alloc(&ctl->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp < 8 && !extra_epi) {
    // ------------------------------ A producers ------------------------------
    const int grp = warp >> 2, ptid = tid & 127;
    int kbg = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const long long m0 = (long long)(t % tiles_m) * TC_BM;
      for (int kb = 0; kb < KB; kb++, kbg++) {
        if ((kbg & 1) != grp) continue;
        const int s = kbg % P_STAGES;
        if (kbg >= P_STAGES) mbar_wait(&ctl->empty[s], ((kbg / P_STAGES) - 1) & 1);
        const int phase = kb < kb_gen ? 0 : 1;
        const int k0 = (phase == 0 ? kb : kb - kb_gen) * TC_BK;
        if (phase == 0) {
          // generated columns: one row per thread, zero the 128-byte row segment then fill the valid columns
          const int r7 = ptid & 7;
          uint8_t* rowp = sA[s] + (ptid >> 3) * 1024 + r7 * 128;
#pragma unroll
          for (int c = 0; c < 8; c++) *reinterpret_cast<uint4*>(rowp + c * 16) = make_uint4(0u, 0u, 0u, 0u);
          const long long m = m0 + ptid;
          if (m < M) {
            gen_row(a.gen, m, [&](int j, float val) {
              const int jj = j - k0;
              if (jj >= 0 && jj < TC_BK)
                *reinterpret_cast<unsigned short*>(rowp + ((((jj >> 3) ^ r7) & 7) << 4) + ((jj & 7) << 1)) =
                    f32_to_bf16_bits(val);
            });
          }
          fence_proxy_async();
        } else if (!a_img && !(dbg & 1)) {
          float4 av[16];
#pragma unroll
          for (int it = 0; it < 16; it++) {
            int idx = it * 128 + ptid;
            av[it] = load_a4(a, 1, m0 + (idx >> 4), M, k0 + ((idx & 15) << 2));
          }
#pragma unroll
          for (int it = 0; it < 16; it++) {
            int idx = it * 128 + ptid;
            sts_kmajor(sA[s], idx >> 4, (idx & 15) << 2, av[it]);
          }
          fence_proxy_async();
        }
        mbar_arrive(&ctl->full[s]);
      }
    }
  } else if (warp == 9) {
    // ------------------------------ bulk-copy loader ------------------------------
    if (lane == 0) {
      int kbg = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int nchunk = t / tiles_m, mt = t % tiles_m;
        const int nvalid = min(256, N - nchunk * 256);
        const int Nc = (nvalid + 15) & ~15;
        const uint32_t bytes = WT ? (uint32_t)Nc * 128u : (uint32_t)((Nc + 63) / 64) * 8192u;
        const uint8_t* src = wimg + (size_t)nchunk * KB * TC_B_BYTES;
        for (int kb = 0; kb < KB; kb++, kbg++) {
          const int s = kbg % P_STAGES;
          if (kbg >= P_STAGES) mbar_wait(&ctl->empty[s], ((kbg / P_STAGES) - 1) & 1);
          const bool img_blk = a_img && kb >= kb_gen && !(dbg & 1);
          mbar_arrive_expect_tx(&ctl->full[s], bytes + (img_blk ? (uint32_t)TC_A_BYTES : 0u));
          if (img_blk)
            bulk_g2s(sA[s], reinterpret_cast<const uint8_t*>(a.mem) + ((size_t)mt * (-a.ldm) + (kb - kb_gen)) * TC_A_BYTES,
                     TC_A_BYTES, &ctl->full[s]);
          bulk_g2s(sB[s], src + (size_t)kb * TC_B_BYTES, bytes, &ctl->full[s]);
        }
      }
    }
  } else if (warp == 8) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      int kbg = 0, it = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, it++) {
        const int nchunk = t / tiles_m;
        const int nvalid = min(256, N - nchunk * 256);
        const int Nc = (nvalid + 15) & ~15;
        const uint32_t idesc = make_idesc(Nc, 0, WT ? 0 : 1);
        const int as = it & 1;
        if (it >= 2) mbar_wait(&ctl->acc_empty[as], ((it >> 1) - 1) & 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * 256;
        for (int kb = 0; kb < KB; kb++, kbg++) {
          const int s = kbg % P_STAGES;
          mbar_wait(&ctl->full[s], (kbg / P_STAGES) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA[s]), b_addr = smem_u32(sB[s]);
          if (!(dbg & 4)) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
              uint64_t ad = make_desc(a_addr + k * 32, 16, 1024);
              uint64_t bd = WT ? make_desc(b_addr + k * 32, 16, 1024) : make_desc(b_addr + k * 2048, 8192, 1024);
              umma_bf16(tmem_d, ad, bd, idesc, (kb > 0 || k > 0) ? 1 : 0);
            }
          }
          umma_commit(&ctl->empty[s]);
        }
        umma_commit(&ctl->acc_full[as]);
      }
      tc_fence_before();
    }
  } else {
    // ------------------------------ epilogue ------------------------------
    const int grp = warp >= 10 ? (warp - 10) >> 2 : 2 + (warp >> 2);   // column-chunk group of this warp
    const int quarter = warp & 3;              // TMEM lane quarter this warp may access
    const float* sb = svec;
    const float* sr = svec + 1024;
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, it++) {
      const long long m = (long long)(t % tiles_m) * TC_BM + quarter * 32 + lane;
      const int n0 = (t / tiles_m) * 256;
      const int nvalid = min(256, N - n0);
      const int Nc = (nvalid + 15) & ~15;
      // image destinations: cover their padding columns inside this CTA's 256-column range with zeros
      int cover = (Nc + 31) & ~31;
      if (e.ldc < 0 && e.C) cover = max(cover, min(256, (-e.ldc) * 64 - n0));
      if (e.ldq < 0 && e.Q && (e.mode == EPI_SOFTPLUS_Q || e.mode == EPI_SWEEP)) cover = max(cover, min(256, (-e.ldq) * 64 - n0));
      const int as = it & 1;
      mbar_wait(&ctl->acc_full[as], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * 256 + ((uint32_t)(quarter * 32) << 16);
      if (!(dbg & 2)) {
        switch (e.mode) {
          case EPI_LINEAR: epilogue_tile<EPI_LINEAR>(e, taddr, m, M, n0, Nc, N, grp, ngroups, cover, sb, sr); break;
          case EPI_RELU: epilogue_tile<EPI_RELU>(e, taddr, m, M, n0, Nc, N, grp, ngroups, cover, sb, sr); break;
          case EPI_SIGMOID: epilogue_tile<EPI_SIGMOID>(e, taddr, m, M, n0, Nc, N, grp, ngroups, cover, sb, sr); break;
          case EPI_SOFTPLUS: epilogue_tile<EPI_SOFTPLUS>(e, taddr, m, M, n0, Nc, N, grp, ngroups, cover, sb, sr); break;
          case EPI_SOFTPLUS_Q: epilogue_tile<EPI_SOFTPLUS_Q>(e, taddr, m, M, n0, Nc, N, grp, ngroups, cover, sb, sr); break;
          case EPI_SPMUL: epilogue_tile<EPI_SPMUL>(e, taddr, m, M, n0, Nc, N, grp, ngroups, cover, sb, sr); break;
          case EPI_SWEEP: epilogue_tile<EPI_SWEEP>(e, taddr, m, M, n0, Nc, N, grp, ngroups, cover, sb, sr); break;
          case EPI_SDF_BWD: epilogue_tile<EPI_SDF_BWD>(e, taddr, m, M, n0, Nc, N, grp, ngroups, cover, sb, sr); break;
          case EPI_RELUMASK: epilogue_tile<EPI_RELUMASK>(e, taddr, m, M, n0, Nc, N, grp, ngroups, cover, sb, sr); break;
          case EPI_LINEAR_ADD: epilogue_tile<EPI_LINEAR_ADD>(e, taddr, m, M, n0, Nc, N, grp, ngroups, cover, sb, sr); break;
          default: break;                    // EPI_SDF_OUT is never routed to the persistent kernel
        }
      }
      tc_fence_before();
      mbar_arrive(&ctl->acc_empty[as]);
    }
  }
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// Weight images: BF16 copies of a layer's weights laid out as the exact shared-memory tiles the MMA reads
// (tile = one 256-column chunk x one 64-deep reduction block, 32 KB, 128B-swizzled), so a B tile is one bulk copy.
// WT : tile rows = output features n, K-major.   !WT: tile rows = reduction index (W rows), MN-major.
// ---------------------------------------------------------------------------------------------
inline size_t wimg_bytes(int N, int kgen, int kmem) {
  return (size_t)cdiv(N, 256) * (cdiv(kgen, TC_BK) + cdiv(kmem, TC_BK)) * TC_B_BYTES;
}
template <bool WT>
__global__ void pack_wimg_kernel(const float* __restrict__ W, int ldw, int wout0, int N, int wred_gen, int kgen,
                                 int wred_mem, int kmem, uint8_t* __restrict__ img) {
  const int kb_gen = (kgen + TC_BK - 1) / TC_BK, kb_mem = (kmem + TC_BK - 1) / TC_BK;
  const int KB = kb_gen + kb_mem;
  const long long total = (long long)((N + 255) / 256) * KB * 2048;   // 16-byte chunks
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int tile = (int)(idx / 2048), within = (int)(idx % 2048);
  const int nchunk = tile / KB, kb = tile % KB;
  const int phase = kb < kb_gen ? 0 : 1;
  const int k0 = (phase == 0 ? kb : kb - kb_gen) * TC_BK;
  const int kmax = phase == 0 ? kgen : kmem;
  const int wred0 = phase == 0 ? wred_gen : wred_mem;
  const int n0 = nchunk * 256;
  float v[8];
  int off;
  if (WT) {
    const int r = within >> 3, c = within & 7;              // row n, 16B chunk along k
    const int n = n0 + r;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      int k = k0 + c * 8 + i;
      v[i] = (n < N && k < kmax) ? __ldg(W + (long long)(wout0 + n) * ldw + wred0 + k) : 0.f;
    }
    off = (r >> 3) * 1024 + (r & 7) * 128 + (((c ^ (r & 7)) & 7) << 4);
  } else {
    const int nb = within >> 9, rem = within & 511;         // 64-column block, k-row, 16B chunk along n
    const int kr = rem >> 3, c = rem & 7;
    const int k = k0 + kr;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      int n = n0 + nb * 64 + c * 8 + i;
      v[i] = (n < N && k < kmax) ? __ldg(W + (long long)(wred0 + k) * ldw + wout0 + n) : 0.f;
    }
    off = nb * 8192 + (kr >> 3) * 1024 + (kr & 7) * 128 + (((c ^ (kr & 7)) & 7) << 4);
  }
  uint2 lo = pack_bf16x4(make_float4(v[0], v[1], v[2], v[3]));
  uint2 hi = pack_bf16x4(make_float4(v[4], v[5], v[6], v[7]));
  *reinterpret_cast<uint4*>(img + (size_t)tile * TC_B_BYTES + off) = make_uint4(lo.x, lo.y, hi.x, hi.y);
}
// Several layers' images in one launch: blockIdx.y = job.
constexpr int PACK_MAX_JOBS = 24;
struct PackJob {
  const float* W; uint8_t* img;
  int ldw, wout0, N, wred_gen, kgen, wred_mem, kmem, for_bwd_data, f16;
};
struct PackJobs { int n; PackJob job[PACK_MAX_JOBS]; };

template <bool WT>
__device__ __forceinline__ void pack_wimg_chunk(const PackJob& j, long long idx) {
  const float* __restrict__ W = j.W;
  const int ldw = j.ldw, wout0 = j.wout0, N = j.N;
  const int kb_gen = (j.kgen + TC_BK - 1) / TC_BK, kb_mem = (j.kmem + TC_BK - 1) / TC_BK;
  const int KB = kb_gen + kb_mem;
  const int tile = (int)(idx / 2048), within = (int)(idx % 2048);
  const int nchunk = tile / KB, kb = tile % KB;
  const int phase = kb < kb_gen ? 0 : 1;
  const int k0 = (phase == 0 ? kb : kb - kb_gen) * TC_BK;
  const int kmax = phase == 0 ? j.kgen : j.kmem;
  const int wred0 = phase == 0 ? j.wred_gen : j.wred_mem;
  const int n0 = nchunk * 256;
  float v[8];
  int off;
  if (WT) {
    const int r = within >> 3, c = within & 7;
    const int n = n0 + r;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      int k = k0 + c * 8 + i;
      v[i] = (n < N && k < kmax) ? __ldg(W + (long long)(wout0 + n) * ldw + wred0 + k) : 0.f;
    }
    off = (r >> 3) * 1024 + (r & 7) * 128 + (((c ^ (r & 7)) & 7) << 4);
  } else {
    const int nb = within >> 9, rem = within & 511;
    const int kr = rem >> 3, c = rem & 7;
    const int k = k0 + kr;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      int n = n0 + nb * 64 + c * 8 + i;
      v[i] = (n < N && k < kmax) ? __ldg(W + (long long)(wred0 + k) * ldw + wout0 + n) : 0.f;
    }
    off = nb * 8192 + (kr >> 3) * 1024 + (kr & 7) * 128 + (((c ^ (kr & 7)) & 7) << 4);
  }
  *reinterpret_cast<uint4*>(j.img + (size_t)tile * TC_B_BYTES + off) = f32x8_to_u16(v, j.f16 != 0);
}
__global__ void pack_wimg_multi_kernel(PackJobs jobs) {
  const PackJob& j = jobs.job[blockIdx.y];
  const long long total = (long long)((j.N + 255) / 256) * ((j.kgen + TC_BK - 1) / TC_BK + (j.kmem + TC_BK - 1) / TC_BK) * 2048;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    if (j.for_bwd_data) pack_wimg_chunk<false>(j, idx);
    else pack_wimg_chunk<true>(j, idx);
  }
}

constexpr int TC_SMEM_BYTES = TC_STAGES * (TC_A_BYTES + TC_B_BYTES) + 1024 + 256;
constexpr int WG_SMEM_BYTES = WG_STAGES * (TC_A_BYTES + TC_B_BYTES) + 1024 + 256;

// development knobs (environment, read once): FNEUS_WG_STAGES = 2 | 4 (ring depth; 2 -> two CTAs per SM),
// FNEUS_WG_NOCONV = 1 (timing experiment only: skips the operand-format conversion, results are wrong)
inline int env_int(const char* name, int dflt) { const char* v = getenv(name); return v ? atoi(v) : dflt; }
inline int& tc_wgrad_stages() { static int v = env_int("FNEUS_WG_STAGES", 2) == 4 ? 4 : 2; return v; }
inline int& tc_wgrad_noconv() { static int v = env_int("FNEUS_WG_NOCONV", 0); return v; }
inline int& tc_wgrad_ctas_per_sm() { static int v = tc_wgrad_stages() == 4 ? 1 : 2; return v; }
inline int tc_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}
inline unsigned long long*& hang_rec_host() { static unsigned long long* p = nullptr; return p; }
inline int tc_prepare() {
  static int done = 0;
  if (done) return 0;
  if (hang_rec_host() == nullptr) {
    // host-mapped record of the wait watchdog (readable after a trap has killed the context)
    unsigned long long* h = nullptr;
    unsigned long long* d = nullptr;
    if (cudaHostAlloc(reinterpret_cast<void**>(&h), 64 * 8, cudaHostAllocMapped) == cudaSuccess && h != nullptr) {
      for (int i = 0; i < 64; i++) h[i] = 0ull;
      if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&d), h, 0) == cudaSuccess &&
          cudaMemcpyToSymbol(g_hang_rec, &d, sizeof(d)) == cudaSuccess)
        hang_rec_host() = h;
    }
    (void)cudaGetLastError();
  }
  cudaError_t e1 = cudaFuncSetAttribute(tc_gemm_mk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  cudaError_t e2 = cudaFuncSetAttribute(tc_gemm_mk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  cudaError_t e3 = cudaFuncSetAttribute(tc_gemm_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  if (e3 == cudaSuccess)
    e3 = cudaFuncSetAttribute(tc_gemm_wgrad_group_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
  if (e3 == cudaSuccess)
    e3 = cudaFuncSetAttribute(tc_gemm_wgrad_group_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES);
  cudaError_t e4 = cudaFuncSetAttribute(tc_gemm_mk_persistent_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES);
  cudaError_t e5 = cudaFuncSetAttribute(tc_gemm_mk_persistent_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES);
  if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess || e4 != cudaSuccess || e5 != cudaSuccess) return 1;
  done = 1;
  return 0;
}

inline void launch_tc_fwd(const ASeg& a, const float* W, int ldw, int wout0, long long M, int N, const Epi& e,
                          cudaStream_t st, const uint8_t* wimg) {
  dim3 grid(cdiv(M, TC_BM), cdiv(N, 256));
  prof_begin(PC_TC_MLP, 2.0 * (double)M * N * (a.gen.ncols + a.kmem), 0.0, st);
  if (wimg && a.gen.ncols + a.kmem > 0 && e.mode != EPI_SDF_OUT && N <= 1024) {
    int ntiles = grid.x * grid.y, sms = tc_num_sms();
    tc_gemm_mk_persistent_kernel<true><<<ntiles < sms ? ntiles : sms, P_THREADS, P_SMEM_BYTES, st>>>(a, (int)M, N, e, wimg, tc_debug_flags());
  } else {
    tc_gemm_mk_kernel<true><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(a, W, ldw, wout0, (int)M, N, e, wimg);
  }
  prof_end(st);
}
inline void launch_tc_bwd_data(const ASeg& a, const float* W, int ldw, int wout0, long long M, int N, const Epi& e,
                               cudaStream_t st, const uint8_t* wimg) {
  dim3 grid(cdiv(M, TC_BM), cdiv(N, 256));
  prof_begin(PC_TC_MLP, 2.0 * (double)M * N * (a.gen.ncols + a.kmem), 0.0, st);
  if (wimg && a.gen.ncols + a.kmem > 0 && e.mode != EPI_SDF_OUT && N <= 1024) {
    int ntiles = grid.x * grid.y, sms = tc_num_sms();
    tc_gemm_mk_persistent_kernel<false><<<ntiles < sms ? ntiles : sms, P_THREADS, P_SMEM_BYTES, st>>>(a, (int)M, N, e, wimg, tc_debug_flags());
  } else {
    tc_gemm_mk_kernel<false><<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(a, W, ldw, wout0, (int)M, N, e, wimg);
  }
  prof_end(st);
}
inline void launch_tc_wgrad(const float* dY, int ldy, const ASeg& a, float* dW, int ldw, int wout0, float* db,
                            long long M, int N, int num_sms, cudaStream_t st) {
  int kchunks = cdiv(a.gen.ncols, 256) + cdiv(a.kmem, 256);
  if (kchunks == 0) return;
  int tiles = cdiv(N, 128) * kchunks;
  int splits = (tc_wgrad_ctas_per_sm() * num_sms + tiles - 1) / tiles;   // RED traffic grows with the split count
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

// Collects the weight-gradient GEMMs of a chain and launches them together (tensor-core mode only).  The split
// count over M is chosen for the whole group: about `waves` full waves of CTAs in total, so that the RED traffic
// (one 128 x 256 FP32 tile per CTA) stays small next to the operand traffic.
inline int& tc_wgrad_group_waves() { static int v = 2; return v; }
struct WgradGroup {
  WgradJobs jobs;
  double flops, bytes;               // algorithmic FLOPs; designed DRAM bytes (both operands once + the FP32 gradient)
  int num_sms;
  void reset(long long M, int sms) { jobs.n = 0; jobs.M = (int)M; flops = 0.0; bytes = 0.0; num_sms = sms; }
  void add(const float* dY, int ldy, const ASeg& a, float* dW, int ldw, int wout0, float* db, int N, cudaStream_t st,
           int y_f16 = 0, int x_f16 = 0) {
    const long long M = jobs.M;
    int kchunks = cdiv(a.gen.ncols, 256) + cdiv(a.kmem, 256);
    if (kchunks == 0 || M <= 0 || N <= 0) return;
    if (jobs.n == WG_MAX_JOBS) flush(st);
    WgradJob& j = jobs.job[jobs.n++];
    j.dY = dY; j.ldy = ldy; j.a = a; j.dW = dW; j.ldw = ldw; j.wout0 = wout0; j.db = db; j.N = N;
    j.y_f16 = y_f16; j.x_f16 = x_f16;
    j.tiles = cdiv(N, 128) * kchunks;
    flops += 2.0 * (double)M * N * (a.gen.ncols + a.kmem);
    bytes += (double)M * ((ldy < 0 ? 2.0 : 4.0) * N + (a.ldm < 0 ? 2.0 : 4.0) * a.kmem + 12.0 * (a.gen.ncols > 0 ? 1 : 0)) +
             4.0 * (double)N * (a.gen.ncols + a.kmem);
  }
  void flush(cudaStream_t st) {
    if (jobs.n == 0) return;
    const long long M = jobs.M;
    int total_tiles = 0, max_tiles = 0;
    for (int i = 0; i < jobs.n; i++) {
      total_tiles += jobs.job[i].tiles;
      if (jobs.job[i].tiles > max_tiles) max_tiles = jobs.job[i].tiles;
    }
    // whole waves: never more CTAs than `waves` rounds of the machine hold (a few CTAs spilling into an extra round cost a
    // full round of latency)
    int splits = env_int("FNEUS_WG_CEIL", 0) ? (tc_wgrad_group_waves() * tc_wgrad_ctas_per_sm() * num_sms + total_tiles - 1) / total_tiles
                                             : (tc_wgrad_group_waves() * tc_wgrad_ctas_per_sm() * num_sms) / total_tiles;
    const int max_s = cdiv(M, 4 * TC_BK);
    if (splits > max_s) splits = max_s;
    if (splits < 1) splits = 1;
    const int mps = round_up(cdiv(M, splits), TC_BK);
    splits = cdiv(M, mps);
    for (int i = 0; i < jobs.n; i++) { jobs.job[i].m_per_split = mps; jobs.job[i].splits = splits; }
    prof_begin(PC_TC_WGRAD, flops, bytes, st);
    jobs.no_conv = tc_wgrad_noconv();
    if (tc_wgrad_stages() == 4)
      tc_gemm_wgrad_group_kernel<4><<<dim3(max_tiles, splits, jobs.n), TC_THREADS, WG_SMEM_BYTES, st>>>(jobs);
    else
      tc_gemm_wgrad_group_kernel<2><<<dim3(max_tiles, splits, jobs.n), TC_THREADS, TC_SMEM_BYTES, st>>>(jobs);
    prof_end(st);
    reset(M, num_sms);
  }
};

// ---------------------------------------------------------------------------------------------
// precision dispatch: 0 = FP32 SIMT (exactness anchor), 1 = BF16 tcgen05 (FP32 accumulate)
// ---------------------------------------------------------------------------------------------
// The precision of a call comes from its configuration struct (include/fneus.h: fneus_*_cfg.precision) and is held in a
// thread-local for the duration of that ABI call (PrecScope) -- no process-global state, calls are re-entrant.
inline int& precision_mode() { static thread_local int m = 0; return m; }
struct PrecScope {
  int saved;
  explicit PrecScope(int p) : saved(precision_mode()) { precision_mode() = (p == 1) ? 1 : 0; }
  ~PrecScope() { precision_mode() = saved; }
};

// A bump allocator over a caller-provided byte region for the weight images of one ABI call.  make_wimg only
// RECORDS the packing job; ImgArena::flush packs every recorded image in one launch and must be called before the
// first GEMM that consumes one (images depend only on the weights, so a pass declares all of them up front).
struct ImgArena {
  uint8_t* base; size_t cap, used;
  PackJobs jobs;
  uint8_t* take(size_t bytes) {
    size_t off = (used + 1023) & ~(size_t)1023;
    if (!base || off + bytes > cap) return nullptr;
    used = off + bytes;
    return base + off;
  }
  void flush(cudaStream_t st) {
    if (jobs.n == 0) return;
    prof_begin(PC_ELEMENTWISE, 0.0, 0.0, st);
    pack_wimg_multi_kernel<<<dim3(32, jobs.n), 256, 0, st>>>(jobs);
    prof_end(st);
    jobs.n = 0;
  }
};
inline ImgArena arena_make(uint8_t* base, size_t cap) {
  ImgArena ar;
  ar.base = base; ar.cap = cap; ar.used = 0; ar.jobs.n = 0;
  return ar;
}
inline const uint8_t* make_wimg(ImgArena& ar, bool for_bwd_data, const float* W, int ldw, int wout0, int N,
                                int wred_gen, int kgen, int wred_mem, int kmem, cudaStream_t st, int f16 = 0) {
  if (precision_mode() != 1) return nullptr;
  uint8_t* img = ar.take(wimg_bytes(N, kgen, kmem));
  if (!img) return nullptr;
  if (ar.jobs.n == PACK_MAX_JOBS) ar.flush(st);
  PackJob& j = ar.jobs.job[ar.jobs.n++];
  j.W = W; j.img = img; j.ldw = ldw; j.wout0 = wout0; j.N = N; j.wred_gen = wred_gen; j.kgen = kgen;
  j.wred_mem = wred_mem; j.kmem = kmem; j.for_bwd_data = for_bwd_data ? 1 : 0; j.f16 = f16;
  return img;
}

inline void launch_gemm_fwd(const ASeg& a, const float* W, int ldw, int wout0, long long M, int N, const Epi& e,
                            cudaStream_t st, const uint8_t* wimg = nullptr) {
  if (M <= 0 || N <= 0) return;
  if (precision_mode() == 1 && tc_prepare() == 0) launch_tc_fwd(a, W, ldw, wout0, M, N, e, st, wimg);
  else launch_simt_fwd(a, W, ldw, wout0, M, N, e, st);
}
inline void launch_gemm_bwd_data(const ASeg& a, const float* W, int ldw, int wout0, long long M, int N,
                                 const Epi& e, cudaStream_t st, const uint8_t* wimg = nullptr) {
  if (M <= 0 || N <= 0) return;
  if (precision_mode() == 1 && tc_prepare() == 0) launch_tc_bwd_data(a, W, ldw, wout0, M, N, e, st, wimg);
  else launch_simt_bwd_data(a, W, ldw, wout0, M, N, e, st);
}
inline void launch_gemm_wgrad(const float* dY, int ldy, const ASeg& a, float* dW, int ldw, int wout0, float* db,
                              long long M, int N, int num_sms, cudaStream_t st) {
  if (M <= 0 || N <= 0) return;
  if (precision_mode() == 1 && tc_prepare() == 0) launch_tc_wgrad(dY, ldy, a, dW, ldw, wout0, db, M, N, num_sms, st);
  else launch_simt_wgrad(dY, ldy, a, dW, ldw, wout0, db, M, N, num_sms, st);
}

}  // namespace fneus

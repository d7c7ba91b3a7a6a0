// Fused layer chains (BF16 tensor-core mode): the layer sequences that a training step runs on the render_core points,
// each as ONE persistent kernel, one 128-point tile at a time per CTA, the running operand on chip between layers
// (shared memory BF16 <-> TMEM FP32), weights streamed as pre-packed images.  The Softplus SDF network:
//
//   forward  : value chain h_{l+1} = softplus(W_l h_l + b_l) (fields.py:74-95), sdf = h_L . W_L[0] and the feature
//              GEMM, then the reverse chain of the analytic gradient q_{l-1} = s_l * (q_l W_l) down to g_0 and the
//              normal (fields.py:101-111 as reverse-mode) -- 2L+1 GEMM steps
//   backward : the double-backward sweep gbar_{l+1} = s_l * (gbar_l W_l^T), e_l = beta (1 - s_l) q_l (gbar_l W_l^T),
//              then the value-path backward abar_{l-1} = s_l * (abar_l W_l) + e_{l-1} -- 2L GEMM steps
//
// and the ReLU networks (RenderingNetwork fields.py:114-175, RefColor fields.py:271-335, Lvis fields.py:338-369):
//
//   forward  : [generated PE block | feature blocks] -> ReLU layers (activation images kept for the backward) ->
//              FP32 output rows (sigmoid)
//   backward : dz_{l-1} = (dz_l W_l) * [h_l > 0] with the h_l blocks bulk-loaded, then the two input-gradient GEMMs
//              (feature columns, generated columns) off the same operand
//
// ALL bulk traffic goes through the async copy engine, never through per-thread global accesses (a thread-per-row
// access touches 32 different 128-byte lines per warp instruction and serialises in L1): weight half-tiles and the
// auxiliary activation blocks (h_l, q_l / e_l, 128 rows x 64 columns = 16 KB) arrive by bulk copies into rings,
// results leave as bulk stores of the operand blocks themselves (their bytes ARE the activation image) and of the
// auxiliary blocks rewritten in place (e_l over h_l).  FP32 row-major outputs (features) are transposed through
// shared memory and written as whole lines.  The positional-encoding part of the skip gradient waits in shared
// memory instead of HBM; q_{L-1} is rebuilt from the pre-activations of layer L-1 still held in TMEM.
//
// Software pipeline: TMEM holds TWO 256-column accumulators.  The epilogue of step s publishes the next operand one
// 64-column block at a time (a_ready[b]); the MMA issuer starts step s+1 on block 0 into the other accumulator while
// the epilogue of step s is still working on blocks 1..3, so the tensor pipe runs underneath the epilogue and only
// the last quarter of a step's MMAs is exposed.
//
// Operand formats (fneus_common.cuh, Fmt16): the forward families (FAM_SDF_FWD, forward ReLU chains) compute on FP16
// operands and store FP16 images; the backward families (FAM_SDF_BWD, backward ReLU chains) compute on BF16 operands,
// store BF16 images and READ the forward pass's FP16 images (h_l for the activation derivative, q_l).
//
//   warp 0      : MMA issuer (+ TMEM alloc: 2 x 256 accumulator columns)
//   warp 1      : weight-image loader, ring of 4 half-tiles (128 output columns x 64 reduction, 16 KB)
//   warp 2      : auxiliary-block loader (2 slots of h + q blocks)
//   warp 3      : storer: bulk shared->global of finished blocks, releases the slots
//   warps 4-19  : epilogue, thread = one row (TMEM lane) x one of four 16-column groups per 64-column block
#pragma once
#include "gemm_tc.cuh"

namespace fneus {

constexpr int CH_WBYTES = 16384;                                  // one weight half-tile: 128 output columns x 64 reduction
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

constexpr int SC_THREADS = 640, SC_WSTAGES = 4, SC_MAXS = 20, SC_EPI_THREADS = 512;
constexpr int SC_NAR = 6;                                         // operand blocks a step may read (a_ready barriers)
constexpr int SC_OPB_RELU = 5;                                    // running-operand blocks of the ReLU family
constexpr int SC_AUXGEN_MAX = 3;                                  // generated blocks that may live in the auxiliary area
// SDF modes (Softplus network) and the ReLU-network modes (RenderingNetwork / RefColor / Lvis chains):
//   SC_RELU : y = max(acc + bias, 0) -> next operand (+ image)        SC_MASK : y = acc * [h > 0] -> next operand (+ image)
//   SC_OUT  : FP32 result rows to HBM (bias, optional sigmoid, optional accumulate); the operand stays as it is
// SC_RELU with act = 2 is a plain linear layer (bias only); SC_MASK without an h image passes the product through, and
// with use_rs adds the rank-1 term rs[m] * rvec[n] before the mask (a 1-wide head next to a wide one: NeRF's alpha_linear).
enum SdfStepMode { SC_SOFTPLUS = 0, SC_FEATQ, SC_SPMUL, SC_G0, SC_SWEEP, SC_SDFBWD, SC_RELU, SC_MASK, SC_OUT };
// SRC_GENMEM: [one generated block | memory blocks (FP32 row-major, or an activation image when ldm < 0)]
// SRC_AUXGEN: up to SC_AUXGEN_MAX generated blocks written once per tile into the AUXILIARY area, where they stay for the
//             whole chain: a step reads its first kb_op operand blocks from the running operand and the remaining
//             KB - kb_op blocks from auxiliary blocks aux_blk0.. (the outside NeRF: PE(pts) is the input of layer 0 AND is
//             concatenated to the hidden state at the skip, PE(views) joins the feature vector; fields.py:233-259)
enum SdfStepSrc { SRC_CHAIN = 0, SRC_PE, SRC_TAN, SRC_MEM, SRC_GENMEM, SRC_AUXGEN };

struct SdfStep {
  const uint8_t* wimg;  // weight image ([1 n-chunk][KB] tiles of 32 KB)
  const float* bias;    // SOFTPLUS / FEATQ
  float* img_out;       // image of the step's result (4 blocks per tile) or null
  const float* h;       // SPMUL / SWEEP / SDFBWD: image of the forward activation the softplus derivative comes from
  const float* q;       // SWEEP: image q_l ; SDFBWD: image e_{l-1}
  float* e_out;         // SWEEP: image e_l (rewritten over the h block in shared memory)
  float* out;           // FEATQ: features FP32 [M, ldo] ; G0: normal FP32 [M, d_in]
  int KB, N, mode, bmn, src;
  int kb_op, aux_blk0;  // operand blocks [0, kb_op) come from the running operand, [kb_op, KB) from auxiliary block aux_blk0 + ..
  int ldo, csplit, append, dot, use_rs, bias_slot;
  int act, accumulate;  // OUT: act 1 = sigmoid ; accumulate: out += result
  int sync_stores;      // storer: wait for full completion of every store so far after this step (later steps read them back)
  int wait_sync;        // aux loader: wait for that completion before this step's first load
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
  float* a0_img;        // SRC_GENMEM: optional image copy of the whole first operand ([M, 1 + kb_mem blocks]);
                        // SRC_AUXGEN: optional image copy of the auxiliary generated blocks ([M, aux_gen_blocks blocks])
  int aux_gen_blocks;   // SRC_AUXGEN: generated blocks in the auxiliary area (columns of `gen`, 64 per block)
  const float* rvec;    // row 0 of the last linear [<= 256]
  const float* b_last;  // its bias
  float* sdf_out;       // [M]
  float sdf_scale;      // out_sign / scale
  const float* rs;      // SDFBWD with use_rs: d_sdf [M]
  float rscale, beta;
  int f16;              // FAM_RELU: operands / weight images / stored images are FP16 (forward chains) or BF16 (backward)
  int dbg;              // record the debug timeline (CTA 0)
  int xflags;           // experiment switches (fneus_debug_flags bits 8..): 1 = every thread arrives, 2 = no suspend hint,
                        // 4 = no accumulator prefetch, 8 = no early start on the first column half
  long long M;
};
// debug timeline (fneus_debug_flags bit 6): clock64 stamps of CTA 0's first epilogue thread and MMA thread
__device__ unsigned long long g_sc_dbg[8192];
// compiled in only with -DFNEUS_SC_TIMELINE: the stamps sit in the innermost (per 64-column block) loop, where every
// predicate costs issue slots (ncu: control flow was 25% of the executed instructions of the forward chain)
#ifdef FNEUS_SC_TIMELINE
#define SC_STAMP(slot) do { if (dbg_on) { g_sc_dbg[dbg_i] = ((unsigned long long)(slot) << 56) | (clock64() & 0x00FFFFFFFFFFFFFFull); dbg_i = dbg_i < 8190 ? dbg_i + 1 : dbg_i; } } while (0)
#else
#define SC_STAMP(slot) do { } while (0)
#endif

struct SCSmem {
  uint64_t wfull[SC_WSTAGES], wempty[SC_WSTAGES];
  uint64_t a_ready[SC_NAR], acc_full, acc_half0, op_free, st_sync;
  uint64_t aux_full[3], aux_empty[3], blk_done[4];   // blk_done: by block counter & 3 (see the epilogue's slot comment)
  uint32_t tmem_base;
};
// Auxiliary slots (h block + q block, 32 KB each) and weight stages per family: the backward chain is HBM-bound and
// needs bytes in flight (Little: 6.5 TB/s x ~1.2 us = 53 KB per SM), so it trades one weight stage for a third slot.
template <int FAM> __host__ __device__ constexpr int sc_nslot() { return FAM == 1 ? 3 : 2; }
// bias rows staged in shared memory: the ReLU family runs the 12-linear NeRF chain
template <int FAM> __host__ __device__ constexpr int sc_bias_slots() { return FAM == 2 ? 12 : 10; }
template <int FAM> __host__ __device__ constexpr int sc_wstages() { return FAM == 1 ? 3 : SC_WSTAGES; }
constexpr int SC_PARK_LD = 41;                                   // odd row stride: thread-per-row accesses hit 32 banks
// Three instantiations, each compiling only its own modes (the union was 127 KB of SASS and ran 15% slower on
// instruction fetch): FAM_SDF_FWD (4 operand blocks = 256 columns, parking area for the skip gradient), FAM_SDF_BWD,
// FAM_RELU (first operand up to 320 columns: [generated | features]).
enum ChainFamily { FAM_SDF_FWD = 0, FAM_SDF_BWD = 1, FAM_RELU = 2 };
template <int FAM>
constexpr int sc_smem_bytes() {
  constexpr int OPB = FAM == FAM_RELU ? SC_OPB_RELU : 4;
  constexpr bool PARK = FAM == FAM_SDF_FWD;
  return OPB * TC_A_BYTES + sc_wstages<FAM>() * CH_WBYTES + sc_nslot<FAM>() * 2 * TC_A_BYTES +
         (sc_bias_slots<FAM>() * 256 + 256 + (PARK ? 512 : 128) + (PARK ? 128 * SC_PARK_LD : 0)) * 4 + 1024 + 256;
}

__device__ __forceinline__ void bf16x8_to_f32(const uint4 u, float* out) { u16x8_to_f32(u, false, out); }
__device__ __forceinline__ void f16x8_to_f32(const uint4 u, float* out) { u16x8_to_f32(u, true, out); }
__device__ __forceinline__ uint4 f32x8_to_bf16(const float* y) { return f32x8_to_u16(y, false); }
// [h > 0] for 8 consecutive 16-bit elements of either format (sign bit clear and magnitude bits non-zero)
__device__ __forceinline__ void u16x8_positive(const uint4 u, bool* pos) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int t = 0; t < 4; t++) {
    pos[2 * t] = (w[t] & 0x8000u) == 0u && (w[t] & 0x7FFFu) != 0u;
    pos[2 * t + 1] = (w[t] & 0x80000000u) == 0u && (w[t] & 0x7FFF0000u) != 0u;
  }
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// Split accumulator load: issue now, wait later (the registers are tied to the wait so nothing reads them early).
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16_wait(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
// Branch-free epilogue forms (results are consumed at BF16 precision):
//   s(h)  = 1 - exp(-beta h)  = 1 - 2^(ksg h)             (saturates to 1 by itself; h >= 0)
//   sp(v) = max(log2(1 + 2^(min(kz v, 43))) * kinv, v)    (softplus >= v, and equals v in FP32 beyond the clamp)
__device__ __forceinline__ float sg_fast(float h, float ksg) { return 1.f - ex2_approx(h * ksg); }
// the same in units of z = kz v (the forward chain stages kz * bias, so z is one FMA off the accumulator and the two
// scale factors fold into one multiply after the max: 7 instructions per element instead of 9)
__device__ __forceinline__ float sp_z(float z) { return fmaxf(lg2_approx(1.f + ex2_approx(fminf(z, 43.f))), z); }
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

// number of 64-column blocks a step's epilogue walks (every role derives it the same way)
__device__ __forceinline__ int sdf_step_blocks(const SdfStep& S) {
  return S.mode == SC_G0 ? 1 : (S.mode == SC_OUT ? (S.N + 63) >> 6 : 4);
}

// The kernel body for CTA `cta` of `nctas` working on chain `g` (g lives in the kernel parameter space).
template <int FAM>
__device__ __forceinline__ void chain_body(const SdfChainArgs& g, const int cta, const int nctas) {
  constexpr int OPB = FAM == FAM_RELU ? SC_OPB_RELU : 4;
  constexpr bool PARK = FAM == FAM_SDF_FWD;
  constexpr bool FWD = FAM == FAM_SDF_FWD, BWD = FAM == FAM_SDF_BWD;
  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* base = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared
                                                 // address space (an integer round trip turned every access into a generic LD.E / ST.E)
  uint8_t* sOp = base;
  uint8_t* sW0 = base + OPB * TC_A_BYTES;
  constexpr int NSLOT = sc_nslot<FAM>(), WST = sc_wstages<FAM>();
  uint8_t* sAux = sW0 + WST * CH_WBYTES;                         // slot i: h block at 2i, q block at 2i+1 (16 KB each)
  constexpr int NBIAS = sc_bias_slots<FAM>();
  float* sbias = reinterpret_cast<float*>(sAux + NSLOT * 2 * TC_A_BYTES);  // [NBIAS][256]
  float* srvec = sbias + NBIAS * 256;                            // [256]
  float* sdot = srvec + 256;                                     // [4][128] partial row-dots of the four column groups (forward family)
  float* spark = sdot + (FWD ? 512 : 128);                       // [128][SC_PARK_LD]: skip part of the input gradient
  SCSmem* ctl = reinterpret_cast<SCSmem*>(spark + (PARK ? 128 * SC_PARK_LD : 0));

  constexpr bool SDF = FAM != FAM_RELU;
  // format of this kernel's own operands, weight images and stored images (compile-time for the SDF families)
  const bool opf16 = FWD ? true : (BWD ? false : g.f16 != 0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long ntiles = (g.M + 127) / 128;
  const float rsqrt2 = 0.70710678118654752440f;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < WST; s++) { mbar_init(&ctl->wfull[s], 1); mbar_init(&ctl->wempty[s], 1); }
#pragma unroll
    for (int i = 0; i < SC_NAR; i++) mbar_init(&ctl->a_ready[i], SC_EPI_THREADS / 32);
    mbar_init(&ctl->acc_full, 1);
    mbar_init(&ctl->acc_half0, 1);
    mbar_init(&ctl->op_free, 1);
    mbar_init(&ctl->st_sync, 1);
#pragma unroll
    for (int i = 0; i < NSLOT; i++) {
      mbar_init(&ctl->aux_full[i], 1);
      mbar_init(&ctl->aux_empty[i], 1);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) mbar_init(&ctl->blk_done[i], SC_EPI_THREADS / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int s = 0; s < g.nsteps; s++) {
    const SdfStep& S = g.st[s];
    if (S.bias != nullptr && S.bias_slot >= 0) {
      // (softplus steps of the forward chain keep kz * bias: see sp_z)
      const float bsc = (FWD && S.mode == SC_SOFTPLUS) ? g.beta * 1.4426950408889634f : 1.f;
      for (int c = tid; c < 256; c += SC_THREADS) sbias[S.bias_slot * 256 + c] = c < S.N ? __ldg(S.bias + c) * bsc : 0.f;
    }
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
      for (long long tile = cta; tile < ntiles; tile += nctas) {
        for (int s = 0; s < g.nsteps; s++) {
          const SdfStep& S = g.st[s];
          const int Nc = (S.N + 15) & ~15;
          for (int kb = 0; kb < S.KB; kb++) {
            for (int h = 0; h * 128 < Nc; h++, kbg++) {
              const int rows = min(128, Nc - h * 128);
              const uint32_t bytes = S.bmn ? (uint32_t)((rows + 63) >> 6) * 8192u : (uint32_t)rows * 128u;
              const int stg = kbg % WST;
              if (kbg >= WST) mbar_wait(&ctl->wempty[stg], ((kbg / WST) - 1) & 1);
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
      for (long long tile = cta; tile < ntiles; tile += nctas) {
        for (int s = 0; s < g.nsteps; s++, lg++) {
          const SdfStep& S = g.st[s];
          const int Nc = (S.N + 15) & ~15;
          const uint32_t acc = tmem_base + (uint32_t)((lg & 1) << 8);
          for (int kb = 0; kb < S.KB; kb++) {
            // operand block kb is ready as soon as the previous step's epilogue has written it: the MMAs of this
            // step trail that epilogue block by block (its accumulator is the other TMEM buffer)
            mbar_wait(&ctl->a_ready[kb], lg & 1);
            tc_fence_after();
            for (int h = 0; h * 128 < Nc; h++, kbg++) {
              const int rows = min(128, Nc - h * 128);
              const uint32_t idesc = make_idesc(rows, 0, S.bmn, opf16, opf16);
              const int stg = kbg % WST;
              mbar_wait(&ctl->wfull[stg], (kbg / WST) & 1);
              tc_fence_after();
              const uint32_t b_addr = smem_u32(sW0 + stg * CH_WBYTES);
              const uint32_t a_addr = (SDF || kb < S.kb_op) ? smem_u32(sOp) + kb * TC_A_BYTES
                                                            : smem_u32(sAux) + (S.aux_blk0 + kb - S.kb_op) * TC_A_BYTES;
#pragma unroll
              for (int k = 0; k < 4; k++) {
                const uint64_t bd = S.bmn ? make_desc(b_addr + k * 2048, 8192, 1024) : make_desc(b_addr + k * 32, 16, 1024);
                umma_bf16(acc + h * 128, make_desc(a_addr + k * 32, 16, 1024), bd, idesc, (kb > 0 || k > 0) ? 1 : 0);
              }
              umma_commit(&ctl->wempty[stg]);
              // output columns 0..127 are complete once the last operand block's first half-tile is done: the epilogue
              // starts on them while the second half-tile's MMAs (the exposed tail of the step) are still running
              if (kb == S.KB - 1 && h == 0) umma_commit(&ctl->acc_half0);
            }
          }
          for (int kb = S.KB; kb < SC_NAR; kb++) mbar_wait(&ctl->a_ready[kb], lg & 1);   // keep the phases in step
          umma_commit(&ctl->acc_full);
        }
      }
      tc_fence_before();
    }
  } else if (warp == 2) {
    // ------------------------------ auxiliary-block loader ------------------------------
    if (lane == 0) {
      int c = 0, nsync = 0;                         // c: global block counter (slot = c % NSLOT)
      for (long long tile = cta; tile < ntiles; tile += nctas) {
        for (int s = 0; s < g.nsteps; s++) {
          const SdfStep& S = g.st[s];
          const int nb = sdf_step_blocks(S);
          const bool has_h = S.h != nullptr, has_q = S.q != nullptr;
          if (S.wait_sync) { mbar_wait(&ctl->st_sync, nsync & 1); nsync++; }
          for (int b = 0; b < nb; b++, c++) {
            const int slot = c % NSLOT;
            if (c >= NSLOT) mbar_wait(&ctl->aux_empty[slot], ((c / NSLOT) - 1) & 1);
            const bool ld = (has_h || has_q) && b * 64 < S.N;
            if (ld) {
              mbar_arrive_expect_tx(&ctl->aux_full[slot], (uint32_t)((has_h ? 1 : 0) + (has_q ? 1 : 0)) * TC_A_BYTES);
              const size_t off = ((size_t)tile * 4 + b) * TC_A_BYTES;
              if (has_h) bulk_g2s(sAux + (2 * slot) * TC_A_BYTES, reinterpret_cast<const uint8_t*>(S.h) + off, TC_A_BYTES, &ctl->aux_full[slot]);
              if (has_q) bulk_g2s(sAux + (2 * slot + 1) * TC_A_BYTES, reinterpret_cast<const uint8_t*>(S.q) + off, TC_A_BYTES, &ctl->aux_full[slot]);
            } else {
              mbar_arrive(&ctl->aux_full[slot]);
            }
          }
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------ storer ------------------------------
    if (lane == 0) {
      int c = 0;
      for (long long tile = cta; tile < ntiles; tile += nctas) {
        for (int s = 0; s < g.nsteps; s++) {
          const SdfStep& S = g.st[s];
          const int nb = sdf_step_blocks(S);
          for (int b = 0; b < nb; b++, c++) {
            const int slot = c % NSLOT;
            // the storer consumes the loader's phase of the slot too (whether the epilogue needed the slot or not): loader
            // and storer then never run more than NSLOT blocks apart, whatever the epilogue does
            mbar_wait(&ctl->aux_full[slot], (c / NSLOT) & 1);
            mbar_wait(&ctl->blk_done[c & 3], (c >> 2) & 1);
            const size_t off = ((size_t)tile * 4 + b) * TC_A_BYTES;
            bool any = false;
            if (S.img_out != nullptr) { bulk_s2g(reinterpret_cast<uint8_t*>(S.img_out) + off, sOp + b * TC_A_BYTES, TC_A_BYTES); any = true; }
            if (S.e_out != nullptr) { bulk_s2g(reinterpret_cast<uint8_t*>(S.e_out) + off, sAux + (2 * slot) * TC_A_BYTES, TC_A_BYTES); any = true; }
            if (any) { bulk_commit(); bulk_wait_read0(); }
            mbar_arrive(&ctl->aux_empty[slot]);
          }
          if (S.sync_stores) { bulk_wait0(); mbar_arrive(&ctl->st_sync); }
          mbar_arrive(&ctl->op_free);
        }
      }
      bulk_wait0();
    }
  } else {
    // ------------------------------ operand builders / epilogue ------------------------------
    const int et = tid - 128;                      // 0..511
    const int ew = et >> 5;                        // epilogue warp 0..15
    const int cg = ew >> 2;                        // 16-column group inside a 64-column block
    const int quarter = warp & 3;                  // TMEM lane quarter this warp may access
    const int r = quarter * 32 + lane;             // row within the tile == TMEM lane
    const int r7 = r & 7;
    const int rowoff = (r >> 3) * 1024 + r7 * 128; // byte offset of the row inside a 128 x 64 BF16 block
    const int ch0 = (((2 * cg) ^ r7) & 7) << 4, ch1 = (((2 * cg + 1) ^ r7) & 7) << 4;   // the thread's two 16-byte chunks
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float beta = g.beta, inv_beta = 1.f / g.beta;
    int lg = 0, c = 0, nfree = 0;                  // nfree: op_free phases consumed
    int slot = 0;                                  // c % NSLOT and (c / NSLOT) & 1, kept incrementally
    uint32_t slot_par = 0;
    constexpr uint32_t whint = 0x989680u;              // suspend-time hint of the mbarrier waits
    constexpr bool all_arrive = false;                 // one elected arrival per warp (the per-thread variant was slower)
    bool a0_pending = false;                       // bulk store of the first operand's image still reading shared memory
#ifdef FNEUS_SC_TIMELINE
    const bool dbg_on = g.dbg != 0 && cta == 0 && et == 0;
    int dbg_i = 0;
#endif
    for (long long tile = cta; tile < ntiles; tile += nctas) {
      const long long m = tile * 128 + r;
      const bool valid = m < g.M;
      for (int s = 0; s < g.nsteps; s++, lg++) {
        const SdfStep& S = g.st[s];
        // The operand blocks of the previous step are being stored: nothing may overwrite them before the storer has
        // read them (one op_free phase per step).
        const bool first_overall = lg == 0;
        // ---- operand for this step, when it does not come from the previous step's epilogue ----
        if (S.src != SRC_CHAIN) {
          if (!first_overall) { mbar_wait_hint(&ctl->op_free, nfree & 1, whint); nfree++; }
          if (a0_pending) {
            if (et == 0) bulk_wait_read0();
            epi_bar();
            a0_pending = false;
          }
          if (SDF && S.src == (FWD ? SRC_PE : SRC_TAN)) {
            // the four threads of a row clear it, then split its components (sincosf is the cost of this block)
            uint8_t* rowA = sOp + rowoff;
            *reinterpret_cast<uint4*>(rowA + ch0) = make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(rowA + ch1) = make_uint4(0u, 0u, 0u, 0u);
            epi_bar();
            if (valid)
              gen_row_part(FWD ? g.gen : g.gen_t, m, cg, 4, [&](int j, float val) {
                if (j < TC_BK)
                  *reinterpret_cast<unsigned short*>(rowA + ((((j >> 3) ^ r7) & 7) << 4) + ((j & 7) << 1)) = f32_to_u16_bits(val, opf16);
              });
            if (g.pe_img != nullptr) {
              epi_bar();
              uint8_t* dst = reinterpret_cast<uint8_t*>(g.pe_img) + (size_t)tile * TC_A_BYTES + rowoff;
              *reinterpret_cast<uint4*>(dst + ch0) = *reinterpret_cast<const uint4*>(rowA + ch0);
              *reinterpret_cast<uint4*>(dst + ch1) = *reinterpret_cast<const uint4*>(rowA + ch1);
            }
          } else if (!SDF && S.src == SRC_AUXGEN) {
            // generated blocks into the auxiliary area (they stay there for the whole chain of this tile); the previous
            // tile's MMAs that read them are complete (its last epilogue has passed acc_full)
#pragma unroll 1
            for (int blk = 0; blk < g.aux_gen_blocks; blk++) {
              uint8_t* rowA = sAux + blk * TC_A_BYTES + rowoff;
              *reinterpret_cast<uint4*>(rowA + ch0) = make_uint4(0u, 0u, 0u, 0u);
              *reinterpret_cast<uint4*>(rowA + ch1) = make_uint4(0u, 0u, 0u, 0u);
            }
            epi_bar();
            if (valid)
              gen_row_part(g.gen, m, cg, 4, [&](int j, float val) {
                if (j < g.aux_gen_blocks * TC_BK)
                  *reinterpret_cast<unsigned short*>(sAux + (j >> 6) * TC_A_BYTES + rowoff + (((((j & 63) >> 3) ^ r7) & 7) << 4) +
                                                     ((j & 7) << 1)) = f32_to_u16_bits(val, opf16);
              });
          } else if (!FWD) {
            uint8_t* sMem = sOp;                                           // first block of the memory segment
            if (!SDF && S.src == SRC_GENMEM) {
              // [one generated block | memory blocks]: the first operand of a ReLU network (fields.py:157,324-331)
              sMem = sOp + TC_A_BYTES;
              uint8_t* rowA = sOp + rowoff;
              *reinterpret_cast<uint4*>(rowA + ch0) = make_uint4(0u, 0u, 0u, 0u);
              *reinterpret_cast<uint4*>(rowA + ch1) = make_uint4(0u, 0u, 0u, 0u);
              epi_bar();
              if (valid)
                gen_row_part(g.gen, m, cg, 4, [&](int j, float val) {
                  if (j < TC_BK)
                    *reinterpret_cast<unsigned short*>(rowA + ((((j >> 3) ^ r7) & 7) << 4) + ((j & 7) << 1)) = f32_to_u16_bits(val, opf16);
                });
            }
            const int kb_mem = (g.kmem + TC_BK - 1) / TC_BK;
            if (g.ldm < 0) {
              // image source: the tile's bytes are the operand blocks
              const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(g.mem) +
                                                                (size_t)tile * (size_t)(-g.ldm) * TC_A_BYTES);
              const int nvec = kb_mem * (TC_A_BYTES / 16);
              // four independent 16-byte loads in flight per thread (two 64-column blocks per round)
              for (int i0 = et; i0 < nvec; i0 += 4 * SC_EPI_THREADS) {
                uint4 v[4];
#pragma unroll
                for (int k = 0; k < 4; k++) v[k] = i0 + k * SC_EPI_THREADS < nvec ? __ldg(src + i0 + k * SC_EPI_THREADS) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                for (int k = 0; k < 4; k++)
                  if (i0 + k * SC_EPI_THREADS < nvec) reinterpret_cast<uint4*>(sMem)[i0 + k * SC_EPI_THREADS] = v[k];
              }
            }
            // FP32 row-major source: each warp converts 8 rows, a lane 8 consecutive columns at a time (whole lines)
            const bool vec_ok = (g.ldm & 3) == 0 && (reinterpret_cast<uintptr_t>(g.mem) & 15) == 0;
            const int nch = g.ldm < 0 ? 0 : kb_mem * 8;
            // all 8 rows of a warp are loaded before any is converted (16 independent 16-byte loads in flight per lane)
            for (int ch = lane; ch < nch; ch += 32) {
              const int cc = ch * 8;
              const bool fastc = vec_ok && cc + 8 <= g.kmem;
#pragma unroll 1
              for (int i0 = 0; i0 < 8; i0 += 4) {
              float4 lo[4], hi[4];
#pragma unroll
              for (int i = 0; i < 4; i++) {
                const long long mm = tile * 128 + ew * 8 + i0 + i;
                const float* src = g.mem + mm * g.ldm + cc;
                if (mm < g.M && fastc) {
                  lo[i] = __ldg(reinterpret_cast<const float4*>(src));
                  hi[i] = __ldg(reinterpret_cast<const float4*>(src + 4));
                } else {
                  float v[8];
#pragma unroll
                  for (int j = 0; j < 8; j++) v[j] = (mm < g.M && cc + j < g.kmem) ? __ldg(src + j) : 0.f;
                  lo[i] = make_float4(v[0], v[1], v[2], v[3]);
                  hi[i] = make_float4(v[4], v[5], v[6], v[7]);
                }
              }
#pragma unroll
              for (int i = 0; i < 4; i++) {
                const int rr = ew * 8 + i0 + i;
                const float v[8] = {lo[i].x, lo[i].y, lo[i].z, lo[i].w, hi[i].x, hi[i].y, hi[i].z, hi[i].w};
                uint8_t* drow = sMem + (rr >> 3) * 1024 + (rr & 7) * 128;
                *reinterpret_cast<uint4*>(drow + (ch >> 3) * TC_A_BYTES + ((((ch & 7) ^ (rr & 7)) & 7) << 4)) = f32x8_to_u16(v, opf16);
              }
              }
            }
          }
          tc_fence_before();
          fence_proxy_async();
          if (!SDF && S.src == SRC_AUXGEN && g.a0_img != nullptr) {
            // image copy of the generated blocks (the weight gradients of the layers that consume them read it back)
            epi_bar();
            if (et == 0) {
              bulk_s2g(reinterpret_cast<uint8_t*>(g.a0_img) + (size_t)tile * g.aux_gen_blocks * TC_A_BYTES, sAux,
                       (uint32_t)g.aux_gen_blocks * TC_A_BYTES);
              bulk_commit();
            }
            a0_pending = true;
          } else if (!SDF && (S.src == SRC_GENMEM || S.src == SRC_MEM) && g.a0_img != nullptr) {
            // image copy of the whole first operand (the layer-0 weight gradient reads it back): one bulk store; its
            // shared-memory reads are over before anybody overwrites the operand (barrier below)
            epi_bar();
            if (et == 0) {
              const int kb0 = (S.src == SRC_GENMEM ? 1 : 0) + (g.kmem + TC_BK - 1) / TC_BK;
              bulk_s2g(reinterpret_cast<uint8_t*>(g.a0_img) + (size_t)tile * kb0 * TC_A_BYTES, sOp, (uint32_t)kb0 * TC_A_BYTES);
              bulk_commit();
            }
            a0_pending = true;
          }
          __syncwarp();
          if (lane == 0 || all_arrive) {
#pragma unroll
            for (int i = 0; i < SC_NAR; i++) mbar_arrive(&ctl->a_ready[i]);
          }
        }

        const int N = S.N, Nc = (N + 15) & ~15, mode = S.mode;
        const int nb = sdf_step_blocks(S);
        const float* sb = sbias + (S.bias_slot >= 0 ? S.bias_slot : 0) * 256;
        const bool hasb = S.bias_slot >= 0;
        const bool s_append = SDF && S.append != 0, s_dot = FWD && S.dot != 0;
        const int csplit = S.csplit;
        // slots are only waited for when the step reads auxiliary blocks or stages through the slot
        // slots are only waited for when the step reads auxiliary blocks or stages through the slot.  The epilogue of a
        // step that needs no slot runs up to four blocks (one step: op_free) ahead of the storer, so "block done" is
        // signalled on FOUR barriers indexed by the block counter: with one barrier per slot (two slots) the epilogue
        // could complete TWO phases of the same barrier while the storer still waited for the first -- a parity wait
        // cannot tell those apart and never returns (seen on cold first launches of the ReLU forward chain, where the first
        // bulk stores take microseconds; diagnosed with the wait watchdog, gemm_tc.cuh mbar_hang).
        const bool uses_slot = S.h != nullptr || S.q != nullptr || mode == SC_FEATQ || mode == SC_G0 || mode == SC_OUT;
        const float oscale = S.oscale;
        const float ksg = -beta * S.hscale * 1.4426950408889634f;       // s(h) = 1 - 2^(ksg h)
        const float kz = beta * 1.4426950408889634f, kinv = 0.6931471805599453f * inv_beta, ko = kinv * oscale;
        const int lim = (mode == SC_SPMUL || mode == SC_SDFBWD) ? csplit : N;
        const float rsv = ((mode == SC_SDFBWD || mode == SC_MASK) && S.use_rs && valid) ? __ldg(g.rs + m) * g.rscale : 0.f;
        float dot = 0.f;
        const uint32_t tacc = taddr + (uint32_t)((lg & 1) << 8);         // this step's accumulator buffer
        const bool feeds_next = s + 1 < g.nsteps && g.st[s + 1].src == SRC_CHAIN;
        const bool pub_blocks = feeds_next && mode != SC_FEATQ && mode != SC_OUT;   // publish the next operand block by block

        SC_STAMP(1);
        mbar_wait_hint(&ctl->acc_half0, lg & 1, whint);
        tc_fence_after();
        bool full_done = false;                                          // acc_full of this step already waited for
        auto need_full = [&]() {
          if (!full_done) {
            mbar_wait_hint(&ctl->acc_full, lg & 1, whint);
            tc_fence_after();
            full_done = true;
          }
        };
        // every mode but the in-place ones touches shared state that the step's last MMAs may still read or write
        if (mode == SC_FEATQ || mode == SC_G0 || mode == SC_OUT) need_full();
        SC_STAMP(2);
        if (S.src == SRC_CHAIN && !first_overall) { mbar_wait_hint(&ctl->op_free, nfree & 1, whint); nfree++; }
        if (a0_pending && mode != SC_OUT) {
          if (et == 0) bulk_wait_read0();
          epi_bar();
          a0_pending = false;
        }
        // an OUT step leaves the operand as it is: the next step's MMAs may start at once
        if (mode == SC_OUT && feeds_next && (lane == 0 || all_arrive)) {
#pragma unroll
          for (int i = 0; i < SC_NAR; i++) mbar_arrive(&ctl->a_ready[i]);
        }
        SC_STAMP(3);

        // Modes that end in the common pack-and-store: the next block's accumulator load is issued before the store /
        // fence / arrive tail of this block, so the TMEM latency hides underneath it.
        const bool can_pf = (mode == SC_SOFTPLUS || mode == SC_SPMUL || mode == SC_SWEEP || mode == SC_SDFBWD ||
                             mode == SC_RELU || mode == SC_MASK);
        uint32_t ar[16];
        bool pf = false;                                                   // ar[] is an in-flight load of this block
        // The per-block loop, instantiated per mode: MODEC >= 0 fixes the step's mode at compile time; FAST is a whole tile
        // with a 256-wide result and no skip / row-dot extras (four unmasked blocks) -- the bulk of the forward chains,
        // where the epilogue's instruction issue is the bound (7% on the SDF forward).  MODEC < 0 is the general loop
        // (runtime mode, ragged tiles, narrow or staged outputs).  The loop stays rolled: unrolling it by four buys nothing
        // (0.679 ms either way on the SDF forward).
        auto run_blocks = [&](auto mode_c, auto fast_c) {
          constexpr int MODEC = decltype(mode_c)::value;
          constexpr bool FAST = decltype(fast_c)::value;
          const int md = MODEC >= 0 ? MODEC : mode;
          const bool pf_ok = MODEC >= 0 ? true : can_pf;
          const bool slot_wait = (MODEC == SC_SOFTPLUS || MODEC == SC_RELU) ? false : uses_slot;
          const int nbl = FAST ? 4 : nb;
#pragma unroll 1
          for (int b = 0; b < nbl; b++, c++, slot = slot + 1 == NSLOT ? 0 : slot + 1, slot_par ^= (slot == 0 ? 1u : 0u)) {
            const int n = b * 64 + cg * 16;
            uint8_t* opb = sOp + b * TC_A_BYTES + rowoff;                    // the row inside operand block b
            uint8_t* hb = sAux + (2 * slot) * TC_A_BYTES + rowoff;           // ... inside the slot's h block
            const uint8_t* qb = sAux + (2 * slot + 1) * TC_A_BYTES + rowoff; // ... inside the slot's q block
            SC_STAMP(4);
            // columns >= 128, and the operand block the step's last half-tile of MMAs is still reading (block KB-1)
            if (b >= 2 || b == S.KB - 1) need_full();
            if (slot_wait) mbar_wait_hint(&ctl->aux_full[slot], slot_par, whint);
            SC_STAMP(5);
            const bool full = FAST ? true : (valid && n + 16 <= lim);                        // no per-element masks needed
            float a[16];
            if (!pf) {
              if (FAST || n < Nc) tmem_ld16_issue(tacc + n, ar);
              else {
#pragma unroll
                for (int j = 0; j < 16; j++) ar[j] = 0u;
              }
            }
            tmem_ld16_wait(ar);
            SC_STAMP(8);
#pragma unroll
            for (int j = 0; j < 16; j++) a[j] = __uint_as_float(ar[j]);
            pf = false;
            if (FWD && md == SC_SOFTPLUS) {
              if (s_dot) {
#pragma unroll
                for (int j = 0; j < 16; j++) {
                  float v = sp_z(fmaf(a[j], kz, sb[n + j])) * kinv;
                  v = (valid && n + j < N) ? v : 0.f;
                  dot = fmaf(v, srvec[n + j], dot);
                  a[j] = v * oscale;
                }
              } else if (full) {
#pragma unroll
                for (int j = 0; j < 16; j++) a[j] = sp_z(fmaf(a[j], kz, sb[n + j])) * ko;
              } else {
#pragma unroll
                for (int j = 0; j < 16; j++) a[j] = (valid && n + j < N) ? sp_z(fmaf(a[j], kz, sb[n + j])) * ko : 0.f;
              }
            } else if (FWD && md == SC_FEATQ) {
              if (S.out == nullptr) {
                // image hand-over (fneus_sdf_cfg.feat_image): the features leave as one more FP16 operand block, from the
                // slot's h block (the storer sends it to S.e_out) -- the colour chain's first operand reads it as is
                float f[16];
#pragma unroll
                for (int j = 0; j < 16; j++) f[j] = (valid && n + j < N) ? a[j] + sb[n + j] : 0.f;
                *reinterpret_cast<uint4*>(hb + ch0) = f32x8_to_u16(f, true);
                *reinterpret_cast<uint4*>(hb + ch1) = f32x8_to_u16(f + 8, true);
              } else {
                // features: FP32 through a swizzled [128][64] staging tile (the slot's 32 KB), written out as whole lines
                float* T = reinterpret_cast<float*>(sAux + (2 * slot) * TC_A_BYTES);
#pragma unroll
                for (int i = 0; i < 4; i++)
                  *reinterpret_cast<float4*>(T + r * 64 + (((cg * 4 + i) ^ (r & 15)) << 2)) =
                      make_float4(a[4 * i] + sb[n + 4 * i], a[4 * i + 1] + sb[n + 4 * i + 1], a[4 * i + 2] + sb[n + 4 * i + 2],
                                  a[4 * i + 3] + sb[n + 4 * i + 3]);
              }
              // q_{L-1} = s(h_L) * W_L[0] from the unrounded activation: the previous step's pre-activations still sit
              // in the other accumulator (this step publishes its operand only when all blocks are done)
              const float* sbp = sbias + (s > 0 && g.st[s - 1].bias_slot >= 0 ? g.st[s - 1].bias_slot : 0) * 256;
              tmem_ld16(taddr + (uint32_t)(((lg + 1) & 1) << 8) + n, a);
#pragma unroll
              for (int j = 0; j < 16; j++)
                a[j] = (valid && n + j < N) ? sg_fast(sp_z(fmaf(a[j], kz, sbp[n + j])) * kinv, ksg) * srvec[n + j] : 0.f;
            } else if (FWD && md == SC_SPMUL) {
              if (!FAST && csplit < N && n + 16 > csplit) {
                // positional-encoding part of the skip gradient: parked in shared memory until G0
#pragma unroll
                for (int j = 0; j < 16; j++)
                  if (PARK && n + j >= csplit && n + j < N && n + j - csplit < SC_PARK_LD)
                    spark[r * SC_PARK_LD + n + j - csplit] = a[j] * oscale;
              }
#pragma unroll
              for (int hf = 0; hf < 2; hf++) {
                float hv[8];
                f16x8_to_f32(*reinterpret_cast<const uint4*>(hb + (hf ? ch1 : ch0)), hv);      // h_l: forward image, FP16
                if (full) {
#pragma unroll
                  for (int j = 0; j < 8; j++) {                                    // s(h) a oscale = ao - 2^(ksg h) ao
                    const float ao = a[hf * 8 + j] * oscale;
                    a[hf * 8 + j] = fmaf(-ex2_approx(hv[j] * ksg), ao, ao);
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 8; j++)
                    a[hf * 8 + j] = (valid && n + hf * 8 + j < csplit) ? sg_fast(hv[j], ksg) * a[hf * 8 + j] * oscale : 0.f;
                }
              }
            } else if (FWD && md == SC_G0) {
              // g_0 = q_0 W_0 (+ the parked skip part): 128 x 64 FP32 staging, then one thread per row forms the normal
              float* T = reinterpret_cast<float*>(sAux + (2 * slot) * TC_A_BYTES);
              if (csplit > 0) {
                // parked columns csplit .. csplit+N-1 of the skip step <-> g_0 columns 0 .. N-1 (written by other
                // column groups of the row: the named barrier orders the shared-memory accesses)
                epi_bar();
#pragma unroll
                for (int j = 0; j < 16; j++)
                  if (PARK && n + j < N && n + j < SC_PARK_LD && csplit + n + j < 256) a[j] += spark[r * SC_PARK_LD + n + j];
              }
#pragma unroll
              for (int i = 0; i < 4; i++)
                *reinterpret_cast<float4*>(T + r * 64 + (((cg * 4 + i) ^ (r & 15)) << 2)) =
                    make_float4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]);
            } else if (BWD && md == SC_SWEEP) {
#pragma unroll
              for (int hf = 0; hf < 2; hf++) {
                float hv[8], qv[8], e[8];
                f16x8_to_f32(*reinterpret_cast<const uint4*>(hb + (hf ? ch1 : ch0)), hv);      // h_{l+1}, q_l: forward
                f16x8_to_f32(*reinterpret_cast<const uint4*>(qb + (hf ? ch1 : ch0)), qv);      // images, FP16
#pragma unroll
                for (int j = 0; j < 8; j++) {
                  const float sg = sg_fast(hv[j], ksg);
                  const float acc = a[hf * 8 + j];
                  a[hf * 8 + j] = sg * acc * oscale;
                  e[j] = fmaf(-beta, sg, beta) * qv[j] * acc;
                }
                if (!full) {                                                       // ragged tile / partial block only
#pragma unroll
                  for (int j = 0; j < 8; j++) {
                    const bool ok = valid && n + hf * 8 + j < N;
                    a[hf * 8 + j] = ok ? a[hf * 8 + j] : 0.f;
                    e[j] = ok ? e[j] : 0.f;
                  }
                }
                *reinterpret_cast<uint4*>(hb + (hf ? ch1 : ch0)) = f32x8_to_bf16(e);   // e_l leaves from where h_{l+1} arrived
              }
            } else if (!SDF && md == SC_RELU) {
              const float floor_v = S.act == 2 ? -3.0e38f : 0.f;            // act 2: plain linear layer
              if (full) {
#pragma unroll
                for (int j = 0; j < 16; j++) a[j] = fmaxf(a[j] + sb[n + j], floor_v);
              } else {
#pragma unroll
                for (int j = 0; j < 16; j++) a[j] = (valid && n + j < N) ? fmaxf(a[j] + sb[n + j], floor_v) : 0.f;
              }
            } else if (!SDF && md == SC_MASK) {
              // dz_{l-1} = (dz_l W_l [+ rs rvec]) * [h_l > 0]: the forward activation block arrived in the slot's h block
              const bool has_h = S.h != nullptr;
#pragma unroll
              for (int hf = 0; hf < 2; hf++) {
                if (S.use_rs) {
#pragma unroll
                  for (int j = 0; j < 8; j++) a[hf * 8 + j] = fmaf(rsv, srvec[n + hf * 8 + j], a[hf * 8 + j]);
                }
                if (has_h) {
                  // [h > 0] straight off the 16-bit pattern moved to the top of an FP32 word: sign and zero-ness survive for
                  // FP16 and BF16 alike (an FP16 pattern read this way is a tiny, possibly denormal, float of the same sign)
                  float hv[8];
                  bf16x8_to_f32(*reinterpret_cast<const uint4*>(hb + (hf ? ch1 : ch0)), hv);
#pragma unroll
                  for (int j = 0; j < 8; j++) a[hf * 8 + j] = hv[j] > 0.f ? a[hf * 8 + j] : 0.f;
                }
                if (!full) {
#pragma unroll
                  for (int j = 0; j < 8; j++) a[hf * 8 + j] = (valid && n + hf * 8 + j < N) ? a[hf * 8 + j] : 0.f;
                }
              }
            } else if (!SDF && md == SC_OUT) {
              if (N <= 16) {
                // narrow result (colours, a specular scalar): the row's thread writes it directly
                if (cg == 0 && b == 0 && valid) {
                  float y[16];
                  if (S.accumulate) row_load16(S.out, S.ldo, m, 0, N, y);
#pragma unroll
                  for (int j = 0; j < 16; j++) {
                    float v = a[j] + (hasb ? sb[j] : 0.f);
                    if (S.act == 1) v = sigmoid_fast(v);
                    y[j] = S.accumulate ? y[j] + v : v;
                  }
                  row_store16(S.out, S.ldo, m, 0, N, y);
                }
              } else if (S.out == nullptr) {
                // wide result handed over as an operand image (the feature gradient, fneus_color_cfg.feat_image): packed
                // into the slot's h block, which the storer sends to S.e_out
                float f[16];
#pragma unroll
                for (int j = 0; j < 16; j++) {
                  float v = a[j] + (hasb ? sb[n + j] : 0.f);
                  if (S.act == 1) v = sigmoid_fast(v);
                  f[j] = (valid && n + j < N) ? v : 0.f;
                }
                *reinterpret_cast<uint4*>(hb + ch0) = f32x8_to_u16(f, opf16);
                *reinterpret_cast<uint4*>(hb + ch1) = f32x8_to_u16(f + 8, opf16);
              } else {
                // wide result: FP32 through the slot's swizzled [128][64] staging tile, written out as whole lines
                float* T = reinterpret_cast<float*>(sAux + (2 * slot) * TC_A_BYTES);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                  float4 v = make_float4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]);
                  if (hasb) { v.x += sb[n + 4 * i]; v.y += sb[n + 4 * i + 1]; v.z += sb[n + 4 * i + 2]; v.w += sb[n + 4 * i + 3]; }
                  if (S.act == 1) { v.x = sigmoid_fast(v.x); v.y = sigmoid_fast(v.y); v.z = sigmoid_fast(v.z); v.w = sigmoid_fast(v.w); }
                  *reinterpret_cast<float4*>(T + r * 64 + (((cg * 4 + i) ^ (r & 15)) << 2)) = v;
                }
              }
            } else if (BWD) {  // SC_SDFBWD
#pragma unroll
              for (int hf = 0; hf < 2; hf++) {
                float hv[8], qv[8];
                f16x8_to_f32(*reinterpret_cast<const uint4*>(hb + (hf ? ch1 : ch0)), hv);      // h_l: forward image, FP16
                bf16x8_to_f32(*reinterpret_cast<const uint4*>(qb + (hf ? ch1 : ch0)), qv);     // e_{l-1}: this pass, BF16
                if (S.use_rs) {                                                    // the sdf row of the last linear (rank-1)
#pragma unroll
                  for (int j = 0; j < 8; j++) a[hf * 8 + j] = fmaf(rsv, srvec[n + hf * 8 + j], a[hf * 8 + j]);
                }
#pragma unroll
                for (int j = 0; j < 8; j++) a[hf * 8 + j] = fmaf(sg_fast(hv[j], ksg) * oscale, a[hf * 8 + j], qv[j]);
                if (!full) {
#pragma unroll
                  for (int j = 0; j < 8; j++) a[hf * 8 + j] = (valid && n + hf * 8 + j < csplit) ? a[hf * 8 + j] : 0.f;
                }
              }
            }
            SC_STAMP(9);
            if (md != SC_G0 && md != SC_OUT) {
              const uint4 p0 = f32x8_to_u16(a, opf16), p1 = f32x8_to_u16(a + 8, opf16);
              if (pf_ok && b + 1 < nbl && (FAST || n + 64 < Nc)) {
                if (b + 1 >= 2) need_full();
                tmem_ld16_issue(tacc + n + 64, ar);
                pf = true;
              }
              *reinterpret_cast<uint4*>(opb + ch0) = p0;
              *reinterpret_cast<uint4*>(opb + ch1) = p1;
            }
            if (!FAST && s_append && b == 3) {
              // skip connection (fields.py:83-84): PE columns (value or tangent form) follow the N hidden columns once
              // every column group has written its chunk of the last block
              epi_bar();
              if (valid)
                gen_row_part(BWD ? g.gen_t : g.gen, m, cg, 4, [&](int j, float val) {
                  const int col = N + j;
                  if (col < 256)
                    *reinterpret_cast<unsigned short*>(sOp + (col >> 6) * TC_A_BYTES + rowoff + (((((col & 63) >> 3) ^ r7) & 7) << 4) +
                                                       ((col & 7) << 1)) = f32_to_u16_bits(val * rsqrt2, opf16);
                });
            }
            if (FWD ? ((md == SC_FEATQ && S.out != nullptr) || md == SC_G0) : (!SDF && md == SC_OUT && N > 16 && S.out != nullptr)) {
              epi_bar();
              const float* T = reinterpret_cast<const float*>(sAux + (2 * slot) * TC_A_BYTES);
              if (!SDF || md == SC_FEATQ) {
                // warp ew writes rows 8 ew .. 8 ew + 7: two rows (2 x 256 B) per instruction
#pragma unroll
                for (int i = 0; i < 4; i++) {
                  const int rr = ew * 8 + 2 * i + (lane >> 4), q4 = lane & 15;
                  const long long mm = tile * 128 + rr;
                  const int col = b * 64 + q4 * 4;
                  if (mm < g.M && col < N) {
                    const float4 v = *reinterpret_cast<const float4*>(T + rr * 64 + ((q4 ^ (rr & 15)) << 2));
                    float* dst = S.out + mm * S.ldo + col;
                    const bool acc_out = md == SC_OUT && S.accumulate;
                    if (col + 4 <= N && (S.ldo & 3) == 0) {
                      if (acc_out) {
                        const float4 o = *reinterpret_cast<const float4*>(dst);
                        *reinterpret_cast<float4*>(dst) = make_float4(o.x + v.x, o.y + v.y, o.z + v.z, o.w + v.w);
                      } else *reinterpret_cast<float4*>(dst) = v;
                    } else {
                      dst[0] = acc_out ? dst[0] + v.x : v.x;
                      if (col + 1 < N) dst[1] = acc_out ? dst[1] + v.y : v.y;
                      if (col + 2 < N) dst[2] = acc_out ? dst[2] + v.z : v.z;
                      if (col + 3 < N) dst[3] = acc_out ? dst[3] + v.w : v.w;
                    }
                  }
                }
              } else if (FWD && et < 128) {
                // thread et = row: normal = J_PE(x)^T g_0 (fields.py:101-111)
                const int rr = et;
                const long long mm = tile * 128 + rr;
                if (mm < g.M) {
                  const GenItem it = g.gen.it[0];
                  for (int cc = 0; cc < it.dim; cc++) {
                    const float v = __ldg(it.src + mm * it.dim + cc) * it.scale;
                    auto G0 = [&](int col) { return T[rr * 64 + ((((col >> 2) ^ (rr & 15)) << 2) | (col & 3))]; };
                    float acc = G0(cc);
                    for (int k = 0; k < it.multires; k++) {
                      const float f = (float)(1u << k);
                      float sn, cs;
                      sincosf(v * f, &sn, &cs);
                      acc += f * cs * G0(it.dim * (1 + 2 * k) + cc) - f * sn * G0(it.dim * (2 + 2 * k) + cc);
                    }
                    S.out[mm * it.dim + cc] = acc;
                  }
                }
              }
            }
            SC_STAMP(10);
            tc_fence_before();
            fence_proxy_async();
            SC_STAMP(11);
            // one arrival per warp: every lane has fenced its own writes, the warp barrier orders them before the arrive
            if (!all_arrive) __syncwarp();
            if (lane == 0 || all_arrive) {
              mbar_arrive(&ctl->blk_done[c & 3]);
              if (pub_blocks) mbar_arrive(&ctl->a_ready[b]);                  // the next step's MMAs may read block b
            }
            SC_STAMP(6);
          }
        };
        {
          using sc_fast = std::true_type;
          const bool fast_ok = tile * 128 + 128 <= g.M && N == 256 && lim == 256 && nb == 4 && !s_dot && !s_append;
          bool done = false;
          if constexpr (FWD) {
            if (fast_ok && mode == SC_SOFTPLUS && !uses_slot) { run_blocks(std::integral_constant<int, SC_SOFTPLUS>{}, sc_fast{}); done = true; }
            else if (fast_ok && mode == SC_SPMUL) { run_blocks(std::integral_constant<int, SC_SPMUL>{}, sc_fast{}); done = true; }
          } else if constexpr (!BWD) {                 // (the backward SDF chain is HBM-bound: nothing to gain there)
            if (fast_ok && mode == SC_RELU && !uses_slot) { run_blocks(std::integral_constant<int, SC_RELU>{}, sc_fast{}); done = true; }
            else if (fast_ok && mode == SC_MASK) { run_blocks(std::integral_constant<int, SC_MASK>{}, sc_fast{}); done = true; }
          }
          if (!done) run_blocks(std::integral_constant<int, -1>{}, std::false_type{});
        }
        need_full();                                                       // keeps the barrier phases in step
        if (feeds_next && mode != SC_OUT && (lane == 0 || all_arrive))
          for (int b = pub_blocks ? nb : 0; b < SC_NAR; b++) mbar_arrive(&ctl->a_ready[b]);
        if (s_dot) {
          // sdf = h_L . W_L[0] + b: the four column groups of a row meet in shared memory
          // (fixed summation order: the value must not depend on which warp arrives first -- inverse-CDF sampling
          // amplifies a last-bit difference of the sdf into visibly different sample depths)
          sdot[cg * 128 + r] = dot;
          epi_bar();
          if (cg == 0 && valid)
            g.sdf_out[m] = ((((sdot[r] + sdot[128 + r]) + sdot[256 + r]) + sdot[384 + r]) + __ldg(g.b_last)) * g.sdf_scale;
          epi_bar();                                                         // the partials are rewritten by the next tile
        }
        SC_STAMP(7);
      }
    }
    if (et == 0) bulk_wait0();
#ifdef FNEUS_SC_TIMELINE
    if (dbg_on) g_sc_dbg[8191] = dbg_i;
#endif
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int FAM>
__global__ void __launch_bounds__(SC_THREADS, 1) sdf_chain_kernel(const __grid_constant__ SdfChainArgs g) {
  chain_body<FAM>(g, blockIdx.x, gridDim.x);
}
// Two independent ReLU chains in one launch (RefColor's diffuse and specular networks, fields.py:303-335, are 8 tiles
// each at 512 rays: alone, either leaves 140 SMs idle): CTAs [0, ctas_a) run chain a, the rest chain b.
struct SdfChainPair { SdfChainArgs a, b; int ctas_a; };
__global__ void __launch_bounds__(SC_THREADS, 1) relu_chain_pair_kernel(const __grid_constant__ SdfChainPair p) {
  const bool first = (int)blockIdx.x < p.ctas_a;
  chain_body<FAM_RELU>(first ? p.a : p.b, first ? blockIdx.x : blockIdx.x - p.ctas_a, first ? p.ctas_a : gridDim.x - p.ctas_a);
}

inline SdfStep sdf_step(int mode, const uint8_t* wimg, int KB, int N, int bmn) {
  SdfStep S;
  memset(&S, 0, sizeof(S));
  S.mode = mode; S.wimg = wimg; S.KB = KB; S.N = N; S.bmn = bmn; S.src = SRC_CHAIN;
  S.kb_op = KB; S.aux_blk0 = 0;
  S.csplit = N; S.bias_slot = -1; S.hscale = 1.f; S.oscale = 1.f; S.ldo = 4;
  return S;
}

inline int sdf_chain_prepare() {
  static int done = 0;
  if (done) return 0;
  if (cudaFuncSetAttribute(sdf_chain_kernel<FAM_SDF_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           sc_smem_bytes<FAM_SDF_FWD>()) != cudaSuccess ||
      cudaFuncSetAttribute(sdf_chain_kernel<FAM_SDF_BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           sc_smem_bytes<FAM_SDF_BWD>()) != cudaSuccess ||
      cudaFuncSetAttribute(sdf_chain_kernel<FAM_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           sc_smem_bytes<FAM_RELU>()) != cudaSuccess ||
      cudaFuncSetAttribute(relu_chain_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           sc_smem_bytes<FAM_RELU>()) != cudaSuccess)
    return 1;
  done = 1;
  return 0;
}
// DRAM bytes one launch of a chain moves BY DESIGN (what bench.py reports next to the ncu figure): per tile and step the
// auxiliary blocks loaded (h, q), the image blocks stored (result, e), FP32 rows read (first operand) and written
// (features, normals, narrow outputs), plus every weight image once (they are re-read from L2 by each tile).
inline double sdf_chain_bytes(const SdfChainArgs& g) {
  const double ntiles = (double)((g.M + 127) / 128);
  double per_tile = 0.0, weights = 0.0;
  for (int s = 0; s < g.nsteps; s++) {
    const SdfStep& S = g.st[s];
    const int nb = S.mode == SC_G0 ? 1 : (S.mode == SC_OUT ? (S.N + 63) >> 6 : 4);
    const int nload = (S.N + 63) >> 6 < nb ? (S.N + 63) >> 6 : nb;
    per_tile += (double)((S.h ? 1 : 0) + (S.q ? 1 : 0)) * nload * TC_A_BYTES;
    per_tile += (double)((S.img_out ? 1 : 0) + (S.e_out ? 1 : 0)) * nb * TC_A_BYTES;
    if (S.out) per_tile += 128.0 * 4.0 * (S.mode == SC_G0 ? 3 : S.N) * (S.accumulate ? 2 : 1);
    if (S.src == SRC_MEM || S.src == SRC_GENMEM)
      per_tile += g.ldm < 0 ? (double)((g.kmem + 63) / 64) * TC_A_BYTES : 128.0 * 4.0 * g.kmem;
    if (S.src != SRC_CHAIN) per_tile += 128.0 * 4.0 * 12;          // raw inputs of the generated columns (<= 4 items x 3)
    weights += (double)S.KB * (double)(((S.N + 15) & ~15) * 128);
  }
  if (g.pe_img) per_tile += TC_A_BYTES;
  if (g.a0_img) per_tile += (double)(g.aux_gen_blocks > 0 ? g.aux_gen_blocks : 1 + (g.kmem + 63) / 64) * TC_A_BYTES;
  if (g.sdf_out) per_tile += 128.0 * 4.0;
  return ntiles * per_tile + weights;
}

inline void sdf_chain_launch(const SdfChainArgs& g, double flops, cudaStream_t st, int family = FAM_SDF_FWD) {
  const long long ntiles = (g.M + 127) / 128;
  const int sms = tc_num_sms();
  const int grid = (int)(ntiles < sms ? ntiles : sms);
  prof_begin(family == FAM_RELU ? PC_CHAIN_RELU : (family == FAM_SDF_BWD ? PC_CHAIN_SDF_BWD : PC_CHAIN_SDF_FWD), flops,
             sdf_chain_bytes(g), st);
  if (family == FAM_RELU) sdf_chain_kernel<FAM_RELU><<<grid, SC_THREADS, sc_smem_bytes<FAM_RELU>(), st>>>(g);
  else if (family == FAM_SDF_BWD) sdf_chain_kernel<FAM_SDF_BWD><<<grid, SC_THREADS, sc_smem_bytes<FAM_SDF_BWD>(), st>>>(g);
  else sdf_chain_kernel<FAM_SDF_FWD><<<grid, SC_THREADS, sc_smem_bytes<FAM_SDF_FWD>(), st>>>(g);
  prof_end(st);
}

inline void relu_chain_pair_launch(const SdfChainArgs& a, const SdfChainArgs& b, double flops, cudaStream_t st) {
  SdfChainPair p;
  p.a = a; p.b = b;
  const int sms = tc_num_sms();
  const long long ta = (a.M + 127) / 128, tb = (b.M + 127) / 128;
  int ca = (int)(ta < sms / 2 ? ta : sms / 2), cb = (int)(tb < sms - ca ? tb : sms - ca);
  p.ctas_a = ca;
  prof_begin(PC_CHAIN_RELU, flops, sdf_chain_bytes(a) + sdf_chain_bytes(b), st);
  relu_chain_pair_kernel<<<ca + cb, SC_THREADS, sc_smem_bytes<FAM_RELU>(), st>>>(p);
  prof_end(st);
}

}  // namespace fneus

// Status decoding and version entry points of the C ABI (include/fneus.h).
#include "gemm_tc.cuh"

namespace fneus { int num_sms(); }

extern "C" {

int fneus_abi_version(void) { return 1; }
int fneus_num_sms(void) { return fneus::num_sms(); }

const char* fneus_status_string(int status) {
  switch (status) {
    case FNEUS_OK: return "ok";
    case FNEUS_ERR_BAD_SHAPE: return "bad shape";
    case FNEUS_ERR_MISALIGNED: return "misaligned pointer";
    case FNEUS_ERR_UNSUPPORTED: return "unsupported configuration";
    case FNEUS_ERR_NULL: return "null pointer";
    case FNEUS_ERR_WORKSPACE: return "workspace too small";
    default: break;
  }
  if (status >= FNEUS_ERR_CUDA_BASE) return cudaGetErrorString((cudaError_t)(status - FNEUS_ERR_CUDA_BASE));
  return "unknown status";
}

}  // extern "C"

extern "C" {

// ---- profiling hooks (bench.py): per-class CUDA-event time, launch counts, algorithmic flops/bytes ----
int fneus_prof_classes(void) { return fneus::PC_COUNT; }

int fneus_prof_enable(int on) {
  fneus::ProfState& s = fneus::prof_state();
  s.on = on != 0;
  return FNEUS_OK;
}

// Synchronises the recorded events, ADDS per-class totals into the caller's arrays (length fneus_prof_classes())
// and clears the record list and the launch counters.
int fneus_prof_collect(double* ms, long long* launches, double* flops, double* bytes) {
  fneus::ProfState& s = fneus::prof_state();
  for (auto& r : s.recs) {
    float t = 0.f;
    cudaError_t e = cudaEventSynchronize(r.b);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&t, r.a, r.b);
    if (e != cudaSuccess) return fneus::fneus_cuda_error((int)e);
    if (ms) ms[r.cls] += t;
    if (flops) flops[r.cls] += r.flops;
    if (bytes) bytes[r.cls] += r.bytes;
    s.pool.push_back(r.a);
    s.pool.push_back(r.b);
  }
  s.recs.clear();
  for (int c = 0; c < fneus::PC_COUNT; c++) {
    if (launches) launches[c] += s.launches[c];
    s.launches[c] = 0;
  }
  return FNEUS_OK;
}

}  // extern "C"

extern "C" {

// debug/bisect switches of the persistent tensor-core kernel (bit0: skip A loads, bit1: skip epilogue, bit2: skip MMA)
int fneus_debug_timeline(unsigned long long* host_dst, int n) {
  if (!host_dst || n <= 0 || n > 8192) return FNEUS_ERR_BAD_SHAPE;
  cudaError_t e = cudaMemcpyFromSymbol(host_dst, fneus::g_sc_dbg, (size_t)n * 8, (size_t)(8192 - n) * 8 * 0);
  return e == cudaSuccess ? FNEUS_OK : fneus::fneus_cuda_error((int)e);
}
// The wait watchdog's record (gemm_tc.cuh mbar_hang): out8[0] != 0 when a kernel of this library trapped on a wait that
// never completed; [1] = blockIdx.x << 32 | threadIdx.x, [2] = gridDim.x << 32 | blockDim.x, [3] = barrier shared-memory
// address << 32 | parity, [4] = blockIdx.y << 32 | blockIdx.z, [5] = 0x600DD06 once complete, [8..55] = the 64-bit words
// of shared memory [barrier - 128, barrier + 256).  64 words.  Works after the CUDA context has died.
int fneus_debug_hang_record(unsigned long long* out64) {
  if (!out64) return FNEUS_ERR_NULL;
  const volatile unsigned long long* h = fneus::hang_rec_host();
  for (int i = 0; i < 64; i++) out64[i] = h ? h[i] : 0ull;
  return FNEUS_OK;
}
int fneus_debug_flags(int flags) {
  fneus::tc_debug_flags() = flags & 0xFF;
  int w = (flags >> 8) & 0xF;                      // bits 8-11: weight-gradient CTAs per SM (tuning knob), 0 = keep
  if (w) fneus::tc_wgrad_ctas_per_sm() = w;
  return FNEUS_OK;
}

// Raw dense-layer contractions in the current precision mode (test hook for the two GEMM engines).
// kind 0: C[M,N] = A[M,K] W[N,K]^T + bias ; kind 1: C[M,N] = A[M,K] W[K,N] ; kind 2: C[N,K] += Y[M,N]^T A[M,K],
// bias[N] += colsum(Y)  (for kind 2 the `W` argument is Y with leading dimension ldw).
int fneus_debug_gemm(int precision, int kind, const float* A, int lda, const float* W, int ldw, float* bias, long long M,
                     int N, int K, float* C, int ldc, void* stream) {
  using namespace fneus;
  PrecScope prec_scope_(precision);
  if (!A || !W || !C) return FNEUS_ERR_NULL;
  if (lda % 4 != 0) return FNEUS_ERR_MISALIGNED;
  cudaStream_t st = (cudaStream_t)stream;
  ASeg a = aseg_mem(A, lda, K);
  Epi e = epi_default();
  e.C = C; e.ldc = ldc;
  if (kind == 0) { e.bias = bias; launch_gemm_fwd(a, W, ldw, 0, M, N, e, st); }
  else if (kind == 1) launch_gemm_bwd_data(a, W, ldw, 0, M, N, e, st);
  else if (kind == 2) {
    if (ldw % 4 != 0) return FNEUS_ERR_MISALIGNED;
    launch_gemm_wgrad(W, ldw, a, C, ldc, 0, bias, M, N, num_sms(), st);
  } else return FNEUS_ERR_UNSUPPORTED;
  FNEUS_CHECK_LAUNCH();
  return FNEUS_OK;
}

}  // extern "C"

// Status decoding and version entry points of the C ABI (include/fneus.h).
#include "fneus_common.cuh"
#include "prof.cuh"

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

// Status decoding and version entry points of the C ABI (include/fneus.h).
#include "fneus_common.cuh"

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

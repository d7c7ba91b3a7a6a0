// Optional per-launch CUDA-event timing and launch counting (bench.py roofline / gpu_launches).
// Host-side only; disabled by default (then only a counter is bumped per launch).
#pragma once
#include <cuda_runtime.h>

#include <vector>

namespace fneus {

// PC_TC_MLP: single-layer tensor-core GEMMs; the fused chains and the grouped weight gradients have their own classes so
// that bench.py can put each kernel on its own roofline (flops = algorithmic FLOPs, bytes = designed DRAM bytes).
enum ProfClass { PC_GEMM_FWD = 0, PC_GEMM_BWD_DATA, PC_GEMM_WGRAD, PC_SAMPLING, PC_COMPOSITE, PC_ELEMENTWISE,
                 PC_TC_MLP, PC_CHAIN_SDF_FWD, PC_CHAIN_SDF_BWD, PC_CHAIN_RELU, PC_TC_WGRAD, PC_COUNT };

struct ProfRec { cudaEvent_t a, b; int cls; double flops, bytes; };
struct ProfState {
  bool on = false;
  std::vector<ProfRec> recs;
  std::vector<cudaEvent_t> pool;
  long long launches[PC_COUNT] = {0};
  ProfRec cur;
  bool open = false;
};
inline ProfState& prof_state() { static ProfState s; return s; }

inline cudaEvent_t prof_event(ProfState& s) {
  if (!s.pool.empty()) { cudaEvent_t e = s.pool.back(); s.pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
inline void prof_begin(int cls, double flops, double bytes, cudaStream_t st) {
  ProfState& s = prof_state();
  s.launches[cls]++;
  if (!s.on) return;
  s.cur.a = prof_event(s); s.cur.b = prof_event(s); s.cur.cls = cls; s.cur.flops = flops; s.cur.bytes = bytes;
  cudaEventRecord(s.cur.a, st);
  s.open = true;
}
inline void prof_end(cudaStream_t st) {
  ProfState& s = prof_state();
  if (!s.on || !s.open) return;
  cudaEventRecord(s.cur.b, st);
  s.recs.push_back(s.cur);
  s.open = false;
}

}  // namespace fneus

// Engine dispatch: which implicit-GEMM kernel family executes the WaveNet contractions.
#include "common.cuh"
#include "model.h"

namespace fwn {

int prepare_engine(Model* m, const Workspace& w, int B, int T, cudaStream_t st) {
  if (m->cfg.precision == FWN_FP32) return 0;
  set_error("mixed-precision (tcgen05) engine is not available in this build");
  return 1;
}

int run_gemm(Model* m, const GemmArgs& g, EpiKind kind, int gemm_id, const FlowPack& fp, cudaStream_t st) {
  m->launches++;
  if (m->cfg.precision == FWN_FP32) return simt_gemm(g, kind, st);
  set_error("mixed-precision (tcgen05) engine is not available in this build");
  return 1;
}

void engine_free(Model* m) {}

}  // namespace fwn

// Engine dispatch: which implicit-GEMM kernel family executes the WaveNet contractions.
//   FWN_FP32       -> CUDA-core fp32 engine (conv_simt.cu), the parity mode
//   FWN_MIXED_BF16 -> tcgen05/TMEM/TMA engine (gemm_tc.cu), the throughput mode
// There is no cross-fallback: a mode either runs on its engine or fails.
#include "common.cuh"
#include "model.h"

namespace fwn {

int tc_prepare(Model* m, const Workspace& w, int B, int T, cudaStream_t st);
int tc_run(Model* m, const GemmArgs& g, EpiKind kind, int gemm_id, const FlowPack& fp, cudaStream_t st);
void tc_free(Model* m);

int prepare_engine(Model* m, const Workspace& w, int B, int T, cudaStream_t st) {
  if (m->cfg.precision == FWN_FP32) return 0;
  return tc_prepare(m, w, B, T, st);
}

int run_gemm(Model* m, const GemmArgs& g, EpiKind kind, int gemm_id, const FlowPack& fp, cudaStream_t st) {
  m->launches++;
  if (m->cfg.precision == FWN_FP32) return simt_gemm(g, kind, st);
  return tc_run(m, g, kind, gemm_id, fp, st);
}

void engine_free(Model* m) { tc_free(m); }

}  // namespace fwn

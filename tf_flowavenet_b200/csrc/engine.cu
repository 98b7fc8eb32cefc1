// Engine dispatch: which implicit-GEMM kernel family executes the WaveNet contractions.
//   FWN_FP32       -> the parity mode: 3-way bf16-split tcgen05 engine (gemm_tc3.cu, fp32-GEMM accuracy) for the WaveNet GEMMs;
//                     CUDA-core fp32 engine (conv_simt.cu) for the front conv and when FWN_FP32_ENGINE=simt
//   FWN_MIXED_BF16 -> tcgen05/TMEM/TMA engine (gemm_tc.cu), the throughput mode
// There is no cross-fallback: a mode either runs on its engine or fails.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "model.h"

namespace fwn {

int tc_prepare(Model* m, const Workspace& w, int B, int T, cudaStream_t st);
int tc_run(Model* m, const GemmArgs& g, EpiKind kind, int gemm_id, const FlowPack& fp, cudaStream_t st);
void tc_free(Model* m);
bool tc3_supported(const GemmArgs& g);
int tc3_gemm(const GemmArgs& g, EpiKind kind, const void* w3, int Kpad, int Npad, int nterms, cudaStream_t st);

// FWN_FP32_ENGINE = tc3 (default) | simt
static bool fp32_split_engine() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FWN_FP32_ENGINE");
    v = (e && !strcmp(e, "simt")) ? 0 : 1;
  }
  return v == 1;
}

int prepare_engine(Model* m, const Workspace& w, int B, int T, cudaStream_t st) {
  if (m->cfg.precision == FWN_FP32) return 0;
  return tc_prepare(m, w, B, T, st);
}

int run_gemm(Model* m, const GemmArgs& g, EpiKind kind, int gemm_id, const FlowPack& fp, cudaStream_t st) {
  m->launches++;
  if (m->cfg.precision == FWN_FP32) {
    const W3& w3 = fp.w3[gemm_id];
    // FWN_SIMT_FAMILIES: debugging aid, bitmask of GEMM families forced onto the CUDA-core engine (1 front, 2 gate, 4 res|skip, 8 final, 16 zero)
    const char* simt_env = getenv("FWN_SIMT_FAMILIES");   // read per call: a debugging script flips it between passes
    const int simt_mask = simt_env ? atoi(simt_env) : 0;
    const int fam = gemm_id == GEMM_FRONT ? 1 : gemm_id == GEMM_FINAL ? 8 : gemm_id == GEMM_ZERO ? 16 : gemm_id >= GEMM_RS0 ? 4 : 2;
    if (w3.p && fp32_split_engine() && !m->force_simt && !(simt_mask & fam) && tc3_supported(g)) return tc3_gemm(g, kind, w3.p, w3.Kpad, w3.Npad, m->cur_terms, st);
    if (gemm_id == GEMM_FRONT) {   // the fp32 [3][nq][F] layout has tap k at row k*nq; Seg.koff carries the plane layout (k*front_k16)
      GemmArgs gg = g;
      for (int s = 0; s < gg.nseg; ++s) gg.seg[s].koff = s * fp.nq;
      return simt_gemm(gg, kind, st);
    }
    return simt_gemm(g, kind, st);
  }
  return tc_run(m, g, kind, gemm_id, fp, st);
}

void engine_free(Model* m) { tc_free(m); }

}  // namespace fwn

// placeholder until the tcgen05 kernel lands (replaced in the next commit)
#include "score.cuh"
namespace cfl {
bool score_umma_supported(int, int, const float*, int64_t) { return false; }
size_t score_umma_qimg_bytes(const ScorePlan&, int) { return 0; }
int score_umma_pack_queries(const ScoreArgs&, void*, cudaStream_t) { return CFL_ERR_UNSUPPORTED; }
int score_umma_launch(const ScoreArgs&, cudaStream_t) { return CFL_ERR_UNSUPPORTED; }
int score_umma_qt(int, int) { return 0; }
}

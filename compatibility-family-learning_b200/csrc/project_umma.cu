// placeholder until the tcgen05 kernel lands (replaced in the next commit)
#include "common.cuh"
namespace cfl {
int project_fwd_umma(const float*, int64_t, int, int64_t, const float*, int, int64_t, const float*,
                     const float*, float, int, float*, int64_t, float*, float*, void*, size_t,
                     cudaStream_t) { return CFL_ERR_UNSUPPORTED; }
size_t project_fwd_umma_workspace(int64_t, int, int) { return 0; }
bool project_fwd_umma_supported(const float*, int64_t, int, int64_t, int) { return false; }
}
extern "C" int cfl_selftest_umma(const float*, const float*, float*, int, int, void*) {
  cfl::set_error("selftest_umma: not built yet");
  return CFL_ERR_UNSUPPORTED;
}

// score_select_tc.cu — tcgen05/TMEM engine for the greedy score+select (placeholder
// until the tf32 filter + fp32 refine kernel lands; reports "unsupported").
#include "pcv_common.cuh"

namespace pcv {
bool score_select_tc_supported(const Table *) { return false; }
size_t score_select_tc_workspace(const Table *, int64_t) { return 0; }
int score_select_tc(const Table *, const float *, int64_t, int64_t *, float *, void *, size_t,
                    cudaStream_t) {
  set_error("score_select: tcgen05 engine not built");
  return PCV_ERR_UNSUPPORTED;
}
}  // namespace pcv

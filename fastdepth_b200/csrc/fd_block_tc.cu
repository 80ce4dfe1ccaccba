// Fused depthwise -> pointwise block kernel on tcgen05 (placeholder until the kernel lands).
#include "fd_common.cuh"
namespace fd {
struct BlockTcPlan {};
bool block_tc_supported(int, const StageGeom&, bool) { return false; }
int block_tc_prepare(int, const BlockArgs&, const float*, float, float, int, void*, BlockTcPlan**) {
    return fail(FD_ERR_UNSUPPORTED, "fused block kernel not built");
}
int block_tc_launch(BlockTcPlan*, cudaStream_t) { return fail(FD_ERR_UNSUPPORTED, "fused block kernel not built"); }
void block_tc_destroy(BlockTcPlan*) {}
const char* block_tc_name(BlockTcPlan*) { return "block_tc"; }
}  // namespace fd

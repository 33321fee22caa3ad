// placeholder: tcgen05 correlation (filled in next)
#include "common.cuh"
namespace stm {
bool corr_tc_supported(const StmCorrDesc&, const char** why) { *why = "not built"; return false; }
int launch_corr_tc(const StmCorrDesc&, const void*, const void*, const void*, const void*, void*, cudaStream_t) { set_error("not built"); return STM_ERR_UNSUPPORTED; }
}

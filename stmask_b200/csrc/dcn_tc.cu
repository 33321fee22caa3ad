// placeholder: tcgen05 deformable conv (filled in next)
#include "common.cuh"
namespace stm {
bool dcn_tc_supported(const StmDcnConv*, const StmDcnProblem*, int, const char** why) { *why = "not built"; return false; }
size_t dcn_tc_workspace(const StmDcnConv*, const StmDcnProblem*, int) { return 0; }
int launch_dcn_tc(const StmDcnConv*, const DcnParams&, void*, size_t, cudaStream_t) { set_error("not built"); return STM_ERR_UNSUPPORTED; }
}

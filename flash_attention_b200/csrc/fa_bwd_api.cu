// fa_bwd_api.cu — backward entry points (placeholder until the sm_100a backward kernels land).
#include "../../include/fa_b200.h"
#include "fa_api_common.h"

extern "C" {
int fa_b200_mha_bwd(const fa_b200_bwd_args*, fa_b200_stream_t) { return fa::set_error("fa_b200 backward is not built in this revision"); }
int fa_b200_mha_varlen_bwd(const fa_b200_bwd_args*, fa_b200_stream_t) { return fa::set_error("fa_b200 backward is not built in this revision"); }
size_t fa_b200_bwd_workspace_bytes(const fa_b200_bwd_args*) { return 0; }
}

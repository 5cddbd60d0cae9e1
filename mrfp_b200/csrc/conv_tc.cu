#include "hrfp.cuh"
namespace mrfp {
bool conv3x3_tc_supported(int, int) { return false; }
int conv3x3_tc_bf16(const __nv_bfloat16*, const __nv_bfloat16*, __nv_bfloat16*, int, int, int, int, int, int, const int*, const int*, double*, cudaStream_t) { return MRFP_ERR_UNSUPPORTED; }
}

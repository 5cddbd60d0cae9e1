// placeholder: real implementation follows
#include "common.cuh"
struct mrfp_hrfp_plan { int dummy; };
extern "C" int mrfp_hrfp_plan_create(mrfp_hrfp_plan_t** plan, int, int, int, int, int, int, const int*, int) { return MRFP_ERR_UNSUPPORTED; }
extern "C" void mrfp_hrfp_plan_destroy(mrfp_hrfp_plan_t*) {}
extern "C" size_t mrfp_hrfp_plan_ws_bytes(const mrfp_hrfp_plan_t*) { return 0; }
extern "C" size_t mrfp_hrfp_plan_saved_bytes(const mrfp_hrfp_plan_t*) { return 0; }
extern "C" size_t mrfp_hrfp_plan_lut_bytes(const mrfp_hrfp_plan_t*) { return 0; }
extern "C" int mrfp_hrfp_plan_write_luts(const mrfp_hrfp_plan_t*, void*, size_t) { return MRFP_ERR_UNSUPPORTED; }
extern "C" int mrfp_hrfp_plan_stage(const mrfp_hrfp_plan_t*, int, int*) { return MRFP_ERR_UNSUPPORTED; }
extern "C" int mrfp_hrfp_fwd(const mrfp_hrfp_plan_t*, const float*, const float* const*, const float* const*, const float* const*,
                  float* const*, float* const*, float, float, const float*, float*, float*, const void*, void*, void*, void*) { return MRFP_ERR_UNSUPPORTED; }
extern "C" int mrfp_hrfp_bwd(const mrfp_hrfp_plan_t*, const float*, const float*, const float* const*, const void*, const void*, float*, void*, void*) { return MRFP_ERR_UNSUPPORTED; }
extern "C" int mrfp_add_f32(const float*, const float*, float*, size_t, void*) { return MRFP_ERR_UNSUPPORTED; }

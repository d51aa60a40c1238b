// conv_tc.cu -- placeholder until the tcgen05 path lands (next commit).
#include "common.cuh"
namespace cattl3 {
bool tc_gather_gemm_supported(const cattl3_ctx*, const GatherGeom&) { return false; }
int tc_gather_gemm_f32(cattl3_ctx*, const GatherGeom&, const float*, const float*, const float*, int, float*) {
	set_error("tcgen05 path not built");
	return CATTL3_ERR_UNSUPPORTED;
}
bool tc_wgrad_supported(const cattl3_ctx*, const GatherGeom&) { return false; }
int tc_wgrad_f32(cattl3_ctx*, const GatherGeom&, const float*, const float*, float*) {
	set_error("tcgen05 path not built");
	return CATTL3_ERR_UNSUPPORTED;
}
}

// vt_resample_brick.cu -- TMA-staged shared-memory brick-cache kernel family (placeholder until implemented).
#include "vt_common.cuh"
int vt_brick_supported(const VtResampleParams &, int) { return 0; }
int vt_launch_brick(const VtResampleParams &, int, cudaStream_t) { return VT_ERR_UNSUPPORTED; }

// Entry points declared in include/cppf_b200.h whose kernels have not landed yet return
// CPPF_ERR_UNSUPPORTED (the Python layer raises).  Each block disappears when its real translation
// unit (shot.cu, heads.cu) is added to the build.
#include "common.cuh"

#ifndef CPPF_HAVE_SHOT
CPPF_API int64_t cppf_shot_workspace_bytes(int64_t) { return 0; }
CPPF_API int cppf_shot_compute(const float *, int64_t, float, float, float *, float *, void *, int64_t, void *) { return CPPF_ERR_UNSUPPORTED; }
CPPF_API int cppf_estimate_normal(const float *, int64_t, float, float *, void *, int64_t, void *) { return CPPF_ERR_UNSUPPORTED; }
CPPF_API int cppf_shot_compute_color(const float *, const float *, int64_t, float, float, float *, void *) { return CPPF_ERR_UNSUPPORTED; }
#endif

#ifndef CPPF_HAVE_HEADS
CPPF_API int cppf_heads_create(int, int, const float *, int64_t, cppf_heads **) { return CPPF_ERR_UNSUPPORTED; }
CPPF_API int cppf_heads_destroy(cppf_heads *) { return CPPF_ERR_UNSUPPORTED; }
CPPF_API int cppf_heads_has_tc(const cppf_heads *) { return 0; }
CPPF_API int64_t cppf_heads_workspace_bytes(const cppf_heads *, int64_t, int64_t, int) { return 0; }
CPPF_API int cppf_heads_forward(const cppf_heads *, int, const float *, int64_t, const void *, int, int64_t, int64_t, const float *,
                                const float *, float *, float *, void *, int64_t, void *) { return CPPF_ERR_UNSUPPORTED; }
#endif

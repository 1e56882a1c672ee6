// bf16 tensor-core (tcgen05 / TMEM) path of the BeyondCPPF heads.  Placeholder until the kernel lands:
// creation reports CPPF_ERR_UNSUPPORTED, which cppf_heads_create treats as "float32 path only".
#include "heads_common.cuh"

using namespace cppf;

extern "C" int cppf_heads_tc_create(const HeadsModel *, const float *, void **state) {
    *state = nullptr;
    return CPPF_ERR_UNSUPPORTED;
}
extern "C" void cppf_heads_tc_destroy(void *) {}
extern "C" int64_t cppf_heads_tc_workspace_bytes(const void *, int64_t, int64_t) { return 0; }
extern "C" int cppf_heads_tc_forward(const void *, const float *, int64_t, const void *, int, int64_t, int64_t, const float *,
                                     const float *, float *, float *, void *, int64_t, void *) {
    return CPPF_ERR_UNSUPPORTED;
}

#include "common.cuh"
namespace tspn {
int span_head_tensor(const float*, const int64_t*, int64_t, int64_t, int64_t, int, int, const float*, const float*, const float*, const float*, int, float*, void*, cudaStream_t) { set_error("tensor span head not built yet"); return TSPN_EBADARG; }
}
extern "C" {
int64_t tspn_span_head_workspace_bytes(int64_t, int, int, int, int) { return 0; }
int64_t tspn_postprocess_workspace_bytes(int64_t, int) { return 0; }
int tspn_postprocess(const int64_t*, int, const float*, const int64_t*, const int64_t*, int, const float*, int, const int32_t*, int, int, int32_t*, int32_t*, void*, void*) { tspn::set_error("postprocess not built yet"); return TSPN_EBADARG; }
}

// simple-knn (distCUDA2) rebuilt -- placeholder until the search kernels land.
#include "gsr_common.cuh"
namespace gsr
{
size_t knn_workspace_bytes(int P) { return P > 0 ? 256 : 0; }
int knn_run(int, const float*, float*, void*, size_t, cudaStream_t)
{
    set_error("gsr_knn_dist2 is not built yet");
    return GSR_ERR_UNSUPPORTED;
}
} // namespace gsr

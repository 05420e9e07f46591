// mlp_tc.cu -- tensor-core (TF32) variant of K6.  Placeholder until the mma path lands.
#include "common.cuh"

namespace nnmd {
int mlp_tf32_run(int, const int32_t *, const void *const *, const void *const *, const void *, int64_t, void *, void *,
                 cudaStream_t) {
    return set_err(NNMD_ERR_INVALID, "mlp: TF32 tensor-core path is not built in this revision");
}
}  // namespace nnmd

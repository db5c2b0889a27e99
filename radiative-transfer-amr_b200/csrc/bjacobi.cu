// Block-Jacobi preconditioner (optional; north-star subsystem 4).
// Placeholder translation unit: the preconditioner is not built yet, so a
// solve that asks for it fails loudly instead of silently running without it.
#include "rtb200_internal.cuh"

int rt_pc_setup(rtb200_ctx* ctx, cudaStream_t) {
    return rt_fail(ctx, RTB200_EUNSUPPORTED, "block-Jacobi preconditioner is not available in this build");
}
int rt_pc_apply(rtb200_ctx* ctx, const double*, double*, cudaStream_t) {
    return rt_fail(ctx, RTB200_EUNSUPPORTED, "block-Jacobi preconditioner is not available in this build");
}
void rt_pc_destroy(rtb200_ctx* ctx) {
    ctx->pc_ready = false;
}

// placeholder until the tcgen05 path lands (replaced in the next milestone)
#include "common.cuh"
int tc_create(isdfb_ctx* ctx) { ISDFB_FAIL(ctx, ISDFB_ERR_ARG, "tensor-core path not built yet"); }
void tc_destroy(isdfb_ctx*) {}
int tc_repack(isdfb_ctx*, cudaStream_t) { return ISDFB_OK; }
int tc_forward(isdfb_ctx* ctx, const float*, const float*, float, int64_t, float*, float*, cudaStream_t) { ISDFB_FAIL(ctx, ISDFB_ERR_ARG, "tensor-core path not built yet"); }
int tc_train(isdfb_ctx* ctx, const float*, const float*, const float*, const float*, const float*, const float*, const float*, const uint8_t*, int64_t, int32_t, const isdfb_loss_cfg*, float*, float*, float*, float*, cudaStream_t) { ISDFB_FAIL(ctx, ISDFB_ERR_ARG, "tensor-core path not built yet"); }

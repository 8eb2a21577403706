// dist.cu -- multi-GPU 2D block-cyclic Cholesky + NLL over NCCL (placeholder until the
// single-GPU path is validated on hardware; the entry points report DGP_ERR_NCCL).
#include "common.cuh"

namespace dgp {
struct DistState {};
void dist_destroy(Ctx *ctx) { (void)ctx; }
}  // namespace dgp

using namespace dgp;
extern "C" {
int32_t dgp_dist_unique_id(void *id128) { (void)id128; return DGP_ERR_NCCL; }
int32_t dgp_dist_init(dgp_ctx *ctx, const void *, int32_t, int32_t, int32_t, int32_t) {
  return ctx ? ctx->fail(DGP_ERR_NCCL, "distributed path not built yet") : DGP_ERR_ARG;
}
int32_t dgp_dist_finalize(dgp_ctx *ctx) { (void)ctx; return DGP_OK; }
int32_t dgp_dist_energy(dgp_ctx *ctx, const dgp_data *, const dgp_kernel_spec *, int32_t, double *, double *) {
  return ctx ? ctx->fail(DGP_ERR_NCCL, "distributed path not built yet") : DGP_ERR_ARG;
}
}

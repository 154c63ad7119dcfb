// placeholder
extern "C" {
int socp_residual_batch(socp_ctx *ctx, const socp_shape *, long, const double *, const double *, const double *, const double *, double *, int) { return fail(ctx, SOCP_ERR_UNSUPPORTED, "not built yet"); }
int socp_fdjac_batch(socp_ctx *ctx, const socp_shape *, long, const double *, const double *, const double *, const double *, double, double *, int) { return fail(ctx, SOCP_ERR_UNSUPPORTED, "not built yet"); }
int socp_solve_batch(socp_ctx *ctx, const socp_shape *, long, const double *, const double *, const double *, double *, double, int, int *, int *, double *, int) { return fail(ctx, SOCP_ERR_UNSUPPORTED, "not built yet"); }
int socp_continuation_param_batch(socp_ctx *ctx, const socp_shape *, long, double *, const double *, const double *, double *, double, int, double, int, const double *, double, int *, int *) { return fail(ctx, SOCP_ERR_UNSUPPORTED, "not built yet"); }
int socp_continuation_boundary_batch(socp_ctx *ctx, const socp_shape *, long, const double *, const double *, const double *, const double *, const double *, double *, double, int, double, double, int *, int *) { return fail(ctx, SOCP_ERR_UNSUPPORTED, "not built yet"); }
}

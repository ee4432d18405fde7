// lattice_schur.cu -- batched per-cell Schur complements and the DDM interface operator.  sm_100a.
#include "common.cuh"

extern "C" int lat_schur_batch(lat_ctx* ctx, const double* xyz, const int32_t* len0, const int32_t* len1,
                               const double* rad, int64_t n_cells, int32_t n_loc_nodes, int32_t n_bnd_nodes,
                               int32_t n_loc_elem, double young, double nu, double kappa, double* S,
                               const int32_t* elem_group, const double* drad_chain, int32_t n_grad, double* dS) {
  if (!ctx) return LAT_ERR_ARG;
  return lat_fail(ctx, LAT_ERR_UNSUPPORTED, "lat_schur_batch not built yet", __FILE__, __LINE__);
}

extern "C" int lat_ddm_matvec(lat_ctx* ctx, const double* S, int64_t s_stride, const int32_t* gidx,
                              const double* u_fixed, int64_t n_cells, int32_t nb, int64_t n_free,
                              const double* x, double* y) {
  if (!ctx) return LAT_ERR_ARG;
  return lat_fail(ctx, LAT_ERR_UNSUPPORTED, "lat_ddm_matvec not built yet", __FILE__, __LINE__);
}

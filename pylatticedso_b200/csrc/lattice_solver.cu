// lattice_solver.cu -- BSR(6x6) FP64 SpMV and the two-kernel PCG iteration.  sm_100a.
//
// Thread layout shared by every kernel here: a warp owns 5 consecutive block
// rows (nodes); lane = 6*g + r handles scalar row r of node g (lanes 30, 31
// idle).  DOF index 6*node + r is therefore consecutive across lanes 0..29, so
// all vector traffic is coalesced, and the 6 lanes of a group read one 288 B
// block of the matrix as 6 x 48 B (three 16 B loads per lane).
#include <vector>

#include "common.cuh"

static constexpr int SPMV_BLOCK = 256;                        // 8 warps
static constexpr int ROWS_PER_WARP = 5;
static constexpr int ROWS_PER_CTA = ROWS_PER_WARP * (SPMV_BLOCK / 32);  // 40 nodes

__device__ __forceinline__ double dot6(const double2& a0, const double2& a1, const double2& a2,
                                       const double2& x0, const double2& x1, const double2& x2,
                                       double acc) {
  acc = fma(a0.x, x0.x, acc);
  acc = fma(a0.y, x0.y, acc);
  acc = fma(a1.x, x1.x, acc);
  acc = fma(a1.y, x1.y, acc);
  acc = fma(a2.x, x2.x, acc);
  acc = fma(a2.y, x2.y, acc);
  return acc;
}

// ---------------------------------------------------------------------------
// plain y = A x
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(SPMV_BLOCK) k_bsr_spmv(const int32_t* __restrict__ rowptr,
                                                         const int32_t* __restrict__ colidx,
                                                         const double* __restrict__ vals, int64_t n_nodes,
                                                         const double* __restrict__ x, double* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int g = lane / 6, r = lane - g * 6;
  const int64_t warp = (int64_t)blockIdx.x * (SPMV_BLOCK / 32) + (threadIdx.x >> 5);
  const int64_t n = warp * ROWS_PER_WARP + g;
  if (g >= ROWS_PER_WARP || n >= n_nodes) return;
  const int lo = rowptr[n], hi = rowptr[n + 1];
  double acc = 0.0;
#pragma unroll 4
  for (int j = lo; j < hi; ++j) {
    const int c = __ldg(colidx + j);
    const double2* vp = reinterpret_cast<const double2*>(vals + (int64_t)j * 36 + r * 6);
    const double2* xp = reinterpret_cast<const double2*>(x + (int64_t)c * 6);
    const double2 a0 = __ldcs(vp), a1 = __ldcs(vp + 1), a2 = __ldcs(vp + 2);  // streamed once: evict-first
    const double2 x0 = __ldg(xp), x1 = __ldg(xp + 1), x2 = __ldg(xp + 2);
    acc = dot6(a0, a1, a2, x0, x1, x2, acc);
  }
  y[n * 6 + r] = acc;
}

int lat_spmv_internal(lat_ctx* ctx, const int32_t* rowptr, const int32_t* colidx, const double* vals,
                      int64_t n_nodes, const double* x, double* y) {
  LAT_LAUNCH(ctx, k_bsr_spmv, (unsigned)ceil_div(n_nodes, ROWS_PER_CTA), SPMV_BLOCK, 0, rowptr, colidx, vals,
             n_nodes, x, y);
  return LAT_OK;
}

extern "C" int lat_bsr_spmv(lat_ctx* ctx, const int32_t* rowptr, const int32_t* colidx,
                            const double* vals, int64_t n_nodes, const double* x, double* y) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, rowptr && colidx && vals && x && y && n_nodes > 0);
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  return lat_spmv_internal(ctx, rowptr, colidx, vals, n_nodes, x, y);
}

// ---------------------------------------------------------------------------
// preconditioner setup
// ---------------------------------------------------------------------------
// dinv layout: Jacobi -> [6n] reciprocal diagonal; block-Jacobi -> [n][6][6] inverse of the diagonal block.
__global__ void k_precond_setup(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                const double* __restrict__ vals, int64_t n_nodes, int precond,
                                double* __restrict__ dinv) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_nodes) return;
  int d = -1;
  for (int j = rowptr[n]; j < rowptr[n + 1]; ++j)
    if (colidx[j] == n) { d = j; break; }
  if (precond == LAT_PC_JACOBI) {
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      const double v = d >= 0 ? vals[(int64_t)d * 36 + r * 7] : 1.0;
      dinv[n * 6 + r] = (v > 0.0) ? 1.0 / v : 1.0;
    }
    return;
  }
  // 6x6 SPD inverse by Gauss-Jordan without pivoting
  double a[6][6], inv[6][6];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      a[i][k] = d >= 0 ? vals[(int64_t)d * 36 + i * 6 + k] : (i == k ? 1.0 : 0.0);
      inv[i][k] = (i == k) ? 1.0 : 0.0;
    }
  bool ok = true;
#pragma unroll
  for (int p = 0; p < 6; ++p) {
    const double piv = a[p][p];
    if (!(piv > 0.0)) ok = false;
    const double ip = 1.0 / piv;
#pragma unroll
    for (int k = 0; k < 6; ++k) { a[p][k] *= ip; inv[p][k] *= ip; }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      if (i == p) continue;
      const double f = a[i][p];
#pragma unroll
      for (int k = 0; k < 6; ++k) { a[i][k] -= f * a[p][k]; inv[i][k] -= f * inv[p][k]; }
    }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int k = 0; k < 6; ++k) dinv[n * 36 + i * 6 + k] = ok ? inv[i][k] : (i == k ? 1.0 : 0.0);
}

// z_r = (M^-1 r)_r for the lane's DOF; rv = the lane's residual entry.  All 32 lanes must call.
template <int PC>
__device__ __forceinline__ double apply_precond(const double* __restrict__ dinv, int64_t n, int g, int r,
                                                bool active, double rv) {
  if (PC == LAT_PC_NONE) return rv;
  if (PC == LAT_PC_JACOBI) return active ? dinv[n * 6 + r] * rv : 0.0;
  double z = 0.0;
  const double* row = dinv + n * 36 + r * 6;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double rk = __shfl_sync(0xffffffffu, rv, g * 6 + k);
    if (active) z = fma(row[k], rk, z);
  }
  return z;
}

// ---------------------------------------------------------------------------
// PCG kernels
// ---------------------------------------------------------------------------
struct PcgParams {
  double tol, mintol, alpha_max;
  int64_t restart_every;
  int32_t maxiter, reference;
};

// init: x = 0, r = b, z = M^-1 r, p[0] = z; rz_old = r.z, bb = b.b
template <int PC>
__global__ void __launch_bounds__(SPMV_BLOCK) k_pcg_init(int64_t n_nodes, const double* __restrict__ b,
                                                         const double* __restrict__ dinv, double* __restrict__ x,
                                                         double* __restrict__ r, double* __restrict__ z,
                                                         double* __restrict__ p0, double* __restrict__ p1,
                                                         PcgScalars* __restrict__ sc, double* __restrict__ partials) {
  const int lane = threadIdx.x & 31;
  const int g = lane / 6, rr_ = lane - g * 6;
  const int64_t warp = (int64_t)blockIdx.x * (SPMV_BLOCK / 32) + (threadIdx.x >> 5);
  const int64_t n = warp * ROWS_PER_WARP + g;
  const bool active = g < ROWS_PER_WARP && n < n_nodes;
  const int64_t i = n * 6 + rr_;
  const double bv = active ? b[i] : 0.0;
  const double zv = apply_precond<PC>(dinv, n, g, rr_, active, bv);
  if (active) { x[i] = 0.0; r[i] = bv; z[i] = zv; p0[i] = zv; p1[i] = zv; }
  double v[2] = {bv * zv, bv * bv}, out[2];
  if (grid_reduce<2, SPMV_BLOCK>(v, partials, &sc->counter[0], out)) {
    sc->rz_old = out[0];
    sc->bb = out[1];
    sc->rr = out[1];
    sc->beta = 0.0;
    sc->pAp = 0.0; sc->pp = 0.0; sc->xx = 0.0; sc->alpha_last = 0.0;
    sc->iters = 0; sc->info_flag2 = 0; sc->breakdown = 0;
    sc->done = (out[1] == 0.0) ? 1 : 0;   // b == 0 -> x = 0
  }
}

// kernel 1 of an iteration: p_new = z + beta p_old (own rows written, neighbour rows recomputed on
// the fly), Ap = A p_new, partial sums p.Ap and p.p.  p buffers ping-pong with the iteration parity.
template <int DUMMY>
__global__ void __launch_bounds__(SPMV_BLOCK) k_pcg_spmv(const int32_t* __restrict__ rowptr,
                                                         const int32_t* __restrict__ colidx,
                                                         const double* __restrict__ vals, int64_t n_nodes,
                                                         const double* __restrict__ z, double* __restrict__ pa,
                                                         double* __restrict__ pb, double* __restrict__ Ap,
                                                         PcgScalars* __restrict__ sc, double* __restrict__ partials,
                                                         PcgParams prm) {
  if (sc->done || sc->iters >= prm.maxiter) return;
  const int k = sc->iters;
  const double beta = sc->beta;
  const double* __restrict__ p_old = (k & 1) ? pb : pa;
  double* __restrict__ p_new = (k & 1) ? pa : pb;
  const int lane = threadIdx.x & 31;
  const int g = lane / 6, r = lane - g * 6;
  const int64_t warp = (int64_t)blockIdx.x * (SPMV_BLOCK / 32) + (threadIdx.x >> 5);
  const int64_t n = warp * ROWS_PER_WARP + g;
  const bool active = g < ROWS_PER_WARP && n < n_nodes;
  double acc = 0.0, pv = 0.0;
  if (active) {
    const int lo = rowptr[n], hi = rowptr[n + 1];
#pragma unroll 4
    for (int j = lo; j < hi; ++j) {
      const int c = __ldg(colidx + j);
      const double2* vp = reinterpret_cast<const double2*>(vals + (int64_t)j * 36 + r * 6);
      const double2* zp = reinterpret_cast<const double2*>(z + (int64_t)c * 6);
      const double2* pp = reinterpret_cast<const double2*>(p_old + (int64_t)c * 6);
      const double2 a0 = __ldcs(vp), a1 = __ldcs(vp + 1), a2 = __ldcs(vp + 2);
      const double2 z0 = zp[0], z1 = zp[1], z2 = zp[2];
      const double2 q0 = pp[0], q1 = pp[1], q2 = pp[2];
      double2 x0, x1, x2;
      x0.x = fma(beta, q0.x, z0.x); x0.y = fma(beta, q0.y, z0.y);
      x1.x = fma(beta, q1.x, z1.x); x1.y = fma(beta, q1.y, z1.y);
      x2.x = fma(beta, q2.x, z2.x); x2.y = fma(beta, q2.y, z2.y);
      acc = dot6(a0, a1, a2, x0, x1, x2, acc);
    }
    const int64_t i = n * 6 + r;
    pv = fma(beta, p_old[i], z[i]);
    p_new[i] = pv;
    Ap[i] = acc;
  }
  double v[2] = {pv * acc, pv * pv}, out[2];
  if (grid_reduce<2, SPMV_BLOCK>(v, partials, &sc->counter[1], out)) {
    sc->pAp = out[0];
    sc->pp = out[1];
  }
}

// kernel 2: alpha = min(rz/pAp, alpha_max); x += alpha p; r -= alpha Ap; z = M^-1 r;
// partial sums r.z, r.r, x.x; the last block evaluates the stop tests and beta.
template <int PC>
__global__ void __launch_bounds__(SPMV_BLOCK) k_pcg_update(int64_t n_nodes, const double* __restrict__ dinv,
                                                           double* __restrict__ x, double* __restrict__ r,
                                                           double* __restrict__ z, double* __restrict__ pa,
                                                           double* __restrict__ pb, const double* __restrict__ Ap,
                                                           PcgScalars* __restrict__ sc, double* __restrict__ partials,
                                                           PcgParams prm) {
  if (sc->done || sc->iters >= prm.maxiter) return;
  const int k = sc->iters;
  double* __restrict__ p_cur = (k & 1) ? pa : pb;  // written by k_pcg_spmv of this iteration
  const double pAp = sc->pAp;
  double alpha = sc->rz_old / pAp;
  if (prm.reference && prm.alpha_max > 0.0) alpha = fmin(alpha, prm.alpha_max);
  const bool restart = prm.reference && prm.restart_every > 0 && k > 0 && (k % prm.restart_every) == 0;
  const int lane = threadIdx.x & 31;
  const int g = lane / 6, rr_ = lane - g * 6;
  const int64_t warp = (int64_t)blockIdx.x * (SPMV_BLOCK / 32) + (threadIdx.x >> 5);
  const int64_t n = warp * ROWS_PER_WARP + g;
  const bool active = g < ROWS_PER_WARP && n < n_nodes;
  const int64_t i = n * 6 + rr_;
  double xv = 0.0, rv = 0.0, pnorm = 0.0;
  if (active) {
    const double pv = p_cur[i];
    xv = fma(alpha, pv, x[i]);
    rv = fma(-alpha, Ap[i], r[i]);
    x[i] = xv;
    r[i] = rv;
    if (restart) {
      // conjugate_gradient_solver.py:89-90: p = z.copy().  With a preconditioner z is still the
      // PREVIOUS z (recomputed only at :111); with M=None z aliases r (:66-69), which :82 has just
      // updated in place, so the restart direction is the NEW residual.
      const double zo = (PC == LAT_PC_NONE) ? rv : z[i];
      p_cur[i] = zo;
      pnorm = zo * zo;
    }
  }
  const double zv = apply_precond<PC>(dinv, n, g, rr_, active, rv);
  if (active) z[i] = zv;
  double v[4] = {rv * zv, rv * rv, xv * xv, pnorm}, out[4];
  if (grid_reduce<4, SPMV_BLOCK>(v, partials, &sc->counter[2], out)) {
    const double rz_new = out[0], rr = out[1], xx = out[2];
    const double pp = restart ? out[3] : sc->pp;
    sc->rr = rr;
    sc->xx = xx;
    sc->alpha_last = alpha;
    sc->iters = k + 1;
    int done = 0;
    if (prm.reference) {
      // :97  residual_norm <= tol * norm_b      :102  direction_norm < mintol * (solution_norm + 1e-12)
      if (sqrt(rr) <= prm.tol * sqrt(sc->bb)) done = 1;
      else if (prm.mintol > 0.0 && sqrt(pp) < prm.mintol * (sqrt(xx) + 1e-12)) done = 1;
      else if (alpha < 1e-6) sc->info_flag2 = 1;  // :107-109
    } else {
      if (rr <= prm.tol * prm.tol * sc->bb) done = 1;
      else if (!(pAp > 0.0) || !(rr == rr)) { done = 1; sc->breakdown = 1; }
    }
    sc->beta = rz_new / sc->rz_old;
    sc->rz_old = rz_new;
    if (done) sc->done = 1;
  }
}

// ---------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------
template <int PC>
static int pcg_run(lat_ctx* ctx, const int32_t* rowptr, const int32_t* colidx, const double* vals,
                   int64_t n_nodes, const double* b, double* x, const lat_pcg_opts* o, lat_pcg_result* res) {
  const int64_t n = 6 * n_nodes;
  const unsigned grid = (unsigned)ceil_div(n_nodes, ROWS_PER_CTA);
  double* r = lat_buf<double>(ctx, "pcg_r", n);
  double* z = lat_buf<double>(ctx, "pcg_z", n);
  double* pa = lat_buf<double>(ctx, "pcg_pa", n);
  double* pb = lat_buf<double>(ctx, "pcg_pb", n);
  double* Ap = lat_buf<double>(ctx, "pcg_Ap", n);
  double* dinv = lat_buf<double>(ctx, "pcg_dinv", PC == LAT_PC_BLOCK6 ? 36 * n_nodes : n);
  double* partials = lat_buf<double>(ctx, "pcg_partials", (size_t)4 * grid + 8);
  PcgScalars* sc = lat_buf<PcgScalars>(ctx, "pcg_scalars", 1);
  if (!r || !z || !pa || !pb || !Ap || !dinv || !partials || !sc)
    return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
  const int64_t launches0 = ctx->launches;
  PcgParams prm;
  prm.tol = o->tol;
  prm.mintol = o->mintol;
  prm.alpha_max = o->alpha_max;
  prm.restart_every = o->restart_every;
  prm.maxiter = o->maxiter;
  prm.reference = o->reference_semantics;
  int check = o->check_every > 0 ? o->check_every : 32;
  if (check > o->maxiter) check = o->maxiter > 0 ? o->maxiter : 1;

  LAT_CUDA(ctx, cudaMemsetAsync(sc, 0, sizeof(PcgScalars), ctx->stream));
  if (PC != LAT_PC_NONE)
    LAT_LAUNCH(ctx, k_precond_setup, (unsigned)ceil_div(n_nodes, 128), 128, 0, rowptr, colidx, vals, n_nodes, PC, dinv);
  LAT_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
  LAT_LAUNCH(ctx, k_pcg_init<PC>, grid, SPMV_BLOCK, 0, n_nodes, b, dinv, x, r, z, pa, pb, sc, partials);

  // one CUDA graph = `check` iterations (2 kernels each); relaunched until the device reports done
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t gexec = nullptr;
  cudaStream_t cap = nullptr;
  LAT_CUDA(ctx, cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
  cudaError_t ce = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
  if (ce == cudaSuccess) {
    for (int it = 0; it < check; ++it) {
      k_pcg_spmv<0><<<grid, SPMV_BLOCK, 0, cap>>>(rowptr, colidx, vals, n_nodes, z, pa, pb, Ap, sc, partials, prm);
      k_pcg_update<PC><<<grid, SPMV_BLOCK, 0, cap>>>(n_nodes, dinv, x, r, z, pa, pb, Ap, sc, partials, prm);
    }
    ce = cudaStreamEndCapture(cap, &graph);
  }
  if (ce == cudaSuccess) ce = cudaGraphInstantiate(&gexec, graph, 0);
  if (ce != cudaSuccess) {
    if (graph) cudaGraphDestroy(graph);
    cudaStreamDestroy(cap);
    return lat_cuda_fail(ctx, ce, "PCG graph capture", __FILE__, __LINE__);
  }

  int rc = LAT_OK;
  // optional per-kernel timing of the first iterations (outside the graph, events between kernels)
  int nprof = o->profile_iters > 0 ? o->profile_iters : 0;
  if (nprof > o->maxiter) nprof = o->maxiter;
  if (nprof > 256) nprof = 256;
  double spmv_ms = 0.0, update_ms = 0.0;
  if (nprof > 0) {
    std::vector<cudaEvent_t> evs(3 * nprof);
    for (auto& e : evs) cudaEventCreate(&e);
    for (int it = 0; it < nprof; ++it) {
      cudaEventRecord(evs[3 * it], ctx->stream);
      k_pcg_spmv<0><<<grid, SPMV_BLOCK, 0, ctx->stream>>>(rowptr, colidx, vals, n_nodes, z, pa, pb, Ap, sc, partials, prm);
      cudaEventRecord(evs[3 * it + 1], ctx->stream);
      k_pcg_update<PC><<<grid, SPMV_BLOCK, 0, ctx->stream>>>(n_nodes, dinv, x, r, z, pa, pb, Ap, sc, partials, prm);
      cudaEventRecord(evs[3 * it + 2], ctx->stream);
      ctx->launches += 2;
    }
    ce = cudaStreamSynchronize(ctx->stream);
    if (ce == cudaSuccess) {
      for (int it = 0; it < nprof; ++it) {
        float a = 0.f, b2 = 0.f;
        cudaEventElapsedTime(&a, evs[3 * it], evs[3 * it + 1]);
        cudaEventElapsedTime(&b2, evs[3 * it + 1], evs[3 * it + 2]);
        spmv_ms += a;
        update_ms += b2;
      }
      spmv_ms /= nprof;
      update_ms /= nprof;
    }
    for (auto& e : evs) cudaEventDestroy(e);
    if (ce != cudaSuccess) rc = lat_cuda_fail(ctx, ce, "PCG profiled iterations", __FILE__, __LINE__);
  }
  const int remaining = o->maxiter - nprof;
  const int nbatch = (int)ceil_div(remaining > 0 ? remaining : 0, check);
  PcgScalars* hs = ctx->h_scal;
  int launched = 0, checked = 0;
  bool finished = (remaining <= 0);
  // software pipeline of depth 2: batch i+1 is enqueued before the status of batch i is read;
  // kernels of a batch enqueued after convergence exit immediately (sc->done), so x is frozen
  // at the converged iterate exactly like the `break` of the reference loop.
  while (!finished && rc == LAT_OK) {
    while (launched < nbatch && launched - checked < 2) {
      ce = cudaGraphLaunch(gexec, ctx->stream);
      if (ce != cudaSuccess) { rc = lat_cuda_fail(ctx, ce, "cudaGraphLaunch", __FILE__, __LINE__); break; }
      ctx->launches += 2 * check;
      cudaMemcpyAsync(&hs[launched & 1], sc, sizeof(PcgScalars), cudaMemcpyDeviceToHost, ctx->stream);
      cudaEventRecord(ctx->ev[2 + (launched & 1)], ctx->stream);
      ++launched;
    }
    if (rc != LAT_OK) break;
    const int s = checked & 1;
    ce = cudaEventSynchronize(ctx->ev[2 + s]);
    if (ce != cudaSuccess) { rc = lat_cuda_fail(ctx, ce, "PCG iteration", __FILE__, __LINE__); break; }
    ++checked;
    if (hs[s].done || hs[s].iters >= o->maxiter || checked == nbatch) finished = true;
  }
  if (rc == LAT_OK) {
    cudaEventRecord(ctx->ev[1], ctx->stream);
    cudaMemcpyAsync(&hs[0], sc, sizeof(PcgScalars), cudaMemcpyDeviceToHost, ctx->stream);
    ce = cudaStreamSynchronize(ctx->stream);
    if (ce != cudaSuccess) rc = lat_cuda_fail(ctx, ce, "PCG final sync", __FILE__, __LINE__);
  }
  cudaGraphExecDestroy(gexec);
  cudaGraphDestroy(graph);
  cudaStreamDestroy(cap);
  if (rc != LAT_OK) return rc;
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
  res->iters = hs[0].iters;
  res->norm_b = sqrt(hs[0].bb);
  res->relres = hs[0].bb > 0.0 ? sqrt(hs[0].rr / hs[0].bb) : 0.0;
  res->info = hs[0].done && !hs[0].breakdown ? 0 : (hs[0].breakdown ? 3 : (hs[0].info_flag2 ? 2 : 1));
  res->solve_ms = ms;
  res->launches = ctx->launches - launches0;
  res->spmv_ms = spmv_ms;
  res->update_ms = update_ms;
  res->profiled = nprof;
  res->reserved = 0;
  return LAT_OK;
}

extern "C" int lat_pcg_bsr(lat_ctx* ctx, const int32_t* rowptr, const int32_t* colidx, const double* vals,
                           int64_t n_nodes, const double* b, double* x, const lat_pcg_opts* opts,
                           lat_pcg_result* result) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, rowptr && colidx && vals && b && x && opts && result && n_nodes > 0);
  LAT_CHECK_ARG(ctx, opts->maxiter >= 0 && opts->tol >= 0.0);
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  switch (opts->precond) {
    case LAT_PC_NONE: return pcg_run<LAT_PC_NONE>(ctx, rowptr, colidx, vals, n_nodes, b, x, opts, result);
    case LAT_PC_JACOBI: return pcg_run<LAT_PC_JACOBI>(ctx, rowptr, colidx, vals, n_nodes, b, x, opts, result);
    case LAT_PC_BLOCK6: return pcg_run<LAT_PC_BLOCK6>(ctx, rowptr, colidx, vals, n_nodes, b, x, opts, result);
    default: return lat_fail(ctx, LAT_ERR_ARG, "unknown preconditioner", __FILE__, __LINE__);
  }
}

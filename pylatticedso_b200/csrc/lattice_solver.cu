// lattice_solver.cu -- BSR(6x6) FP64 SpMV and the two-kernel PCG iteration.  sm_100a.
//
// Thread layout shared by every kernel here: a warp owns 5 consecutive block
// rows (nodes); lane = 6*g + r handles scalar row r of node g (lanes 30, 31
// idle).  DOF index 6*node + r is therefore consecutive across lanes 0..29, so
// all vector traffic is coalesced, and the 6 lanes of a group read one 288 B
// block of the matrix as 6 x 48 B (three 16 B loads per lane).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>

#include <stddef.h>

#include "common.cuh"
#include "matfree.cuh"

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
// "Transposed piece" mapping of a 6x6 block onto the six lanes of a row group (round 2).
// A block is 18 pieces of 16 bytes (piece p = scalar row p/3, column pair p%3).  Lane r of the group loads pieces
// r, 6+r, 12+r: the six lanes together read 96 CONTIGUOUS bytes per instruction (the row-per-lane mapping read
// 6 x 16 B at stride 48 B = three 128-byte lines per block and instruction, which made the product kernels L1-wavefront
// bound: profiles/r02_ncu_persist_before_l2_policy.txt), and every lane needs only ONE 16-byte piece of x (column pair c = r%3)
// instead of all 48 bytes.  The lane accumulates partial sums of scalar rows h, 2+h, 4+h (h = r/3) over the blocks of the
// row; at the end of the row the three lanes with the same h are added in a fixed order and lane (h, c) keeps the
// total of scalar row 2c+h.  16 instead of 24 registers per block in flight, ~6 instead of ~13 L1 wavefronts per block.
// ---------------------------------------------------------------------------
struct Piece3 { double2 a0, a1, a2, x; };
__device__ __forceinline__ int lane_dof(int r) { return 2 * (r % 3) + r / 3; }   // scalar row the lane ends up owning
template <bool STREAM>
__device__ __forceinline__ Piece3 piece_load(const double* __restrict__ vals, int64_t j, int r, const double* xcol, int c) {
  Piece3 q;
  const double2* vp = reinterpret_cast<const double2*>(vals + j * 36 + 2 * r);
  if (STREAM) { q.a0 = __ldcs(vp); q.a1 = __ldcs(vp + 6); q.a2 = __ldcs(vp + 12); }
  else { q.a0 = __ldg(vp); q.a1 = __ldg(vp + 6); q.a2 = __ldg(vp + 12); }
  q.x = *reinterpret_cast<const double2*>(xcol + 2 * c);
  return q;
}
__device__ __forceinline__ void piece_fma(const Piece3& q, double& s0, double& s1, double& s2) {
  s0 = fma(q.a0.x, q.x.x, s0); s0 = fma(q.a0.y, q.x.y, s0);
  s1 = fma(q.a1.x, q.x.x, s1); s1 = fma(q.a1.y, q.x.y, s1);
  s2 = fma(q.a2.x, q.x.x, s2); s2 = fma(q.a2.y, q.x.y, s2);
}
// All 32 lanes call.  Returns the total of the scalar row lane_dof(r) of the lane's node.
__device__ __forceinline__ double piece_finish(int g, int r, double s0, double s1, double s2) {
  const int h = r / 3, c = r - 3 * h;
  const int base = (g < 5 ? g : 4) * 6 + 3 * h;
  const double a0 = __shfl_sync(0xffffffffu, s0, base), a1 = __shfl_sync(0xffffffffu, s0, base + 1), a2 = __shfl_sync(0xffffffffu, s0, base + 2);
  const double b0 = __shfl_sync(0xffffffffu, s1, base), b1 = __shfl_sync(0xffffffffu, s1, base + 1), b2 = __shfl_sync(0xffffffffu, s1, base + 2);
  const double c0 = __shfl_sync(0xffffffffu, s2, base), c1 = __shfl_sync(0xffffffffu, s2, base + 1), c2 = __shfl_sync(0xffffffffu, s2, base + 2);
  const double t0 = (a0 + a1) + a2, t1 = (b0 + b1) + b2, t2 = (c0 + c1) + c2;
  return c == 0 ? t0 : (c == 1 ? t1 : t2);
}
// acc over the blocks [lo, hi) of one block row, three blocks (all loads first) per trip.  NC: gather x past L1.
template <bool STREAM, bool CG_GATHER>
__device__ __forceinline__ void piece_row(const int32_t* colidx, const double* __restrict__ vals, int lo, int hi,
                                          int r, const double* x, double& s0, double& s1, double& s2) {
  const int c = r % 3;
  for (int j = lo; j < hi; j += 3) {
    const bool p1 = j + 1 < hi, p2 = j + 2 < hi;
    const int j1 = p1 ? j + 1 : j, j2 = p2 ? j + 2 : j;
    const int c0 = colidx[j], c1 = colidx[j1], c2 = colidx[j2];   // global (read-only path) or shared memory
    Piece3 q0, q1, q2;
    if (CG_GATHER) {
      q0 = piece_load<STREAM>(vals, j, r, x, 0);  q0.x = __ldcg(reinterpret_cast<const double2*>(x + (int64_t)c0 * 6 + 2 * c));
      q1 = piece_load<STREAM>(vals, j1, r, x, 0); q1.x = __ldcg(reinterpret_cast<const double2*>(x + (int64_t)c1 * 6 + 2 * c));
      q2 = piece_load<STREAM>(vals, j2, r, x, 0); q2.x = __ldcg(reinterpret_cast<const double2*>(x + (int64_t)c2 * 6 + 2 * c));
    } else {
      q0 = piece_load<STREAM>(vals, j, r, x + (int64_t)c0 * 6, c);
      q1 = piece_load<STREAM>(vals, j1, r, x + (int64_t)c1 * 6, c);
      q2 = piece_load<STREAM>(vals, j2, r, x + (int64_t)c2 * 6, c);
    }
    piece_fma(q0, s0, s1, s2);
    if (p1) piece_fma(q1, s0, s1, s2);
    if (p2) piece_fma(q2, s0, s1, s2);
  }
}

// Matrix loads with an explicit L2 eviction policy (createpolicy + ld.global.L2::cache_hint), past L1.
__device__ __forceinline__ double2 ld_policy(const double2* p, unsigned long long pol) {
  double2 v;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ Piece3 piece_load_policy(const double* __restrict__ vals, int64_t j, int r, const double* xcol, int c,
                                                    unsigned long long pol) {
  Piece3 q;
  const double2* vp = reinterpret_cast<const double2*>(vals + j * 36 + 2 * r);
  q.a0 = ld_policy(vp, pol); q.a1 = ld_policy(vp + 6, pol); q.a2 = ld_policy(vp + 12, pol);
  q.x = *reinterpret_cast<const double2*>(xcol + 2 * c);
  return q;
}
__device__ __forceinline__ void piece_row_policy(const int32_t* colidx, const double* __restrict__ vals, int lo, int hi, int r,
                                                 const double* x, double& s0, double& s1, double& s2, unsigned long long pol) {
  const int c = r % 3;
  for (int j = lo; j < hi; j += 3) {
    const bool p1 = j + 1 < hi, p2 = j + 2 < hi;
    const int j1 = p1 ? j + 1 : j, j2 = p2 ? j + 2 : j;
    const int c0 = colidx[j], c1 = colidx[j1], c2 = colidx[j2];
    const Piece3 q0 = piece_load_policy(vals, j, r, x + (int64_t)c0 * 6, c, pol);
    const Piece3 q1 = piece_load_policy(vals, j1, r, x + (int64_t)c1 * 6, c, pol);
    const Piece3 q2 = piece_load_policy(vals, j2, r, x + (int64_t)c2 * 6, c, pol);
    piece_fma(q0, s0, s1, s2);
    if (p1) piece_fma(q1, s0, s1, s2);
    if (p2) piece_fma(q2, s0, s1, s2);
  }
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
  const bool active = g < ROWS_PER_WARP && n < n_nodes;
  int lo = 0, hi = 0;
  if (active) { lo = rowptr[n]; hi = rowptr[n + 1]; }
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  piece_row<true, false>(colidx, vals, lo, hi, r, x, s0, s1, s2);      // matrix streamed once: evict-first
  const double tot = piece_finish(g, r, s0, s1, s2);
  if (active) y[n * 6 + lane_dof(r)] = tot;
}

int lat_spmv_internal(lat_ctx* ctx, const int32_t* rowptr, const int32_t* colidx, const double* vals,
                      int64_t n_nodes, const double* x, double* y) {
  // one-shot products (lifting, reactions) use the direct-load kernel: no partition / host sync needed
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
// dinv layout: Jacobi -> [6n] reciprocal diagonal; block-Jacobi -> [n][21] packed upper triangle of the
// (symmetric) inverse of the diagonal block.
// Invert one diagonal block and store it in the dinv layout above.
__device__ __forceinline__ void precond_store(double (&a)[6][6], int64_t n, int precond, double* __restrict__ dinv) {
  if (precond == LAT_PC_JACOBI) {
#pragma unroll
    for (int r = 0; r < 6; ++r) dinv[n * 6 + r] = (a[r][r] > 0.0) ? 1.0 / a[r][r] : 1.0;
    return;
  }
  // 6x6 SPD inverse by Gauss-Jordan without pivoting
  double inv[6][6];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int k = 0; k < 6; ++k) inv[i][k] = (i == k) ? 1.0 : 0.0;
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
  // the inverse of an SPD block is symmetric: store the 21 upper-triangular entries of (inv + inv^T)/2,
  // packed row-major (entry (i,j), i <= j, at i*(11-i)/2 + j) -- 168 instead of 288 B per node per iteration
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int k = i; k < 6; ++k)
      dinv[n * 21 + (i * (11 - i)) / 2 + k] = ok ? 0.5 * (inv[i][k] + inv[k][i]) : (i == k ? 1.0 : 0.0);
}

__global__ void k_precond_setup(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                const double* __restrict__ vals, int64_t n_nodes, int precond,
                                double* __restrict__ dinv) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_nodes) return;
  int d = -1;
  for (int j = rowptr[n]; j < rowptr[n + 1]; ++j)
    if (colidx[j] == n) { d = j; break; }
  double a[6][6];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int k = 0; k < 6; ++k) a[i][k] = d >= 0 ? vals[(int64_t)d * 36 + i * 6 + k] : (i == k ? 1.0 : 0.0);
  precond_store(a, n, precond, dinv);
}

// matrix-free operator: the diagonal blocks are regenerated from the geometry
__global__ void k_mf_precond(MfOp op, int64_t n_nodes, int precond, double* __restrict__ dinv) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_nodes) return;
  double a[6][6];
  mf_diag_block(op, n, a);
  precond_store(a, n, precond, dinv);
}

// The preconditioner entries of the lane's row can be fetched BEFORE the residual is known: doing so
// explicitly puts them in the same memory round trip as the vector loads (nvcc only hoisted half of them).
template <int PC>
struct PrecondRow {
  double d[PC == LAT_PC_BLOCK6 ? 6 : 1];
};
template <int PC>
__device__ __forceinline__ PrecondRow<PC> load_precond(const double* __restrict__ dinv, int64_t n, int r, bool active) {
  PrecondRow<PC> p;
  if (PC == LAT_PC_JACOBI) p.d[0] = active ? __ldg(dinv + n * 6 + r) : 0.0;
  if (PC == LAT_PC_BLOCK6) {
    const double* blk = dinv + n * 21;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const int i = r < k ? r : k, j = r < k ? k : r;
      p.d[k] = active ? __ldg(blk + (i * (11 - i)) / 2 + j) : 0.0;
    }
  }
  return p;
}
// z_r = (M^-1 r)_r for the lane's DOF; rv = the lane's residual entry.  All 32 lanes must call.
template <int PC>
__device__ __forceinline__ double apply_precond(const PrecondRow<PC>& p, int g, double rv) {
  if (PC == LAT_PC_NONE) return rv;
  if (PC == LAT_PC_JACOBI) return p.d[0] * rv;
  double z = 0.0;
#pragma unroll
  for (int k = 0; k < 6; ++k) z = fma(p.d[k], __shfl_sync(0xffffffffu, rv, g * 6 + k), z);
  return z;
}
template <int PC>
__device__ __forceinline__ double apply_precond(const double* __restrict__ dinv, int64_t n, int g, int r,
                                                bool active, double rv) {
  const PrecondRow<PC> p = load_precond<PC>(dinv, n, r, active);
  return apply_precond<PC>(p, g, rv);
}

// ---------------------------------------------------------------------------
// PCG kernels
// ---------------------------------------------------------------------------
struct PcgParams {
  double tol, mintol, alpha_max;
  int64_t restart_every;
  int32_t maxiter, reference;
  int32_t dist, l2_keep;  // dist = 1: kernels only store LOCAL sums; the host all-reduces and runs k_pcg_finalize_*
                          // l2_keep = k > 0 (k_cg_spmv<.., .., true>): the matrix blocks of k/16 of the rows stay L2-resident
  unsigned long long seq_base;  // peer-memory path: (solve epoch << 32), so flags of earlier solves never match
  unsigned long long push_base; // fused-halo path: halo pushes completed by earlier solves on this arena
};

// Stop tests, beta and bookkeeping of one iteration from the GLOBAL sums (single thread).
__device__ __forceinline__ void pcg_finish_iteration(PcgScalars* sc, const PcgParams& prm, double rz_new, double rr,
                                                     double xx, double pnorm) {
  const int k = sc->iters;
  const double pAp = sc->pAp;
  double alpha = sc->rz_old / pAp;
  if (prm.reference && prm.alpha_max > 0.0) alpha = fmin(alpha, prm.alpha_max);
  const bool restart = prm.reference && prm.restart_every > 0 && k > 0 && (k % prm.restart_every) == 0;
  const double pp = restart ? pnorm : sc->pp;
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

__device__ __forceinline__ void pcg_finish_init(PcgScalars* sc, double rz, double bb) {
  sc->rz_old = rz;
  sc->bb = bb;
  sc->rr = bb;
  sc->beta = 0.0;
  sc->pAp = 0.0; sc->pp = 0.0; sc->xx = 0.0; sc->alpha_last = 0.0;
  sc->iters = 0; sc->info_flag2 = 0; sc->breakdown = 0;
  sc->done = (bb == 0.0) ? 1 : 0;   // b == 0 -> x = 0
}

__global__ void k_pcg_finalize_init(PcgScalars* sc) { pcg_finish_init(sc, sc->sums[0], sc->sums[1]); }
__global__ void k_pcg_finalize_update(PcgScalars* sc, PcgParams prm) {
  if (sc->done || sc->iters >= prm.maxiter) return;
  pcg_finish_iteration(sc, prm, sc->sums[0], sc->sums[1], sc->sums[2], sc->sums[3]);
}
// ghost entries of the search direction: p_new = z + beta p_old on [first, last)
__global__ void k_pcg_ghost_p(int64_t first, int64_t last, const double* __restrict__ z, double* __restrict__ pa,
                              double* __restrict__ pb, const PcgScalars* __restrict__ sc, PcgParams prm) {
  if (sc->done || sc->iters >= prm.maxiter) return;
  const int k = sc->iters;
  const double beta = sc->beta;
  const double* p_old = (k & 1) ? pb : pa;
  double* p_new = (k & 1) ? pa : pb;
  const int64_t i = first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < last) p_new[i] = fma(beta, p_old[i], z[i]);
}
__global__ void k_pack_halo(const int32_t* __restrict__ idx, int64_t n, const double* __restrict__ v, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n * 6) out[i] = v[(int64_t)idx[i / 6] * 6 + (i % 6)];
}

// init: x = 0, r = b, z = M^-1 r, p[0] = z; rz_old = r.z, bb = b.b
template <int PC>
__global__ void __launch_bounds__(SPMV_BLOCK) k_pcg_init(int64_t n_nodes, const double* __restrict__ b,
                                                         const double* __restrict__ dinv, double* __restrict__ x,
                                                         double* __restrict__ r, double* __restrict__ z,
                                                         double* __restrict__ p0, double* __restrict__ p1,
                                                         PcgScalars* __restrict__ sc, double* __restrict__ partials,
                                                         int dist) {
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
    if (dist) { sc->sums[0] = out[0]; sc->sums[1] = out[1]; }
    else pcg_finish_init(sc, out[0], out[1]);
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
    if (prm.dist) { sc->sums[0] = out[0]; sc->sums[1] = out[1]; sc->sums[2] = out[2]; sc->sums[3] = out[3]; }
    else pcg_finish_iteration(sc, prm, out[0], out[1], out[2], out[3]);
  }
}

// ---------------------------------------------------------------------------
// Chronopoulos-Gear PCG (default for the textbook mode): ONE gather and ONE global reduction per
// iteration.   u = M^-1 r,  w = A u,  gamma = (r,u),  delta = (w,u)
//   beta = gamma/gamma_old,  alpha = gamma / (delta - beta*gamma/alpha_old)
//   p = u + beta p,  s = w + beta s,  x += alpha p,  r -= alpha s
// Kernel A (k_cg_spmv): w = A u with the three dots and, in the last block, the scalar recurrences
// and the stop test.  Kernel B (k_cg_update): the five vector updates and the preconditioner, no
// reduction at all.  Mathematically the same Krylov iteration as classic PCG; the classic two-reduction
// form (k_pcg_spmv / k_pcg_update) is kept for reference_semantics = 1, whose clamp / restart rules
// are defined on the classic recurrences.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void cg_finish(PcgScalars* sc, const PcgParams& prm, double gamma, double delta, double rr) {
  sc->rr = rr;
  if (sc->first) {  // set-up pass: w0 = A u0
    if (sc->first == 1) sc->bb = rr;   // r0 = b (first == 2: restart from x, |b| is kept)
    else if (rr <= prm.tol * prm.tol * sc->bb) { sc->first = 0; sc->done = 1; return; }
    sc->first = 0;
    sc->beta = 0.0;
    sc->gamma = gamma;
    sc->alpha = gamma / delta;
    if (rr == 0.0) sc->done = 1;
    else if (!(delta > 0.0)) { sc->done = 1; sc->breakdown = 1; }
    return;
  }
  sc->iters += 1;
  if (rr <= prm.tol * prm.tol * sc->bb) { sc->done = 1; return; }
  const double beta = gamma / sc->gamma;
  const double denom = delta - beta * gamma / sc->alpha;
  if (!(denom > 0.0) || !(rr == rr)) { sc->done = 1; sc->breakdown = 1; return; }
  sc->beta = beta;
  sc->alpha = gamma / denom;
  sc->gamma = gamma;
}

__global__ void k_cg_finalize(PcgScalars* sc, PcgParams prm) {
  if (sc->done || sc->iters >= prm.maxiter) return;
  cg_finish(sc, prm, sc->sums[0], sc->sums[1], sc->sums[2]);
}

// Which rows a product kernel covers.  Default: all of [0, n_nodes).  Multi-GPU overlap: the INTERIOR launch
// passes skip[] (1 = row reads a ghost column: left to the boundary launch, which waits for the halo), the
// BOUNDARY launch passes the list of those rows.  p_stride / p_offset place the per-CTA partial sums of both
// launches in one array.
struct RowSet {
  const int32_t* rows = nullptr;   // non-null: n_nodes entries, row = rows[i]
  const uint8_t* skip = nullptr;   // non-null: rows with skip[row] != 0 are not touched
  int p_stride = 0, p_offset = 0;
  // fused halo (peer-memory path): need[row] = bit mask of the neighbours whose ghosts the row reads.  A CTA
  // that owns such a row waits (bounded) until the neighbour's cumulative entry counter in OUR arena reaches
  // (push_base + seq + 1) * per_push[k]; all other CTAs start at once.  Ghost columns (>= n_own) are then read
  // past L1 (a line straddling the owned/ghost border may have been cached before the data arrived).
  const uint8_t* need = nullptr;
  const uint8_t* cta_need = nullptr;          // OR of need[] over the rows of each CTA (uniform test, no barrier)
  const unsigned long long* cnt = nullptr;    // my arena: halo_cnt[source rank]
  int n_nb = 0, nb_rank[2] = {0, 0};
  unsigned long long per_push[2] = {0, 0};
  int64_t n_own = INT64_MAX;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// All threads of a CTA whose mask `need` (CTA-uniform) is non-zero call.
__device__ __forceinline__ void halo_wait(const RowSet& rs, int need, PcgScalars* sc, const PcgParams& prm) {
  if ((int)threadIdx.x < rs.n_nb && ((need >> threadIdx.x) & 1)) {
    // (constant indices only: a dynamically indexed kernel-parameter array would be copied to local memory)
    const unsigned long long per = threadIdx.x == 0 ? rs.per_push[0] : rs.per_push[1];
    const int src = threadIdx.x == 0 ? rs.nb_rank[0] : rs.nb_rank[1];
    const unsigned long long want = (prm.push_base + (unsigned long long)sc->seq + 1ull) * per;
    long long spins = 0;
    while (ld_acquire_sys_u64(rs.cnt + src) < want) {
      if (++spins > (1ll << 24)) { sc->p2p_timeout = 1; break; }
      __nanosleep(20);
    }
  }
  __syncthreads();
}

// Fused halo push (peer-memory path): the kernel that produces u stores the boundary entries straight into the
// neighbours' ghost sections and adds the number of entries to the neighbour's counter -- no halo kernel.
struct HaloPush {
  const int2* dst = nullptr;                  // per owned node: destination node in neighbour 0 / 1 (-1: none)
  const uint8_t* cta_push = nullptr;          // per CTA of the update kernel: does any of its rows push?
  double* peer_u[2] = {nullptr, nullptr};
  unsigned long long* peer_cnt[2] = {nullptr, nullptr};   // &peer_arena->halo_cnt[my_rank]
};
// All threads of the CTA call (two CTA-wide counts).  d = the thread's destinations, comp its DOF, val its entry.
__device__ __forceinline__ void halo_push(const HaloPush& hp, int2 d, int comp, double val) {
  if (d.x >= 0) hp.peer_u[0][(int64_t)d.x * 6 + comp] = val;
  if (d.y >= 0) hp.peer_u[1][(int64_t)d.y * 6 + comp] = val;
  const int c0 = __syncthreads_count(d.x >= 0);   // barrier: the CTA's peer stores happen-before thread 0's release
  const int c1 = __syncthreads_count(d.y >= 0);
  if (threadIdx.x == 0 && (c0 | c1)) {
    asm volatile("fence.acq_rel.sys;" ::: "memory");
    if (c0) asm volatile("red.relaxed.sys.global.add.u64 [%0], %1;" ::"l"(hp.peer_cnt[0]), "l"((unsigned long long)c0) : "memory");
    if (c1) asm volatile("red.relaxed.sys.global.add.u64 [%0], %1;" ::"l"(hp.peer_cnt[1]), "l"((unsigned long long)c1) : "memory");
  }
}

// GHOST = true is the fused-halo instantiation (waits, ghost-reading rows load their columns past L1); the
// default instantiation carries none of it -- a per-column branch in the gather loop cost 25 % on one GPU.
// SUBSET = true enables rs.rows / rs.skip (opt-in overlap path); without it the row extent and own entries
// are loaded before the status word is looked at -- a skip[] load in front of them serialised one more
// memory round trip per CTA (+9 % at 15 waves of CTAs).
template <bool GHOST = false, bool SUBSET = false, bool L2KEEP = false>
__global__ void __launch_bounds__(SPMV_BLOCK) k_cg_spmv(
const int32_t* __restrict__ rowptr,
                                                        const int32_t* __restrict__ colidx,
                                                        const double* __restrict__ vals, int64_t n_nodes,
                                                        const double* __restrict__ u, const double* __restrict__ r,
                                                        double* __restrict__ w, PcgScalars* __restrict__ sc,
                                                        double* __restrict__ partials, PcgParams prm,
                                                        RowSet rs = RowSet()) {
  const int lane = threadIdx.x & 31;
  const int g = lane / 6, rr_ = lane - g * 6;
  const int64_t warp = (int64_t)blockIdx.x * (SPMV_BLOCK / 32) + (threadIdx.x >> 5);
  int64_t n = warp * ROWS_PER_WARP + g;
  bool active = g < ROWS_PER_WARP && n < n_nodes;
  if (SUBSET) {
    if (rs.rows) n = active ? rs.rows[n] : 0;
    if (rs.skip && active && rs.skip[n]) active = false;
  }
  // independent loads first: the row extent and the own-row entries do not depend on the status word
  int lo = 0, hi = 0;
  double uo = 0.0, ro = 0.0;
  const int64_t i = n * 6 + lane_dof(rr_);     // the scalar row this lane owns in the transposed-piece mapping
  if (active) { lo = __ldg(rowptr + n); hi = __ldg(rowptr + n + 1); uo = u[i]; ro = r[i]; }
  if (sc->done || sc->iters >= prm.maxiter) return;
  int need_row = 0;
  if (GHOST) {
    const int cta_need = rs.cta_need[blockIdx.x];
    if (cta_need) {
      need_row = active ? rs.need[n] : 0;
      halo_wait(rs, cta_need, sc, prm);
    }
  }
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  if (GHOST && need_row) piece_row<true, true>(colidx, vals, lo, hi, rr_, u, s0, s1, s2);   // ghost-reading row: gathers past L1
  else if (L2KEEP) {
    // matrices of up to a few L2 sizes: ~44 MB of the blocks (the same 16-row classes in every iteration) are loaded
    // evict_last and stay in L2 between iterations, the rest streams evict_first (as in k_pcg_persist)
    unsigned long long pol;
    if (((n >> 4) & 15) < prm.l2_keep) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    piece_row_policy(colidx, vals, lo, hi, rr_, u, s0, s1, s2, pol);
  } else piece_row<true, false>(colidx, vals, lo, hi, rr_, u, s0, s1, s2);
  const double acc = piece_finish(g, rr_, s0, s1, s2);
  if (active) w[i] = acc;
  double v[3] = {ro * uo, acc * uo, ro * ro};
  block_partials<3, SPMV_BLOCK>(v, partials, rs.p_stride, rs.p_offset);
}

// Matrix-free twin of k_cg_spmv: w = A u regenerated from the geometry (matfree.cuh, one thread per node),
// same three dots.
template <bool GHOST = false, bool SUBSET = false>
__global__ void __launch_bounds__(MF_BLOCK, 9) k_cg_spmv_mf(MfOp op, int64_t n_nodes, const double* __restrict__ u,
                                                         const double* __restrict__ r, double* __restrict__ w,
                                                         PcgScalars* __restrict__ sc, double* __restrict__ partials,
                                                         PcgParams prm, RowSet rs = RowSet()) {
  int64_t n = (int64_t)blockIdx.x * MF_BLOCK + threadIdx.x;
  bool active = n < n_nodes;
  if (SUBSET) {
    if (rs.rows) n = active ? rs.rows[n] : 0;
    if (rs.skip && active && rs.skip[n]) active = false;
  }
  MfU rr;
  rr.a = rr.b = rr.c = make_double2(0.0, 0.0);
  if (active) rr = mf_load_u(r, n);
  if (sc->done || sc->iters >= prm.maxiter) return;
  int need_row = 0;
  if (GHOST) {
    const int cta_need = rs.cta_need[blockIdx.x];
    if (cta_need) {
      need_row = active ? rs.need[n] : 0;
      halo_wait(rs, cta_need, sc, prm);
    }
  }
  double v[3] = {0.0, 0.0, 0.0};
  if (active) {
    double uo[6], f[6];
    if (GHOST && need_row) mf_node_product<true, true>(op, n, u, uo, f);   // reads ghosts: gathers past L1
    else mf_node_product<true, false>(op, n, u, uo, f);
    mf_store6(w, n, f);
    const double ro[6] = {rr.a.x, rr.a.y, rr.b.x, rr.b.y, rr.c.x, rr.c.y};
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      v[0] = fma(ro[k], uo[k], v[0]);
      v[1] = fma(f[k], uo[k], v[1]);
      v[2] = fma(ro[k], ro[k], v[2]);
    }
  }
  block_partials<3, MF_BLOCK>(v, partials, rs.p_stride, rs.p_offset);
}

static constexpr int CG_REDUCE_BLOCK = 512;
// One CTA: fixed-order sum of the per-CTA partials, then the scalar recurrences / stop test.
__global__ void __launch_bounds__(CG_REDUCE_BLOCK) k_cg_reduce(const double* __restrict__ partials, int n_part,
                                                               PcgScalars* __restrict__ sc, PcgParams prm) {
  double out[3];
  sum_partials<3, CG_REDUCE_BLOCK>(partials, n_part, out);   // issued before the (dependent) status check
  if (sc->done || sc->iters >= prm.maxiter) return;
  if (threadIdx.x == 0) {
    if (prm.dist) { sc->sums[0] = out[0]; sc->sums[1] = out[1]; sc->sums[2] = out[2]; }
    else cg_finish(sc, prm, out[0], out[1], out[2]);
  }
}

template <int PC>
__global__ void __launch_bounds__(SPMV_BLOCK) k_cg_update(int64_t n_nodes, const double* __restrict__ dinv,
                                                          double* __restrict__ x, double* __restrict__ r,
                                                          double* __restrict__ u, const double* __restrict__ w,
                                                          double* __restrict__ p, double* __restrict__ s,
                                                          const PcgScalars* __restrict__ sc, PcgParams prm,
                                                          HaloPush hp = HaloPush()) {
  const int lane = threadIdx.x & 31;
  const int g = lane / 6, rr_ = lane - g * 6;
  const int64_t warp = (int64_t)blockIdx.x * (SPMV_BLOCK / 32) + (threadIdx.x >> 5);
  const int64_t n = warp * ROWS_PER_WARP + g;
  const bool active = g < ROWS_PER_WARP && n < n_nodes;
  const int64_t i = n * 6 + rr_;
  double uv = 0.0, wv = 0.0, pv = 0.0, sv = 0.0, xv = 0.0, rv = 0.0;
  if (active) { uv = u[i]; wv = w[i]; pv = p[i]; sv = s[i]; xv = x[i]; rv = r[i]; }
  int2 dst = make_int2(-1, -1);
  const bool pushes = hp.dst && hp.cta_push[blockIdx.x];   // CTA-uniform
  if (pushes && active) dst = hp.dst[n];
  const PrecondRow<PC> pr = load_precond<PC>(dinv, n, rr_, active);   // same round trip as the vectors
  const int done = sc->done, iters = sc->iters;
  const double alpha = sc->alpha, beta = sc->beta;
  if (done || iters >= prm.maxiter) return;
  pv = fma(beta, pv, uv);
  sv = fma(beta, sv, wv);
  xv = fma(alpha, pv, xv);
  rv = fma(-alpha, sv, rv);
  const double zn = apply_precond<PC>(pr, g, rv);
  if (active) { p[i] = pv; s[i] = sv; x[i] = xv; r[i] = rv; u[i] = zn; }
  if (pushes) halo_push(hp, dst, rr_, zn);
}

// init for the CG variant: x = 0, r = b, u = M^-1 b, p = s = 0
template <int PC>
__global__ void __launch_bounds__(SPMV_BLOCK) k_cg_init(int64_t n_nodes, const double* __restrict__ b,
                                                        const double* __restrict__ dinv, double* __restrict__ x,
                                                        double* __restrict__ r, double* __restrict__ u,
                                                        double* __restrict__ p, double* __restrict__ s) {
  const int lane = threadIdx.x & 31;
  const int g = lane / 6, rr_ = lane - g * 6;
  const int64_t warp = (int64_t)blockIdx.x * (SPMV_BLOCK / 32) + (threadIdx.x >> 5);
  const int64_t n = warp * ROWS_PER_WARP + g;
  const bool active = g < ROWS_PER_WARP && n < n_nodes;
  const int64_t i = n * 6 + rr_;
  const double bv = active ? b[i] : 0.0;
  const double zv = apply_precond<PC>(dinv, n, g, rr_, active, bv);
  if (active) { x[i] = 0.0; r[i] = bv; u[i] = zv; p[i] = 0.0; s[i] = 0.0; }
}

// ---------------------------------------------------------------------------
// True-residual safeguard of the Chronopoulos-Gear form.  The recurrences r -= alpha s, s = w + beta s
// can drift from b - A x (and a lost halo would break them silently), so a converged solve is verified:
// r_true = b - A x; if |r_true| > 2 tol |b| the iteration restarts from x with r = r_true (at most twice).
// ---------------------------------------------------------------------------
template <int PC>
__global__ void __launch_bounds__(SPMV_BLOCK) k_cg_restart(int64_t n_nodes, const double* __restrict__ b,
                                                           const double* __restrict__ Ax, const double* __restrict__ dinv,
                                                           double* __restrict__ r, double* __restrict__ u,
                                                           double* __restrict__ p, double* __restrict__ s,
                                                           double* __restrict__ partials) {
  const int lane = threadIdx.x & 31;
  const int g = lane / 6, rr_ = lane - g * 6;
  const int64_t warp = (int64_t)blockIdx.x * (SPMV_BLOCK / 32) + (threadIdx.x >> 5);
  const int64_t n = warp * ROWS_PER_WARP + g;
  const bool active = g < ROWS_PER_WARP && n < n_nodes;
  const int64_t i = n * 6 + rr_;
  const double rv = active ? b[i] - Ax[i] : 0.0;
  const double zv = apply_precond<PC>(dinv, n, g, rr_, active, rv);
  if (active) { r[i] = rv; u[i] = zv; p[i] = 0.0; s[i] = 0.0; }
  double v[1] = {rv * rv};
  block_partials<1, SPMV_BLOCK>(v, partials);
}

// One CTA: |r_true|^2 (local, or local part awaiting an all-reduce when `dist`) and the accept/restart decision.
__global__ void __launch_bounds__(CG_REDUCE_BLOCK) k_cg_true_residual(const double* __restrict__ partials, int n_part,
                                                                      PcgScalars* __restrict__ sc, PcgParams prm, int decide) {
  double out[1];
  if (decide == 0 || decide == 2) sum_partials<1, CG_REDUCE_BLOCK>(partials, n_part, out);
  if (threadIdx.x != 0) return;
  if (decide == 0) { sc->sums[3] = out[0]; return; }           // dist: local sum only, all-reduced by the host
  const double rr = (decide == 2) ? out[0] : sc->sums[3];
  sc->true_rr = rr;
  if (rr > 4.0 * prm.tol * prm.tol * sc->bb) {  // |r_true| > 2 tol |b|
    if (sc->restarts < 2 && sc->iters < prm.maxiter) {
      sc->done = 0;
      sc->first = 2;
      sc->restarts += 1;
    } else {
      sc->breakdown = 3;   // restart budget exhausted: reported as info = 5, never as "converged"
    }
  }
}

// ---------------------------------------------------------------------------
// TMA-staged SpMV (sm_90+ bulk async copy, used on sm_100a)
// ---------------------------------------------------------------------------
// A CTA owns a contiguous range of block rows whose matrix blocks form ONE contiguous
// byte range of `vals` (<= cap_blocks * 288 B).  One elected thread arms an mbarrier
// and issues cp.async.bulk (global -> shared) for the whole range; the other threads
// meanwhile fetch the column indices.  With 4 CTAs resident per SM, ~200 KB of matrix
// per SM are in flight while other CTAs compute, so HBM latency is covered by
// CTA-level parallelism instead of per-thread dependent loads (the v1 kernel issued
// rowptr -> colidx -> vals -> vector loads back to back in every thread).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// row partition: CTA i covers rows [cta_row0[i], cta_row0[i+1]) = the rows whose first block index
// falls in [i*T, (i+1)*T)  ->  at most T + max_row_len - 1 blocks per CTA.
__global__ void k_spmv_partition(const int32_t* __restrict__ rowptr, int64_t n_nodes, int T, int G,
                                 int32_t* __restrict__ cta_row0, int32_t* __restrict__ max_nb) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= G) {
    const int64_t target = i * (int64_t)T;
    int64_t lo = 0, hi = n_nodes;  // first row r with rowptr[r] >= target
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (rowptr[mid] < target) lo = mid + 1; else hi = mid;
    }
    cta_row0[i] = (i == G) ? (int32_t)n_nodes : (int32_t)lo;
  }
  // longest row (grid-stride)
  int m = 0;
  for (int64_t r = i; r < n_nodes; r += (int64_t)gridDim.x * blockDim.x) m = max(m, rowptr[r + 1] - rowptr[r]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(max_nb, m);
}

static constexpr int TMA_T = 160;  // target blocks per CTA (46 KB of matrix)

// MODE 0: y = A x.   MODE 1: PCG kernel 1 (p_new = z + beta p_old; Ap; p.Ap, p.p)
template <int MODE>
__global__ void __launch_bounds__(SPMV_BLOCK, 4)
k_spmv_tma(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, const double* __restrict__ vals,
           int64_t n_nodes, const int32_t* __restrict__ cta_row0, int cap_blocks, const double* __restrict__ xz,
           double* __restrict__ pa, double* __restrict__ pb, double* __restrict__ y, PcgScalars* __restrict__ sc,
           double* __restrict__ partials, PcgParams prm) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_bar;
  double* s_vals = reinterpret_cast<double*>(smem_raw);
  int32_t* s_col = reinterpret_cast<int32_t*>(smem_raw + (size_t)cap_blocks * 288);
  int k = 0;
  double beta = 0.0;
  const double* __restrict__ p_old = nullptr;
  double* __restrict__ p_new = nullptr;
  if (MODE == 1) {
    if (sc->done || sc->iters >= prm.maxiter) return;
    k = sc->iters;
    beta = sc->beta;
    p_old = (k & 1) ? pb : pa;
    p_new = (k & 1) ? pa : pb;
  }
  const int r0 = cta_row0[blockIdx.x], r1 = cta_row0[blockIdx.x + 1];
  const int b0 = rowptr[r0], b1 = rowptr[r1];
  const int nblk = b1 - b0;
  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    if (nblk > 0) {
      const uint32_t bytes = (uint32_t)nblk * 288u;
      mbar_expect_tx(&s_bar, bytes);
      const unsigned char* src = reinterpret_cast<const unsigned char*>(vals + (int64_t)b0 * 36);
      unsigned char* dst = smem_raw;
      for (uint32_t off = 0; off < bytes; off += 32768u) {
        const uint32_t n = min(32768u, bytes - off);
        bulk_g2s(dst + off, src + off, n, &s_bar);
      }
    }
  }
  for (int i = threadIdx.x; i < nblk; i += SPMV_BLOCK) s_col[i] = __ldg(colidx + b0 + i);
  __syncthreads();
  if (nblk > 0) mbar_wait(&s_bar, 0);

  const int lane = threadIdx.x & 31;
  const int g = lane / 6, r = lane - g * 6;
  double dot_pAp = 0.0, dot_pp = 0.0;
  for (int base = r0 + (threadIdx.x >> 5) * ROWS_PER_WARP; base < r1; base += ROWS_PER_CTA) {
    const int n = base + g;
    if (g < ROWS_PER_WARP && n < r1) {
      const int lo = rowptr[n] - b0, hi = rowptr[n + 1] - b0;
      double acc = 0.0;
#pragma unroll 2
      for (int j = lo; j < hi; ++j) {
        const int c = s_col[j];
        const double2* vp = reinterpret_cast<const double2*>(s_vals + j * 36 + r * 6);
        const double2 a0 = vp[0], a1 = vp[1], a2 = vp[2];
        double2 x0, x1, x2;
        if (MODE == 1) {
          const double2* zp = reinterpret_cast<const double2*>(xz + (int64_t)c * 6);
          const double2* pp = reinterpret_cast<const double2*>(p_old + (int64_t)c * 6);
          const double2 z0 = zp[0], z1 = zp[1], z2 = zp[2];
          const double2 q0 = pp[0], q1 = pp[1], q2 = pp[2];
          x0.x = fma(beta, q0.x, z0.x); x0.y = fma(beta, q0.y, z0.y);
          x1.x = fma(beta, q1.x, z1.x); x1.y = fma(beta, q1.y, z1.y);
          x2.x = fma(beta, q2.x, z2.x); x2.y = fma(beta, q2.y, z2.y);
        } else {
          const double2* xp = reinterpret_cast<const double2*>(xz + (int64_t)c * 6);
          x0 = __ldg(xp); x1 = __ldg(xp + 1); x2 = __ldg(xp + 2);
        }
        acc = dot6(a0, a1, a2, x0, x1, x2, acc);
      }
      const int64_t i = (int64_t)n * 6 + r;
      if (MODE == 1) {
        const double pv = fma(beta, p_old[i], xz[i]);
        p_new[i] = pv;
        dot_pAp = fma(pv, acc, dot_pAp);
        dot_pp = fma(pv, pv, dot_pp);
      }
      y[i] = acc;
    }
  }
  if (MODE == 1) {
    double v[2] = {dot_pAp, dot_pp}, out[2];
    if (grid_reduce<2, SPMV_BLOCK>(v, partials, &sc->counter[1], out)) {
      sc->pAp = out[0];
      sc->pp = out[1];
    }
  }
}

struct SpmvPlan {
  bool tma = false;
  int grid = 0, cap_blocks = 0, smem = 0;
  int32_t* cta_row0 = nullptr;
};

// Builds (or reuses) the row partition of a BSR matrix for the TMA kernels.  [syncs once per new matrix]
static int spmv_plan(lat_ctx* ctx, const int32_t* rowptr, int64_t n_nodes, SpmvPlan* plan) {
  plan->tma = false;
  int32_t* meta = lat_buf<int32_t>(ctx, "spmv_meta", 8);
  if (!meta) return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
  int32_t nnzb32 = 0;
  LAT_CUDA(ctx, cudaMemsetAsync(meta, 0, 8 * sizeof(int32_t), ctx->stream));
  LAT_CUDA(ctx, cudaMemcpyAsync(&nnzb32, rowptr + n_nodes, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  LAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const int64_t nnzb = nnzb32;
  if (nnzb <= 0) return LAT_OK;
  const int G = (int)ceil_div(nnzb, TMA_T);
  int32_t* cta_row0 = lat_buf<int32_t>(ctx, "spmv_cta_row0", (size_t)G + 2);
  if (!cta_row0) return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
  const int64_t work = n_nodes > G + 1 ? n_nodes : G + 1;
  int64_t pg = ceil_div(work, 256);
  if (pg > 4096) pg = 4096;
  if (pg * 256 < G + 1) pg = ceil_div(G + 1, 256);
  LAT_LAUNCH(ctx, k_spmv_partition, (unsigned)pg, 256, 0, rowptr, n_nodes, TMA_T, G, cta_row0, meta);
  int32_t max_nb = 0;
  LAT_CUDA(ctx, cudaMemcpyAsync(&max_nb, meta, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  LAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const int cap = TMA_T + max_nb;
  const int smem = ((cap * 292 + 15) / 16) * 16;
  if (smem > 200 * 1024) return LAT_OK;  // pathological row length: keep the direct-load kernels
  static bool attr_done = false;
  if (!attr_done || true) {
    LAT_CUDA(ctx, cudaFuncSetAttribute(k_spmv_tma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    LAT_CUDA(ctx, cudaFuncSetAttribute(k_spmv_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }
  plan->tma = true;
  plan->grid = G;
  plan->cap_blocks = cap;
  plan->smem = smem;
  plan->cta_row0 = cta_row0;
  return LAT_OK;
}

// ---------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------
#include "coarse.cuh"

static CoarseLaunch coarse_launch(lat_ctx* ctx) {
  CoarseLaunch cl;
  const CoarseSpace& cs = ctx->coarse;
  cl.n_agg = cs.n_agg;
  cl.n_pieces = cs.n_pieces;
  cl.fused = cs.fused;
  cl.agg_ptr = (const int32_t*)ctx->bufs["coarse_ptr"].p;
  cl.piece_ptr = (const int32_t*)ctx->bufs["coarse_piece_ptr"].p;
  cl.piece_agg = (const int32_t*)ctx->bufs["coarse_piece_agg"].p;
  cl.agg_piece = (const int32_t*)ctx->bufs["coarse_agg_piece"].p;
  cl.nodes = (const CoarseNode*)ctx->bufs["coarse_nodes"].p;
  cl.einv = cs.einv;
  cl.part = (double*)ctx->bufs["coarse_part"].p;
  cl.rc = (double*)ctx->bufs["coarse_rc"].p;
  cl.yc = (double*)ctx->bufs["coarse_yc"].p;
  return cl;
}

template <int PC>
static int pcg_run(lat_ctx* ctx, const int32_t* rowptr, const int32_t* colidx, const double* vals,
                   int64_t n_nodes, const double* b, double* x, const lat_pcg_opts* o, lat_pcg_result* res,
                   const MfOp* mf = nullptr) {
  const int64_t n = 6 * n_nodes;
  const unsigned grid = (unsigned)ceil_div(n_nodes, ROWS_PER_CTA);
  const unsigned mf_grid = (unsigned)ceil_div(n_nodes, MF_BLOCK);   // matrix-free product: one thread per node
  // The TMA-staged SpMV is opt-in (bit 1 of `reserved`): on B200 it measured 5-10 % SLOWER than the
  // direct-load kernel (tools/ab_spmv.py, profiles/r01_spmv_ab.txt) because the kernel is bound by
  // L2->SM traffic of the vector gathers, not by HBM latency.
  SpmvPlan plan;
  if ((o->reserved & 2) && !mf) {
    const int prc = spmv_plan(ctx, rowptr, n_nodes, &plan);
    if (prc) return prc;
  }
  const size_t max_grid = (plan.tma && (unsigned)plan.grid > grid) ? (size_t)plan.grid : (size_t)grid;
  double* r = lat_buf<double>(ctx, "pcg_r", n);
  double* z = lat_buf<double>(ctx, "pcg_z", n);
  double* pa = lat_buf<double>(ctx, "pcg_pa", n);
  double* pb = lat_buf<double>(ctx, "pcg_pb", n);
  double* Ap = lat_buf<double>(ctx, "pcg_Ap", n);
  double* dinv = lat_buf<double>(ctx, "pcg_dinv", PC == LAT_PC_BLOCK6 ? 21 * n_nodes : n);
  double* partials = lat_buf<double>(ctx, "pcg_partials", (size_t)4 * max_grid + 8);
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
  prm.dist = 0;
  prm.l2_keep = 0;
  prm.seq_base = 0;
  prm.push_base = 0;
  int check = o->check_every > 0 ? o->check_every : 32;
  if (check > o->maxiter) check = o->maxiter > 0 ? o->maxiter : 1;
  if (!mf && rowptr) {
    // L2-resident part of the matrix (see k_cg_spmv<.., .., true>): ~44 MB, for matrices of up to one L2 size only --
    // measured (tools/ab_spmv_l2keep.py): 98 MB 35.7 -> 32.5 us per iteration, 168 MB 55.4 -> 56.2, 397 MB 130.5 -> 131.8
    // (the update kernel streams the vectors between two products and evicts what a larger matrix leaves room for)
    int32_t last = 0;
    LAT_CUDA(ctx, cudaMemcpyAsync(&last, rowptr + n_nodes, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    LAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const double mat_mb = (double)last * 288.0 / 1e6;
    int keep = (mat_mb > 0.0 && mat_mb <= 128.0) ? (int)(16.0 * 44.0 / mat_mb + 0.5) : 0;
    keep = keep < 0 ? 0 : (keep > 8 ? 8 : keep);
    const char* env_keep = getenv("LAT_SPMV_L2KEEP");
    prm.l2_keep = env_keep ? atoi(env_keep) : keep;
  }

  // bit 3 of `reserved` forces the classic two-reduction recurrences in the textbook mode
  const bool cgv = mf || (!o->reference_semantics && !(o->reserved & 8) && !plan.tma);
  // two-level preconditioner (coarse.cuh): registered on the context by lat_coarse_setup / lat_coarse_set_inverse
  const CoarseSpace& cs = ctx->coarse;
  const bool coarse = cs.active;
  if (coarse && cs.n_nodes != n_nodes)
    return lat_fail(ctx, LAT_ERR_STATE, "the registered coarse space belongs to another system: call lat_coarse_setup / lat_coarse_set_inverse(NULL)", __FILE__, __LINE__);
  if (coarse && !cgv)
    return lat_fail(ctx, LAT_ERR_UNSUPPORTED, "the coarse correction runs in the Chronopoulos-Gear iteration only (no reference_semantics / classic / TMA variant)", __FILE__, __LINE__);
  const CoarseLaunch cl = coarse ? coarse_launch(ctx) : CoarseLaunch();
  const int coarse_launches = coarse ? cl.launches() : 0;
  // u (= z here) += Z Einv Z^T r, between the kernel that produced u = D^-1 r and the product that consumes it
  auto launch_coarse = [&](cudaStream_t st) {
    if (coarse) cl.run(st, r, z, sc, prm.maxiter);
  };
  const int per_iter = (cgv ? 3 : 2) + coarse_launches;
  auto launch_spmv = [&](cudaStream_t st) {
    if (mf)
      k_cg_spmv_mf<false><<<mf_grid, MF_BLOCK, 0, st>>>(*mf, n_nodes, z, r, Ap, sc, partials, prm, RowSet());
    else if (cgv && prm.l2_keep > 0)
      k_cg_spmv<false, false, true><<<grid, SPMV_BLOCK, 0, st>>>(rowptr, colidx, vals, n_nodes, z, r, Ap, sc, partials, prm, RowSet());
    else if (cgv)
      k_cg_spmv<false><<<grid, SPMV_BLOCK, 0, st>>>(rowptr, colidx, vals, n_nodes, z, r, Ap, sc, partials, prm, RowSet());
    else if (plan.tma)
      k_spmv_tma<1><<<plan.grid, SPMV_BLOCK, plan.smem, st>>>(rowptr, colidx, vals, n_nodes, plan.cta_row0,
                                                             plan.cap_blocks, z, pa, pb, Ap, sc, partials, prm);
    else
      k_pcg_spmv<0><<<grid, SPMV_BLOCK, 0, st>>>(rowptr, colidx, vals, n_nodes, z, pa, pb, Ap, sc, partials, prm);
  };
  auto launch_update = [&](cudaStream_t st) {
    if (cgv)
      k_cg_update<PC><<<grid, SPMV_BLOCK, 0, st>>>(n_nodes, dinv, x, r, z, Ap, pa, pb, sc, prm);
    else
      k_pcg_update<PC><<<grid, SPMV_BLOCK, 0, st>>>(n_nodes, dinv, x, r, z, pa, pb, Ap, sc, partials, prm);
  };
  // one iteration = (spmv, update) in the classic form, (update, spmv) in the Chronopoulos-Gear form
  auto launch_reduce = [&](cudaStream_t st) {
    if (cgv) k_cg_reduce<<<1, CG_REDUCE_BLOCK, 0, st>>>(partials, (int)(mf ? mf_grid : grid), sc, prm);
  };
  auto launch_iteration = [&](cudaStream_t st) {
    if (cgv) { launch_update(st); launch_coarse(st); launch_spmv(st); launch_reduce(st); }
    else { launch_spmv(st); launch_update(st); }
  };
  LAT_CUDA(ctx, cudaMemsetAsync(sc, 0, sizeof(PcgScalars), ctx->stream));
  if (PC != LAT_PC_NONE) {
    if (mf) LAT_LAUNCH(ctx, k_mf_precond, (unsigned)ceil_div(n_nodes, 128), 128, 0, *mf, n_nodes, PC, dinv);
    else LAT_LAUNCH(ctx, k_precond_setup, (unsigned)ceil_div(n_nodes, 128), 128, 0, rowptr, colidx, vals, n_nodes, PC, dinv);
  }
  LAT_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
  if (cgv) {
    const int32_t one = 1;
    LAT_CUDA(ctx, cudaMemcpyAsync(&sc->first, &one, sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    LAT_LAUNCH(ctx, k_cg_init<PC>, grid, SPMV_BLOCK, 0, n_nodes, b, dinv, x, r, z, pa, pb);
    launch_coarse(ctx->stream);
    launch_spmv(ctx->stream);  // set-up pass: w0 = A u0, gamma0, delta0, |b|^2
    launch_reduce(ctx->stream);
    ctx->launches += 2 + coarse_launches;
  } else {
    LAT_LAUNCH(ctx, k_pcg_init<PC>, grid, SPMV_BLOCK, 0, n_nodes, b, dinv, x, r, z, pa, pb, sc, partials, 0);
  }

  // one CUDA graph = `check` iterations (2 kernels each); relaunched until the device reports done
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t gexec = nullptr;
  cudaStream_t cap = nullptr;
  LAT_CUDA(ctx, cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
  cudaError_t ce = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
  if (ce == cudaSuccess) {
    for (int it = 0; it < check; ++it) launch_iteration(cap);
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
    std::vector<cudaEvent_t> evs(3 * nprof), evs_end(nprof);
    for (auto& e : evs) cudaEventCreate(&e);
    for (auto& e : evs_end) cudaEventCreate(&e);
    for (int it = 0; it < nprof; ++it) {
      // events bracket each kernel; slot 0->1 is always the SpMV kernel, 1->2 / 2->0' the update kernel
      if (cgv) {
        cudaEventRecord(evs[3 * it + 1], ctx->stream);
        launch_update(ctx->stream);
        launch_coarse(ctx->stream);                    // timed with the update kernel
        ctx->launches += coarse_launches;
        cudaEventRecord(evs[3 * it + 2], ctx->stream);
        cudaEventRecord(evs[3 * it], ctx->stream);
        launch_spmv(ctx->stream);
        cudaEventRecord(evs_end[it], ctx->stream);   // closes the SpMV kernel; the reduce kernel is timed apart
        launch_reduce(ctx->stream);
        ++ctx->launches;
      } else {
        cudaEventRecord(evs[3 * it], ctx->stream);
        launch_spmv(ctx->stream);
        cudaEventRecord(evs[3 * it + 1], ctx->stream);
        launch_update(ctx->stream);
        cudaEventRecord(evs[3 * it + 2], ctx->stream);
      }
      ctx->launches += 2;
    }
    ce = cudaStreamSynchronize(ctx->stream);
    if (ce == cudaSuccess) {
      for (int it = 0; it < nprof; ++it) {
        float a = 0.f, b2 = 0.f;
        if (cgv) {
          cudaEventElapsedTime(&a, evs[3 * it], evs_end[it]);            // SpMV kernel
          cudaEventElapsedTime(&b2, evs[3 * it + 1], evs[3 * it + 2]);   // update kernel
        } else {
          cudaEventElapsedTime(&a, evs[3 * it], evs[3 * it + 1]);
          cudaEventElapsedTime(&b2, evs[3 * it + 1], evs[3 * it + 2]);
        }
        spmv_ms += a;
        update_ms += b2;
      }
      spmv_ms /= nprof;
      update_ms /= nprof;
    }
    for (auto& e : evs) cudaEventDestroy(e);
    for (auto& e : evs_end) cudaEventDestroy(e);
    if (ce != cudaSuccess) rc = lat_cuda_fail(ctx, ce, "PCG profiled iterations", __FILE__, __LINE__);
  }
  const int remaining = o->maxiter - nprof;
  const int nbatch = (int)ceil_div(remaining > 0 ? remaining : 0, check);
  PcgScalars* hs = ctx->h_scal;
  // software pipeline of depth 2: batch i+1 is enqueued before the status of batch i is read;
  // kernels of a batch enqueued after convergence exit immediately (sc->done), so x is frozen
  // at the converged iterate exactly like the `break` of the reference loop.
  auto run_batches = [&](int64_t budget) -> int {
    const int nb = (int)ceil_div(budget > 0 ? budget : 0, check);
    int launched = 0, checked = 0;
    bool finished = (nb <= 0);
    while (!finished) {
      while (launched < nb && launched - checked < 2) {
        ce = cudaGraphLaunch(gexec, ctx->stream);
        if (ce != cudaSuccess) return lat_cuda_fail(ctx, ce, "cudaGraphLaunch", __FILE__, __LINE__);
        ctx->launches += per_iter * check;
        cudaMemcpyAsync(&hs[launched & 1], sc, sizeof(PcgScalars), cudaMemcpyDeviceToHost, ctx->stream);
        cudaEventRecord(ctx->ev[2 + (launched & 1)], ctx->stream);
        ++launched;
      }
      const int sl = checked & 1;
      ce = cudaEventSynchronize(ctx->ev[2 + sl]);
      if (ce != cudaSuccess) return lat_cuda_fail(ctx, ce, "PCG iteration", __FILE__, __LINE__);
      ++checked;
      if (hs[sl].done || hs[sl].iters >= o->maxiter || checked == nb) finished = true;
    }
    cudaMemcpyAsync(&hs[0], sc, sizeof(PcgScalars), cudaMemcpyDeviceToHost, ctx->stream);
    ce = cudaStreamSynchronize(ctx->stream);
    if (ce != cudaSuccess) return lat_cuda_fail(ctx, ce, "PCG sync", __FILE__, __LINE__);
    return LAT_OK;
  };
  if (rc == LAT_OK) rc = run_batches(remaining);
  // true-residual safeguard of the Chronopoulos-Gear form (see k_cg_restart)
  double true_rr = -1.0;
  while (rc == LAT_OK && cgv && hs[0].done && !hs[0].breakdown && hs[0].bb > 0.0) {
    if (mf) { k_mf_apply<true><<<mf_grid, MF_BLOCK, 0, ctx->stream>>>(*mf, n_nodes, x, Ap); ++ctx->launches; }
    else rc = lat_spmv_internal(ctx, rowptr, colidx, vals, n_nodes, x, Ap);
    if (rc) break;
    k_cg_restart<PC><<<grid, SPMV_BLOCK, 0, ctx->stream>>>(n_nodes, b, Ap, dinv, r, z, pa, pb, partials);
    k_cg_true_residual<<<1, CG_REDUCE_BLOCK, 0, ctx->stream>>>(partials, (int)grid, sc, prm, 2);
    ctx->launches += 2;
    cudaMemcpyAsync(&hs[0], sc, sizeof(PcgScalars), cudaMemcpyDeviceToHost, ctx->stream);
    ce = cudaStreamSynchronize(ctx->stream);
    if (ce != cudaSuccess) { rc = lat_cuda_fail(ctx, ce, "PCG residual check", __FILE__, __LINE__); break; }
    true_rr = hs[0].true_rr;
    if (hs[0].done) break;       // accepted
    // restart from x: set-up pass, then iterate on the remaining budget
    launch_coarse(ctx->stream);
    launch_spmv(ctx->stream);
    launch_reduce(ctx->stream);
    ctx->launches += 2 + coarse_launches;
    rc = run_batches((int64_t)o->maxiter - hs[0].iters);
  }
  if (rc == LAT_OK) {
    cudaEventRecord(ctx->ev[1], ctx->stream);
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
  res->info = hs[0].done && !hs[0].breakdown ? 0 : (hs[0].breakdown == 3 ? 5 : (hs[0].breakdown ? 3 : (hs[0].info_flag2 ? 2 : 1)));
  res->solve_ms = ms;
  res->launches = ctx->launches - launches0;
  res->spmv_ms = spmv_ms;
  res->update_ms = update_ms;
  res->profiled = nprof;
  res->reserved = hs[0].restarts | 0x100 | (coarse ? 0x400 : 0);   // bit 8: the iteration batches ran as CUDA graphs; bit 10: two-level
  res->true_relres = (true_rr >= 0.0 && hs[0].bb > 0.0) ? sqrt(true_rr / hs[0].bb) : -1.0;
  return LAT_OK;
}

// ---------------------------------------------------------------------------
// persistent on-chip PCG (pcg_persist.cuh): host side
// ---------------------------------------------------------------------------
#include "pcg_persist.cuh"

// Multi-GPU set-up of the persistent kernel, filled by pcg_run_dist_impl (peer-memory arena, fused-halo tables).
struct PersistDistSetup {
  double* u = nullptr;                        // the ghosted u inside this rank's arena
  int nranks = 1, my_rank = 0, n_nb = 0;
  long long ghost_first[2] = {0, 0};
  int ghost_entries[2] = {0, 0};
  unsigned long long push_base = 0, seq_base = 0;
  const int2* push_dst = nullptr;
  unsigned long long* peer_ll[2] = {nullptr, nullptr};
  const unsigned long long* my_ll = nullptr;
  unsigned char* const* peers = nullptr;
  unsigned long long pushes = 0;              // out: productions of u (= halo pushes) of the solve
};

// Returns LAT_OK with *used = true when the solve ran in the persistent kernel; *used = false (and LAT_OK) when the
// system does not fit the shared memory of the device or cooperative launch is unavailable -- the caller then
// runs the three-kernel iteration.  With `dist` the decision is collective (every rank's slab must fit: one tiny
// all-reduce), n_nodes counts the owned rows and u lives in the arena.
template <int PC>
static int pcg_run_persist(lat_ctx* ctx, const int32_t* rowptr, const int32_t* colidx, const double* vals,
                           int64_t n_nodes, const double* b, double* x, const lat_pcg_opts* o, lat_pcg_result* res,
                           bool* used, PersistDistSetup* dist = nullptr) {
  *used = false;
  int coop = 0, smem_optin = 0;
  LAT_CUDA(ctx, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device));
  LAT_CUDA(ctx, cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
  const int G = ctx->sm_count < PERSIST_INBOX_STRIDE ? ctx->sm_count : PERSIST_INBOX_STRIDE;
  bool fits = coop && n_nodes < (int64_t)INT32_MAX / 8;
  // cheap size screen before any work: 5 vectors of the average CTA must fit at all
  if ((double)n_nodes / G * (PERSIST_NVEC * 48 + persist_pc_width(PC) * 8) > (double)smem_optin) fits = false;
  if ((n_nodes + PERSIST_CHUNK - 1) / PERSIST_CHUNK > (int64_t)64 * G) fits = false;   // s_chunk_base holds 64 chunks per CTA
  if (!fits && !dist) return LAT_OK;
  int rows_cap = 1, blk_cap = 1;
  bool pc_smem = false;
  size_t smem = 0;
  if (fits) {
    int32_t* maxima = lat_buf<int32_t>(ctx, "persist_max", 4);
    if (!maxima) return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
    LAT_CUDA(ctx, cudaMemsetAsync(maxima, 0, 4 * sizeof(int32_t), ctx->stream));
    LAT_LAUNCH(ctx, k_persist_caps, (unsigned)ceil_div(G, 128), 128, 0, rowptr, n_nodes, G, maxima);
    int32_t hmax[2] = {0, 0};
    LAT_CUDA(ctx, cudaMemcpyAsync(hmax, maxima, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    LAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    rows_cap = hmax[0] > 0 ? hmax[0] : 1;
    blk_cap = hmax[1] > 0 ? hmax[1] : 1;
    // Shared memory and L1 share 256 KB per SM, and the product phase needs its L1: the preconditioner rows are
    // cached on chip only while the total stays below ~160 KB (LAT_PERSIST_PCSMEM_KB overrides the limit).
    const size_t smem_base = (size_t)rows_cap * (PERSIST_NVEC * 48) + (size_t)3 * G * 8 + (size_t)(2 * rows_cap + 1) * 4 +
                             (size_t)blk_cap * 4 + (size_t)((rows_cap + 15) / 16) * 16 + 32;
    const size_t smem_pc = (size_t)rows_cap * persist_pc_width(PC) * 8;
    const char* env_kb = getenv("LAT_PERSIST_PCSMEM_KB");
    const size_t pc_limit = (size_t)(env_kb ? atoi(env_kb) : 160) * 1024;
    pc_smem = smem_base + smem_pc + 2560 <= (size_t)smem_optin && smem_base + smem_pc <= pc_limit;
    smem = smem_base + (pc_smem ? smem_pc : 0);
    if (smem + 2560 > (size_t)smem_optin) fits = false;     // does not fit on chip: three-kernel path
    // ... and above ~176 KB of shared memory the product phase has too little L1 left for its gathers of u: BCC 24^3
    // (0.84 M DOF, 187 KB) runs at 61.3 us per iteration here against 53.2 us in the three-kernel iteration
    // (tools/ab_persist.py); LAT_PERSIST_FORCE=1 keeps the on-chip kernel up to the hardware limit
    if (smem > (size_t)176 * 1024 && !getenv("LAT_PERSIST_FORCE")) fits = false;
  }
  const void* kfn = dist ? (const void*)k_pcg_persist<PC, true> : (const void*)k_pcg_persist<PC, false>;
  if (fits) {
    LAT_CUDA(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    if (dist) LAT_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcg_persist<PC, true>, PERSIST_BLOCK, smem));
    else LAT_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcg_persist<PC, false>, PERSIST_BLOCK, smem));
    if (per_sm < 1) fits = false;
  }
  if (dist && dist->nranks > 1) {
    // collective decision: one rank in the persistent kernel and another in the three-kernel iteration would wait
    // for each other for ever
    double* vote = lat_buf<double>(ctx, "persist_vote", 1);
    if (!vote) return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
    const double mine = fits ? 0.0 : 1.0;
    LAT_CUDA(ctx, cudaMemcpyAsync(vote, &mine, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (int rc = lat_allreduce_sum(ctx, vote, 1)) return rc;
    double against = 1.0;
    LAT_CUDA(ctx, cudaMemcpyAsync(&against, vote, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    LAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (against > 0.0) fits = false;
  }
  if (!fits) return LAT_OK;

  const int64_t n = 6 * n_nodes;
  double* u = dist ? dist->u : lat_buf<double>(ctx, "pcg_z", n);
  double* dinv = lat_buf<double>(ctx, "pcg_dinv", PC == LAT_PC_BLOCK6 ? 21 * n_nodes : n);
  PcgScalars* sc = lat_buf<PcgScalars>(ctx, "pcg_scalars", 1);
  unsigned long long* mail = lat_buf<unsigned long long>(ctx, "persist_mail", (size_t)G * 8);
  unsigned int* flags = lat_buf<unsigned int>(ctx, "persist_flags", (size_t)G * PERSIST_INBOX_STRIDE);
  if (!u || !dinv || !sc || !mail || !flags) return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
  const int64_t launches0 = ctx->launches - 1;
  LAT_CUDA(ctx, cudaMemsetAsync(sc, 0, sizeof(PcgScalars), ctx->stream));
  LAT_CUDA(ctx, cudaMemsetAsync(mail, 0, (size_t)G * 8 * sizeof(unsigned long long), ctx->stream));
  LAT_CUDA(ctx, cudaMemsetAsync(flags, 0, (size_t)G * PERSIST_INBOX_STRIDE * sizeof(unsigned int), ctx->stream));
  if (PC != LAT_PC_NONE)
    LAT_LAUNCH(ctx, k_precond_setup, (unsigned)ceil_div(n_nodes, 128), 128, 0, rowptr, colidx, vals, n_nodes, PC, dinv);
  float* dinv_full = nullptr;
  if (PC == LAT_PC_BLOCK6 && !pc_smem) {
    dinv_full = lat_buf<float>(ctx, "pcg_dinv_full", (size_t)36 * n_nodes);
    if (!dinv_full) return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
    LAT_LAUNCH(ctx, k_persist_expand_dinv, (unsigned)ceil_div(36 * n_nodes, 256), 256, 0, dinv, n_nodes, dinv_full);
  }
  PersistArgs a;
  a.dinv_full = dinv_full;
  a.rowptr = rowptr; a.colidx = colidx; a.vals = vals; a.n_nodes = n_nodes; a.b = b; a.x = x; a.u = u; a.dinv = dinv;
  a.sc = sc;
  a.prm.tol = o->tol; a.prm.mintol = 0.0; a.prm.alpha_max = 0.0; a.prm.restart_every = 0; a.prm.maxiter = o->maxiter;
  a.prm.reference = 0; a.prm.dist = 0; a.prm.l2_keep = 0; a.prm.seq_base = 0; a.prm.push_base = 0;
  a.mail = mail; a.flags = flags; a.rows_cap = rows_cap; a.blk_cap = blk_cap; a.pc_smem = pc_smem ? 1 : 0;
  a.nranks = 1; a.my_rank = 0; a.n_nb = 0;
  a.ghost_first[0] = a.ghost_first[1] = 0; a.ghost_entries[0] = a.ghost_entries[1] = 0;
  a.push_base = 0; a.seq_base = 0;
  a.push_dst = nullptr; a.peer_ll[0] = a.peer_ll[1] = nullptr; a.my_ll = nullptr; a.peers = nullptr;
  if (dist) {
    a.nranks = dist->nranks; a.my_rank = dist->my_rank; a.n_nb = dist->n_nb;
    for (int k = 0; k < 2; ++k) {
      a.ghost_first[k] = dist->ghost_first[k]; a.ghost_entries[k] = dist->ghost_entries[k];
      a.peer_ll[k] = dist->peer_ll[k];
    }
    a.push_base = dist->push_base; a.seq_base = dist->seq_base; a.prm.seq_base = dist->seq_base; a.prm.push_base = dist->push_base;
    a.push_dst = dist->push_dst; a.my_ll = dist->my_ll; a.peers = dist->peers;
  }
  a.trace = nullptr;
  a.trace_iters = 0;
  // L2 residency of part of the matrix: the blocks of l2_keep / 16 of the rows are loaded with an evict_last policy and
  // stay in the 126 MB L2 from one iteration to the next, the rest is streamed evict_first.  Target ~44 MB resident
  // (config 1, 98 MB matrix: 7 / 16 -> product phase 15.4 -> 11.5 us, iteration 30.2 -> 24.6 us; 12 / 16 and more thrash,
  // profiles/r02_persist_l2keep_ab.txt).  LAT_PERSIST_L2KEEP overrides (0 = plain loads).
  {
    int64_t nnzb_h = 0;
    {
      int32_t last = 0;
      LAT_CUDA(ctx, cudaMemcpyAsync(&last, rowptr + n_nodes, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
      LAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      nnzb_h = last;
    }
    const double mat_mb = (double)nnzb_h * 288.0 / 1e6;
    int keep = mat_mb > 0.0 ? (int)(16.0 * 44.0 / mat_mb + 0.5) : 8;
    keep = keep < 0 ? 0 : (keep > 8 ? 8 : keep);
    const char* env_keep = getenv("LAT_PERSIST_L2KEEP");
    a.l2_keep = env_keep ? atoi(env_keep) : keep;
  }
  const char* env_tr = getenv("LAT_PERSIST_TRACE");     // LAT_PERSIST_TRACE=n: phase breakdown of the first n iterations on stderr
  if (env_tr && atoi(env_tr) > 0) {
    a.trace_iters = atoi(env_tr) > 256 ? 256 : atoi(env_tr);
    a.trace = lat_buf<long long>(ctx, "persist_trace", (size_t)G * a.trace_iters * 9 + 2 * G);
    if (!a.trace) return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
    LAT_CUDA(ctx, cudaMemsetAsync(a.trace, 0, ((size_t)G * a.trace_iters * 9 + 2 * G) * sizeof(long long), ctx->stream));
  }
  void* kargs[] = {&a};
  LAT_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
  LAT_CUDA(ctx, cudaLaunchCooperativeKernel(kfn, dim3(G), dim3(PERSIST_BLOCK), kargs, smem, ctx->stream));
  ctx->launches++;
  LAT_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
  PcgScalars* hs = ctx->h_scal;
  LAT_CUDA(ctx, cudaMemcpyAsync(&hs[0], sc, sizeof(PcgScalars), cudaMemcpyDeviceToHost, ctx->stream));
  LAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
  if (a.trace) {
    std::vector<long long> tr((size_t)G * a.trace_iters * 9 + 2 * G);
    cudaMemcpy(tr.data(), a.trace, tr.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    int clk_khz = 1;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, ctx->device);
    const int n_it = hs[0].iters < a.trace_iters ? hs[0].iters : a.trace_iters;
    double fine[5] = {0, 0, 0, 0, 0}; double sum[4] = {0, 0, 0, 0}, prod_min = 0, prod_max = 0, upd_min = 0, upd_max = 0;
    int cnt = 0;
    for (int it = 2; it < n_it; ++it) {
      double pmin = 1e30, pmax = 0, umin = 1e30, umax = 0;
      for (int c = 0; c < G; ++c) {
        const long long* t = &tr[((size_t)c * a.trace_iters + it) * 9];
        if (!t[4]) continue;
        const double d[4] = {(double)(t[1] - t[0]), (double)(t[2] - t[1]), (double)(t[3] - t[2]), (double)(t[4] - t[3])};
        for (int k = 0; k < 4; ++k) sum[k] += d[k];
        fine[0] += (double)(t[5] - t[3]); fine[1] += (double)(t[6] - t[5]); fine[2] += (double)(t[7] - t[6]); fine[3] += (double)(t[8] - t[7]); fine[4] += (double)(t[4] - t[8]);
        pmin = d[0] < pmin ? d[0] : pmin; pmax = d[0] > pmax ? d[0] : pmax;
        umin = d[2] < umin ? d[2] : umin; umax = d[2] > umax ? d[2] : umax;
        ++cnt;
      }
      prod_min += pmin; prod_max += pmax; upd_min += umin; upd_max += umax;
    }
    const double us = 1e3 / (double)clk_khz;   // cycles -> us at the nominal SM clock
    if (const char* fn = getenv("LAT_PERSIST_TRACE_FILE")) {   // per-CTA means: cta rows blocks product_us update_us
      if (FILE* fp = fopen(fn, "w")) {
        for (int c = 0; c < G; ++c) {
          double p = 0, u2 = 0; int k = 0;
          for (int it = 2; it < n_it; ++it) {
            const long long* t = &tr[((size_t)c * a.trace_iters + it) * 9];
            if (!t[4]) continue;
            p += (double)(t[1] - t[0]); u2 += (double)(t[3] - t[2]); ++k;
          }
          fprintf(fp, "%d %lld %lld %.3f %.3f\n", c, tr[(size_t)G * a.trace_iters * 9 + 2 * c], tr[(size_t)G * a.trace_iters * 9 + 2 * c + 1],
                  k ? p / k * us : 0.0, k ? u2 / k * us : 0.0);
        }
        fclose(fp);
      }
    }
    const int nit = n_it > 2 ? n_it - 2 : 1;
    if (cnt > 0)
      fprintf(stderr, "[lat persist trace] mean over %d CTAs x %d iterations (us at %.0f MHz): product %.2f (fastest CTA %.2f, slowest %.2f) | "
                      "reduce/barrier B %.2f | update %.2f (fastest %.2f, slowest %.2f) | barrier A %.2f\n",
              G, nit, clk_khz / 1e3, sum[0] / cnt * us, prod_min / nit * us, prod_max / nit * us, sum[1] / cnt * us,
              sum[2] / cnt * us, upd_min / nit * us, upd_max / nit * us, sum[3] / cnt * us);
    if (cnt > 0)
      fprintf(stderr, "[lat persist trace] barrier A in detail: first bar.sync %.2f | release fence + flag %.2f | poll + bar.sync %.2f | acquire fence %.2f | last bar.sync %.2f\n",
              fine[0] / cnt * us, fine[1] / cnt * us, fine[2] / cnt * us, fine[3] / cnt * us, fine[4] / cnt * us);
  }
  res->iters = hs[0].iters;
  res->norm_b = sqrt(hs[0].bb);
  res->relres = hs[0].bb > 0.0 ? sqrt(hs[0].rr / hs[0].bb) : 0.0;
  res->info = hs[0].done && !hs[0].breakdown ? 0 : (hs[0].breakdown == 2 ? 4 : (hs[0].breakdown == 3 ? 5 : (hs[0].breakdown ? 3 : 1)));
  res->solve_ms = ms;
  res->launches = ctx->launches - launches0;
  res->spmv_ms = 0.0;
  res->update_ms = 0.0;
  res->profiled = 0;
  res->reserved = hs[0].restarts | 0x200;    // bit 9: solved by the persistent on-chip kernel
  if (dist) dist->pushes = hs[0].counter[0];
  res->true_relres = (hs[0].true_rr >= 0.0 && hs[0].bb > 0.0) ? sqrt(hs[0].true_rr / hs[0].bb) : -1.0;
  *used = true;
  return LAT_OK;
}

extern "C" int lat_pcg_bsr(lat_ctx* ctx, const int32_t* rowptr, const int32_t* colidx, const double* vals,
                           int64_t n_nodes, const double* b, double* x, const lat_pcg_opts* opts,
                           lat_pcg_result* result) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, rowptr && colidx && vals && b && x && opts && result && n_nodes > 0);
  LAT_CHECK_ARG(ctx, opts->maxiter >= 0 && opts->tol >= 0.0);
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  // Textbook mode, nothing experimental requested: try the persistent on-chip kernel first (bit 7 of `reserved`
  // opts out; profile_iters > 0 asks for per-kernel timings, which only the three-kernel iteration has).
  const bool want_persist = !opts->reference_semantics && !(opts->reserved & (2 | 8 | 128)) && opts->profile_iters <= 0 &&
                            !ctx->coarse.active;   // the coarse correction lives in the three-kernel iteration
  if (want_persist) {
    bool used = false;
    int rc = LAT_OK;
    switch (opts->precond) {
      case LAT_PC_NONE: rc = pcg_run_persist<LAT_PC_NONE>(ctx, rowptr, colidx, vals, n_nodes, b, x, opts, result, &used); break;
      case LAT_PC_JACOBI: rc = pcg_run_persist<LAT_PC_JACOBI>(ctx, rowptr, colidx, vals, n_nodes, b, x, opts, result, &used); break;
      case LAT_PC_BLOCK6: rc = pcg_run_persist<LAT_PC_BLOCK6>(ctx, rowptr, colidx, vals, n_nodes, b, x, opts, result, &used); break;
      default: return lat_fail(ctx, LAT_ERR_ARG, "unknown preconditioner", __FILE__, __LINE__);
    }
    if (rc != LAT_OK || used) return rc;
  }
  switch (opts->precond) {
    case LAT_PC_NONE: return pcg_run<LAT_PC_NONE>(ctx, rowptr, colidx, vals, n_nodes, b, x, opts, result);
    case LAT_PC_JACOBI: return pcg_run<LAT_PC_JACOBI>(ctx, rowptr, colidx, vals, n_nodes, b, x, opts, result);
    case LAT_PC_BLOCK6: return pcg_run<LAT_PC_BLOCK6>(ctx, rowptr, colidx, vals, n_nodes, b, x, opts, result);
    default: return lat_fail(ctx, LAT_ERR_ARG, "unknown preconditioner", __FILE__, __LINE__);
  }
}

// ---------------------------------------------------------------------------
// matrix-free operator (matfree.cuh): resident set-up, products, PCG
// ---------------------------------------------------------------------------
static int mf_get(lat_ctx* ctx, MfOp* op) {
  if (ctx->mf_nnodes < 0)
    return lat_fail(ctx, LAT_ERR_STATE, "no resident matrix-free operator: call lat_matfree_setup", __FILE__, __LINE__);
  const double G = ctx->mf_young / (2.0 * (1.0 + ctx->mf_nu));
  op->adjptr = (const int32_t*)ctx->bufs["pat_adjptr"].p;
  op->inc = (const MfInc*)ctx->bufs["mf_inc"].p;
  op->node4 = (const double*)ctx->bufs["mf_node4"].p;
  op->E = ctx->mf_young;
  op->Gk = G * ctx->mf_kappa;
  op->G2mE = 2.0 * G - ctx->mf_young;
  op->inv4pi = 1.0 / (4.0 * 3.14159265358979323846);
  return LAT_OK;
}

extern "C" int lat_matfree_setup(lat_ctx* ctx, const double* x, const double* y, const double* z,
                                 const int32_t* en0, const int32_t* en1, const double* rad, int64_t n_elem,
                                 int64_t n_nodes, double young, double nu, double kappa, const uint8_t* fixed) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, x && y && z && en0 && en1 && rad && n_elem > 0 && n_nodes > 0);
  if (ctx->pat_nnzb < 0 || ctx->pat_nelem != n_elem || ctx->pat_nnodes != n_nodes)
    return lat_fail(ctx, LAT_ERR_STATE, "resident pattern does not match this mesh: call lat_bsr_pattern_build", __FILE__, __LINE__);
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  MfInc* inc = lat_buf<MfInc>(ctx, "mf_inc", (size_t)2 * n_elem);
  double* node4 = lat_buf<double>(ctx, "mf_node4", (size_t)4 * n_nodes);
  if (!inc || !node4) return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
  LAT_LAUNCH(ctx, k_mf_setup_inc, (unsigned)ceil_div(2 * n_elem, 256), 256, 0, (const int32_t*)ctx->bufs["pat_adj_other"].p,
             (const int32_t*)ctx->bufs["pat_adj_el"].p, x, y, z, en0, en1, rad, 2 * n_elem, inc);
  LAT_LAUNCH(ctx, k_mf_setup_nodes, (unsigned)ceil_div(n_nodes, 256), 256, 0, x, y, z, fixed, n_nodes, node4);
  ctx->mf_nnodes = n_nodes;
  ctx->mf_nelem = n_elem;
  ctx->mf_young = young;
  ctx->mf_nu = nu;
  ctx->mf_kappa = kappa;
  return LAT_OK;
}

extern "C" int lat_matfree_apply(lat_ctx* ctx, const double* u, double* y, int eliminated) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, u && y && u != y);
  LAT_CHECK_ARG(ctx, ((reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(y)) & 31) == 0);   // 256-bit accesses
  MfOp op;
  if (int rc = mf_get(ctx, &op)) return rc;
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  const unsigned grid = (unsigned)ceil_div(ctx->mf_nnodes, MF_BLOCK);
  if (eliminated) LAT_LAUNCH(ctx, k_mf_apply<true>, grid, MF_BLOCK, 0, op, ctx->mf_nnodes, u, y);
  else LAT_LAUNCH(ctx, k_mf_apply<false>, grid, MF_BLOCK, 0, op, ctx->mf_nnodes, u, y);
  return LAT_OK;
}

extern "C" int lat_matfree_rhs(lat_ctx* ctx, const double* g, const double* f, double* b) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, g && b && g != b);
  LAT_CHECK_ARG(ctx, ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(b)) & 31) == 0);
  MfOp op;
  if (int rc = mf_get(ctx, &op)) return rc;
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  LAT_LAUNCH(ctx, k_mf_rhs, (unsigned)ceil_div(ctx->mf_nnodes, MF_BLOCK), MF_BLOCK, 0, op, ctx->mf_nnodes, g, f, b);
  return LAT_OK;
}

extern "C" int lat_pcg_matfree(lat_ctx* ctx, const double* b, double* x, const lat_pcg_opts* opts,
                               lat_pcg_result* result) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, b && x && opts && result);
  LAT_CHECK_ARG(ctx, opts->maxiter >= 0 && opts->tol >= 0.0);
  LAT_CHECK_ARG(ctx, (reinterpret_cast<uintptr_t>(x) & 31) == 0);
  if (opts->reference_semantics)
    return lat_fail(ctx, LAT_ERR_UNSUPPORTED, "the reference clamp/restart rules are defined on the assembled path (lat_pcg_bsr)", __FILE__, __LINE__);
  MfOp op;
  if (int rc = mf_get(ctx, &op)) return rc;
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t nn = ctx->mf_nnodes;
  switch (opts->precond) {
    case LAT_PC_NONE: return pcg_run<LAT_PC_NONE>(ctx, nullptr, nullptr, nullptr, nn, b, x, opts, result, &op);
    case LAT_PC_JACOBI: return pcg_run<LAT_PC_JACOBI>(ctx, nullptr, nullptr, nullptr, nn, b, x, opts, result, &op);
    case LAT_PC_BLOCK6: return pcg_run<LAT_PC_BLOCK6>(ctx, nullptr, nullptr, nullptr, nn, b, x, opts, result, &op);
    default: return lat_fail(ctx, LAT_ERR_ARG, "unknown preconditioner", __FILE__, __LINE__);
  }
}


// ---------------------------------------------------------------------------
// two-level preconditioner (coarse.cuh): set-up entry points
// ---------------------------------------------------------------------------
extern "C" int lat_coarse_setup(lat_ctx* ctx, const double* x, const double* y, const double* z, int64_t n_nodes,
                                const int32_t* node_agg, const int32_t* agg_ptr, const int32_t* agg_nodes, int32_t n_agg,
                                const uint8_t* fixed, const double* centers) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, x && y && z && node_agg && agg_ptr && agg_nodes && n_nodes > 0 && n_agg > 0 &&
                         n_nodes < ((int64_t)1 << COARSE_NODE_BITS));
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  ctx->coarse = CoarseSpace();
  CoarseNode* by_agg = lat_buf<CoarseNode>(ctx, "coarse_nodes", (size_t)n_nodes);
  CoarseNode* by_node = lat_buf<CoarseNode>(ctx, "coarse_bynode", (size_t)n_nodes);
  int32_t* ptr = lat_buf<int32_t>(ctx, "coarse_ptr", (size_t)n_agg + 1);
  double* rc = lat_buf<double>(ctx, "coarse_rc", (size_t)6 * n_agg);
  double* yc = lat_buf<double>(ctx, "coarse_yc", (size_t)6 * n_agg);
  double* cen = lat_buf<double>(ctx, "coarse_centers", (size_t)3 * n_agg);
  if (!by_agg || !by_node || !ptr || !rc || !yc || !cen) return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
  std::vector<int32_t> h_ptr((size_t)n_agg + 1);
  LAT_CUDA(ctx, cudaMemcpyAsync(ptr, agg_ptr, ((size_t)n_agg + 1) * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));
  LAT_CUDA(ctx, cudaMemcpyAsync(h_ptr.data(), agg_ptr, ((size_t)n_agg + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  LAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (h_ptr[0] != 0 || h_ptr[n_agg] < 1 || h_ptr[n_agg] > n_nodes)
    return lat_fail(ctx, LAT_ERR_ARG, "agg_ptr must run from 0 to the number of listed nodes (1 .. n_nodes)", __FILE__, __LINE__);
  // pieces: every aggregate cut into chunks of at most COARSE_PIECE entries
  std::vector<int32_t> piece_ptr(1, 0), piece_agg, agg_piece((size_t)n_agg + 1, 0);
  for (int32_t a = 0; a < n_agg; ++a) {
    if (h_ptr[a + 1] < h_ptr[a]) return lat_fail(ctx, LAT_ERR_ARG, "agg_ptr must be non-decreasing", __FILE__, __LINE__);
    for (int32_t k = h_ptr[a]; k < h_ptr[a + 1]; k += COARSE_PIECE) {
      piece_ptr.push_back(std::min(k + COARSE_PIECE, h_ptr[a + 1]));
      piece_agg.push_back(a);
    }
    agg_piece[a + 1] = (int32_t)piece_agg.size();
  }
  const int32_t n_pieces = (int32_t)piece_agg.size();
  int32_t* d_pp = lat_buf<int32_t>(ctx, "coarse_piece_ptr", (size_t)n_pieces + 1);
  int32_t* d_pa = lat_buf<int32_t>(ctx, "coarse_piece_agg", (size_t)std::max(n_pieces, 1));
  int32_t* d_ap = lat_buf<int32_t>(ctx, "coarse_agg_piece", (size_t)n_agg + 1);
  double* part = lat_buf<double>(ctx, "coarse_part", (size_t)6 * std::max(n_pieces, 1));
  if (!d_pp || !d_pa || !d_ap || !part) return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
  LAT_CUDA(ctx, cudaMemcpyAsync(d_pp, piece_ptr.data(), ((size_t)n_pieces + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  LAT_CUDA(ctx, cudaMemcpyAsync(d_pa, piece_agg.data(), (size_t)n_pieces * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  LAT_CUDA(ctx, cudaMemcpyAsync(d_ap, agg_piece.data(), ((size_t)n_agg + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  LAT_LAUNCH(ctx, k_coarse_setup, (unsigned)n_agg, COARSE_BLOCK, 0, ptr, agg_nodes, x, y, z, fixed, centers, cen, by_agg);
  LAT_LAUNCH(ctx, k_coarse_bynode, (unsigned)ceil_div(n_nodes, 256), 256, 0, node_agg, x, y, z, fixed, (const double*)cen, n_nodes, by_node);
  LAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // the host vectors above are read by the async copies
  ctx->coarse.n_pieces = n_pieces;
  bool one_each = n_pieces == n_agg;
  for (int32_t a = 0; a < n_agg && one_each; ++a) one_each = agg_piece[a + 1] - agg_piece[a] == 1;
  ctx->coarse.fused = one_each;
  ctx->coarse.n_nodes = n_nodes;
  ctx->coarse.n_agg = n_agg;
  return LAT_OK;
}

extern "C" int lat_coarse_galerkin(lat_ctx* ctx, const int32_t* rowptr, const int32_t* colidx, const double* vals,
                                   int64_t n_nodes, double* E) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, rowptr && colidx && vals && E);
  if (ctx->coarse.n_nodes < n_nodes || n_nodes <= 0)      // n_nodes = block rows given (the owned rows of a sharded matrix)
    return lat_fail(ctx, LAT_ERR_STATE, "no coarse space for this system: call lat_coarse_setup", __FILE__, __LINE__);
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t n_c = 6 * (int64_t)ctx->coarse.n_agg;
  LAT_CUDA(ctx, cudaMemsetAsync(E, 0, (size_t)n_c * n_c * sizeof(double), ctx->stream));
  LAT_LAUNCH(ctx, k_coarse_galerkin, (unsigned)ceil_div(n_nodes, 128), 128, 0, rowptr, colidx, vals, n_nodes,
             (const CoarseNode*)ctx->bufs["coarse_bynode"].p, n_c, E);
  return LAT_OK;
}

extern "C" int lat_coarse_set_inverse(lat_ctx* ctx, const double* einv) {
  if (!ctx) return LAT_ERR_ARG;
  if (einv && ctx->coarse.n_nodes < 0)
    return lat_fail(ctx, LAT_ERR_STATE, "no coarse space: call lat_coarse_setup", __FILE__, __LINE__);
  ctx->coarse.einv = einv;
  ctx->coarse.active = einv != nullptr;
  return LAT_OK;
}

extern "C" int lat_coarse_apply(lat_ctx* ctx, const double* r, double* u) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, r && u && r != u);
  LAT_CHECK_ARG(ctx, ((reinterpret_cast<uintptr_t>(r) | reinterpret_cast<uintptr_t>(u)) & 15) == 0);
  const CoarseSpace& cs = ctx->coarse;
  if (!cs.active) return lat_fail(ctx, LAT_ERR_STATE, "no coarse inverse registered: call lat_coarse_set_inverse", __FILE__, __LINE__);
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  const CoarseLaunch cl = coarse_launch(ctx);
  cl.run(ctx->stream, r, u, nullptr, 0);
  ctx->launches += cl.launches();
  LAT_CUDA(ctx, cudaPeekAtLastError());
  return LAT_OK;
}

// ===========================================================================
// multi-GPU: NCCL (dlopen'ed -- the library torch already loaded), halo exchange, distributed PCG
// ===========================================================================
#include <dlfcn.h>

namespace {
struct NcclUniqueId { char internal[128]; };
typedef int (*fn_GetUniqueId)(NcclUniqueId*);
typedef int (*fn_CommInitRank)(void**, int, NcclUniqueId, int);
typedef int (*fn_CommDestroy)(void*);
typedef int (*fn_AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_SendRecv)(void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_Send)(const void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_Group)(void);
typedef const char* (*fn_ErrStr)(int);
struct NcclApi {
  void* h = nullptr;
  fn_GetUniqueId GetUniqueId = nullptr;
  fn_CommInitRank CommInitRank = nullptr;
  fn_CommDestroy CommDestroy = nullptr;
  fn_AllReduce AllReduce = nullptr;
  fn_Send Send = nullptr;
  fn_SendRecv Recv = nullptr;
  fn_Group GroupStart = nullptr, GroupEnd = nullptr;
  fn_ErrStr ErrStr = nullptr;
  bool ok = false;
} g_nccl;
const int NCCL_FLOAT64 = 8, NCCL_SUM = 0;

bool nccl_load() {
  if (g_nccl.ok) return true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    g_nccl.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.h) break;
  }
  if (!g_nccl.h) return false;
  g_nccl.GetUniqueId = (fn_GetUniqueId)dlsym(g_nccl.h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (fn_CommInitRank)dlsym(g_nccl.h, "ncclCommInitRank");
  g_nccl.CommDestroy = (fn_CommDestroy)dlsym(g_nccl.h, "ncclCommDestroy");
  g_nccl.AllReduce = (fn_AllReduce)dlsym(g_nccl.h, "ncclAllReduce");
  g_nccl.Send = (fn_Send)dlsym(g_nccl.h, "ncclSend");
  g_nccl.Recv = (fn_SendRecv)dlsym(g_nccl.h, "ncclRecv");
  g_nccl.GroupStart = (fn_Group)dlsym(g_nccl.h, "ncclGroupStart");
  g_nccl.GroupEnd = (fn_Group)dlsym(g_nccl.h, "ncclGroupEnd");
  g_nccl.ErrStr = (fn_ErrStr)dlsym(g_nccl.h, "ncclGetErrorString");
  g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.AllReduce && g_nccl.Send &&
              g_nccl.Recv && g_nccl.GroupStart && g_nccl.GroupEnd;
  return g_nccl.ok;
}

int nccl_fail(lat_ctx* ctx, int rc, const char* what, int line) {
  char buf[512];
  snprintf(buf, sizeof buf, "lattice_b200: NCCL error %d (%s) in %s (line %d)", rc,
           g_nccl.ErrStr ? g_nccl.ErrStr(rc) : "?", what, line);
  ctx->err = buf;
  return 1000 + rc;
}
#define LAT_NCCL(ctx, call)                                   \
  do {                                                        \
    int _r = (call);                                          \
    if (_r != 0) return nccl_fail((ctx), _r, #call, __LINE__); \
  } while (0)
}  // namespace

extern "C" int lat_nccl_unique_id(void* id128) {
  if (!id128) return LAT_ERR_ARG;
  if (!nccl_load()) return LAT_ERR_UNSUPPORTED;
  NcclUniqueId id;
  const int rc = g_nccl.GetUniqueId(&id);
  if (rc != 0) return 1000 + rc;
  memcpy(id128, &id, sizeof id);
  return LAT_OK;
}

extern "C" int lat_comm_create(lat_ctx* ctx, const void* id128, int nranks, int rank) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, id128 && nranks >= 1 && rank >= 0 && rank < nranks);
  if (!nccl_load()) return lat_fail(ctx, LAT_ERR_UNSUPPORTED, "libnccl.so.2 not found", __FILE__, __LINE__);
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  NcclUniqueId id;
  memcpy(&id, id128, sizeof id);
  void* comm = nullptr;
  LAT_NCCL(ctx, g_nccl.CommInitRank(&comm, nranks, id, rank));
  ctx->nccl_comm = comm;
  ctx->nranks = nranks;
  ctx->rank = rank;
  return LAT_OK;
}

extern "C" int lat_comm_destroy(lat_ctx* ctx) {
  if (!ctx) return LAT_ERR_ARG;
  if (ctx->nccl_comm) {
    cudaStreamSynchronize(ctx->stream);
    g_nccl.CommDestroy(ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
  }
  return LAT_OK;
}

extern "C" int lat_allreduce_sum(lat_ctx* ctx, double* buf, int64_t n) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, buf && n > 0);
  if (ctx->nranks == 1 || !ctx->nccl_comm) return LAT_OK;
  LAT_NCCL(ctx, g_nccl.AllReduce(buf, buf, (size_t)n, NCCL_FLOAT64, NCCL_SUM, ctx->nccl_comm, ctx->stream));
  return LAT_OK;
}

// sends/receives of one halo exchange; must be called between ncclGroupStart/End
static int halo_comm(lat_ctx* ctx, const lat_halo* h, double* vec, double* sendbuf) {
  int64_t so = 0, ro = 0;
  for (int i = 0; i < h->n_neighbors; ++i) {
    if (h->send_count[i] > 0)
      LAT_NCCL(ctx, g_nccl.Send(sendbuf + so * 6, (size_t)h->send_count[i] * 6, NCCL_FLOAT64, h->peer[i], ctx->nccl_comm, ctx->stream));
    if (h->recv_count[i] > 0)
      LAT_NCCL(ctx, g_nccl.Recv(vec + (h->n_owned + ro) * 6, (size_t)h->recv_count[i] * 6, NCCL_FLOAT64, h->peer[i], ctx->nccl_comm, ctx->stream));
    so += h->send_count[i];
    ro += h->recv_count[i];
  }
  return LAT_OK;
}

static int halo_pack(lat_ctx* ctx, const lat_halo* h, const double* vec, double** sendbuf_out) {
  int64_t tot_send = 0;
  for (int i = 0; i < h->n_neighbors; ++i) tot_send += h->send_count[i];
  double* sendbuf = lat_buf<double>(ctx, "halo_send", (size_t)tot_send * 6 + 8);
  if (!sendbuf) return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
  if (tot_send > 0)
    LAT_LAUNCH(ctx, k_pack_halo, (unsigned)ceil_div(tot_send * 6, 256), 256, 0, h->send_idx, tot_send, vec, sendbuf);
  *sendbuf_out = sendbuf;
  return LAT_OK;
}

static int halo_exchange(lat_ctx* ctx, const lat_halo* h, double* vec) {
  if (ctx->nranks == 1 || h->n_neighbors == 0) return LAT_OK;
  double* sendbuf = nullptr;
  int rc = halo_pack(ctx, h, vec, &sendbuf);
  if (rc) return rc;
  LAT_NCCL(ctx, g_nccl.GroupStart());
  rc = halo_comm(ctx, h, vec, sendbuf);
  if (rc) { g_nccl.GroupEnd(); return rc; }
  LAT_NCCL(ctx, g_nccl.GroupEnd());
  return LAT_OK;
}

extern "C" int lat_halo_exchange(lat_ctx* ctx, const lat_halo* halo, double* vec) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, halo && vec);
  LAT_CHECK_ARG(ctx, ctx->nranks == 1 || ctx->nccl_comm != nullptr);
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  return halo_exchange(ctx, halo, vec);
}

// ===========================================================================
// NVLink peer-memory path: halo push and all-reduce INSIDE our own kernels
// ===========================================================================
// Every rank owns one cudaMalloc'ed arena, IPC-mapped into all ranks of the node:
//   [ mailbox: 2 parities x MAXR x {v0,v1,v2,seq} | halo flags: MAXR x uint64 | u: 6 n_local doubles ]
// * halo:  k_p2p_halo (one CTA) stores the owned boundary entries of u straight into the neighbour's ghost section
//          (st.global on the mapped peer pointer), then publishes the sequence number with a system-scope
//          release and waits for the neighbours' flags; the SpMV kernel that follows is the plain one.
// * all-reduce: k_p2p_reduce sums the per-CTA partials and writes the 3 local sums into EVERY rank's mailbox
//          (slot = my rank) as 8-byte {32 data bits | 32 flag bits} words (one NVLink hop, no fence / flag
//          round trip), polls its own mailbox until every word carries this sequence's flag and adds the
//          contributions in rank order -- identical bits on every rank, no NCCL launch.
// All spins are bounded; a timeout sets breakdown = 2 and stops the solve instead of hanging the GPU.
static constexpr int P2P_MAXR = 16;
static constexpr int P2P_SLOTS = 32;          // halo flag slots per source rank (one per pushing CTA, wrapped)
// "LL" mailbox word (as in NCCL's low-latency protocol): 32 data bits + 32 flag bits in ONE 8-byte store, which
// is atomic, so no fence / separate flag round trip is needed: a double travels as two such words.
struct P2PArenaHdr {
  unsigned long long mail[2][P2P_MAXR][6];                  // [parity][source rank][3 doubles x {lo, hi}]
  unsigned long long halo_flag[P2P_MAXR][P2P_SLOTS];        // [source rank][slot]
  unsigned long long halo_cnt[P2P_MAXR];                    // fused-halo path: cumulative entries received per source rank
  unsigned long long n_local_nodes, pad[3];                 // of the arena's owner: places the halo inbox behind u
};
// Arena = header | u [6 n_local] | 256 B | halo inbox of the persistent kernel: two LL words per entry of u.
static inline size_t p2p_ll_offset(int64_t n_local) { return sizeof(P2PArenaHdr) + (size_t)n_local * 6 * sizeof(double) + 256; }
// ... | inbox of the coarse all-reduce (two-level preconditioner): [source rank][P2P_COARSE_CAP entries][2 LL words]
static constexpr int P2P_COARSE_CAP = 6 * 2048;
static inline size_t p2p_coarse_offset(int64_t n_local) { return p2p_ll_offset(n_local) + (size_t)n_local * 6 * 2 * sizeof(unsigned long long); }
static_assert(P2P_MAXR == 16, "halo_cnt replaces the former 128-byte pad");
static_assert(sizeof(P2PArenaHdr) % 32 == 0, "u behind the header is read with 256-bit loads");
static_assert(offsetof(P2PArenaHdr, mail) == 0 && sizeof(PersistRankMail) == sizeof(P2PArenaHdr::mail) && P2P_MAXR == 16,
              "pcg_persist.cuh posts its rank totals into the mail words at the head of the arena");
struct P2P {
  int nranks = 1, rank = 0;
  unsigned long long epoch = 0;
  unsigned char* arena = nullptr;           // my arena (device pointer)
  size_t arena_bytes = 0;
  int64_t n_local = 0;
  unsigned char* peer[P2P_MAXR] = {nullptr};  // mapped arenas of all ranks (peer[rank] == arena)
  // per neighbour (halo order): destination ghost offset (nodes) inside the peer's u
  int n_nb = 0;
  int nb_rank[P2P_MAXR];
  int64_t nb_dst_off[P2P_MAXR];
  bool attached = false;
  // device copies for kernels
  unsigned char** d_peer = nullptr;
  unsigned long long pushes_done = 0;   // fused-halo path: pushes completed by earlier solves (same on all ranks)
  unsigned long long ll_pushes = 0;     // persistent kernel: productions of u so far (tags of the halo inbox words)
  int64_t peer_n_local[P2P_MAXR] = {0}; // n_local of every rank (read from the mapped headers)
  // halo push/wait runs on a side stream next to the product of the interior rows (fork/join by events)
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Push the owned boundary entries of u into the ghost sections of the neighbours, then publish `seq`.
// send_idx: local node ids, neighbour-major; nb_first[k]..nb_first[k+1]: range of neighbour k.
struct P2PPushArgs {
  int n_nb;
  int nb_rank[4];
  int nb_first[5];
  int64_t nb_dst_node0[4];   // first destination node (peer local numbering) of my segment
  int my_rank;
};
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Push the owned boundary entries of u into the neighbours' ghost sections (grid-stride over the send list,
// st.global on the IPC-mapped peer pointers).  Every CTA then publishes `seq` into ITS flag slots of every
// neighbour's arena with one release store each (no local ticket, no second phase), and finally waits
// (bounded) until all slots of all neighbours in OUR arena carry `seq`.  The SpMV that follows is the plain
// k_cg_spmv: the kernel boundary orders it after this wait and starts with a clean L1.
__global__ void __launch_bounds__(256) k_p2p_halo(const int32_t* __restrict__ send_idx, const double* __restrict__ u,
                                                  unsigned char* const* __restrict__ peers, P2PPushArgs a, size_t u_off,
                                                  PcgScalars* __restrict__ sc, PcgParams prm) {
  if (sc->done || sc->iters >= prm.maxiter) return;
  const unsigned long long seq = prm.seq_base + (unsigned long long)sc->seq + 1ull;  // sequence of the upcoming SpMV
  const int total = a.nb_first[a.n_nb];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total * 6; i += gridDim.x * blockDim.x) {
    const int e = i / 6, d = i - e * 6;
    int k = 0;
    while (k + 1 < a.n_nb && e >= a.nb_first[k + 1]) ++k;
    double* dst = reinterpret_cast<double*>(peers[a.nb_rank[k]] + u_off);
    dst[(a.nb_dst_node0[k] + (e - a.nb_first[k])) * 6 + d] = u[(int64_t)send_idx[e] * 6 + d];
  }
  __syncthreads();   // the CTA's peer stores happen-before thread 0's release stores (cumulativity)
  // ONE system-scope fence, then relaxed flag stores spread over the lanes of warp 0 (a release store per
  // slot would serialise one fence + NVLink round trip per slot: measured 83 instead of 67 us per iteration)
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) __threadfence_system();
    __syncwarp();
    for (int k = 0; k < a.n_nb; ++k) {
      P2PArenaHdr* hdr = reinterpret_cast<P2PArenaHdr*>(peers[a.nb_rank[k]]);
      const int sl = blockIdx.x + (int)threadIdx.x * gridDim.x;
      if (sl < P2P_SLOTS)
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(&hdr->halo_flag[a.my_rank][sl]), "l"(seq) : "memory");
    }
  }
  // wait for the neighbours: lane l of warp 0 polls slot l of every neighbour
  if (threadIdx.x < P2P_SLOTS) {
    const P2PArenaHdr* mine = reinterpret_cast<const P2PArenaHdr*>(peers[a.my_rank]);
    for (int k = 0; k < a.n_nb; ++k) {
      long long spins = 0;
      while (ld_acquire_sys(&mine->halo_flag[a.nb_rank[k]][threadIdx.x]) < seq) {
        if (++spins > (1ll << 24)) { sc->p2p_timeout = 1; break; }
        __nanosleep(20);
      }
    }
  }
}

// One CTA: local sums -> every rank's mailbox -> wait for all -> rank-ordered total -> recurrences.
__global__ void __launch_bounds__(1024) k_p2p_reduce(const double* __restrict__ partials, int n_part,
                                                     PcgScalars* __restrict__ sc, PcgParams prm,
                                                     unsigned char* const* __restrict__ peers, int nranks, int my_rank) {
  if (sc->done || sc->iters >= prm.maxiter) return;
  __shared__ double s_loc[3];
  __shared__ double s_tot[P2P_MAXR][3];
  __shared__ int s_ok;
  double out[3];
  sum_partials<3, 1024>(partials, n_part, out);
  const unsigned long long seq = prm.seq_base + (unsigned long long)sc->seq + 1ull;
  // 32-bit flag in the upper half of each word: 12 bits of the solve epoch + 20 bits of the iteration
  // counter, never 0 (the arena starts zeroed) and never equal to what the previous solves left behind
  const unsigned long long flag = ((((seq >> 32) & 0xfffull) << 20) | ((seq & 0xfffffull) + 1ull) | 0x80000000ull) << 32;
  const int par = (int)(seq & 1ull);
  if (threadIdx.x == 0) { s_loc[0] = out[0]; s_loc[1] = out[1]; s_loc[2] = out[2]; s_ok = 1; }
  __syncthreads();
  // thread (q, w): word w (0..5) of my 3 sums into rank q's mailbox -- nranks * 6 independent 8-byte stores
  if (threadIdx.x < nranks * 6) {
    const int q = threadIdx.x / 6, w = threadIdx.x - q * 6;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(s_loc[w >> 1]);
    const unsigned long long half = (w & 1) ? (bits >> 32) : (bits & 0xffffffffull);
    P2PArenaHdr* hdr = reinterpret_cast<P2PArenaHdr*>(peers[q]);
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(&hdr->mail[par][my_rank][w]), "l"(flag | half) : "memory");
  }
  // thread (q, w) polls word w of rank q's contribution in MY mailbox until it carries this sequence's flag
  unsigned long long word = 0;
  if (threadIdx.x < nranks * 6) {
    const int q = threadIdx.x / 6, w = threadIdx.x - q * 6;
    const P2PArenaHdr* mine = reinterpret_cast<const P2PArenaHdr*>(peers[my_rank]);
    long long spins = 0;
    for (;;) {
      word = ld_relaxed_sys(&mine->mail[par][q][w]);
      if ((word & 0xffffffff00000000ull) == flag) break;
      if (++spins > (1ll << 24)) { s_ok = 0; break; }
      __nanosleep(20);
    }
  }
  // reassemble the doubles: lanes (q, 2k) and (q, 2k+1) hold lo / hi of sum k of rank q
  const unsigned long long other = __shfl_down_sync(0xffffffffu, word, 1);
  if (threadIdx.x < nranks * 6 && (threadIdx.x % 6) % 2 == 0) {
    const int q = threadIdx.x / 6, w = threadIdx.x - q * 6;
    const unsigned long long bits = (word & 0xffffffffull) | ((other & 0xffffffffull) << 32);
    s_tot[q][w >> 1] = __longlong_as_double((long long)bits);
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  double tot[3] = {0.0, 0.0, 0.0};
  for (int q = 0; q < nranks; ++q) { tot[0] += s_tot[q][0]; tot[1] += s_tot[q][1]; tot[2] += s_tot[q][2]; }   // rank order
  sc->seq = sc->seq + 1;
  if (!s_ok || sc->p2p_timeout) { sc->done = 1; sc->breakdown = 2; return; }
  cg_finish(sc, prm, tot[0], tot[1], tot[2]);
}

// Coarse all-reduce of the two-level preconditioner over peer memory: rc[q] = sum over the ranks of (sum of this rank's
// piece sums of entry q).  One thread per coarse entry: it adds its rank's pieces, writes the value as two LL words
// ({32 data | 32 flag}, one 16-byte store) into the inbox of EVERY rank (its own included), then polls its own inbox for
// the words of all ranks and adds them in rank order (bit-identical totals on all ranks).  The flag carries the solve
// epoch and a per-solve sequence number of the correction (2 iters + 1; set-up pass / restart: 2 iters), so words of
// earlier corrections never match.  Two corrections are always separated by the dot-product all-reduce of an iteration,
// which orders "every rank has read correction s" before "any rank writes correction s + 1": one inbox suffices.
struct CoarseP2P {
  size_t inbox_off[P2P_MAXR];   // byte offset of the coarse inbox inside each rank's arena
  int nranks, my_rank;
};
__global__ void __launch_bounds__(256) k_coarse_gather_p2p(const int32_t* __restrict__ agg_piece, const double* __restrict__ part,
                                                          double* __restrict__ rc, int n_c, PcgScalars* __restrict__ sc,
                                                          PcgParams prm, unsigned char* const* __restrict__ peers, CoarseP2P cp) {
  if (sc->done || sc->iters >= prm.maxiter) return;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_c) return;
  const int a = q / 6, p = q - a * 6;
  double s = 0.0;
  for (int k = agg_piece[a]; k < agg_piece[a + 1]; ++k) s += part[(int64_t)k * 6 + p];
  const unsigned long long it = 2ull * (unsigned long long)sc->iters + (sc->first ? 0ull : 1ull);
  const unsigned long long tag = (((((prm.seq_base >> 32) & 0x7ffull) << 20) | (it & 0xfffffull)) | 0x80000000ull) << 32;
  const unsigned long long bits = (unsigned long long)__double_as_longlong(s);
  const unsigned long long w0 = tag | (bits & 0xffffffffull), w1 = tag | (bits >> 32);
  for (int r = 0; r < cp.nranks; ++r) {
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(peers[r] + cp.inbox_off[r]) + ((size_t)cp.my_rank * P2P_COARSE_CAP + q) * 2;
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(w0), "l"(w1) : "memory");
  }
  const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(peers[cp.my_rank] + cp.inbox_off[cp.my_rank]);
  double tot = 0.0;
  for (int r = 0; r < cp.nranks; ++r) {
    const unsigned long long* src = mine + ((size_t)r * P2P_COARSE_CAP + q) * 2;
    unsigned long long v0, v1;
    long long spins = 0;
    for (;;) {
      asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v0), "=l"(v1) : "l"(src) : "memory");
      if ((v0 & 0xffffffff00000000ull) == tag && (v1 & 0xffffffff00000000ull) == tag) break;
      if (++spins > (1ll << 24)) { sc->p2p_timeout = 1; break; }
      if (spins > 64) __nanosleep(40);
    }
    tot += __longlong_as_double((long long)((v0 & 0xffffffffull) | (v1 << 32)));     // rank order
  }
  rc[q] = tot;
}

extern "C" int lat_p2p_arena_create(lat_ctx* ctx, int64_t n_local, void* handle64) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, n_local > 0 && handle64);
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!ctx->p2p) ctx->p2p = new P2P();
  P2P* p = ctx->p2p;
  if (p->attached) return lat_fail(ctx, LAT_ERR_STATE, "p2p arena already attached: destroy the comm first", __FILE__, __LINE__);
  if (p->arena) { cudaFree(p->arena); p->arena = nullptr; }
  p->arena_bytes = p2p_coarse_offset(n_local) + (size_t)P2P_MAXR * P2P_COARSE_CAP * 2 * sizeof(unsigned long long);
  LAT_CUDA(ctx, cudaMalloc(&p->arena, p->arena_bytes));
  LAT_CUDA(ctx, cudaMemset(p->arena, 0, p->arena_bytes));
  const unsigned long long nl = (unsigned long long)n_local;
  LAT_CUDA(ctx, cudaMemcpy(p->arena + offsetof(P2PArenaHdr, n_local_nodes), &nl, sizeof nl, cudaMemcpyHostToDevice));
  p->n_local = n_local;
  cudaIpcMemHandle_t h;
  LAT_CUDA(ctx, cudaIpcGetMemHandle(&h, p->arena));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handle64, &h, 64);
  return LAT_OK;
}

extern "C" int lat_p2p_attach(lat_ctx* ctx, const void* handles, int nranks, int rank, int n_neighbors,
                              const int32_t* nb_rank, const int64_t* nb_dst_node0) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, handles && nranks >= 1 && nranks <= P2P_MAXR && rank >= 0 && rank < nranks);
  LAT_CHECK_ARG(ctx, n_neighbors >= 0 && n_neighbors <= 4 && (n_neighbors == 0 || (nb_rank && nb_dst_node0)));
  P2P* p = ctx->p2p;
  if (!p || !p->arena) return lat_fail(ctx, LAT_ERR_STATE, "call lat_p2p_arena_create first", __FILE__, __LINE__);
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  p->nranks = nranks;
  p->rank = rank;
  for (int q = 0; q < nranks; ++q) {
    if (q == rank) { p->peer[q] = p->arena; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, reinterpret_cast<const unsigned char*>(handles) + 64 * q, 64);
    void* ptr = nullptr;
    LAT_CUDA(ctx, cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    p->peer[q] = reinterpret_cast<unsigned char*>(ptr);
  }
  for (int q = 0; q < nranks; ++q) {   // every owner wrote its header before the handles were exchanged
    unsigned long long nl = 0;
    LAT_CUDA(ctx, cudaMemcpy(&nl, p->peer[q] + offsetof(P2PArenaHdr, n_local_nodes), sizeof nl, cudaMemcpyDeviceToHost));
    p->peer_n_local[q] = (int64_t)nl;
  }
  p->n_nb = n_neighbors;
  for (int k = 0; k < n_neighbors; ++k) { p->nb_rank[k] = nb_rank[k]; p->nb_dst_off[k] = nb_dst_node0[k]; }
  if (!p->d_peer) LAT_CUDA(ctx, cudaMalloc(&p->d_peer, P2P_MAXR * sizeof(unsigned char*)));
  LAT_CUDA(ctx, cudaMemcpy(p->d_peer, p->peer, P2P_MAXR * sizeof(unsigned char*), cudaMemcpyHostToDevice));
  p->attached = true;
  return LAT_OK;
}

extern "C" int lat_p2p_destroy(lat_ctx* ctx) {
  if (!ctx || !ctx->p2p) return LAT_OK;
  P2P* p = ctx->p2p;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (int q = 0; q < p->nranks; ++q)
    if (q != p->rank && p->peer[q]) cudaIpcCloseMemHandle(p->peer[q]);
  if (p->d_peer) cudaFree(p->d_peer);
  if (p->arena) cudaFree(p->arena);
  if (p->side) cudaStreamDestroy(p->side);
  if (p->ev_fork) cudaEventDestroy(p->ev_fork);
  if (p->ev_join) cudaEventDestroy(p->ev_join);
  delete p;
  ctx->p2p = nullptr;
  return LAT_OK;
}

// Rows of the owned block that read at least one ghost column: they must wait for the halo, all others
// can be multiplied while the halo is in flight.  flag[row] = bit mask of the neighbours (halo order) whose
// ghosts the row reads; ghosts are stored neighbour-major behind the owned nodes.
struct GhostMap {
  int n_nb;
  int64_t first[5];   // first local node of neighbour k's ghost segment; first[n_nb] = n_local
};
__device__ __forceinline__ int ghost_bit(const GhostMap& gm, int64_t col) {
  int k = 0;
  while (k + 1 < gm.n_nb && col >= gm.first[k + 1]) ++k;
  return 1 << (k < 8 ? k : 7);
}
__global__ void k_mark_boundary_bsr(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                    int64_t n_own, GhostMap gm, uint8_t* __restrict__ flag) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_own) return;
  int f = 0;
  for (int j = rowptr[n]; j < rowptr[n + 1]; ++j)
    if (colidx[j] >= n_own) f |= ghost_bit(gm, colidx[j]);
  flag[n] = (uint8_t)f;
}
__global__ void k_mark_boundary_mf(MfOp op, int64_t n_own, GhostMap gm, uint8_t* __restrict__ flag) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_own) return;
  int f = 0;
  for (int j = op.adjptr[n]; j < op.adjptr[n + 1]; ++j)
    if (op.inc[j].other >= n_own) f |= ghost_bit(gm, op.inc[j].other);
  flag[n] = (uint8_t)f;
}
// Fused-halo path: per owned node the destination node inside neighbour 0 / 1 (-1: not sent there).
__global__ void k_build_push_dst(const int32_t* __restrict__ send_idx, P2PPushArgs a, int2* __restrict__ dst) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.nb_first[a.n_nb]) return;
  int k = 0;
  while (k + 1 < a.n_nb && e >= a.nb_first[k + 1]) ++k;
  const int v = (int)(a.nb_dst_node0[k] + (e - a.nb_first[k]));
  if (k == 0) dst[send_idx[e]].x = v;
  else dst[send_idx[e]].y = v;
}
// per-CTA OR of a per-row mask (rows_per_cta consecutive rows per CTA) / of "row has a push destination"
__global__ void k_cta_or_rows(const uint8_t* __restrict__ rowmask, int64_t n_rows, int rows_per_cta, uint8_t* __restrict__ out) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c * rows_per_cta >= n_rows) return;
  int f = 0;
  for (int64_t n = c * rows_per_cta; n < (c + 1) * rows_per_cta && n < n_rows; ++n) f |= rowmask[n];
  out[c] = (uint8_t)f;
}
__global__ void k_cta_or_push(const int2* __restrict__ dst, int64_t n_rows, int rows_per_cta, uint8_t* __restrict__ out) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c * rows_per_cta >= n_rows) return;
  int f = 0;
  for (int64_t n = c * rows_per_cta; n < (c + 1) * rows_per_cta && n < n_rows; ++n) f |= (dst[n].x >= 0) | ((dst[n].y >= 0) << 1);
  out[c] = (uint8_t)f;
}
// Stand-alone push with the counter protocol (set-up pass and restarts, where u comes from k_cg_init / k_cg_restart).
__global__ void __launch_bounds__(256) k_p2p_push_cnt(const int32_t* __restrict__ send_idx, const double* __restrict__ u,
                                                      unsigned char* const* __restrict__ peers, P2PPushArgs a, size_t u_off,
                                                      const PcgScalars* __restrict__ sc, PcgParams prm) {
  if (sc->done || sc->iters >= prm.maxiter) return;
  __shared__ unsigned int s_c[4];
  if (threadIdx.x < 4) s_c[threadIdx.x] = 0;
  __syncthreads();
  const int total = a.nb_first[a.n_nb];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total * 6; i += gridDim.x * blockDim.x) {
    const int e = i / 6, d = i - e * 6;
    int k = 0;
    while (k + 1 < a.n_nb && e >= a.nb_first[k + 1]) ++k;
    double* dst = reinterpret_cast<double*>(peers[a.nb_rank[k]] + u_off);
    dst[(a.nb_dst_node0[k] + (e - a.nb_first[k])) * 6 + d] = u[(int64_t)send_idx[e] * 6 + d];
    atomicAdd(&s_c[k], 1u);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.acq_rel.sys;" ::: "memory");
    for (int k = 0; k < a.n_nb; ++k)
      if (s_c[k]) {
        P2PArenaHdr* hdr = reinterpret_cast<P2PArenaHdr*>(peers[a.nb_rank[k]]);
        asm volatile("red.relaxed.sys.global.add.u64 [%0], %1;" ::"l"(&hdr->halo_cnt[a.my_rank]), "l"((unsigned long long)s_c[k]) : "memory");
      }
  }
}
// One CTA: ascending list of the flagged rows (deterministic order -> deterministic partial sums).
__global__ void __launch_bounds__(1024) k_compact_rows(const uint8_t* __restrict__ flag, int64_t n, int32_t* __restrict__ rows,
                                                       int64_t* __restrict__ count) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int64_t c0 = 0; c0 < n; c0 += 1024) {
    const int64_t i = c0 + threadIdx.x;
    const bool f = i < n && flag[i] != 0;
    const unsigned m = __ballot_sync(0xffffffffu, f);
    if (lane == 0) s_warp[wid] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
    for (int k = 0; k < 32; ++k) { const int c = s_warp[k]; if (k < wid) before += c; total += c; }
    const int base = s_base;
    if (f) rows[base + before + __popc(m & ((1u << lane) - 1u))] = (int32_t)i;
    __syncthreads();
    if (threadIdx.x == 0) s_base = base + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = s_base;
}

template <int PC>
static int pcg_run_dist_impl(lat_ctx* ctx, const int32_t* rowptr, const int32_t* colidx, const double* vals,
                             const lat_halo* h, const double* b, double* x, const lat_pcg_opts* o, lat_pcg_result* res,
                             const MfOp* mf);

// The iteration batches (kernels and, in the NCCL path, the collectives) are captured into a CUDA graph ON the
// stream they run on.  The legacy default stream -- what a caller without an own stream hands to lat_ctx_create --
// cannot be captured, so the whole solve then runs on the ctx-owned non-blocking stream, ordered after the caller's
// stream by an event; the solve ends with a synchronisation, which orders the caller's later work after it.
template <int PC>
static int pcg_run_dist(lat_ctx* ctx, const int32_t* rowptr, const int32_t* colidx, const double* vals,
                        const lat_halo* h, const double* b, double* x, const lat_pcg_opts* o, lat_pcg_result* res,
                        const MfOp* mf = nullptr) {
  cudaStream_t caller = ctx->stream;
  const bool legacy = caller == nullptr || caller == cudaStreamLegacy;
  if (legacy) {
    if (!ctx->work) LAT_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->work, cudaStreamNonBlocking));
    if (!ctx->ev_order) LAT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_order, cudaEventDisableTiming));
    LAT_CUDA(ctx, cudaEventRecord(ctx->ev_order, caller));
    LAT_CUDA(ctx, cudaStreamWaitEvent(ctx->work, ctx->ev_order, 0));
    ctx->stream = ctx->work;
  }
  const int rc = pcg_run_dist_impl<PC>(ctx, rowptr, colidx, vals, h, b, x, o, res, mf);
  if (legacy) {
    cudaStreamSynchronize(ctx->work);
    ctx->stream = caller;
  }
  return rc;
}

template <int PC>
static int pcg_run_dist_impl(lat_ctx* ctx, const int32_t* rowptr, const int32_t* colidx, const double* vals,
                             const lat_halo* h, const double* b, double* x, const lat_pcg_opts* o, lat_pcg_result* res,
                             const MfOp* mf) {
  // Chronopoulos-Gear arrangement: per iteration ONE halo exchange (u) and ONE all-reduce (3 doubles).
  const int64_t n_own = h->n_owned, n_loc = h->n_local;
  const int64_t n = 6 * n_loc;
  const unsigned grid = (unsigned)ceil_div(n_own, ROWS_PER_CTA);
  const unsigned mf_grid = (unsigned)ceil_div(n_own, MF_BLOCK);
  const int n_part = (int)(mf ? mf_grid : grid);   // CTAs of the product kernel = partial sums to add
  // bit 4 of `reserved`: halo push + all-reduce through NVLink peer memory inside our kernels (no NCCL)
  P2P* pp = ctx->p2p;
  const bool p2p = (o->reserved & 16) && pp && pp->attached && pp->nranks == ctx->nranks && pp->n_local == n_loc &&
                   pp->n_nb == h->n_neighbors;
  if ((o->reserved & 16) && !p2p)
    return lat_fail(ctx, LAT_ERR_STATE, "peer-memory path requested but no matching arena is attached", __FILE__, __LINE__);
  double* r = lat_buf<double>(ctx, "pcg_r", n);
  double* u = p2p ? reinterpret_cast<double*>(pp->arena + sizeof(P2PArenaHdr)) : lat_buf<double>(ctx, "pcg_z", n);
  double* p = lat_buf<double>(ctx, "pcg_pa", n);
  double* sv = lat_buf<double>(ctx, "pcg_pb", n);
  double* w = lat_buf<double>(ctx, "pcg_Ap", n);
  double* dinv = lat_buf<double>(ctx, "pcg_dinv", PC == LAT_PC_BLOCK6 ? 21 * n_own : 6 * n_own);
  double* partials = lat_buf<double>(ctx, "pcg_partials", (size_t)4 * 2 * grid + 8);   // interior + boundary launches
  PcgScalars* sc = lat_buf<PcgScalars>(ctx, "pcg_scalars", 1);
  if (!r || !u || !p || !sv || !w || !dinv || !partials || !sc)
    return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
  if (o->reference_semantics)
    return lat_fail(ctx, LAT_ERR_UNSUPPORTED, "reference_semantics is single-GPU only", __FILE__, __LINE__);
  const int64_t launches0 = ctx->launches;
  PcgParams prm;
  prm.tol = o->tol; prm.mintol = 0.0; prm.alpha_max = 0.0; prm.restart_every = 0;
  prm.maxiter = o->maxiter; prm.reference = 0; prm.dist = 1; prm.l2_keep = 0; prm.seq_base = 0; prm.push_base = 0;
  int check = o->check_every > 0 ? o->check_every : 32;
  const bool multi = ctx->nranks > 1 && ctx->nccl_comm != nullptr;
  // Two-level preconditioner (coarse.cuh) across ranks: every rank restricts the residual of its OWNED nodes, the coarse
  // residual (6 n_agg doubles) is all-reduced, every rank multiplies the rows of the dense inverse that belong to
  // aggregates it owns nodes of and corrects its owned u.  The correction changes u after the update kernel, so the halo
  // cannot be pushed by that kernel: the separate halo kernel / NCCL exchange is used, and the persistent kernel is skipped.
  const bool coarse = ctx->coarse.active;
  if (coarse && ctx->coarse.n_nodes != n_loc)
    return lat_fail(ctx, LAT_ERR_STATE, "the registered coarse space belongs to another system: call lat_coarse_setup / lat_coarse_set_inverse(NULL)", __FILE__, __LINE__);
  CoarseLaunch cl = coarse ? coarse_launch(ctx) : CoarseLaunch();
  const int64_t coarse_nc = 6 * (int64_t)ctx->coarse.n_agg;
  // peer-memory path: the coarse residual is all-reduced by k_coarse_gather_p2p (LL words into every rank's inbox)
  // instead of NCCL -- ~100 us less per iteration, measured on 2 GPUs
  const bool coarse_p2p = coarse && p2p && multi && coarse_nc <= P2P_COARSE_CAP && !getenv("LAT_COARSE_NCCL");
  CoarseP2P ccp;
  memset(&ccp, 0, sizeof ccp);
  if (coarse_p2p) {
    cl.fused = false;               // the piece sums feed the cross-rank gather
    ccp.nranks = pp->nranks;
    ccp.my_rank = pp->rank;
    for (int q = 0; q < pp->nranks; ++q) ccp.inbox_off[q] = p2p_coarse_offset(pp->peer_n_local[q]);
  }
  // the product: assembled BSR rows or the matrix-free operator (owned rows only, ghosts are read)
  auto launch_product = [&]() -> int {
    if (mf) LAT_LAUNCH(ctx, k_cg_spmv_mf<false>, mf_grid, MF_BLOCK, 0, *mf, n_own, u, r, w, sc, partials, prm, RowSet());
    else LAT_LAUNCH(ctx, k_cg_spmv<false>, grid, SPMV_BLOCK, 0, rowptr, colidx, vals, n_own, u, r, w, sc, partials, prm, RowSet());
    return LAT_OK;
  };
  P2PPushArgs pa;
  memset(&pa, 0, sizeof pa);
  int push_total = 0;
  if (p2p) {
    pp->epoch += 1;
    prm.seq_base = pp->epoch << 32;
    pa.n_nb = h->n_neighbors;
    pa.my_rank = pp->rank;
    for (int k = 0; k < h->n_neighbors; ++k) {
      if (h->peer[k] != pp->nb_rank[k])
        return lat_fail(ctx, LAT_ERR_STATE, "halo neighbour order differs from the attached arena", __FILE__, __LINE__);
      pa.nb_rank[k] = h->peer[k];
      pa.nb_first[k] = push_total;
      pa.nb_dst_node0[k] = pp->nb_dst_off[k];
      push_total += h->send_count[k];
    }
    pa.nb_first[h->n_neighbors] = push_total;
  }
  const size_t u_off = sizeof(P2PArenaHdr);
  unsigned halo_grid = (unsigned)ceil_div((int64_t)push_total * 6, 256 * 8);   // ~8 entries per thread
  if (halo_grid < 1) halo_grid = 1;
  if (halo_grid > (unsigned)P2P_SLOTS) halo_grid = P2P_SLOTS;
  {  // the halo send buffer must exist before any stream capture (allocation is not capturable)
    int64_t tot_send = 0;
    for (int i = 0; i < h->n_neighbors; ++i) tot_send += h->send_count[i];
    if (!lat_buf<double>(ctx, "halo_send", (size_t)tot_send * 6 + 8))
      return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
  }
  // Overlap (peer-memory path): the rows that read ghost columns are listed once per solve; per iteration the
  // halo push/wait runs on a side stream while the main stream multiplies all other rows, then the listed
  // rows follow.  OPT-IN (bit 5 of `reserved`): measured on 2 x B200 it is no faster than the sequential
  // halo -> product (profiles/r01_overlap_ab.txt): NVLink moves a 1 MB halo in ~3 us, what is left is launch and
  // flag latency, and the fork/join plus the extra boundary launch cost as much as the overlap hides.
  GhostMap gmap;
  memset(&gmap, 0, sizeof gmap);
  gmap.n_nb = h->n_neighbors < 4 ? h->n_neighbors : 4;
  {
    int64_t off = n_own;
    for (int k = 0; k < gmap.n_nb; ++k) { gmap.first[k] = off; off += h->recv_count[k]; }
    gmap.first[gmap.n_nb] = n_loc;
  }
  int64_t n_bnd = 0;
  unsigned bnd_grid = 0;
  uint8_t* bflag = nullptr;
  int32_t* brows = nullptr;
  const bool overlap = p2p && h->n_neighbors > 0 && (o->reserved & 32) && !coarse;
  if (overlap) {
    bflag = lat_buf<uint8_t>(ctx, "pcg_bflag", n_own);
    brows = lat_buf<int32_t>(ctx, "pcg_brows", n_own);
    int64_t* d_cnt = lat_buf<int64_t>(ctx, "pcg_bcount", 1);
    if (!bflag || !brows || !d_cnt) return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
    if (mf) LAT_LAUNCH(ctx, k_mark_boundary_mf, (unsigned)ceil_div(n_own, 256), 256, 0, *mf, n_own, gmap, bflag);
    else LAT_LAUNCH(ctx, k_mark_boundary_bsr, (unsigned)ceil_div(n_own, 256), 256, 0, rowptr, colidx, n_own, gmap, bflag);
    LAT_LAUNCH(ctx, k_compact_rows, 1, 1024, 0, bflag, n_own, brows, d_cnt);
    LAT_CUDA(ctx, cudaMemcpyAsync(ctx->h_i64, d_cnt, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    LAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    n_bnd = ctx->h_i64[0];
    bnd_grid = (unsigned)ceil_div(n_bnd, mf ? MF_BLOCK : ROWS_PER_CTA);
    if (!pp->side) {
      // highest priority: the few CTAs of the halo kernel must be scheduled ahead of the product's thousands,
      // otherwise the push (and with it every neighbour) waits for the interior product to drain
      int prio_lo = 0, prio_hi = 0;
      LAT_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
      LAT_CUDA(ctx, cudaStreamCreateWithPriority(&pp->side, cudaStreamNonBlocking, prio_hi));
      LAT_CUDA(ctx, cudaEventCreateWithFlags(&pp->ev_fork, cudaEventDisableTiming));
      LAT_CUDA(ctx, cudaEventCreateWithFlags(&pp->ev_join, cudaEventDisableTiming));
    }
  }
  // Fused halo (default for <= 2 neighbours, i.e. slab partitions; bit 6 of `reserved` selects the separate
  // halo kernel instead): no halo kernel in the iteration.  The
  // update kernel pushes the boundary entries of the new u straight into the neighbours' ghost sections and bumps
  // their entry counters; only the product CTAs that own a ghost-reading row wait for the counter.
  const bool fused = p2p && !(o->reserved & 64) && h->n_neighbors >= 1 && h->n_neighbors <= 2 && !overlap && !coarse;
  HaloPush hpush;
  RowSet wait_rs;
  if (fused) {
    if (!bflag) {
      bflag = lat_buf<uint8_t>(ctx, "pcg_bflag", n_own);
      if (!bflag) return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
      if (mf) LAT_LAUNCH(ctx, k_mark_boundary_mf, (unsigned)ceil_div(n_own, 256), 256, 0, *mf, n_own, gmap, bflag);
      else LAT_LAUNCH(ctx, k_mark_boundary_bsr, (unsigned)ceil_div(n_own, 256), 256, 0, rowptr, colidx, n_own, gmap, bflag);
    }
    int2* pdst = lat_buf<int2>(ctx, "pcg_push_dst", n_own);
    if (!pdst) return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
    LAT_CUDA(ctx, cudaMemsetAsync(pdst, 0xff, (size_t)n_own * sizeof(int2), ctx->stream));
    if (push_total > 0) LAT_LAUNCH(ctx, k_build_push_dst, (unsigned)ceil_div(push_total, 256), 256, 0, h->send_idx, pa, pdst);
    const int rpc = mf ? MF_BLOCK : ROWS_PER_CTA;              // rows per CTA of the product kernel
    const unsigned n_cta_prod = mf ? mf_grid : grid;
    uint8_t* cta_need = lat_buf<uint8_t>(ctx, "pcg_cta_need", n_cta_prod);
    uint8_t* cta_push = lat_buf<uint8_t>(ctx, "pcg_cta_push", grid);
    if (!cta_need || !cta_push) return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
    LAT_LAUNCH(ctx, k_cta_or_rows, (unsigned)ceil_div(n_cta_prod, 128), 128, 0, bflag, n_own, rpc, cta_need);
    LAT_LAUNCH(ctx, k_cta_or_push, (unsigned)ceil_div(grid, 128), 128, 0, pdst, n_own, (int)ROWS_PER_CTA, cta_push);
    hpush.dst = pdst;
    hpush.cta_push = cta_push;
    wait_rs.cta_need = cta_need;
    wait_rs.need = bflag;
    wait_rs.cnt = reinterpret_cast<const P2PArenaHdr*>(pp->arena)->halo_cnt;
    wait_rs.n_nb = h->n_neighbors;
    wait_rs.n_own = n_own;
    for (int k = 0; k < h->n_neighbors; ++k) {
      hpush.peer_u[k] = reinterpret_cast<double*>(pp->peer[h->peer[k]] + u_off);
      hpush.peer_cnt[k] = &reinterpret_cast<P2PArenaHdr*>(pp->peer[h->peer[k]])->halo_cnt[pp->rank];
      wait_rs.nb_rank[k] = h->peer[k];
      wait_rs.per_push[k] = (unsigned long long)h->recv_count[k] * 6ull;
    }
    prm.push_base = pp->pushes_done;
  }
  auto push_now = [&]() -> int {   // u was produced by a kernel without the fused push
    if (fused) LAT_LAUNCH(ctx, k_p2p_push_cnt, halo_grid, 256, 0, h->send_idx, u, pp->d_peer, pa, u_off, sc, prm);
    return LAT_OK;
  };
  const int n_part_ov = n_part + (int)bnd_grid;
  auto launch_product_overlapped = [&]() -> int {
    // fork: halo on the side stream
    LAT_CUDA(ctx, cudaEventRecord(pp->ev_fork, ctx->stream));
    LAT_CUDA(ctx, cudaStreamWaitEvent(pp->side, pp->ev_fork, 0));
    k_p2p_halo<<<halo_grid, 256, 0, pp->side>>>(h->send_idx, u, pp->d_peer, pa, u_off, sc, prm);
    ++ctx->launches;
    LAT_CUDA(ctx, cudaEventRecord(pp->ev_join, pp->side));
    RowSet in_rs, bd_rs;
    in_rs.skip = bflag; in_rs.p_stride = n_part_ov; in_rs.p_offset = 0;
    bd_rs.rows = brows; bd_rs.p_stride = n_part_ov; bd_rs.p_offset = n_part;
    if (mf) LAT_LAUNCH(ctx, (k_cg_spmv_mf<false, true>), mf_grid, MF_BLOCK, 0, *mf, n_own, u, r, w, sc, partials, prm, in_rs);
    else LAT_LAUNCH(ctx, (k_cg_spmv<false, true>), grid, SPMV_BLOCK, 0, rowptr, colidx, vals, n_own, u, r, w, sc, partials, prm, in_rs);
    // join: the listed rows need the ghosts
    LAT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, pp->ev_join, 0));
    if (bnd_grid > 0) {
      if (mf) LAT_LAUNCH(ctx, (k_cg_spmv_mf<false, true>), bnd_grid, MF_BLOCK, 0, *mf, n_bnd, u, r, w, sc, partials, prm, bd_rs);
      else LAT_LAUNCH(ctx, (k_cg_spmv<false, true>), bnd_grid, SPMV_BLOCK, 0, rowptr, colidx, vals, n_bnd, u, r, w, sc, partials, prm, bd_rs);
    }
    return LAT_OK;
  };
  // Slabs that fit the shared memory of every rank: the whole solve in ONE persistent kernel per GPU, halo and
  // all-reduce inside its two grid barriers (pcg_persist.cuh, "Multi-GPU").  Bit 7 of `reserved` opts out.
  if (fused && !mf && multi && !(o->reserved & 128) && o->profile_iters <= 0) {
    PersistDistSetup ds;
    ds.u = u;
    ds.nranks = pp->nranks; ds.my_rank = pp->rank; ds.n_nb = h->n_neighbors;
    for (int k = 0; k < h->n_neighbors; ++k) {
      ds.ghost_first[k] = 6 * gmap.first[k];
      ds.ghost_entries[k] = 6 * h->recv_count[k];
      ds.peer_ll[k] = reinterpret_cast<unsigned long long*>(pp->peer[h->peer[k]] + p2p_ll_offset(pp->peer_n_local[h->peer[k]]));
    }
    ds.push_base = pp->ll_pushes; ds.seq_base = prm.seq_base;
    ds.push_dst = hpush.dst; ds.peers = pp->d_peer;
    ds.my_ll = reinterpret_cast<const unsigned long long*>(pp->arena + p2p_ll_offset(pp->n_local));
    bool used = false;
    const int prc = pcg_run_persist<PC>(ctx, rowptr, colidx, vals, n_own, b, x, o, res, &used, &ds);
    if (prc != LAT_OK) return prc;
    if (used) {
      pp->ll_pushes += ds.pushes;
      return LAT_OK;
    }
  }
  LAT_CUDA(ctx, cudaMemsetAsync(sc, 0, sizeof(PcgScalars), ctx->stream));
  const int32_t one = 1;
  LAT_CUDA(ctx, cudaMemcpyAsync(&sc->first, &one, sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  // Only the OWNED part of u is initialised here.  In the peer-memory path a neighbour that is ahead may
  // already have pushed its first halo of this solve into our ghost section: zeroing it would lose data
  // (observed: 1634 instead of 1246 iterations and a wrong iterate on the first solve).
  if (!p2p) LAT_CUDA(ctx, cudaMemsetAsync(u, 0, n * sizeof(double), ctx->stream));
  if (PC != LAT_PC_NONE) {
    if (mf) LAT_LAUNCH(ctx, k_mf_precond, (unsigned)ceil_div(n_own, 128), 128, 0, *mf, n_own, PC, dinv);
    else LAT_LAUNCH(ctx, k_precond_setup, (unsigned)ceil_div(n_own, 128), 128, 0, rowptr, colidx, vals, n_own, PC, dinv);
  }
  LAT_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
  LAT_LAUNCH(ctx, k_cg_init<PC>, grid, SPMV_BLOCK, 0, n_own, b, dinv, x, r, u, p, sv);
  auto coarse_correction = [&]() -> int {    // u += Z Einv Z^T r on the owned nodes (r, u of the current iterate)
    if (!coarse) return LAT_OK;
    if (coarse_p2p) {
      cl.restrict_to_coarse(ctx->stream, r, sc, prm.maxiter, false);
      k_coarse_gather_p2p<<<(unsigned)ceil_div(coarse_nc, 256), 256, 0, ctx->stream>>>(cl.agg_piece, cl.part, cl.rc, (int)coarse_nc, sc,
                                                                                    prm, pp->d_peer, ccp);
    } else {
      cl.restrict_to_coarse(ctx->stream, r, sc, prm.maxiter);
      if (multi) { if (int arc = lat_allreduce_sum(ctx, cl.rc, coarse_nc)) return arc; }
    }
    cl.correct(ctx->stream, u, sc, prm.maxiter);
    ctx->launches += cl.launches();
    cudaError_t ce2 = cudaPeekAtLastError();
    if (ce2 != cudaSuccess) return lat_cuda_fail(ctx, ce2, "coarse correction", __FILE__, __LINE__);
    return LAT_OK;
  };
  if (int crc0 = coarse_correction()) return crc0;

  // halo(u) -> w = A u, partial dots -> local sums -> all-reduce -> scalar recurrences / stop test
  cudaEvent_t prof_ev[2] = {nullptr, nullptr};
  cudaEvent_t tr_ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  double tr_ms[4] = {0.0, 0.0, 0.0, 0.0};
  const bool trace = getenv("LAT_P2P_TRACE") != nullptr;
  bool prof_on = false;
  double prof_ms = 0.0;
  int prof_n = 0;
  auto spmv_and_reduce = [&]() -> int {
    if (fused) {
      if (prof_on) cudaEventRecord(prof_ev[0], ctx->stream);
      if (mf) LAT_LAUNCH(ctx, k_cg_spmv_mf<true>, mf_grid, MF_BLOCK, 0, *mf, n_own, u, r, w, sc, partials, prm, wait_rs);
      else LAT_LAUNCH(ctx, k_cg_spmv<true>, grid, SPMV_BLOCK, 0, rowptr, colidx, vals, n_own, u, r, w, sc, partials, prm, wait_rs);
      if (prof_on) cudaEventRecord(prof_ev[1], ctx->stream);
      LAT_LAUNCH(ctx, k_p2p_reduce, 1, 1024, 0, partials, n_part, sc, prm, pp->d_peer, pp->nranks, pp->rank);
      if (prof_on) {
        cudaEventSynchronize(prof_ev[1]);
        float ms1 = 0.f;
        cudaEventElapsedTime(&ms1, prof_ev[0], prof_ev[1]);   // includes the halo wait of the boundary CTAs
        prof_ms += ms1;
        ++prof_n;
      }
      return LAT_OK;
    }
    if (p2p && overlap && !prof_on) {
      if (int prc = launch_product_overlapped()) return prc;
      LAT_LAUNCH(ctx, k_p2p_reduce, 1, 1024, 0, partials, n_part_ov, sc, prm, pp->d_peer, pp->nranks, pp->rank);
      return LAT_OK;
    }
    if (p2p) {
      if (prof_on && trace) cudaEventRecord(tr_ev[1], ctx->stream);
      if (h->n_neighbors > 0)
        LAT_LAUNCH(ctx, k_p2p_halo, halo_grid, 256, 0, h->send_idx, u, pp->d_peer, pa, u_off, sc, prm);
      if (prof_on && trace) cudaEventRecord(tr_ev[2], ctx->stream);
      if (prof_on) cudaEventRecord(prof_ev[0], ctx->stream);
      if (int prc = launch_product()) return prc;
      if (prof_on) cudaEventRecord(prof_ev[1], ctx->stream);
      LAT_LAUNCH(ctx, k_p2p_reduce, 1, 1024, 0, partials, n_part, sc, prm, pp->d_peer, pp->nranks, pp->rank);
      if (prof_on && trace) cudaEventRecord(tr_ev[4], ctx->stream);
      if (prof_on) {
        cudaEventSynchronize(prof_ev[1]);
        float ms1 = 0.f;
        cudaEventElapsedTime(&ms1, prof_ev[0], prof_ev[1]);
        prof_ms += ms1;
        ++prof_n;
        if (trace) {
          cudaEventSynchronize(tr_ev[4]);
          float a = 0.f, b2 = 0.f, c2 = 0.f;
          cudaEventElapsedTime(&a, tr_ev[0], tr_ev[1]);   // update kernel
          cudaEventElapsedTime(&b2, tr_ev[1], tr_ev[2]);  // halo push + wait
          cudaEventElapsedTime(&c2, prof_ev[1], tr_ev[4]); // reduce + mailbox all-reduce
          tr_ms[0] += a; tr_ms[1] += b2; tr_ms[2] += ms1; tr_ms[3] += c2;
        }
      }
      return LAT_OK;
    }
    int rc2 = multi ? halo_exchange(ctx, h, u) : LAT_OK;
    if (rc2) return rc2;
    if (int prc = launch_product()) return prc;
    LAT_LAUNCH(ctx, k_cg_reduce, 1, CG_REDUCE_BLOCK, 0, partials, n_part, sc, prm);
    rc2 = lat_allreduce_sum(ctx, sc->sums, 3);
    if (rc2) return rc2;
    LAT_LAUNCH(ctx, k_cg_finalize, 1, 1, 0, sc, prm);
    return LAT_OK;
  };
  auto one_iteration = [&]() -> int {
    if (prof_on && trace) cudaEventRecord(tr_ev[0], ctx->stream);
    LAT_LAUNCH(ctx, k_cg_update<PC>, grid, SPMV_BLOCK, 0, n_own, dinv, x, r, u, w, p, sv, sc, prm, hpush);
    if (int crc1 = coarse_correction()) return crc1;
    return spmv_and_reduce();
  };
  int rc = push_now();
  if (rc) return rc;
  rc = spmv_and_reduce();  // set-up pass
  if (rc) return rc;

  // `check` iterations (kernels AND NCCL operations) are captured into one CUDA graph
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t gexec = nullptr;
  bool use_graph = (o->reserved & 4) == 0 && o->maxiter >= check;
  int64_t per_iter = 0;
  if (use_graph) {
    const int64_t l0 = ctx->launches;
    cudaError_t ce = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal);
    int crc = LAT_OK;
    if (ce == cudaSuccess) {
      for (int q = 0; q < check && crc == LAT_OK; ++q) crc = one_iteration();
      ce = cudaStreamEndCapture(ctx->stream, &graph);
    }
    per_iter = (ctx->launches - l0) / check;
    ctx->launches = l0;
    if (ce != cudaSuccess || crc != LAT_OK || graph == nullptr) {
      cudaGetLastError();
      if (graph) cudaGraphDestroy(graph);
      graph = nullptr;
      use_graph = false;
    } else if (cudaGraphInstantiate(&gexec, graph, 0) != cudaSuccess) {
      cudaGetLastError();
      cudaGraphDestroy(graph);
      graph = nullptr;
      use_graph = false;
    }
  }
  PcgScalars* hs = ctx->h_scal;
  auto run_batches = [&](int64_t budget) -> int {
    int64_t it = 0;
    bool finished = budget <= 0;
    while (!finished) {
      const int batch = (int)((budget - it) < check ? (budget - it) : check);
      if (use_graph && batch == check) {
        cudaError_t ce = cudaGraphLaunch(gexec, ctx->stream);
        if (ce != cudaSuccess) return lat_cuda_fail(ctx, ce, "cudaGraphLaunch", __FILE__, __LINE__);
        ctx->launches += per_iter * check;
      } else {
        for (int q = 0; q < batch; ++q) {
          const int rc3 = one_iteration();
          if (rc3) return rc3;
        }
      }
      it += batch;
      cudaMemcpyAsync(&hs[0], sc, sizeof(PcgScalars), cudaMemcpyDeviceToHost, ctx->stream);
      cudaError_t ce = cudaStreamSynchronize(ctx->stream);
      if (ce != cudaSuccess) return lat_cuda_fail(ctx, ce, "distributed PCG iteration", __FILE__, __LINE__);
      if (hs[0].done || hs[0].iters >= o->maxiter || it >= budget) finished = true;
    }
    return LAT_OK;
  };
  // optional: the first profile_iters iterations run outside the graph with events around the SpMV kernel
  int nprof = (p2p && o->profile_iters > 0) ? o->profile_iters : 0;
  if (nprof > 64) nprof = 64;
  if (nprof > o->maxiter) nprof = o->maxiter;
  if (nprof > 0) {
    cudaEventCreate(&prof_ev[0]);
    cudaEventCreate(&prof_ev[1]);
    for (auto& e : tr_ev) cudaEventCreate(&e);
    prof_on = true;
    for (int q = 0; q < nprof && rc == LAT_OK; ++q) rc = one_iteration();
    prof_on = false;
    cudaEventDestroy(prof_ev[0]);
    cudaEventDestroy(prof_ev[1]);
    for (auto& e : tr_ev) cudaEventDestroy(e);
    if (trace && prof_n > 0)
      fprintf(stderr, "[lat p2p trace] rank %d: update %.1f us | halo push+wait %.1f us | spmv %.1f us | reduce+allreduce %.1f us (mean of %d)\n",
              ctx->rank, 1e3 * tr_ms[0] / prof_n, 1e3 * tr_ms[1] / prof_n, 1e3 * tr_ms[2] / prof_n, 1e3 * tr_ms[3] / prof_n, prof_n);
    if (rc) return rc;
  }
  rc = run_batches((int64_t)o->maxiter - nprof);
  // true-residual safeguard (see k_cg_restart): x needs its ghosts, |r_true|^2 is all-reduced
  double true_rr = -1.0;
  while (rc == LAT_OK && hs[0].done && !hs[0].breakdown && hs[0].bb > 0.0) {
    if (multi) { rc = halo_exchange(ctx, h, x); if (rc) break; }
    if (mf) { k_mf_apply<true><<<mf_grid, MF_BLOCK, 0, ctx->stream>>>(*mf, n_own, x, w); ++ctx->launches; }
    else rc = lat_spmv_internal(ctx, rowptr, colidx, vals, n_own, x, w);
    if (rc) break;
    k_cg_restart<PC><<<grid, SPMV_BLOCK, 0, ctx->stream>>>(n_own, b, w, dinv, r, u, p, sv, partials);
    k_cg_true_residual<<<1, CG_REDUCE_BLOCK, 0, ctx->stream>>>(partials, (int)grid, sc, prm, 0);
    rc = lat_allreduce_sum(ctx, sc->sums + 3, 1);
    if (rc) break;
    k_cg_true_residual<<<1, CG_REDUCE_BLOCK, 0, ctx->stream>>>(partials, (int)grid, sc, prm, 1);
    ctx->launches += 3;
    cudaMemcpyAsync(&hs[0], sc, sizeof(PcgScalars), cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t ce = cudaStreamSynchronize(ctx->stream);
    if (ce != cudaSuccess) { rc = lat_cuda_fail(ctx, ce, "distributed PCG residual check", __FILE__, __LINE__); break; }
    true_rr = hs[0].true_rr;
    if (hs[0].done) break;
    rc = coarse_correction();
    if (rc) break;
    rc = push_now();
    if (rc) break;
    rc = spmv_and_reduce();   // restart from x: set-up pass
    if (rc) break;
    rc = run_batches((int64_t)o->maxiter - hs[0].iters);
  }
  if (gexec) cudaGraphExecDestroy(gexec);
  if (graph) cudaGraphDestroy(graph);
  if (rc) return rc;
  LAT_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
  LAT_CUDA(ctx, cudaMemcpyAsync(&hs[0], sc, sizeof(PcgScalars), cudaMemcpyDeviceToHost, ctx->stream));
  LAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (fused) pp->pushes_done += (unsigned long long)hs[0].seq;   // identical on every rank (global stop test)
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
  res->iters = hs[0].iters;
  res->norm_b = sqrt(hs[0].bb);
  res->relres = hs[0].bb > 0.0 ? sqrt(hs[0].rr / hs[0].bb) : 0.0;
  res->info = hs[0].done && !hs[0].breakdown ? 0 : (hs[0].breakdown == 2 ? 4 : (hs[0].breakdown == 3 ? 5 : (hs[0].breakdown ? 3 : 1)));
  res->solve_ms = ms;
  res->launches = ctx->launches - launches0;
  res->spmv_ms = prof_n > 0 ? prof_ms / prof_n : 0.0; res->update_ms = 0.0; res->profiled = prof_n; res->reserved = hs[0].restarts | (use_graph ? 0x100 : 0) | (coarse ? 0x400 : 0);
  res->true_relres = (true_rr >= 0.0 && hs[0].bb > 0.0) ? sqrt(true_rr / hs[0].bb) : -1.0;
  return LAT_OK;
}

extern "C" int lat_pcg_bsr_dist(lat_ctx* ctx, const int32_t* rowptr, const int32_t* colidx, const double* vals,
                                const lat_halo* halo, const double* b, double* x, const lat_pcg_opts* opts,
                                lat_pcg_result* result) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, rowptr && colidx && vals && halo && b && x && opts && result);
  LAT_CHECK_ARG(ctx, halo->n_owned > 0 && halo->n_local >= halo->n_owned && opts->maxiter >= 0);
  LAT_CHECK_ARG(ctx, ctx->nranks == 1 || ctx->nccl_comm != nullptr);
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  switch (opts->precond) {
    case LAT_PC_NONE: return pcg_run_dist<LAT_PC_NONE>(ctx, rowptr, colidx, vals, halo, b, x, opts, result);
    case LAT_PC_JACOBI: return pcg_run_dist<LAT_PC_JACOBI>(ctx, rowptr, colidx, vals, halo, b, x, opts, result);
    case LAT_PC_BLOCK6: return pcg_run_dist<LAT_PC_BLOCK6>(ctx, rowptr, colidx, vals, halo, b, x, opts, result);
    default: return lat_fail(ctx, LAT_ERR_ARG, "unknown preconditioner", __FILE__, __LINE__);
  }
}

extern "C" int lat_pcg_matfree_dist(lat_ctx* ctx, const lat_halo* halo, const double* b, double* x,
                                    const lat_pcg_opts* opts, lat_pcg_result* result) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, halo && b && x && opts && result);
  LAT_CHECK_ARG(ctx, halo->n_owned > 0 && halo->n_local >= halo->n_owned && opts->maxiter >= 0);
  LAT_CHECK_ARG(ctx, ctx->nranks == 1 || ctx->nccl_comm != nullptr);
  LAT_CHECK_ARG(ctx, (reinterpret_cast<uintptr_t>(x) & 31) == 0);
  MfOp op;
  if (int rc = mf_get(ctx, &op)) return rc;
  if (ctx->mf_nnodes != halo->n_local)
    return lat_fail(ctx, LAT_ERR_STATE, "resident matrix-free operator was set up for a different local mesh", __FILE__, __LINE__);
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  switch (opts->precond) {
    case LAT_PC_NONE: return pcg_run_dist<LAT_PC_NONE>(ctx, nullptr, nullptr, nullptr, halo, b, x, opts, result, &op);
    case LAT_PC_JACOBI: return pcg_run_dist<LAT_PC_JACOBI>(ctx, nullptr, nullptr, nullptr, halo, b, x, opts, result, &op);
    case LAT_PC_BLOCK6: return pcg_run_dist<LAT_PC_BLOCK6>(ctx, nullptr, nullptr, nullptr, halo, b, x, opts, result, &op);
    default: return lat_fail(ctx, LAT_ERR_ARG, "unknown preconditioner", __FILE__, __LINE__);
  }
}

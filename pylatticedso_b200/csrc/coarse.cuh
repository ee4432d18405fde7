// Two-level preconditioner of the PCG solvers: 6x6 block-Jacobi + a rigid-body-mode coarse space.
//
//   M^-1 = D^-1 + Z E^+ Z^T,      E = Z^T A Z
//
// The nodes are grouped into aggregates (boxes of cells, chosen by the host).  Each aggregate carries the six
// rigid-body modes of its nodes about a reference point c_a (its centroid, or the box centre in a sharded solve -- the
// coarse space is the same for any point):   u_i = t + w x (x_i - c_a),  theta_i = w,  with the constrained DOFs masked
// out, so Z has 6 columns per aggregate and one 6x6 block T_i per node,
//     T_i = M_i [ I  -[d_i]x ; 0  I ],   d_i = x_i - c_a,  M_i = diag(free DOF mask).
// The reference hands SuperLU's factorisation of the whole interface matrix to its PCG as the preconditioner
// (lattice_sim.py:1333-1415) and converges in a few iterations; a host factorisation cannot run inside the device
// loop, and block-Jacobi alone leaves the long-wavelength modes of a stretch-dominated lattice to the Krylov space
// (888 iterations on the 100^3 octet lattice; 139 with the coarse space).  E is small (6 n_agg, dense), its inverse is
// formed once by the host through a library factorisation, and every iteration adds, between the update kernel (which
// produces r and u = D^-1 r) and the product kernel (which needs the final u),
//     k_coarse_restrict : rc = Z^T r          one CTA per aggregate (piece), fixed-order reduction
//     k_coarse_solve    : u += Z (Einv rc)    one CTA per aggregate: 6 rows of the dense inverse, then its nodes
// (aggregates of more than COARSE_PIECE nodes: + k_coarse_gather / k_coarse_prolong, see "per-iteration kernels").
// The node passes stream r once and u twice (48 B per node each) plus a 16-byte table entry per node, the coarse product
// streams Einv once (8 n_c^2 B).  The additive form keeps M^-1 symmetric positive definite whatever the aggregates are,
// so the Chronopoulos-Gear recurrences, the stop test and the true-residual safeguard are unchanged.
// Sharded solves (lattice_solver.cu, pcg_run_dist_impl): the listed nodes are the rank's OWNED nodes, the coarse residual
// is all-reduced by k_coarse_gather_p2p (LL words into every rank's peer-memory inbox) or NCCL.
#pragma once
#include "common.cuh"

static constexpr int COARSE_BLOCK = 256;         // node passes
static constexpr int COARSE_SOLVE_BLOCK = 512;   // dense coarse product: 6 rows of Einv per CTA
static constexpr int COARSE_PIECE = 1024;        // nodes per piece of an aggregate
static constexpr int COARSE_NPT = COARSE_PIECE / COARSE_BLOCK;   // nodes per thread of a node pass, all in flight together

// One node of an aggregate: 16 bytes, so that a pass over the nodes adds 16 B to the 48 B of the node's vector entries.
// The lever arm d = x_i - c_a is kept in FP32 -- Z is DEFINED with the rounded d (Galerkin product, restriction and
// prolongation all read this table), so the coarse operator stays exactly Z^T A Z; the modes are rigid to 1e-7.
struct CoarseNode {
  float dx, dy, dz;
  uint32_t node_mask;  // bits 0-25: node (by-aggregate table) or aggregate (by-node table); bit 26+k set: DOF k is free
};
static_assert(sizeof(CoarseNode) == 16, "CoarseNode is read as one 16-byte word");
static constexpr int COARSE_NODE_BITS = 26;

#ifdef __CUDACC__

struct CoarseNodeD {   // decoded
  double dx, dy, dz;
  int32_t node;
  uint32_t mask;
};
__device__ __forceinline__ CoarseNodeD coarse_load(const CoarseNode* __restrict__ p) {
  const int4 a = __ldg(reinterpret_cast<const int4*>(p));
  CoarseNodeD c;
  c.dx = (double)__int_as_float(a.x);
  c.dy = (double)__int_as_float(a.y);
  c.dz = (double)__int_as_float(a.z);
  c.node = (int32_t)((uint32_t)a.w & ((1u << COARSE_NODE_BITS) - 1u));
  c.mask = (uint32_t)a.w >> COARSE_NODE_BITS;
  return c;
}


// T^T v for v = (f, m) masked:  (f, m + d x f)
__device__ __forceinline__ void coarse_restrict_node(const CoarseNodeD& c, const double (&v)[6], double (&acc)[6]) {
  double f[3], m[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    f[k] = (c.mask >> k) & 1u ? v[k] : 0.0;
    m[k] = (c.mask >> (3 + k)) & 1u ? v[3 + k] : 0.0;
  }
  acc[0] += f[0]; acc[1] += f[1]; acc[2] += f[2];
  acc[3] += m[0] + (c.dy * f[2] - c.dz * f[1]);
  acc[4] += m[1] + (c.dz * f[0] - c.dx * f[2]);
  acc[5] += m[2] + (c.dx * f[1] - c.dy * f[0]);
}

// T y for y = (t, w) masked:  (t + w x d, w)
__device__ __forceinline__ void coarse_prolong_node(const CoarseNodeD& c, const double (&y)[6], double (&out)[6]) {
  out[0] = y[0] + (y[4] * c.dz - y[5] * c.dy);
  out[1] = y[1] + (y[5] * c.dx - y[3] * c.dz);
  out[2] = y[2] + (y[3] * c.dy - y[4] * c.dx);
  out[3] = y[3]; out[4] = y[4]; out[5] = y[5];
#pragma unroll
  for (int k = 0; k < 6; ++k)
    if (!((c.mask >> k) & 1u)) out[k] = 0.0;
}

// block-wide fixed-order sum of 6 values; totals valid in every thread after the call
template <int BLOCK>
__device__ __forceinline__ void coarse_block_sum6(double (&v)[6], double* s_part /* [6][BLOCK/32] */, double* s_tot /* [6] */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const double w = warp_sum(v[i]);
    if (lane == 0) s_part[i * (BLOCK / 32) + wid] = w;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < BLOCK / 32; ++k) s += s_part[threadIdx.x * (BLOCK / 32) + k];
    s_tot[threadIdx.x] = s;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 6; ++i) v[i] = s_tot[i];
}

// Reference points of the aggregates (given, or the centroid of the listed nodes) and the aggregate-ordered node table.
// Any reference point spans the same coarse space (translations + rotations about ANY point); ranks of a sharded solve
// must agree on it, so they pass the box centres.
__global__ void __launch_bounds__(COARSE_BLOCK) k_coarse_setup(const int32_t* __restrict__ agg_ptr,
                                                              const int32_t* __restrict__ agg_nodes,
                                                              const double* __restrict__ x, const double* __restrict__ y,
                                                              const double* __restrict__ z, const uint8_t* __restrict__ fixed,
                                                              const double* __restrict__ centers_in, double* __restrict__ centers,
                                                              CoarseNode* __restrict__ by_agg) {
  __shared__ double s_part[6 * (COARSE_BLOCK / 32)], s_tot[6];
  const int a = blockIdx.x;
  const int lo = agg_ptr[a], hi = agg_ptr[a + 1];
  double cx, cy, cz;
  if (centers_in) {
    cx = centers_in[3 * a]; cy = centers_in[3 * a + 1]; cz = centers_in[3 * a + 2];
  } else {
    double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int k = lo + threadIdx.x; k < hi; k += COARSE_BLOCK) {
      const int n = agg_nodes[k];
      v[0] += x[n]; v[1] += y[n]; v[2] += z[n];
    }
    coarse_block_sum6<COARSE_BLOCK>(v, s_part, s_tot);
    const double inv = hi > lo ? 1.0 / (double)(hi - lo) : 0.0;
    cx = v[0] * inv; cy = v[1] * inv; cz = v[2] * inv;
  }
  if (threadIdx.x == 0) { centers[3 * a] = cx; centers[3 * a + 1] = cy; centers[3 * a + 2] = cz; }
  for (int k = lo + threadIdx.x; k < hi; k += COARSE_BLOCK) {
    const int n = agg_nodes[k];
    CoarseNode c;
    c.dx = (float)(x[n] - cx); c.dy = (float)(y[n] - cy); c.dz = (float)(z[n] - cz);
    uint32_t m = 0;
    for (int d = 0; d < 6; ++d)
      if (!fixed || !fixed[(int64_t)n * 6 + d]) m |= 1u << d;
    c.node_mask = (m << COARSE_NODE_BITS) | (uint32_t)n;
    by_agg[k] = c;
  }
}

// node-ordered table over ALL local nodes (ghosts of a sharded solve included: the Galerkin product needs the columns):
// lever arm, free-DOF mask and, in the index field, the aggregate
__global__ void __launch_bounds__(256) k_coarse_bynode(const int32_t* __restrict__ node_agg, const double* __restrict__ x,
                                                      const double* __restrict__ y, const double* __restrict__ z,
                                                      const uint8_t* __restrict__ fixed, const double* __restrict__ centers,
                                                      int64_t n_nodes, CoarseNode* __restrict__ by_node) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_nodes) return;
  const int a = node_agg[n];
  CoarseNode c;
  c.dx = (float)(x[n] - centers[3 * a]); c.dy = (float)(y[n] - centers[3 * a + 1]); c.dz = (float)(z[n] - centers[3 * a + 2]);
  uint32_t m = 0;
  for (int d = 0; d < 6; ++d)
    if (!fixed || !fixed[n * 6 + d]) m |= 1u << d;
  c.node_mask = (m << COARSE_NODE_BITS) | (uint32_t)a;
  by_node[n] = c;
}

// E += Z^T A Z: one thread per BSR block (i, j):  T_i^T A_ij T_j  added to the 6x6 block (agg_i, agg_j) of the
// dense E with FP64 reductions (set-up only; E is zeroed by the caller)
__global__ void __launch_bounds__(128) k_coarse_galerkin(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                                        const double* __restrict__ vals, int64_t n_nodes,
                                                        const CoarseNode* __restrict__ by_node, int64_t n_c,
                                                        double* __restrict__ E) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes) return;
  const CoarseNodeD ci = coarse_load(by_node + i);
  double acc[36];
  int cur = -1;
  auto flush = [&]() {
    if (cur < 0) return;
    double* e = E + ((int64_t)ci.node * 6) * n_c + (int64_t)cur * 6;
#pragma unroll
    for (int p = 0; p < 6; ++p)
#pragma unroll
      for (int q = 0; q < 6; ++q)
        if (acc[p * 6 + q] != 0.0) atomicAdd(e + (int64_t)p * n_c + q, acc[p * 6 + q]);
  };
  for (int j = rowptr[i]; j < rowptr[i + 1]; ++j) {
    const int c = colidx[j];
    const CoarseNodeD cj = coarse_load(by_node + c);
    if (cj.node != cur) {
      flush();
      cur = cj.node;
#pragma unroll
      for (int k = 0; k < 36; ++k) acc[k] = 0.0;
    }
    const double* a = vals + (int64_t)j * 36;
    // B = (M_i A M_j) T_j : columns 0..2 unchanged, column 3+k += A[:, 0:3] (e_k x d_j)
    double B[6][6];
#pragma unroll
    for (int p = 0; p < 6; ++p) {
      const bool rp = (ci.mask >> p) & 1u;
      double row[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) row[q] = (rp && ((cj.mask >> q) & 1u)) ? a[p * 6 + q] : 0.0;
      B[p][0] = row[0]; B[p][1] = row[1]; B[p][2] = row[2];
      // e_x x d = (0, -dz, dy); e_y x d = (dz, 0, -dx); e_z x d = (-dy, dx, 0)
      B[p][3] = row[3] + (-row[1] * cj.dz + row[2] * cj.dy);
      B[p][4] = row[4] + (row[0] * cj.dz - row[2] * cj.dx);
      B[p][5] = row[5] + (-row[0] * cj.dy + row[1] * cj.dx);
    }
    // G = T_i^T B : rows 0..2 unchanged, row 3+k += (d_i x B[0:3, :])_k
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      acc[0 * 6 + q] += B[0][q];
      acc[1 * 6 + q] += B[1][q];
      acc[2 * 6 + q] += B[2][q];
      acc[3 * 6 + q] += B[3][q] + (ci.dy * B[2][q] - ci.dz * B[1][q]);
      acc[4 * 6 + q] += B[4][q] + (ci.dz * B[0][q] - ci.dx * B[2][q]);
      acc[5 * 6 + q] += B[5][q] + (ci.dx * B[1][q] - ci.dy * B[0][q]);
    }
  }
  flush();
}

// ---- per-iteration kernels ---------------------------------------------------------------------------------
// Aggregates are cut into PIECES of at most COARSE_PIECE consecutive entries of the aggregate-ordered node table, so
// that the two passes over the nodes (restrict, prolong) have enough CTAs whatever the aggregate size (4 000 nodes per
// aggregate on the 100^3 octet lattice).  With one piece per aggregate the piece sums ARE rc and the prolongation is
// done by the CTA that computed the aggregate's coarse solution (2 launches); otherwise 4 launches:
//   restrict (piece) -> gather (aggregate: fixed-order sum of its pieces) -> solve (aggregate) -> prolong (piece)

// part[6 piece + p] = sum over the nodes of the piece of (T_i^T r_i)_p.  `sc` (may be null): early exit of a converged solve.
__global__ void __launch_bounds__(COARSE_BLOCK) k_coarse_restrict(const int32_t* __restrict__ piece_ptr,
                                                                 const CoarseNode* __restrict__ by_agg,
                                                                 const double* __restrict__ r, double* __restrict__ part,
                                                                 const PcgScalars* __restrict__ sc, int maxiter) {
  __shared__ double s_part[6 * (COARSE_BLOCK / 32)], s_tot[6];
  if (sc && (sc->done || sc->iters >= maxiter)) return;
  const int lo = piece_ptr[blockIdx.x], hi = piece_ptr[blockIdx.x + 1];
  double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  // a piece has at most COARSE_PIECE entries: the table words of all of this thread's nodes are requested first, then
  // all their vector entries (two memory round trips per CTA instead of two per node)
  CoarseNodeD c[COARSE_NPT];
#pragma unroll
  for (int j = 0; j < COARSE_NPT; ++j) {
    const int k = lo + threadIdx.x + j * COARSE_BLOCK;
    c[j] = coarse_load(by_agg + (k < hi ? k : lo));
    if (k >= hi) c[j].mask = 0u;                      // masked out: contributes nothing
  }
  double2 rv[COARSE_NPT][3];
#pragma unroll
  for (int j = 0; j < COARSE_NPT; ++j) {
    const double2* rp = reinterpret_cast<const double2*>(r + (int64_t)c[j].node * 6);
    rv[j][0] = rp[0]; rv[j][1] = rp[1]; rv[j][2] = rp[2];
  }
#pragma unroll
  for (int j = 0; j < COARSE_NPT; ++j) {
    const double v[6] = {rv[j][0].x, rv[j][0].y, rv[j][1].x, rv[j][1].y, rv[j][2].x, rv[j][2].y};
    coarse_restrict_node(c[j], v, acc);
  }
  coarse_block_sum6<COARSE_BLOCK>(acc, s_part, s_tot);
  if (threadIdx.x < 6) part[(int64_t)blockIdx.x * 6 + threadIdx.x] = s_tot[threadIdx.x];
}

// rc[6a + p] = sum of part[6 piece + p] over the pieces of aggregate a, in piece order (one thread per entry)
__global__ void __launch_bounds__(256) k_coarse_gather(const int32_t* __restrict__ agg_piece, const double* __restrict__ part,
                                                      double* __restrict__ rc, int n_c, const PcgScalars* __restrict__ sc,
                                                      int maxiter) {
  if (sc && (sc->done || sc->iters >= maxiter)) return;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_c) return;
  const int a = q / 6, p = q - a * 6;
  double s = 0.0;
  for (int k = agg_piece[a]; k < agg_piece[a + 1]; ++k) s += part[(int64_t)k * 6 + p];
  rc[q] = s;
}

// y = (Einv rc)[6a .. 6a+5].  PROLONG: u_i += T_i y for the nodes of aggregate a; otherwise yc[6a + p] = y_p.
template <bool PROLONG>
__global__ void __launch_bounds__(COARSE_SOLVE_BLOCK) k_coarse_solve(const int32_t* __restrict__ agg_ptr,
                                                                    const CoarseNode* __restrict__ by_agg,
                                                                    const double* __restrict__ einv, const double* __restrict__ rc,
                                                                    int64_t n_c, double* __restrict__ yc, double* __restrict__ u,
                                                                    const PcgScalars* __restrict__ sc, int maxiter) {
  __shared__ double s_part[6 * (COARSE_SOLVE_BLOCK / 32)], s_tot[6];
  if (sc && (sc->done || sc->iters >= maxiter)) return;
  const int a = blockIdx.x;
  if (agg_ptr[a + 1] == agg_ptr[a]) return;      // no node of this aggregate is prolonged here (sharded solve): y is not needed
  double y[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  const double* row = einv + (int64_t)a * 6 * n_c;
  // n_c = 6 n_agg is even and the rows start 16-byte aligned: 128-bit loads, two column pairs in flight per thread
  const int64_t n2 = n_c >> 1;
  const double2* rc2 = reinterpret_cast<const double2*>(rc);
#pragma unroll 2
  for (int64_t q = threadIdx.x; q < n2; q += COARSE_SOLVE_BLOCK) {
    const double2 v = rc2[q];
#pragma unroll
    for (int p = 0; p < 6; ++p) {
      const double2 e = __ldg(reinterpret_cast<const double2*>(row + (int64_t)p * n_c) + q);
      y[p] = fma(e.x, v.x, fma(e.y, v.y, y[p]));
    }
  }
  coarse_block_sum6<COARSE_SOLVE_BLOCK>(y, s_part, s_tot);
  if (!PROLONG) {
    if (threadIdx.x < 6) yc[(int64_t)a * 6 + threadIdx.x] = s_tot[threadIdx.x];
    return;
  }
  const int lo = agg_ptr[a], hi = agg_ptr[a + 1];
  for (int k = lo + threadIdx.x; k < hi; k += COARSE_SOLVE_BLOCK) {
    const CoarseNodeD c = coarse_load(by_agg + k);
    double d[6];
    coarse_prolong_node(c, y, d);
    double2* up = reinterpret_cast<double2*>(u + (int64_t)c.node * 6);
    double2 u0 = up[0], u1 = up[1], u2 = up[2];
    u0.x += d[0]; u0.y += d[1]; u1.x += d[2]; u1.y += d[3]; u2.x += d[4]; u2.y += d[5];
    up[0] = u0; up[1] = u1; up[2] = u2;
  }
}

// u_i += T_i yc[6a .. 6a+5] for the nodes of one piece of aggregate a
__global__ void __launch_bounds__(COARSE_BLOCK) k_coarse_prolong(const int32_t* __restrict__ piece_ptr,
                                                                const int32_t* __restrict__ piece_agg,
                                                                const CoarseNode* __restrict__ by_agg,
                                                                const double* __restrict__ yc, double* __restrict__ u,
                                                                const PcgScalars* __restrict__ sc, int maxiter) {
  if (sc && (sc->done || sc->iters >= maxiter)) return;
  const int lo = piece_ptr[blockIdx.x], hi = piece_ptr[blockIdx.x + 1];
  const double* yp = yc + (int64_t)piece_agg[blockIdx.x] * 6;
  const double y[6] = {yp[0], yp[1], yp[2], yp[3], yp[4], yp[5]};
  CoarseNodeD c[COARSE_NPT];
#pragma unroll
  for (int j = 0; j < COARSE_NPT; ++j) {
    const int k = lo + threadIdx.x + j * COARSE_BLOCK;
    c[j] = coarse_load(by_agg + (k < hi ? k : lo));
    if (k >= hi) c[j].node = -1;
  }
  double2 uv[COARSE_NPT][3];
#pragma unroll
  for (int j = 0; j < COARSE_NPT; ++j) {
    if (c[j].node < 0) continue;
    const double2* up = reinterpret_cast<const double2*>(u + (int64_t)c[j].node * 6);
    uv[j][0] = up[0]; uv[j][1] = up[1]; uv[j][2] = up[2];
  }
#pragma unroll
  for (int j = 0; j < COARSE_NPT; ++j) {
    if (c[j].node < 0) continue;
    double d[6];
    coarse_prolong_node(c[j], y, d);
    double2* up = reinterpret_cast<double2*>(u + (int64_t)c[j].node * 6);
    uv[j][0].x += d[0]; uv[j][0].y += d[1]; uv[j][1].x += d[2]; uv[j][1].y += d[3]; uv[j][2].x += d[4]; uv[j][2].y += d[5];
    up[0] = uv[j][0]; up[1] = uv[j][1]; up[2] = uv[j][2];
  }
}

// The launches of one coarse correction u += Z Einv Z^T r on stream st; returns their number.
struct CoarseLaunch {
  int n_agg = 0, n_pieces = 0;
  bool fused = false;   // every aggregate is exactly one piece
  const int32_t *agg_ptr = nullptr, *piece_ptr = nullptr, *piece_agg = nullptr, *agg_piece = nullptr;
  const CoarseNode* nodes = nullptr;
  const double* einv = nullptr;
  double *part = nullptr, *rc = nullptr, *yc = nullptr;
  int launches() const { return fused ? 2 : 4; }
  // rc = Z^T r over the listed nodes (a sharded solve all-reduces rc between the two halves).  gather = false (non-fused
  // layout only): stop after the per-piece sums; the caller's own gather kernel adds them up across ranks.
  void restrict_to_coarse(cudaStream_t st, const double* r, const PcgScalars* sc, int maxiter, bool gather = true) const {
    const int64_t n_c = 6 * (int64_t)n_agg;
    if (fused) {
      k_coarse_restrict<<<n_agg, COARSE_BLOCK, 0, st>>>(agg_ptr, nodes, r, rc, sc, maxiter);
    } else {
      k_coarse_restrict<<<n_pieces, COARSE_BLOCK, 0, st>>>(piece_ptr, nodes, r, part, sc, maxiter);
      if (gather) k_coarse_gather<<<(unsigned)((n_c + 255) / 256), 256, 0, st>>>(agg_piece, part, rc, (int)n_c, sc, maxiter);
    }
  }
  // u += Z (Einv rc) on the listed nodes
  void correct(cudaStream_t st, double* u, const PcgScalars* sc, int maxiter) const {
    const int64_t n_c = 6 * (int64_t)n_agg;
    if (fused) {
      k_coarse_solve<true><<<n_agg, COARSE_SOLVE_BLOCK, 0, st>>>(agg_ptr, nodes, einv, rc, n_c, yc, u, sc, maxiter);
    } else {
      k_coarse_solve<false><<<n_agg, COARSE_SOLVE_BLOCK, 0, st>>>(agg_ptr, nodes, einv, rc, n_c, yc, u, sc, maxiter);
      k_coarse_prolong<<<n_pieces, COARSE_BLOCK, 0, st>>>(piece_ptr, piece_agg, nodes, yc, u, sc, maxiter);
    }
  }
  void run(cudaStream_t st, const double* r, double* u, const PcgScalars* sc, int maxiter) const {
    restrict_to_coarse(st, r, sc, maxiter);
    correct(st, u, sc, maxiter);
  }
};

#endif  // __CUDACC__

// lattice_schur.cu -- batched per-cell Schur complements (dense partial Cholesky, one CTA per
// cell) and the DDM interface operator (batched S_c GEMV with gather/scatter).  sm_100a.
#include <vector>

#include "common.cuh"

static constexpr int SCHUR_BLOCK = 256;
// boundary DOFs per cell: BCC 48, Octet 84, Kelvin 144, Original2 156 (the reference's largest geometry); the per-lane
// tiles of k_ddm_matvec are sized from nb at launch (1..6 tiles of 32), the dense kernels are limited by shared memory
static constexpr int SCHUR_MAX_NB = 192;
static constexpr int GRAD_PER_THREAD = 36;   // dS entries a thread accumulates per pass (nB <= 96: one pass)

// position of local node l in the factorisation order: interior nodes first, boundary nodes last
__device__ __forceinline__ int node_pos(int l, int nn, int nbn) { return l >= nbn ? l - nbn : nn - nbn + l; }

// The six strain 12-vectors of an element (oracle.strain_vectors; simulation_base.py:141-156) --
// only needed for the rank-6 form of dK_e used by the sensitivity contraction.
__device__ void strain_vectors_dev(const double* xa, const double* xb, double* B /*[6][12]*/, double* Lout) {
  double d[3] = {xb[0] - xa[0], xb[1] - xa[1], xb[2] - xa[2]};
  const double L = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]), iL = 1.0 / L;
  double t[3] = {d[0] * iL, d[1] * iL, d[2] * iL};
  // frame rule of beam_model.py:199-216
  double e1[3] = {1, 0, 0};
  if (fabs(t[1]) < fabs(t[0])) { e1[0] = 0; e1[1] = 1; }
  const double te1 = t[0] * e1[0] + t[1] * e1[1] + t[2] * e1[2];
  double e2[3] = {e1[0], e1[1], e1[2]};
  if (fabs(t[2]) < fabs(te1)) { e2[0] = 0; e2[1] = 0; e2[2] = 1; }
  double a1[3] = {t[1] * e2[2] - t[2] * e2[1], t[2] * e2[0] - t[0] * e2[2], t[0] * e2[1] - t[1] * e2[0]};
  double n1 = 1.0 / sqrt(a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2]);
  a1[0] *= n1; a1[1] *= n1; a1[2] *= n1;
  double a2[3] = {t[1] * a1[2] - t[2] * a1[1], t[2] * a1[0] - t[0] * a1[2], t[0] * a1[1] - t[1] * a1[0]};
  double n2 = 1.0 / sqrt(a2[0] * a2[0] + a2[1] * a2[1] + a2[2] * a2[2]);
  a2[0] *= n2; a2[1] *= n2; a2[2] *= n2;
  for (int i = 0; i < 72; ++i) B[i] = 0.0;
  for (int k = 0; k < 3; ++k) {
    B[0 * 12 + k] = -t[k] * iL;  B[0 * 12 + 6 + k] = t[k] * iL;
    B[1 * 12 + k] = -a1[k] * iL; B[1 * 12 + 6 + k] = a1[k] * iL; B[1 * 12 + 3 + k] = -0.5 * a2[k]; B[1 * 12 + 9 + k] = -0.5 * a2[k];
    B[2 * 12 + k] = -a2[k] * iL; B[2 * 12 + 6 + k] = a2[k] * iL; B[2 * 12 + 3 + k] = 0.5 * a1[k];  B[2 * 12 + 9 + k] = 0.5 * a1[k];
    B[3 * 12 + 3 + k] = -t[k] * iL;  B[3 * 12 + 9 + k] = t[k] * iL;
    B[4 * 12 + 3 + k] = -a1[k] * iL; B[4 * 12 + 9 + k] = a1[k] * iL;
    B[5 * 12 + 3 + k] = -a2[k] * iL; B[5 * 12 + 9 + k] = a2[k] * iL;
  }
  *Lout = L;
}

// ---------------------------------------------------------------------------------------------------
// Strut pre-pass: exact static condensation of straight element chains
// ---------------------------------------------------------------------------------------------------
// The reference meshes every strut with ~18 elements (gmsh rule h = 0.05 cell size), i.e. 822 of the 870 DOFs
// of a BCC cell sit on degree-2 nodes inside straight struts.  In the frame of a straight strut of circular
// section the stiffness splits into an axial spring series (sum L/ES), a torsion spring series (sum L/GJ) and
// ONE planar Timoshenko beam (both bending planes are identical) on (W, Phi = i*Theta):
//     M_WW(re,ce) = sA GS/L    M_WPhi(re,ce) = (re == 0 ? -GS/2 : GS/2)    M_PhiPhi(re,ce) = GS L/4 + sA EI/L
// whose chain is condensed with 2x2 pivots by ONE THREAD per (cell, strut): O(m) scalar work instead of m - 1
// dense 6x6 pivots.  Back in 3-D the condensed 12x12 "super-element" is
//     ww = M_WW (I - tt) + k_ax tt,  w-th = M_WPhi [t]x,  th-w = -M_PhiW [t]x,  th-th = M_PhiPhi (I - tt) + k_tor tt
// and the cell keeps its joints only (BCC: 9 nodes for ANY subdivision).  Checker: oracle.condensed_strut /
// schur_via_chain_condensation, equal to the reference's 30 stored Schur matrices to 1.3e-12.
struct SupCoef {
  double tx, ty, tz, kax, ktor;
  double m[10];   // symmetric 4x4 on (W_A, Phi_A, W_B, Phi_B), upper triangle row-major
  double pad;
};
__host__ __device__ __forceinline__ int sym4(int r, int c) {   // index into m[] of entry (r, c)
  const int i = r < c ? r : c, j = r < c ? c : r;
  return i * 4 - (i * (i - 1)) / 2 + (j - i);
}

__global__ void __launch_bounds__(128) k_chain_condense(
    const double* __restrict__ xyz, const int32_t* __restrict__ len0, const int32_t* __restrict__ len1,
    const double* __restrict__ rad, int64_t n_cells, int nn, int ne, const int32_t* __restrict__ chain_ptr,
    const int32_t* __restrict__ chain_elem, const int32_t* __restrict__ chain_flip, int n_chains, double young,
    double nu, double kappa, SupCoef* __restrict__ sup) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_cells * n_chains) return;
  const int64_t c = w / n_chains;
  const int ch = (int)(w - c * n_chains);
  const double* cx = xyz + c * (int64_t)nn * 3;
  const double* cr = rad + c * (int64_t)ne;
  const double PI = 3.14159265358979323846, G = young / (2.0 * (1.0 + nu));
  double flex_ax = 0.0, flex_tor = 0.0;
  double tx = 0.0, ty = 0.0, tz = 0.0;
  // condensed planar beam so far, blocks on (A, k): aa, ak, kk (2x2 each; aa, kk symmetric)
  double aa[2][2], ak[2][2], kk[2][2];
  for (int q = chain_ptr[ch]; q < chain_ptr[ch + 1]; ++q) {
    const int e = chain_elem[q];
    const bool fl = chain_flip[q] != 0;
    const int a = fl ? len1[e] : len0[e], b = fl ? len0[e] : len1[e];
    const double dx = cx[b * 3] - cx[a * 3], dy = cx[b * 3 + 1] - cx[a * 3 + 1], dz = cx[b * 3 + 2] - cx[a * 3 + 2];
    const double L = sqrt(dx * dx + dy * dy + dz * dz), iL = 1.0 / L;
    const double r = cr[e];
    const double S = PI * r * r, I = PI * r * r * r * r * 0.25;
    const double ES = young * S, GS = G * kappa * S, EI = young * I, GJ = G * 2.0 * I;
    flex_ax += L / ES;
    flex_tor += L / GJ;
    // element planar beam on (k, n): diagonal blocks [[GS/L, -+GS/2], [-+GS/2, GS L/4 + EI/L]], coupling below
    const double g1 = GS * iL, g2 = 0.5 * GS, dp = 0.25 * GS * L + EI * iL, dm = 0.25 * GS * L - EI * iL;
    if (q == chain_ptr[ch]) {
      tx = dx * iL; ty = dy * iL; tz = dz * iL;
      aa[0][0] = g1; aa[0][1] = -g2; aa[1][0] = -g2; aa[1][1] = dp;
      ak[0][0] = -g1; ak[0][1] = -g2; ak[1][0] = g2; ak[1][1] = dm;
      kk[0][0] = g1; kk[0][1] = g2; kk[1][0] = g2; kk[1][1] = dp;
      continue;
    }
    // pivot P = kk + E_kk with E_kk = [[g1, -g2], [-g2, dp]];  E_kn = [[-g1, -g2], [g2, dm]];  E_nn = [[g1, g2], [g2, dp]]
    const double p00 = kk[0][0] + g1, p01 = kk[0][1] - g2, p11 = kk[1][1] + dp;
    const double idet = 1.0 / (p00 * p11 - p01 * p01);
    const double i00 = p11 * idet, i01 = -p01 * idet, i11 = p00 * idet;     // P^-1 (symmetric)
    const double ekn[2][2] = {{-g1, -g2}, {g2, dm}};
    // X = ak P^-1 (2x2),  Y = E_kn^T P^-1 ... work with explicit products
    double x[2][2], y[2][2];   // x = ak * Pinv,  y = ekn^T * Pinv
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      x[i][0] = ak[i][0] * i00 + ak[i][1] * i01;
      x[i][1] = ak[i][0] * i01 + ak[i][1] * i11;
      y[i][0] = ekn[0][i] * i00 + ekn[1][i] * i01;
      y[i][1] = ekn[0][i] * i01 + ekn[1][i] * i11;
    }
    double naa[2][2], nan_[2][2], nnn[2][2];
    const double enn[2][2] = {{g1, g2}, {g2, dp}};
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        naa[i][j] = aa[i][j] - (x[i][0] * ak[j][0] + x[i][1] * ak[j][1]);        // aa - ak P^-1 ak^T
        nan_[i][j] = -(x[i][0] * ekn[0][j] + x[i][1] * ekn[1][j]);               // -ak P^-1 E_kn
        nnn[i][j] = enn[i][j] - (y[i][0] * ekn[0][j] + y[i][1] * ekn[1][j]);     // E_nn - E_kn^T P^-1 E_kn
      }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) { aa[i][j] = naa[i][j]; ak[i][j] = nan_[i][j]; kk[i][j] = nnn[i][j]; }
  }
  SupCoef o;
  o.tx = tx; o.ty = ty; o.tz = tz;
  o.kax = 1.0 / flex_ax;
  o.ktor = 1.0 / flex_tor;
  o.m[sym4(0, 0)] = aa[0][0]; o.m[sym4(0, 1)] = 0.5 * (aa[0][1] + aa[1][0]); o.m[sym4(1, 1)] = aa[1][1];
  o.m[sym4(0, 2)] = ak[0][0]; o.m[sym4(0, 3)] = ak[0][1]; o.m[sym4(1, 2)] = ak[1][0]; o.m[sym4(1, 3)] = ak[1][1];
  o.m[sym4(2, 2)] = kk[0][0]; o.m[sym4(2, 3)] = 0.5 * (kk[0][1] + kk[1][0]); o.m[sym4(3, 3)] = kk[1][1];
  o.pad = 0.0;
  sup[w] = o;
}

// entry (i, j) of the 6x6 block (row end re, column end ce) of a super-element
__device__ __forceinline__ double sup_block_entry(const SupCoef& s, int re, int ce, int i, int j) {
  const bool rw = i < 3, cw = j < 3;
  const int a = rw ? i : i - 3, b = cw ? j : j - 3;
  const double ta = (a == 0) ? s.tx : (a == 1 ? s.ty : s.tz);
  const double tb = (b == 0) ? s.tx : (b == 1 ? s.ty : s.tz);
  const double d = (a == b) ? 1.0 : 0.0;
  const double mv = s.m[sym4(2 * re + (rw ? 0 : 1), 2 * ce + (cw ? 0 : 1))];
  const double sA = (re == ce) ? 1.0 : -1.0;
  if (rw && cw) return mv * (d - ta * tb) + sA * s.kax * ta * tb;
  if (!rw && !cw) return mv * (d - ta * tb) + sA * s.ktor * ta * tb;
  double sk = 0.0;
  if (a != b) {
    const int k = 3 - a - b;
    const double tk = (k == 0) ? s.tx : (k == 1 ? s.ty : s.tz);
    sk = ((b - a + 3) % 3 == 1) ? -tk : tk;  // [t]x entry (a, b)
  }
  return rw ? mv * sk : -mv * sk;
}

// ---------------------------------------------------------------------------------------------------
// Joint-only GLOBAL assembly: the same super-elements in the BSR matrix of the whole lattice
// ---------------------------------------------------------------------------------------------------
// With no load and no constraint on strut-interior nodes (the reference applies both to lattice points only),
// static condensation of every strut is exact: the joint-only system K_J u_J = f_J has the joint displacements
// and reactions of the full system with 5x (2 elements per strut) to ~70x (the reference's 18) fewer DOFs and a
// far better condition number.  One thread per block row, as k_assemble_rows (256-bit stores, deterministic).
__device__ __forceinline__ void sup_block_accum(const SupCoef& s, int re, int ce, double (&acc)[36]) {
  double mWW, mWP, mPW, mPP;   // constant indices only (a runtime index into s.m[] would go through local memory)
  if (re == 0 && ce == 0) { mWW = s.m[0]; mWP = s.m[1]; mPW = s.m[1]; mPP = s.m[4]; }
  else if (re == 0 && ce == 1) { mWW = s.m[2]; mWP = s.m[3]; mPW = s.m[5]; mPP = s.m[6]; }
  else if (re == 1 && ce == 0) { mWW = s.m[2]; mWP = s.m[5]; mPW = s.m[3]; mPP = s.m[6]; }
  else { mWW = s.m[7]; mWP = s.m[8]; mPW = s.m[8]; mPP = s.m[9]; }
  const double sA = (re == ce) ? 1.0 : -1.0;
  const double t[3] = {s.tx, s.ty, s.tz};
  const double ax = sA * s.kax, tor = sA * s.ktor;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const double tt = t[a] * t[b];
      const double d = (a == b) ? 1.0 : 0.0;
      double sk = 0.0;   // [t]x entry (a, b)
      if (a == 0 && b == 1) sk = -t[2];
      if (a == 0 && b == 2) sk = t[1];
      if (a == 1 && b == 0) sk = t[2];
      if (a == 1 && b == 2) sk = -t[0];
      if (a == 2 && b == 0) sk = -t[1];
      if (a == 2 && b == 1) sk = t[0];
      acc[a * 6 + b] += mWW * (d - tt) + ax * tt;
      acc[a * 6 + 3 + b] += mWP * sk;
      acc[(a + 3) * 6 + b] += -mPW * sk;
      acc[(a + 3) * 6 + 3 + b] += mPP * (d - tt) + tor * tt;
    }
  }
}
__device__ __forceinline__ void store_block_sup(double* __restrict__ dst, const double (&q)[36], bool accumulate) {
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    double a = q[4 * k], b = q[4 * k + 1], c = q[4 * k + 2], d = q[4 * k + 3];
    if (accumulate) {
      double oa, ob, oc, od;
      asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(oa), "=d"(ob), "=d"(oc), "=d"(od) : "l"(dst + 4 * k) : "memory");
      a += oa; b += ob; c += oc; d += od;
    }
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * k), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
  }
}
__global__ void __launch_bounds__(128) k_assemble_rows_sup(
    const SupCoef* __restrict__ sup, const int32_t* __restrict__ adjptr, const int32_t* __restrict__ adj_other,
    const int32_t* __restrict__ adj_el, const int32_t* __restrict__ rowptr, int64_t n_nodes, double* __restrict__ vals) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_nodes) return;
  const int lo = adjptr[n], hi = adjptr[n + 1];
  if (hi == lo) return;
  int w = rowptr[n];
  int wdiag = -1;
  double dacc[36];
#pragma unroll
  for (int k = 0; k < 36; ++k) dacc[k] = 0.0;
  int prev = -1;
  for (int i = lo; i < hi; ++i) {
    const int other = adj_other[i];
    const int ee = adj_el[i];
    const SupCoef s = sup[ee >> 1];
    const int end = ee & 1;
    sup_block_accum(s, end, end, dacc);
    double q[36];
#pragma unroll
    for (int k = 0; k < 36; ++k) q[k] = 0.0;
    sup_block_accum(s, end, end ^ 1, q);
    const bool dup = (other == prev);
    if (!dup) {
      if (wdiag < 0 && other > (int)n) { wdiag = w; ++w; }
      ++w;
    }
    store_block_sup(vals + (int64_t)(w - 1) * 36, q, dup);
    prev = other;
  }
  if (wdiag < 0) wdiag = w;
  store_block_sup(vals + (int64_t)wdiag * 36, dacc, false);
}

// xyz: [n_nodes_full][3] (AoS), len0/len1/rad: the FULL (subdivided) mesh; chains as in lat_schur_batch_chains with
// one "cell".  The resident pattern must be the one of the JOINT mesh (chain_a, chain_b over n_joints nodes).
extern "C" int lat_assemble_bsr_struts(lat_ctx* ctx, const double* xyz, const int32_t* len0, const int32_t* len1,
                                       const double* rad, int64_t n_nodes_full, int64_t n_elem,
                                       const int32_t* chain_ptr, const int32_t* chain_elem, const int32_t* chain_flip,
                                       int64_t n_chains, int64_t n_joints, double young, double nu, double kappa,
                                       double* vals) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, xyz && len0 && len1 && rad && chain_ptr && chain_elem && chain_flip && vals);
  LAT_CHECK_ARG(ctx, n_nodes_full > 0 && n_elem > 0 && n_chains > 0 && n_joints > 0 && n_joints <= n_nodes_full);
  LAT_CHECK_ARG(ctx, n_nodes_full < ((int64_t)1 << 30) && n_elem < ((int64_t)1 << 30));
  LAT_CHECK_ARG(ctx, (reinterpret_cast<uintptr_t>(vals) & 31) == 0);
  if (ctx->pat_nnzb < 0 || ctx->pat_nelem != n_chains || ctx->pat_nnodes != n_joints)
    return lat_fail(ctx, LAT_ERR_STATE, "resident pattern is not the joint mesh (chain ends over n_joints nodes): call lat_bsr_pattern_build", __FILE__, __LINE__);
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  SupCoef* sup = lat_buf<SupCoef>(ctx, "strut_sup", (size_t)n_chains);
  if (!sup) return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
  LAT_LAUNCH(ctx, k_chain_condense, (unsigned)ceil_div(n_chains, 128), 128, 0, xyz, len0, len1, rad, (int64_t)1,
             (int)n_nodes_full, (int)n_elem, chain_ptr, chain_elem, chain_flip, (int)n_chains, young, nu, kappa, sup);
  LAT_LAUNCH(ctx, k_assemble_rows_sup, (unsigned)ceil_div(n_joints, 128), 128, 0, sup,
             (const int32_t*)ctx->bufs["pat_adjptr"].p, (const int32_t*)ctx->bufs["pat_adj_other"].p,
             (const int32_t*)ctx->bufs["pat_adj_el"].p, (const int32_t*)ctx->bufs["pat_rowptr"].p, n_joints, vals);
  return LAT_OK;
}

// Back-substitution of the joint-only solve: the displacements of the strut-interior nodes from the two joint
// displacements of their strut.  One thread per strut: axial and torsional components follow the spring series
// (linear in the accumulated flexibility), the two bending planes share one block-tridiagonal 2x2 system that is
// swept forward exactly as in k_chain_condense (pivot inverses kept in local memory) and solved backwards for
// both planes.  Planes: (W, Phi) = (w.a1, -th.a2) and (w.a2, th.a1) for any orthonormal a1, a2 = t x a1.
static constexpr int STRUT_MAX_SEG = 64;
struct PlanarElem { double g1, g2, dp, dm, fax, ftor; };
__device__ __forceinline__ PlanarElem planar_elem(const double* __restrict__ xyz, const int32_t* __restrict__ len0,
                                                  const int32_t* __restrict__ len1, const double* __restrict__ rad,
                                                  int e, double young, double G, double kappa) {
  const int a = len0[e], b = len1[e];
  const double dx = xyz[b * 3] - xyz[a * 3], dy = xyz[b * 3 + 1] - xyz[a * 3 + 1], dz = xyz[b * 3 + 2] - xyz[a * 3 + 2];
  const double L = sqrt(dx * dx + dy * dy + dz * dz), iL = 1.0 / L;
  const double r = rad[e];
  const double PI = 3.14159265358979323846;
  const double S = PI * r * r, I = PI * r * r * r * r * 0.25;
  const double ES = young * S, GS = G * kappa * S, EI = young * I, GJ = G * 2.0 * I;
  PlanarElem p;
  p.g1 = GS * iL; p.g2 = 0.5 * GS; p.dp = 0.25 * GS * L + EI * iL; p.dm = 0.25 * GS * L - EI * iL;
  p.fax = L / ES; p.ftor = L / GJ;
  return p;
}
__global__ void __launch_bounds__(64) k_strut_recover(
    const double* __restrict__ xyz, const int32_t* __restrict__ len0, const int32_t* __restrict__ len1,
    const double* __restrict__ rad, const int32_t* __restrict__ chain_ptr, const int32_t* __restrict__ chain_elem,
    const int32_t* __restrict__ chain_flip, const int32_t* __restrict__ chain_a, const int32_t* __restrict__ chain_b,
    int64_t n_chains, double young, double nu, double kappa, const double* __restrict__ uj, double* __restrict__ ufull) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_chains) return;
  const int q0 = chain_ptr[s], m = chain_ptr[s + 1] - q0;
  if (m < 2 || m > STRUT_MAX_SEG) return;
  const double G = young / (2.0 * (1.0 + nu));
  const int A = chain_a[s], B = chain_b[s];
  // strut direction (walking order) and a frame
  double t[3];
  {
    const int e = chain_elem[q0];
    const bool fl = chain_flip[q0] != 0;
    const int a = fl ? len1[e] : len0[e], b = fl ? len0[e] : len1[e];
    const double dx = xyz[b * 3] - xyz[a * 3], dy = xyz[b * 3 + 1] - xyz[a * 3 + 1], dz = xyz[b * 3 + 2] - xyz[a * 3 + 2];
    const double iL = rsqrt(dx * dx + dy * dy + dz * dz);
    t[0] = dx * iL; t[1] = dy * iL; t[2] = dz * iL;
    const double n2 = 1.0 / sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);   // rsqrt is approximate: renormalise
    t[0] *= n2; t[1] *= n2; t[2] *= n2;
  }
  double a1[3], a2[3];
  if (fabs(t[0]) < 0.9) { a1[0] = 0.0; a1[1] = t[2]; a1[2] = -t[1]; }      // t x ex
  else { a1[0] = -t[2]; a1[1] = 0.0; a1[2] = t[0]; }                        // t x ey
  {
    const double n1 = 1.0 / sqrt(a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2]);
    a1[0] *= n1; a1[1] *= n1; a1[2] *= n1;
  }
  a2[0] = t[1] * a1[2] - t[2] * a1[1]; a2[1] = t[2] * a1[0] - t[0] * a1[2]; a2[2] = t[0] * a1[1] - t[1] * a1[0];
  double uA[6], uB[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) { uA[k] = uj[(int64_t)A * 6 + k]; uB[k] = uj[(int64_t)B * 6 + k]; }
  auto dot3 = [](const double* p, const double* q) { return p[0] * q[0] + p[1] * q[1] + p[2] * q[2]; };
  // planar unknowns of the two planes at the joints: x = (W, Phi)
  const double xA[2][2] = {{dot3(uA, a1), -dot3(uA + 3, a2)}, {dot3(uA, a2), dot3(uA + 3, a1)}};
  double xn[2][2] = {{dot3(uB, a1), -dot3(uB + 3, a2)}, {dot3(uB, a2), dot3(uB + 3, a1)}};   // x_{k+1}, starts at B
  const double wAt = dot3(uA, t), wBt = dot3(uB, t), tAt = dot3(uA + 3, t), tBt = dot3(uB + 3, t);
  // forward sweep
  double pinv[STRUT_MAX_SEG][3], akk[STRUT_MAX_SEG][4];
  double fax_tot = 0.0, ftor_tot = 0.0;
  double kk[2][2], ak[2][2];
  for (int k = 0; k < m; ++k) {
    const PlanarElem p = planar_elem(xyz, len0, len1, rad, chain_elem[q0 + k], young, G, kappa);
    fax_tot += p.fax; ftor_tot += p.ftor;
    if (k == 0) {
      kk[0][0] = p.g1; kk[0][1] = p.g2; kk[1][0] = p.g2; kk[1][1] = p.dp;
      ak[0][0] = -p.g1; ak[0][1] = -p.g2; ak[1][0] = p.g2; ak[1][1] = p.dm;
      continue;
    }
    const double p00 = kk[0][0] + p.g1, p01 = kk[0][1] - p.g2, p11 = kk[1][1] + p.dp;
    const double idet = 1.0 / (p00 * p11 - p01 * p01);
    const double i00 = p11 * idet, i01 = -p01 * idet, i11 = p00 * idet;
    pinv[k][0] = i00; pinv[k][1] = i01; pinv[k][2] = i11;
    akk[k][0] = ak[0][0]; akk[k][1] = ak[0][1]; akk[k][2] = ak[1][0]; akk[k][3] = ak[1][1];
    const double ekn[2][2] = {{-p.g1, -p.g2}, {p.g2, p.dm}};
    const double enn[2][2] = {{p.g1, p.g2}, {p.g2, p.dp}};
    double x[2][2], y[2][2], nak[2][2], nkk[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      x[i][0] = ak[i][0] * i00 + ak[i][1] * i01;
      x[i][1] = ak[i][0] * i01 + ak[i][1] * i11;
      y[i][0] = ekn[0][i] * i00 + ekn[1][i] * i01;
      y[i][1] = ekn[0][i] * i01 + ekn[1][i] * i11;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        nak[i][j] = -(x[i][0] * ekn[0][j] + x[i][1] * ekn[1][j]);
        nkk[i][j] = enn[i][j] - (y[i][0] * ekn[0][j] + y[i][1] * ekn[1][j]);
      }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) { ak[i][j] = nak[i][j]; kk[i][j] = nkk[i][j]; }
  }
  // backward: x_k = -P_k^-1 (ak_k^T x_A + E_kn(element k) x_{k+1}),  k = m-1 .. 1
  double fax_suffix = 0.0, ftor_suffix = 0.0;   // flexibility of the elements k .. m-1
  for (int k = m - 1; k >= 1; --k) {
    const PlanarElem p = planar_elem(xyz, len0, len1, rad, chain_elem[q0 + k], young, G, kappa);
    fax_suffix += p.fax; ftor_suffix += p.ftor;
    const double i00 = pinv[k][0], i01 = pinv[k][1], i11 = pinv[k][2];
    double xk[2][2];
#pragma unroll
    for (int pl = 0; pl < 2; ++pl) {
      // rhs = ak_k^T x_A + E_kn x_{k+1}
      const double r0 = akk[k][0] * xA[pl][0] + akk[k][2] * xA[pl][1] + (-p.g1) * xn[pl][0] + (-p.g2) * xn[pl][1];
      const double r1 = akk[k][1] * xA[pl][0] + akk[k][3] * xA[pl][1] + p.g2 * xn[pl][0] + p.dm * xn[pl][1];
      xk[pl][0] = -(i00 * r0 + i01 * r1);
      xk[pl][1] = -(i01 * r0 + i11 * r1);
    }
    const double cax = 1.0 - fax_suffix / fax_tot, ctor = 1.0 - ftor_suffix / ftor_tot;   // share of elements 0 .. k-1
    const double wt = wAt + (wBt - wAt) * cax, tt = tAt + (tBt - tAt) * ctor;
    // node between element k-1 and k (walking order): the end node of element k-1
    const int ep = chain_elem[q0 + k - 1];
    const int node = chain_flip[q0 + k - 1] ? len0[ep] : len1[ep];
    double* o = ufull + (int64_t)node * 6;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      o[c] = wt * t[c] + xk[0][0] * a1[c] + xk[1][0] * a2[c];
      o[3 + c] = tt * t[c] + xk[1][1] * a1[c] - xk[0][1] * a2[c];
    }
#pragma unroll
    for (int pl = 0; pl < 2; ++pl) { xn[pl][0] = xk[pl][0]; xn[pl][1] = xk[pl][1]; }
  }
}

// u_full[0 .. 6 n_joints) must already hold the joint displacements (joints are the first nodes of the full mesh);
// this fills the strut-interior nodes.  max_chain_len <= 64.
extern "C" int lat_strut_recover(lat_ctx* ctx, const double* xyz, const int32_t* len0, const int32_t* len1,
                                 const double* rad, const int32_t* chain_ptr, const int32_t* chain_elem,
                                 const int32_t* chain_flip, const int32_t* chain_a, const int32_t* chain_b,
                                 int64_t n_chains, int32_t max_chain_len, double young, double nu, double kappa,
                                 const double* u_joints, double* u_full) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, xyz && len0 && len1 && rad && chain_ptr && chain_elem && chain_flip && chain_a && chain_b && u_joints && u_full);
  LAT_CHECK_ARG(ctx, n_chains > 0);
  if (max_chain_len > STRUT_MAX_SEG)
    return lat_fail(ctx, LAT_ERR_UNSUPPORTED, "lat_strut_recover: more than 64 elements per strut", __FILE__, __LINE__);
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  LAT_LAUNCH(ctx, k_strut_recover, (unsigned)ceil_div(n_chains, 64), 64, 0, xyz, len0, len1, rad, chain_ptr, chain_elem,
             chain_flip, chain_a, chain_b, n_chains, young, nu, kappa, u_joints, u_full);
  return LAT_OK;
}

// One CTA per cell (grid-stride over cells).  A = dense cell stiffness in factorisation order
// (interior DOFs first), lower triangle used.  Partial right-looking Cholesky over the nI interior
// pivots; the rows of column k that are exactly zero are skipped (the cell graph is a set of strut
// chains joined at a few nodes, so most of the column is structurally zero), which keeps the cost
// proportional to the fill, not to n^3.  The trailing nB x nB block is then the Schur complement.
// SUPER = true: the "elements" are the condensed struts of k_chain_condense (sup[n_cells][ne]) on the joint-only
// cell; xyz / rad are not read and there are no sensitivities in this mode.
template <bool SMEM, bool SUPER = false>
__global__ void __launch_bounds__(SCHUR_BLOCK) k_schur_dense(
    const double* __restrict__ xyz, const int32_t* __restrict__ len0, const int32_t* __restrict__ len1,
    const double* __restrict__ rad, int64_t n_cells, int nn, int nbn, int ne, double young, double nu, double kappa,
    double* __restrict__ S, double* __restrict__ ws, const int32_t* __restrict__ elem_group,
    const double* __restrict__ drad_chain, int n_grad, double* __restrict__ dS,
    const SupCoef* __restrict__ sup = nullptr) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const int n = 6 * nn, nB = 6 * nbn, nI = n - nB;
  const int ld = n | 1;  // odd leading dimension: column walks are bank-conflict free
  // shared layout
  size_t off = 0;
  double* A = SMEM ? reinterpret_cast<double*>(sm_raw) : (ws + (size_t)blockIdx.x * n * ld);
  if (SMEM) off += (size_t)n * ld * sizeof(double);
  ElemCoef* s_coef = reinterpret_cast<ElemCoef*>(sm_raw + off);
  SupCoef* s_sup = reinterpret_cast<SupCoef*>(sm_raw + off);
  off += (size_t)ne * (SUPER ? sizeof(SupCoef) : sizeof(ElemCoef));
  double* s_diag = reinterpret_cast<double*>(sm_raw + off); off += (size_t)n * sizeof(double);
  double* s_v = reinterpret_cast<double*>(sm_raw + off); off += (size_t)6 * SCHUR_MAX_NB * sizeof(double);
  double* s_B = reinterpret_cast<double*>(sm_raw + off); off += 80 * sizeof(double);  // 72 strain entries + L + weights
  int32_t* s_list = reinterpret_cast<int32_t*>(sm_raw + off); off += (size_t)n * sizeof(int32_t);
  int32_t* s_e0 = reinterpret_cast<int32_t*>(sm_raw + off); off += (size_t)ne * sizeof(int32_t);
  int32_t* s_e1 = reinterpret_cast<int32_t*>(sm_raw + off); off += (size_t)ne * sizeof(int32_t);
  __shared__ int s_cnt;
  __shared__ int s_bad;
  const int tid = threadIdx.x;

  for (int e = tid; e < ne; e += SCHUR_BLOCK) { s_e0[e] = len0[e]; s_e1[e] = len1[e]; }

  for (int64_t c = blockIdx.x; c < n_cells; c += gridDim.x) {
    const double* cx = xyz + c * (int64_t)nn * 3;
    const double* cr = rad + c * (int64_t)ne;
    __syncthreads();
    if (tid == 0) { s_cnt = 0; s_bad = 0; }
    for (int e = tid; e < ne; e += SCHUR_BLOCK) {
      if (SUPER) {
        s_sup[e] = sup[c * (int64_t)ne + e];
      } else {
        const int a = s_e0[e], b = s_e1[e];
        s_coef[e] = elem_coef(cx[a * 3], cx[a * 3 + 1], cx[a * 3 + 2], cx[b * 3], cx[b * 3 + 1], cx[b * 3 + 2], cr[e],
                              young, nu, kappa, false);
      }
    }
    for (int i = tid; i < n * ld; i += SCHUR_BLOCK) A[i] = 0.0;
    __syncthreads();
    // ---- assembly: work item = (row node a, entry k of a 6x6 block); fixed element order -> deterministic
    for (int w = tid; w < nn * 36; w += SCHUR_BLOCK) {
      const int a = w / 36, k = w - a * 36, i = k / 6, j = k - i * 6;
      const int ra = node_pos(a, nn, nbn) * 6 + i;
      double diag = 0.0;
      for (int e = 0; e < ne; ++e) {
        const int e0 = s_e0[e], e1 = s_e1[e];
        if (e0 == a) {
          diag += SUPER ? sup_block_entry(s_sup[e], 0, 0, i, j) : elem_block_entry(s_coef[e], 0, 0, i, j);
          A[ra * ld + node_pos(e1, nn, nbn) * 6 + j] +=
              SUPER ? sup_block_entry(s_sup[e], 0, 1, i, j) : elem_block_entry(s_coef[e], 0, 1, i, j);
        } else if (e1 == a) {
          diag += SUPER ? sup_block_entry(s_sup[e], 1, 1, i, j) : elem_block_entry(s_coef[e], 1, 1, i, j);
          A[ra * ld + node_pos(e0, nn, nbn) * 6 + j] +=
              SUPER ? sup_block_entry(s_sup[e], 1, 0, i, j) : elem_block_entry(s_coef[e], 1, 0, i, j);
        }
      }
      A[ra * ld + node_pos(a, nn, nbn) * 6 + j] += diag;
    }
    // ---- partial Cholesky over the interior pivots
    for (int k = 0; k < nI; ++k) {
      __syncthreads();
      const double p = A[k * ld + k];
      if (!(p > 0.0)) { if (tid == 0) s_bad = 1; }
      const double inv = rsqrt(p);
      const double invr = inv * (1.5 - 0.5 * p * inv * inv);  // one Newton step on rsqrt -> full FP64 accuracy
      for (int i = k + 1 + tid; i < n; i += SCHUR_BLOCK) {
        const double v = A[i * ld + k];
        if (v != 0.0) {
          A[i * ld + k] = v * invr;
          s_list[atomicAdd(&s_cnt, 1)] = i;
        }
      }
      if (tid == 0) s_diag[k] = p * invr;  // sqrt(p)
      __syncthreads();
      const int m = s_cnt;
      for (int q = tid; q < m * m; q += SCHUR_BLOCK) {
        const int ii = q / m, jj = q - ii * m;
        const int i = s_list[ii], j = s_list[jj];
        if (i >= j) A[i * ld + j] -= A[i * ld + k] * A[j * ld + k];
      }
      __syncthreads();
      if (tid == 0) s_cnt = 0;
    }
    __syncthreads();
    const bool bad = s_bad != 0;
    // ---- S = trailing block (lower triangle mirrored), row-major [nB][nB]
    double* Sc = S + c * (int64_t)nB * nB;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    for (int q = tid; q < nB * nB; q += SCHUR_BLOCK) {
      const int bi = q / nB, bj = q - bi * nB;
      const int hi = bi > bj ? bi : bj, lo = bi > bj ? bj : bi;
      Sc[q] = bad ? qnan : A[(nI + hi) * ld + nI + lo];
    }
    if (SUPER || dS == nullptr || n_grad <= 0) continue;
    // ---- sensitivities: dS_g = E^T dK_g E with E = [-X; I], X = K_II^-1 K_IB.
    // rows nI.. of A hold Y^T = K_BI L^-T; back-substitute in place to X^T = Y^T L^-1.
    __syncthreads();
    for (int b = tid; b < nB; b += SCHUR_BLOCK) {
      double* row = A + (size_t)(nI + b) * ld;
      for (int k = nI - 1; k >= 0; --k) {
        double v = row[k];
        for (int i = k + 1; i < nI; ++i) v -= A[i * ld + k] * row[i];
        row[k] = v / s_diag[k];
      }
    }
    __syncthreads();
    for (int gsel = 0; gsel < n_grad; ++gsel)
    for (int q0 = 0; q0 < nB * nB; q0 += GRAD_PER_THREAD * SCHUR_BLOCK) {   // passes over the entries of dS (one for nB <= 96)
      double acc[GRAD_PER_THREAD];
#pragma unroll
      for (int u = 0; u < GRAD_PER_THREAD; ++u) acc[u] = 0.0;
      for (int e = 0; e < ne; ++e) {
        if (elem_group[e] != gsel) continue;  // uniform across the CTA
        const int a = s_e0[e], b = s_e1[e];
        if (tid == 0) {
          double L;
          strain_vectors_dev(cx + a * 3, cx + b * 3, s_B, &L);
          const double r = cr[e];
          const double PI = 3.14159265358979323846, G = young / (2.0 * (1.0 + nu));
          const double dSr = 2.0 * PI * r, dIr = PI * r * r * r, ch = drad_chain ? drad_chain[e] : 1.0;
          const double wES = young * dSr, wGS = G * kappa * dSr, wGJ = G * 2.0 * dIr, wEI = young * dIr;
          s_B[72] = ch * L * wES; s_B[73] = ch * L * wGS; s_B[74] = ch * L * wGS;
          s_B[75] = ch * L * wGJ; s_B[76] = ch * L * wEI; s_B[77] = ch * L * wEI;
        }
        __syncthreads();
        // v_i[b] = sum_d B_i[d] * E[dof_d][b]
        const int pa = node_pos(a, nn, nbn) * 6, pb = node_pos(b, nn, nbn) * 6;
        for (int w = tid; w < 6 * nB; w += SCHUR_BLOCK) {
          const int i = w / nB, bb = w - i * nB;
          double v = 0.0;
#pragma unroll
          for (int d = 0; d < 12; ++d) {
            const int q = (d < 6 ? pa + d : pb + d - 6);
            const double Eqb = (q < nI) ? -A[(size_t)(nI + bb) * ld + q] : ((q - nI == bb) ? 1.0 : 0.0);
            v = fma(s_B[i * 12 + d], Eqb, v);
          }
          s_v[i * SCHUR_MAX_NB + bb] = v;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < GRAD_PER_THREAD; ++u) {
          const int q = q0 + tid + u * SCHUR_BLOCK;
          if (q < nB * nB) {
            const int b1 = q / nB, b2 = q - b1 * nB;
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < 6; ++i) s = fma(s_B[72 + i] * s_v[i * SCHUR_MAX_NB + b1], s_v[i * SCHUR_MAX_NB + b2], s);
            acc[u] += s;
          }
        }
        __syncthreads();
      }
      double* dSc = dS + (c * (int64_t)n_grad + gsel) * (int64_t)nB * nB;
#pragma unroll
      for (int u = 0; u < GRAD_PER_THREAD; ++u) {
        const int q = q0 + tid + u * SCHUR_BLOCK;
        if (q < nB * nB) dSc[q] = bad ? qnan : acc[u];
      }
    }
  }
}

extern "C" int lat_schur_batch(lat_ctx* ctx, const double* xyz, const int32_t* len0, const int32_t* len1,
                               const double* rad, int64_t n_cells, int32_t n_loc_nodes, int32_t n_bnd_nodes,
                               int32_t n_loc_elem, double young, double nu, double kappa, double* S,
                               const int32_t* elem_group, const double* drad_chain, int32_t n_grad, double* dS) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, xyz && len0 && len1 && rad && S);
  LAT_CHECK_ARG(ctx, n_cells >= 0 && n_loc_nodes > 0 && n_bnd_nodes > 0 && n_bnd_nodes <= n_loc_nodes && n_loc_elem > 0);
  LAT_CHECK_ARG(ctx, 6 * n_bnd_nodes <= SCHUR_MAX_NB);
  LAT_CHECK_ARG(ctx, dS == nullptr || (elem_group != nullptr && n_grad > 0));
  if (n_cells == 0) return LAT_OK;
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  const int n = 6 * n_loc_nodes, ld = n | 1;
  const size_t aux = (size_t)n_loc_elem * sizeof(ElemCoef) + (size_t)n * 8 + 6 * SCHUR_MAX_NB * 8 + 80 * 8 + (size_t)n * 4 +
                     2 * (size_t)n_loc_elem * 4 + 64;
  const size_t a_bytes = (size_t)n * ld * sizeof(double);
  const bool in_smem = a_bytes + aux <= 220 * 1024;
  const size_t smem = in_smem ? a_bytes + aux : aux;
  if (smem > 220 * 1024) return lat_fail(ctx, LAT_ERR_UNSUPPORTED, "cell mesh too large for lat_schur_batch", __FILE__, __LINE__);
  int per_sm = in_smem ? (int)((220 * 1024) / (smem + 1024)) : 2;
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  int64_t grid = (int64_t)ctx->sm_count * per_sm;
  if (grid > n_cells) grid = n_cells;
  if (in_smem) {
    LAT_CUDA(ctx, cudaFuncSetAttribute(k_schur_dense<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LAT_LAUNCH(ctx, k_schur_dense<true>, (unsigned)grid, SCHUR_BLOCK, smem, xyz, len0, len1, rad, n_cells, n_loc_nodes,
               n_bnd_nodes, n_loc_elem, young, nu, kappa, S, nullptr, elem_group, drad_chain, n_grad, dS);
  } else {
    double* ws = lat_buf<double>(ctx, "schur_ws", (size_t)grid * n * ld);
    if (!ws) return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
    LAT_CUDA(ctx, cudaFuncSetAttribute(k_schur_dense<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LAT_LAUNCH(ctx, k_schur_dense<false>, (unsigned)grid, SCHUR_BLOCK, smem, xyz, len0, len1, rad, n_cells, n_loc_nodes,
               n_bnd_nodes, n_loc_elem, young, nu, kappa, S, ws, elem_group, drad_chain, n_grad, dS);
  }
  return LAT_OK;
}

// Same result through the strut pre-pass: chains of collinear elements between joints are condensed first
// (k_chain_condense), then the dense kernel runs on the joint-only cell.  chain_a / chain_b: the two joints of
// each chain in the REDUCED numbering (boundary joints first, in the order of S's rows).
extern "C" int lat_schur_batch_chains(lat_ctx* ctx, const double* xyz, const int32_t* len0, const int32_t* len1,
                                      const double* rad, int64_t n_cells, int32_t n_loc_nodes, int32_t n_loc_elem,
                                      const int32_t* chain_ptr, const int32_t* chain_elem, const int32_t* chain_flip,
                                      const int32_t* chain_a, const int32_t* chain_b, int32_t n_chains,
                                      int32_t n_joints, int32_t n_bnd_nodes, double young, double nu, double kappa,
                                      double* S) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, xyz && len0 && len1 && rad && S && chain_ptr && chain_elem && chain_flip && chain_a && chain_b);
  LAT_CHECK_ARG(ctx, n_cells >= 0 && n_loc_nodes > 0 && n_loc_elem > 0 && n_chains > 0);
  LAT_CHECK_ARG(ctx, n_bnd_nodes > 0 && n_bnd_nodes <= n_joints && n_joints <= n_loc_nodes && 6 * n_bnd_nodes <= SCHUR_MAX_NB);
  if (n_cells == 0) return LAT_OK;
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  SupCoef* sup = lat_buf<SupCoef>(ctx, "schur_sup", (size_t)n_cells * n_chains);
  if (!sup) return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
  LAT_LAUNCH(ctx, k_chain_condense, (unsigned)ceil_div(n_cells * n_chains, 128), 128, 0, xyz, len0, len1, rad, n_cells,
             n_loc_nodes, n_loc_elem, chain_ptr, chain_elem, chain_flip, n_chains, young, nu, kappa, sup);
  const int n = 6 * n_joints, ld = n | 1;
  const size_t aux = (size_t)n_chains * sizeof(SupCoef) + (size_t)n * 8 + 6 * SCHUR_MAX_NB * 8 + 80 * 8 + (size_t)n * 4 +
                     2 * (size_t)n_chains * 4 + 64;
  const size_t a_bytes = (size_t)n * ld * sizeof(double);
  const bool in_smem = a_bytes + aux <= 220 * 1024;
  const size_t smem = in_smem ? a_bytes + aux : aux;
  if (smem > 220 * 1024) return lat_fail(ctx, LAT_ERR_UNSUPPORTED, "cell mesh too large for lat_schur_batch_chains", __FILE__, __LINE__);
  int per_sm = in_smem ? (int)((220 * 1024) / (smem + 1024)) : 2;
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  int64_t grid = (int64_t)ctx->sm_count * per_sm;
  if (grid > n_cells) grid = n_cells;
  if (in_smem) {
    LAT_CUDA(ctx, cudaFuncSetAttribute(k_schur_dense<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LAT_LAUNCH(ctx, (k_schur_dense<true, true>), (unsigned)grid, SCHUR_BLOCK, smem, nullptr, chain_a, chain_b, nullptr, n_cells,
               n_joints, n_bnd_nodes, n_chains, young, nu, kappa, S, nullptr, nullptr, nullptr, 0, nullptr, sup);
  } else {
    double* ws = lat_buf<double>(ctx, "schur_ws", (size_t)grid * n * ld);
    if (!ws) return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
    LAT_CUDA(ctx, cudaFuncSetAttribute(k_schur_dense<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LAT_LAUNCH(ctx, (k_schur_dense<false, true>), (unsigned)grid, SCHUR_BLOCK, smem, nullptr, chain_a, chain_b, nullptr, n_cells,
               n_joints, n_bnd_nodes, n_chains, young, nu, kappa, S, ws, nullptr, nullptr, 0, nullptr, sup);
  }
  return LAT_OK;
}


// ===========================================================================
// A7, star cells: warp-level Schur complement (+ analytic dS/dr) of cells with ONE interior joint
// ===========================================================================
// After the strut pre-pass a BCC cell is 8 corner joints + 1 centre joint joined by 8 super-elements: 54 DOFs with SIX
// interior pivots.  k_schur_dense<SUPER> spends a 256-thread CTA, a 54x55 shared-memory matrix and ~20 CTA barriers on
// that (0.05 of HBM); here a HALF-WARP owns a cell:
//   S(k, l) = delta_kl D_k - O_k Kcc^-1 O_l^T,   Kcc = sum_k C_k
// with D_k / C_k the corner / centre diagonal 6x6 blocks and O_k the corner-row, centre-column block of strut k.  Lane k
// builds the blocks of strut k from its 16 SupCoef scalars, one lane inverts the 6x6 Kcc in registers, lane h then owns
// rows h, h+16, ... of S and writes them with 128-bit stores.  No n x n matrix, no CTA barrier.
// Sensitivities: the strut pre-pass is differentiated in forward mode (Dual below) w.r.t. the radius parameter of the
// strut's group, giving dC, dO, dD, and
//   dS(k, l) = delta_kl dD_k - dW_k O_l^T - W_k dO_l^T,   W_k = O_k Kcc^-1,   dW_k = dO_k Kcc^-1 - W_k dKcc Kcc^-1
// -- the gradients no longer need the dense 822-interior-DOF route (0.07 M cells/s in round 1).
struct Dual {
  double v, d;
  __host__ __device__ Dual() : v(0.0), d(0.0) {}
  __host__ __device__ Dual(double a) : v(a), d(0.0) {}
  __host__ __device__ Dual(double a, double b) : v(a), d(b) {}
};
__host__ __device__ __forceinline__ Dual operator+(Dual a, Dual b) { return Dual(a.v + b.v, a.d + b.d); }
__host__ __device__ __forceinline__ Dual operator-(Dual a, Dual b) { return Dual(a.v - b.v, a.d - b.d); }
__host__ __device__ __forceinline__ Dual operator-(Dual a) { return Dual(-a.v, -a.d); }
__host__ __device__ __forceinline__ Dual operator*(Dual a, Dual b) { return Dual(a.v * b.v, a.d * b.v + a.v * b.d); }
__host__ __device__ __forceinline__ Dual operator/(Dual a, Dual b) { const double q = a.v / b.v; return Dual(q, (a.d - q * b.d) / b.v); }

// k_chain_condense in dual numbers: sup = value, dsup = d/d(rho) with r_e = r_e(rho), dr_e/drho = drad_chain[e]
// (1.5 on penalised segments, NULL -> 1).  Same recursion, same order of operations for the value part.
__global__ void __launch_bounds__(128) k_chain_condense_dual(
    const double* __restrict__ xyz, const int32_t* __restrict__ len0, const int32_t* __restrict__ len1,
    const double* __restrict__ rad, const double* __restrict__ drad_chain, int64_t n_cells, int nn, int ne,
    const int32_t* __restrict__ chain_ptr, const int32_t* __restrict__ chain_elem, const int32_t* __restrict__ chain_flip,
    int n_chains, double young, double nu, double kappa, SupCoef* __restrict__ sup, SupCoef* __restrict__ dsup) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_cells * n_chains) return;
  const int64_t c = w / n_chains;
  const int ch = (int)(w - c * n_chains);
  const double* cx = xyz + c * (int64_t)nn * 3;
  const double* cr = rad + c * (int64_t)ne;
  const double PI = 3.14159265358979323846, G = young / (2.0 * (1.0 + nu));
  Dual flex_ax, flex_tor;
  double tx = 0.0, ty = 0.0, tz = 0.0;
  Dual aa[2][2], ak[2][2], kk[2][2];
  for (int q = chain_ptr[ch]; q < chain_ptr[ch + 1]; ++q) {
    const int e = chain_elem[q];
    const bool fl = chain_flip[q] != 0;
    const int a = fl ? len1[e] : len0[e], b = fl ? len0[e] : len1[e];
    const double dx = cx[b * 3] - cx[a * 3], dy = cx[b * 3 + 1] - cx[a * 3 + 1], dz = cx[b * 3 + 2] - cx[a * 3 + 2];
    const double L = sqrt(dx * dx + dy * dy + dz * dz), iL = 1.0 / L;
    const Dual r(cr[e], drad_chain ? drad_chain[e] : 1.0);
    const Dual S = Dual(PI) * r * r, I = Dual(PI * 0.25) * r * r * r * r;
    const Dual ES = Dual(young) * S, GS = Dual(G * kappa) * S, EI = Dual(young) * I, GJ = Dual(G * 2.0) * I;
    flex_ax = flex_ax + Dual(L) / ES;
    flex_tor = flex_tor + Dual(L) / GJ;
    const Dual g1 = GS * Dual(iL), g2 = Dual(0.5) * GS, dp = Dual(0.25 * L) * GS + EI * Dual(iL), dm = Dual(0.25 * L) * GS - EI * Dual(iL);
    if (q == chain_ptr[ch]) {
      tx = dx * iL; ty = dy * iL; tz = dz * iL;
      aa[0][0] = g1; aa[0][1] = -g2; aa[1][0] = -g2; aa[1][1] = dp;
      ak[0][0] = -g1; ak[0][1] = -g2; ak[1][0] = g2; ak[1][1] = dm;
      kk[0][0] = g1; kk[0][1] = g2; kk[1][0] = g2; kk[1][1] = dp;
      continue;
    }
    const Dual p00 = kk[0][0] + g1, p01 = kk[0][1] - g2, p11 = kk[1][1] + dp;
    const Dual idet = Dual(1.0) / (p00 * p11 - p01 * p01);
    const Dual i00 = p11 * idet, i01 = -(p01 * idet), i11 = p00 * idet;
    const Dual ekn[2][2] = {{-g1, -g2}, {g2, dm}};
    const Dual enn[2][2] = {{g1, g2}, {g2, dp}};
    Dual x[2][2], y[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      x[i][0] = ak[i][0] * i00 + ak[i][1] * i01;
      x[i][1] = ak[i][0] * i01 + ak[i][1] * i11;
      y[i][0] = ekn[0][i] * i00 + ekn[1][i] * i01;
      y[i][1] = ekn[0][i] * i01 + ekn[1][i] * i11;
    }
    Dual naa[2][2], nak[2][2], nkk[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        naa[i][j] = aa[i][j] - (x[i][0] * ak[j][0] + x[i][1] * ak[j][1]);
        nak[i][j] = -(x[i][0] * ekn[0][j] + x[i][1] * ekn[1][j]);
        nkk[i][j] = enn[i][j] - (y[i][0] * ekn[0][j] + y[i][1] * ekn[1][j]);
      }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) { aa[i][j] = naa[i][j]; ak[i][j] = nak[i][j]; kk[i][j] = nkk[i][j]; }
  }
  const Dual kax = Dual(1.0) / flex_ax, ktor = Dual(1.0) / flex_tor;
  SupCoef o, d;
  o.tx = tx; o.ty = ty; o.tz = tz; d.tx = tx; d.ty = ty; d.tz = tz;
  o.kax = kax.v; d.kax = kax.d; o.ktor = ktor.v; d.ktor = ktor.d;
  const Dual m01 = Dual(0.5) * (aa[0][1] + aa[1][0]), m23 = Dual(0.5) * (kk[0][1] + kk[1][0]);
  const Dual mm[10] = {aa[0][0], m01, ak[0][0], ak[0][1], aa[1][1], ak[1][0], ak[1][1], kk[0][0], m23, kk[1][1]};
#pragma unroll
  for (int k = 0; k < 10; ++k) { o.m[k] = mm[k].v; d.m[k] = mm[k].d; }
  o.pad = 0.0; d.pad = 0.0;
  sup[w] = o;
  dsup[w] = d;
}

// Topology of a star cell, shared by all cells of the batch (device arrays of n_struts entries):
//   corner[k]  boundary joint (= block row of S) at the far end of strut k;  cend[k]  which end (0 / 1) is the centre
//   strut_of[j] strut whose corner is boundary joint j;  group[k]  radius group of strut k (sensitivities)
// a half-warp per cell; 8 cells (128 threads, 60 KB of shared memory for 8 struts) per CTA for the values, 4 with the
// sensitivities (7 instead of 3 block sets per cell: 8 cells would need 134 KB = ONE resident CTA, 4 warps per SM)
__host__ __device__ constexpr int star_cells_per_cta(bool grad) { return grad ? 4 : 8; }
static constexpr int STAR_MAX_STRUTS = 16;           // one strut per lane of the half-warp

__device__ __forceinline__ void sup_block(const SupCoef& s, int re, int ce, double* dst /*[36] shared*/) {
  double q[36];
#pragma unroll
  for (int k = 0; k < 36; ++k) q[k] = 0.0;
  sup_block_accum(s, re, ce, q);
#pragma unroll
  for (int k = 0; k < 36; ++k) dst[k] = q[k];
}

// Four (at the row tail: fewer) consecutive doubles of a row of S: one 256-bit store when the piece is 32-byte aligned
// (always for an even number of struts: rows are multiples of 96 bytes then)
__device__ __forceinline__ void star_store4(double* dst, const double (&v)[4], int ncol) {
  if (ncol == 4 && (reinterpret_cast<uintptr_t>(dst) & 31) == 0) {
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
  } else if (ncol == 4 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    reinterpret_cast<double2*>(dst)[0] = make_double2(v[0], v[1]);
    reinterpret_cast<double2*>(dst)[1] = make_double2(v[2], v[3]);
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (q < ncol) dst[q] = v[q];
  }
}

template <bool GRAD>
__global__ void __launch_bounds__(16 * star_cells_per_cta(GRAD), 3) k_schur_star(
    const SupCoef* __restrict__ sup, const SupCoef* __restrict__ dsup, const int32_t* __restrict__ cend,
    const int32_t* __restrict__ strut_of, const int32_t* __restrict__ group, int64_t n_cells, int ns, int n_grad,
    double* __restrict__ S, double* __restrict__ dS) {
  extern __shared__ __align__(16) double star_smem[];
  // per cell slot: O, W, D (+ dO, dW, dD, dW of the current group), then Kinv[36], T[36]
  const int per_cell = (GRAD ? 7 : 3) * ns * 36 + 72;
  const int slot = threadIdx.x >> 4, h = threadIdx.x & 15;
  double* base = star_smem + (size_t)slot * per_cell;
  double* sO = base;
  double* sW = sO + ns * 36;
  double* sD = sW + ns * 36;
  double* sdO = GRAD ? sD + ns * 36 : nullptr;
  double* sdW = GRAD ? sdO + ns * 36 : nullptr;
  double* sdD = GRAD ? sdW + ns * 36 : nullptr;
  double* sG = GRAD ? sdD + ns * 36 : nullptr;
  double* sK = base + (GRAD ? 7 : 3) * ns * 36;   // Kcc^-1
  double* sT = sK + 36;                           // scratch 6x6
  const int nB = 6 * ns;
  __shared__ int s_strut_of[STAR_MAX_STRUTS], s_group[STAR_MAX_STRUTS];
  if (threadIdx.x < STAR_MAX_STRUTS) {
    s_strut_of[threadIdx.x] = (int)threadIdx.x < ns ? strut_of[threadIdx.x] : 0;
    s_group[threadIdx.x] = (GRAD && (int)threadIdx.x < ns) ? group[threadIdx.x] : 0;
  }
  __syncthreads();
  constexpr int CPC = star_cells_per_cta(GRAD);
  for (int64_t cell0 = (int64_t)blockIdx.x * CPC; cell0 < n_cells; cell0 += (int64_t)gridDim.x * CPC) {
    const int64_t cell = cell0 + slot;
    const bool live = cell < n_cells;
    // 1. strut blocks: lane k builds O_k, D_k and (into W's space) C_k
    if (live && h < ns) {
      const SupCoef s = sup[cell * ns + h];
      const int ec = cend[h], eb = ec ^ 1;
      sup_block(s, eb, ec, sO + h * 36);
      sup_block(s, eb, eb, sD + h * 36);
      sup_block(s, ec, ec, sW + h * 36);
      if (GRAD) {
        const SupCoef d = dsup[cell * ns + h];
        sup_block(d, eb, ec, sdO + h * 36);
        sup_block(d, eb, eb, sdD + h * 36);
        sup_block(d, ec, ec, sdW + h * 36);
      }
    }
    __syncwarp();
    // 2. Kcc = sum_k C_k (fixed order), 3. inverse by Gauss-Jordan without pivoting (SPD) in lane 0's registers
    for (int e = h; e < 36; e += 16) {
      double acc = 0.0;
      for (int k = 0; k < ns; ++k) acc += sW[k * 36 + e];
      sT[e] = acc;
    }
    __syncwarp();
    if (h == 0) {
      double a[6][6], inv[6][6];
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int k = 0; k < 6; ++k) { a[i][k] = sT[i * 6 + k]; inv[i][k] = (i == k) ? 1.0 : 0.0; }
#pragma unroll
      for (int p = 0; p < 6; ++p) {
        const double ip = 1.0 / a[p][p];
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
        for (int k = 0; k < 6; ++k) sK[i * 6 + k] = 0.5 * (inv[i][k] + inv[k][i]);
    }
    __syncwarp();
    // 4. W_k = O_k Kcc^-1 (lane k)
    if (h < ns) {
      double w[36];
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          double acc = 0.0;
#pragma unroll
          for (int m = 0; m < 6; ++m) acc = fma(sO[h * 36 + i * 6 + m], sK[m * 6 + j], acc);
          w[i * 6 + j] = acc;
        }
#pragma unroll
      for (int k = 0; k < 36; ++k) sW[h * 36 + k] = w[k];
    }
    __syncwarp();
    // 5. S row by row, the whole half-warp on ONE row: lane h owns columns 4h .. 4h+3 (+64, ...), so a row leaves as
    //    contiguous 32-byte pieces -- 384 contiguous bytes per row and instruction for n_s = 8.  (First version: every
    //    lane owned whole rows; a store instruction then touched 32 different lines and the kernel ran at 98 % of the
    //    L1 store-wavefront limit, 0.46 of HBM: profiles/r02_ncu_schur_star.txt.)
    //    S[i][6 l' + b] = delta D_k[a][b] - sum_m W_k[a][m] O_l[b][m]
    for (int c0 = 4 * h; c0 < nB; c0 += 64) {
      int ofs[4], jcol[4], bq[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int col = c0 + q < nB ? c0 + q : nB - 1;
        jcol[q] = col / 6;
        bq[q] = col - 6 * jcol[q];
        ofs[q] = strut_of[jcol[q]] * 36 + bq[q] * 6;
      }
      const int ncol = nB - c0 < 4 ? nB - c0 : 4;
      if (live) {
        // the lane's four columns are the same for every row: their O_l[b][:] stay in registers (24 doubles), the
        // row's W_k[a][:] is a broadcast read
        double oc[4][6];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int m = 0; m < 6; ++m) oc[q][m] = sO[ofs[q] + m];
        double* out = S + cell * (int64_t)nB * nB + c0;
        for (int j = 0; j < ns; ++j) {
          const int k = s_strut_of[j];
#pragma unroll
          for (int a = 0; a < 6; ++a) {
            const double* wrow = sW + k * 36 + a * 6;
            double v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              double acc = (jcol[q] == j) ? sD[k * 36 + a * 6 + bq[q]] : 0.0;
#pragma unroll
              for (int m = 0; m < 6; ++m) acc = fma(-wrow[m], oc[q][m], acc);
              v[q] = acc;
            }
            star_store4(out, v, ncol);
            out += nB;
          }
        }
      }
    }
    if (GRAD) {
      for (int gsel = 0; gsel < n_grad; ++gsel) {
        __syncwarp();
        // dKcc = sum_{k in g} dC_k  (dC sits in sdW)
        for (int e = h; e < 36; e += 16) {
          double acc = 0.0;
          for (int k = 0; k < ns; ++k)
            if (group[k] == gsel) acc += sdW[k * 36 + e];
          sT[e] = acc;
        }
        __syncwarp();
        // T = dKcc Kcc^-1 (36 entries over the lanes), kept in registers then written back
        double t_loc[3] = {0.0, 0.0, 0.0};
        for (int q = 0, e = h; e < 36; e += 16, ++q) {
          const int i = e / 6, jj = e - 6 * i;
          double acc = 0.0;
#pragma unroll
          for (int m = 0; m < 6; ++m) acc = fma(sT[i * 6 + m], sK[m * 6 + jj], acc);
          t_loc[q] = acc;
        }
        __syncwarp();
        for (int q = 0, e = h; e < 36; e += 16, ++q) sT[e] = t_loc[q];
        __syncwarp();
        // dW_k = [k in g] dO_k Kcc^-1 - W_k T  (lane k, into its own scratch block: sdW keeps holding dC_k, which the
        // next group needs again)
        if (h < ns) {
          const bool kin = group[h] == gsel;
          double w[36];
#pragma unroll
          for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int jj = 0; jj < 6; ++jj) {
              double acc = 0.0;
#pragma unroll
              for (int m = 0; m < 6; ++m) {
                if (kin) acc = fma(sdO[h * 36 + i * 6 + m], sK[m * 6 + jj], acc);
                acc = fma(-sW[h * 36 + i * 6 + m], sT[m * 6 + jj], acc);
              }
              w[i * 6 + jj] = acc;
            }
#pragma unroll
          for (int e = 0; e < 36; ++e) sG[h * 36 + e] = w[e];
        }
        __syncwarp();
        for (int c0 = 4 * h; c0 < nB; c0 += 64) {
          int ofs[4], jcol[4], bq[4];
          bool lin[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int col = c0 + q < nB ? c0 + q : nB - 1;
            jcol[q] = col / 6;
            bq[q] = col - 6 * jcol[q];
            const int l = strut_of[jcol[q]];
            ofs[q] = l * 36 + bq[q] * 6;
            lin[q] = group[l] == gsel;
          }
          const int ncol = nB - c0 < 4 ? nB - c0 : 4;
          if (live) {
            double oc[4][6], doc[4][6];
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
              for (int m = 0; m < 6; ++m) { oc[q][m] = sO[ofs[q] + m]; doc[q][m] = lin[q] ? sdO[ofs[q] + m] : 0.0; }
            double* out = dS + (cell * n_grad + gsel) * (int64_t)nB * nB + c0;
            for (int j = 0; j < ns; ++j) {
              const int k = s_strut_of[j];
              const bool kin = s_group[k] == gsel;
#pragma unroll
              for (int a = 0; a < 6; ++a) {
                const double* wrow = sW + k * 36 + a * 6;
                const double* dwrow = sG + k * 36 + a * 6;
                double v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  double acc = (jcol[q] == j && kin) ? sdD[k * 36 + a * 6 + bq[q]] : 0.0;
#pragma unroll
                  for (int m = 0; m < 6; ++m) { acc = fma(-dwrow[m], oc[q][m], acc); acc = fma(-wrow[m], doc[q][m], acc); }
                  v[q] = acc;
                }
                star_store4(out, v, ncol);
                out += nB;
              }
            }
          }
        }
      }
    }
    __syncwarp();
  }
}

// Cells WITHOUT an interior joint (every joint lies on the cell boundary: Octet, 14 joints / 36 struts): after the strut
// pre-pass nothing is left to eliminate, S is the assembled joint-only cell matrix.  One warp per cell: lane k builds the
// (a, b) coupling block of strut k and lane j the diagonal block of joint j into shared memory, then the warp writes S
// row by row -- lane h owns columns 4h..4h+3 (+128, ...) of every row, a block row of six rows at a time, 256-bit
// stores of contiguous pieces; pair[i][j] = strut that joins joints i and j (its stored block is (a -> b): the other
// direction reads it transposed) or -1.  The dense kernel needs 5.5-9.4 ms for 64 000 Octet cells (0.06-0.10 of HBM).
static constexpr int DIRECT_WARPS = 8;    // 8 warps (1.84 ms) against 4 (2.11 ms) and 2 (2.12 ms) for 64 000 Octet cells
__global__ void __launch_bounds__(32 * DIRECT_WARPS) k_schur_direct(
    const SupCoef* __restrict__ sup, const int32_t* __restrict__ inc_ptr, const int16_t* __restrict__ inc,
    const int16_t* __restrict__ pair, int64_t n_cells, int ns, int nj, double* __restrict__ S, int64_t cell_stride,
    const int32_t* __restrict__ group, int gsel) {
  extern __shared__ __align__(16) double direct_smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int per_cell = (ns + nj) * 36;
  double* sB = direct_smem + (size_t)wid * per_cell;      // [ns][36] coupling blocks (row joint a, column joint b)
  double* sD = sB + ns * 36;                               // [nj][36] diagonal blocks
  __shared__ double sZ[36];                                // the block of a joint pair without a strut
  if (threadIdx.x < 36) sZ[threadIdx.x] = 0.0;
  __syncthreads();
  const int nB = 6 * nj;
  for (int64_t cell = (int64_t)blockIdx.x * DIRECT_WARPS + wid; cell < n_cells; cell += (int64_t)gridDim.x * DIRECT_WARPS) {
    __syncwarp();
    // sensitivities: S is linear in the super-elements, so dS/dr_g is the SAME assembly applied to the derivative
    // super-elements of the struts of radius group g (group != nullptr: all other struts contribute nothing)
    for (int k = lane; k < ns; k += 32) {
      if (group && group[k] != gsel) {
        for (int e = 0; e < 36; ++e) sB[k * 36 + e] = 0.0;
      } else sup_block(sup[cell * ns + k], 0, 1, sB + k * 36);
    }
    for (int j = lane; j < nj; j += 32) {
      double q[36];
#pragma unroll
      for (int e = 0; e < 36; ++e) q[e] = 0.0;
      // the struts of joint j in ascending strut order (fixed summation order); inc = +(k+1): end a, -(k+1): end b
      for (int t = inc_ptr[j]; t < inc_ptr[j + 1]; ++t) {
        const int p = inc[t];
        const int e = p > 0 ? 0 : 1, k = p > 0 ? p - 1 : -p - 1;
        if (group && group[k] != gsel) continue;
        sup_block_accum(sup[cell * ns + k], e, e, q);
      }
#pragma unroll
      for (int e = 0; e < 36; ++e) sD[j * 36 + e] = q[e];
    }
    __syncwarp();
    for (int c0 = 4 * lane; c0 < nB; c0 += 128) {
      int jq[4], bq[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int col = c0 + q < nB ? c0 + q : nB - 1;
        jq[q] = col / 6;
        bq[q] = col - 6 * jq[q];
      }
      const int ncol = nB - c0 < 4 ? nB - c0 : 4;
      double* out = S + cell * cell_stride + c0;
      // rows are multiples of 32 bytes when nB % 4 == 0: then every full piece of this lane is a 256-bit store
      const bool fast = ncol == 4 && (nB & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 31) == 0;
      for (int i = 0; i < nj; ++i) {
        // the lane's (at most two) column joints against row joint i: base + a * sa + offset, or the zero block
        const double* src[4];
        int sa[4], ob[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int j = jq[q];
          const int p = j == i ? 0 : pair[i * nj + j];      // strut k: +(k+1) if i is its end a, -(k+1) if i is its end b
          if (j == i) { src[q] = sD + i * 36; sa[q] = 6; ob[q] = bq[q]; }
          else if (p > 0) { src[q] = sB + (p - 1) * 36; sa[q] = 6; ob[q] = bq[q]; }
          else if (p < 0) { src[q] = sB + (-p - 1) * 36; sa[q] = 1; ob[q] = 6 * bq[q]; }     // transposed
          else { src[q] = sZ; sa[q] = 0; ob[q] = 0; }
        }
#pragma unroll
        for (int a = 0; a < 6; ++a) {
          double v[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) v[q] = src[q][a * sa[q] + ob[q]];
          if (fast) asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(out), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
          else star_store4(out, v, ncol);
          out += nB;
        }
      }
    }
  }
}

// Host: is this chain topology a star?  (one interior joint = joint n_bnd, every strut joins it to a distinct boundary
// joint, every boundary joint has a strut).  Fills the device tables of k_schur_star.
static bool star_topology(const std::vector<int32_t>& ca, const std::vector<int32_t>& cb, int n_joints, int n_bnd,
                          std::vector<int32_t>* cend, std::vector<int32_t>* strut_of) {
  const int ns = (int)ca.size();
  if (n_joints != n_bnd + 1 || ns != n_bnd || ns > STAR_MAX_STRUTS) return false;
  cend->assign(ns, 0);
  strut_of->assign(n_bnd, -1);
  for (int k = 0; k < ns; ++k) {
    int corner;
    if (ca[k] == n_bnd && cb[k] < n_bnd) { (*cend)[k] = 0; corner = cb[k]; }
    else if (cb[k] == n_bnd && ca[k] < n_bnd) { (*cend)[k] = 1; corner = ca[k]; }
    else return false;
    if ((*strut_of)[corner] >= 0) return false;
    (*strut_of)[corner] = k;
  }
  return true;
}

extern "C" int lat_schur_batch_struts(lat_ctx* ctx, const double* xyz, const int32_t* len0, const int32_t* len1,
                                      const double* rad, int64_t n_cells, int32_t n_loc_nodes, int32_t n_loc_elem,
                                      const int32_t* chain_ptr, const int32_t* chain_elem, const int32_t* chain_flip,
                                      const int32_t* chain_a, const int32_t* chain_b, int32_t n_chains, int32_t n_joints,
                                      int32_t n_bnd_nodes, double young, double nu, double kappa, double* S,
                                      const int32_t* chain_group, const double* drad_chain, int32_t n_grad, double* dS) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, xyz && len0 && len1 && rad && S && chain_ptr && chain_elem && chain_flip && chain_a && chain_b);
  LAT_CHECK_ARG(ctx, n_cells >= 0 && n_loc_nodes > 0 && n_loc_elem > 0 && n_chains > 0);
  LAT_CHECK_ARG(ctx, n_bnd_nodes > 0 && n_bnd_nodes <= n_joints && n_joints <= n_loc_nodes && 6 * n_bnd_nodes <= SCHUR_MAX_NB);
  LAT_CHECK_ARG(ctx, dS == nullptr || (chain_group != nullptr && n_grad > 0));
  if (n_cells == 0) return LAT_OK;
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  std::vector<int32_t> ca(n_chains), cb(n_chains), cend, strut_of;
  LAT_CUDA(ctx, cudaMemcpyAsync(ca.data(), chain_a, n_chains * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  LAT_CUDA(ctx, cudaMemcpyAsync(cb.data(), chain_b, n_chains * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  LAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (n_joints == n_bnd_nodes && n_chains <= 4096) {
    // no interior joint: S is the assembled joint-only matrix (k_schur_direct); two struts between the same pair of
    // joints are left to the general path
    std::vector<int16_t> pair((size_t)n_joints * n_joints, 0);
    bool simple = true;
    for (int k = 0; k < n_chains && simple; ++k) {
      const int a = ca[k], b = cb[k];
      if (a < 0 || b < 0 || a >= n_joints || b >= n_joints || a == b || pair[(size_t)a * n_joints + b] != 0) { simple = false; break; }
      pair[(size_t)a * n_joints + b] = (int16_t)(k + 1);
      pair[(size_t)b * n_joints + a] = (int16_t)(-(k + 1));
    }
    if (simple) {
      const int ns = n_chains;
      SupCoef* sup = lat_buf<SupCoef>(ctx, "schur_sup", (size_t)n_cells * ns);
      SupCoef* dsup = dS ? lat_buf<SupCoef>(ctx, "schur_dsup", (size_t)n_cells * ns) : nullptr;
      if (dS && !dsup) return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
      int16_t* dpair = lat_buf<int16_t>(ctx, "schur_pair", pair.size());
      int32_t* dincp = lat_buf<int32_t>(ctx, "schur_incp", (size_t)n_joints + 1);
      int16_t* dinc = lat_buf<int16_t>(ctx, "schur_inc", (size_t)2 * ns);
      if (!sup || !dpair || !dincp || !dinc) return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
      std::vector<int32_t> incp(n_joints + 1, 0);
      std::vector<int16_t> incv((size_t)2 * ns);
      for (int k = 0; k < ns; ++k) { incp[ca[k] + 1]++; incp[cb[k] + 1]++; }
      for (int j = 0; j < n_joints; ++j) incp[j + 1] += incp[j];
      {
        std::vector<int32_t> fill(incp.begin(), incp.end() - 1);
        for (int k = 0; k < ns; ++k) { incv[fill[ca[k]]++] = (int16_t)(k + 1); incv[fill[cb[k]]++] = (int16_t)(-(k + 1)); }
      }
      LAT_CUDA(ctx, cudaMemcpyAsync(dpair, pair.data(), pair.size() * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream));
      LAT_CUDA(ctx, cudaMemcpyAsync(dincp, incp.data(), incp.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
      LAT_CUDA(ctx, cudaMemcpyAsync(dinc, incv.data(), incv.size() * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream));
      LAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // the host tables go out of scope
      const size_t smem = (size_t)DIRECT_WARPS * (ns + n_joints) * 36 * sizeof(double);
      if (smem <= 200 * 1024) {
        const unsigned cgrid = (unsigned)ceil_div(n_cells * ns, 128);
        if (dS)
          LAT_LAUNCH(ctx, k_chain_condense_dual, cgrid, 128, 0, xyz, len0, len1, rad, drad_chain, n_cells, n_loc_nodes, n_loc_elem,
                     chain_ptr, chain_elem, chain_flip, ns, young, nu, kappa, sup, dsup);
        else
          LAT_LAUNCH(ctx, k_chain_condense, cgrid, 128, 0, xyz, len0, len1, rad, n_cells, n_loc_nodes, n_loc_elem, chain_ptr,
                     chain_elem, chain_flip, ns, young, nu, kappa, sup);
        LAT_CUDA(ctx, cudaFuncSetAttribute(k_schur_direct, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int64_t grid = ceil_div(n_cells, DIRECT_WARPS);
        const int64_t cap = (int64_t)ctx->sm_count * 16;
        if (grid > cap) grid = cap;
        const int64_t nB2 = (int64_t)36 * n_joints * n_joints;
        LAT_LAUNCH(ctx, k_schur_direct, (unsigned)grid, 32 * DIRECT_WARPS, smem, sup, dincp, dinc, dpair, n_cells, ns, n_joints, S, nB2,
                   nullptr, 0);
        for (int g = 0; dS && g < n_grad; ++g)
          LAT_LAUNCH(ctx, k_schur_direct, (unsigned)grid, 32 * DIRECT_WARPS, smem, dsup, dincp, dinc, dpair, n_cells, ns, n_joints,
                     dS + g * nB2, (int64_t)n_grad * nB2, chain_group, g);
        return LAT_OK;
      }
    }
  }
  if (!star_topology(ca, cb, n_joints, n_bnd_nodes, &cend, &strut_of)) {
    if (dS) return lat_fail(ctx, LAT_ERR_UNSUPPORTED, "sensitivities through the strut pre-pass need a star cell (one interior joint): use lat_schur_batch", __FILE__, __LINE__);
    return lat_schur_batch_chains(ctx, xyz, len0, len1, rad, n_cells, n_loc_nodes, n_loc_elem, chain_ptr, chain_elem, chain_flip,
                                  chain_a, chain_b, n_chains, n_joints, n_bnd_nodes, young, nu, kappa, S);
  }
  const int ns = n_chains;
  SupCoef* sup = lat_buf<SupCoef>(ctx, "schur_sup", (size_t)n_cells * ns);
  SupCoef* dsup = dS ? lat_buf<SupCoef>(ctx, "schur_dsup", (size_t)n_cells * ns) : nullptr;
  int32_t* tab = lat_buf<int32_t>(ctx, "schur_star_tab", (size_t)3 * STAR_MAX_STRUTS);
  if (!sup || (dS && !dsup) || !tab) return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
  std::vector<int32_t> host_tab(3 * STAR_MAX_STRUTS, 0);
  for (int k = 0; k < ns; ++k) { host_tab[k] = cend[k]; host_tab[STAR_MAX_STRUTS + k] = strut_of[k]; }
  if (dS) {
    std::vector<int32_t> grp(ns);
    LAT_CUDA(ctx, cudaMemcpyAsync(grp.data(), chain_group, ns * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    LAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < ns; ++k) host_tab[2 * STAR_MAX_STRUTS + k] = grp[k];
  }
  LAT_CUDA(ctx, cudaMemcpyAsync(tab, host_tab.data(), host_tab.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  LAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // host_tab goes out of scope
  const unsigned cgrid = (unsigned)ceil_div(n_cells * ns, 128);
  if (dS)
    LAT_LAUNCH(ctx, k_chain_condense_dual, cgrid, 128, 0, xyz, len0, len1, rad, drad_chain, n_cells, n_loc_nodes, n_loc_elem,
               chain_ptr, chain_elem, chain_flip, ns, young, nu, kappa, sup, dsup);
  else
    LAT_LAUNCH(ctx, k_chain_condense, cgrid, 128, 0, xyz, len0, len1, rad, n_cells, n_loc_nodes, n_loc_elem, chain_ptr,
               chain_elem, chain_flip, ns, young, nu, kappa, sup);
  const int cpc = star_cells_per_cta(dS != nullptr);
  const size_t smem = (size_t)cpc * ((dS ? 7 : 3) * ns * 36 + 72) * sizeof(double);
  int64_t grid = ceil_div(n_cells, cpc);
  const int64_t cap = (int64_t)ctx->sm_count * 8;
  if (grid > cap) grid = cap;
  if (dS) {
    LAT_CUDA(ctx, cudaFuncSetAttribute(k_schur_star<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LAT_LAUNCH(ctx, k_schur_star<true>, (unsigned)grid, 16 * cpc, smem, sup, dsup, tab, tab + STAR_MAX_STRUTS,
               tab + 2 * STAR_MAX_STRUTS, n_cells, ns, (int)n_grad, S, dS);
  } else {
    LAT_CUDA(ctx, cudaFuncSetAttribute(k_schur_star<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LAT_LAUNCH(ctx, k_schur_star<false>, (unsigned)grid, 16 * cpc, smem, sup, nullptr, tab, tab + STAR_MAX_STRUTS,
               tab + 2 * STAR_MAX_STRUTS, n_cells, ns, 0, S, nullptr);
  }
  return LAT_OK;
}

// ===========================================================================
// A8: DDM interface operator  y = sum_c B_c S_c B_c^T x
// ===========================================================================
// One warp per cell.  The warp gathers the cell's boundary displacements (lane j holds
// u_c[j], u_c[j+32], u_c[j+64]), streams S_c row by row with coalesced loads, reduces each
// row product with shuffles, and scatter-adds the 6 n_bnd results into y with FP64 atomics.
template <int NT>
__global__ void __launch_bounds__(256) k_ddm_matvec(const double* __restrict__ S, int64_t s_stride,
                                                    const int32_t* __restrict__ gidx, const double* __restrict__ u_fixed,
                                                    int64_t n_cells, int nb, const double* __restrict__ x,
                                                    double* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= n_cells) return;
  const int32_t* gi = gidx + c * nb;
  const double* Sc = S + c * s_stride;
  double xr[NT];
  int gl[NT];
#pragma unroll
  for (int q = 0; q < NT; ++q) {
    const int j = lane + 32 * q;
    gl[q] = -2;
    xr[q] = 0.0;
    if (j < nb) {
      gl[q] = gi[j];
      xr[q] = gl[q] >= 0 ? x[gl[q]] : (u_fixed ? u_fixed[c * nb + j] : 0.0);
    }
  }
  double yr[NT];
#pragma unroll
  for (int q = 0; q < NT; ++q) yr[q] = 0.0;
  // four rows per step: their loads are issued together and the four shuffle reductions interleave
  // (eight rows per step measured slower again: 0.53 / 0.81 against 0.61 / 0.87 of HBM at nb = 48 / 84)
  for (int i0 = 0; i0 < nb; i0 += 4) {
    double s[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = i0 + k;
      s[k] = 0.0;
      if (i < nb) {
        const double* row = Sc + (int64_t)i * nb;
#pragma unroll
        for (int q = 0; q < NT; ++q) {
          const int j = lane + 32 * q;
          if (j < nb) s[k] = fma(__ldcs(row + j), xr[q], s[k]);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int k = 0; k < 4; ++k) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = i0 + k;
      if (i < nb && lane == (i & 31)) {
#pragma unroll
        for (int q = 0; q < NT; ++q)
          if ((i >> 5) == q) yr[q] = s[k];
      }
    }
  }
#pragma unroll
  for (int q = 0; q < NT; ++q)
    if (gl[q] >= 0) atomicAdd(&y[gl[q]], yr[q]);
}

extern "C" int lat_ddm_matvec(lat_ctx* ctx, const double* S, int64_t s_stride, const int32_t* gidx,
                              const double* u_fixed, int64_t n_cells, int32_t nb, int64_t n_free,
                              const double* x, double* y) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, S && gidx && x && y && n_cells >= 0 && nb > 0 && nb <= SCHUR_MAX_NB && n_free > 0);
  LAT_CHECK_ARG(ctx, s_stride == 0 || s_stride >= (int64_t)nb * nb);
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  LAT_CUDA(ctx, cudaMemsetAsync(y, 0, n_free * sizeof(double), ctx->stream));
  if (n_cells == 0) return LAT_OK;
  const unsigned grid = (unsigned)ceil_div(n_cells, 8);
  switch ((nb + 31) / 32) {
    case 1: LAT_LAUNCH(ctx, k_ddm_matvec<1>, grid, 256, 0, S, s_stride, gidx, u_fixed, n_cells, nb, x, y); break;
    case 2: LAT_LAUNCH(ctx, k_ddm_matvec<2>, grid, 256, 0, S, s_stride, gidx, u_fixed, n_cells, nb, x, y); break;
    case 3: LAT_LAUNCH(ctx, k_ddm_matvec<3>, grid, 256, 0, S, s_stride, gidx, u_fixed, n_cells, nb, x, y); break;
    case 4: LAT_LAUNCH(ctx, k_ddm_matvec<4>, grid, 256, 0, S, s_stride, gidx, u_fixed, n_cells, nb, x, y); break;
    case 5: LAT_LAUNCH(ctx, k_ddm_matvec<5>, grid, 256, 0, S, s_stride, gidx, u_fixed, n_cells, nb, x, y); break;
    default: LAT_LAUNCH(ctx, k_ddm_matvec<6>, grid, 256, 0, S, s_stride, gidx, u_fixed, n_cells, nb, x, y); break;
  }
  return LAT_OK;
}


// ===========================================================================
// A9: assembled interface operator  K_G = sum_c P_c^T S_c P_c  as BSR(6x6)
// ===========================================================================
// LatticeSim.build_preconditioner / Cell.build_local_preconditioner (lattice_sim.py:1351-1415,
// cell.py:783-827) assemble the same matrix as COO triplets on the free DOFs and hand it to SuperLU.
// Here it is assembled on ALL interface DOFs (6 per boundary node, Dirichlet rows eliminated afterwards
// with lat_apply_dirichlet) so that the BSR PCG can solve the interface problem directly.
// One warp per (cell, row node a): for every column node b the 6x6 block of S_c is scatter-added
// (FP64 RED) into the BSR block (node_a, node_b) located by binary search.
__global__ void __launch_bounds__(256) k_assemble_cells(const double* __restrict__ S, int64_t s_stride,
                                                        const int32_t* __restrict__ cell_nodes, int64_t n_cells, int nbn,
                                                        const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                                        double* __restrict__ vals, int32_t* __restrict__ missing) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= n_cells * nbn) return;
  const int64_t c = wid / nbn;
  const int a = (int)(wid - c * nbn);
  const int nb = 6 * nbn;
  const int32_t* cn = cell_nodes + c * nbn;
  const int na = cn[a];
  if (na < 0) return;
  const double* Sc = S + c * s_stride;
  const int lo = rowptr[na], hi = rowptr[na + 1];
  for (int b = 0; b < nbn; ++b) {
    const int nbk = cn[b];
    if (nbk < 0) continue;
    int l = lo, h = hi;
    while (l < h) { const int mid = (l + h) >> 1; if (colidx[mid] < nbk) l = mid + 1; else h = mid; }
    if (l >= hi || colidx[l] != nbk) { if (lane == 0) atomicAdd(missing, 1); continue; }
    double* dst = vals + (int64_t)l * 36;
    for (int k = lane; k < 36; k += 32) {
      const int i = k / 6, j = k - i * 6;
      atomicAdd(dst + k, Sc[(int64_t)(a * 6 + i) * nb + b * 6 + j]);
    }
  }
}

extern "C" int lat_assemble_cells_bsr(lat_ctx* ctx, const double* S, int64_t s_stride, const int32_t* cell_nodes,
                                      int64_t n_cells, int32_t n_bnd_nodes, const int32_t* rowptr, const int32_t* colidx,
                                      int64_t nnzb, double* vals) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, S && cell_nodes && rowptr && colidx && vals && n_cells >= 0 && n_bnd_nodes > 0 && nnzb > 0);
  LAT_CHECK_ARG(ctx, 6 * n_bnd_nodes <= SCHUR_MAX_NB);
  LAT_CHECK_ARG(ctx, s_stride == 0 || s_stride >= (int64_t)36 * n_bnd_nodes * n_bnd_nodes);
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  int32_t* missing = lat_buf<int32_t>(ctx, "cells_missing", 4);
  if (!missing) return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
  LAT_CUDA(ctx, cudaMemsetAsync(missing, 0, 4 * sizeof(int32_t), ctx->stream));
  LAT_CUDA(ctx, cudaMemsetAsync(vals, 0, (size_t)nnzb * 36 * sizeof(double), ctx->stream));
  if (n_cells == 0) return LAT_OK;
  LAT_LAUNCH(ctx, k_assemble_cells, (unsigned)ceil_div(n_cells * n_bnd_nodes, 8), 256, 0, S, s_stride, cell_nodes, n_cells,
             (int)n_bnd_nodes, rowptr, colidx, vals, missing);
  int32_t h = 0;
  LAT_CUDA(ctx, cudaMemcpyAsync(&h, missing, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  LAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (h != 0) return lat_fail(ctx, LAT_ERR_ARG, "pattern does not contain every (node, node) pair of the cells", __FILE__, __LINE__);
  return LAT_OK;
}


// Gather form of the same assembly: a PLAN (built once per interface pattern by the host: ddm.InterfaceProblem) lists for
// every BSR block the (cell, row node, column node) contributions, sorted by block.  Eight lanes per block add its
// contributions in plan order -- fixed summation order (bit-reproducible matrix, unlike the FP64 reductions above), no
// atomics, every S entry read once and every block written once: the assembly of a design iteration (new radii -> new
// S, same plan) moves 8 nB^2 per cell + 288 B per block.
// contrib[k] = (cell * nbn + a) * nbn + b.  A group of lanes per block adds its contributions in plan order.
// Eight lanes per block (four blocks per warp, so four independent chains blk_ptr -> contrib -> S per warp are in flight):
// lane sl of a group owns entries sl, sl + 8, ..., sl + 32 (< 36) of the 6x6 block; two contributions are fetched per trip.
__global__ void __launch_bounds__(256) k_assemble_cells_gather(const double* __restrict__ S, int64_t s_stride, int nbn,
                                                               const int32_t* __restrict__ blk_ptr, const int64_t* __restrict__ contrib,
                                                               int64_t nnzb, double* __restrict__ vals) {
  const int sl = threadIdx.x & 7;
  const int64_t l = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  if (l >= nnzb) return;
  const int nb = 6 * nbn;
  const int lo = __ldg(blk_ptr + l), hi = __ldg(blk_ptr + l + 1);
  int off[5];
#pragma unroll
  for (int m = 0; m < 5; ++m) {
    const int e = sl + 8 * m;
    off[m] = e < 36 ? (e / 6) * nb + (e % 6) : 0;
  }
  const bool last = sl < 4;                    // entries 32..35
  double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  auto base_of = [&](int64_t d) -> const double* {
    const int64_t ca = d / nbn;
    const int b = (int)(d - ca * nbn);
    const int64_t c = ca / nbn;
    const int a = (int)(ca - c * nbn);
    return S + c * s_stride + (int64_t)(a * 6) * nb + b * 6;
  };
  int k = lo;
  for (; k + 1 < hi; k += 2) {
    const int64_t d0 = __ldg(contrib + k), d1 = __ldg(contrib + k + 1);
    const double* s0 = base_of(d0);
    const double* s1 = base_of(d1);
    double v0[5], v1[5];
#pragma unroll
    for (int m = 0; m < 5; ++m) {
      const bool on = m < 4 || last;
      v0[m] = on ? __ldg(s0 + off[m]) : 0.0;
      v1[m] = on ? __ldg(s1 + off[m]) : 0.0;
    }
#pragma unroll
    for (int m = 0; m < 5; ++m) { acc[m] += v0[m]; acc[m] += v1[m]; }      // plan order
  }
  if (k < hi) {
    const double* s0 = base_of(__ldg(contrib + k));
#pragma unroll
    for (int m = 0; m < 5; ++m)
      if (m < 4 || last) acc[m] += __ldg(s0 + off[m]);
  }
  double* dst = vals + l * 36;
#pragma unroll
  for (int m = 0; m < 4; ++m) dst[sl + 8 * m] = acc[m];
  if (last) dst[32 + sl] = acc[4];
}

extern "C" int lat_assemble_cells_bsr_plan(lat_ctx* ctx, const double* S, int64_t s_stride, int32_t n_bnd_nodes,
                                           const int32_t* blk_ptr, const int64_t* contrib, int64_t nnzb, double* vals) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, S && blk_ptr && contrib && vals && n_bnd_nodes > 0 && nnzb > 0);
  LAT_CHECK_ARG(ctx, 6 * n_bnd_nodes <= SCHUR_MAX_NB);
  LAT_CHECK_ARG(ctx, s_stride == 0 || s_stride >= (int64_t)36 * n_bnd_nodes * n_bnd_nodes);
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  LAT_LAUNCH(ctx, k_assemble_cells_gather, (unsigned)ceil_div(nnzb, 32), 256, 0, S, s_stride, (int)n_bnd_nodes, blk_ptr, contrib, nnzb, vals);
  return LAT_OK;
}


// ===========================================================================
// A11 (cell form): q[c][j] = v_c^T dS_{m(c,j)} u_c
// ===========================================================================
// The per-(cell, geometry) term of LatticeOpti.calculate_gradient (lattice_opti.py:752-761: u_cell @ (dS @ u_cell));
// the reference caches one dS per unique (geometry, radii) key (lattice_sim.py:857-883), hence the index table.
// One warp per (cell, j): rows of the matrix are read coalesced, each row product is reduced with a butterfly and
// weighted by v_i; fixed order -> reproducible.
__global__ void __launch_bounds__(256) k_cell_quadform(const double* __restrict__ mats, const int32_t* __restrict__ mat_index,
                                                       const double* __restrict__ U, const double* __restrict__ V,
                                                       int64_t n_pairs, int n_grad, int nb, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= n_pairs) return;
  const int64_t c = p / n_grad;
  const int m = mat_index[p];
  if (m < 0) { if (lane == 0) out[p] = 0.0; return; }
  const double* M = mats + (int64_t)m * nb * nb;
  const double* u = U + c * nb;
  const double* v = V ? V + c * nb : u;
  double acc = 0.0;
  for (int i = 0; i < nb; ++i) {
    double s = 0.0;
    for (int k = lane; k < nb; k += 32) s = fma(__ldg(M + (int64_t)i * nb + k), u[k], s);
    s = warp_sum(s);
    acc = fma(v[i], s, acc);
  }
  if (lane == 0) out[p] = acc;
}

extern "C" int lat_cell_quadform(lat_ctx* ctx, const double* mats, int64_t n_mats, const int32_t* mat_index, const double* U,
                                 const double* V, int64_t n_cells, int32_t n_grad, int32_t nb, double* out) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, mats && mat_index && U && out && n_mats > 0 && n_cells >= 0 && n_grad > 0 && nb > 0);
  if (n_cells == 0) return LAT_OK;
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t n_pairs = n_cells * n_grad;
  LAT_LAUNCH(ctx, k_cell_quadform, (unsigned)ceil_div(n_pairs, 8), 256, 0, mats, mat_index, U, V, n_pairs, (int)n_grad, (int)nb, out);
  return LAT_OK;
}

// matfree.cuh -- matrix-free stiffness operator: y = K u regenerated from the beam geometry.
//
// The assembled BSR operator streams 292 B per 6x6 block and per product; a B200 has ~37 TFLOP/s of FP64
// against ~6.5 TB/s of HBM, i.e. ~45 flops per streamed double.  For a circular section every 3x3
// sub-block of K_e is a*I + b*t t^T + c*[t]x (common.cuh, ElemCoef), so the action of one element on
// one of its end nodes needs only the unit vector t', six scalars and ~75 FMAs:
//
//   dw = w_i - w_j,  st = th_i + th_j,  dt = th_i - th_j,   t' = (x_j - x_i)/L   (own node i, other j)
//   f_w  = aI dw + aT t'(t'.dw) - c (t' x st)
//   f_th = bI (st - t'(t'.st)) + dI dt + dT t'(t'.dt) + c (t' x dw)
//
// (beam_model.py:197-216 + material_definition.py:147 restated; the end index drops out because the
// signs of the coupling blocks flip together with t).  Per incidence (element end) the kernel streams
// 32 B {other node, S, L, 1/L} and gathers 32 B of node data + 48 B of u -- about a tenth of the bytes of
// the assembled product -- and no matrix is stored at all (15 GB at octet 100^3).
//
// Dirichlet rows/columns are eliminated on the fly from a 6-bit per-node mask:  A = P K P + (I - P).
//
// Thread layout: ONE THREAD PER NODE walks the node's incidence list (sorted by neighbour) two records at a
// time.  A first version with the solver's 6-lanes-per-node layout + a shared-memory transpose measured
// 23 us at BCC 20^3 m=2 (80 % of the nodes have 2 incidences -> 4 of 6 lanes idle, 7 waves of CTAs, each a
// chain of 4 dependent L2 round trips); one thread per node is a single wave with no idle lanes.
#pragma once
#include <cstdint>
#include <climits>
#include "common.cuh"

struct __align__(32) MfInc {   // one element end, in the node-sorted incidence order of the pattern
  int32_t other, pad;
  double S, L, iL;             // section area, length, 1/length
};

struct MfOp {
  const int32_t* adjptr;       // [n_nodes + 1]
  const MfInc* inc;            // [2 * n_elem]
  const double* node4;         // [n_nodes][4]  x, y, z, Dirichlet mask (low 6 bits of the bit pattern)
  double E, Gk, G2mE, inv4pi;  // young, G*kappa, 2G - E, 1/(4 pi)
};

__device__ __forceinline__ void mf_ld256(const void* p, double& a, double& b, double& c, double& d) {
  // not volatile, no memory clobber: read-only data, the compiler is free to hoist / overlap these loads
  asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

struct MfCoef { double aI, aT, c, bI, dI, dT; };

__device__ __forceinline__ MfCoef mf_coef(const MfOp& op, double S, double L, double iL) {
  MfCoef k;
  const double GS = op.Gk * S, ES = op.E * S, Iv = S * S * op.inv4pi;   // I = pi r^4 / 4 = S^2 / (4 pi)
  k.aI = GS * iL;
  k.aT = (ES - GS) * iL;
  k.c = 0.5 * GS;
  k.bI = 0.25 * GS * L;
  k.dI = op.E * Iv * iL;
  k.dT = op.G2mE * Iv * iL;
  return k;
}

// Build the incidence records from the resident pattern (adj_el = 2*element + end).
__global__ void k_mf_setup_inc(const int32_t* __restrict__ adj_other, const int32_t* __restrict__ adj_el,
                               const double* __restrict__ x, const double* __restrict__ y,
                               const double* __restrict__ z, const int32_t* __restrict__ en0,
                               const int32_t* __restrict__ en1, const double* __restrict__ rad, int64_t n_inc,
                               MfInc* __restrict__ inc) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_inc) return;
  const int e = adj_el[i] >> 1;
  const int a = en0[e], c = en1[e];
  const double dx = x[c] - x[a], dy = y[c] - y[a], dz = z[c] - z[a];
  const double L = sqrt(dx * dx + dy * dy + dz * dz);
  const double r = rad[e];
  MfInc v;
  v.other = adj_other[i];
  v.pad = 0;
  v.S = 3.14159265358979323846 * r * r;   // material_definition.py:147
  v.L = L;
  v.iL = 1.0 / L;
  inc[i] = v;
}

__global__ void k_mf_setup_nodes(const double* __restrict__ x, const double* __restrict__ y,
                                 const double* __restrict__ z, const uint8_t* __restrict__ fixed, int64_t n_nodes,
                                 double* __restrict__ node4) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_nodes) return;
  long long m = 0;
  if (fixed)
    for (int r = 0; r < 6; ++r) m |= (long long)(fixed[n * 6 + r] != 0) << r;
  node4[n * 4 + 0] = x[n];
  node4[n * 4 + 1] = y[n];
  node4[n * 4 + 2] = z[n];
  node4[n * 4 + 3] = __longlong_as_double(m);
}

// One incidence: the action of the element (own node -> other node) on the own node, added to f.
struct MfRec { double ow, S, L, iL; };     // raw 32 B incidence record (ow: bit pattern, low word = other)
struct MfNode { double x, y, z, m; };      // raw 32 B node record
struct MfU { double2 a, b, c; };           // 6 DOFs of one node

__device__ __forceinline__ MfRec mf_load_rec(const MfOp& op, int j) {
  MfRec r;
  mf_ld256(op.inc + j, r.ow, r.S, r.L, r.iL);
  return r;
}
__device__ __forceinline__ MfNode mf_load_node(const MfOp& op, int64_t n) {
  MfNode v;
  mf_ld256(op.node4 + n * 4, v.x, v.y, v.z, v.m);
  return v;
}
// 6 consecutive doubles at v + 6 n as ONE 32 B + ONE 16 B access instead of three 16 B ones: 48 n is 32 B
// aligned for even n, 48 n + 16 for odd n.  The L1 cost of a scattered per-thread access is per instruction
// (one wavefront per lane and instruction), so this is a third fewer wavefronts on the gathers.
// same access past L1 (ghost entries written by a peer GPU while this kernel may already be running)
__device__ __forceinline__ MfU mf_load_u_cg(const double* __restrict__ v, int64_t n) {
  const char* base = reinterpret_cast<const char*>(v + n * 6);
  const bool odd = n & 1;
  const double2 a = __ldcg(reinterpret_cast<const double2*>(base + (odd ? 0 : 32)));
  double b0, b1, b2, b3;
  asm volatile("ld.global.cg.v4.f64 {%0, %1, %2, %3}, [%4];"
               : "=d"(b0), "=d"(b1), "=d"(b2), "=d"(b3) : "l"(base + (odd ? 16 : 0)) : "memory");
  MfU r;
  r.a = odd ? a : make_double2(b0, b1);
  r.b = odd ? make_double2(b0, b1) : make_double2(b2, b3);
  r.c = odd ? make_double2(b2, b3) : a;
  return r;
}
__device__ __forceinline__ MfU mf_load_u(const double* __restrict__ v, int64_t n) {
  const char* base = reinterpret_cast<const char*>(v + n * 6);
  const bool odd = n & 1;
  const double2 a = *reinterpret_cast<const double2*>(base + (odd ? 0 : 32));
  double b0, b1, b2, b3;
  // (no volatile / memory clobber: the vectors read by a product are never written by the same kernel)
  asm("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(b0), "=d"(b1), "=d"(b2), "=d"(b3) : "l"(base + (odd ? 16 : 0)));
  MfU r;
  r.a = odd ? a : make_double2(b0, b1);
  r.b = odd ? make_double2(b0, b1) : make_double2(b2, b3);
  r.c = odd ? make_double2(b2, b3) : a;
  return r;
}

template <bool MASKED>
__device__ __forceinline__ void mf_incidence(const MfOp& op, const MfRec& rec, const MfNode& nj, const MfU& uu,
                                             double xo, double yo, double zo, const double (&uo)[6],
                                             double (&f)[6]) {
  double uj[6] = {uu.a.x, uu.a.y, uu.b.x, uu.b.y, uu.c.x, uu.c.y};
  if (MASKED) {
    const int m = __double2loint(nj.m);
#pragma unroll
    for (int k = 0; k < 6; ++k) uj[k] = ((m >> k) & 1) ? 0.0 : uj[k];
  }
  const MfCoef k = mf_coef(op, rec.S, rec.L, rec.iL);
  const double tx = (nj.x - xo) * rec.iL, ty = (nj.y - yo) * rec.iL, tz = (nj.z - zo) * rec.iL;
  const double dwx = uo[0] - uj[0], dwy = uo[1] - uj[1], dwz = uo[2] - uj[2];
  const double sx = uo[3] + uj[3], sy = uo[4] + uj[4], sz = uo[5] + uj[5];
  const double dtx = uo[3] - uj[3], dty = uo[4] - uj[4], dtz = uo[5] - uj[5];
  const double t_dw = fma(tx, dwx, fma(ty, dwy, tz * dwz));
  const double t_s = fma(tx, sx, fma(ty, sy, tz * sz));
  const double t_dt = fma(tx, dtx, fma(ty, dty, tz * dtz));
  // t' x st and t' x dw
  const double c1x = ty * sz - tz * sy, c1y = tz * sx - tx * sz, c1z = tx * sy - ty * sx;
  const double c2x = ty * dwz - tz * dwy, c2y = tz * dwx - tx * dwz, c2z = tx * dwy - ty * dwx;
  const double qa = k.aT * t_dw;
  f[0] += fma(k.aI, dwx, fma(qa, tx, -k.c * c1x));
  f[1] += fma(k.aI, dwy, fma(qa, ty, -k.c * c1y));
  f[2] += fma(k.aI, dwz, fma(qa, tz, -k.c * c1z));
  const double qt = k.dT * t_dt - k.bI * t_s;
  f[3] += fma(k.bI, sx, fma(k.dI, dtx, fma(qt, tx, k.c * c2x)));
  f[4] += fma(k.bI, sy, fma(k.dI, dty, fma(qt, ty, k.c * c2y)));
  f[5] += fma(k.bI, sz, fma(k.dI, dtz, fma(qt, tz, k.c * c2z)));
}

static constexpr int MF_BLOCK = 64;    // small CTAs: 9 per SM at <= 113 registers cover 85 k nodes in one wave

// (A u)[6 n .. 6 n + 5] for ONE THREAD PER NODE.  The node's incidences are walked two at a time so that two
// independent gather chains (record -> node + u of the other end) are in flight per thread; joints and
// strut-interior nodes are numbered apart (mesh.py), so the degree is nearly uniform inside a warp.
// uo returns the node's own u (unmasked); f the product (identity rows already substituted when MASKED).
template <bool MASKED, bool CG = false>
__device__ __forceinline__ void mf_node_product(const MfOp& op, int64_t n, const double* __restrict__ u,
                                                double (&uo)[6], double (&f)[6]) {
  const int lo = __ldg(op.adjptr + n), hi = __ldg(op.adjptr + n + 1);
  const MfNode no = mf_load_node(op, n);
  const MfU uu = mf_load_u(u, n);
  uo[0] = uu.a.x; uo[1] = uu.a.y; uo[2] = uu.b.x; uo[3] = uu.b.y; uo[4] = uu.c.x; uo[5] = uu.c.y;
  const int mo = MASKED ? __double2loint(no.m) : 0;
  double um[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) { um[k] = ((mo >> k) & 1) ? 0.0 : uo[k]; f[k] = 0.0; }
  int j = lo;
  for (; j + 1 < hi; j += 2) {
    const MfRec ra = mf_load_rec(op, j), rb = mf_load_rec(op, j + 1);
    const int oa = __double2loint(ra.ow), ob = __double2loint(rb.ow);
    const MfNode na = mf_load_node(op, oa), nb = mf_load_node(op, ob);
    const MfU ua = CG ? mf_load_u_cg(u, oa) : mf_load_u(u, oa);   // CG: past L1 (fused-halo path, rows reading ghosts)
    const MfU ub = CG ? mf_load_u_cg(u, ob) : mf_load_u(u, ob);
    mf_incidence<MASKED>(op, ra, na, ua, no.x, no.y, no.z, um, f);
    mf_incidence<MASKED>(op, rb, nb, ub, no.x, no.y, no.z, um, f);
  }
  if (j < hi) {
    const MfRec ra = mf_load_rec(op, j);
    const int oa = __double2loint(ra.ow);
    const MfNode na = mf_load_node(op, oa);
    const MfU ua = CG ? mf_load_u_cg(u, oa) : mf_load_u(u, oa);
    mf_incidence<MASKED>(op, ra, na, ua, no.x, no.y, no.z, um, f);
  }
#pragma unroll
  for (int k = 0; k < 6; ++k)
    if ((mo >> k) & 1) f[k] = uo[k];   // eliminated DOF: identity row
}

__device__ __forceinline__ void mf_store6(double* __restrict__ y, int64_t n, const double (&f)[6]) {
  char* base = reinterpret_cast<char*>(y + n * 6);
  const bool odd = n & 1;
  *reinterpret_cast<double2*>(base + (odd ? 0 : 32)) = odd ? make_double2(f[0], f[1]) : make_double2(f[4], f[5]);
  const double b0 = odd ? f[2] : f[0], b1 = odd ? f[3] : f[1], b2 = odd ? f[4] : f[2], b3 = odd ? f[5] : f[3];
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(base + (odd ? 16 : 0)), "d"(b0), "d"(b1), "d"(b2), "d"(b3) : "memory");
}

// plain y = A u  (MASKED: Dirichlet-eliminated operator; otherwise the raw stiffness K)
template <bool MASKED>
__global__ void __launch_bounds__(MF_BLOCK) k_mf_apply(MfOp op, int64_t n_nodes, const double* __restrict__ u,
                                                       double* __restrict__ y) {
  const int64_t n = (int64_t)blockIdx.x * MF_BLOCK + threadIdx.x;
  if (n >= n_nodes) return;
  double uo[6], f[6];
  mf_node_product<MASKED>(op, n, u, uo, f);
  mf_store6(y, n, f);
}

// b = P (f - K g) + (I - P) g   (lifting of prescribed displacements; g is read on fixed DOFs only)
__global__ void __launch_bounds__(MF_BLOCK) k_mf_rhs(MfOp op, int64_t n_nodes, const double* __restrict__ gfull,
                                                     const double* __restrict__ fext, double* __restrict__ b) {
  const int64_t n = (int64_t)blockIdx.x * MF_BLOCK + threadIdx.x;
  if (n >= n_nodes) return;
  double go[6], kg[6];
  mf_node_product<false>(op, n, gfull, go, kg);   // raw K g
  const int m = __double2loint(op.node4[n * 4 + 3]);
  double out[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) out[r] = ((m >> r) & 1) ? go[r] : (fext ? fext[n * 6 + r] : 0.0) - kg[r];
  mf_store6(b, n, out);
}

// Diagonal 6x6 block of A = P K P + (I - P) for one node, row-major in a[][].
__device__ __forceinline__ void mf_diag_block(const MfOp& op, int64_t n, double (&a)[6][6]) {
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int k = 0; k < 6; ++k) a[i][k] = 0.0;
  const int lo = op.adjptr[n], hi = op.adjptr[n + 1];
  double xo, yo, zo, mm;
  mf_ld256(op.node4 + n * 4, xo, yo, zo, mm);
  const int mo = __double2loint(mm);
  for (int j = lo; j < hi; ++j) {
    double ow, S, L, iL, xj, yj, zj, mj;
    mf_ld256(op.inc + j, ow, S, L, iL);
    mf_ld256(op.node4 + (int64_t)__double2loint(ow) * 4, xj, yj, zj, mj);
    const MfCoef k = mf_coef(op, S, L, iL);
    const double t[3] = {(xj - xo) * iL, (yj - yo) * iL, (zj - zo) * iL};
    // own-end quadrant: [aI I + aT tt,  -c [t']x ; +c [t']x,  (bI + dI) I + (dT - bI) tt]
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const double tt = t[p] * t[q];
        const double d = (p == q) ? 1.0 : 0.0;
        double sk = 0.0;   // [t']x entry (p, q)
        if (p == 0 && q == 1) sk = -t[2];
        if (p == 0 && q == 2) sk = t[1];
        if (p == 1 && q == 0) sk = t[2];
        if (p == 1 && q == 2) sk = -t[0];
        if (p == 2 && q == 0) sk = -t[1];
        if (p == 2 && q == 1) sk = t[0];
        a[p][q] += k.aI * d + k.aT * tt;
        a[p][q + 3] += -k.c * sk;
        a[p + 3][q] += k.c * sk;
        a[p + 3][q + 3] += (k.bI + k.dI) * d + (k.dT - k.bI) * tt;
      }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int k = 0; k < 6; ++k)
      if (((mo >> i) & 1) || ((mo >> k) & 1)) a[i][k] = (i == k) ? 1.0 : 0.0;
  if (hi == lo)   // unconnected node: identity block, like a missing diagonal in the assembled path
#pragma unroll
    for (int i = 0; i < 6; ++i) a[i][i] = 1.0;
}

// lattice_surrogate.cu -- SURVEY section 8(f) row N4: the reduced-basis / surrogate Schur pipeline on the device.  sm_100a.
//
//   greedy reduced basis of Schur snapshots      greedy_algorithm.py:35-155     lat_greedy_basis, lat_basis_project,
//                                                                               lat_upper_solve
//   thin-plate-spline RBF fit / value / gradient utils_rbf.py:22-144            lat_rbf_fit, lat_rbf_eval
//   nearest-neighbour / 1-D linear look-up       lattice_sim.py:755-807,939-946 lat_alpha_lookup
//   S_q = reshape_F(basis @ alpha_q)             lattice_sim.py:921-978,1056-1082  lat_basis_prepare, lat_basis_expand
//
// The last one is the only genuine dense FP64 contraction of the whole hot path ([n_q x k] x [k x n_B^2], k = 3..40,
// n_B^2 = 1296..7056, n_q = every cell of the lattice): it runs on the FP64 tensor cores (mma.sync m8n8k4 = DMMA) and
// is bound by writing S (8 n_B^2 bytes per cell) for small k and by the DMMA pipe from k ~ 40 on.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace {

constexpr int RED_BLOCK = 256;

// fixed-order CTA sum / max (same value in every thread)
template <int BLOCK>
__device__ __forceinline__ double cta_sum(double v) {
  __shared__ double s[BLOCK / 32];
  __shared__ double tot;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();                       // s / tot may still be read from a previous call
  if (lane == 0) s[wid] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < BLOCK / 32; ++k) t += s[k];
    tot = t;
  }
  __syncthreads();
  return tot;
}
template <int BLOCK>
__device__ __forceinline__ double cta_max(double v) {
  __shared__ double s[BLOCK / 32];
  __shared__ double tot;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if (lane == 0) s[wid] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = s[0];
    for (int k = 1; k < BLOCK / 32; ++k) t = fmax(t, s[k]);
    tot = t;
  }
  __syncthreads();
  return tot;
}

// ---------------------------------------------------------------------------------------------------------------
// 1. greedy reduced basis.  D[s][:] = snapshot s (normalised), deflated in place; stat[s] = {max |D_s|, sum |D_s|}.
// ---------------------------------------------------------------------------------------------------------------
// One CTA per snapshot: 2-norm, D = snap / norm, stats (greedy_algorithm.py:98-105).
__global__ void __launch_bounds__(RED_BLOCK) k_greedy_prepare(const double* __restrict__ snaps, int64_t len, double* __restrict__ D,
                                                              double* __restrict__ norms, double* __restrict__ stat) {
  const int64_t s = blockIdx.x;
  const double* v = snaps + s * len;
  double a = 0.0;
  for (int64_t i = threadIdx.x; i < len; i += RED_BLOCK) a = fma(v[i], v[i], a);
  const double nrm = sqrt(cta_sum<RED_BLOCK>(a));
  double mx = 0.0, sm = 0.0;
  for (int64_t i = threadIdx.x; i < len; i += RED_BLOCK) {
    const double d = v[i] / nrm;
    D[s * len + i] = d;
    mx = fmax(mx, fabs(d));
    sm += fabs(d);
  }
  mx = cta_max<RED_BLOCK>(mx);
  sm = cta_sum<RED_BLOCK>(sm);
  if (threadIdx.x == 0) { norms[s] = nrm; stat[2 * s] = mx; stat[2 * s + 1] = sm; }
}
// One CTA: s_I = first arg max of the column inf-norms (:114-115), basis[count] = D[s_I] / |D[s_I]|_2 (:117);
// scal[0] = max_s sum|D_s| (the convergence measure of :105 / :121 as it stands BEFORE this step), scal[1] = s_I.
__global__ void __launch_bounds__(RED_BLOCK) k_greedy_pick(const double* __restrict__ D, const double* __restrict__ stat, int n_snap,
                                                           int64_t len, double* __restrict__ newvec, int32_t* __restrict__ mainelem,
                                                           double* __restrict__ scal) {
  __shared__ int s_pick;
  if (threadIdx.x == 0) {
    int best = 0;
    double bm = stat[0], one = stat[1];
    for (int s = 1; s < n_snap; ++s) {
      if (stat[2 * s] > bm) { bm = stat[2 * s]; best = s; }
      one = fmax(one, stat[2 * s + 1]);
    }
    s_pick = best;
    scal[0] = one;
    scal[1] = (double)best;
    *mainelem = best;
  }
  __syncthreads();
  const double* v = D + (int64_t)s_pick * len;
  double a = 0.0;
  for (int64_t i = threadIdx.x; i < len; i += RED_BLOCK) a = fma(v[i], v[i], a);
  const double nrm = sqrt(cta_sum<RED_BLOCK>(a));
  for (int64_t i = threadIdx.x; i < len; i += RED_BLOCK) newvec[i] = v[i] / nrm;
}
// One CTA per snapshot: c_s = D_s . newvec (dgemv, :118); D_s -= c_s newvec (dger, :119); new stats.
__global__ void __launch_bounds__(RED_BLOCK) k_greedy_deflate(double* __restrict__ D, const double* __restrict__ newvec, int64_t len,
                                                              double* __restrict__ coef_row, double* __restrict__ stat) {
  const int64_t s = blockIdx.x;
  double* v = D + s * len;
  double a = 0.0;
  for (int64_t i = threadIdx.x; i < len; i += RED_BLOCK) a = fma(v[i], newvec[i], a);
  const double c = cta_sum<RED_BLOCK>(a);
  double mx = 0.0, sm = 0.0;
  for (int64_t i = threadIdx.x; i < len; i += RED_BLOCK) {
    const double d = fma(-c, newvec[i], v[i]);
    v[i] = d;
    mx = fmax(mx, fabs(d));
    sm += fabs(d);
  }
  mx = cta_max<RED_BLOCK>(mx);
  sm = cta_sum<RED_BLOCK>(sm);
  if (threadIdx.x == 0) { coef_row[s] = c; stat[2 * s] = mx; stat[2 * s + 1] = sm; }
}
__global__ void k_max_sumabs(const double* __restrict__ stat, int n_snap, double* __restrict__ out) {
  double one = 0.0;
  for (int s = 0; s < n_snap; ++s) one = fmax(one, stat[2 * s + 1]);
  *out = one;
}

// out[i][j] = A_i . B_j over len entries: one CTA per (i, j) (Gram matrix and right-hand sides of the projection).
__global__ void __launch_bounds__(RED_BLOCK) k_dots(const double* __restrict__ A, const double* __restrict__ B, int64_t len, int nb,
                                                    double* __restrict__ out, int ldo) {
  const double* a = A + (int64_t)blockIdx.y * len;
  const double* b = B + (int64_t)blockIdx.x * len;
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < len; i += RED_BLOCK) s = fma(a[i], b[i], s);
  s = cta_sum<RED_BLOCK>(s);
  if (threadIdx.x == 0) out[(int64_t)blockIdx.y * ldo + blockIdx.x] = s;
}

__global__ void k_transpose(const double* __restrict__ in, int rows, int cols, double* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)rows * cols) return;
  const int r = (int)(t / cols), c = (int)(t % cols);
  out[(size_t)c * rows + r] = in[t];
}

// ---------------------------------------------------------------------------------------------------------------
// small dense solvers, ONE CTA each (set-up algebra: k x k Gram systems, the (N+d+1)^2 TPS system)
// ---------------------------------------------------------------------------------------------------------------
// SPD solve G X = R in place (Cholesky, right looking).  G [n][n] row-major, R [n][m] row-major.
__global__ void __launch_bounds__(1024) k_chol_solve(double* __restrict__ G, int n, double* __restrict__ R, int m, int* __restrict__ fail) {
  __shared__ double piv, floor_;
  if (threadIdx.x == 0) {                      // pivots below 1e-13 of the largest diagonal entry = numerically rank deficient
    double dmax = 0.0;
    for (int j = 0; j < n; ++j) dmax = fmax(dmax, G[(size_t)j * n + j]);
    floor_ = 1e-13 * dmax;
  }
  __syncthreads();
  for (int j = 0; j < n; ++j) {
    if (threadIdx.x == 0) {
      const double d = G[(size_t)j * n + j];
      if (!(d > floor_)) *fail = j + 1;
      piv = sqrt(d > floor_ ? d : 1.0);
    }
    __syncthreads();
    const double p = piv;
    for (int i = j + threadIdx.x; i < n; i += blockDim.x) G[(size_t)i * n + j] /= p;      // column j of L (incl. diagonal)
    __syncthreads();
    const int rem = n - j - 1;
    for (int64_t t = threadIdx.x; t < (int64_t)rem * rem; t += blockDim.x) {
      const int i = j + 1 + (int)(t / rem), c = j + 1 + (int)(t % rem);
      if (c <= i) G[(size_t)i * n + c] = fma(-G[(size_t)i * n + j], G[(size_t)c * n + j], G[(size_t)i * n + c]);
    }
    __syncthreads();
  }
  // forward L y = R, backward L^T x = y; one thread per right-hand side
  for (int c = threadIdx.x; c < m; c += blockDim.x) {
    for (int i = 0; i < n; ++i) {
      double s = R[(size_t)i * m + c];
      for (int k = 0; k < i; ++k) s = fma(-G[(size_t)i * n + k], R[(size_t)k * m + c], s);
      R[(size_t)i * m + c] = s / G[(size_t)i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
      double s = R[(size_t)i * m + c];
      for (int k = i + 1; k < n; ++k) s = fma(-G[(size_t)k * n + i], R[(size_t)k * m + c], s);
      R[(size_t)i * m + c] = s / G[(size_t)i * n + i];
    }
  }
}
// Upper-triangular solve U X = R in place (dtrtrs of greedy_algorithm.py:129); U [n][ldu] row-major (upper part used).
__global__ void k_upper_solve(const double* __restrict__ U, int n, int ldu, double* __restrict__ R, int m) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m) return;
  for (int i = n - 1; i >= 0; --i) {
    double s = R[(size_t)i * m + c];
    for (int k = i + 1; k < n; ++k) s = fma(-U[(size_t)i * ldu + k], R[(size_t)k * m + c], s);
    R[(size_t)i * m + c] = s / U[(size_t)i * ldu + i];
  }
}
// General solve A X = B in place by LU with partial pivoting (np.linalg.solve of utils_rbf.py:58).  A [n][n], B [n][m].
__global__ void __launch_bounds__(1024) k_lu_solve(double* __restrict__ A, int n, double* __restrict__ B, int m, int* __restrict__ fail) {
  __shared__ double s_val[32];
  __shared__ int s_idx[32];
  __shared__ int s_p;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int j = 0; j < n; ++j) {
    // pivot: first row of maximal |A[i][j]|, i >= j
    double bv = -1.0;
    int bi = n;
    for (int i = j + threadIdx.x; i < n; i += blockDim.x) {
      const double v = fabs(A[(size_t)i * n + j]);
      if (v > bv) { bv = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { s_val[wid] = bv; s_idx[wid] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int k = 1; k < nw; ++k)
        if (s_val[k] > bv || (s_val[k] == bv && s_idx[k] < bi)) { bv = s_val[k]; bi = s_idx[k]; }
      if (!(bv > 0.0)) { *fail = j + 1; bi = j; }
      s_p = bi;
    }
    __syncthreads();
    const int p = s_p;
    if (p != j) {
      for (int c = threadIdx.x; c < n + m; c += blockDim.x) {
        double* a = c < n ? &A[(size_t)j * n + c] : &B[(size_t)j * m + (c - n)];
        double* b = c < n ? &A[(size_t)p * n + c] : &B[(size_t)p * m + (c - n)];
        const double t = *a; *a = *b; *b = t;
      }
    }
    __syncthreads();
    const double d = A[(size_t)j * n + j];
    if (d != 0.0) {
      const int rem = n - j - 1, wcols = rem + m;        // trailing columns of A, then all of B
      for (int i = j + 1 + threadIdx.x; i < n; i += blockDim.x) A[(size_t)i * n + j] /= d;
      __syncthreads();
      for (int64_t t = threadIdx.x; t < (int64_t)rem * wcols; t += blockDim.x) {
        const int i = j + 1 + (int)(t / wcols), cc = (int)(t % wcols);
        const double l = A[(size_t)i * n + j];
        if (cc < rem) A[(size_t)i * n + j + 1 + cc] = fma(-l, A[(size_t)j * n + j + 1 + cc], A[(size_t)i * n + j + 1 + cc]);
        else B[(size_t)i * m + (cc - rem)] = fma(-l, B[(size_t)j * m + (cc - rem)], B[(size_t)i * m + (cc - rem)]);
      }
    }
    __syncthreads();
  }
  for (int c = threadIdx.x; c < m; c += blockDim.x)
    for (int i = n - 1; i >= 0; --i) {
      double s = B[(size_t)i * m + c];
      for (int k = i + 1; k < n; ++k) s = fma(-A[(size_t)i * n + k], B[(size_t)k * m + c], s);
      B[(size_t)i * m + c] = s / A[(size_t)i * n + i];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// 2. thin-plate-spline RBF
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double tps_dist(const double* __restrict__ a, const double* __restrict__ b, int d) {
  double r2 = 0.0;
  for (int k = 0; k < d; ++k) { const double t = a[k] - b[k]; r2 = fma(t, t, r2); }
  return sqrt(r2);
}
// A = [[Phi + reg I, P], [P^T, 0]], rhs = [Y; 0]   (utils_rbf.py:43-56)
__global__ void k_tps_system(const double* __restrict__ X, int N, int d, double reg, const double* __restrict__ Y, int m,
                             double* __restrict__ A, double* __restrict__ rhs) {
  const int n = N + d + 1;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < (int64_t)n * n; t += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(t / n), j = (int)(t % n);
    double v = 0.0;
    if (i < N && j < N) {
      const double r = tps_dist(X + (size_t)i * d, X + (size_t)j * d, d);
      v = r > 0.0 ? r * r * log(r) : 0.0;
      if (i == j) v += reg;
    } else if (i < N) v = (j == N) ? 1.0 : X[(size_t)i * d + (j - N - 1)];
    else if (j < N) v = (i == N) ? 1.0 : X[(size_t)j * d + (i - N - 1)];
    A[t] = v;
  }
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < (int64_t)n * m; t += (int64_t)gridDim.x * blockDim.x)
    rhs[t] = (t / m) < N ? Y[t] : 0.0;
}
// One CTA per 8 queries: the kernel values phi(r) (and (2 log r + 1)(x - x_i) for the gradient) of a chunk of 128
// centres go through shared memory once and are shared by all m outputs.   f [M][m]; grad [M][d][m].
constexpr int TPS_Q = 8, TPS_C = 128, TPS_DMAX = 8;
__global__ void __launch_bounds__(256) k_tps_eval(const double* __restrict__ X, int N, int d, const double* __restrict__ wcp, int m,
                                                  const double* __restrict__ xq, int64_t M, double* __restrict__ f,
                                                  double* __restrict__ grad) {
  extern __shared__ double s_dyn[];
  double* s_phi = s_dyn;                                // [TPS_Q][TPS_C]
  double* s_g = s_dyn + TPS_Q * TPS_C;                  // [TPS_Q][d][TPS_C]  (gradient only)
  const int64_t q0 = (int64_t)blockIdx.x * TPS_Q;
  const int nq = (int)(M - q0 < TPS_Q ? M - q0 : TPS_Q);
  const int n_out = m * (grad ? 1 + d : 1);             // outputs per query handled by the thread loop
  // accumulators: thread t owns outputs t, t + 256, ... of the (query, output) grid -- at most 4 per thread kept in
  // registers; larger problems loop over output tiles.
  for (int o0 = 0; o0 < nq * n_out; o0 += 256 * 4) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int c0 = 0; c0 < N; c0 += TPS_C) {
      const int nc = N - c0 < TPS_C ? N - c0 : TPS_C;
      __syncthreads();
      for (int t = threadIdx.x; t < nq * nc; t += 256) {
        const int q = t / nc, c = t - q * nc;
        const double* xc = X + (size_t)(c0 + c) * d;
        const double* xx = xq + (q0 + q) * d;
        const double r = tps_dist(xx, xc, d);
        const double lg = r > 0.0 ? log(r) : 0.0;
        s_phi[q * TPS_C + c] = r > 0.0 ? r * r * lg : 0.0;
        if (grad) {
          const double fac = r > 0.0 ? 2.0 * lg + 1.0 : 0.0;
          for (int k = 0; k < d; ++k) s_g[(q * d + k) * TPS_C + c] = fac * (xx[k] - xc[k]);
        }
      }
      __syncthreads();
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int o = o0 + u * 256 + threadIdx.x;
        if (o >= nq * n_out) continue;
        const int q = o / n_out, w = o - q * n_out;       // w < m: value j = w;  else gradient (k, j)
        const int j = w % m, k = w / m - 1;
        const double* row = (k < 0) ? s_phi + q * TPS_C : s_g + (q * d + k) * TPS_C;
        double a = acc[u];
        for (int c = 0; c < nc; ++c) a = fma(row[c], __ldg(wcp + (size_t)(c0 + c) * m + j), a);
        acc[u] = a;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int o = o0 + u * 256 + threadIdx.x;
      if (o >= nq * n_out) continue;
      const int q = o / n_out, w = o - q * n_out;
      const int j = w % m, k = w / m - 1;
      const double* xx = xq + (q0 + q) * d;
      if (k < 0) {
        double a = acc[u] + wcp[(size_t)N * m + j];                          // polynomial tail c0 + c . x
        for (int e = 0; e < d; ++e) a = fma(xx[e], wcp[(size_t)(N + 1 + e) * m + j], a);
        if (f) f[(q0 + q) * m + j] = a;
      } else {
        grad[((q0 + q) * d + k) * m + j] = acc[u] + wcp[(size_t)(N + 1 + k) * m + j];
      }
    }
  }
}
// mode 0: nearest centre (first minimum); mode 1: 1-D piecewise linear, clamped outside (np.interp), centres in any order.
__global__ void k_alpha_lookup(int mode, const double* __restrict__ X, int N, int d, const double* __restrict__ alpha, int m,
                               const double* __restrict__ xq, int64_t M, double* __restrict__ out) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= M) return;
  const double* xx = xq + q * d;
  if (mode == 0) {
    int best = 0;
    double bd = INFINITY;
    for (int i = 0; i < N; ++i) {
      const double r = tps_dist(xx, X + (size_t)i * d, d);
      if (r < bd) { bd = r; best = i; }
    }
    for (int j = 0; j < m; ++j) out[q * m + j] = alpha[(size_t)best * m + j];
    return;
  }
  // left neighbour = largest centre <= x, right = smallest centre > x  (np.interp on the sorted centres)
  const double x = xx[0];
  int lo = -1, hi = -1, mn = 0, mx = 0;
  for (int i = 0; i < N; ++i) {
    const double c = X[i];
    if (c < X[mn]) mn = i;
    if (c > X[mx]) mx = i;
    if (c <= x && (lo < 0 || c > X[lo])) lo = i;
    if (c > x && (hi < 0 || c < X[hi])) hi = i;
  }
  for (int j = 0; j < m; ++j) {
    double v;
    if (lo < 0) v = alpha[(size_t)mn * m + j];
    else if (hi < 0) v = alpha[(size_t)mx * m + j];
    else {
      const double x0 = X[lo], x1 = X[hi], y0 = alpha[(size_t)lo * m + j], y1 = alpha[(size_t)hi * m + j];
      v = (x == x0) ? y0 : __dadd_rn(__dmul_rn((y1 - y0) / (x1 - x0), x - x0), y0);     // numpy: slope * (x - x0) + y0, unfused
    }
    out[q * m + j] = v;
  }
}

// N-parameter "linear" surrogate: piecewise-linear interpolation on a Delaunay triangulation of the centres, nearest
// centre outside the convex hull (scipy LinearNDInterpolator / NearestNDInterpolator, lattice_sim.py:794-807).  The
// triangulation is set-up geometry (built once on the host by the same Qhull call scipy makes); the kernel locates the
// simplex of every query with the triangulation's own barycentric transforms: b = T_s (x - r_s), b_d = 1 - sum b.
// simplices: int32 [ns][d+1]; transform: [ns][d+1][d] (rows 0..d-1 = T_s, row d = r_s), as scipy.spatial.Delaunay stores it.
__global__ void k_alpha_simplex(const int32_t* __restrict__ simplices, const double* __restrict__ transform, int ns, int d,
                                const double* __restrict__ X, int N, const double* __restrict__ alpha, int m,
                                const double* __restrict__ xq, int64_t M, double* __restrict__ out) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= M) return;
  const double* xx = xq + q * d;
  const double eps = 100.0 * 2.220446049250313e-16;          // Qhull's inside test (scipy _qhull: 100 * DBL_EPSILON)
  double bary[TPS_DMAX + 1];
  int found = -1;
  for (int s = 0; s < ns && found < 0; ++s) {
    const double* T = transform + (size_t)s * (d + 1) * d;
    double last = 1.0;
    bool inside = true;
    for (int i = 0; i < d; ++i) {
      double b = 0.0;
      for (int j = 0; j < d; ++j) b = fma(T[i * d + j], xx[j] - T[d * d + j], b);
      bary[i] = b;
      last -= b;
      inside = inside && (b >= -eps) && (b <= 1.0 + eps);     // NaN transforms (degenerate simplices) fail both tests
    }
    bary[d] = last;
    inside = inside && (last >= -eps) && (last <= 1.0 + eps);
    if (inside) found = s;
  }
  if (found >= 0) {
    const int32_t* vtx = simplices + (size_t)found * (d + 1);
    for (int j = 0; j < m; ++j) {
      double v = 0.0;
      for (int i = 0; i <= d; ++i) v = fma(bary[i], alpha[(size_t)vtx[i] * m + j], v);
      out[q * m + j] = v;
    }
    return;
  }
  int best = 0;
  double bd = INFINITY;
  for (int i = 0; i < N; ++i) {
    const double r = tps_dist(xx, X + (size_t)i * d, d);
    if (r < bd) { bd = r; best = i; }
  }
  for (int j = 0; j < m; ++j) out[q * m + j] = alpha[(size_t)best * m + j];
}

// ---------------------------------------------------------------------------------------------------------------
// 3. S = basis @ alpha on the FP64 tensor cores
// ---------------------------------------------------------------------------------------------------------------
// basisP[kk][a * n + b] = basis[(a + n b)][kk]  (the reference reshapes basis @ alpha in FORTRAN order,
// lattice_sim.py:973-976; n = 0: no permutation); rows k..Kp-1 are zero.
__global__ void k_basis_prepare(const double* __restrict__ basis, int64_t len, int k, int n, int Kp, double* __restrict__ basisP) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)Kp * len) return;
  const int kk = (int)(t / len);
  const int64_t o = t - (int64_t)kk * len;
  int64_t src = o;
  if (n > 0) { const int64_t a = o / n, b = o - a * n; src = a + (int64_t)n * b; }
  basisP[t] = kk < k ? basis[src * k + kk] : 0.0;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void st256(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// out[q][l] (+)= sum_kk alphas[q][k0 + kk] basisP[k0 + kk][l].  One warp = 8 MT rows (MT 8-row DMMA tiles) whose
// coefficients stay in registers in A-fragment layout (lane (r8, c4) holds alpha[q0 + 8 mt + r8][4 kk + c4]); the warp
// walks its column range in groups of 16: B fragments straight from basisP through L1 (the eight warps of a CTA walk
// the same columns), two 8-column tiles per group with the columns interleaved so that every lane ends up with FOUR
// CONSECUTIVE outputs of one row -> one 256-bit store per tile row, 128 contiguous bytes per row and instruction.
//   tile t in {0,1}, tile column n  <->  output column l0 + 4 (n / 2) + 2 t + (n % 2)
template <int KT, int MT, bool ACCUM>
__global__ void __launch_bounds__(256) k_basis_expand(const double* __restrict__ basisP, const double* __restrict__ alphas, int64_t M,
                                                      int k, int lda, int k0, int Kp, int64_t L, int cols_per_cta,
                                                      double* __restrict__ out) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int r8 = lane >> 2, c4 = lane & 3;
  const int64_t q0 = ((int64_t)blockIdx.x * 8 + wid) * (8 * MT);
  if (q0 >= M) return;
  double a[MT][KT];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int64_t q = q0 + 8 * mt + r8;
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) {
      const int kc = k0 + 4 * kk + c4;
      a[mt][kk] = (q < M && kc < k) ? __ldg(alphas + q * lda + kc) : 0.0;
    }
  }
  const int64_t l_begin = (int64_t)blockIdx.y * cols_per_cta;
  const int64_t l_end = l_begin + cols_per_cta < L ? l_begin + cols_per_cta : L;
  const int bcol = 4 * (r8 >> 1) + (r8 & 1);
  // B fragments of a 16-column group: 2 KT loads straight from the (L2 / L1 resident) basis.  Up to 10 k-steps the NEXT
  // group's fragments are fetched before the current group's DMMAs are issued (software pipelining: with one CTA per
  // SM at these register counts nothing else hides the load latency: 42 % long-scoreboard stalls without it).
  constexpr bool PREFETCH = KT <= 10;
  double bn[PREFETCH ? KT : 1][2];
  auto load_b = [&](int64_t l0, double (&bb)[PREFETCH ? KT : 1][2]) {
    const bool bval = l0 + bcol < L && l0 < l_end;         // L % 4 == 0: the +2 column is valid with it
    const double* bp = basisP + (size_t)(k0 + c4) * L + l0 + bcol;
#pragma unroll
    for (int kk = 0; kk < (PREFETCH ? KT : 1); ++kk) {
      const bool bk = bval && k0 + 4 * kk + c4 < Kp;        // KT may be rounded up past the padded basis
      bb[kk][0] = bk ? __ldg(bp + (size_t)(4 * kk) * L) : 0.0;
      bb[kk][1] = bk ? __ldg(bp + (size_t)(4 * kk) * L + 2) : 0.0;
    }
  };
  if (PREFETCH) load_b(l_begin, bn);
  for (int64_t l0 = l_begin; l0 < l_end; l0 += 16) {
    double c[MT][2][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) { c[mt][0][0] = c[mt][0][1] = c[mt][1][0] = c[mt][1][1] = 0.0; }
    if (PREFETCH) {
      double b[PREFETCH ? KT : 1][2];
#pragma unroll
      for (int kk = 0; kk < (PREFETCH ? KT : 1); ++kk) { b[kk][0] = bn[kk][0]; b[kk][1] = bn[kk][1]; }
      load_b(l0 + 16, bn);
#pragma unroll
      for (int kk = 0; kk < (PREFETCH ? KT : 1); ++kk) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          dmma884(c[mt][0][0], c[mt][0][1], a[mt][kk], b[kk][0]);
          dmma884(c[mt][1][0], c[mt][1][1], a[mt][kk], b[kk][1]);
        }
      }
    } else {
      const bool bval = l0 + bcol < L;
      const double* bp = basisP + (size_t)(k0 + c4) * L + l0 + bcol;
#pragma unroll
      for (int kk = 0; kk < KT; ++kk) {
        const bool bk = bval && k0 + 4 * kk + c4 < Kp;
        const double b0 = bk ? __ldg(bp + (size_t)(4 * kk) * L) : 0.0;
        const double b1 = bk ? __ldg(bp + (size_t)(4 * kk) * L + 2) : 0.0;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          dmma884(c[mt][0][0], c[mt][0][1], a[mt][kk], b0);
          dmma884(c[mt][1][0], c[mt][1][1], a[mt][kk], b1);
        }
      }
    }
    const int64_t lc = l0 + 4 * c4;
    if (lc < L) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int64_t q = q0 + 8 * mt + r8;
        if (q < M) {
          double* o = out + q * L + lc;
          if (ACCUM) {
            const double4 old = *reinterpret_cast<const double4*>(o);
            st256(o, old.x + c[mt][0][0], old.y + c[mt][0][1], old.z + c[mt][1][0], old.w + c[mt][1][1]);
          } else st256(o, c[mt][0][0], c[mt][0][1], c[mt][1][0], c[mt][1][1]);
        }
      }
    }
  }
}

template <int KT, int MT>
int launch_expand(lat_ctx* ctx, const double* basisP, const double* alphas, int64_t M, int k, int lda, int k0, int Kp, int64_t L,
                  double* out) {
  // column split: enough CTAs for two waves when there are few queries, otherwise long column runs per CTA
  const int64_t gx = ceil_div(M, 64 * MT);
  int cols = 1024;                                    // a multiple of 16 at every halving
  while (cols > 16 && gx * ceil_div(L, cols) < 2 * ctx->sm_count) cols /= 2;
  const dim3 grid((unsigned)gx, (unsigned)ceil_div(L, cols));
  if (k0 == 0) LAT_LAUNCH(ctx, (k_basis_expand<KT, MT, false>), grid, 256, 0, basisP, alphas, M, k, lda, k0, Kp, L, cols, out);
  else LAT_LAUNCH(ctx, (k_basis_expand<KT, MT, true>), grid, 256, 0, basisP, alphas, M, k, lda, k0, Kp, L, cols, out);
  return LAT_OK;
}
// Rows per warp: 32 (four tiles share every B fragment).  16 rows per warp (half the registers, three CTAs per SM instead
// of one at 10 k-steps) was measured and is SLOWER from 5 k-steps on (k = 38: 1.71 vs 1.41 ms): the kernel is not
// occupancy bound, the B fragments are (profiles/r02_surrogate_ab.txt).  LAT_EXPAND_MT=2 selects it for experiments.
template <int KT>
int launch_expand_kt(lat_ctx* ctx, const double* basisP, const double* alphas, int64_t M, int k, int lda, int k0, int Kp, int64_t L,
                     double* out) {
  const char* env = getenv("LAT_EXPAND_MT");
  const int mt = env ? atoi(env) : 4;
  if (mt == 4) return launch_expand<KT, 4>(ctx, basisP, alphas, M, k, lda, k0, Kp, L, out);
  return launch_expand<KT, 2>(ctx, basisP, alphas, M, k, lda, k0, Kp, L, out);
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------
extern "C" int lat_greedy_basis(lat_ctx* ctx, const double* snaps, int64_t n_snap, int64_t len, double tol, double* basis,
                                double* coef, int32_t* mainelem, double* norms, int32_t* k_out) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, snaps && basis && coef && mainelem && norms && k_out && n_snap > 0 && n_snap < (1 << 30) && len > 0);
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  double* D = lat_buf<double>(ctx, "gr_D", (size_t)n_snap * len);
  double* stat = lat_buf<double>(ctx, "gr_stat", (size_t)n_snap * 2);
  double* scal = lat_buf<double>(ctx, "gr_scal", 4);
  if (!D || !stat || !scal) return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
  LAT_LAUNCH(ctx, k_greedy_prepare, (unsigned)n_snap, RED_BLOCK, 0, snaps, len, D, norms, stat);
  double h[2] = {0.0, 0.0};
  double atol = 0.0;
  int count = 0;
  bool cvg = false;
  while (!cvg && count < n_snap) {                                                   // greedy_algorithm.py:112
    LAT_LAUNCH(ctx, k_greedy_pick, 1, RED_BLOCK, 0, D, stat, (int)n_snap, len, basis + (size_t)count * len, mainelem + count, scal);
    LAT_LAUNCH(ctx, k_greedy_deflate, (unsigned)n_snap, RED_BLOCK, 0, D, basis + (size_t)count * len, len,
               coef + (size_t)count * n_snap, stat);
    LAT_LAUNCH(ctx, k_max_sumabs, 1, 1, 0, stat, (int)n_snap, scal + 2);
    LAT_CUDA(ctx, cudaMemcpyAsync(h, scal, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    LAT_CUDA(ctx, cudaMemcpyAsync(h + 1, scal + 2, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    LAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (count == 0) atol = tol * h[0];                                               // :105 (measure before the first step)
    ++count;
    cvg = h[1] < atol;                                                               // :121
  }
  *k_out = count;
  return LAT_OK;
}

extern "C" int lat_upper_solve(lat_ctx* ctx, const double* U, int32_t n, int32_t ldu, double* R, int32_t m) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, U && R && n > 0 && ldu >= n && m > 0);
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  LAT_LAUNCH(ctx, k_upper_solve, (unsigned)ceil_div(m, 128), 128, 0, U, n, ldu, R, m);
  return LAT_OK;
}

extern "C" int lat_basis_project(lat_ctx* ctx, const double* basis, int32_t k, int64_t len, const double* V, int64_t n,
                                 double* alphas) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, basis && V && alphas && k > 0 && k <= 4096 && len > 0 && n > 0 && n < 65536);
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  double* G = lat_buf<double>(ctx, "gr_gram", (size_t)k * k);
  double* R = lat_buf<double>(ctx, "gr_rhs", (size_t)k * n);
  int* fail = lat_buf<int>(ctx, "gr_fail", 1);
  if (!G || !R || !fail) return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
  LAT_CUDA(ctx, cudaMemsetAsync(fail, 0, sizeof(int), ctx->stream));
  LAT_LAUNCH(ctx, k_dots, dim3((unsigned)k, (unsigned)k), RED_BLOCK, 0, basis, basis, len, k, G, k);
  LAT_LAUNCH(ctx, k_dots, dim3((unsigned)n, (unsigned)k), RED_BLOCK, 0, basis, V, len, (int)n, R, (int)n);     // R[i][s] = B_i . V_s
  LAT_LAUNCH(ctx, k_chol_solve, 1, 1024, 0, G, k, R, (int)n, fail);
  LAT_LAUNCH(ctx, k_transpose, (unsigned)ceil_div((int64_t)k * n, 256), 256, 0, R, k, (int)n, alphas);      // alphas [n][k] = R^T
  int hfail = 0;
  LAT_CUDA(ctx, cudaMemcpyAsync(&hfail, fail, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  LAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (hfail) return lat_fail(ctx, LAT_ERR_ARG, "basis is rank deficient (Gram matrix not positive definite)", __FILE__, __LINE__);
  return LAT_OK;
}

extern "C" int lat_rbf_fit(lat_ctx* ctx, const double* x_train, int32_t N, int32_t d, const double* y, int32_t m, double reg,
                           double* wcp) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, x_train && y && wcp && N > 0 && d > 0 && d <= TPS_DMAX && m > 0 && N + d + 1 <= 4096);     // one-CTA LU
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  const int n = N + d + 1;
  double* A = lat_buf<double>(ctx, "rbf_A", (size_t)n * n);
  int* fail = lat_buf<int>(ctx, "gr_fail", 1);
  if (!A || !fail) return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
  LAT_CUDA(ctx, cudaMemsetAsync(fail, 0, sizeof(int), ctx->stream));
  LAT_LAUNCH(ctx, k_tps_system, (unsigned)ceil_div((int64_t)n * n, 256), 256, 0, x_train, N, d, reg, y, m, A, wcp);
  LAT_LAUNCH(ctx, k_lu_solve, 1, 1024, 0, A, n, wcp, m, fail);
  int hfail = 0;
  LAT_CUDA(ctx, cudaMemcpyAsync(&hfail, fail, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  LAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (hfail) return lat_fail(ctx, LAT_ERR_ARG, "singular thin-plate-spline system (duplicate centres?)", __FILE__, __LINE__);
  return LAT_OK;
}

extern "C" int lat_rbf_eval(lat_ctx* ctx, const double* x_train, int32_t N, int32_t d, const double* wcp, int32_t m,
                            const double* xq, int64_t M, double* f, double* grad) {
  if (!ctx) return LAT_ERR_ARG;
  if (M == 0) return LAT_OK;
  LAT_CHECK_ARG(ctx, x_train && wcp && xq && (f || grad) && N > 0 && d > 0 && d <= TPS_DMAX && m > 0 && M > 0);
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t smem = sizeof(double) * TPS_Q * TPS_C * (grad ? 1 + d : 1);
  LAT_CUDA(ctx, cudaFuncSetAttribute(k_tps_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * TPS_Q * TPS_C * (1 + TPS_DMAX))));
  LAT_LAUNCH(ctx, k_tps_eval, (unsigned)ceil_div(M, TPS_Q), 256, smem, x_train, N, d, wcp, m, xq, M, f, grad);
  return LAT_OK;
}

extern "C" int lat_alpha_lookup(lat_ctx* ctx, int32_t mode, const double* x_train, int32_t N, int32_t d, const double* alpha_train,
                                int32_t m, const double* xq, int64_t M, double* out) {
  if (!ctx) return LAT_ERR_ARG;
  if (mode == 1 && d != 1)
    return lat_fail(ctx, LAT_ERR_UNSUPPORTED, "the linear surrogate is implemented for one parameter (np.interp branch, lattice_sim.py:781-792)",
                    __FILE__, __LINE__);
  if (M == 0) return LAT_OK;
  LAT_CHECK_ARG(ctx, x_train && alpha_train && xq && out && N > 0 && d > 0 && m > 0 && M > 0 && (mode == 0 || mode == 1));
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  LAT_LAUNCH(ctx, k_alpha_lookup, (unsigned)ceil_div(M, 128), 128, 0, mode, x_train, N, d, alpha_train, m, xq, M, out);
  return LAT_OK;
}

extern "C" int lat_alpha_simplex(lat_ctx* ctx, const int32_t* simplices, const double* transform, int32_t n_simplices, int32_t d,
                                 const double* x_train, int32_t N, const double* alpha_train, int32_t m, const double* xq,
                                 int64_t M, double* out) {
  if (!ctx) return LAT_ERR_ARG;
  if (M == 0) return LAT_OK;
  LAT_CHECK_ARG(ctx, simplices && transform && x_train && alpha_train && xq && out && n_simplices > 0 && d > 0 && d <= TPS_DMAX &&
                         N > 0 && m > 0 && M > 0);
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  LAT_LAUNCH(ctx, k_alpha_simplex, (unsigned)ceil_div(M, 128), 128, 0, simplices, transform, n_simplices, d, x_train, N,
             alpha_train, m, xq, M, out);
  return LAT_OK;
}

extern "C" int lat_basis_prepare(lat_ctx* ctx, const double* basis, int64_t len, int32_t k, int32_t n_fortran, double* basisP) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, basis && basisP && len > 0 && k > 0 && (n_fortran == 0 || (int64_t)n_fortran * n_fortran == len));
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  const int Kp = 4 * (int)ceil_div(k, 4);
  LAT_LAUNCH(ctx, k_basis_prepare, (unsigned)ceil_div((int64_t)Kp * len, 256), 256, 0, basis, len, k, n_fortran, Kp, basisP);
  return LAT_OK;
}

extern "C" int lat_basis_expand(lat_ctx* ctx, const double* basisP, int32_t k, int64_t len, const double* alphas, int64_t M,
                                int32_t lda, double* out) {
  if (!ctx) return LAT_ERR_ARG;
  if (M == 0) return LAT_OK;                                       // an empty batch has no buffers to check
  LAT_CHECK_ARG(ctx, basisP && alphas && out && k > 0 && lda >= k && M > 0);
  LAT_CHECK_ARG(ctx, len > 0 && len % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 31) == 0);      // 256-bit stores
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  const int Kp = 4 * (int)ceil_div(k, 4);
  for (int k0 = 0; k0 < Kp; k0 += 64) {
    const int kt = (Kp - k0 < 64 ? Kp - k0 : 64) / 4;
    int rc;
    switch (kt) {
      case 1: rc = launch_expand_kt<1>(ctx, basisP, alphas, M, k, lda, k0, Kp, len, out); break;
      case 2: rc = launch_expand_kt<2>(ctx, basisP, alphas, M, k, lda, k0, Kp, len, out); break;
      case 3: rc = launch_expand_kt<3>(ctx, basisP, alphas, M, k, lda, k0, Kp, len, out); break;
      case 4: rc = launch_expand_kt<4>(ctx, basisP, alphas, M, k, lda, k0, Kp, len, out); break;
      case 5: rc = launch_expand_kt<5>(ctx, basisP, alphas, M, k, lda, k0, Kp, len, out); break;
      case 6: rc = launch_expand_kt<6>(ctx, basisP, alphas, M, k, lda, k0, Kp, len, out); break;
      case 7: case 8: rc = launch_expand_kt<8>(ctx, basisP, alphas, M, k, lda, k0, Kp, len, out); break;
      case 9: case 10: rc = launch_expand_kt<10>(ctx, basisP, alphas, M, k, lda, k0, Kp, len, out); break;
      case 11: case 12: rc = launch_expand_kt<12>(ctx, basisP, alphas, M, k, lda, k0, Kp, len, out); break;
      default: rc = launch_expand_kt<16>(ctx, basisP, alphas, M, k, lda, k0, Kp, len, out); break;
    }
    if (rc != LAT_OK) return rc;
  }
  return LAT_OK;
}

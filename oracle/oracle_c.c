/* CPU ORACLE in C (test / baseline infrastructure, NOT product code).
 *
 * Plain-C restatement, with OpenMP, of the pieces of the pyLatticeSim hot path that the CPU baseline of
 * bench.py times on all host cores: element stiffness, value assembly into a CSR pattern, CSR mat-vec and
 * the reference's PCG.  Same pinning status as oracle/lattice_oracle.py (it is checked against that module,
 * which the reference's 30 stored Schur matrices pin): tests/test_oracle_c.py.
 * Only tests/, __graft_entry__.smoke() and bench.py may load this.
 *
 *   gcc -O3 -fopenmp -shared -fPIC -o oracle/_build/liboracle_c.so oracle/oracle_c.c -lm
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

/* K_e = L * sum_i D_i b_i b_i^T with the six strain vectors of simulation_base.py:141-156, the frame of
 * beam_model.py:197-216, the section constants of material_definition.py:142-156 and the degree-1 shear
 * quadrature of simulation_base.py:190-197,220-225.  Ke: [ne][12][12]. */
void orc_elem_stiffness(const double* xyz, const int* en, const double* rad, long ne, double E, double nu,
                        double kappa, double* Ke) {
  const double PI = 3.14159265358979323846;
  const double G = E / (2.0 * (1.0 + nu));
#pragma omp parallel for schedule(static)
  for (long e = 0; e < ne; ++e) {
    const double* x1 = xyz + 3 * (long)en[2 * e];
    const double* x2 = xyz + 3 * (long)en[2 * e + 1];
    double d[3] = {x2[0] - x1[0], x2[1] - x1[1], x2[2] - x1[2]};
    const double L = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]), iL = 1.0 / L;
    double t[3] = {d[0] * iL, d[1] * iL, d[2] * iL};
    double e1[3] = {1, 0, 0};
    if (fabs(t[1]) < fabs(t[0])) { e1[0] = 0; e1[1] = 1; }
    const double te1 = t[0] * e1[0] + t[1] * e1[1] + t[2] * e1[2];
    double e2[3] = {e1[0], e1[1], e1[2]};
    if (fabs(t[2]) < fabs(te1)) { e2[0] = 0; e2[1] = 0; e2[2] = 1; }
    double a1[3], a2[3];
    cross3(t, e2, a1);
    double n1 = 1.0 / sqrt(a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2]);
    a1[0] *= n1; a1[1] *= n1; a1[2] *= n1;
    cross3(t, a1, a2);
    double n2 = 1.0 / sqrt(a2[0] * a2[0] + a2[1] * a2[1] + a2[2] * a2[2]);
    a2[0] *= n2; a2[1] *= n2; a2[2] *= n2;
    const double r = rad[e], S = PI * r * r, I = PI * r * r * r * r / 4.0, J = 2.0 * I;
    const double D[6] = {E * S, G * kappa * S, G * kappa * S, G * J, E * I, E * I};
    double B[6][12];
    memset(B, 0, sizeof B);
    for (int k = 0; k < 3; ++k) {
      B[0][k] = -t[k] * iL;  B[0][6 + k] = t[k] * iL;
      B[1][k] = -a1[k] * iL; B[1][6 + k] = a1[k] * iL; B[1][3 + k] = -0.5 * a2[k]; B[1][9 + k] = -0.5 * a2[k];
      B[2][k] = -a2[k] * iL; B[2][6 + k] = a2[k] * iL; B[2][3 + k] = 0.5 * a1[k];  B[2][9 + k] = 0.5 * a1[k];
      B[3][3 + k] = -t[k] * iL;  B[3][9 + k] = t[k] * iL;
      B[4][3 + k] = -a1[k] * iL; B[4][9 + k] = a1[k] * iL;
      B[5][3 + k] = -a2[k] * iL; B[5][9 + k] = a2[k] * iL;
    }
    double* K = Ke + 144 * e;
    for (int a = 0; a < 12; ++a)
      for (int b = 0; b < 12; ++b) {
        double s = 0.0;
        for (int i = 0; i < 6; ++i) s += D[i] * B[i][a] * B[i][b];
        K[a * 12 + b] = L * s;
      }
  }
}

/* data[pos(i,j)] += Ke entries, pattern (indptr, indices sorted per row) given; data must be zeroed.
 * fem_petsc.assemble_matrix, simulation_base.py:480-481. */
void orc_assemble_csr(long ne, const int* en, const double* Ke, const int* indptr, const int* indices, double* data) {
#pragma omp parallel for schedule(static)
  for (long e = 0; e < ne; ++e) {
    int dof[12];
    for (int k = 0; k < 6; ++k) { dof[k] = 6 * en[2 * e] + k; dof[6 + k] = 6 * en[2 * e + 1] + k; }
    for (int a = 0; a < 12; ++a) {
      const int row = dof[a];
      int lo = indptr[row];
      const int hi = indptr[row + 1];
      for (int b = 0; b < 12; ++b) {
        int l = lo, h = hi;
        const int col = dof[b];
        while (l < h) { int m = (l + h) >> 1; if (indices[m] < col) l = m + 1; else h = m; }
#pragma omp atomic
        data[l] += Ke[144 * e + a * 12 + b];
      }
    }
  }
}

void orc_csr_matvec(long n, const int* indptr, const int* indices, const double* data, const double* x, double* y) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) {
    double s = 0.0;
    for (int k = indptr[i]; k < indptr[i + 1]; ++k) s += data[k] * x[indices[k]];
    y[i] = s;
  }
}

static double dotp(long n, const double* a, const double* b) {
  double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (long i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

/* conjugate_gradient_solver.py:59-122 with M = diag(dinv) (dinv == NULL: no preconditioner, z aliases r).
 * Returns info (0 converged, 1 maxiter, 2 alpha < 1e-6 seen); *iters = iterations performed. */
int orc_pcg(long n, const int* indptr, const int* indices, const double* data, const double* b, const double* dinv,
            double* x, int maxiter, double tol, double mintol, long restart_every, double alpha_max, int* iters) {
  double* r = (double*)malloc(sizeof(double) * n);
  double* z = dinv ? (double*)malloc(sizeof(double) * n) : r;
  double* p = (double*)malloc(sizeof(double) * n);
  double* Ap = (double*)malloc(sizeof(double) * n);
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) { x[i] = 0.0; r[i] = b[i]; if (dinv) z[i] = dinv[i] * b[i]; p[i] = dinv ? dinv[i] * b[i] : b[i]; }
  double rz_old = dotp(n, r, z);
  const double norm_b = sqrt(dotp(n, b, b));
  int info = 1, it = 0;
  for (int k = 0; k < maxiter; ++k) {
    it = k + 1;
    orc_csr_matvec(n, indptr, indices, data, p, Ap);
    double alpha = rz_old / dotp(n, p, Ap);
    if (alpha > alpha_max) alpha = alpha_max;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) { x[i] += alpha * p[i]; r[i] -= alpha * Ap[i]; }
    if (restart_every > 0 && k % restart_every == 0 && k > 0) memcpy(p, z, sizeof(double) * n);
    const double rn = sqrt(dotp(n, r, r)), dn = sqrt(dotp(n, p, p)), sn = sqrt(dotp(n, x, x));
    if (rn <= tol * norm_b) { info = 0; break; }
    if (dn < mintol * (sn + 1e-12)) { info = 0; break; }
    if (alpha < 1e-6) info = 2;
    if (dinv) {
#pragma omp parallel for schedule(static)
      for (long i = 0; i < n; ++i) z[i] = dinv[i] * r[i];
    }
    const double rz_new = dotp(n, r, z);
    const double beta = rz_new / rz_old;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
    rz_old = rz_new;
  }
  *iters = it;
  free(r);
  if (dinv) free(z);
  free(p);
  free(Ap);
  return info;
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm of bench.py sets the count explicitly. */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

// pcg_persist.cuh -- the whole PCG solve as ONE persistent cooperative kernel, vector state on chip.  sm_100a.
//
// Why: at BASELINE config 1 (98 MB matrix, 0.49 M DOF) the three-kernel iteration (k_cg_update, k_cg_spmv,
// k_cg_reduce) spends ~40 % of its 35 us outside the matrix stream: 56 MB of vector traffic per iteration, three
// launch ramps / tails of ~2-5 us each and a one-CTA reduce kernel (profiles/r01_ncu_full_v4_kernels.txt).  A B200 has
// 148 x 227 KB = 33 MB of shared memory, enough to hold x, r, p, s and w of a 0.5-0.7 M DOF system.  So:
//
//   * one CTA per SM (cooperative launch), each owning a CONTIGUOUS range of block rows balanced by bytes;
//   * x, r, p, s, w of the owned rows live in shared memory for the whole solve; the CTA's slice of rowptr and
//     colidx is cached in shared memory too, so the product phase issues only independent loads (matrix blocks,
//     gathered u) -- no rowptr -> colidx -> u dependency chain;
//   * only u = M^-1 r (gathered by other CTAs) lives in global memory, i.e. in L2;
//   * per iteration two grid-wide synchronisations, both flag based (no atomics on a shared counter):
//       barrier A  (after the update phase: the new u is visible)  -- one epoch flag per CTA, everyone polls all flags;
//       barrier B  (after the product phase: dot products)         -- every CTA publishes its three partial sums as
//                  six 8-byte {32 data bits | 32 epoch bits} words (the "LL" trick of the peer-memory all-reduce),
//                  every CTA polls all words and adds them in CTA order: barrier and all-reduce in ONE L2 round
//                  trip, identical bits in every CTA, so all CTAs take the same branch (converged / restart);
//   * the Chronopoulos-Gear recurrences, the stop test, the true-residual safeguard and its restart run inside the
//     kernel; the host launches once and reads the status block.
//
// HBM traffic per iteration: the matrix (+ block-Jacobi inverse); the vectors never leave the SM.  All spins are
// bounded: a CTA that waits longer than ~1 s reports info = 4 and every CTA leaves.
//
// Included by lattice_solver.cu after PcgScalars / PcgParams / dot6 / load_precond / apply_precond.
#pragma once

static constexpr int PERSIST_BLOCK = 1024;             // one CTA per SM
static constexpr int PERSIST_NW = PERSIST_BLOCK / 32;
static constexpr int PERSIST_NVEC = 5;                 // r, p, s, w, x in shared memory
static constexpr long long PERSIST_SPIN_LIMIT = 1ll << 23;

struct PersistArgs {
  const int32_t* rowptr;
  const int32_t* colidx;
  const double* vals;
  int64_t n_nodes;
  const double* b;
  double* x;               // out: solution (written at the end and before a residual check)
  double* u;               // global: preconditioned residual, gathered by every CTA
  const double* dinv;      // preconditioner (layout of k_precond_setup)
  PcgScalars* sc;          // status block read by the host
  PcgParams prm;
  unsigned long long* mail;   // [G][8] LL words (6 used), zeroed before the launch
  unsigned int* flags;        // [G]    barrier-A epochs, zeroed before the launch
  const int32_t* cta_row0;    // [G+1]  row partition
  int rows_cap, blk_cap;      // shared-memory capacities (rows, blocks) the launch was sized for
};

// Row partition balanced by bytes: cost(row) = 292 B per block + ROW_COST for the row's vector / preconditioner share.
static constexpr int64_t PERSIST_ROW_COST = 320;
__global__ void k_persist_partition(const int32_t* __restrict__ rowptr, int64_t n_nodes, int G,
                                    int32_t* __restrict__ cta_row0, int32_t* __restrict__ maxima) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > G) return;
  const int64_t total = 292ll * rowptr[n_nodes] + PERSIST_ROW_COST * n_nodes;
  const int64_t target = (total * c) / G;
  int64_t lo = 0, hi = n_nodes;   // first row r with cost_prefix(r) >= target
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (292ll * rowptr[mid] + PERSIST_ROW_COST * mid < target) lo = mid + 1; else hi = mid;
  }
  cta_row0[c] = (c == G) ? (int32_t)n_nodes : (int32_t)lo;
}
__global__ void k_persist_maxima(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ cta_row0, int G,
                                 int32_t* __restrict__ maxima) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= G) return;
  const int r0 = cta_row0[c], r1 = cta_row0[c + 1];
  atomicMax(&maxima[0], r1 - r0);
  atomicMax(&maxima[1], rowptr[r1] - rowptr[r0]);
}

__device__ __forceinline__ unsigned long long ld_relaxed_gpu_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_gpu_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_relaxed_gpu_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_gpu_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct PersistShared {
  double red[3][PERSIST_NW];
  double loc[3];
  double tot[3];
  int timeout;
};

// Barrier A: every global store of this CTA (the new u / x of its rows) is visible to every CTA afterwards.
// Same structure as a cooperative-groups grid sync, with one flag per CTA instead of a shared counter.
__device__ __forceinline__ void persist_barrier(unsigned int* flags, unsigned int epoch, int G, PersistShared& sh) {
  __syncthreads();                                   // the CTA's stores are issued ...
  if (threadIdx.x == 0) {
    __threadfence();                                 // ... and ordered before the flag (cumulative at gpu scope)
    st_relaxed_gpu_u32(flags + blockIdx.x, epoch);
  }
  for (int q = threadIdx.x; q < G; q += PERSIST_BLOCK) {
    long long spins = 0;
    while (ld_relaxed_gpu_u32(flags + q) < epoch) {
      if (++spins > PERSIST_SPIN_LIMIT) { sh.timeout = 1; break; }
      if (spins > 4096) __nanosleep(64);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) __threadfence();             // acquire side: drops the SM's L1 lines (CCTL.IVALL)
  __syncthreads();
}

// Barrier B: grid-wide sum of three values, result in sh.tot[] for every thread of every CTA (same bits everywhere).
__device__ __forceinline__ void persist_reduce(double (&v)[3], unsigned long long* mail, unsigned int epoch, int G,
                                               PersistShared& sh, double* s_scratch /* >= 3*G doubles */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double w = warp_sum(v[i]);
    if (lane == 0) sh.red[i][wid] = w;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < PERSIST_NW; ++k) s += sh.red[threadIdx.x][k];
    sh.loc[threadIdx.x] = s;
  }
  __syncthreads();
  const unsigned long long flag = (unsigned long long)epoch << 32;
  if (threadIdx.x < 6) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(sh.loc[threadIdx.x >> 1]);
    const unsigned long long half = (threadIdx.x & 1) ? (bits >> 32) : (bits & 0xffffffffull);
    st_relaxed_gpu_u64(mail + (size_t)blockIdx.x * 8 + threadIdx.x, flag | half);
  }
  // thread (q, w) polls word w of CTA q
  for (int t = threadIdx.x; t < G * 6; t += PERSIST_BLOCK) {
    const int q = t / 6, w = t - q * 6;
    unsigned long long word;
    long long spins = 0;
    for (;;) {
      word = ld_relaxed_gpu_u64(mail + (size_t)q * 8 + w);
      if ((word >> 32) == (unsigned long long)epoch) break;
      if (++spins > PERSIST_SPIN_LIMIT) { sh.timeout = 1; break; }
      if (spins > 4096) __nanosleep(64);
    }
    reinterpret_cast<unsigned int*>(s_scratch)[(size_t)(q * 3 + (w >> 1)) * 2 + (w & 1)] = (unsigned int)(word & 0xffffffffull);
  }
  __syncthreads();
  // fixed-shape sum over the CTAs: warp i adds value i of CTA lane, lane+32, ... then a butterfly
  if (wid < 3) {
    double s = 0.0;
    for (int q = lane; q < G; q += 32) s += s_scratch[q * 3 + wid];
    s = warp_sum(s);
    if (lane == 0) sh.tot[wid] = s;
  }
  __syncthreads();
}

template <int PC>
__global__ void __launch_bounds__(PERSIST_BLOCK, 1) k_pcg_persist(PersistArgs a) {
  extern __shared__ __align__(16) unsigned char persist_smem[];
  __shared__ PersistShared sh;
  const int G = gridDim.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int g = lane / 6, rr_ = lane - g * 6;
  const int R0 = a.cta_row0[blockIdx.x], R1 = a.cta_row0[blockIdx.x + 1];
  const int nrows = R1 - R0;
  const int b0 = a.rowptr[R0];
  const int nblk = a.rowptr[R1] - b0;
  const size_t vec = (size_t)a.rows_cap * 6;
  double* s_r = reinterpret_cast<double*>(persist_smem);
  double* s_p = s_r + vec;
  double* s_s = s_p + vec;
  double* s_w = s_s + vec;
  double* s_x = s_w + vec;
  double* s_scratch = s_x + vec;                                    // 3 * G doubles
  int32_t* s_rp = reinterpret_cast<int32_t*>(s_scratch + 3 * G);    // nrows + 1 block offsets relative to b0
  int32_t* s_col = s_rp + a.rows_cap + 1;
  if (threadIdx.x == 0) sh.timeout = 0;
  for (int i = threadIdx.x; i <= nrows; i += PERSIST_BLOCK) s_rp[i] = a.rowptr[R0 + i] - b0;
  for (int j = threadIdx.x; j < nblk; j += PERSIST_BLOCK) s_col[j] = a.colidx[b0 + j];
  const double* __restrict__ vals = a.vals + (size_t)b0 * 36;
  double* __restrict__ u = a.u;
  const PcgParams prm = a.prm;
  const double tol2 = prm.tol * prm.tol;
  __syncthreads();

  // ---- init: x = 0, r = b, u = M^-1 b, p = s = 0
  for (int base = wid * 5; base < nrows; base += PERSIST_NW * 5) {
    const int lr = base + g;
    const bool active = g < 5 && lr < nrows;
    const int64_t n = R0 + lr;
    const int li = lr * 6 + rr_;
    const double bv = active ? a.b[n * 6 + rr_] : 0.0;
    const double zv = apply_precond<PC>(a.dinv, n, g, rr_, active, bv);
    if (active) { s_x[li] = 0.0; s_r[li] = bv; s_p[li] = 0.0; s_s[li] = 0.0; u[n * 6 + rr_] = zv; }
  }
  unsigned int epochA = 0, epochB = 0;
  persist_barrier(a.flags, ++epochA, G, sh);

  int first = 1, iters = 0, done = 0, breakdown = 0, restarts = 0;
  double gamma_old = 0.0, alpha = 0.0, beta = 0.0, bb = 0.0, rr = 0.0, true_rr = -1.0;
  for (;;) {
    if (sh.timeout) break;
    // ---- product phase: w = A u for the owned rows, partial (r,u), (w,u), (r,r)
    double v[3] = {0.0, 0.0, 0.0};
    for (int base = wid * 5; base < nrows; base += PERSIST_NW * 5) {
      const int lr = base + g;
      const bool active = g < 5 && lr < nrows;
      if (active) {
        const int lo = s_rp[lr], hi = s_rp[lr + 1];
        const int64_t n = R0 + lr;
        const double uo = u[n * 6 + rr_];
        double acc = 0.0;
#pragma unroll 4
        for (int j = lo; j < hi; ++j) {
          const int c = s_col[j];
          const double2* vp = reinterpret_cast<const double2*>(vals + (size_t)j * 36 + rr_ * 6);
          const double2* xp = reinterpret_cast<const double2*>(u + (int64_t)c * 6);
          const double2 a0 = __ldg(vp), a1 = __ldg(vp + 1), a2 = __ldg(vp + 2);
          const double2 x0 = xp[0], x1 = xp[1], x2 = xp[2];
          acc = dot6(a0, a1, a2, x0, x1, x2, acc);
        }
        const int li = lr * 6 + rr_;
        const double ro = s_r[li];
        s_w[li] = acc;
        v[0] = fma(ro, uo, v[0]);
        v[1] = fma(acc, uo, v[1]);
        v[2] = fma(ro, ro, v[2]);
      }
    }
    persist_reduce(v, a.mail, ++epochB, G, sh, s_scratch);
    if (sh.timeout) break;
    const double gamma = sh.tot[0], delta = sh.tot[1];
    rr = sh.tot[2];
    // ---- scalar recurrences and stop test (cg_finish), evaluated identically by every thread
    if (first) {
      if (first == 1) bb = rr;
      if (first == 2 && rr <= tol2 * bb) done = 1;
      else {
        beta = 0.0; gamma_old = gamma; alpha = gamma / delta;
        if (rr == 0.0) done = 1;
        else if (!(delta > 0.0)) { done = 1; breakdown = 1; }
      }
      first = 0;
    } else {
      iters += 1;
      if (rr <= tol2 * bb) done = 1;
      else {
        const double bt = gamma / gamma_old;
        const double denom = delta - bt * gamma / alpha;
        if (!(denom > 0.0) || !(rr == rr)) { done = 1; breakdown = 1; }
        else { beta = bt; alpha = gamma / denom; gamma_old = gamma; }
      }
    }
    if (done || iters >= prm.maxiter) {
      if (!(done && !breakdown && bb > 0.0)) break;
      // ---- true-residual safeguard: r_true = b - A x
      for (int i = threadIdx.x; i < nrows * 6; i += PERSIST_BLOCK) a.x[(size_t)R0 * 6 + i] = s_x[i];
      persist_barrier(a.flags, ++epochA, G, sh);
      if (sh.timeout) break;
      double t[3] = {0.0, 0.0, 0.0};
      for (int base = wid * 5; base < nrows; base += PERSIST_NW * 5) {
        const int lr = base + g;
        const bool active = g < 5 && lr < nrows;
        if (active) {
          const int lo = s_rp[lr], hi = s_rp[lr + 1];
          const int64_t n = R0 + lr;
          double acc = 0.0;
#pragma unroll 4
          for (int j = lo; j < hi; ++j) {
            const int c = s_col[j];
            const double2* vp = reinterpret_cast<const double2*>(vals + (size_t)j * 36 + rr_ * 6);
            const double2* xp = reinterpret_cast<const double2*>(a.x + (int64_t)c * 6);
            const double2 a0 = __ldg(vp), a1 = __ldg(vp + 1), a2 = __ldg(vp + 2);
            const double2 x0 = xp[0], x1 = xp[1], x2 = xp[2];
            acc = dot6(a0, a1, a2, x0, x1, x2, acc);
          }
          const double rt = a.b[n * 6 + rr_] - acc;
          s_w[lr * 6 + rr_] = rt;
          t[0] = fma(rt, rt, t[0]);
        }
      }
      persist_reduce(t, a.mail, ++epochB, G, sh, s_scratch);
      if (sh.timeout) break;
      true_rr = sh.tot[0];
      if (!(true_rr > 4.0 * tol2 * bb)) break;                       // accepted
      if (restarts >= 2 || iters >= prm.maxiter) { breakdown = 3; break; }
      // restart from x with r = r_true
      restarts += 1;
      for (int base = wid * 5; base < nrows; base += PERSIST_NW * 5) {
        const int lr = base + g;
        const bool active = g < 5 && lr < nrows;
        const int64_t n = R0 + lr;
        const int li = lr * 6 + rr_;
        const double rv = active ? s_w[li] : 0.0;
        const double zv = apply_precond<PC>(a.dinv, n, g, rr_, active, rv);
        if (active) { s_r[li] = rv; s_p[li] = 0.0; s_s[li] = 0.0; u[n * 6 + rr_] = zv; }
      }
      persist_barrier(a.flags, ++epochA, G, sh);
      done = 0;
      first = 2;
      continue;
    }
    // ---- update phase: p = u + beta p; s = w + beta s; x += alpha p; r -= alpha s; u = M^-1 r
    for (int base = wid * 5; base < nrows; base += PERSIST_NW * 5) {
      const int lr = base + g;
      const bool active = g < 5 && lr < nrows;
      const int64_t n = R0 + lr;
      const int li = lr * 6 + rr_;
      const PrecondRow<PC> pr = load_precond<PC>(a.dinv, n, rr_, active);
      double rv = 0.0;
      if (active) {
        const double uv = u[n * 6 + rr_];
        const double pv = fma(beta, s_p[li], uv);
        const double sv = fma(beta, s_s[li], s_w[li]);
        s_p[li] = pv;
        s_s[li] = sv;
        s_x[li] = fma(alpha, pv, s_x[li]);
        rv = fma(-alpha, sv, s_r[li]);
        s_r[li] = rv;
      }
      const double zn = apply_precond<PC>(pr, g, rv);
      if (active) u[n * 6 + rr_] = zn;
    }
    persist_barrier(a.flags, ++epochA, G, sh);
  }
  __syncthreads();
  const int timed_out = sh.timeout;
  // ---- epilogue: solution and status
  for (int i = threadIdx.x; i < nrows * 6; i += PERSIST_BLOCK) a.x[(size_t)R0 * 6 + i] = s_x[i];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    PcgScalars* sc = a.sc;
    sc->iters = iters;
    sc->bb = bb;
    sc->rr = rr;
    sc->done = done;
    sc->breakdown = timed_out ? 2 : breakdown;
    sc->restarts = restarts;
    sc->true_rr = true_rr;
    sc->seq = (int)epochB;
  }
  if (timed_out && threadIdx.x == 0) a.sc->p2p_timeout = 1;
}

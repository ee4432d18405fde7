// Shared declarations for the lattice_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>

#include "lattice_b200.h"

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

// Device-resident scalars of one PCG solve (also copied to pinned host memory when polled).
struct PcgScalars {
  double rz_old, pAp, pp, rr, xx, bb, beta, alpha_last;
  int32_t done, info_flag2, iters, breakdown;
  unsigned int counter[4];  // last-block-done tickets (one per kernel family)
  double sums[4];           // local partial sums awaiting the all-reduce (multi-GPU path)
  double gamma, alpha;      // Chronopoulos-Gear variant: (r,u) and the current step length
  int32_t first, pad;
  int32_t seq, p2p_timeout; // peer-memory path: sequence number of the last completed reduction
  double true_rr;           // |b - A x|^2 measured by the safeguard
  int32_t restarts, pad2;
};

// Rigid-body-mode coarse space of the two-level preconditioner (coarse.cuh), registered by lat_coarse_setup.
struct CoarseSpace {
  int64_t n_nodes = -1;
  int32_t n_agg = 0, n_pieces = 0;
  bool active = false;            // an inverse is registered: the PCG drivers add the coarse correction
  bool fused = false;             // every aggregate is exactly one piece: 2 launches per correction instead of 4
  const double* einv = nullptr;   // [6 n_agg][6 n_agg] row-major (borrowed)
  // ctx-owned buffers: "coarse_nodes" [n_nodes] CoarseNode in aggregate order, "coarse_bynode" [n_nodes] (agg, d, mask)
  // by node, "coarse_ptr" [n_agg+1], "coarse_piece_ptr" [n_pieces+1], "coarse_piece_agg" [n_pieces],
  // "coarse_agg_piece" [n_agg+1], "coarse_part" [6 n_pieces], "coarse_rc" / "coarse_yc" [6 n_agg]
};

struct lat_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  std::string err;
  int64_t launches = 0;
  std::map<std::string, DevBuf> bufs;
  // resident sparsity pattern (lat_bsr_pattern_build)
  int64_t pat_nelem = -1, pat_nnodes = -1, pat_nnzb = -1;
  // resident matrix-free operator (lat_matfree_setup): material constants; arrays live in bufs "mf_*"
  int64_t mf_nnodes = -1, mf_nelem = -1;
  double mf_young = 0.0, mf_nu = 0.0, mf_kappa = 0.0;
  // pinned host staging
  PcgScalars* h_scal = nullptr;  // 2 slots
  int64_t* h_i64 = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  // ctx-owned non-blocking stream: stands in for the caller's stream wherever stream capture is needed and the
  // caller handed us the legacy default stream (which cannot be captured); ordered after it by ev_order
  cudaStream_t work = nullptr;
  cudaEvent_t ev_order = nullptr;
  // multi-GPU
  void* nccl_comm = nullptr;
  int nranks = 1, rank = 0;
  // NVLink peer-memory path (lat_p2p_*): one arena per rank, mapped into every rank
  struct P2P* p2p = nullptr;
  CoarseSpace coarse;
};

int lat_fail(lat_ctx* ctx, int code, const char* what, const char* file, int line);
int lat_cuda_fail(lat_ctx* ctx, cudaError_t e, const char* what, const char* file, int line);
void* lat_buf_raw(lat_ctx* ctx, const char* name, size_t bytes);

template <class T>
static inline T* lat_buf(lat_ctx* ctx, const char* name, size_t count) {
  return reinterpret_cast<T*>(lat_buf_raw(ctx, name, count * sizeof(T)));
}

#define LAT_CHECK_ARG(ctx, cond)                                              \
  do {                                                                        \
    if (!(cond)) return lat_fail((ctx), LAT_ERR_ARG, #cond, __FILE__, __LINE__); \
  } while (0)

#define LAT_CUDA(ctx, call)                                                   \
  do {                                                                        \
    cudaError_t _e = (call);                                                  \
    if (_e != cudaSuccess) return lat_cuda_fail((ctx), _e, #call, __FILE__, __LINE__); \
  } while (0)

// Launch on the ctx stream, count it, and surface launch-configuration errors.
#define LAT_LAUNCH(ctx, kernel, grid, block, smem, ...)                       \
  do {                                                                        \
    kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);          \
    (ctx)->launches++;                                                        \
    cudaError_t _e = cudaPeekAtLastError();                                   \
    if (_e != cudaSuccess) return lat_cuda_fail((ctx), _e, #kernel, __FILE__, __LINE__); \
  } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic grid reduction of NV values: every block writes its partial sums, the last block
// to finish (ticket counter) adds the partials in a fixed order.  Returns true in thread 0 of that
// last block, with the totals in out[].  partials: [NV][gridDim.x].  All threads must call this.
// Cost note (profiles/r01_fence_ab.txt): the gpu-scope release before the ticket keeps every CTA
// resident ~1.5 us longer; kernels on the critical path of the PCG iteration therefore use
// block_partials() + a separate one-CTA reduce kernel instead.
template <int NV, int BLOCK>
__device__ __forceinline__ bool grid_reduce(double (&v)[NV], double* partials, unsigned int* ticket,
                                            double (&out)[NV]) {
  __shared__ double s_part[NV][BLOCK / 32];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double w = warp_sum(v[i]);
    if (lane == 0) s_part[i][wid] = w;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double s = 0.0;
      for (int k = 0; k < BLOCK / 32; ++k) s += s_part[i][k];
      partials[(size_t)i * gridDim.x + blockIdx.x] = s;
    }
    // RELEASE ticket: orders this thread's partial stores before the increment (MEMBAR.ALL.GPU +
    // ATOMG) without the CCTL.IVALL L1 invalidate that __threadfence() adds on sm_100
    unsigned int t;
    asm volatile("atom.release.gpu.global.add.u32 %0, [%1], 1;" : "=r"(t) : "l"(ticket) : "memory");
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return false;
  // last block only: acquire, then a fixed-order tree over the partials (independent of arrival order)
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double s = 0.0;
    for (unsigned int k = threadIdx.x; k < gridDim.x; k += BLOCK)
      s += __ldcg(&partials[(size_t)i * gridDim.x + k]);
    s = warp_sum(s);
    if (lane == 0) s_part[i][wid] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double s = 0.0;
      for (int k = 0; k < BLOCK / 32; ++k) s += s_part[i][k];
      out[i] = s;
    }
    *ticket = 0u;  // re-arm for the next launch
    return true;
  }
  return false;
}

// Block-level half of the reduction only: partials[i][blockIdx.x] = sum over the CTA of v[i].
// Plain stores, no ordering needed -- the consumer is a later kernel.
template <int NV, int BLOCK>
__device__ __forceinline__ void block_partials(double (&v)[NV], double* partials, int stride = 0, int offset = 0) {
  // layout [NV][stride]; this CTA's slot is offset + blockIdx.x (stride 0: this kernel's own grid).  Two
  // kernels can fill one array (interior rows at offset 0, boundary rows behind them).
  __shared__ double s_bp[NV][BLOCK / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double w = warp_sum(v[i]);
    if (lane == 0) s_bp[i][wid] = w;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < BLOCK / 32; ++k) s += s_bp[threadIdx.x][k];
    const size_t st = stride > 0 ? (size_t)stride : (size_t)gridDim.x;
    partials[(size_t)threadIdx.x * st + offset + blockIdx.x] = s;
  }
}

// Fixed-order sum of partials[NV][n_part] by ONE CTA of BLOCK threads; totals valid in thread 0.
template <int NV, int BLOCK>
__device__ __forceinline__ void sum_partials(const double* __restrict__ partials, int n_part, double (&out)[NV]) {
  __shared__ double s_sp[NV][BLOCK / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double s = 0.0;
    for (int k = threadIdx.x; k < n_part; k += BLOCK) s += partials[(size_t)i * n_part + k];
    s = warp_sum(s);
    if (lane == 0) s_sp[i][wid] = s;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double s = 0.0;
    if (threadIdx.x == 0)
      for (int k = 0; k < BLOCK / 32; ++k) s += s_sp[i][k];
    out[i] = s;
  }
}

// Coefficients that define the 12x12 stiffness of one element in global axes.
// Every 3x3 sub-block of K_e is a*I + b*t t^T + c*[t]x (circular section:
// EI1 == EI2 and GS1 == GS2, so a1 a1^T + a2 a2^T = I - t t^T and
// a1 a2^T - a2 a1^T = -[t]x; see DESIGN.md "element algebra").
struct ElemCoef {
  double tx, ty, tz;
  double aI, aT;  // ww block:      A  = aI*I + aT*tt          (aI = GS/L, aT = (ES-GS)/L)
  double c;       // w-theta block: C  = -c*[t]x               (c  = GS/2)
  double bI;      // theta-theta shear part: Bm = bI*(I - tt)  (bI = GS*L/4)
  double dI, dT;  // theta-theta bending/torsion: Dm = dI*I + dT*tt (dI = EI/L, dT = (GJ-EI)/L)
};

__device__ __forceinline__ ElemCoef elem_coef(double x0, double y0, double z0, double x1, double y1,
                                              double z1, double r, double young, double nu,
                                              double kappa, bool drad) {
  ElemCoef e;
  const double dx = x1 - x0, dy = y1 - y0, dz = z1 - z0;
  const double L = sqrt(dx * dx + dy * dy + dz * dz);
  const double iL = 1.0 / L;
  e.tx = dx * iL;
  e.ty = dy * iL;
  e.tz = dz * iL;
  const double PI = 3.14159265358979323846;
  const double G = young / (2.0 * (1.0 + nu));
  double S, I;
  if (!drad) {
    S = PI * r * r;            // material_definition.py:147
    I = PI * r * r * r * r * 0.25;
  } else {
    S = 2.0 * PI * r;          // material_definition.py:216-217 (normal beams)
    I = PI * r * r * r;
  }
  const double ES = young * S, GS = G * kappa * S, EI = young * I, GJ = G * 2.0 * I;
  e.aI = GS * iL;
  e.aT = (ES - GS) * iL;
  e.c = 0.5 * GS;
  e.bI = 0.25 * GS * L;
  e.dI = EI * iL;
  e.dT = (GJ - EI) * iL;
  return e;
}

// 6x6 block (row node end `re`, column node end `ce`, 0 = first node of the element)
// accumulated into acc[36] (row-major) with weight w.
__device__ __forceinline__ void elem_block_accum(const ElemCoef& e, int re, int ce, double w,
                                                 double (&acc)[36]) {
  const double sA = (re == ce) ? 1.0 : -1.0;
  const double sC = (re == 0) ? -e.c : e.c;   // upper-right block = sC*[t]x  (C = -c[t]x for row end 0)
  const double sT = (ce == 0) ? e.c : -e.c;   // lower-left  block = sT*[t]x  (C^T = +c[t]x for col end 0)
  const double t[3] = {e.tx, e.ty, e.tz};
  const double aI = sA * e.aI, aT = sA * e.aT;
  const double qI = e.bI + sA * e.dI, qT = sA * e.dT - e.bI;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const double tt = t[a] * t[b];
      const double d = (a == b) ? 1.0 : 0.0;
      // [t]x entries: Sk[a][b] = -eps_{abk} t_k
      double sk = 0.0;
      if (a == 0 && b == 1) sk = -t[2];
      if (a == 0 && b == 2) sk = t[1];
      if (a == 1 && b == 0) sk = t[2];
      if (a == 1 && b == 2) sk = -t[0];
      if (a == 2 && b == 0) sk = -t[1];
      if (a == 2 && b == 1) sk = t[0];
      acc[a * 6 + b] += w * (aI * d + aT * tt);
      acc[a * 6 + 3 + b] += w * (sC * sk);
      acc[(a + 3) * 6 + b] += w * (sT * sk);
      acc[(a + 3) * 6 + 3 + b] += w * (qI * d + qT * tt);
    }
  }
}

// Single entry (i, j), 0 <= i, j < 6, of the 6x6 block (row node end re, column node end ce).
__device__ __forceinline__ double elem_block_entry(const ElemCoef& e, int re, int ce, int i, int j) {
  const double sA = (re == ce) ? 1.0 : -1.0;
  const bool rw = i < 3, cw = j < 3;
  const int a = rw ? i : i - 3, b = cw ? j : j - 3;
  const double ta = (a == 0) ? e.tx : (a == 1 ? e.ty : e.tz);
  const double tb = (b == 0) ? e.tx : (b == 1 ? e.ty : e.tz);
  const double d = (a == b) ? 1.0 : 0.0;
  if (rw && cw) return sA * (e.aI * d + e.aT * ta * tb);
  if (!rw && !cw) return (e.bI + sA * e.dI) * d + (sA * e.dT - e.bI) * ta * tb;
  double sk = 0.0;
  if (a != b) {
    const int k = 3 - a - b;
    const double tk = (k == 0) ? e.tx : (k == 1 ? e.ty : e.tz);
    sk = ((b - a + 3) % 3 == 1) ? -tk : tk;  // [t]x entry (a, b)
  }
  if (rw) return ((re == 0) ? -e.c : e.c) * sk;
  return ((ce == 0) ? e.c : -e.c) * sk;
}

#endif  // __CUDACC__

// lattice_core.cu -- context, element stiffness, sparsity pattern, assembly,
// Dirichlet elimination, CSR export, compliance sensitivity.  sm_100a.
#include <cstdio>
#include <cstring>

#include "common.cuh"

// ===========================================================================
// context
// ===========================================================================
int lat_fail(lat_ctx* ctx, int code, const char* what, const char* file, int line) {
  if (ctx) {
    char buf[512];
    snprintf(buf, sizeof buf, "lattice_b200: %s (%s:%d)", what, file, line);
    ctx->err = buf;
  }
  return code;
}

int lat_cuda_fail(lat_ctx* ctx, cudaError_t e, const char* what, const char* file, int line) {
  if (ctx) {
    char buf[768];
    snprintf(buf, sizeof buf, "lattice_b200: CUDA error %d (%s) in %s (%s:%d)", (int)e,
             cudaGetErrorString(e), what, file, line);
    ctx->err = buf;
  }
  return (int)e > 0 ? (int)e : 999;
}

void* lat_buf_raw(lat_ctx* ctx, const char* name, size_t bytes) {
  DevBuf& b = ctx->bufs[name];
  if (bytes == 0) bytes = 16;
  if (b.cap < bytes) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 16 + 256;
    if (cudaMalloc(&b.p, want) != cudaSuccess) {
      cudaGetLastError();
      b.p = nullptr;
      return nullptr;
    }
    b.cap = want;
  }
  return b.p;
}

extern "C" int lat_version(void) { return 100; }

extern "C" int lat_ctx_create(int device, void* stream, lat_ctx** out) {
  if (!out) return LAT_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return e != cudaSuccess ? (int)e : 100 /* cudaErrorNoDevice */;
  if (device < 0 || device >= ndev) return LAT_ERR_ARG;
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return (int)e;
  lat_ctx* ctx = new lat_ctx();
  ctx->device = device;
  ctx->stream = reinterpret_cast<cudaStream_t>(stream);
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  if (cudaMallocHost(&ctx->h_scal, 2 * sizeof(PcgScalars)) != cudaSuccess ||
      cudaMallocHost(&ctx->h_i64, 16 * sizeof(int64_t)) != cudaSuccess) {
    delete ctx;
    return 2 /* cudaErrorMemoryAllocation */;
  }
  for (int i = 0; i < 4; ++i) cudaEventCreate(&ctx->ev[i]);
  *out = ctx;
  return LAT_OK;
}

extern "C" int lat_ctx_destroy(lat_ctx* ctx) {
  if (!ctx) return LAT_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& kv : ctx->bufs)
    if (kv.second.p) cudaFree(kv.second.p);
  if (ctx->h_scal) cudaFreeHost(ctx->h_scal);
  if (ctx->h_i64) cudaFreeHost(ctx->h_i64);
  for (int i = 0; i < 4; ++i)
    if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  if (ctx->work) cudaStreamDestroy(ctx->work);
  if (ctx->ev_order) cudaEventDestroy(ctx->ev_order);
  delete ctx;
  return LAT_OK;
}

extern "C" const char* lat_last_error(lat_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

extern "C" int lat_ctx_sync(lat_ctx* ctx) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  LAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return LAT_OK;
}

extern "C" int64_t lat_launch_count(lat_ctx* ctx) { return ctx ? ctx->launches : -1; }

// ===========================================================================
// A1: element stiffness, element-major output
// ===========================================================================
// One warp per 32 elements: each lane derives the coefficients of its element,
// parks them in shared memory, then the warp streams the 32*144 entries out so
// that consecutive lanes write consecutive doubles (coalesced 256 B per request).
__global__ void __launch_bounds__(128) k_elem_stiffness(
    const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
    const int32_t* __restrict__ en0, const int32_t* __restrict__ en1,
    const double* __restrict__ rad, int64_t n_elem, double young, double nu, double kappa,
    int drad, double* __restrict__ Ke) {
  __shared__ ElemCoef s_coef[4][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t warp = (int64_t)blockIdx.x * 4 + wid;
  const int64_t e0 = warp * 32;
  if (e0 >= n_elem) return;
  const int64_t e = e0 + lane;
  if (e < n_elem) {
    const int a = en0[e], b = en1[e];
    s_coef[wid][lane] = elem_coef(x[a], y[a], z[a], x[b], y[b], z[b], rad[e], young, nu, kappa, drad != 0);
  }
  __syncwarp();
  const int cnt = (int)min((int64_t)32, n_elem - e0);
  double* out = Ke + e0 * 144;
  for (int idx = lane; idx < cnt * 144; idx += 32) {
    const int le = idx / 144, ij = idx - le * 144;
    const int i = ij / 12, j = ij - i * 12;
    const ElemCoef& c = s_coef[wid][le];
    const int bi = i / 3, a = i - bi * 3, bj = j / 3, b = j - bj * 3;  // blocks: 0 w1, 1 th1, 2 w2, 3 th2
    const int re = bi >> 1, ce = bj >> 1;
    const bool rw = (bi & 1) == 0, cw = (bj & 1) == 0;
    const double sA = (re == ce) ? 1.0 : -1.0;
    const double t[3] = {c.tx, c.ty, c.tz};
    const double tt = t[a] * t[b];
    const double d = (a == b) ? 1.0 : 0.0;
    double sk = 0.0;
    if (a != b) {
      const int k = 3 - a - b;                        // remaining axis
      const double sgn = ((b - a + 3) % 3 == 1) ? -1.0 : 1.0;  // Sk[a][b] = -eps_{abk} t_k
      sk = sgn * t[k];
    }
    double v;
    if (rw && cw) v = sA * (c.aI * d + c.aT * tt);
    else if (rw && !cw) v = ((re == 0) ? -c.c : c.c) * sk;
    else if (!rw && cw) v = ((ce == 0) ? c.c : -c.c) * sk;
    else v = (c.bI + sA * c.dI) * d + (sA * c.dT - c.bI) * tt;
    out[idx] = v;
  }
}

// Same output, one thread per ELEMENT: the four 6x6 quadrants are generated with compile-time indices
// (no div/mod per entry) and each 96 B row of K_e leaves as three 256-bit stores.  The per-entry kernel above
// is issue-bound at 0.32 of HBM; it stays as the fallback for an output that is not 32 B aligned.
__device__ __forceinline__ void st256_ke(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__global__ void __launch_bounds__(128) k_elem_stiffness_rows(
    const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
    const int32_t* __restrict__ en0, const int32_t* __restrict__ en1,
    const double* __restrict__ rad, int64_t n_elem, double young, double nu, double kappa,
    int drad, double* __restrict__ Ke) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elem) return;
  const int a = en0[e], b = en1[e];
  const ElemCoef co = elem_coef(x[a], y[a], z[a], x[b], y[b], z[b], rad[e], young, nu, kappa, drad != 0);
  double* out = Ke + e * 144;
#pragma unroll
  for (int re = 0; re < 2; ++re) {
    double q0[36], q1[36];
#pragma unroll
    for (int k = 0; k < 36; ++k) { q0[k] = 0.0; q1[k] = 0.0; }
    elem_block_accum(co, re, 0, 1.0, q0);
    elem_block_accum(co, re, 1, 1.0, q1);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      double* row = out + (re * 6 + i) * 12;
      st256_ke(row, q0[i * 6], q0[i * 6 + 1], q0[i * 6 + 2], q0[i * 6 + 3]);
      st256_ke(row + 4, q0[i * 6 + 4], q0[i * 6 + 5], q1[i * 6], q1[i * 6 + 1]);
      st256_ke(row + 8, q1[i * 6 + 2], q1[i * 6 + 3], q1[i * 6 + 4], q1[i * 6 + 5]);
    }
  }
}

extern "C" int lat_elem_stiffness(lat_ctx* ctx, const double* x, const double* y, const double* z,
                                  const int32_t* en0, const int32_t* en1, const double* rad,
                                  int64_t n_elem, double young, double nu, double kappa, int drad,
                                  double* Ke) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, n_elem >= 0);
  if (n_elem == 0) return LAT_OK;
  LAT_CHECK_ARG(ctx, x && y && z && en0 && en1 && rad && Ke);
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t grid = ceil_div(n_elem, 128);
  if ((reinterpret_cast<uintptr_t>(Ke) & 31) == 0)
    LAT_LAUNCH(ctx, k_elem_stiffness_rows, (unsigned)grid, 128, 0, x, y, z, en0, en1, rad, n_elem, young, nu,
               kappa, drad, Ke);
  else
    LAT_LAUNCH(ctx, k_elem_stiffness, (unsigned)grid, 128, 0, x, y, z, en0, en1, rad, n_elem, young, nu,
               kappa, drad, Ke);
  return LAT_OK;
}

// ===========================================================================
// exclusive scan of int32 counts (three-kernel block scan)
// ===========================================================================
static constexpr int SCAN_BLOCK = 1024;

__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_block(const int32_t* __restrict__ in, int64_t n,
                                                           int32_t* __restrict__ out,
                                                           int32_t* __restrict__ sums) {
  __shared__ int32_t s_w[SCAN_BLOCK / 32];
  const int64_t i = (int64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int32_t v = (i < n) ? in[i] : 0;
  int32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_w[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int32_t w = s_w[lane];
    int32_t winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int32_t t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    s_w[lane] = winc - w;  // exclusive warp offsets
    if (lane == 31) sums[blockIdx.x] = winc;
  }
  __syncthreads();
  if (i < n) out[i] = inc - v + s_w[wid];
}

__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_sums(int32_t* __restrict__ sums, int64_t nb,
                                                          int32_t* __restrict__ total) {
  // single block: sequential chunks of SCAN_BLOCK
  __shared__ int32_t s_w[SCAN_BLOCK / 32];
  __shared__ int32_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int64_t base = 0; base < nb; base += SCAN_BLOCK) {
    const int64_t i = base + threadIdx.x;
    int32_t v = (i < nb) ? sums[i] : 0;
    int32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      int32_t w = s_w[lane];
      int32_t winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int32_t t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
      }
      s_w[lane] = winc - w;
    }
    __syncthreads();
    const int32_t carry = s_carry;
    if (i < nb) sums[i] = inc - v + s_w[wid] + carry;
    __syncthreads();
    if (threadIdx.x == SCAN_BLOCK - 1) s_carry = carry + inc + s_w[wid];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = s_carry;
}

__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_add(int32_t* __restrict__ out, int64_t n,
                                                         const int32_t* __restrict__ sums,
                                                         const int32_t* __restrict__ total) {
  const int64_t i = (int64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
  if (i < n) out[i] += sums[blockIdx.x];
  if (i == 0) out[n] = *total;  // out has n+1 entries
}

// out[0..n] = exclusive scan of in[0..n-1]; out[n] = total.  in may alias out.
static int scan_exclusive(lat_ctx* ctx, const int32_t* in, int64_t n, int32_t* out) {
  const int64_t nb = ceil_div(n, SCAN_BLOCK);
  int32_t* sums = lat_buf<int32_t>(ctx, "scan_sums", (size_t)nb + 2);
  if (!sums) return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
  int32_t* total = sums + nb;
  LAT_LAUNCH(ctx, k_scan_block, (unsigned)nb, SCAN_BLOCK, 0, in, n, out, sums);
  LAT_LAUNCH(ctx, k_scan_sums, 1, SCAN_BLOCK, 0, sums, nb, total);
  LAT_LAUNCH(ctx, k_scan_add, (unsigned)nb, SCAN_BLOCK, 0, out, n, sums, total);
  return LAT_OK;
}

// ===========================================================================
// A3: sparsity pattern (counting sort by row node + per-row sort/unique + scans)
// ===========================================================================
__global__ void k_count_deg(const int32_t* __restrict__ en0, const int32_t* __restrict__ en1,
                            int64_t n_elem, int64_t n_nodes, int32_t* __restrict__ deg,
                            int32_t* __restrict__ bad) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elem) return;
  const int a = en0[e], b = en1[e];
  if (a < 0 || b < 0 || a >= n_nodes || b >= n_nodes || a == b) {
    atomicAdd(bad, 1);
    return;
  }
  atomicAdd(&deg[a], 1);
  atomicAdd(&deg[b], 1);
}

__global__ void k_fill_adj(const int32_t* __restrict__ en0, const int32_t* __restrict__ en1,
                           int64_t n_elem, const int32_t* __restrict__ adjptr,
                           int32_t* __restrict__ cursor, int32_t* __restrict__ adj_other,
                           int32_t* __restrict__ adj_el) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elem) return;
  const int a = en0[e], b = en1[e];
  int s = adjptr[a] + atomicAdd(&cursor[a], 1);
  adj_other[s] = b;
  adj_el[s] = (int32_t)(e * 2);
  s = adjptr[b] + atomicAdd(&cursor[b], 1);
  adj_other[s] = a;
  adj_el[s] = (int32_t)(e * 2 + 1);
}

// Per node: sort the incidence list by (other node, element) -- this removes the
// arbitrary order left by the atomics -- and count the distinct neighbours.
__global__ void k_sort_rows(int64_t n_nodes, const int32_t* __restrict__ adjptr,
                            int32_t* __restrict__ adj_other, int32_t* __restrict__ adj_el,
                            int32_t* __restrict__ nb) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_nodes) return;
  const int lo = adjptr[n], hi = adjptr[n + 1];
  for (int i = lo + 1; i < hi; ++i) {
    const int ko = adj_other[i], ke = adj_el[i];
    int j = i - 1;
    while (j >= lo && (adj_other[j] > ko || (adj_other[j] == ko && adj_el[j] > ke))) {
      adj_other[j + 1] = adj_other[j];
      adj_el[j + 1] = adj_el[j];
      --j;
    }
    adj_other[j + 1] = ko;
    adj_el[j + 1] = ke;
  }
  // a node without any element keeps an EMPTY row (like scipy's COO->CSR of the element
  // triplets, and like dolfinx, whose mesh does not contain unconnected gmsh points)
  int cnt = (hi > lo) ? 1 : 0;  // self
  for (int i = lo; i < hi; ++i)
    if (i == lo || adj_other[i] != adj_other[i - 1]) ++cnt;
  nb[n] = cnt;
}

__global__ void k_fill_cols(int64_t n_nodes, const int32_t* __restrict__ adjptr,
                            const int32_t* __restrict__ adj_other, const int32_t* __restrict__ rowptr,
                            int32_t* __restrict__ colidx, int32_t* __restrict__ blockrow,
                            int32_t* __restrict__ blk_lo, int32_t* __restrict__ blk_hi,
                            int32_t* __restrict__ diagpos) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_nodes) return;
  const int lo = adjptr[n], hi = adjptr[n + 1];
  if (hi == lo) { diagpos[n] = -1; return; }
  int w = rowptr[n];
  bool self_done = false;
  int i = lo;
  while (i < hi) {
    const int c = adj_other[i];
    if (!self_done && c > (int)n) {
      colidx[w] = (int)n; blockrow[w] = (int)n; blk_lo[w] = lo; blk_hi[w] = hi; diagpos[n] = w;
      ++w;
      self_done = true;
    }
    int j = i;
    while (j < hi && adj_other[j] == c) ++j;
    colidx[w] = c; blockrow[w] = (int)n; blk_lo[w] = i; blk_hi[w] = j;
    ++w;
    i = j;
  }
  if (!self_done) {
    colidx[w] = (int)n; blockrow[w] = (int)n; blk_lo[w] = lo; blk_hi[w] = hi; diagpos[n] = w;
  }
}

__device__ __forceinline__ int row_find(const int32_t* __restrict__ colidx, int lo, int hi, int c) {
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (colidx[mid] < c) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void k_elem_block(const int32_t* __restrict__ en0, const int32_t* __restrict__ en1,
                             int64_t n_elem, const int32_t* __restrict__ rowptr,
                             const int32_t* __restrict__ colidx, const int32_t* __restrict__ diagpos,
                             int32_t* __restrict__ elem_block) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elem) return;
  const int a = en0[e], b = en1[e];
  int4 r;
  r.x = diagpos[a];
  r.y = row_find(colidx, rowptr[a], rowptr[a + 1], b);
  r.z = row_find(colidx, rowptr[b], rowptr[b + 1], a);
  r.w = diagpos[b];
  reinterpret_cast<int4*>(elem_block)[e] = r;
}

extern "C" int lat_bsr_pattern_build(lat_ctx* ctx, const int32_t* en0, const int32_t* en1,
                                     int64_t n_elem, int64_t n_nodes, int64_t* nnzb) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, nnzb != nullptr);
  LAT_CHECK_ARG(ctx, n_elem >= 0 && n_nodes > 0);
  LAT_CHECK_ARG(ctx, n_elem == 0 || (en0 && en1));
  LAT_CHECK_ARG(ctx, n_nodes < (int64_t)1 << 30 && n_elem < (int64_t)1 << 30);
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  ctx->pat_nnzb = -1;
  ctx->mf_nnodes = -1;   // a resident matrix-free operator refers to the old incidence lists
  int32_t* deg = lat_buf<int32_t>(ctx, "pat_deg", n_nodes + 1);
  int32_t* adjptr = lat_buf<int32_t>(ctx, "pat_adjptr", n_nodes + 1);
  int32_t* cursor = lat_buf<int32_t>(ctx, "pat_cursor", n_nodes + 1);
  int32_t* adj_other = lat_buf<int32_t>(ctx, "pat_adj_other", 2 * n_elem + 1);
  int32_t* adj_el = lat_buf<int32_t>(ctx, "pat_adj_el", 2 * n_elem + 1);
  int32_t* rowptr = lat_buf<int32_t>(ctx, "pat_rowptr", n_nodes + 1);
  int32_t* diagpos = lat_buf<int32_t>(ctx, "pat_diagpos", n_nodes + 1);
  int32_t* bad = lat_buf<int32_t>(ctx, "pat_bad", 4);
  if (!deg || !adjptr || !cursor || !adj_other || !adj_el || !rowptr || !diagpos || !bad)
    return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
  LAT_CUDA(ctx, cudaMemsetAsync(deg, 0, (n_nodes + 1) * sizeof(int32_t), ctx->stream));
  LAT_CUDA(ctx, cudaMemsetAsync(cursor, 0, (n_nodes + 1) * sizeof(int32_t), ctx->stream));
  LAT_CUDA(ctx, cudaMemsetAsync(bad, 0, 4 * sizeof(int32_t), ctx->stream));
  const int TB = 256;
  const unsigned ge = (unsigned)ceil_div(n_elem > 0 ? n_elem : 1, TB);
  const unsigned gn = (unsigned)ceil_div(n_nodes, TB);
  if (n_elem > 0) LAT_LAUNCH(ctx, k_count_deg, ge, TB, 0, en0, en1, n_elem, n_nodes, deg, bad);
  int rc = scan_exclusive(ctx, deg, n_nodes, adjptr);
  if (rc) return rc;
  if (n_elem > 0) LAT_LAUNCH(ctx, k_fill_adj, ge, TB, 0, en0, en1, n_elem, adjptr, cursor, adj_other, adj_el);
  LAT_LAUNCH(ctx, k_sort_rows, gn, TB, 0, n_nodes, adjptr, adj_other, adj_el, deg /* reuse as nb */);
  rc = scan_exclusive(ctx, deg, n_nodes, rowptr);
  if (rc) return rc;
  int32_t h[2];
  LAT_CUDA(ctx, cudaMemcpyAsync(&h[0], rowptr + n_nodes, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  LAT_CUDA(ctx, cudaMemcpyAsync(&h[1], bad, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  LAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (h[1] != 0) return lat_fail(ctx, LAT_ERR_ARG, "element connectivity out of range or degenerate (en0 == en1)", __FILE__, __LINE__);
  const int64_t nz = h[0];
  int32_t* colidx = lat_buf<int32_t>(ctx, "pat_colidx", nz);
  int32_t* blockrow = lat_buf<int32_t>(ctx, "pat_blockrow", nz);
  int32_t* blk_lo = lat_buf<int32_t>(ctx, "pat_blk_lo", nz);
  int32_t* blk_hi = lat_buf<int32_t>(ctx, "pat_blk_hi", nz);
  int32_t* elem_block = lat_buf<int32_t>(ctx, "pat_elem_block", 4 * n_elem + 4);
  if (!colidx || !blockrow || !blk_lo || !blk_hi || !elem_block)
    return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
  LAT_LAUNCH(ctx, k_fill_cols, gn, TB, 0, n_nodes, adjptr, adj_other, rowptr, colidx, blockrow, blk_lo, blk_hi, diagpos);
  if (n_elem > 0) LAT_LAUNCH(ctx, k_elem_block, ge, TB, 0, en0, en1, n_elem, rowptr, colidx, diagpos, elem_block);
  ctx->pat_nelem = n_elem;
  ctx->pat_nnodes = n_nodes;
  ctx->pat_nnzb = nz;
  *nnzb = nz;
  return LAT_OK;
}

extern "C" int lat_bsr_pattern_export(lat_ctx* ctx, int32_t* rowptr, int32_t* colidx, int32_t* elem_block) {
  if (!ctx) return LAT_ERR_ARG;
  if (ctx->pat_nnzb < 0) return lat_fail(ctx, LAT_ERR_STATE, "no pattern resident: call lat_bsr_pattern_build first", __FILE__, __LINE__);
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  if (rowptr)
    LAT_CUDA(ctx, cudaMemcpyAsync(rowptr, ctx->bufs["pat_rowptr"].p, (ctx->pat_nnodes + 1) * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));
  if (colidx)
    LAT_CUDA(ctx, cudaMemcpyAsync(colidx, ctx->bufs["pat_colidx"].p, ctx->pat_nnzb * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));
  if (elem_block && ctx->pat_nelem > 0)
    LAT_CUDA(ctx, cudaMemcpyAsync(elem_block, ctx->bufs["pat_elem_block"].p, 4 * ctx->pat_nelem * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));
  return LAT_OK;
}

// ---------------------------------------------------------------------------
// scalar CSR view
// ---------------------------------------------------------------------------
__global__ void k_csr_structure(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                int64_t n_nodes, int32_t* __restrict__ indptr, int32_t* __restrict__ indices) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row > 6 * n_nodes) return;
  if (row == 6 * n_nodes) { indptr[row] = rowptr[n_nodes] * 36; return; }
  const int64_t n = row / 6;
  const int r = (int)(row - n * 6);
  const int lo = rowptr[n], hi = rowptr[n + 1];
  const int start = lo * 36 + r * 6 * (hi - lo);
  indptr[row] = start;
  if (indices) {
    int w = start;
    for (int j = lo; j < hi; ++j) {
      const int c6 = colidx[j] * 6;
#pragma unroll
      for (int k = 0; k < 6; ++k) indices[w++] = c6 + k;
    }
  }
}

__global__ void k_bsr_to_csr_values(const int32_t* __restrict__ rowptr, int64_t n_nodes,
                                    const double* __restrict__ bsr, double* __restrict__ csr) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= 6 * n_nodes) return;
  const int64_t n = row / 6;
  const int r = (int)(row - n * 6);
  const int64_t lo = rowptr[n], hi = rowptr[n + 1];
  int64_t w = lo * 36 + (int64_t)r * 6 * (hi - lo);
  for (int64_t j = lo; j < hi; ++j) {
#pragma unroll
    for (int k = 0; k < 6; ++k) csr[w++] = bsr[j * 36 + r * 6 + k];
  }
}

extern "C" int lat_csr_structure(lat_ctx* ctx, const int32_t* rowptr, const int32_t* colidx,
                                 int64_t n_nodes, int32_t* indptr, int32_t* indices) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, rowptr && colidx && indptr && n_nodes > 0);
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  // int32 CSR offsets (scipy's index type for this size) hold at most 2^31-1 scalar entries = 59.6 M blocks:
  // refuse instead of wrapping (the BSR path itself has no such limit).  [syncs]
  int32_t nnzb32 = 0;
  LAT_CUDA(ctx, cudaMemcpyAsync(&nnzb32, rowptr + n_nodes, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  LAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (nnzb32 < 0 || (int64_t)nnzb32 * 36 > (int64_t)INT32_MAX)
    return lat_fail(ctx, LAT_ERR_UNSUPPORTED, "scalar CSR structure needs more than 2^31-1 entries: use the BSR arrays", __FILE__, __LINE__);
  LAT_LAUNCH(ctx, k_csr_structure, (unsigned)ceil_div(6 * n_nodes + 1, 256), 256, 0, rowptr, colidx, n_nodes, indptr, indices);
  return LAT_OK;
}

extern "C" int lat_bsr_to_csr_values(lat_ctx* ctx, const int32_t* rowptr, int64_t n_nodes,
                                     const double* bsr_vals, double* csr_vals) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, rowptr && bsr_vals && csr_vals && n_nodes > 0);
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  LAT_LAUNCH(ctx, k_bsr_to_csr_values, (unsigned)ceil_div(6 * n_nodes, 256), 256, 0, rowptr, n_nodes, bsr_vals, csr_vals);
  return LAT_OK;
}

// ===========================================================================
// A1+A3 fused: element generation + assembly
// ===========================================================================
// Gather mode: one thread per BSR block.  The thread walks the (sorted)
// incidence range of its block, regenerates each contributing element's
// coefficients from 64 B of geometry and accumulates the relevant 6x6 quadrant
// in registers -- no K_e is ever materialised, no atomics, fixed summation
// order.  The 36 values are staged through shared memory (stride 37: conflict
// free for FP64) so that the CTA writes its 128 consecutive blocks (36 KB of
// contiguous HBM) with fully coalesced 128 B lines.
static constexpr int ASM_BLOCK = 128;

__global__ void __launch_bounds__(ASM_BLOCK) k_assemble_gather(
    const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
    const int32_t* __restrict__ en0, const int32_t* __restrict__ en1,
    const double* __restrict__ rad, const double* __restrict__ chain,
    const int32_t* __restrict__ blk_lo, const int32_t* __restrict__ blk_hi,
    const int32_t* __restrict__ blockrow, const int32_t* __restrict__ colidx,
    const int32_t* __restrict__ adj_el, int64_t nnzb, double young, double nu, double kappa,
    int drad, double* __restrict__ vals) {
  __shared__ double s_out[ASM_BLOCK * 37];
  const int64_t b0 = (int64_t)blockIdx.x * ASM_BLOCK;
  const int64_t b = b0 + threadIdx.x;
  if (b < nnzb) {
    double acc[36];
#pragma unroll
    for (int k = 0; k < 36; ++k) acc[k] = 0.0;
    const int lo = blk_lo[b], hi = blk_hi[b];
    const bool diag = blockrow[b] == colidx[b];
    for (int i = lo; i < hi; ++i) {
      const int ee = adj_el[i];
      const int e = ee >> 1, end = ee & 1;
      const int a = en0[e], c = en1[e];
      const ElemCoef co = elem_coef(x[a], y[a], z[a], x[c], y[c], z[c], rad[e], young, nu, kappa, drad != 0);
      const double w = (drad && chain) ? chain[e] : 1.0;
      elem_block_accum(co, end, diag ? end : (end ^ 1), w, acc);
    }
#pragma unroll
    for (int k = 0; k < 36; ++k) s_out[threadIdx.x * 37 + k] = acc[k];
  }
  __syncthreads();
  const int nblk = (int)min((int64_t)ASM_BLOCK, nnzb - b0);
  double* out = vals + b0 * 36;
  for (int i = threadIdx.x; i < nblk * 36; i += ASM_BLOCK) {
    const int q = i / 36;
    out[i] = s_out[q * 37 + (i - q * 36)];
  }
}

// Row mode: one thread per block ROW.  The thread walks the node's incidence list (sorted by neighbour, then
// element); each incident element's coefficients are generated ONCE per end (twice per element instead of the
// four times of the block-per-thread kernel), the off-diagonal quadrant is stored straight from registers
// (32-byte stores into the row's contiguous 288 B blocks) and the diagonal quadrant is accumulated in
// registers and written last.  No idle lanes while one lane sums a joint's 8-14 diagonal contributions,
// no atomics, fixed summation order (bit-reproducible).
// sm_100a has 256-bit global accesses (STG.E.256): nine per 288 B block instead of eighteen 128-bit ones.  A row
// thread's stores are scattered across lanes (one line per lane), so the LSU cost is per instruction, not per byte.
__device__ __forceinline__ void st256(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void ld256(const double* p, double& a, double& b, double& c, double& d) {
  asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p) : "memory");
}
__device__ __forceinline__ void store_block(double* __restrict__ dst, const double (&q)[36], bool accumulate) {
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    double a = q[4 * k], b = q[4 * k + 1], c = q[4 * k + 2], d = q[4 * k + 3];
    if (accumulate) {
      double oa, ob, oc, od;
      ld256(dst + 4 * k, oa, ob, oc, od);
      a += oa; b += ob; c += oc; d += od;
    }
    st256(dst + 4 * k, a, b, c, d);
  }
}

__global__ void __launch_bounds__(128) k_assemble_rows(
    const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
    const int32_t* __restrict__ en0, const int32_t* __restrict__ en1,
    const double* __restrict__ rad, const double* __restrict__ chain,
    const int32_t* __restrict__ adjptr, const int32_t* __restrict__ adj_other, const int32_t* __restrict__ adj_el,
    const int32_t* __restrict__ rowptr, int64_t n_nodes, double young, double nu, double kappa,
    int drad, double* __restrict__ vals) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_nodes) return;
  const int lo = adjptr[n], hi = adjptr[n + 1];
  if (hi == lo) return;                      // unconnected node: empty row
  int w = rowptr[n];                         // next block to write in this row
  int wdiag = -1;
  double dacc[36];
#pragma unroll
  for (int k = 0; k < 36; ++k) dacc[k] = 0.0;
  int prev = -1;
  for (int i = lo; i < hi; ++i) {
    const int other = adj_other[i];
    const int ee = adj_el[i];
    const int e = ee >> 1, end = ee & 1;
    const int a = en0[e], c = en1[e];
    const ElemCoef co = elem_coef(x[a], y[a], z[a], x[c], y[c], z[c], rad[e], young, nu, kappa, drad != 0);
    const double wt = (drad && chain) ? chain[e] : 1.0;
    elem_block_accum(co, end, end, wt, dacc);
    double q[36];
#pragma unroll
    for (int k = 0; k < 36; ++k) q[k] = 0.0;
    elem_block_accum(co, end, end ^ 1, wt, q);
    const bool dup = (other == prev);        // second strut between the same two nodes: add to the same block
    if (!dup) {
      if (wdiag < 0 && other > (int)n) { wdiag = w; ++w; }   // the diagonal block sits here in column order
      ++w;
    }
    store_block(vals + (int64_t)(w - 1) * 36, q, dup);
    prev = other;
  }
  if (wdiag < 0) wdiag = w;                  // all neighbours have smaller indices
  store_block(vals + (int64_t)wdiag * 36, dacc, false);
}

// Atomic mode: one thread per element, 4 quadrants scatter-added with FP64 RED.
// The two diagonal quadrants are warp-aggregated first: lanes whose quadrant
// targets the same BSR block (consecutive elements of a strut, struts of one
// joint) are summed with shuffles and only the group leader issues atomics.
__device__ __forceinline__ void scatter_quadrant(double* __restrict__ vals, int blk, bool valid,
                                                 const double (&q)[36], bool aggregate) {
  double* dst = vals + (int64_t)(valid ? blk : 0) * 36;
  const int lane = threadIdx.x & 31;
  if (!aggregate) {
    if (valid) {
#pragma unroll
      for (int k = 0; k < 36; ++k) atomicAdd(dst + k, q[k]);
    }
    return;
  }
  // every lane of the warp reaches this point (no early exit in the caller)
  const unsigned peers = __match_any_sync(0xffffffffu, valid ? blk : -1 - lane);
  const int leader = __ffs(peers) - 1;
  if (peers == (1u << lane)) {
    if (valid) {
#pragma unroll
      for (int k = 0; k < 36; ++k) atomicAdd(dst + k, q[k]);
    }
    return;
  }
#pragma unroll
  for (int k = 0; k < 36; ++k) {
    double s = 0.0;
    for (unsigned m = peers; m; m &= m - 1) s += __shfl_sync(peers, q[k], __ffs(m) - 1);
    if (lane == leader) atomicAdd(dst + k, s);
  }
}

__global__ void __launch_bounds__(128) k_assemble_atomic(
    const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
    const int32_t* __restrict__ en0, const int32_t* __restrict__ en1,
    const double* __restrict__ rad, const double* __restrict__ chain,
    const int32_t* __restrict__ elem_block, int64_t n_elem, double young, double nu, double kappa,
    int drad, double* __restrict__ vals) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = e < n_elem;
  const int64_t ee = valid ? e : 0;
  const int a = en0[ee], c = en1[ee];
  const ElemCoef co = elem_coef(x[a], y[a], z[a], x[c], y[c], z[c], rad[ee], young, nu, kappa, drad != 0);
  const double w = (drad && chain) ? chain[ee] : 1.0;
  const int4 blk = reinterpret_cast<const int4*>(elem_block)[ee];
  const int target[4] = {blk.x, blk.y, blk.z, blk.w};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    double acc[36];
#pragma unroll
    for (int k = 0; k < 36; ++k) acc[k] = 0.0;
    elem_block_accum(co, q >> 1, q & 1, w, acc);
    scatter_quadrant(vals, target[q], valid, acc, q == 0 || q == 3);
  }
}

extern "C" int lat_assemble_bsr(lat_ctx* ctx, const double* x, const double* y, const double* z,
                                const int32_t* en0, const int32_t* en1, const double* rad,
                                const double* chain, int64_t n_elem, int64_t n_nodes, double young,
                                double nu, double kappa, int mode, int drad, double* vals) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, x && y && z && en0 && en1 && rad && vals);
  LAT_CHECK_ARG(ctx, mode == LAT_ASM_GATHER || mode == LAT_ASM_ATOMIC || mode == LAT_ASM_ROWS);
  if (ctx->pat_nnzb < 0 || ctx->pat_nelem != n_elem || ctx->pat_nnodes != n_nodes)
    return lat_fail(ctx, LAT_ERR_STATE, "resident pattern does not match this mesh: call lat_bsr_pattern_build", __FILE__, __LINE__);
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t nz = ctx->pat_nnzb;
  // the row kernel uses 256-bit stores; a misaligned values array takes the (bit-identical) gather kernel
  if (mode == LAT_ASM_ROWS && (reinterpret_cast<uintptr_t>(vals) & 31) != 0) mode = LAT_ASM_GATHER;
  if (mode == LAT_ASM_ROWS) {
    LAT_LAUNCH(ctx, k_assemble_rows, (unsigned)ceil_div(n_nodes, 128), 128, 0, x, y, z, en0, en1, rad, chain,
               (const int32_t*)ctx->bufs["pat_adjptr"].p, (const int32_t*)ctx->bufs["pat_adj_other"].p,
               (const int32_t*)ctx->bufs["pat_adj_el"].p, (const int32_t*)ctx->bufs["pat_rowptr"].p, n_nodes, young, nu,
               kappa, drad, vals);
  } else if (mode == LAT_ASM_GATHER) {
    LAT_LAUNCH(ctx, k_assemble_gather, (unsigned)ceil_div(nz, ASM_BLOCK), ASM_BLOCK, 0, x, y, z, en0, en1, rad,
               chain, (const int32_t*)ctx->bufs["pat_blk_lo"].p, (const int32_t*)ctx->bufs["pat_blk_hi"].p,
               (const int32_t*)ctx->bufs["pat_blockrow"].p, (const int32_t*)ctx->bufs["pat_colidx"].p,
               (const int32_t*)ctx->bufs["pat_adj_el"].p, nz, young, nu, kappa, drad, vals);
  } else {
    LAT_CUDA(ctx, cudaMemsetAsync(vals, 0, (size_t)nz * 36 * sizeof(double), ctx->stream));
    LAT_LAUNCH(ctx, k_assemble_atomic, (unsigned)ceil_div(n_elem, 128), 128, 0, x, y, z, en0, en1, rad, chain,
               (const int32_t*)ctx->bufs["pat_elem_block"].p, n_elem, young, nu, kappa, drad, vals);
  }
  return LAT_OK;
}

// ===========================================================================
// A4: Dirichlet elimination
// ===========================================================================
__global__ void k_node_mask(const uint8_t* __restrict__ fixed, int64_t n_nodes, uint8_t* __restrict__ mask) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_nodes) return;
  unsigned m = 0;
#pragma unroll
  for (int d = 0; d < 6; ++d) m |= (fixed[n * 6 + d] ? 1u : 0u) << d;
  mask[n] = (uint8_t)m;
}

__global__ void __launch_bounds__(256) k_dirichlet_vals(const int32_t* __restrict__ blockrow,
                                                        const int32_t* __restrict__ colidx,
                                                        const uint8_t* __restrict__ mask, int64_t nnzb,
                                                        const double* __restrict__ vin, double* __restrict__ vout) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nnzb * 36) return;
  const int64_t b = i / 36;
  const int k = (int)(i - b * 36);
  const int a = k / 6, c = k - a * 6;
  const int rn = blockrow[b], cn = colidx[b];
  const unsigned rm = mask[rn], cm = mask[cn];
  double v = vin[i];
  if (((rm >> a) & 1u) | ((cm >> c) & 1u)) v = (rn == cn && a == c) ? 1.0 : 0.0;
  vout[i] = v;
}

// Same elimination, one thread per 6x6 BLOCK with nine 256-bit loads and stores (the per-entry kernel above
// spends its time on index arithmetic: 3.0 TB/s).  In place, a block that touches no constrained DOF is not
// read at all.  vin may alias vout.
__global__ void __launch_bounds__(128) k_dirichlet_blocks(const int32_t* __restrict__ blockrow,
                                                          const int32_t* __restrict__ colidx,
                                                          const uint8_t* __restrict__ mask, int64_t nnzb,
                                                          const double* vin, double* vout) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nnzb) return;
  const int rn = blockrow[b], cn = colidx[b];
  const unsigned rm = mask[rn], cm = mask[cn];
  if (vin == vout && (rm | cm) == 0u) return;
  const double* src = vin + b * 36;
  double* dst = vout + b * 36;
  double v[36];
#pragma unroll
  for (int k = 0; k < 9; ++k)
    asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];"
                 : "=d"(v[4 * k]), "=d"(v[4 * k + 1]), "=d"(v[4 * k + 2]), "=d"(v[4 * k + 3]) : "l"(src + 4 * k) : "memory");
  if (rm | cm) {
#pragma unroll
    for (int e = 0; e < 36; ++e) {
      const int a = e / 6, c = e - a * 6;
      if (((rm >> a) & 1u) | ((cm >> c) & 1u)) v[e] = (rn == cn && a == c) ? 1.0 : 0.0;
    }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k)
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * k), "d"(v[4 * k]), "d"(v[4 * k + 1]),
                 "d"(v[4 * k + 2]), "d"(v[4 * k + 3]) : "memory");
}

__global__ void k_blockrow_from_rowptr(const int32_t* __restrict__ rowptr, int64_t n_nodes, int32_t* __restrict__ blockrow) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_nodes) return;
  for (int j = rowptr[n]; j < rowptr[n + 1]; ++j) blockrow[j] = (int32_t)n;
}

__global__ void k_dirichlet_rhs(const uint8_t* __restrict__ fixed, const double* __restrict__ g,
                                const double* __restrict__ f, const double* __restrict__ Kg, int64_t n,
                                double* __restrict__ b) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  b[i] = fixed[i] ? g[i] : (f[i] - Kg[i]);
}

__global__ void k_mask_vec(const uint8_t* __restrict__ fixed, const double* __restrict__ g, int64_t n, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = fixed[i] ? g[i] : 0.0;
}

__global__ void k_set_bc(const uint8_t* __restrict__ fixed, const double* __restrict__ g, int64_t n, double* __restrict__ u) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && fixed[i]) u[i] = g[i];
}

extern "C" int lat_set_dirichlet_values(lat_ctx* ctx, const uint8_t* fixed, const double* g, int64_t n_dof, double* u) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, fixed && g && u && n_dof > 0);
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  LAT_LAUNCH(ctx, k_set_bc, (unsigned)ceil_div(n_dof, 256), 256, 0, fixed, g, n_dof, u);
  return LAT_OK;
}

int lat_spmv_internal(lat_ctx* ctx, const int32_t* rowptr, const int32_t* colidx, const double* vals,
                      int64_t n_nodes, const double* x, double* y);

extern "C" int lat_apply_dirichlet(lat_ctx* ctx, const int32_t* rowptr, const int32_t* colidx,
                                   int64_t n_nodes, const double* vals, const uint8_t* fixed,
                                   const double* g, const double* f, double* vals_bc, double* b) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, rowptr && colidx && vals && fixed && g && f && n_nodes > 0);
  LAT_CHECK_ARG(ctx, vals_bc || b);
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t n = 6 * n_nodes;
  if (b) {
    // lifting with the UNCONSTRAINED operator: b = f - K g_c, then b[c] = g[c]
    double* gm = lat_buf<double>(ctx, "bc_gmask", n);
    double* Kg = lat_buf<double>(ctx, "bc_Kg", n);
    if (!gm || !Kg) return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
    LAT_LAUNCH(ctx, k_mask_vec, (unsigned)ceil_div(n, 256), 256, 0, fixed, g, n, gm);
    int rc = lat_spmv_internal(ctx, rowptr, colidx, vals, n_nodes, gm, Kg);
    if (rc) return rc;
    LAT_LAUNCH(ctx, k_dirichlet_rhs, (unsigned)ceil_div(n, 256), 256, 0, fixed, g, f, Kg, n, b);
  }
  if (vals_bc) {
    int32_t nz32 = 0;
    LAT_CUDA(ctx, cudaMemcpyAsync(&nz32, rowptr + n_nodes, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    LAT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int64_t nz = nz32;
    uint8_t* mask = lat_buf<uint8_t>(ctx, "bc_nodemask", n_nodes);
    int32_t* blockrow = lat_buf<int32_t>(ctx, "bc_blockrow", nz);
    if (!mask || !blockrow) return lat_fail(ctx, 2, "workspace allocation failed", __FILE__, __LINE__);
    LAT_LAUNCH(ctx, k_node_mask, (unsigned)ceil_div(n_nodes, 256), 256, 0, fixed, n_nodes, mask);
    LAT_LAUNCH(ctx, k_blockrow_from_rowptr, (unsigned)ceil_div(n_nodes, 256), 256, 0, rowptr, n_nodes, blockrow);
    if (((reinterpret_cast<uintptr_t>(vals) | reinterpret_cast<uintptr_t>(vals_bc)) & 31) == 0)
      LAT_LAUNCH(ctx, k_dirichlet_blocks, (unsigned)ceil_div(nz, 128), 128, 0, blockrow, colidx, mask, nz, vals, vals_bc);
    else
      LAT_LAUNCH(ctx, k_dirichlet_vals, (unsigned)ceil_div(nz * 36, 256), 256, 0, blockrow, colidx, mask, nz, vals, vals_bc);
  }
  return LAT_OK;
}

// ===========================================================================
// A11: compliance sensitivity
// ===========================================================================
// q_e = L [ dES (t.dw/L)^2 + dGS (|s|^2 - (t.s)^2) + dGJ (t.dth/L)^2
//           + dEI (|dth/L|^2 - (t.dth/L)^2) ],   s = dw/L - thbar x t
// (frame independent for a circular section; equals u_e^T dK_e/dr u_e).
// With a second vector lambda the bilinear form lambda_e^T dK_e/dr u_e is used.
__global__ void __launch_bounds__(128) k_compliance_grad(
    const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
    const int32_t* __restrict__ en0, const int32_t* __restrict__ en1,
    const double* __restrict__ rad, const double* __restrict__ chain,
    const int32_t* __restrict__ group, int64_t n_elem, double young, double nu, double kappa,
    const double* __restrict__ u, const double* __restrict__ lam, int64_t n_groups,
    double* __restrict__ g, double* __restrict__ q_elem) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int grp = -1;
  double q = 0.0;
  if (e < n_elem) {
    grp = group ? group[e] : 0;
    const int a = en0[e], b = en1[e];
    const double dx = x[b] - x[a], dy = y[b] - y[a], dz = z[b] - z[a];
    const double L = sqrt(dx * dx + dy * dy + dz * dz), iL = 1.0 / L;
    const double t[3] = {dx * iL, dy * iL, dz * iL};
    const double r = rad[e];
    const double PI = 3.14159265358979323846;
    const double G = young / (2.0 * (1.0 + nu));
    const double dS = 2.0 * PI * r, dI = PI * r * r * r;
    const double dES = young * dS, dGS = G * kappa * dS, dEI = young * dI, dGJ = G * 2.0 * dI;
    double eps[2][4];  // per vector: axial, torsion (scalars) -- shear/bending handled via dot products
    double sv[2][3], kv[2][3];
    const double* vec[2] = {u, lam ? lam : u};
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const double* p = vec[v];
      double w0[3], th0[3], w1[3], th1[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        w0[k] = p[(int64_t)a * 6 + k]; th0[k] = p[(int64_t)a * 6 + 3 + k];
        w1[k] = p[(int64_t)b * 6 + k]; th1[k] = p[(int64_t)b * 6 + 3 + k];
      }
      double dw[3], dth[3], tb[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        dw[k] = (w1[k] - w0[k]) * iL; dth[k] = (th1[k] - th0[k]) * iL; tb[k] = 0.5 * (th0[k] + th1[k]);
      }
      // s = dw/L - thbar x t
      sv[v][0] = dw[0] - (tb[1] * t[2] - tb[2] * t[1]);
      sv[v][1] = dw[1] - (tb[2] * t[0] - tb[0] * t[2]);
      sv[v][2] = dw[2] - (tb[0] * t[1] - tb[1] * t[0]);
      kv[v][0] = dth[0]; kv[v][1] = dth[1]; kv[v][2] = dth[2];
      eps[v][0] = t[0] * dw[0] + t[1] * dw[1] + t[2] * dw[2];
      eps[v][1] = t[0] * sv[v][0] + t[1] * sv[v][1] + t[2] * sv[v][2];
      eps[v][2] = t[0] * dth[0] + t[1] * dth[1] + t[2] * dth[2];
      eps[v][3] = 0.0;
    }
    const double ss = sv[0][0] * sv[1][0] + sv[0][1] * sv[1][1] + sv[0][2] * sv[1][2];
    const double kk = kv[0][0] * kv[1][0] + kv[0][1] * kv[1][1] + kv[0][2] * kv[1][2];
    q = L * (dES * eps[0][0] * eps[1][0] + dGS * (ss - eps[0][1] * eps[1][1]) +
             dGJ * eps[0][2] * eps[1][2] + dEI * (kk - eps[0][2] * eps[1][2]));
    if (chain) q *= chain[e];
    if (q_elem) q_elem[e] = -q;
    if (grp >= n_groups) grp = -1;
  }
  // warp-segmented reduction: lanes with the same group are summed, the leader issues one atomic
  const unsigned act = __activemask();
  const unsigned peers = __match_any_sync(act, grp);
  const int lane = threadIdx.x & 31;
  double s = 0.0;
  for (unsigned m = peers; m; m &= m - 1) s += __shfl_sync(peers, q, __ffs(m) - 1);
  if (grp >= 0 && lane == __ffs(peers) - 1) atomicAdd(&g[grp], -s);
}

extern "C" int lat_compliance_grad(lat_ctx* ctx, const double* x, const double* y, const double* z,
                                   const int32_t* en0, const int32_t* en1, const double* rad,
                                   const double* chain, const int32_t* group, int64_t n_elem,
                                   double young, double nu, double kappa, const double* u,
                                   const double* lambda, int64_t n_groups, double* g, double* q_elem) {
  if (!ctx) return LAT_ERR_ARG;
  LAT_CHECK_ARG(ctx, x && y && z && en0 && en1 && rad && u && g && n_groups > 0 && n_elem >= 0);
  LAT_CUDA(ctx, cudaSetDevice(ctx->device));
  LAT_CUDA(ctx, cudaMemsetAsync(g, 0, n_groups * sizeof(double), ctx->stream));
  if (n_elem == 0) return LAT_OK;
  LAT_LAUNCH(ctx, k_compliance_grad, (unsigned)ceil_div(n_elem, 128), 128, 0, x, y, z, en0, en1, rad, chain, group,
             n_elem, young, nu, kappa, u, lambda, n_groups, g, q_elem);
  return LAT_OK;
}

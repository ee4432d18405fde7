// Standalone micro-benchmark used to choose the SpMV thread mapping (not part of the library).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o spmv_bench spmv_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

__global__ void k_read(const double2* __restrict__ v, size_t n2, double* out) {
  double s = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
    double2 a = __ldcs(v + i); s += a.x + a.y;
  }
  if (s == 123.456) out[0] = s;
}

// V0/V1/V2: 6 lanes per row, 5 rows per warp (library v1 mapping)
template <int MODE>
__global__ void __launch_bounds__(256) k_v1(const int* __restrict__ rowptr, const int* __restrict__ colidx,
                                             const double* __restrict__ vals, int n_nodes, const double* __restrict__ x,
                                             const double* __restrict__ x2, double* __restrict__ y, double* __restrict__ y2) {
  const int lane = threadIdx.x & 31, g = lane / 6, r = lane - g * 6;
  const long warp = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long n = warp * 5 + g;
  if (g >= 5 || n >= n_nodes) return;
  const int lo = rowptr[n], hi = rowptr[n + 1];
  double acc = 0;
#pragma unroll 4
  for (int j = lo; j < hi; ++j) {
    const double2* vp = reinterpret_cast<const double2*>(vals + (long)j * 36 + r * 6);
    const double2 a0 = __ldcs(vp), a1 = __ldcs(vp + 1), a2 = __ldcs(vp + 2);
    if (MODE == 0) { acc += a0.x + a0.y + a1.x + a1.y + a2.x + a2.y; continue; }
    const int c = __ldg(colidx + j);
    const double2* xp = reinterpret_cast<const double2*>(x + (long)c * 6);
    double2 x0 = __ldg(xp), x1 = __ldg(xp + 1), x2v = __ldg(xp + 2);
    if (MODE == 2) {
      const double2* qp = reinterpret_cast<const double2*>(x2 + (long)c * 6);
      double2 q0 = __ldg(qp), q1 = __ldg(qp + 1), q2 = __ldg(qp + 2);
      x0.x += 0.5 * q0.x; x0.y += 0.5 * q0.y; x1.x += 0.5 * q1.x; x1.y += 0.5 * q1.y; x2v.x += 0.5 * q2.x; x2v.y += 0.5 * q2.y;
    }
    acc += a0.x * x0.x + a0.y * x0.y + a1.x * x1.x + a1.y * x1.y + a2.x * x2v.x + a2.y * x2v.y;
  }
  y[n * 6 + r] = acc;
  if (MODE == 2) y2[n * 6 + r] = acc * 0.5;
}

// V3: warp streams its rows' blocks flat: lane l loads double2 #l of an 18-double2 block pair layout.
// Each warp handles ROWS consecutive rows; the blocks of those rows are contiguous -> the warp reads
// them with perfectly coalesced 512 B requests (32 lanes x 16 B), multiplies by the gathered x and
// does a segmented reduction: 3 lanes per matrix row-of-block, blocks summed per row via smem atomics-free pass.
template <int MODE>
__global__ void __launch_bounds__(256) k_flat(const int* __restrict__ rowptr, const int* __restrict__ colidx,
                                               const double* __restrict__ vals, int n_nodes, const double* __restrict__ x,
                                               const double* __restrict__ x2, double* __restrict__ y, double* __restrict__ y2) {
  // CTA handles 64 rows; per-row accumulators in shared memory (6 doubles each)
  __shared__ double s_y[64 * 6];
  __shared__ int s_rp[65];
  const int row0 = blockIdx.x * 64;
  const int nrow = min(64, n_nodes - row0);
  for (int i = threadIdx.x; i <= nrow; i += 256) s_rp[i] = rowptr[row0 + i];
  for (int i = threadIdx.x; i < 64 * 6; i += 256) s_y[i] = 0.0;
  __syncthreads();
  const int b0 = s_rp[0], b1 = s_rp[nrow];
  const long q0 = (long)b0 * 18, q1 = (long)b1 * 18;  // double2 units
  const double2* v2 = reinterpret_cast<const double2*>(vals);
  for (long q = q0 + threadIdx.x; q < q1; q += 256) {
    const double2 a = __ldcs(v2 + q);
    const int blk = (int)(q / 18), w = (int)(q - (long)blk * 18);   // w: 0..17 -> row w/3, col pair w%3
    const int rr = w / 3, cp = w - rr * 3;
    double v;
    if (MODE == 0) v = a.x + a.y;
    else {
      const int c = __ldg(colidx + blk);
      const double2 xv = __ldg(reinterpret_cast<const double2*>(x + (long)c * 6) + cp);
      v = a.x * xv.x + a.y * xv.y;
      if (MODE == 2) { const double2 qv = __ldg(reinterpret_cast<const double2*>(x2 + (long)c * 6) + cp); v += 0.5 * (a.x * qv.x + a.y * qv.y); }
    }
    // find row of blk by binary search in s_rp
    int lo = 0, hi = nrow;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (s_rp[mid] <= blk) lo = mid; else hi = mid; }
    atomicAdd(&s_y[lo * 6 + rr], v);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nrow * 6; i += 256) { y[(long)row0 * 6 + i] = s_y[i]; if (MODE == 2) y2[(long)row0 * 6 + i] = 0.5 * s_y[i]; }
}

// Chunked variant: CH blocks of a row are loaded back to back (clamped indices, zero weights past the end)
// before any is consumed -> CH*6 independent 16 B loads in flight per lane even for 3-block rows.
template <int CH, int DOTS>
__global__ void __launch_bounds__(256) k_chunk(const int* __restrict__ rowptr, const int* __restrict__ colidx,
                                                const double* __restrict__ vals, int n_nodes, const double* __restrict__ x,
                                                const double* __restrict__ rr, double* __restrict__ y, double* __restrict__ part) {
  const int lane = threadIdx.x & 31, g = lane / 6, r = lane - g * 6;
  const long warp = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long n = warp * 5 + g;
  const bool active = g < 5 && n < n_nodes;
  double acc = 0, d0 = 0, d1 = 0, d2 = 0;
  if (active) {
    const int lo = rowptr[n], hi = rowptr[n + 1];
    for (int j = lo; j < hi; j += CH) {
      double2 a[CH][3], xv[CH][3];
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        const int jj = min(j + k, hi - 1);
        const int c = __ldg(colidx + jj);
        const double2* vp = reinterpret_cast<const double2*>(vals + (long)jj * 36 + r * 6);
        const double2* xp = reinterpret_cast<const double2*>(x + (long)c * 6);
        a[k][0] = __ldcs(vp); a[k][1] = __ldcs(vp + 1); a[k][2] = __ldcs(vp + 2);
        xv[k][0] = __ldg(xp); xv[k][1] = __ldg(xp + 1); xv[k][2] = __ldg(xp + 2);
      }
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        double t = a[k][0].x * xv[k][0].x + a[k][0].y * xv[k][0].y + a[k][1].x * xv[k][1].x + a[k][1].y * xv[k][1].y +
                   a[k][2].x * xv[k][2].x + a[k][2].y * xv[k][2].y;
        acc += (j + k < hi) ? t : 0.0;
      }
    }
    const long i = n * 6 + r;
    y[i] = acc;
    if (DOTS) { const double u = x[i], rv = rr[i]; d0 = rv * u; d1 = acc * u; d2 = rv * rv; }
  }
  if (DOTS) {
    // block reduction + partials (no last-block pass here: measures the cheap part of the epilogue)
    __shared__ double sp[3][8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { d0 += __shfl_xor_sync(~0u, d0, o); d1 += __shfl_xor_sync(~0u, d1, o); d2 += __shfl_xor_sync(~0u, d2, o); }
    if (lane == 0) { sp[0][threadIdx.x >> 5] = d0; sp[1][threadIdx.x >> 5] = d1; sp[2][threadIdx.x >> 5] = d2; }
    __syncthreads();
    if (threadIdx.x < 3) { double s2 = 0; for (int k = 0; k < 8; ++k) s2 += sp[threadIdx.x][k]; part[threadIdx.x * gridDim.x + blockIdx.x] = s2; }
  }
}

// Clone of the library's k_cg_spmv with switches: STATUS (dependent status-word check), MINB (min CTAs/SM),
// LATE (own-row loads after the loop instead of before).
template <int STATUS, int MINB, int LATE>
__global__ void __launch_bounds__(256, MINB) k_cgclone(const int* __restrict__ rowptr, const int* __restrict__ colidx,
                                                        const double* __restrict__ vals, int n_nodes, const double* __restrict__ u,
                                                        const double* __restrict__ r, double* __restrict__ w,
                                                        const int* __restrict__ status, double* __restrict__ part) {
  __shared__ double sp[3][8];
  const int lane = threadIdx.x & 31, g = lane / 6, rr_ = lane - g * 6;
  const long warp = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const long n = warp * 5 + g;
  const bool active = g < 5 && n < n_nodes;
  int lo = 0, hi = 0; double uo = 0, ro = 0;
  const long i = n * 6 + rr_;
  if (active) { lo = __ldg(rowptr + n); hi = __ldg(rowptr + n + 1); if (!LATE) { uo = u[i]; ro = r[i]; } }
  if (STATUS) { if (status[0] || status[1] >= 1000000) return; }
  double acc = 0;
#pragma unroll 4
  for (int j = lo; j < hi; ++j) {
    const int c = __ldg(colidx + j);
    const double2* vp = reinterpret_cast<const double2*>(vals + (long)j * 36 + rr_ * 6);
    const double2* xp = reinterpret_cast<const double2*>(u + (long)c * 6);
    const double2 a0 = __ldcs(vp), a1 = __ldcs(vp + 1), a2 = __ldcs(vp + 2);
    const double2 x0 = xp[0], x1 = xp[1], x2 = xp[2];
    acc += a0.x * x0.x + a0.y * x0.y + a1.x * x1.x + a1.y * x1.y + a2.x * x2.x + a2.y * x2.y;
  }
  if (active) { w[i] = acc; if (LATE) { uo = u[i]; ro = r[i]; } }
  double d0 = ro * uo, d1 = acc * uo, d2 = ro * ro;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { d0 += __shfl_xor_sync(~0u, d0, o); d1 += __shfl_xor_sync(~0u, d1, o); d2 += __shfl_xor_sync(~0u, d2, o); }
  if (lane == 0) { sp[0][threadIdx.x >> 5] = d0; sp[1][threadIdx.x >> 5] = d1; sp[2][threadIdx.x >> 5] = d2; }
  __syncthreads();
  if (threadIdx.x < 3) { double s2 = 0; for (int k = 0; k < 8; ++k) s2 += sp[threadIdx.x][k]; part[threadIdx.x * gridDim.x + blockIdx.x] = s2; }
}

// Persistent variant: grid = SMs x k CTAs; each CTA owns a contiguous range of rows (equal nnz), warps walk it
// in tiles of 5 rows.  MODE as above.
template <int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS) k_persist(const int* __restrict__ rowptr, const int* __restrict__ colidx,
                                                      const double* __restrict__ vals, int n_nodes, const int* __restrict__ cta_row0,
                                                      const double* __restrict__ x, const double* __restrict__ x2,
                                                      double* __restrict__ y, double* __restrict__ y2) {
  const int lane = threadIdx.x & 31, g = lane / 6, r = lane - g * 6;
  const int r0 = cta_row0[blockIdx.x], r1 = cta_row0[blockIdx.x + 1];
  for (int base = r0 + (threadIdx.x >> 5) * 5; base < r1; base += (THREADS / 32) * 5) {
    const int n = base + g;
    if (g >= 5 || n >= r1) continue;
    const int lo = rowptr[n], hi = rowptr[n + 1];
    double acc = 0;
#pragma unroll 4
    for (int j = lo; j < hi; ++j) {
      const double2* vp = reinterpret_cast<const double2*>(vals + (long)j * 36 + r * 6);
      const double2 a0 = __ldcs(vp), a1 = __ldcs(vp + 1), a2 = __ldcs(vp + 2);
      if (MODE == 0) { acc += a0.x + a0.y + a1.x + a1.y + a2.x + a2.y; continue; }
      const int c = __ldg(colidx + j);
      const double2* xp = reinterpret_cast<const double2*>(x + (long)c * 6);
      double2 x0 = __ldg(xp), x1 = __ldg(xp + 1), x2v = __ldg(xp + 2);
      acc += a0.x * x0.x + a0.y * x0.y + a1.x * x1.x + a1.y * x1.y + a2.x * x2v.x + a2.y * x2v.y;
    }
    y[(long)n * 6 + r] = acc;
  }
}

int main(int argc, char** argv) {
  int N = argc > 1 ? atoi(argv[1]) : 440000; int NB = argc > 2 ? atoi(argv[2]) : 9;
  std::vector<int> rp, ci;
  bool from_file = false;
  if (argc > 1 && atoi(argv[1]) == 0) {   // argv[1] = pattern file: int32 N, int32 nnzb, rowptr[N+1], colidx[nnzb]
    FILE* f = fopen(argv[1], "rb"); if (!f) { printf("cannot open %s\n", argv[1]); return 1; }
    int hdr[2]; if (fread(hdr, 4, 2, f) != 2) return 1; N = hdr[0];
    rp.resize(N + 1); ci.resize(hdr[1]);
    if (fread(rp.data(), 4, N + 1, f) != (size_t)N + 1 || fread(ci.data(), 4, hdr[1], f) != (size_t)hdr[1]) return 1;
    fclose(f); from_file = true;
  }
  if (!from_file) { rp.assign(N + 1, 0); }
  const bool lattice_like = NB < 0;
  if (!from_file)   // NB < 0: BCC m=2 like structure: first 21% rows have 9 blocks, the rest 3
  for (int i = 0; i < N; ++i) {
    std::vector<int> c;
    if (lattice_like) NB = (i < (int)(0.2124 * N)) ? 9 : 3;
    for (int k = 0; k < NB; ++k) { long o = (long)i + (k - NB / 2) * 997L; o = ((o % N) + N) % N; c.push_back((int)o); }
    std::sort(c.begin(), c.end()); c.erase(std::unique(c.begin(), c.end()), c.end());
    for (int v : c) ci.push_back(v); rp[i + 1] = (int)ci.size();
  }
  size_t nnzb = ci.size();
  double *vals, *x, *x2, *y, *y2; int *drp, *dci;
  CK(cudaMalloc(&vals, nnzb * 288)); CK(cudaMalloc(&x, N * 48)); CK(cudaMalloc(&x2, N * 48)); CK(cudaMalloc(&y, N * 48)); CK(cudaMalloc(&y2, N * 48));
  CK(cudaMalloc(&drp, (N + 1) * 4)); CK(cudaMalloc(&dci, nnzb * 4));
  CK(cudaMemset(vals, 0, nnzb * 288)); CK(cudaMemset(x, 0, N * 48)); CK(cudaMemset(x2, 0, N * 48));
  CK(cudaMemcpy(drp, rp.data(), (N + 1) * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dci, ci.data(), nnzb * 4, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double mb = nnzb * 292.0 / 1e6;
  printf("N=%d nnzb=%zu matrix=%.1f MB\n", N, nnzb, mb);
  auto timeit = [&](const char* name, auto fn, double bytes) {
    for (int i = 0; i < 3; ++i) fn();
    cudaEventRecord(e0); for (int i = 0; i < 20; ++i) fn(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 20;
    printf("%-28s %8.1f us  %7.0f GB/s\n", name, ms * 1e3, bytes / ms / 1e6);
  };
  double bytes_m = nnzb * 288.0, bytes_1 = nnzb * 292.0 + N * 100.0, bytes_2 = nnzb * 292.0 + N * (4 + 4 * 48.0);
  timeit("read-only stream", [&] { k_read<<<148 * 16, 256>>>((const double2*)vals, nnzb * 18, y); }, bytes_m);
  int grid = (N + 39) / 40;
  timeit("v1 matrix only", [&] { k_v1<0><<<grid, 256>>>(drp, dci, vals, N, x, x2, y, y2); }, bytes_m);
  timeit("v1 spmv", [&] { k_v1<1><<<grid, 256>>>(drp, dci, vals, N, x, x2, y, y2); }, bytes_1);
  timeit("v1 pcg-like (2 gathers)", [&] { k_v1<2><<<grid, 256>>>(drp, dci, vals, N, x, x2, y, y2); }, bytes_2);
  int gridf = (N + 63) / 64;
  timeit("flat matrix only", [&] { k_flat<0><<<gridf, 256>>>(drp, dci, vals, N, x, x2, y, y2); }, bytes_m);
  timeit("flat spmv", [&] { k_flat<1><<<gridf, 256>>>(drp, dci, vals, N, x, x2, y, y2); }, bytes_1);
  timeit("flat pcg-like", [&] { k_flat<2><<<gridf, 256>>>(drp, dci, vals, N, x, x2, y, y2); }, bytes_2);
  double* part; CK(cudaMalloc(&part, 3 * (grid + 1) * 8));
  double bytes_c = nnzb * 292.0 + N * (4 + 3 * 48.0);
  timeit("chunk2 spmv", [&] { k_chunk<2, 0><<<grid, 256>>>(drp, dci, vals, N, x, x2, y, part); }, bytes_1);
  timeit("chunk3 spmv", [&] { k_chunk<3, 0><<<grid, 256>>>(drp, dci, vals, N, x, x2, y, part); }, bytes_1);
  timeit("chunk4 spmv", [&] { k_chunk<4, 0><<<grid, 256>>>(drp, dci, vals, N, x, x2, y, part); }, bytes_1);
  timeit("chunk3 spmv + 3 dots", [&] { k_chunk<3, 1><<<grid, 256>>>(drp, dci, vals, N, x, x2, y, part); }, bytes_c);
  timeit("chunk4 spmv + 3 dots", [&] { k_chunk<4, 1><<<grid, 256>>>(drp, dci, vals, N, x, x2, y, part); }, bytes_c);
  int* dstat; CK(cudaMalloc(&dstat, 16)); CK(cudaMemset(dstat, 0, 16));
  timeit("cgclone status=1 minb=1 early", [&] { k_cgclone<1, 1, 0><<<grid, 256>>>(drp, dci, vals, N, x, x2, y, dstat, part); }, bytes_c);
  timeit("cgclone status=0 minb=1 early", [&] { k_cgclone<0, 1, 0><<<grid, 256>>>(drp, dci, vals, N, x, x2, y, dstat, part); }, bytes_c);
  timeit("cgclone status=1 minb=6 early", [&] { k_cgclone<1, 6, 0><<<grid, 256>>>(drp, dci, vals, N, x, x2, y, dstat, part); }, bytes_c);
  timeit("cgclone status=1 minb=8 early", [&] { k_cgclone<1, 8, 0><<<grid, 256>>>(drp, dci, vals, N, x, x2, y, dstat, part); }, bytes_c);
  timeit("cgclone status=1 minb=1 late", [&] { k_cgclone<1, 1, 1><<<grid, 256>>>(drp, dci, vals, N, x, x2, y, dstat, part); }, bytes_c);
  timeit("cgclone status=1 minb=6 late", [&] { k_cgclone<1, 6, 1><<<grid, 256>>>(drp, dci, vals, N, x, x2, y, dstat, part); }, bytes_c);
  if (argc > 3) return 0;
  // persistent variants
  for (int k : {2, 4}) {
    for (int threads : {256, 512}) {
      int G = 148 * k;
      std::vector<int> c0(G + 1);
      for (int i = 0; i <= G; ++i) {
        long target = (long)nnzb * i / G;
        int r = (int)(std::lower_bound(rp.begin(), rp.end(), (int)target) - rp.begin());
        c0[i] = i == G ? N : std::min(r, N);
      }
      int* dc0; CK(cudaMalloc(&dc0, (G + 1) * 4)); CK(cudaMemcpy(dc0, c0.data(), (G + 1) * 4, cudaMemcpyHostToDevice));
      char nm[64];
      snprintf(nm, 64, "persist k=%d t=%d matrix", k, threads);
      if (threads == 256) timeit(nm, [&] { k_persist<0, 256><<<G, 256>>>(drp, dci, vals, N, dc0, x, x2, y, y2); }, bytes_m);
      else timeit(nm, [&] { k_persist<0, 512><<<G, 512>>>(drp, dci, vals, N, dc0, x, x2, y, y2); }, bytes_m);
      snprintf(nm, 64, "persist k=%d t=%d spmv", k, threads);
      if (threads == 256) timeit(nm, [&] { k_persist<1, 256><<<G, 256>>>(drp, dci, vals, N, dc0, x, x2, y, y2); }, bytes_1);
      else timeit(nm, [&] { k_persist<1, 512><<<G, 512>>>(drp, dci, vals, N, dc0, x, x2, y, y2); }, bytes_1);
      cudaFree(dc0);
    }
  }
  return 0;
}

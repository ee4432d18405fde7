// pcg_persist.cuh -- the whole PCG solve as ONE persistent cooperative kernel, vector state on chip.  sm_100a.
//
// Why: at BASELINE config 1 (98 MB matrix, 0.49 M DOF) the three-kernel iteration (k_cg_update, k_cg_spmv,
// k_cg_reduce) spends ~40 % of its 35 us outside the matrix stream: 56 MB of vector traffic per iteration, three
// launch ramps / tails of ~2-5 us each and a one-CTA reduce kernel (profiles/r01_ncu_full_v4_kernels.txt).  A B200 has
// 148 x 227 KB = 33 MB of shared memory, enough to hold x, r, p, s and w of a 0.5-0.7 M DOF system.  So:
//
//   * one CTA per SM (cooperative launch); the block rows are dealt to the CTAs in chunks of 16 rows, round robin, so
//     every CTA holds the same mix of long rows (joints) and short rows (strut-interior nodes): with contiguous
//     byte-balanced ranges the joint-only CTAs needed 28 us per product and the others 18 us, and the reverse in the
//     update phase (profiles/r02_persist_trace_contiguous_partition.txt) -- and every phase ends in a grid barrier;
//   * r, p, s, w of the owned rows live in shared memory for the whole solve -- and so does the preconditioner while the
//     total stays below ~160 KB (the product phase needs the rest of the 256 KB as L1 for its gathers of u; at config 1
//     the Jacobi diagonal fits, the block-Jacobi rows do not and are read from L2 as full FP32 6x6 blocks, three 8-byte
//     loads per lane); the CTA's slice of rowptr and colidx is cached in shared memory too, so the product phase
//     issues only independent loads (matrix blocks, gathered u) -- no rowptr -> colidx -> u chain;
//   * ~44 MB of the matrix stay in L2 from one iteration to the next: the blocks of a fixed set of 16-row classes are
//     loaded with an evict_last policy, the rest streams evict_first, all past L1 (createpolicy +
//     ld.global.L1::no_allocate.L2::cache_hint): DRAM traffic 111 -> 40 MB per iteration, product phase 15.4 -> 11.5 us
//     (profiles/r02_persist_l2keep_ab.txt);
//   * x stays in global memory (one read-modify-write per iteration, nobody waits for it);
//   * only u = M^-1 r (gathered by other CTAs) lives in global memory, i.e. in L2;
//   * product and block-Jacobi preconditioner both use the transposed-piece lane mapping (piece_row / piece_finish
//     in lattice_solver.cu): 96 contiguous bytes per row group and instruction; the inverse diagonal blocks are
//     stored as full 6x6 blocks for this kernel so that u = M^-1 r is the same block product with x = r from shared
//     memory (the packed 21-entry layout of k_cg_update needs six scattered 8-byte loads per lane);
//   * per iteration two grid-wide synchronisations, both flag based (no atomics on a shared counter):
//       barrier A  (after the update phase: the new u is visible)  -- one epoch flag per CTA, everyone polls all flags;
//       barrier B  (after the product phase: dot products)         -- every CTA publishes its three partial sums as
//                  six 8-byte {32 data bits | 32 epoch bits} words (the "LL" trick of the peer-memory all-reduce),
//                  every CTA polls all words and adds them in CTA order: barrier and all-reduce in ONE L2 round
//                  trip, identical bits in every CTA, so all CTAs take the same branch (converged / restart);
//   * the Chronopoulos-Gear recurrences, the stop test, the true-residual safeguard and its restart run inside the
//     kernel; the host launches once and reads the status block.
//
// HBM traffic per iteration: the non-resident part of the matrix; the vectors never leave the SM.  Config 1: 23.0 us per
// iteration (product 11.5 | reduce / barrier B 3.4 | update 5.0 | barrier A 2.7) against 35.5 us for the three-kernel
// iteration of round 1.  All spins are bounded: a CTA that waits longer than ~1 s reports info = 4 and every CTA leaves.
//
// Multi-GPU (template DIST, slab partitions with <= 2 neighbours, every rank's slab on chip): the same kernel runs on
// every GPU, one cooperative launch per rank, and the two grid synchronisations of an iteration also carry the exchange:
//   * the update phase sends every boundary entry of the new u to the neighbour that mirrors it as two 8-byte
//     {32 data bits | 32 tag bits} words (tag = number of the production of u) into the neighbour's halo INBOX: data
//     and flag travel in one NVLink store, so the sender needs no system-scope release fence (one NVLink round trip
//     per CTA: the first version -- plain stores + fence.acq_rel.sys + a counter -- spent 10-13 us in barrier A
//     instead of 4.4, profiles/r02_persist_dist_trace.txt) and the receiver no acquire fence; at the start of barrier A
//     every CTA unpacks its share of the inbox into the ghost section of u (plain local stores, published by the
//     barrier's own gpu-scope release like the rest of u), so the barrier returns when the local u AND the ghosts are
//     complete;
//   * barrier B first reduces inside the GPU as before, then CTA 0 posts the three local totals to every rank's
//     mailbox (LL words, one hop) and every CTA of every rank adds the rank totals in rank order: identical bits on
//     all CTAs of all ranks, so all of them take the same branches and execute the same number of barriers;
//   * the true-residual check multiplies u := x (pushed like any other u), so x needs no ghosts.
// No NCCL call and no host round trip inside the solve; all cross-GPU spins are bounded (info = 4 on time-out).
//
// Included by lattice_solver.cu after PcgScalars / PcgParams / dot6 / load_precond / apply_precond and the P2P arena.
#pragma once

static constexpr int PERSIST_BLOCK = 1024;             // one CTA per SM
static constexpr int PERSIST_NW = PERSIST_BLOCK / 32;
static constexpr int PERSIST_NVEC = 4;                 // r, p, s, w in shared memory (x: global)
__host__ __device__ constexpr int persist_pc_width(int pc) { return pc == LAT_PC_BLOCK6 ? 21 : (pc == LAT_PC_JACOBI ? 6 : 0); }   // doubles per row
static constexpr int PERSIST_CHUNK = 16;               // rows per chunk of the round-robin row distribution
static constexpr long long PERSIST_SPIN_LIMIT = 1ll << 23;

struct PersistArgs {
  const int32_t* rowptr;
  const int32_t* colidx;
  const double* vals;
  int64_t n_nodes;
  const double* b;
  double* x;               // out: solution (written at the end and before a residual check)
  double* u;               // global: preconditioned residual, gathered by every CTA
  const double* dinv;      // preconditioner in the layout of k_precond_setup (copied to shared memory at start)
  const float* dinv_full;  // block-Jacobi rows NOT cached in shared memory: the inverse blocks as full 6x6 in FP32 (144 B:
                           // a lane fetches its row with three 8-byte loads, k_persist_expand_dinv); else nullptr.  A
                           // preconditioner needs no more than single precision; the rounded blocks stay symmetric.
  PcgScalars* sc;          // status block read by the host
  PcgParams prm;
  unsigned long long* mail;   // [G][8] LL words (6 used), zeroed before the launch
  unsigned int* flags;        // [G][PERSIST_INBOX_STRIDE] barrier-A inboxes, zeroed before the launch
  int rows_cap, blk_cap;      // shared-memory capacities (rows, blocks) the launch was sized for
  int pc_smem;                // 1: preconditioner rows cached in shared memory; 0: read from global every iteration
                              //    (shared memory and L1 share 256 KB per SM: above ~164 KB of shared memory the
                              //    product phase loses its L1 and slowed from 15.5 to 26 us at config 1)
  long long* trace;           // optional [G][trace_iters][9] SM clock stamps of CTA thread 0 (+ [G][2] rows, blocks)
  int trace_iters;
  int l2_keep;                // 0: plain loads; k = 1..16: the matrix blocks of k / 16 of the rows are loaded with an L2
                              // evict_last policy (they stay L2-resident from one iteration to the next), the rest evict_first
  // ---- multi-GPU (DIST = true): slab partition over NVLink peer memory, see "Multi-GPU" below.  u is the ghosted
  // vector inside this rank's arena ([owned | ghosts]); n_nodes counts the OWNED rows.
  int nranks, my_rank, n_nb;
  long long ghost_first[2];             // first ghost entry (6 x node) of neighbour k in my u
  int ghost_entries[2];                 // entries neighbour k sends me with every new u (6 x ghost nodes)
  unsigned long long push_base;         // productions of u completed by earlier solves (same on every rank): halo tags
  unsigned long long seq_base;          // epoch of this solve << 32 (flag of the rank-level mailbox words)
  const int2* push_dst;                 // per owned node: ghost node in neighbour 0 / 1 that mirrors it (-1: none)
  unsigned long long* peer_ll[2];       // the neighbours' halo inboxes (IPC-mapped): two LL words per entry of THEIR u
  const unsigned long long* my_ll;      // my halo inbox
  unsigned char* const* peers;          // all arenas (device array), for the rank-level all-reduce
};

// Rows of CTA c: chunks c, c + G, c + 2G, ... of PERSIST_CHUNK consecutive rows.
__host__ __device__ __forceinline__ int64_t persist_global_row(int lr, int cta, int G) {
  return ((int64_t)(lr / PERSIST_CHUNK) * G + cta) * PERSIST_CHUNK + (lr % PERSIST_CHUNK);
}
__host__ __device__ __forceinline__ int persist_rows_of(int64_t n_nodes, int cta, int G) {
  const int64_t n_chunks = (n_nodes + PERSIST_CHUNK - 1) / PERSIST_CHUNK;
  if (cta >= n_chunks) return 0;
  const int64_t mine = (n_chunks - 1 - cta) / G + 1;                  // chunks cta, cta+G, ...
  const int64_t last = cta + (mine - 1) * G;                          // my last chunk: partial if it is the global last
  const int64_t last_rows = (last == n_chunks - 1) ? n_nodes - last * PERSIST_CHUNK : PERSIST_CHUNK;
  return (int)((mine - 1) * PERSIST_CHUNK + last_rows);
}
// Packed symmetric inverse blocks (21 FP64 entries) -> full row-major 6x6 blocks in FP32.
__global__ void k_persist_expand_dinv(const double* __restrict__ packed, int64_t n_nodes, float* __restrict__ full) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_nodes * 36) return;
  const int64_t n = t / 36;
  const int e = (int)(t - n * 36), r = e / 6, c = e - r * 6;
  const int i = r < c ? r : c, j = r < c ? c : r;
  full[t] = (float)packed[n * 21 + (i * (11 - i)) / 2 + j];
}

// Shared-memory need of the launch: the largest number of rows / blocks any CTA holds.
__global__ void k_persist_caps(const int32_t* __restrict__ rowptr, int64_t n_nodes, int G, int32_t* __restrict__ maxima) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= G) return;
  const int nrows = persist_rows_of(n_nodes, c, G);
  int nblk = 0;
  for (int lr = 0; lr < nrows; lr += PERSIST_CHUNK) {
    const int64_t r0 = persist_global_row(lr, c, G);
    const int64_t r1 = r0 + PERSIST_CHUNK < n_nodes ? r0 + PERSIST_CHUNK : n_nodes;
    nblk += rowptr[r1] - rowptr[r0];
  }
  atomicMax(&maxima[0], nrows);
  atomicMax(&maxima[1], nblk);
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



// Barrier A: every global store of this CTA (the new u / x of its rows) is visible to every CTA afterwards.
// All-to-all flags with PRIVATE inboxes: CTA c stores its epoch into slot c of every CTA's inbox (G 4-byte stores by G
// threads, each behind its own release fence) and polls only its own inbox (5 lines nobody else reads).  The first
// version had one flag per CTA that all G x G pollers read: 5 hot lines, 5.5 us per barrier
// (profiles/r02_persist_trace_contiguous_partition.txt); the acquire fence afterwards drops the SM's L1 lines (CCTL.IVALL).
static constexpr int PERSIST_INBOX_STRIDE = 160;     // u32 slots per inbox row (G <= 160), 640 B = 5 lines
static constexpr long long PERSIST_SPIN_LIMIT_SYS = 1ll << 25;   // cross-GPU waits: the peers may start their kernel later
struct PersistShared {
  double red[3][PERSIST_NW];
  double loc[3];
  double tot[3];
  int timeout;
};
__device__ __forceinline__ unsigned int persist_tag(const PersistArgs& a, unsigned int epoch) {
  return ((unsigned int)(a.push_base + (unsigned long long)epoch) & 0x7fffffffu) | 0x80000000u;     // never 0
}
// DIST, receiver side: CTA c unpacks entries [c * per, (c + 1) * per) of each neighbour's part of the inbox.
__device__ __forceinline__ void persist_unpack(const PersistArgs& a, unsigned int tag, int G, PersistShared& sh) {
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    if (k >= a.n_nb) break;
    const int E = k == 0 ? a.ghost_entries[0] : a.ghost_entries[1];
    const long long first = k == 0 ? a.ghost_first[0] : a.ghost_first[1];
    const int per = (E + G - 1) / G;
    const int lo = blockIdx.x * per, hi = lo + per < E ? lo + per : E;
    for (int e = lo + threadIdx.x; e < hi; e += PERSIST_BLOCK) {
      const unsigned long long* src = a.my_ll + (size_t)(first + e) * 2;
      unsigned long long w0, w1;
      long long spins = 0;
      for (;;) {
        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(src) : "memory");
        if ((unsigned int)(w0 >> 32) == tag && (unsigned int)(w1 >> 32) == tag) break;
        if (++spins > PERSIST_SPIN_LIMIT_SYS) { sh.timeout = 1; break; }
        if (spins > 4096) __nanosleep(64);
      }
      a.u[first + e] = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
    }
  }
}
template <bool DIST>
__device__ __forceinline__ void persist_barrier(unsigned int* inbox, unsigned int epoch, int G, PersistShared& sh,
                                                long long* stamp, const PersistArgs& a) {
  __syncthreads();                                   // the CTA's stores are issued ...
  if (stamp && threadIdx.x == 0) stamp[0] = clock64();
  if (DIST) {
    persist_unpack(a, persist_tag(a, epoch), G, sh); // ... the ghosts this CTA is responsible for are in u ...
    __syncthreads();
  }
  if ((int)threadIdx.x < G) {
    asm volatile("fence.acq_rel.gpu;" ::: "memory");  // ... and ordered before this thread's flag (cumulative at gpu scope)
    st_relaxed_gpu_u32(inbox + (size_t)threadIdx.x * PERSIST_INBOX_STRIDE + blockIdx.x, epoch);
    if (stamp && threadIdx.x == 0) stamp[1] = clock64();
    long long spins = 0;
    while (ld_relaxed_gpu_u32(inbox + (size_t)blockIdx.x * PERSIST_INBOX_STRIDE + threadIdx.x) < epoch) {
      if (++spins > PERSIST_SPIN_LIMIT) { sh.timeout = 1; break; }
      if (spins > 4096) __nanosleep(64);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (stamp) stamp[2] = clock64();
    asm volatile("fence.acq_rel.gpu;" ::: "memory");   // acquire side: drops the SM's L1 lines (CCTL.IVALL)
    if (stamp) stamp[3] = clock64();
  }
  __syncthreads();
}

// Barrier B: grid-wide sum of three values, result in sh.tot[] for every thread of every CTA (same bits everywhere).
// One shared mailbox (56 lines read by every CTA).  Private inboxes as in barrier A were tried and are SLOWER here
// (148 x 888 eight-byte stores per reduction: 3.8 -> 6.2 us); so were L1-bypassing gathers without the acquire fence.
// DIST: the local totals then travel once over NVLink -- CTA 0 posts them into every rank's arena (the mail words of
// P2PArenaHdr: [parity][source rank][6], same LL format, flag = solve epoch + reduction number) and EVERY CTA of every
// rank adds the rank totals in rank order.
struct PersistRankMail { unsigned long long mail[2][16][6]; };     // = the head of P2PArenaHdr (static_assert there)
template <bool DIST>
__device__ __forceinline__ void persist_reduce(double (&v)[3], unsigned long long* mail, unsigned int epoch, int G,
                                               PersistShared& sh, double* s_scratch /* >= 3*G doubles */,
                                               const PersistArgs& a) {
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
  if (DIST && a.nranks > 1) {
    const unsigned long long seq = a.seq_base + (unsigned long long)epoch;
    // 12 bits of the solve epoch + 20 bits of the reduction number, never 0 (the arena starts zeroed)
    const unsigned long long rflag = ((((seq >> 32) & 0xfffull) << 20) | ((seq & 0xfffffull) + 1ull) | 0x80000000ull) << 32;
    const int par = (int)(epoch & 1u);
    const int nw = a.nranks * 6;
    if (blockIdx.x == 0 && (int)threadIdx.x < nw) {
      const int q = threadIdx.x / 6, w = threadIdx.x - q * 6;
      const unsigned long long bits = (unsigned long long)__double_as_longlong(sh.tot[w >> 1]);
      const unsigned long long half = (w & 1) ? (bits >> 32) : (bits & 0xffffffffull);
      PersistRankMail* hdr = reinterpret_cast<PersistRankMail*>(a.peers[q]);
      asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(&hdr->mail[par][a.my_rank][w]), "l"(rflag | half) : "memory");
    }
    __syncthreads();                                   // sh.tot is overwritten below
    if ((int)threadIdx.x < nw) {
      const int q = threadIdx.x / 6, w = threadIdx.x - q * 6;
      const PersistRankMail* mine = reinterpret_cast<const PersistRankMail*>(a.peers[a.my_rank]);
      unsigned long long word;
      long long spins = 0;
      for (;;) {
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(word) : "l"(&mine->mail[par][q][w]) : "memory");
        if ((word & 0xffffffff00000000ull) == rflag) break;
        if (++spins > PERSIST_SPIN_LIMIT_SYS) { sh.timeout = 1; break; }
        if (spins > 4096) __nanosleep(64);
      }
      reinterpret_cast<unsigned int*>(s_scratch)[(size_t)(q * 3 + (w >> 1)) * 2 + (w & 1)] = (unsigned int)(word & 0xffffffffull);
    }
    __syncthreads();
    if (threadIdx.x < 3) {
      double s = 0.0;
      for (int q = 0; q < a.nranks; ++q) s += s_scratch[q * 3 + threadIdx.x];     // rank order
      sh.tot[threadIdx.x] = s;
    }
    __syncthreads();
  }
}

// s_w = (A v) on the CTA's rows: six lanes per block row in the transposed-piece layout, three blocks (12 x 16 B loads
// per lane) in flight; lane (g, r) leaves the total of scalar row lane_dof(r) in s_w.  All warps call.
__device__ __forceinline__ void persist_product(const double* __restrict__ vals, const double* v, const int32_t* s_rp,
                                                const int32_t* s_gb, const int32_t* s_col, double* s_w, int nrows, int l2_keep = 0) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int g = lane / 6, rr_ = lane - g * 6;
  unsigned long long pol_keep = 0, pol_stream = 0;
  if (l2_keep > 0) {
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
  }
  for (int base = wid * 5; base < nrows; base += PERSIST_NW * 5) {
    const int lr = base + g;
    const bool active = g < 5 && lr < nrows;
    const int lp = active ? s_rp[lr] : 0, nb = active ? s_rp[lr + 1] - lp : 0;
    const int gb = active ? s_gb[lr] : 0;             // global index of the row's first block
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    if (l2_keep > 0) {
      // the same rows are kept in every iteration: 16-row chunks, lr / 16 cycles through 16 classes
      const unsigned long long pol = ((lr >> 4) & 15) < l2_keep ? pol_keep : pol_stream;
      piece_row_policy(s_col + (lp - gb), vals, gb, gb + nb, rr_, v, s0, s1, s2, pol);
    } else {
      piece_row<false, false>(s_col + (lp - gb), vals, gb, gb + nb, rr_, v, s0, s1, s2);
    }
    const double tot = piece_finish(g, rr_, s0, s1, s2);
    if (active) s_w[lr * 6 + lane_dof(rr_)] = tot;
  }
}

// z_d = (M^-1 v)_d for the lane's scalar row d of local row lr; M^-1 and v both in shared memory (no shuffles).
template <int PC>
__device__ __forceinline__ double persist_precond(const double* s_m, const double* s_v, int lr, int d) {
  if (PC == LAT_PC_NONE) return s_v[lr * 6 + d];
  if (PC == LAT_PC_JACOBI) return s_m[lr * 6 + d] * s_v[lr * 6 + d];
  const double* m = s_m + lr * 21;
  const double* v = s_v + lr * 6;
  double z = 0.0;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const int i = d < k ? d : k, j = d < k ? k : d;
    z = fma(m[(i * (11 - i)) / 2 + j], v[k], z);
  }
  return z;
}

// Same with the preconditioner row read from global memory (layout of k_precond_setup).
template <int PC>
__device__ __forceinline__ double persist_precond_global(const double* __restrict__ dinv, int64_t n, const double* s_v, int lr, int d) {
  if (PC == LAT_PC_NONE) return s_v[lr * 6 + d];
  if (PC == LAT_PC_JACOBI) return __ldg(dinv + n * 6 + d) * s_v[lr * 6 + d];
  const double* m = dinv + n * 21;
  const double* v = s_v + lr * 6;
  double z = 0.0;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const int i = d < k ? d : k, j = d < k ? k : d;
    z = fma(__ldg(m + (i * (11 - i)) / 2 + j), v[k], z);
  }
  return z;
}

// Global loads of one trip of the update phase (two row groups A, B): own u, x and -- when they are not cached in
// shared memory -- the preconditioner rows, all issued before the dependent math.  (Issuing the first trip's loads before
// the grid reduction was tried: the values spill across the reduction and the phase got slower, 30.2 -> 34.6 us.)
struct UpdLoads {
  double uA, uB, xA, xB;
  double mA[6], mB[6];
};
template <int PC>
__device__ __forceinline__ UpdLoads persist_upd_load(const PersistArgs& a, bool pcs, int64_t iA, int64_t iB, bool actA, bool actB, int dof) {
  UpdLoads L;
  L.uA = actA ? a.u[iA] : 0.0;
  L.uB = actB ? a.u[iB] : 0.0;
  L.xA = actA ? a.x[iA] : 0.0;
  L.xB = actB ? a.x[iB] : 0.0;
#pragma unroll
  for (int k = 0; k < 6; ++k) { L.mA[k] = 0.0; L.mB[k] = 0.0; }
  if (!pcs && PC == LAT_PC_BLOCK6) {
    // row `dof` of the full 6x6 inverse block: 24 contiguous bytes, three 8-byte loads.  (The packed 21-entry FP64
    // layout needs six scattered 8-byte loads per lane: 12 instructions x ~10 L1 wavefronts per warp and trip made the
    // block-Jacobi update 3.9 us slower than the Jacobi one; full FP64 blocks cost the product phase the L2 space they
    // take, profiles/r02_persist_l2keep_ab.txt.)
    const float2* pa = reinterpret_cast<const float2*>(a.dinv_full + (iA / 6) * 36 + dof * 6);
    const float2* pb = reinterpret_cast<const float2*>(a.dinv_full + (iB / 6) * 36 + dof * 6);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float2 va = actA ? __ldg(pa + k) : make_float2(0.f, 0.f);
      const float2 vb = actB ? __ldg(pb + k) : make_float2(0.f, 0.f);
      L.mA[2 * k] = (double)va.x; L.mA[2 * k + 1] = (double)va.y;
      L.mB[2 * k] = (double)vb.x; L.mB[2 * k + 1] = (double)vb.y;
    }
  }
  if (!pcs && PC == LAT_PC_JACOBI) { L.mA[0] = actA ? __ldg(a.dinv + iA) : 0.0; L.mB[0] = actB ? __ldg(a.dinv + iB) : 0.0; }
  return L;
}

// Store one entry of a freshly produced u: the owned copy and, for the nodes a neighbour mirrors, two LL words into
// that neighbour's halo inbox (slot = entry index in THEIR u).
template <bool DIST>
__device__ __forceinline__ void persist_store_u(const PersistArgs& a, const uint8_t* s_pflag, int lr, int64_t n, int dof, double val,
                                                unsigned int tag) {
  a.u[n * 6 + dof] = val;
  if (DIST && s_pflag[lr]) {
    const int2 d = a.push_dst[n];
    const unsigned long long bits = (unsigned long long)__double_as_longlong(val);
    const unsigned long long t = (unsigned long long)tag << 32;
    const unsigned long long w0 = t | (bits & 0xffffffffull), w1 = t | (bits >> 32);
    if (d.x >= 0) asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(a.peer_ll[0] + ((size_t)d.x * 6 + dof) * 2), "l"(w0), "l"(w1) : "memory");
    if (d.y >= 0) asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(a.peer_ll[1] + ((size_t)d.y * 6 + dof) * 2), "l"(w0), "l"(w1) : "memory");
  }
}

template <int PC, bool DIST = false>
__global__ void __launch_bounds__(PERSIST_BLOCK, 1) k_pcg_persist(PersistArgs a) {
  extern __shared__ __align__(16) unsigned char persist_smem[];
  __shared__ PersistShared sh;
  __shared__ int s_chunk_base[64];
  const int G = gridDim.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int g = lane / 6, rr_ = lane - g * 6;
  const int dof = lane_dof(rr_);                       // scalar row this lane owns in every phase
  const int cta = blockIdx.x;
  const int nrows = persist_rows_of(a.n_nodes, cta, G);
  const int nchunks = (nrows + PERSIST_CHUNK - 1) / PERSIST_CHUNK;
  const size_t vec = (size_t)a.rows_cap * 6;
  constexpr int PCW = persist_pc_width(PC);
  double* s_r = reinterpret_cast<double*>(persist_smem);
  double* s_p = s_r + vec;
  double* s_s = s_p + vec;
  double* s_w = s_s + vec;
  double* s_m = s_w + vec;                                          // [rows_cap][PCW] preconditioner rows
  double* s_scratch = s_m + (a.pc_smem ? (size_t)a.rows_cap * PCW : 0);   // 3 * G doubles
  int32_t* s_rp = reinterpret_cast<int32_t*>(s_scratch + 3 * G);    // [rows_cap + 1] local block offset of each row
  int32_t* s_gb = s_rp + a.rows_cap + 1;                            // [rows_cap]     global index of the row's first block
  int32_t* s_col = s_gb + a.rows_cap;                               // [blk_cap]      column indices in local block order
  uint8_t* s_pflag = reinterpret_cast<uint8_t*>(s_col + a.blk_cap); // [rows_cap]     DIST: row is mirrored by a neighbour
  if (threadIdx.x == 0) sh.timeout = 0;
  // local block offsets: per chunk (contiguous rows -> one subtraction), then a short serial scan
  for (int k = threadIdx.x; k < nchunks; k += PERSIST_BLOCK) {
    const int64_t r0 = persist_global_row(k * PERSIST_CHUNK, cta, G);
    const int64_t r1 = r0 + PERSIST_CHUNK < a.n_nodes ? r0 + PERSIST_CHUNK : a.n_nodes;
    s_chunk_base[k] = a.rowptr[r1] - a.rowptr[r0];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int k = 0; k < nchunks; ++k) { const int c = s_chunk_base[k]; s_chunk_base[k] = acc; acc += c; }
    s_rp[nrows] = acc;
  }
  __syncthreads();
  for (int lr = threadIdx.x; lr < nrows; lr += PERSIST_BLOCK) {
    const int64_t n = persist_global_row(lr, cta, G);
    const int64_t n0 = persist_global_row(lr - lr % PERSIST_CHUNK, cta, G);
    const int gb = a.rowptr[n];
    s_gb[lr] = gb;
    s_rp[lr] = s_chunk_base[lr / PERSIST_CHUNK] + (gb - a.rowptr[n0]);
  }
  __syncthreads();
  const int nblk = s_rp[nrows];
  for (int lr = wid; lr < nrows; lr += PERSIST_NW) {              // one warp per row: its column indices
    const int lp = s_rp[lr], nb = s_rp[lr + 1] - lp, gb = s_gb[lr];
    for (int k = lane; k < nb; k += 32) s_col[lp + k] = a.colidx[gb + k];
  }
  if (DIST) {
    for (int lr = threadIdx.x; lr < nrows; lr += PERSIST_BLOCK) {
      const int2 d = a.push_dst[persist_global_row(lr, cta, G)];
      s_pflag[lr] = (uint8_t)((d.x >= 0) | ((d.y >= 0) << 1));
    }
  }
  const double* __restrict__ vals = a.vals;
  double* __restrict__ u = a.u;
  const PcgParams prm = a.prm;
  const double tol2 = prm.tol * prm.tol;
  __syncthreads();

  // ---- init: preconditioner rows -> shared memory; x = 0, r = b, u = M^-1 b, p = s = 0
  const bool pcs = a.pc_smem != 0;
  if (PCW > 0 && pcs) {
    for (int lr = wid; lr < nrows; lr += PERSIST_NW) {
      const int64_t n = persist_global_row(lr, cta, G);
      for (int k = lane; k < PCW; k += 32) s_m[lr * PCW + k] = a.dinv[n * PCW + k];
    }
  }
  for (int base = wid * 5; base < nrows; base += PERSIST_NW * 5) {
    const int lr = base + g;
    if (g < 5 && lr < nrows) {
      const int64_t n = persist_global_row(lr, cta, G);
      const int li = lr * 6 + dof;
      s_r[li] = a.b[n * 6 + dof]; s_p[li] = 0.0; s_s[li] = 0.0;
      a.x[n * 6 + dof] = 0.0;
    }
  }
  __syncthreads();
  for (int base = wid * 5; base < nrows; base += PERSIST_NW * 5) {
    const int lr = base + g;
    if (g < 5 && lr < nrows) {
      const int64_t n = persist_global_row(lr, cta, G);
      persist_store_u<DIST>(a, s_pflag, lr, n, dof,
                            pcs ? persist_precond<PC>(s_m, s_r, lr, dof) : persist_precond_global<PC>(a.dinv, n, s_r, lr, dof),
                            persist_tag(a, 1u));
    }
  }
  unsigned int epochA = 0, epochB = 0;
  if (a.trace && threadIdx.x == 0) {
    a.trace[(size_t)G * a.trace_iters * 9 + 2 * blockIdx.x] = nrows;
    a.trace[(size_t)G * a.trace_iters * 9 + 2 * blockIdx.x + 1] = nblk;
  }
  persist_barrier<DIST>(a.flags, ++epochA, G, sh, nullptr, a);

  int first = 1, iters = 0, done = 0, breakdown = 0, restarts = 0, trace_it = 0;
  double gamma_old = 0.0, alpha = 0.0, beta = 0.0, bb = 0.0, rr = 0.0, true_rr = -1.0;
  for (;;) {
    if (sh.timeout) break;
    // ---- product phase: w = A u for the owned rows, partial (r,u), (w,u), (r,r)
    double v[3] = {0.0, 0.0, 0.0};
    const bool tracing = a.trace != nullptr && trace_it < a.trace_iters;
    long long* tr = tracing ? a.trace + ((size_t)blockIdx.x * a.trace_iters + trace_it) * 9 : nullptr;
    if (tracing && threadIdx.x == 0) tr[0] = clock64();
    persist_product(vals, u, s_rp, s_gb, s_col, s_w, nrows, a.l2_keep);
    // every lane reads back exactly the entries it wrote (same (row, dof) mapping): no barrier needed
    for (int base = wid * 5; base < nrows; base += PERSIST_NW * 5) {
      const int lr = base + g;
      if (g < 5 && lr < nrows) {
        const int e = lr * 6 + dof;
        const double ro = s_r[e], wv = s_w[e], uo = u[persist_global_row(lr, cta, G) * 6 + dof];
        v[0] = fma(ro, uo, v[0]);
        v[1] = fma(wv, uo, v[1]);
        v[2] = fma(ro, ro, v[2]);
      }
    }
    if (tracing) { __syncthreads(); if (threadIdx.x == 0) tr[1] = clock64(); }
    persist_reduce<DIST>(v, a.mail, ++epochB, G, sh, s_scratch, a);
    if (tracing && threadIdx.x == 0) tr[2] = clock64();
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
      // ---- true-residual safeguard: r_true = b - A x  (x is global and, after the last barrier A, visible everywhere)
      double t[3] = {0.0, 0.0, 0.0};
      if (DIST) {
        // x has no ghosts: multiply u := x, exchanged like every other u (u is rebuilt from r on a restart)
        for (int base = wid * 5; base < nrows; base += PERSIST_NW * 5) {
          const int lr = base + g;
          if (g < 5 && lr < nrows) {
            const int64_t n = persist_global_row(lr, cta, G);
            persist_store_u<DIST>(a, s_pflag, lr, n, dof, a.x[n * 6 + dof], persist_tag(a, epochA + 1u));
          }
        }
        persist_barrier<DIST>(a.flags, ++epochA, G, sh, nullptr, a);
        if (sh.timeout) break;
      }
      persist_product(vals, DIST ? (const double*)u : (const double*)a.x, s_rp, s_gb, s_col, s_w, nrows);
      for (int base = wid * 5; base < nrows; base += PERSIST_NW * 5) {
        const int lr = base + g;
        if (g < 5 && lr < nrows) {
          const int e = lr * 6 + dof;
          const double rt = a.b[persist_global_row(lr, cta, G) * 6 + dof] - s_w[e];
          s_w[e] = rt;
          t[0] = fma(rt, rt, t[0]);
        }
      }
      persist_reduce<DIST>(t, a.mail, ++epochB, G, sh, s_scratch, a);
      if (sh.timeout) break;
      true_rr = sh.tot[0];
      if (!(true_rr > 4.0 * tol2 * bb)) break;                       // accepted
      if (restarts >= 2 || iters >= prm.maxiter) { breakdown = 3; break; }
      // restart from x with r = r_true
      restarts += 1;
      for (int base = wid * 5; base < nrows; base += PERSIST_NW * 5) {
        const int lr = base + g;
        if (g < 5 && lr < nrows) { const int li = lr * 6 + dof; s_r[li] = s_w[li]; s_p[li] = 0.0; s_s[li] = 0.0; }
      }
      __syncthreads();
      for (int base = wid * 5; base < nrows; base += PERSIST_NW * 5) {
        const int lr = base + g;
        if (g < 5 && lr < nrows) {
          const int64_t n = persist_global_row(lr, cta, G);
          persist_store_u<DIST>(a, s_pflag, lr, n, dof,
                                pcs ? persist_precond<PC>(s_m, s_r, lr, dof) : persist_precond_global<PC>(a.dinv, n, s_r, lr, dof),
                                persist_tag(a, epochA + 1u));
        }
      }
      persist_barrier<DIST>(a.flags, ++epochA, G, sh, nullptr, a);
      done = 0;
      first = 2;
      continue;
    }
    // ---- update phase: p = u + beta p; s = w + beta s; x += alpha p; r -= alpha s; u = M^-1 r
    // two row groups per trip; the only global loads (own u, x) of both are issued first, the rest is shared memory
    const unsigned int utag = DIST ? persist_tag(a, epochA + 1u) : 0u;
    for (int base = wid * 5; base < nrows; base += 2 * PERSIST_NW * 5) {
      const int lrA = base + g, lrB = base + PERSIST_NW * 5 + g;
      const bool actA = g < 5 && lrA < nrows, actB = g < 5 && lrB < nrows;
      const int64_t iA = actA ? persist_global_row(lrA, cta, G) * 6 + dof : 0;
      const int64_t iB = actB ? persist_global_row(lrB, cta, G) * 6 + dof : 0;
      const int liA = lrA * 6 + dof, liB = lrB * 6 + dof;
      const UpdLoads L = persist_upd_load<PC>(a, pcs, iA, iB, actA, actB, dof);
      const double uA = L.uA, uB = L.uB, xA = L.xA, xB = L.xB;
      const double* mA = L.mA;
      const double* mB = L.mB;
      if (actA) {
        const double pv = fma(beta, s_p[liA], uA);
        const double sv = fma(beta, s_s[liA], s_w[liA]);
        s_p[liA] = pv;
        s_s[liA] = sv;
        a.x[iA] = fma(alpha, pv, xA);
        s_r[liA] = fma(-alpha, sv, s_r[liA]);
      }
      if (actB) {
        const double pv = fma(beta, s_p[liB], uB);
        const double sv = fma(beta, s_s[liB], s_w[liB]);
        s_p[liB] = pv;
        s_s[liB] = sv;
        a.x[iB] = fma(alpha, pv, xB);
        s_r[liB] = fma(-alpha, sv, s_r[liB]);
      }
      __syncwarp();                                     // the six entries of r of each node are in shared memory
      double zA = 0.0, zB = 0.0;
      if (pcs || PC == LAT_PC_NONE) {
        if (actA) zA = persist_precond<PC>(s_m, s_r, lrA, dof);
        if (actB) zB = persist_precond<PC>(s_m, s_r, lrB, dof);
      } else if (PC == LAT_PC_JACOBI) {
        if (actA) zA = mA[0] * s_r[liA];
        if (actB) zB = mB[0] * s_r[liB];
      } else {
#pragma unroll
        for (int k = 0; k < 6; ++k) { zA = fma(mA[k], s_r[lrA * 6 + k], zA); zB = fma(mB[k], s_r[lrB * 6 + k], zB); }
      }
      if (actA) persist_store_u<DIST>(a, s_pflag, lrA, iA / 6, dof, zA, utag);
      if (actB) persist_store_u<DIST>(a, s_pflag, lrB, iB / 6, dof, zB, utag);
    }
    if (tracing) { __syncthreads(); if (threadIdx.x == 0) tr[3] = clock64(); }
    persist_barrier<DIST>(a.flags, ++epochA, G, sh, tracing ? tr + 5 : nullptr, a);
    if (tracing && threadIdx.x == 0) tr[4] = clock64();
    ++trace_it;
  }
  __syncthreads();
  const int timed_out = sh.timeout;
  // ---- epilogue: status (x is already in global memory)
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
    sc->counter[0] = epochA;          // productions of u (= halo pushes) of this solve
  }
  if (timed_out && threadIdx.x == 0) a.sc->p2p_timeout = 1;
}

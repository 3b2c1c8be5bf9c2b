// Batched blocked LU with partial pivoting + in-place multi-RHS solve, FP64, sm_100a.
//
// Replaces the reference's `jnp.linalg.inv(X) @ Y` pairs (local_solve/_uniform_2D_DtN.py:257-259,
// merge/_schur_complement.py:146,222,234): X^-1 Y comes from P X = L U and two blocked triangular
// solves, never from an explicit inverse of X.
//
//   1. Factorisation, right-looking, NB=128 outer / IB=32 inner blocking, with LOOK-AHEAD: the
//      latency-bound chain (panel kernels, 32-wide inner updates, inversion of the 128x128 unit
//      lower block) runs on an internal high-priority stream one block column ahead of the
//      rank-128 DMMA trailing update on the caller's stream.
//   2. Row interchanges applied to the right-hand sides.
//   3. L Z = P B and U X = Z by RECURSIVE blocked substitution: diagonal 128-blocks are inverted
//      once (batched over all blocks), everything else is GEMMs whose K grows with the recursion
//      level, so the right-hand sides — the bulk of the flops in the HPS merges — run at the
//      large-K rate of the DMMA kernel and are read/written O(log) times instead of n/128 times.
//
// Two ways to factor a 128-wide block column:
//   SPECULATIVE (matrices that partial pivoting leaves alone below the diagonal block: the merges' D, and — with
//   threshold pivoting — the leaves' A_ii): diagblk_kernel (128 x 128 diagonal block, one CTA, the row panel in
//   registers) -> trtri_pair_kernel (both triangular inverses) -> ONE DMMA product L21 = A21 U11^-1 ->
//   spec_commit_kernel (commit + the check that makes it equivalent to (threshold) partial pivoting; info = -2 otherwise).
//   No cross-CTA exchange at all: 190 us instead of 830 us per block column at n = 19 200.
//   PIVOTED (everything else, and the fallback): blockcol_kernel (one CTA per matrix, <= 196 rows), blockcol2_kernel
//   (rows split over G co-scheduled CTAs — cluster for G <= 8, cooperative launch above — that keep their rows x 128
//   chunk in shared memory and elect each column's pivot through L2 with self-validating 16-byte units), panel_kernel
//   (IB-column panels for block columns too tall or too many to be co-resident; one group barrier per column).
// Other kernels:
//   laswp_kernel   row interchanges on a column range (rows are contiguous: coalesced).
//   inner_trsm     32x32 unit-lower solve inside the outer panel (panel_kernel path).
//   trtri kernels  invert NBxNB triangular diagonal blocks in shared memory (substitutions in registers) so that
//                  every triangular solve becomes a DMMA GEMM (in place, single tile row).
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace hps {

namespace {

constexpr int NB = 128;          // outer block
constexpr int IB = 32;           // inner panel width
constexpr int PANEL_ROWS = 768;  // rows of the panel one CTA keeps in shared memory
constexpr int PANEL_LD = IB + 1;
constexpr int PANEL_THREADS = 512;
constexpr int MAX_G = 128;

struct Cand {          // one CTA's pivot candidate for the current column
  double val;          // |a|, negative when the CTA has no eligible row
  int row;             // panel-relative row index
  int pad;
  double content[IB];  // that row's IB panel entries
};

struct PanelScratch {  // per matrix
  Cand cand[2][MAX_G];
  double diag[2][IB];
  unsigned int counter;  // SYNC_GRID barrier
  unsigned int pad[3];
};

enum { SYNC_NONE = 0, SYNC_CLUSTER = 1, SYNC_GRID = 2 };

struct PanelArgs {
  double* A; int64_t lda, sA;
  int n, jj, ib, G;
  int* ipiv;            // [batch][n]
  int* info;            // [batch]
  PanelScratch* scratch;  // [batch]
};

template <int SYNC>
__device__ __forceinline__ void group_barrier(PanelScratch* sc, int G, unsigned& epoch) {
  if (SYNC == SYNC_CLUSTER) {
    __threadfence();
    cg::this_cluster().sync();
  } else if (SYNC == SYNC_GRID) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(&sc->counter, 1u);
      const unsigned target = (epoch + 1u) * (unsigned)G;
      volatile unsigned* c = &sc->counter;
      while (*c < target) { __nanosleep(20); }
      __threadfence();
    }
    ++epoch;
    __syncthreads();
  }
}

template <int SYNC>
__global__ void __launch_bounds__(PANEL_THREADS, 1) panel_kernel(PanelArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = blockIdx.x, mat = blockIdx.y;
  const int G = a.G, ib = a.ib;
  const int rows = a.n - a.jj;
  const int rpc = (rows + G - 1) / G;
  const int r0 = min(rows, g * rpc), r1 = min(rows, r0 + rpc), nr = r1 - r0;
  double* A = a.A + (int64_t)mat * a.sA + (int64_t)a.jj * a.lda + a.jj;
  PanelScratch* sc = a.scratch + mat;
  int* ipiv = a.ipiv + (int64_t)mat * a.n + a.jj;

  double* tile = sm;                                   // [nr][PANEL_LD]
  double* prow = sm + (size_t)PANEL_ROWS * PANEL_LD;   // [IB]
  double* red_val = prow + IB;                         // [16]
  int* red_idx = reinterpret_cast<int*>(red_val + 16); // [16]
  __shared__ int s_piv;                                // panel-relative pivot row of this column

  for (int idx = tid; idx < nr * ib; idx += PANEL_THREADS) {
    const int r = idx / ib, c = idx - r * ib;
    tile[r * PANEL_LD + c] = A[(int64_t)(r0 + r) * a.lda + c];
  }
  __syncthreads();

  unsigned epoch = 0;
  for (int c = 0; c < ib; ++c) {
    // ---- local arg-max of |a[r][c]| over rows >= c (lowest row wins ties) ----
    double best = -1.0;
    int bidx = 0x7fffffff;
    for (int r = tid; r < nr; r += PANEL_THREADS) {
      if (r0 + r >= c) {
        const double v = fabs(tile[r * PANEL_LD + c]);
        if (v > best) { best = v; bidx = r; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
    }
    if (lane == 0) { red_val[warp] = best; red_idx[warp] = bidx; }
    __syncthreads();
    if (warp == 0) {
      best = lane < PANEL_THREADS / 32 ? red_val[lane] : -1.0;
      bidx = lane < PANEL_THREADS / 32 ? red_idx[lane] : 0x7fffffff;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
        if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
      }
      best = __shfl_sync(0xffffffffu, best, 0);
      bidx = __shfl_sync(0xffffffffu, bidx, 0);
      if (SYNC == SYNC_NONE) {
        if (lane == 0) s_piv = (best >= 0.0) ? r0 + bidx : c;
        if (best >= 0.0 && lane < ib) prow[lane] = tile[bidx * PANEL_LD + lane];
      } else {
        Cand* cd = &sc->cand[c & 1][g];
        if (lane == 0) { cd->val = best; cd->row = (best >= 0.0) ? r0 + bidx : -1; }
        if (best >= 0.0 && lane < ib) cd->content[lane] = tile[bidx * PANEL_LD + lane];
        if (c >= r0 && c < r1 && lane < ib) sc->diag[c & 1][lane] = tile[(c - r0) * PANEL_LD + lane];
      }
    }
    if (SYNC == SYNC_NONE) {
      __syncthreads();
    } else {
      group_barrier<SYNC>(sc, G, epoch);
      if (warp == 0) {
        // every CTA elects the same winner: largest value, lowest row on ties
        double wv = -1.0; int wg = 0, wr = 0x7fffffff;
        for (int k = lane; k < G; k += 32) {
          const double v = *(volatile const double*)&sc->cand[c & 1][k].val;
          const int r = *(volatile const int*)&sc->cand[c & 1][k].row;
          if (v >= 0.0 && (v > wv || (v == wv && r < wr))) { wv = v; wg = k; wr = r; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double ov = __shfl_xor_sync(0xffffffffu, wv, o);
          const int og = __shfl_xor_sync(0xffffffffu, wg, o);
          const int orr = __shfl_xor_sync(0xffffffffu, wr, o);
          if (ov > wv || (ov == wv && orr < wr)) { wv = ov; wg = og; wr = orr; }
        }
        if (lane == 0) s_piv = (wv >= 0.0) ? wr : c;
        if (lane < ib) prow[lane] = *(volatile const double*)&sc->cand[c & 1][wg].content[lane];
      }
      __syncthreads();
    }
    const int p = s_piv;
    // ---- interchange rows c and p inside the panel ----
    if (p != c && tid < ib) {
      if (SYNC == SYNC_NONE) {
        tile[(p - r0) * PANEL_LD + tid] = tile[(c - r0) * PANEL_LD + tid];
        tile[(c - r0) * PANEL_LD + tid] = prow[tid];
      } else {
        if (p >= r0 && p < r1) tile[(p - r0) * PANEL_LD + tid] = *(volatile const double*)&sc->diag[c & 1][tid];
        if (c >= r0 && c < r1) tile[(c - r0) * PANEL_LD + tid] = prow[tid];
      }
    }
    if (g == 0 && tid == 0) ipiv[c] = a.jj + p;
    __syncthreads();
    const double piv = prow[c];
    if (piv == 0.0) {
      if (g == 0 && tid == 0 && a.info[mat] == 0) a.info[mat] = a.jj + c + 1;
    } else {
      // ---- scale the column and apply the rank-1 update to the rest of the panel ----
      for (int r = tid; r < nr; r += PANEL_THREADS) {
        if (r0 + r > c) {
          double* row = tile + r * PANEL_LD;
          const double l = row[c] / piv;
          row[c] = l;
          for (int cc = c + 1; cc < ib; ++cc) row[cc] = fma(-l, prow[cc], row[cc]);
        }
      }
    }
    __syncthreads();
  }

  for (int idx = tid; idx < nr * ib; idx += PANEL_THREADS) {
    const int r = idx / ib, c = idx - r * ib;
    A[(int64_t)(r0 + r) * a.lda + c] = tile[r * PANEL_LD + c];
  }
}


// =====================================================================================
// Block-column kernel
// =====================================================================================
constexpr int BC_ROWS = 196;     // rows of the block column one CTA keeps in shared memory
constexpr int BC_LD = NB + 1;    // odd leading dimension: column walks are bank-conflict free
constexpr int BC_THREADS = 512;
constexpr int BC_UW = NB - IB;   // widest U12 block
constexpr int BC_MAX_G = 148;

struct BcArgs {
  double* A; int64_t lda, sA;
  int n, j, jb, G, rpc;     // rpc: rows per CTA (>= jb when G > 1, so CTA 0 owns every pivot row)
  int* ipiv;                // [batch][n]
  int* info;                // [batch]
  char* scratch;            // per matrix: BcCand2[2][Gcap], BcChunk diag[2][NB], u12[IB][BC_UW], u12 flag
  size_t scratch_stride;
  int Gcap;
};

// Tagged exchange (blockcol2_kernel): every 16-byte unit carries its own epoch, so no fence is needed anywhere.
struct alignas(16) BcChunk { double v; unsigned long long tag; };
struct alignas(16) BcCand2 {
  BcChunk content[NB];
  double val;          // |a|, negative when the CTA has no eligible row
  int row;             // block-column-relative row index
  unsigned flag;       // epoch of the column this candidate belongs to
};
__host__ __device__ inline size_t bc_scratch_bytes(int Gcap) {  // large enough for either layout
  return (size_t)2 * Gcap * sizeof(BcCand2) + (size_t)2 * NB * sizeof(BcChunk) + (size_t)IB * BC_UW * sizeof(double) + 256;
}
inline size_t bc_smem_bytes(int rpc) {
  return sizeof(double) * ((((size_t)rpc * BC_LD + 1) & ~(size_t)1) + (size_t)IB * BC_UW + NB + 16) + sizeof(int) * 16;
}

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
// The candidate header travels as one aligned 16-byte access (one L2 sector transaction), so a reader that sees
// the new epoch in .y's upper half also sees the value and row stored with it.
template <typename C>
__device__ __forceinline__ void st_header(C* c, double val, int row, unsigned flag) {
  const unsigned long long lo = (unsigned long long)__double_as_longlong(val);
  const unsigned long long hi = (unsigned long long)(unsigned)row | ((unsigned long long)flag << 32);
  asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};\n" ::"l"(&c->val), "l"(lo), "l"(hi) : "memory");
}
__device__ __forceinline__ void st_chunk(BcChunk* p, double v, unsigned tag) {
  asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};\n" ::"l"(p), "l"((unsigned long long)__double_as_longlong(v)),
               "l"((unsigned long long)tag)
               : "memory");
}
__device__ __forceinline__ double ld_chunk_spin(const BcChunk* p, unsigned tag) {  // until the unit carries this epoch
  unsigned long long lo, hi;
  do {
    asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];\n" : "=l"(lo), "=l"(hi) : "l"(p) : "memory");
  } while ((unsigned)hi != tag);
  return __longlong_as_double((long long)lo);
}
template <typename C>
__device__ __forceinline__ void ld_header(const C* c, double& val, int& row, unsigned& flag) {
  unsigned long long lo, hi;
  asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];\n" : "=l"(lo), "=l"(hi) : "l"(&c->val) : "memory");
  val = __longlong_as_double((long long)lo);
  row = (int)(unsigned)(hi & 0xffffffffull);
  flag = (unsigned)(hi >> 32);
}

// One CTA per matrix (block columns of at most BC_ROWS rows): everything stays in shared memory.
__global__ void __launch_bounds__(BC_THREADS, 1) blockcol_kernel(BcArgs a) {
  extern __shared__ __align__(16) double sm[];
  constexpr int LD = BC_LD;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int mat = blockIdx.y;
  const int jb = a.jb;
  const int rows = a.n - a.j, nr = rows;
  double* A = a.A + (int64_t)mat * a.sA + (int64_t)a.j * a.lda + a.j;
  int* ipiv = a.ipiv + (int64_t)mat * a.n + a.j;

  double* tile = sm;                                   // [rows][LD]
  double* U = tile + (((size_t)a.rpc * LD + 1) & ~(size_t)1);  // [IB][BC_UW], 16-byte aligned for the double2 loads
  double* prow = U + IB * BC_UW;                       // [NB]
  double* red_val = prow + NB;                         // [16]
  int* red_idx = reinterpret_cast<int*>(red_val + 16); // [16]

  for (int idx = tid; idx < nr * jb; idx += BC_THREADS) {
    const int r = idx / jb, c = idx - r * jb;
    tile[r * LD + c] = A[(int64_t)r * a.lda + c];
  }
  __syncthreads();

  for (int c = 0; c < jb; ++c) {
    const int c0 = (c / IB) * IB, pe = min(c0 + IB, jb);
    // ---- arg-max of |a[r][c]| over rows >= c (lowest row wins ties) ----
    double best = -1.0;
    int bidx = 0x7fffffff;
    for (int r = tid; r < nr; r += BC_THREADS) {
      if (r >= c) {
        const double v = fabs(tile[r * LD + c]);
        if (v > best) { best = v; bidx = r; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
    }
    if (lane == 0) { red_val[warp] = best; red_idx[warp] = bidx; }
    __syncthreads();
    best = -1.0; bidx = 0x7fffffff;
#pragma unroll
    for (int w = 0; w < BC_THREADS / 32; ++w) {  // every thread reduces the 16 partials itself
      const double ov = red_val[w];
      const int oi = red_idx[w];
      if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
    }
    const int p = (best >= 0.0) ? bidx : c;  // block-column-relative pivot row
    if (tid < jb) {  // each thread swaps its own column: no cross-thread hazard
      const double xp = tile[p * LD + tid], xc = tile[c * LD + tid];
      tile[p * LD + tid] = xc;
      tile[c * LD + tid] = xp;
      prow[tid] = xp;
    }
    __syncthreads();
    if (tid == 0) ipiv[c] = a.j + p;
    const double piv = prow[c];
    if (piv == 0.0) {
      if (tid == 0 && a.info[mat] == 0) a.info[mat] = a.j + c + 1;
    } else {
      // ---- scale the column, rank-1 update of the rest of the inner panel ----
      for (int r = tid; r < nr; r += BC_THREADS) {
        if (r > c) {
          double* row = tile + r * LD;
          const double l = row[c] / piv;
          row[c] = l;
          double x[IB];  // loads first, stores last: the compiler cannot prove prow and row distinct
#pragma unroll
          for (int t = 0; t < IB; ++t) x[t] = (c0 + t > c && c0 + t < pe) ? fma(-l, prow[c0 + t], row[c0 + t]) : 0.0;
#pragma unroll
          for (int t = 0; t < IB; ++t)
            if (c0 + t > c && c0 + t < pe) row[c0 + t] = x[t];
        }
      }
    }
    __syncthreads();

    if (c == pe - 1 && pe < jb) {
      // ---- inner panel finished: U12 = L11^-1 A12, then rows >= pe get A22 -= L21 U12 ----
      const int W = jb - pe;  // pe < jb means this panel is IB wide
      if (tid < W) {  // one column of U12 per thread, forward substitution in registers
        double x[IB];
#pragma unroll
        for (int r = 0; r < IB; ++r) x[r] = tile[(c0 + r) * LD + pe + tid];
#pragma unroll
        for (int r = 1; r < IB; ++r) {
          double s = x[r];
#pragma unroll
          for (int t = 0; t < r; ++t) s = fma(-tile[(c0 + r) * LD + c0 + t], x[t], s);
          x[r] = s;
        }
#pragma unroll
        for (int r = 0; r < IB; ++r) {
          tile[(c0 + r) * LD + pe + tid] = x[r];
          U[r * BC_UW + tid] = x[r];
        }
      }
      __syncthreads();
      // 4x4 register tiles: a warp covers 16 rows x 32 columns per pass
      const int ly = lane >> 3, lx = lane & 7;
      const int nstrips = (nr + 15) >> 4, ncp = (W + 31) >> 5;
      for (int s = warp; s < nstrips; s += BC_THREADS / 32) {
        const int rb = s * 16 + ly * 4;
        if (s * 16 + 15 < pe) continue;  // whole strip above the trailing block
        const double* ap[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) ap[i] = tile + min(rb + i, a.rpc - 1) * LD + c0;
        for (int cp = 0; cp < ncp; ++cp) {
          const int cb = cp * 32 + lx * 4;
          double acc[4][4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) acc[i][jj] = 0.0;
          if (cb < BC_UW) {
#pragma unroll 8
            for (int k = 0; k < IB; ++k) {
              const double2 b01 = *reinterpret_cast<const double2*>(U + k * BC_UW + cb);
              const double2 b23 = *reinterpret_cast<const double2*>(U + k * BC_UW + cb + 2);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const double av = ap[i][k];
                acc[i][0] = fma(av, b01.x, acc[i][0]);
                acc[i][1] = fma(av, b01.y, acc[i][1]);
                acc[i][2] = fma(av, b23.x, acc[i][2]);
                acc[i][3] = fma(av, b23.y, acc[i][3]);
              }
            }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = rb + i;
            if (r < nr && r >= pe) {
#pragma unroll
              for (int jj = 0; jj < 4; ++jj)
                if (cb + jj < W) tile[r * LD + pe + cb + jj] -= acc[i][jj];
            }
          }
        }
      }
      __syncthreads();
    }
  }

  for (int idx = tid; idx < nr * jb; idx += BC_THREADS) {
    const int r = idx / jb, c = idx - r * jb;
    A[(int64_t)r * a.lda + c] = tile[r * LD + c];
  }
}

// Block column shared by G co-scheduled CTAs, second generation.  ncu on the kernel above (n = 19200, G = 96) showed
// 71 % of its 6.9 us per column in the pivot election: a fence before the header store, ONE warp polling all G
// headers four per lane (each re-poll a full L2 round trip), a second fence, and 15 of 16 warps parked at the barrier.
// Here every 16-byte unit that crosses CTAs carries its own epoch (BcChunk / header), so nothing is fenced and no
// barrier sits between the content stores and the header store; thread k < G polls header k and nothing else, the
// polling warps reduce by shuffles and every thread finishes the <= 5 partials itself; the next column's local
// arg-max is taken from the registers of the rank-1 update (no shared-memory scan, no barrier after the update).
// Three CTA barriers per column instead of six.
constexpr int BC_NPW = (BC_MAX_G + 31) / 32;  // polling warps
#ifdef HPS_BC_TIMING  // tools/panel_lab.cu: SM-clock stamps of thread 0 at the phase boundaries of every column
__device__ long long* g_bc_timing = nullptr;  // [G][NB][8]
#define BC_STAMP(k) do { if (g_bc_timing && tid == 0) g_bc_timing[((size_t)g * NB + c) * 8 + (k)] = clock64(); } while (0)
#else
#define BC_STAMP(k) do { } while (0)
#endif
__global__ void __launch_bounds__(BC_THREADS, 1) blockcol2_kernel(BcArgs a) {
  extern __shared__ __align__(16) double sm[];
  constexpr int LD = BC_LD;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = blockIdx.x, mat = blockIdx.y;
  const int G = a.G, jb = a.jb;
  const int rows = a.n - a.j;
  const int r0 = min(rows, g * a.rpc), r1 = min(rows, r0 + a.rpc), nr = r1 - r0;
  double* A = a.A + (int64_t)mat * a.sA + (int64_t)a.j * a.lda + a.j;
  int* ipiv = a.ipiv + (int64_t)mat * a.n + a.j;

  double* tile = sm;                                   // [rpc][LD]
  double* U = tile + (((size_t)a.rpc * LD + 1) & ~(size_t)1);  // [IB][BC_UW]
  double* prow = U + IB * BC_UW;                       // [NB]
  double* red_val = prow + NB;                         // [16]
  int* red_idx = reinterpret_cast<int*>(red_val + 16); // [16]
  __shared__ double s_pv[BC_NPW];
  __shared__ int s_pr[BC_NPW], s_pg[BC_NPW];

  char* sc = a.scratch + (size_t)mat * a.scratch_stride;
  BcCand2* cands = reinterpret_cast<BcCand2*>(sc);                                         // [2][Gcap]
  BcChunk* diag = reinterpret_cast<BcChunk*>(sc + (size_t)2 * a.Gcap * sizeof(BcCand2));   // [2][NB]
  double* u12g = reinterpret_cast<double*>(diag + 2 * NB);                                 // [IB][BC_UW]
  unsigned* u12_flag = reinterpret_cast<unsigned*>(u12g + IB * BC_UW);

  for (int idx = tid; idx < nr * jb; idx += BC_THREADS) {
    const int r = idx / jb, c = idx - r * jb;
    tile[r * LD + c] = A[(int64_t)(r0 + r) * a.lda + c];
  }
  __syncthreads();

  double cv = -1.0;       // this thread's candidate for the next column, taken from the rows it has just updated
  int cr = 0x7fffffff;
  bool scan = true;       // no carried candidate: scan the column in shared memory
  for (int c = 0; c < jb; ++c) {
    const int c0 = (c / IB) * IB, pe = min(c0 + IB, jb);
    BC_STAMP(0);
    // ---- local arg-max of |a[r][c]| over rows >= c (lowest row wins ties) ----
    double best = -1.0;
    int bidx = 0x7fffffff;
    if (scan) {
      for (int r = tid; r < nr; r += BC_THREADS) {
        if (r0 + r >= c) {
          const double v = fabs(tile[r * LD + c]);
          if (v > best) { best = v; bidx = r; }
        }
      }
    } else {
      best = cv; bidx = cr;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
    }
    if (lane == 0) { red_val[warp] = best; red_idx[warp] = bidx; }
    __syncthreads();  // (1) also orders the previous column's update before the content stores below
    best = -1.0; bidx = 0x7fffffff;
#pragma unroll
    for (int w = 0; w < BC_THREADS / 32; ++w) {
      const double ov = red_val[w];
      const int oi = red_idx[w];
      if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
    }

    // ---- publish: candidate row, the row that will be displaced (row c), header; no fence, no barrier ----
    BC_STAMP(1);
    const unsigned epoch = (unsigned)(a.j + c + 1);
    BcCand2* mine = cands + (size_t)(c & 1) * a.Gcap + g;
    BcChunk* dg = diag + (c & 1) * NB;
    if (tid < jb) {
      if (best >= 0.0) st_chunk(&mine->content[tid], tile[bidx * LD + tid], epoch);
    } else if (tid >= NB && tid < NB + jb) {
      if (c >= r0 && c < r1) st_chunk(&dg[tid - NB], tile[(c - r0) * LD + tid - NB], epoch);
    } else if (tid == 2 * NB) {
      st_header(mine, best, (best >= 0.0) ? r0 + bidx : -1, epoch);
    }
    // ---- election: thread k polls CTA k's header; largest value, lowest row on ties ----
    double wv = -1.0;
    int wr = 0x7fffffff, wg = 0;
    if (warp < BC_NPW) {
      if (tid < G) {
        double hv; int hr; unsigned hf;
        const BcCand2* his = cands + (size_t)(c & 1) * a.Gcap + tid;
        do { ld_header(his, hv, hr, hf); } while (hf != epoch);
        if (hv >= 0.0) { wv = hv; wr = hr; wg = tid; }
        BC_STAMP(2);  // thread 0: CTA 0's header seen
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, wv, o);
        const int og = __shfl_xor_sync(0xffffffffu, wg, o);
        const int orr = __shfl_xor_sync(0xffffffffu, wr, o);
        if (ov > wv || (ov == wv && orr < wr)) { wv = ov; wg = og; wr = orr; }
      }
      if (lane == 0) { s_pv[warp] = wv; s_pr[warp] = wr; s_pg[warp] = wg; }
    }
    __syncthreads();  // (2)
    BC_STAMP(3);  // every header seen by this CTA
    wv = -1.0; wr = 0x7fffffff; wg = 0;
#pragma unroll
    for (int w = 0; w < BC_NPW; ++w) {
      const double ov = s_pv[w];
      const int orr = s_pr[w], og = s_pg[w];
      if (ov > wv || (ov == wv && orr < wr)) { wv = ov; wg = og; wr = orr; }
    }
    const bool has = wv >= 0.0;      // false only for a column of NaNs (or an empty one): reported as singular
    const int p = has ? wr : c;      // block-column-relative pivot row
    if (has) {
      if (tid < jb) {
        const double x = ld_chunk_spin(&cands[(size_t)(c & 1) * a.Gcap + wg].content[tid], epoch);
        prow[tid] = x;
        if (c >= r0 && c < r1) tile[(c - r0) * LD + tid] = x;  // the pivot row moves up to row c ...
      } else if (tid >= NB && tid < NB + jb) {
        if (p != c && p >= r0 && p < r1) tile[(p - r0) * LD + tid - NB] = ld_chunk_spin(&dg[tid - NB], epoch);  // ... row c takes its place
      }
    }
    __syncthreads();  // (3)
    BC_STAMP(4);  // pivot row fetched
    if (g == 0 && tid == 0) ipiv[c] = a.j + p;
    const double piv = has ? prow[c] : 0.0;
    cv = -1.0; cr = 0x7fffffff;
    scan = true;
    if (piv == 0.0) {
      if (g == 0 && tid == 0 && a.info[mat] == 0) a.info[mat] = a.j + c + 1;
    } else {
      // ---- scale the column, rank-1 update of the rest of the inner panel; keep |a[r][c+1]| for the next election ----
      for (int r = tid; r < nr; r += BC_THREADS) {
        if (r0 + r > c) {
          double* row = tile + r * LD;
          const double l = row[c] / piv;
          row[c] = l;
          double nv = -1.0;
#pragma unroll
          for (int h = 0; h < IB; h += 16) {  // loads first, stores last (prow / row may alias for the compiler); 16 at a time
            double x[16];
#pragma unroll
            for (int t = 0; t < 16; ++t)
              x[t] = (c0 + h + t > c && c0 + h + t < pe) ? fma(-l, prow[c0 + h + t], row[c0 + h + t]) : 0.0;
#pragma unroll
            for (int t = 0; t < 16; ++t) {
              if (c0 + h + t > c && c0 + h + t < pe) row[c0 + h + t] = x[t];
              if (c0 + h + t == c + 1) nv = fabs(x[t]);
            }
          }
          if (nv > cv) { cv = nv; cr = r; }
        }
      }
      scan = (c + 1 >= pe);  // the first column of the next inner panel comes out of the trailing update below
    }
    BC_STAMP(5);  // thread 0's own row updated

    if (c == pe - 1 && pe < jb) {
      __syncthreads();
      // ---- inner panel finished: U12 = L11^-1 A12, then rows >= pe get A22 -= L21 U12 ----
      const int W = jb - pe;
      if (g == 0) {
        if (tid < W) {
          double x[IB];
#pragma unroll
          for (int r = 0; r < IB; ++r) x[r] = tile[(c0 + r) * LD + pe + tid];
#pragma unroll
          for (int r = 1; r < IB; ++r) {
            double s = x[r];
#pragma unroll
            for (int t = 0; t < r; ++t) s = fma(-tile[(c0 + r) * LD + c0 + t], x[t], s);
            x[r] = s;
          }
#pragma unroll
          for (int r = 0; r < IB; ++r) {
            tile[(c0 + r) * LD + pe + tid] = x[r];
            U[r * BC_UW + tid] = x[r];
            u12g[r * BC_UW + tid] = x[r];
          }
        }
        __syncthreads();
        if (tid == 0) st_release_u32(u12_flag, (unsigned)(a.j + pe));
      } else {
        if (tid == 0) {
          while (ld_acquire_u32(u12_flag) != (unsigned)(a.j + pe)) { }
        }
        __syncthreads();
        for (int idx = tid; idx < IB * W; idx += BC_THREADS) {
          const int r = idx / W, x = idx - r * W;
          U[r * BC_UW + x] = __ldcg(&u12g[r * BC_UW + x]);
        }
        __syncthreads();
      }
      const int ly = lane >> 3, lx = lane & 7;
      const int nstrips = (nr + 15) >> 4, ncp = (W + 31) >> 5;
      for (int s = warp; s < nstrips; s += BC_THREADS / 32) {
        const int rb = s * 16 + ly * 4;
        if (r0 + s * 16 + 15 < pe) continue;
        const double* ap[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) ap[i] = tile + min(rb + i, a.rpc - 1) * LD + c0;
        for (int cp = 0; cp < ncp; ++cp) {
          const int cb = cp * 32 + lx * 4;
          double acc[4][4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) acc[i][jj] = 0.0;
          if (cb < BC_UW) {
#pragma unroll 8
            for (int k = 0; k < IB; ++k) {
              const double2 b01 = *reinterpret_cast<const double2*>(U + k * BC_UW + cb);
              const double2 b23 = *reinterpret_cast<const double2*>(U + k * BC_UW + cb + 2);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const double av = ap[i][k];
                acc[i][0] = fma(av, b01.x, acc[i][0]);
                acc[i][1] = fma(av, b01.y, acc[i][1]);
                acc[i][2] = fma(av, b23.x, acc[i][2]);
                acc[i][3] = fma(av, b23.y, acc[i][3]);
              }
            }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = rb + i;
            if (r < nr && r0 + r >= pe) {
#pragma unroll
              for (int jj = 0; jj < 4; ++jj)
                if (cb + jj < W) tile[r * LD + pe + cb + jj] -= acc[i][jj];
            }
          }
        }
      }
      __syncthreads();
    }
  }
  __syncthreads();
  for (int idx = tid; idx < nr * jb; idx += BC_THREADS) {
    const int r = idx / jb, c = idx - r * jb;
    A[(int64_t)(r0 + r) * a.lda + c] = tile[r * LD + c];
  }
}

constexpr size_t PANEL_SMEM = sizeof(double) * ((size_t)PANEL_ROWS * PANEL_LD + IB + 16) + sizeof(int) * 16;

// rows k0..k1-1 of every matrix are exchanged with rows ipiv[k] on columns [c0, c0+ncols)
__global__ void laswp_kernel(double* A, int64_t lda, int64_t sA, int c0, int ncols, const int* ipiv, int n_ipiv,
                             int k0, int k1) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncols) return;
  double* a = A + (int64_t)blockIdx.y * sA + c0 + col;
  const int* piv = ipiv + (int64_t)blockIdx.y * n_ipiv;
  for (int k = k0; k < k1; ++k) {
    const int p = piv[k];
    if (p != k) {
      const double t = a[(int64_t)k * lda];
      a[(int64_t)k * lda] = a[(int64_t)p * lda];
      a[(int64_t)p * lda] = t;
    }
  }
}

// interchanges k0..k1-1 on columns [c0, c0 + ncols) of `batch` matrices sA apart (no profiling, no checks)
inline void launch_laswp(cudaStream_t st, int batch, double* A, int64_t lda, int64_t sA, int c0, int ncols, const int* ipiv,
                         int n_ipiv, int k0, int k1) {
  // (A "composed" variant — the swaps of a block column merged into one move list and applied through shared memory —
  // measured 87 us per call against 74 us for this kernel inside a factorisation, profiles/r02_laswp_block_vs_serial.txt;
  // it lives in tools/lu_lab_kernels.cuh.  With unpivoted merge matrices the interchanges are no-ops anyway.)
  laswp_kernel<<<dim3((ncols + 255) / 256, batch), 256, 0, st>>>(A, lda, sA, c0, ncols, ipiv, n_ipiv, k0, k1);
}

// X := L^-1 X for the unit-lower ib x ib block at A[jj,jj], X = A[jj:jj+ib, c0:c0+ncols)
__global__ void inner_trsm_kernel(double* A, int64_t lda, int64_t sA, int jj, int ib, int c0, int ncols) {
  __shared__ double L[IB][IB + 1];
  double* a = A + (int64_t)blockIdx.y * sA;
  for (int idx = threadIdx.x; idx < ib * ib; idx += blockDim.x) {
    const int r = idx / ib, c = idx - r * ib;
    L[r][c] = a[(int64_t)(jj + r) * lda + jj + c];
  }
  __syncthreads();
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncols) return;
  double x[IB];
  double* xp = a + (int64_t)jj * lda + c0 + col;
#pragma unroll
  for (int r = 0; r < IB; ++r) x[r] = (r < ib) ? xp[(int64_t)r * lda] : 0.0;
#pragma unroll
  for (int r = 1; r < IB; ++r) {
    double s = x[r];
#pragma unroll
    for (int t = 0; t < r; ++t) s = fma(-L[r][t], x[t], s);
    x[r] = s;
  }
#pragma unroll
  for (int r = 1; r < IB; ++r)
    if (r < ib) xp[(int64_t)r * lda] = x[r];
}

// Dense inverse of the nb x nb UNIT-LOWER triangular block at A[j,j] into W (ld = NB, zero elsewhere).
// Recursive blocked inversion inside one CTA: the four 32x32 diagonal blocks are inverted
// column-per-lane by four warps with no barrier at all, then the off-diagonal blocks follow from
//   inv([[A,0],[C,B]]) = [[A^-1, 0], [-B^-1 (C A^-1), B^-1]]
// as small shared-memory products at the 32 and 64 level: 3 barriers-separated phases instead of
// the 128 sequential column sweeps of an unblocked trti2.  (Partial pivoting keeps these blocks well
// conditioned; the upper blocks need the column-wise kernel below.)
constexpr int TRI_THREADS = 256;
constexpr int TRI_LD = NB + 1;

// dst (M x N) = alpha * A (M x K) * B (K x N), all in shared memory, executed by the whole CTA
__device__ __forceinline__ void smem_mm(double* dst, int ldd, const double* A, int lda, const double* B, int ldb, int M,
                                        int N, int K, double alpha) {
  for (int idx = threadIdx.x; idx < M * N; idx += TRI_THREADS) {
    const int r = idx / N, c = idx - r * N;
    double s = 0.0;
    for (int k = 0; k < K; ++k) s = fma(A[r * lda + k], B[k * ldb + c], s);
    dst[r * ldd + c] = alpha * s;
  }
}

__device__ __forceinline__ void trtri_lower_body(const double* A, int64_t lda, int64_t sA, int j0, int n,
                                                                  double* W, int64_t sW, int bx, int by) {
  extern __shared__ __align__(16) double sm[];
  constexpr int LD = TRI_LD;
  double* Ts = sm;                 // [NB][LD]   the triangle, inverted in place
  double* Ws = sm + NB * LD;       // [64][65]   product scratch
  const int j = j0 + bx * NB;      // diagonal block handled by this CTA
  const int nb = min(NB, n - j);
  const double* a = A + (int64_t)by * sA + (int64_t)j * lda + j;
  // ragged blocks are padded with the identity
  for (int idx = threadIdx.x; idx < NB * NB; idx += TRI_THREADS) {
    const int rr = idx / NB, cc = idx - rr * NB;
    double v = (rr == cc) ? 1.0 : 0.0;
    if (rr < nb && cc < nb && cc < rr) v = a[(int64_t)rr * lda + cc];
    Ts[rr * LD + cc] = v;
  }
  __syncthreads();
  // ---- phase 1: the four 32x32 diagonal blocks, one column per lane ----
  if (threadIdx.x < 128) {
    // column c of the inverse of a unit-lower 32 x 32 block, in this lane's registers (static indices, two accumulators)
    const int blk = threadIdx.x >> 5, c = threadIdx.x & 31, base = blk * 32;
    const double* T = Ts + base * LD + base;
    double x[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int t = 0; t < 32; ++t) {
        if (t < r) {
          if (t & 1) s1 = fma(T[r * LD + t], x[t], s1);
          else s0 = fma(T[r * LD + t], x[t], s0);
        }
      }
      x[r] = (r == c) ? 1.0 : ((r > c) ? -(s0 + s1) : 0.0);
    }
    __syncwarp();  // every lane has read the block before any lane overwrites it
#pragma unroll
    for (int r = 0; r < 32; ++r) Ts[(base + r) * LD + base + c] = x[r];
  }
  __syncthreads();
  // ---- phase 2: 32-level off-diagonal blocks of both 64x64 halves ----
  for (int half = 0; half < 2; ++half) {  // W = L21 * inv(L11)
    const int b0 = half * 64, b1 = b0 + 32;
    smem_mm(Ws + half * 32 * 65, 65, Ts + b1 * LD + b0, LD, Ts + b0 * LD + b0, LD, 32, 32, 32, 1.0);
  }
  __syncthreads();
  for (int half = 0; half < 2; ++half) {  // X21 = -inv(L22) * W
    const int b0 = half * 64, b1 = b0 + 32;
    smem_mm(Ts + b1 * LD + b0, LD, Ts + b1 * LD + b1, LD, Ws + half * 32 * 65, 65, 32, 32, 32, -1.0);
  }
  __syncthreads();
  // ---- phase 3: the 64-level off-diagonal block ----
  smem_mm(Ws, 65, Ts + 64 * LD, LD, Ts, LD, 64, 64, 64, 1.0);                       // W = L21 * inv(L11)
  __syncthreads();
  smem_mm(Ts + 64 * LD, LD, Ts + 64 * LD + 64, LD, Ws, 65, 64, 64, 64, -1.0);       // X21 = -inv(L22) * W
  __syncthreads();
  double* w = W + (int64_t)by * sW + (int64_t)(j / NB) * NB * NB;
  for (int idx = threadIdx.x; idx < nb * nb; idx += TRI_THREADS) {
    const int rr = idx / nb, cc = idx - rr * nb;
    w[rr * NB + cc] = Ts[rr * LD + cc];
  }
}
__global__ void __launch_bounds__(TRI_THREADS) trtri_lower_kernel(const double* A, int64_t lda, int64_t sA, int j0, int n,
                                                                  double* W, int64_t sW) {
  trtri_lower_body(A, lda, sA, j0, n, W, sW, blockIdx.x, blockIdx.y);
}

// Inverse of the nb x nb UPPER-triangular diagonal blocks (with their diagonal) by BACK SUBSTITUTION, U X = I:
// every column of the inverse carries the backward error of a triangular solve.  The blocked formula
//   inv([[A,B],[0,C]]) = [[A^-1, -A^-1 (B C^-1)], [0, C^-1]]
// used for the unit-lower blocks is NOT accurate enough here: the pivot rows of the ItI leaf systems mix impedance
// rows with operator rows three orders of magnitude larger, |A^-1||B||C^-1| is then far larger than |X12| and the
// solve loses two digits against LAPACK (tools/lu_accuracy.py, reproduced in NumPy).  Organisation: 32x32 blocks,
// four phases by block distance d = J - I.  d = 0: the diagonal blocks, one column per lane.  d >= 1:
//   W = sum_{K=I+1..J} U[I][K] X[K][J]   (all threads; plain products with finished blocks of X, no inverses)
//   U[I][I] X[I][J] = -W                 (one column per lane, 32 sequential rows)
// i.e. exactly the terms of the column-wise substitution, re-associated.  X[I][J], I < J, is kept in the unused
// block (J, I) of the tile's lower triangle, the diagonal blocks of X in Xd.
// Back substitution U x = b for ONE column held in this lane's registers (static indices: both loops are unrolled).
// On entry x[r] = b[r]; rows above `top` are the only ones that can be non-zero.  U: 32 x 32 upper block at T (ld),
// rdiag[r] = 1 / U[r][r].  Two accumulators halve the dependent FMA chain of each row's dot product.
__device__ __forceinline__ void backsub32(const double* __restrict__ T, int ld, const double* __restrict__ rdiag,
                                          double (&x)[32], int top) {
#pragma unroll
  for (int rr = 0; rr < 32; ++rr) {
    const int r = 31 - rr;
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int t = 0; t < 32; ++t) {
      if (t > r) {
        if (t & 1) s1 = fma(T[r * ld + t], x[t], s1);
        else s0 = fma(T[r * ld + t], x[t], s0);
      }
    }
    x[r] = (r <= top) ? (x[r] - (s0 + s1)) * rdiag[r] : 0.0;
  }
}

__device__ __forceinline__ void trtri_upper_body(const double* A, int64_t lda, int64_t sA, int j0, int n,
                                                  double* W, int64_t sW, int bx, int by) {
  extern __shared__ __align__(16) double sm[];
  constexpr int LD = TRI_LD;
  double* Ts = sm;                  // [NB][LD]     U in the upper triangle, X[I][J] (I < J) in block (J, I)
  double* Xd = Ts + NB * LD;        // [4][32][33]  diagonal blocks of X
  double* Ws = Xd + 4 * 32 * 33;    // [3][32][33]  right-hand sides of the current phase
  double* rdiag = Ws + 3 * 32 * 33; // [NB]         reciprocals of U's diagonal
  const int j = j0 + bx * NB;
  const int nb = min(NB, n - j);
  const double* a = A + (int64_t)by * sA + (int64_t)j * lda + j;
  for (int idx = threadIdx.x; idx < NB * NB; idx += TRI_THREADS) {  // ragged blocks are padded with the identity
    const int rr = idx / NB, cc = idx - rr * NB;
    double v = (rr == cc) ? 1.0 : 0.0;
    if (rr < nb && cc < nb) v = (cc >= rr) ? a[(int64_t)rr * lda + cc] : 0.0;
    Ts[rr * LD + cc] = v;
  }
  __syncthreads();
  if (threadIdx.x < NB) rdiag[threadIdx.x] = 1.0 / Ts[threadIdx.x * LD + threadIdx.x];
  __syncthreads();
  // ---- d = 0: the diagonal blocks, one column per lane, the column in registers ----
  if (threadIdx.x < 128) {
    const int blk = threadIdx.x >> 5, c = threadIdx.x & 31, base = blk * 32;
    double x[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) x[r] = (r == c) ? 1.0 : 0.0;
    backsub32(Ts + base * LD + base, LD, rdiag + base, x, c);
    double* xd = Xd + blk * 32 * 33 + c;
#pragma unroll
    for (int r = 0; r < 32; ++r) xd[r * 33] = x[r];
  }
  __syncthreads();
#pragma unroll 1  // one copy of the unrolled substitution below (three would push its register array into local memory)
  for (int d = 1; d < 4; ++d) {
    const int nblk = 4 - d;
    for (int idx = threadIdx.x; idx < nblk * 1024; idx += TRI_THREADS) {
      const int I = idx >> 10, r = (idx >> 5) & 31, c = idx & 31, J = I + d;
      const double* urow = Ts + (32 * I + r) * LD;
      double s = 0.0;
      for (int K = I + 1; K <= J; ++K) {
        const double* xb = (K == J) ? (Xd + J * 32 * 33 + c) : (Ts + (32 * J) * LD + 32 * K + c);
        const int xs = (K == J) ? 33 : LD;
#pragma unroll 8
        for (int k = 0; k < 32; ++k) s = fma(urow[32 * K + k], xb[k * xs], s);
      }
      Ws[I * 32 * 33 + r * 33 + c] = s;
    }
    __syncthreads();
    if (threadIdx.x < nblk * 32) {
      const int I = threadIdx.x >> 5, c = threadIdx.x & 31, J = I + d;
      const double* w = Ws + I * 32 * 33 + c;
      double x[32];
#pragma unroll
      for (int r = 0; r < 32; ++r) x[r] = -w[r * 33];
      backsub32(Ts + (32 * I) * LD + 32 * I, LD, rdiag + 32 * I, x, 31);
      double* xo = Ts + (32 * J) * LD + 32 * I + c;  // X[I][J][t][c] at xo[t * LD]
#pragma unroll
      for (int r = 0; r < 32; ++r) xo[r * LD] = x[r];
    }
    __syncthreads();
  }
  double* w = W + (int64_t)by * sW + (int64_t)(j / NB) * NB * NB;
  for (int idx = threadIdx.x; idx < nb * nb; idx += TRI_THREADS) {
    const int rr = idx / nb, cc = idx - rr * nb;
    const int I = rr >> 5, J = cc >> 5;
    double v = 0.0;
    if (I == J) v = Xd[I * 32 * 33 + (rr & 31) * 33 + (cc & 31)];   // zero below the diagonal by construction
    else if (I < J) v = Ts[(32 * J + (rr & 31)) * LD + 32 * I + (cc & 31)];
    w[rr * NB + cc] = v;
  }
}
__global__ void __launch_bounds__(TRI_THREADS) trtri_upper_kernel(const double* A, int64_t lda, int64_t sA, int j0, int n,
                                                                  double* W, int64_t sW) {
  trtri_upper_body(A, lda, sA, j0, n, W, sW, blockIdx.x, blockIdx.y);
}
constexpr size_t TRTRI_UPPER_SMEM = sizeof(double) * (NB * (NB + 1) + 7 * 32 * 33 + NB);

constexpr size_t TRTRI_SMEM = sizeof(double) * (NB * (NB + 1) + 64 * 65 + 4 * 32 * 33);
constexpr size_t TRTRI_PAIR_SMEM = TRTRI_SMEM > TRTRI_UPPER_SMEM ? TRTRI_SMEM : TRTRI_UPPER_SMEM;
// both inverses of ONE diagonal block (at A[j,j]) in one launch: CTA x = 0 the unit-lower one, x = 1 the upper one
__global__ void __launch_bounds__(TRI_THREADS) trtri_pair_kernel(const double* A, int64_t lda, int64_t sA, int j, int n,
                                                                 double* Wl, double* Wu, int64_t sW) {
  if (blockIdx.x == 0) trtri_lower_body(A, lda, sA, j, n, Wl, sW, 0, blockIdx.y);
  else trtri_upper_body(A, lda, sA, j, n, Wu, sW, 0, blockIdx.y);
}

// dst[b][r][c] = src[b][r][c] for an (rows x cols) block
__global__ void copy_block_kernel(double* dst, int64_t ldd, int64_t sD, const double* src, int64_t lds, int64_t sS,
                                  int rows, int cols) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  const int r = idx / cols, c = idx - r * cols;
  dst[(int64_t)blockIdx.y * sD + (int64_t)r * ldd + c] = src[(int64_t)blockIdx.y * sS + (int64_t)r * lds + c];
}


struct LuWorkspace {
  int* ipiv;
  double* Linv;  // [batch][nblk][NB][NB] inverses of the unit-lower diagonal blocks
  double* Uinv;  // [batch][nblk][NB][NB] inverses of the upper diagonal blocks
  double* tmp;   // [batch][NB][16] staging for narrow right-hand sides
  PanelScratch* scratch;
  char* bc_scratch;  // block-column kernel exchange area, bc_stride bytes per matrix (flags must start at 0)
  size_t bc_stride;
  int bc_Gcap;
  double* P;         // [batch][n][NB] the speculative path's L21 before it is committed
};

// CTAs per matrix the block-column kernel may need for an n x n factorisation (0: single-CTA only)
inline int bc_gcap(int n) {
  const int G0 = (n + BC_ROWS - 1) / BC_ROWS;
  return G0 <= 1 ? 0 : std::min(G0, BC_MAX_G);
}

bool carve(Arena& ar, int batch, int n, LuWorkspace& w) {
  const size_t nblk = (n + NB - 1) / NB;
  w.ipiv = ar.take<int>((size_t)batch * n);
  w.Linv = ar.take<double>((size_t)batch * nblk * NB * NB);
  w.Uinv = ar.take<double>((size_t)batch * nblk * NB * NB);
  w.tmp = ar.take<double>((size_t)batch * NB * 16);
  w.scratch = ar.take<PanelScratch>((size_t)batch);
  w.bc_Gcap = bc_gcap(n);
  w.bc_stride = w.bc_Gcap ? align_up(bc_scratch_bytes(w.bc_Gcap), 256) : 0;
  w.bc_scratch = w.bc_Gcap ? ar.take<char>((size_t)batch * w.bc_stride) : reinterpret_cast<char*>(w.scratch);
  w.P = ar.take<double>((size_t)batch * n * NB);
  return w.ipiv && w.Linv && w.Uinv && w.tmp && w.scratch && w.bc_scratch && w.P;
}

// co-resident CTAs of a cooperative kernel on the current device (queried once per device)
template <typename K>
int coop_capacity(K kernel, int threads, size_t smem, std::atomic<int>& cache, int& out) {
  int cap = cache.load(std::memory_order_acquire);
  if (cap < 0) {
    int dev = 0, sms = 0, per_sm = 0;
    HPS_CUDA(cudaGetDevice(&dev));
    HPS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    HPS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    cap = sms * per_sm;
    cache.store(cap, std::memory_order_release);
  }
  out = cap;
  return 0;
}

int launch_panel_impl(cudaStream_t st, int batch, PanelArgs pa);
int launch_panel(cudaStream_t st, int batch, PanelArgs pa) {
  prof_begin(PROF_PANEL, st, (double)batch * (pa.n - pa.jj) * pa.ib);
  const int rc = launch_panel_impl(st, batch, pa);
  prof_end(PROF_PANEL, st);
  return rc;
}
int launch_panel_impl(cudaStream_t st, int batch, PanelArgs pa) {
  const int rows = pa.n - pa.jj;
  int G = (rows + PANEL_ROWS - 1) / PANEL_ROWS;
  if (G <= 1) {
    pa.G = 1;
    panel_kernel<SYNC_NONE><<<dim3(1, batch), PANEL_THREADS, PANEL_SMEM, st>>>(pa);
    HPS_LAUNCH_CHECK("panel_kernel<none>");
    return 0;
  }
  if (G <= 8) {
    G = G <= 2 ? 2 : (G <= 4 ? 4 : 8);
    pa.G = G;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(G, batch);
    cfg.blockDim = dim3(PANEL_THREADS);
    cfg.dynamicSmemBytes = PANEL_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = G;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    HPS_CUDA(cudaLaunchKernelEx(&cfg, panel_kernel<SYNC_CLUSTER>, pa));
    ++g_launches;
    return 0;
  }
  if (G > MAX_G) return fail_arg(3, "matrix too tall for the panel kernel (n > 98304)");
  pa.G = G;
  DeviceState* ds = nullptr;
  HPS_TRY(device_state(ds));
  int capacity = 0;
  HPS_TRY(coop_capacity(panel_kernel<SYNC_GRID>, PANEL_THREADS, PANEL_SMEM, ds->coop_capacity, capacity));
  const int per_launch = capacity / G;
  if (per_launch < 1) return fail_arg(3, "panel does not fit a cooperative launch");
  for (int b0 = 0; b0 < batch; b0 += per_launch) {
    const int nb = min(per_launch, batch - b0);
    PanelArgs sub = pa;
    sub.A = pa.A + (int64_t)b0 * pa.sA;
    sub.ipiv = pa.ipiv + (int64_t)b0 * pa.n;
    sub.info = pa.info + b0;
    sub.scratch = pa.scratch + b0;
    for (int b = 0; b < nb; ++b)
      HPS_CUDA(cudaMemsetAsync(&sub.scratch[b].counter, 0, sizeof(unsigned), st));
    void* args[] = {&sub};
    HPS_CUDA(cudaLaunchCooperativeKernel((void*)panel_kernel<SYNC_GRID>, dim3(G, nb), dim3(PANEL_THREADS), args,
                                         PANEL_SMEM, st));
    ++g_launches;
  }
  return 0;
}

int laswp(cudaStream_t st, int batch, double* A, int64_t lda, int64_t sA, int c0, int ncols, const int* ipiv, int n,
          int k0, int k1) {
  if (ncols <= 0 || k1 <= k0) return 0;
  prof_begin(PROF_LASWP, st, (double)batch * ncols * (k1 - k0));
  launch_laswp(st, batch, A, lda, sA, c0, ncols, ipiv, n, k0, k1);
  prof_end(PROF_LASWP, st);
  HPS_LAUNCH_CHECK("laswp_kernel");
  return 0;
}

// LU with partial pivoting of the jb x jb diagonal block at A[j,j] of every matrix (jb <= NB), the pivot search
// restricted to the block: one CTA of 128 threads per matrix, thread r owns row r.  The row's entries in the current
// 32-column panel live in REGISTERS (the column loop is fully unrolled, so every index is static): a column step is
// one shuffle reduction, two barriers of four warps and up to 31 FMAs per thread — about 300 cycles instead of the
// ~2 us the shared-memory panel of blockcol_kernel<false> needs on a 128-row block.  The block itself sits in shared
// memory for the (rare) row interchanges, U12 = L11^-1 A12 and the trailing update, whose L21 operand is the register
// panel just computed.
constexpr int DB_THREADS = 128;
constexpr int DB_LD = NB + 1;
constexpr int DB_UW = NB - IB;
constexpr size_t DB_SMEM = sizeof(double) * ((size_t)NB * DB_LD + 1 + 2 * (IB + 1) + (size_t)IB * DB_UW);
// PIVOT = false: the diagonal is taken as the pivot without looking (no reduction, one barrier per column, the
// owner of the pivot row hands out the reciprocal); a multiplier above 1 in magnitude sets info = -2 like the check
// on the rows below the block (spec_commit_kernel), so the result is accepted only where partial pivoting would have
// made the same choice.
template <bool PIVOT>
__global__ void __launch_bounds__(DB_THREADS) diagblk_kernel(double* __restrict__ A, int64_t lda, int64_t sA, int n, int j,
                                                             int jb, int* __restrict__ ipiv_all, int* __restrict__ info) {
  extern __shared__ __align__(16) double sm[];
  constexpr int LD = DB_LD;
  double* T = sm;                                  // [NB][LD]
  double* U = T + (((size_t)NB * LD + 1) & ~(size_t)1);  // [IB][DB_UW], 16-byte aligned
  double* prow = U + IB * DB_UW;                   // [2][IB + 1]: pivot row of the panel (+ the reciprocal of the pivot)
  __shared__ double s_val[DB_THREADS / 32];
  __shared__ int s_row[DB_THREADS / 32];
  const int r = threadIdx.x, lane = r & 31, warp = r >> 5, mat = blockIdx.x;
  double* a = A + (int64_t)mat * sA + (int64_t)j * lda + j;
  int* ipiv = ipiv_all + (int64_t)mat * n + j;
  const bool live = r < jb;
  bool bad = false;  // PIVOT = false: a multiplier larger than 1 was seen
  for (int rr = 0; rr < jb; ++rr)
    if (live) T[rr * LD + r] = a[(int64_t)rr * lda + r];
  __syncthreads();
  for (int c0 = 0; c0 < jb; c0 += IB) {
    const int pw = min(IB, jb - c0);  // panel width
    double x[IB];
#pragma unroll
    for (int k = 0; k < IB; ++k) x[k] = (live && k < pw) ? T[r * LD + c0 + k] : 0.0;
#pragma unroll
    for (int c = 0; c < IB; ++c) {
      if (c < pw) {  // uniform
        const int col = c0 + c;
        if (!PIVOT) {
          double* pr = prow + (c & 1) * (IB + 1);
          if (r == col) {
#pragma unroll
            for (int k = 0; k < IB; ++k) pr[k] = x[k];
            pr[IB] = 1.0 / x[c];
          }
          if (r == 0) ipiv[col] = j + col;
          __syncthreads();
          const double piv = pr[c];
          if (piv == 0.0) {
            if (r == 0 && info[mat] == 0) info[mat] = j + col + 1;
          } else if (live && r > col) {
            const double l = x[c] * pr[IB];
            if (!(fabs(l) <= 1.0 + 1e-8)) bad = true;
            x[c] = l;
#pragma unroll
            for (int k = c + 1; k < IB; ++k) x[k] = fma(-l, pr[k], x[k]);
          }
          continue;
        }
        // ---- arg-max of |a[r][col]| over rows >= col (lowest row wins ties) ----
        double v = (live && r >= col) ? fabs(x[c]) : -1.0;
        int vr = r;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double ov = __shfl_xor_sync(0xffffffffu, v, o);
          const int orr = __shfl_xor_sync(0xffffffffu, vr, o);
          if (ov > v || (ov == v && orr < vr)) { v = ov; vr = orr; }
        }
        if (lane == 0) { s_val[warp] = v; s_row[warp] = vr; }
        __syncthreads();
        v = s_val[0]; vr = s_row[0];
#pragma unroll
        for (int w = 1; w < DB_THREADS / 32; ++w) {
          const double ov = s_val[w];
          const int orr = s_row[w];
          if (ov > v || (ov == v && orr < vr)) { v = ov; vr = orr; }
        }
        const int p = (v >= 0.0) ? vr : col;
        if (p != col) {  // uniform: rows col and p change places (registers of the two owners + the block in smem)
          if (r == col || r == p) {
#pragma unroll
            for (int k = 0; k < IB; ++k)
              if (k < pw) T[r * LD + c0 + k] = x[k];
          }
          __syncthreads();
          if (live) { const double t1 = T[col * LD + r], t2 = T[p * LD + r]; T[col * LD + r] = t2; T[p * LD + r] = t1; }
          __syncthreads();
          if (r == col || r == p) {
#pragma unroll
            for (int k = 0; k < IB; ++k)
              if (k < pw) x[k] = T[r * LD + c0 + k];
          }
        }
        if (r == 0) ipiv[col] = j + p;
        double* pr = prow + (c & 1) * (IB + 1);  // double-buffered: the previous column's readers may still be at work
        if (r == col) {
#pragma unroll
          for (int k = 0; k < IB; ++k) pr[k] = x[k];
        }
        __syncthreads();
        const double piv = pr[c];
        if (piv == 0.0) {
          if (r == 0 && info[mat] == 0) info[mat] = j + col + 1;
        } else if (live && r > col) {
          const double l = x[c] / piv;
          x[c] = l;
#pragma unroll
          for (int k = c + 1; k < IB; ++k) x[k] = fma(-l, pr[k], x[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < IB; ++k)
      if (live && k < pw) T[r * LD + c0 + k] = x[k];
    __syncthreads();
    const int pe = c0 + pw, W = jb - pe;
    if (W > 0) {  // pw == IB here
      // ---- U12 = L11^-1 A12: one column per thread, forward substitution in registers ----
      if (r < W) {
        double y[IB];
#pragma unroll
        for (int i = 0; i < IB; ++i) y[i] = T[(c0 + i) * LD + pe + r];
#pragma unroll
        for (int i = 1; i < IB; ++i) {
          double s = y[i];
#pragma unroll
          for (int t = 0; t < i; ++t) s = fma(-T[(c0 + i) * LD + c0 + t], y[t], s);
          y[i] = s;
        }
#pragma unroll
        for (int i = 0; i < IB; ++i) { T[(c0 + i) * LD + pe + r] = y[i]; U[i * DB_UW + r] = y[i]; }
      }
      __syncthreads();
      // ---- A22 -= L21 U12 on rows >= pe: this thread's row of L21 is the register panel ----
      if (live && r >= pe) {
        double* trow = T + r * LD + pe;
        int w = 0;
        for (; w + 1 < W; w += 2) {
          double s0 = 0.0, s1 = 0.0;
#pragma unroll
          for (int k = 0; k < IB; ++k) {
            const double2 u = *reinterpret_cast<const double2*>(U + k * DB_UW + w);
            s0 = fma(x[k], u.x, s0);
            s1 = fma(x[k], u.y, s1);
          }
          trow[w] -= s0;
          trow[w + 1] -= s1;
        }
        if (w < W) {
          double s0 = 0.0;
#pragma unroll
          for (int k = 0; k < IB; ++k) s0 = fma(x[k], U[k * DB_UW + w], s0);
          trow[w] -= s0;
        }
      }
      __syncthreads();
    }
  }
  for (int rr = 0; rr < jb; ++rr)
    if (live) a[(int64_t)rr * lda + r] = T[rr * LD + r];
  if (!PIVOT && __syncthreads_or(bad) && r == 0) atomicCAS(&info[mat], 0, -2);
}

// opt the large-shared-memory kernels in, once per device
int configure_lu_kernels() {
  DeviceState* ds = nullptr;
  HPS_TRY(device_state(ds));
  if (ds->lu_configured.load(std::memory_order_acquire)) return 0;  // idempotent: a race only repeats the calls
    HPS_CUDA(cudaFuncSetAttribute(blockcol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bc_smem_bytes(BC_ROWS)));
    HPS_CUDA(cudaFuncSetAttribute(blockcol2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bc_smem_bytes(BC_ROWS)));
    HPS_CUDA(cudaFuncSetAttribute(panel_kernel<SYNC_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PANEL_SMEM));
    HPS_CUDA(cudaFuncSetAttribute(panel_kernel<SYNC_CLUSTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PANEL_SMEM));
    HPS_CUDA(cudaFuncSetAttribute(panel_kernel<SYNC_GRID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PANEL_SMEM));
    HPS_CUDA(cudaFuncSetAttribute(trtri_lower_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRTRI_SMEM));
    HPS_CUDA(cudaFuncSetAttribute(trtri_upper_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRTRI_UPPER_SMEM));
    HPS_CUDA(cudaFuncSetAttribute(trtri_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRTRI_PAIR_SMEM));
    HPS_CUDA(cudaFuncSetAttribute(diagblk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DB_SMEM));
    HPS_CUDA(cudaFuncSetAttribute(diagblk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DB_SMEM));
  ds->lu_configured.store(true, std::memory_order_release);
  return 0;
}

// One launch per block column (several when a cooperative batch does not fit the GPU at once).
// Returns 1 through `done` when the kernel was used, 0 when the block column is too tall.
int launch_blockcol(cudaStream_t st, int batch, int n, double* A, int64_t lda, int64_t sA, int j, int jb, LuWorkspace& w,
                    int* info, bool& done) {
  done = false;
  const int rows = n - j;
  const int G = (rows + BC_ROWS - 1) / BC_ROWS;
  if (G > 1 && (G > w.bc_Gcap || G > BC_MAX_G)) return 0;
  if (G > 1) {
    // Several CTAs per matrix pay an exchange through L2 per column: worth it when the factorisation is
    // latency-bound (one wave of CTAs: big single matrices), not when many matrices queue for the SMs —
    // those go through the narrower legacy panels, which need a quarter of the CTAs per matrix.
    DeviceState* ds = nullptr;
    HPS_TRY(device_state(ds));
    int sms = ds->sm_count.load(std::memory_order_acquire);
    if (sms == 0) {
      int dev = 0;
      HPS_CUDA(cudaGetDevice(&dev));
      HPS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      ds->sm_count.store(sms, std::memory_order_release);
    }
    if ((long long)batch * G > sms) return 0;
  }
  BcArgs a;
  a.A = A; a.lda = lda; a.sA = sA; a.n = n; a.j = j; a.jb = jb; a.G = G;
  a.rpc = (G == 1) ? rows : std::max(NB, (rows + G - 1) / G);
  a.ipiv = w.ipiv; a.info = info; a.scratch = w.bc_scratch; a.scratch_stride = w.bc_stride; a.Gcap = w.bc_Gcap;
  const size_t smem = bc_smem_bytes(a.rpc);
  if (G == 1) {
    prof_begin(PROF_PANEL, st, (double)batch * rows * jb);
    blockcol_kernel<<<dim3(1, batch), BC_THREADS, smem, st>>>(a);
    prof_end(PROF_PANEL, st);
    HPS_LAUNCH_CHECK("blockcol_kernel<single>");
    done = true;
    return 0;
  }
  void (*shared_kernel)(BcArgs) = blockcol2_kernel;
  if (G <= 8) {  // one thread-block cluster per matrix: co-scheduled by the hardware
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(G, batch);
    cfg.blockDim = dim3(BC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = G;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    prof_begin(PROF_PANEL, st, (double)batch * rows * jb);
    HPS_CUDA(cudaLaunchKernelEx(&cfg, shared_kernel, a));
    prof_end(PROF_PANEL, st);
    ++g_launches;
    done = true;
    return 0;
  }
  // cooperative launch: every CTA of the grid is resident, so the flag polling cannot deadlock
  int dev = 0, sms = 0, per_sm = 0;
  HPS_CUDA(cudaGetDevice(&dev));
  HPS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  HPS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, shared_kernel, BC_THREADS, smem));
  const int per_launch = (sms * per_sm) / G;
  if (per_launch < 1) return 0;
  prof_begin(PROF_PANEL, st, (double)batch * rows * jb);
  for (int b0 = 0; b0 < batch; b0 += per_launch) {
    const int nb = std::min(per_launch, batch - b0);
    BcArgs sub = a;
    sub.A = A + (int64_t)b0 * sA;
    sub.ipiv = w.ipiv + (int64_t)b0 * n;
    sub.info = info + b0;
    sub.scratch = w.bc_scratch + (size_t)b0 * w.bc_stride;
    void* args[] = {&sub};
    HPS_CUDA(cudaLaunchCooperativeKernel((void*)shared_kernel, dim3(G, nb), dim3(BC_THREADS), args, smem, st));
    ++g_launches;
  }
  prof_end(PROF_PANEL, st);
  done = true;
  return 0;
}

// X := Tinv * X for a jb-row block X (in place).  Wide X goes through the DMMA GEMM (one tile
// row, so in-place is safe); narrow X is staged through tmp because the row-per-warp kernel
// would read rows other warps have already overwritten.
int tri_mult(cudaStream_t st, int batch, int jb, const double* Tinv, int64_t sW, double* X, int64_t ld,
             int64_t stride, int ncols, double* tmp) {
  if (ncols <= 0) return 0;
  if (ncols >= 16) return dgemm(st, jb, ncols, jb, 1.0, Tinv, NB, sW, X, ld, stride, 0.0, X, ld, stride, batch);
  const int64_t sT = (int64_t)NB * 16;
  HPS_TRY(dgemm_skinny(st, jb, ncols, jb, 1.0, Tinv, NB, sW, X, ld, stride, 0.0, nullptr, 0, 0, tmp, ncols, sT, batch));
  copy_block_kernel<<<dim3((jb * ncols + 255) / 256, batch), 256, 0, st>>>(X, ld, stride, tmp, ncols, sT, jb, ncols);
  HPS_LAUNCH_CHECK("copy_block_kernel");
  return 0;
}

struct Mat {  // batched row-major matrix view
  double* p;
  int64_t ld, stride;
  double* at(int64_t r, int64_t c) const { return p + r * ld + c; }
};

// A21 := P (rows x jb), and the check that makes the speculation below equivalent to partial pivoting: every
// multiplier must satisfy |l| <= 1, otherwise info := -2 (kept if info already reports a zero pivot).
__global__ void __launch_bounds__(256) spec_commit_kernel(double* __restrict__ A21, int64_t lda, int64_t sA,
                                                          const double* __restrict__ P, int64_t sP, int rows, int jb,
                                                          int* __restrict__ info, double bound) {
  const int mat = blockIdx.y;
  const double* p = P + (int64_t)mat * sP;
  double* a = A21 + (int64_t)mat * sA;
  const int64_t total = (int64_t)rows * jb;
  bool bad = false;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / jb, c = e - r * jb;
    const double v = p[e];
    bad |= !(fabs(v) <= bound);  // also catches NaN
    a[r * lda + c] = v;
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicCAS(&info[mat], 0, -2);
}

// Process-wide switch of the speculative block columns (hps_lu_set_speculative; HPS_LU_SPEC=0 starts with it off)
std::atomic<int> g_lu_speculate{[] { const char* e = std::getenv("HPS_LU_SPEC"); return (e && e[0] == '0') ? 0 : 1; }()};

// Block column j WITHOUT any cross-CTA exchange, for matrices that partial pivoting leaves alone below the diagonal
// block (the HPS merge matrices D; measured: the 128 per-column elections through L2 cost 3 us each and make up
// two thirds of the pivoted kernel, profiles/r02_blockcol_phase_timing.txt):
//   1. one CTA factors the jb x jb diagonal block (pivot search restricted to the block),
//   2. its unit-lower and upper inverses (the solves need both anyway),
//   3. L21 = A21 U11^-1 as ONE tensor-core product over all rows below,
//   4. commit + verification that every multiplier is <= 1 in magnitude — exactly the condition under which
//      partial pivoting over the whole column would have chosen the same pivots.  Otherwise info = -2 and the
//      caller repeats the operation with hps_lu_set_speculative(0).
int factor_block_column_spec(cudaStream_t st, int batch, int n, const Mat& A, int j, int jb, LuWorkspace& w, int* info,
                             bool in_block_pivoting, double bound) {
  const int nblk = (n + NB - 1) / NB;
  const int64_t sW = (int64_t)nblk * NB * NB;
  static const bool force_pivot = [] { const char* e = std::getenv("HPS_DIAGBLK"); return e && e[0] == 'p'; }();
  const bool pivot_in_block = in_block_pivoting || force_pivot;
  prof_begin(PROF_PANEL, st, (double)batch * jb * jb);
  for (int b0 = 0; b0 < batch; b0 += 65535) {
    const int nb = std::min(65535, batch - b0);
    if (pivot_in_block)
      diagblk_kernel<true><<<nb, DB_THREADS, DB_SMEM, st>>>(A.p + (int64_t)b0 * A.stride, A.ld, A.stride, n, j, jb,
                                                            w.ipiv + (int64_t)b0 * n, info + b0);
    else
      diagblk_kernel<false><<<nb, DB_THREADS, DB_SMEM, st>>>(A.p + (int64_t)b0 * A.stride, A.ld, A.stride, n, j, jb,
                                                             w.ipiv + (int64_t)b0 * n, info + b0);
  }
  prof_end(PROF_PANEL, st);
  HPS_LAUNCH_CHECK("diagonal block LU");
  prof_begin(PROF_TRTRI, st, (double)batch * NB * NB * NB * 2 / 3);
  trtri_pair_kernel<<<dim3(2, batch), TRI_THREADS, TRTRI_PAIR_SMEM, st>>>(A.p, A.ld, A.stride, j, n, w.Linv, w.Uinv, sW);
  prof_end(PROF_TRTRI, st);
  HPS_LAUNCH_CHECK("trtri kernels");
  const int rows = n - (j + jb);
  if (rows > 0) {
    const int64_t sP = (int64_t)n * NB;
    HPS_TRY(dgemm(st, rows, jb, jb, 1.0, A.at(j + jb, j), A.ld, A.stride, w.Uinv + (int64_t)(j / NB) * NB * NB, NB, sW, 0.0,
                  w.P, jb, sP, batch));
    const int64_t total = (int64_t)rows * jb;
    spec_commit_kernel<<<dim3((unsigned)std::min<int64_t>((total + 255) / 256, 1184), batch), 256, 0, st>>>(
        A.at(j + jb, j), A.ld, A.stride, w.P, sP, rows, jb, info, bound);
    HPS_LAUNCH_CHECK("spec_commit_kernel");
  }
  return 0;
}

// Factor the outer block column j (inner IB panels + updates inside the block column) and invert
// its unit-lower diagonal block into Linv[j/NB].  Touches columns [j, j+jb) only.
// speculate: 0 = pivoted kernels, 1 = speculative with the diagonal taken as pivot, 2 = speculative with partial
// pivoting inside the diagonal block (matrices that interchange rows locally, e.g. the adaptive interface systems)
// spec_bound: largest multiplier accepted below the diagonal block.  1 (+ rounding): exactly partial pivoting's
// choices; 4: THRESHOLD pivoting with u = 1/4 (the block's pivot is kept when it is within a factor 4 of the column
// maximum — the classical relaxation of sparse direct solvers), used for the leaves' A_ii, which partial pivoting
// does reshuffle (71 of 1000 rows at p = 12, by at most 100 positions) although the unpivoted multipliers stay <= 1.16.
constexpr double SPEC_BOUND_EXACT = 1.0 + 1e-8, SPEC_BOUND_THRESHOLD = 4.0;
int factor_block_column(cudaStream_t st, int batch, int n, const Mat& A, int j, int jb, LuWorkspace& w, int* info,
                        int speculate = 0, double spec_bound = SPEC_BOUND_EXACT) {
  if (speculate) return factor_block_column_spec(st, batch, n, A, j, jb, w, info, speculate == 2, spec_bound);
  bool done = false;
  HPS_TRY(launch_blockcol(st, batch, n, A.p, A.ld, A.stride, j, jb, w, info, done));
  for (int jj = j; !done && jj < j + jb; jj += IB) {
    const int ib = min(IB, j + jb - jj);
    PanelArgs pa;
    pa.A = A.p; pa.lda = A.ld; pa.sA = A.stride; pa.n = n; pa.jj = jj; pa.ib = ib; pa.G = 1;
    pa.ipiv = w.ipiv; pa.info = info; pa.scratch = w.scratch;
    HPS_TRY(launch_panel(st, batch, pa));
    HPS_TRY(laswp(st, batch, A.p, A.ld, A.stride, j, jj - j, w.ipiv, n, jj, jj + ib));
    const int right = j + jb - (jj + ib);
    if (right > 0) {
      HPS_TRY(laswp(st, batch, A.p, A.ld, A.stride, jj + ib, right, w.ipiv, n, jj, jj + ib));
      prof_begin(PROF_INNER, st, (double)batch * right * ib * ib);
      inner_trsm_kernel<<<dim3((right + 127) / 128, batch), 128, 0, st>>>(A.p, A.ld, A.stride, jj, ib, jj + ib, right);
      prof_end(PROF_INNER, st);
      HPS_LAUNCH_CHECK("inner_trsm_kernel");
      const int below = n - (jj + ib);
      if (below > 0)
        HPS_TRY(dgemm(st, below, right, ib, -1.0, A.at(jj + ib, jj), A.ld, A.stride, A.at(jj, jj + ib), A.ld, A.stride,
                      1.0, A.at(jj + ib, jj + ib), A.ld, A.stride, batch));
    }
  }
  const int nblk = (n + NB - 1) / NB;
  prof_begin(PROF_TRTRI, st, (double)batch * NB * NB * NB / 3);
  trtri_lower_kernel<<<dim3(1, batch), TRI_THREADS, TRTRI_SMEM, st>>>(A.p, A.ld, A.stride, j, n, w.Linv, (int64_t)nblk * NB * NB);
  prof_end(PROF_TRTRI, st);
  HPS_LAUNCH_CHECK("trtri_lower_kernel");
  return 0;
}

// Apply block column j's interchanges, U12 = L11^-1 A12 and the rank-jb update to columns
// [c0, c0+nc) of A (rows >= j).
int update_columns(cudaStream_t st, int batch, int n, const Mat& A, int j, int jb, int c0, int nc, LuWorkspace& w) {
  if (nc <= 0) return 0;
  const int nblk = (n + NB - 1) / NB;
  const int64_t sW = (int64_t)nblk * NB * NB;
  const double* Linv = w.Linv + (int64_t)(j / NB) * NB * NB;
  HPS_TRY(laswp(st, batch, A.p, A.ld, A.stride, c0, nc, w.ipiv, n, j, j + jb));
  if (c0 >= j + jb) {  // right of the block column: needs U12 and the trailing update
    HPS_TRY(tri_mult(st, batch, jb, Linv, sW, A.at(j, c0), A.ld, A.stride, nc, w.tmp));
    const int below = n - (j + jb);
    if (below > 0)
      HPS_TRY(dgemm(st, below, nc, jb, -1.0, A.at(j + jb, j), A.ld, A.stride, A.at(j, c0), A.ld, A.stride, 1.0,
                    A.at(j + jb, c0), A.ld, A.stride, batch));
  }
  return 0;
}

// L Z = B in place on rows [r0, r1) of X (unit lower, diagonal blocks pre-inverted).  With `structured` the
// leading zero rows described by X.seg_first_row are never touched: a node of the recursion works on the columns
// that can be non-zero in its row range, which is a prefix of X because the segments are sorted.
int trsm_lower(cudaStream_t st, int batch, int n, const Mat& A, const LuWorkspace& w, const RhsDesc& X, int r0, int r1,
               bool structured = false) {
  const int nblk = (n + NB - 1) / NB;
  const int64_t sW = (int64_t)nblk * NB * NB;
  const int nc = structured ? X.active_cols(r1) : X.ncols;  // columns with a possibly non-zero row below r1
  if (nc <= 0) return 0;
  if (r1 - r0 <= NB)
    return tri_mult(st, batch, r1 - r0, w.Linv + (int64_t)(r0 / NB) * NB * NB, sW, X.ptr + (int64_t)r0 * X.ld, X.ld,
                    X.stride, nc, w.tmp);
  const int blocks = (r1 - r0 + NB - 1) / NB;
  const int mid = r0 + (blocks / 2) * NB;
  HPS_TRY(trsm_lower(st, batch, n, A, w, X, r0, mid, structured));
  if (!structured || X.n_seg <= 0) {  // (a right-hand side without declared structure inside a structured solve)
    HPS_TRY(dgemm(st, r1 - mid, X.ncols, mid - r0, -1.0, A.at(mid, r0), A.ld, A.stride, X.ptr + (int64_t)r0 * X.ld, X.ld,
                  X.stride, 1.0, X.ptr + (int64_t)mid * X.ld, X.ld, X.stride, batch));
  } else {
    // X[mid:r1) -= L[mid:r1, r0:mid) Z[r0:mid), column group by column group: a group whose first non-zero row is f
    // has Z[r0:f) == 0, so its product starts at row max(r0, f) (rounded down to an even row: 16-byte alignment);
    // groups that share the same start go out as one product, groups with f >= mid contribute nothing
    int k = 0;
    while (k < X.n_seg) {
      const int f = X.seg_first_row[k];
      if (f >= mid) break;
      const int ks = std::max(r0, f & ~1);
      int k2 = k + 1;
      while (k2 < X.n_seg && std::max(r0, X.seg_first_row[k2] & ~1) == ks && X.seg_first_row[k2] < mid) ++k2;
      const int c_lo = k * X.seg_cols, c_hi = (k2 == X.n_seg) ? X.ncols : k2 * X.seg_cols;
      HPS_TRY(dgemm(st, r1 - mid, c_hi - c_lo, mid - ks, -1.0, A.at(mid, ks), A.ld, A.stride,
                    X.ptr + (int64_t)ks * X.ld + c_lo, X.ld, X.stride, 1.0, X.ptr + (int64_t)mid * X.ld + c_lo, X.ld, X.stride,
                    batch));
      k = k2;
    }
  }
  return trsm_lower(st, batch, n, A, w, X, mid, r1, structured);
}

// info[0] := -1 (if still 0) when some ipiv[k] != k: the structured forward substitution of a run that could not
// ask the host (lu_dist_run) was not valid and the caller must repeat the solve without the structure
__global__ void pivots_moved_info_kernel(const int* __restrict__ ipiv, int n, int* __restrict__ info) {
  bool moved = false;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) moved |= ipiv[i] != i;
  if (__syncthreads_or(moved) && threadIdx.x == 0) atomicCAS(info, 0, -1);
}

// flag[0] != 0 iff some ipiv[k] != k (the factorisation moved rows)
__global__ void pivots_moved_kernel(const int* __restrict__ ipiv, int64_t total, int n, int* __restrict__ flag) {
  bool moved = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    moved |= ipiv[i] != (int)(i % n);
  if (__syncthreads_or(moved) && threadIdx.x == 0) atomicOr(flag, 1);
}

// U X = Z in place on rows [r0, r1) of X
int trsm_upper(cudaStream_t st, int batch, int n, const Mat& A, const LuWorkspace& w, const RhsDesc& X, int r0, int r1) {
  const int nblk = (n + NB - 1) / NB;
  const int64_t sW = (int64_t)nblk * NB * NB;
  if (r1 - r0 <= NB)
    return tri_mult(st, batch, r1 - r0, w.Uinv + (int64_t)(r0 / NB) * NB * NB, sW, X.ptr + (int64_t)r0 * X.ld, X.ld,
                    X.stride, X.ncols, w.tmp);
  const int blocks = (r1 - r0 + NB - 1) / NB;
  const int mid = r0 + (blocks / 2) * NB;
  HPS_TRY(trsm_upper(st, batch, n, A, w, X, mid, r1));
  HPS_TRY(dgemm(st, mid - r0, X.ncols, r1 - mid, -1.0, A.at(r0, mid), A.ld, A.stride, X.ptr + (int64_t)mid * X.ld, X.ld,
                X.stride, 1.0, X.ptr + (int64_t)r0 * X.ld, X.ld, X.stride, batch));
  return trsm_upper(st, batch, n, A, w, X, r0, mid);
}

}  // namespace

size_t lu_workspace_bytes(int batch, int n) {
  const size_t nblk = (n + NB - 1) / NB;
  return align_up((size_t)batch * n * sizeof(int), 256) + 2 * align_up((size_t)batch * nblk * NB * NB * sizeof(double), 256) +
         align_up((size_t)batch * NB * 16 * sizeof(double), 256) + align_up((size_t)batch * sizeof(PanelScratch), 256) +
         align_up((size_t)batch * align_up(bc_scratch_bytes(std::max(1, bc_gcap(n))), 256), 256) +
         align_up((size_t)batch * n * NB * sizeof(double), 256) + 1024;
}

int lu_solve(cudaStream_t st, int batch, int n, double* Ap, int64_t lda, int64_t sA, int n_rhs, const RhsDesc* rhs,
             void* ws, size_t ws_bytes, int* info, int flags) {
  static const bool force_spec = [] { const char* e = std::getenv("HPS_LU_FORCE_SPEC"); return e && e[0] == '1'; }();  // tools/bench_lu.py
  const bool spec_on = ((flags & LU_NO_PIVOT_EXPECTED) || force_spec) && g_lu_speculate.load(std::memory_order_relaxed) != 0;
  const int speculate = spec_on ? ((flags & LU_PIVOT_IN_BLOCK) ? 2 : 1) : 0;
  const double spec_bound = (flags & LU_THRESHOLD_PIVOTING) ? SPEC_BOUND_THRESHOLD : SPEC_BOUND_EXACT;
  if (batch <= 0 || n <= 0) return 0;
  if (batch > 65535) return fail_arg(2, "lu_solve: batch > 65535");
  Arena ar(ws, ws_bytes);
  LuWorkspace w;
  if (!carve(ar, batch, n, w)) return fail_arg(11, "lu_solve: workspace too small");

  HPS_TRY(configure_lu_kernels());
  Aux* aux = nullptr;
  HPS_TRY(aux_for_stream(st, aux));
  cudaStream_t s0 = st, s1 = aux->stream;
  const Mat A{Ap, lda, sA};
  const int nblk = (n + NB - 1) / NB;

  HPS_CUDA(cudaMemsetAsync(info, 0, sizeof(int) * batch, s0));
  if (w.bc_Gcap) HPS_CUDA(cudaMemsetAsync(w.bc_scratch, 0, (size_t)batch * w.bc_stride, s0));  // epoch flags start at 0
  // ---- 1. factorisation with look-ahead --------------------------------------------------
  // s1 owns the block column being factored; s0 applies finished block columns to everything
  // to their right (and the interchanges to everything to their left).
  HPS_CUDA(cudaEventRecord(aux->fork, s0));
  HPS_CUDA(cudaStreamWaitEvent(s1, aux->fork, 0));
  HPS_TRY(factor_block_column(s1, batch, n, A, 0, min(NB, n), w, info, speculate, spec_bound));
  HPS_CUDA(cudaEventRecord(aux->panel_done[0], s1));
  for (int b = 0; b < nblk; ++b) {
    const int j = b * NB, jb = min(NB, n - j);
    const int next = j + jb, nextb = (next < n) ? min(NB, n - next) : 0;
    HPS_CUDA(cudaStreamWaitEvent(s0, aux->panel_done[b & 1], 0));
    if (nextb > 0) {
      // look-ahead: bring the next block column up to date and factor it on s1.  It was last
      // written by s0's update for block b-1, which must have finished.
      if (b > 0) HPS_CUDA(cudaStreamWaitEvent(s1, aux->update_done[(b - 1) & 1], 0));
      HPS_TRY(update_columns(s1, batch, n, A, j, jb, next, nextb, w));
      HPS_TRY(factor_block_column(s1, batch, n, A, next, nextb, w, info, speculate, spec_bound));
      HPS_CUDA(cudaEventRecord(aux->panel_done[(b + 1) & 1], s1));
    }
    // s0: interchanges on the columns to the left (L in LAPACK form), then the rest of the
    // trailing matrix
    // The block column that s1 factors at the NEXT step goes first and gets its own event, so the chain never waits
    // for the bulk of this trailing update.
    const int rest0 = next + nextb, pri = (rest0 < n) ? min(NB, n - rest0) : 0;
    HPS_TRY(update_columns(s0, batch, n, A, j, jb, rest0, pri, w));
    HPS_CUDA(cudaEventRecord(aux->update_done[b & 1], s0));
    HPS_TRY(update_columns(s0, batch, n, A, j, jb, rest0 + pri, n - (rest0 + pri), w));
    HPS_TRY(update_columns(s0, batch, n, A, j, jb, 0, j, w));
  }
  // s1 has nothing pending beyond panel_done[(nblk-1)&1], which s0 has waited on.
  if (n_rhs == 0) return 0;

  // ---- 2. inverses of U's diagonal blocks (all at once), interchanges on the right-hand sides --
  prof_begin(PROF_TRTRI, s0, (double)batch * nblk * NB * NB * NB / 3);
  trtri_upper_kernel<<<dim3(nblk, batch), TRI_THREADS, TRTRI_UPPER_SMEM, s0>>>(A.p, A.ld, A.stride, 0, n, w.Uinv, (int64_t)nblk * NB * NB);
  prof_end(PROF_TRTRI, s0);
  HPS_LAUNCH_CHECK("trtri_upper_kernel");
  // Right-hand sides with declared leading zero rows: the shortcut is valid only if no interchange happened (the
  // HPS merge matrices D never pivot in practice).  ONE host round trip decides; it costs a pipeline bubble of a
  // few tens of microseconds and removes ~37 % of an oct merge's forward-substitution flops.
  bool structured = false;
  for (int k = 0; k < n_rhs; ++k) structured |= rhs[k].n_seg > 0;
  static const bool no_struct = [] { const char* e = std::getenv("HPS_LU_STRUCT"); return e && e[0] == '0'; }();
  if (structured && no_struct) structured = false;
  if (structured) {
    int* flag = reinterpret_cast<int*>(w.tmp);  // tmp is idle until the substitutions start
    HPS_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), s0));
    const int64_t total = (int64_t)batch * n;
    pivots_moved_kernel<<<(unsigned)std::min<int64_t>((total + 255) / 256, 1024), 256, 0, s0>>>(w.ipiv, total, n, flag);
    HPS_LAUNCH_CHECK("pivots_moved_kernel");
    HPS_CUDA(cudaMemcpyAsync(aux->host_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, s0));
    HPS_CUDA(cudaStreamSynchronize(s0));
    structured = *aux->host_flag == 0;
  }
  for (int k = 0; k < n_rhs; ++k) {
    if (!structured)  // with `structured` every ipiv[k] == k: nothing to interchange
      HPS_TRY(laswp(s0, batch, rhs[k].ptr, rhs[k].ld, rhs[k].stride, 0, rhs[k].ncols, w.ipiv, n, 0, n));
    // ---- 3. recursive substitutions --------------------------------------------------------
    HPS_TRY(trsm_lower(s0, batch, n, A, w, rhs[k], 0, n, structured));
    HPS_TRY(trsm_upper(s0, batch, n, A, w, rhs[k], 0, n));
  }
  return 0;
}


// =====================================================================================
// Step-wise entry points for the DISTRIBUTED factorisation of the multi-GPU root merge.
// Every rank holds the full n x n matrix but owns (keeps up to date) only every
// `block_stride`-th block column.  Per block column b the owner factors it, the caller broadcasts
// the packed column (NCCL, from Python), every rank applies it to the block columns it owns.
// After the last step every rank holds the complete P A = L U and solves its own right-hand sides.
// All calls share one workspace (same pointer, lu_workspace_bytes(1, n)).
// =====================================================================================

namespace {
__global__ void pack_block_kernel(int n, int jb, const double* __restrict__ A, int64_t lda, int j,
                                  const int* __restrict__ ipiv, const double* __restrict__ Linv, double* __restrict__ buf,
                                  int unpack, double* __restrict__ A_out, int* __restrict__ ipiv_out,
                                  double* __restrict__ Linv_out) {
  const int64_t n_col = (int64_t)n * jb;
  const int64_t total = n_col + jb + (int64_t)NB * NB;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    if (e < n_col) {
      const int64_t r = e / jb, c = e - r * jb;
      if (unpack) A_out[r * lda + j + c] = buf[e]; else buf[e] = A[r * lda + j + c];
    } else if (e < n_col + jb) {
      const int k = (int)(e - n_col);
      if (unpack) ipiv_out[j + k] = (int)buf[e]; else buf[e] = (double)ipiv[j + k];
    } else {
      const int64_t k = e - n_col - jb;
      if (unpack) Linv_out[k] = buf[e]; else buf[e] = Linv[k];
    }
  }
}
}  // namespace

void lu_set_speculative(int on) { g_lu_speculate.store(on ? 1 : 0, std::memory_order_relaxed); }

size_t lu_dist_block_buffer_doubles(int n) { return (size_t)n * NB + NB + (size_t)NB * NB; }

static int dist_setup(void* ws, size_t ws_bytes, int n, LuWorkspace& w) {
  Arena ar(ws, ws_bytes);
  if (!carve(ar, 1, n, w)) return fail_arg(5, "lu_dist: workspace too small");
  HPS_TRY(configure_lu_kernels());
  return 0;
}

// owner: factor block column b (must be up to date) and pack it, its pivots and the inverse of its
// unit-lower diagonal block into buf
int lu_dist_factor_pack(cudaStream_t st, int n, double* A, int64_t lda, int b, void* ws, size_t ws_bytes, int* info,
                        double* buf) {
  LuWorkspace w;
  HPS_TRY(dist_setup(ws, ws_bytes, n, w));
  const int j = b * NB, jb = min(NB, n - j);
  if (j >= n) return fail_arg(4, "block index out of range");
  const Mat Am{A, lda, 0};
  // the exchange flags of the block-column kernel must not hold a stale epoch from another use of ws
  if (w.bc_Gcap) HPS_CUDA(cudaMemsetAsync(w.bc_scratch, 0, w.bc_stride, st));
  HPS_TRY(factor_block_column(st, 1, n, Am, j, jb, w, info));
  const double* Linv = w.Linv + (int64_t)b * NB * NB;
  pack_block_kernel<<<1024, 256, 0, st>>>(n, jb, A, lda, j, w.ipiv, Linv, buf, 0, nullptr, nullptr, nullptr);
  HPS_LAUNCH_CHECK("pack_block_kernel");
  return 0;
}

// non-owner: install the broadcast block column
int lu_dist_unpack(cudaStream_t st, int n, double* A, int64_t lda, int b, void* ws, size_t ws_bytes, const double* buf) {
  LuWorkspace w;
  HPS_TRY(dist_setup(ws, ws_bytes, n, w));
  const int j = b * NB, jb = min(NB, n - j);
  double* Linv = w.Linv + (int64_t)b * NB * NB;
  pack_block_kernel<<<1024, 256, 0, st>>>(n, jb, nullptr, lda, j, nullptr, nullptr, const_cast<double*>(buf), 1, A, w.ipiv,
                                          Linv);
  HPS_LAUNCH_CHECK("pack_block_kernel(unpack)");
  return 0;
}

// every rank: apply block column b to the block columns first_block + i*block_stride (i < n_blocks,
// all > b) that it owns, and its interchanges to the columns on the left
int lu_dist_update(cudaStream_t st, int n, double* A, int64_t lda, int b, int first_block, int n_blocks,
                   int block_stride, int apply_left, void* ws, size_t ws_bytes) {
  LuWorkspace w;
  HPS_TRY(dist_setup(ws, ws_bytes, n, w));
  const int j = b * NB, jb = min(NB, n - j);
  const Mat Am{A, lda, 0};
  if (apply_left) HPS_TRY(laswp(st, 1, A, lda, 0, 0, j, w.ipiv, n, j, j + jb));  // L in LAPACK form
  if (n_blocks <= 0) return 0;
  if (first_block <= b) return fail_arg(6, "owned blocks must lie to the right of b");
  const double* Linv = w.Linv + (int64_t)b * NB * NB;
  const int below = n - (j + jb);
  // full-width owned blocks as one strided batch; a ragged last block separately
  int full = n_blocks;
  const int last_c0 = (first_block + (n_blocks - 1) * block_stride) * NB;
  const bool ragged = last_c0 + NB > n;
  if (ragged) --full;
  const int64_t sBlk = (int64_t)block_stride * NB;
  auto apply = [&](int c0, int nc, int batch) -> int {
    prof_begin(PROF_LASWP, st, (double)batch * nc * jb);
    launch_laswp(st, batch, A, lda, sBlk, c0, nc, w.ipiv, 0, j, j + jb);
    prof_end(PROF_LASWP, st);
    HPS_LAUNCH_CHECK("laswp_kernel");
    HPS_TRY(tri_mult(st, batch, jb, Linv, 0, Am.at(j, c0), lda, sBlk, nc, w.tmp));
    if (below > 0)
      HPS_TRY(dgemm(st, below, nc, jb, -1.0, Am.at(j + jb, j), lda, 0, Am.at(j, c0), lda, sBlk, 1.0, Am.at(j + jb, c0), lda,
                    sBlk, batch));
    return 0;
  };
  if (full > 0) HPS_TRY(apply(first_block * NB, NB, full));
  if (ragged) HPS_TRY(apply(last_c0, n - last_c0, 1));
  return 0;
}

// every rank, after the last block column: U's diagonal inverses, interchanges on the right-hand
// sides and the two recursive substitutions
int lu_dist_solve(cudaStream_t st, int n, double* A, int64_t lda, int n_rhs, const RhsDesc* rhs, void* ws,
                  size_t ws_bytes) {
  LuWorkspace w;
  HPS_TRY(dist_setup(ws, ws_bytes, n, w));
  const Mat Am{A, lda, 0};
  const int nblk = (n + NB - 1) / NB;
  prof_begin(PROF_TRTRI, st, (double)nblk * NB * NB * NB / 3);
  trtri_upper_kernel<<<dim3(nblk, 1), TRI_THREADS, TRTRI_UPPER_SMEM, st>>>(A, lda, 0, 0, n, w.Uinv, (int64_t)nblk * NB * NB);
  prof_end(PROF_TRTRI, st);
  HPS_LAUNCH_CHECK("trtri_upper_kernel");
  for (int k = 0; k < n_rhs; ++k) {
    HPS_TRY(laswp(st, 1, rhs[k].ptr, rhs[k].ld, rhs[k].stride, 0, rhs[k].ncols, w.ipiv, n, 0, n));
    HPS_TRY(trsm_lower(st, 1, n, Am, w, rhs[k], 0, n));
    HPS_TRY(trsm_upper(st, 1, n, Am, w, rhs[k], 0, n));
  }
  return 0;
}


// =====================================================================================
// Library-owned communicator: one SYMMETRIC device segment per rank (cudaMalloc + CUDA IPC), mapped into
// every peer of the box, so that kernels store straight into the peers' HBM over NVLink / NVSwitch.
// The distributed root factorisation below uses it for a fused "copy the factored block column to
// every peer + signal": no pack buffer, no NCCL broadcast, no unpack, and the whole block-column loop
// is issued from C.
//
// Segment layout (identical on every rank):
//   [0, 4 KB)        ready[8]   barrier words, 128 bytes apart (written by the peers)
//   [4 KB, 64 KB)    blk_flag[] one epoch word per block column (written by the column's owner)
//   [64 KB, ...)     payload: ipiv[n] | Linv[nblk][NB][NB] | A[n][n]   (distributed LU)
// Epochs grow by one per collective operation, so no flag is ever reset.
// =====================================================================================
constexpr size_t COMM_READY_OFF = 0, COMM_FLAG_OFF = 4096, COMM_PAYLOAD_OFF = 65536;
constexpr int COMM_MAX_WORLD = 8;
constexpr int COMM_MAX_BLOCKS = (int)((COMM_PAYLOAD_OFF - COMM_FLAG_OFF) / sizeof(unsigned));

struct CommPtrs { char* peer[COMM_MAX_WORLD]; };

struct Comm {
  int rank = 0, world = 1, device = 0;
  size_t bytes = 0;          // size of the local segment
  char* local = nullptr;
  CommPtrs ptrs{};           // ptrs.peer[rank] == local; the others are IPC mappings (null until attached)
  bool attached = false;
  unsigned epoch = 0;
  unsigned* done = nullptr;  // completion counter of the copy kernel (local, not shared)
};

namespace {

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

// Every rank tells every peer "I have reached epoch e" and waits until all peers said so.  <<<1, 32>>>
__global__ void comm_barrier_kernel(CommPtrs p, int rank, int world, unsigned epoch) {
  const int t = threadIdx.x;
  if (t < world) {
    unsigned* remote = reinterpret_cast<unsigned*>(p.peer[t] + COMM_READY_OFF) + rank * 32;
    st_release_sys(remote, epoch);
    const unsigned* mine = reinterpret_cast<const unsigned*>(p.peer[rank] + COMM_READY_OFF) + t * 32;
    while ((int)(ld_acquire_sys(mine) - epoch) < 0) { }
  }
}

// <<<1, 1>>>: stream-ordered wait until block column b of this epoch has arrived in the local segment
__global__ void comm_wait_block_kernel(const unsigned* flag, unsigned epoch) {
  while ((int)(ld_acquire_sys(flag) - epoch) < 0) { }
}

// Fused copy + signal: the owner stores its factored block column (all n rows: U above, L below), the
// pivots and the inverted unit-lower diagonal block into every peer's segment, then releases flag b there.
// Each element is read once and written world-1 times; the last CTA to finish raises the flags.
__global__ void __launch_bounds__(256) comm_bcast_blockcol_kernel(CommPtrs p, int rank, int world, unsigned mask, int n,
                                                                  int j, int jb, size_t ipiv_off, size_t linv_off,
                                                                  size_t A_off, int b, unsigned epoch, unsigned* done) {
  const char* src = p.peer[rank];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool vec = ((n | j | jb) & 1) == 0;
  if (vec) {
    const int h = jb >> 1;
    const int64_t total = (int64_t)n * h;
    for (int64_t e = t0; e < total; e += stride) {
      const int64_t r = e / h, c = e - r * h;
      const size_t off = A_off + ((size_t)r * n + j) * sizeof(double) + (size_t)c * sizeof(double2);
      const double2 v = *reinterpret_cast<const double2*>(src + off);
      for (int q = 0; q < world; ++q)
        if ((mask >> q) & 1u) *reinterpret_cast<double2*>(p.peer[q] + off) = v;
    }
  } else {
    const int64_t total = (int64_t)n * jb;
    for (int64_t e = t0; e < total; e += stride) {
      const int64_t r = e / jb, c = e - r * jb;
      const size_t off = A_off + ((size_t)r * n + j + c) * sizeof(double);
      const double v = *reinterpret_cast<const double*>(src + off);
      for (int q = 0; q < world; ++q)
        if ((mask >> q) & 1u) *reinterpret_cast<double*>(p.peer[q] + off) = v;
    }
  }
  for (int64_t e = t0; e < jb; e += stride) {
    const size_t off = ipiv_off + (size_t)(j + e) * sizeof(int);
    const int v = *reinterpret_cast<const int*>(src + off);
    for (int q = 0; q < world; ++q)
      if ((mask >> q) & 1u) *reinterpret_cast<int*>(p.peer[q] + off) = v;
  }
  for (int64_t e = t0; e < (int64_t)NB * NB / 2; e += stride) {
    const size_t off = linv_off + (size_t)e * sizeof(double2);
    const double2 v = *reinterpret_cast<const double2*>(src + off);
    for (int q = 0; q < world; ++q)
      if ((mask >> q) & 1u) *reinterpret_cast<double2*>(p.peer[q] + off) = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(done, 1u);
    if (prev == gridDim.x - 1) {
      *done = 0;  // launches of this kernel are stream-ordered
      __threadfence_system();
      for (int q = 0; q < world; ++q)
        if ((mask >> q) & 1u) st_release_sys(reinterpret_cast<unsigned*>(p.peer[q] + COMM_FLAG_OFF) + b, epoch);
    }
  }
}

constexpr int DIST_NSLOT = 2 * COMM_MAX_WORLD + 2;
constexpr size_t SLOT_LINV_OFF = 1024, SLOT_PANEL_OFF = SLOT_LINV_OFF + (size_t)NB * NB * sizeof(double);

// PACK = true : slot <- (ipiv[j:j+jb], Linv_b, A[:, j:j+jb])   (owner, after the factorisation of block column b)
// PACK = false: the reverse (receiver, after the flag of block column b)
template <bool PACK>
__global__ void __launch_bounds__(256) comm_slot_kernel(double* __restrict__ A, int n, int j, int jb, int* __restrict__ ipiv_j,
                                                        double* __restrict__ Linv_b, char* __restrict__ slot) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double* panel = reinterpret_cast<double*>(slot + SLOT_PANEL_OFF);
  if (((n | jb) & 1) == 0) {  // j is a multiple of NB: 16-byte accesses on both sides
    const int h = jb >> 1;
    const int64_t total = (int64_t)n * h;
    for (int64_t e = t0; e < total; e += stride) {
      const int64_t r = e / h, c = e - r * h;
      double2* a = reinterpret_cast<double2*>(A + r * n + j) + c;
      double2* q = reinterpret_cast<double2*>(panel + r * jb) + c;
      if (PACK) *q = *a; else *a = *q;
    }
  } else {
    const int64_t total = (int64_t)n * jb;
    for (int64_t e = t0; e < total; e += stride) {
      const int64_t r = e / jb, c = e - r * jb;
      if (PACK) panel[e] = A[r * n + j + c]; else A[r * n + j + c] = panel[e];
    }
  }
  int* sp = reinterpret_cast<int*>(slot);
  for (int64_t e = t0; e < jb; e += stride) {
    if (PACK) sp[e] = ipiv_j[e]; else ipiv_j[e] = sp[e];
  }
  double2* sl = reinterpret_cast<double2*>(slot + SLOT_LINV_OFF);
  double2* li = reinterpret_cast<double2*>(Linv_b);
  for (int64_t e = t0; e < (int64_t)NB * NB / 2; e += stride) {
    if (PACK) sl[e] = li[e]; else li[e] = sl[e];
  }
}

// <<<1, 32>>>: raise flag b on the peers in `mask` (after stream-ordered copy-engine transfers)
__global__ void comm_signal_kernel(CommPtrs p, int world, unsigned mask, int b, unsigned epoch) {
  const int q = threadIdx.x;
  if (q < world && ((mask >> q) & 1u)) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<unsigned*>(p.peer[q] + COMM_FLAG_OFF) + b, epoch);
  }
}

// Staging slots of the block-column exchange.  A factored block column is a strided window of the row-major
// matrix (n rows of jb doubles): the copy engines move such a window at ~30 GB/s (per-row overhead) but a
// CONTIGUOUS buffer at NVLink speed, so the owner packs the column (+ pivots + inverted diagonal block) into a
// slot, one contiguous peer copy per destination carries it into the SAME slot of the peer's segment, and the
// peer unpacks it into its matrix (two local HBM passes of 20 MB: ~10 us each).  Slot b % DIST_NSLOT is reused by
// block b + DIST_NSLOT; by then every rank has factored a block column later than b, which it could only do after
// receiving AND unpacking block b (stream order on its chain stream), so DIST_NSLOT >= world + 1 is enough.

struct DistLayout {
  size_t ipiv_off, linv_off, uinv_off, A_off, slot_off, slot_bytes, bytes;
};
DistLayout dist_layout(int n) {
  const size_t nblk = (n + NB - 1) / NB;
  DistLayout L;
  L.ipiv_off = COMM_PAYLOAD_OFF;
  L.linv_off = L.ipiv_off + align_up((size_t)n * sizeof(int), 256);
  L.uinv_off = L.linv_off + align_up(nblk * NB * NB * sizeof(double), 256);  // local only (never sent)
  L.A_off = L.uinv_off + align_up(nblk * NB * NB * sizeof(double), 256);
  L.slot_off = L.A_off + align_up((size_t)n * n * sizeof(double), 256);
  L.slot_bytes = align_up(SLOT_PANEL_OFF + (size_t)n * NB * sizeof(double), 256);
  L.bytes = L.slot_off + (size_t)DIST_NSLOT * L.slot_bytes;
  return L;
}

// block column b applied to `batch` column groups of width nc starting at X (leading dimension ld, groups
// sX apart): interchanges, U12 = L11^-1 X[j:j+jb], X[j+jb:] -= L21 U12.  Used for the owned block columns
// of A and, identically, for the right-hand sides carried along as extra columns.
int dist_apply(cudaStream_t st, int n, const Mat& Am, int j, int jb, const LuWorkspace& w, int b, double* X, int64_t ld,
               int64_t sX, int nc, int batch) {
  if (nc <= 0 || batch <= 0) return 0;
  const double* Linv = w.Linv + (int64_t)b * NB * NB;
  const int below = n - (j + jb);
  prof_begin(PROF_LASWP, st, (double)batch * nc * jb);
  launch_laswp(st, batch, X, ld, sX, 0, nc, w.ipiv, 0, j, j + jb);
  prof_end(PROF_LASWP, st);
  HPS_LAUNCH_CHECK("laswp_kernel");
  HPS_TRY(tri_mult(st, batch, jb, Linv, 0, X + (int64_t)j * ld, ld, sX, nc, w.tmp));
  if (below > 0)
    HPS_TRY(dgemm(st, below, nc, jb, -1.0, Am.at(j + jb, j), Am.ld, 0, X + (int64_t)j * ld, ld, sX, 1.0,
                  X + (int64_t)(j + jb) * ld, ld, sX, batch));
  return 0;
}

}  // namespace

int comm_create(int rank, int world, Comm** out) {
  if (!out) return fail_arg(3, "null output pointer");
  if (world < 1 || world > COMM_MAX_WORLD || rank < 0 || rank >= world) return fail_arg(1, "rank / world out of range (world <= 8)");
  Comm* c = new Comm();
  c->rank = rank; c->world = world;
  HPS_CUDA(cudaGetDevice(&c->device));
  HPS_CUDA(cudaMalloc(&c->done, 256));
  HPS_CUDA(cudaMemset(c->done, 0, 256));
  *out = c;
  return 0;
}

int comm_detach(Comm* c) {
  if (!c) return fail_arg(1, "null communicator");
  for (int q = 0; q < c->world; ++q) {
    if (q != c->rank && c->ptrs.peer[q]) HPS_CUDA(cudaIpcCloseMemHandle(c->ptrs.peer[q]));
    c->ptrs.peer[q] = nullptr;
  }
  c->attached = false;
  return 0;
}

// Local (re)allocation.  *changed = 1 when the segment was replaced: the caller must then run the
// export / all-gather / attach sequence on every rank (sizes are the same everywhere, so all ranks change together).
// Call comm_detach on every rank (and synchronise the ranks) BEFORE a segment that peers have mapped is replaced.
int comm_reserve(Comm* c, size_t bytes, int* changed) {
  if (!c || !changed) return fail_arg(1, "null argument");
  *changed = 0;
  if (c->local && c->bytes >= bytes) return 0;
  if (c->attached) return fail_arg(1, "detach the peers before growing the segment");
  HPS_CUDA(cudaDeviceSynchronize());
  if (c->local) HPS_CUDA(cudaFree(c->local));
  c->local = nullptr;
  HPS_CUDA(cudaMalloc(&c->local, bytes));
  HPS_CUDA(cudaMemset(c->local, 0, COMM_PAYLOAD_OFF));  // flags and barrier words start at epoch 0
  HPS_CUDA(cudaDeviceSynchronize());
  c->bytes = bytes;
  c->epoch = 0;
  *changed = 1;
  return 0;
}

int comm_export(Comm* c, void* handle64) {
  if (!c || !c->local || !handle64) return fail_arg(1, "no local segment to export");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  HPS_CUDA(cudaIpcGetMemHandle(&h, c->local));
  memcpy(handle64, &h, 64);
  return 0;
}

int comm_attach(Comm* c, const void* handles) {
  if (!c || !c->local || (!handles && c->world > 1)) return fail_arg(1, "no local segment / handles");
  for (int q = 0; q < c->world; ++q) {
    if (q == c->rank) { c->ptrs.peer[q] = c->local; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char*>(handles) + (size_t)q * 64, 64);
    void* ptr = nullptr;
    HPS_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    c->ptrs.peer[q] = static_cast<char*>(ptr);
  }
  c->attached = true;
  return 0;
}

int comm_destroy(Comm* c) {
  if (!c) return 0;
  comm_detach(c);
  if (c->local) cudaFree(c->local);
  if (c->done) cudaFree(c->done);
  delete c;
  return 0;
}

size_t lu_dist_segment_bytes(int n) { return dist_layout(n).bytes; }

int lu_dist_matrix_ptr(Comm* c, int n, double** A) {
  if (!c || !A) return fail_arg(1, "null argument");
  const DistLayout L = dist_layout(n);
  if (!c->local || c->bytes < L.bytes) return fail_arg(2, "segment too small: hps_comm_reserve(hps_lu_dist_segment_bytes(n)) first");
  *A = reinterpret_cast<double*>(c->local + L.A_off);
  return 0;
}

// Whole distributed factorisation + solves, enqueued on `st` (and an internal look-ahead stream); no host sync.
// A (n x n, lda = n) must have been assembled at lu_dist_matrix_ptr on EVERY rank.  Block column b is factored by
// rank b % world (block-column kernel), stored into every peer's segment and signalled; every rank applies it to
// the block columns it owns AND to its own right-hand sides, which ride along as extra trailing columns — the
// forward substitution therefore finishes with the factorisation, hidden under the panel chain.  What remains
// afterwards is U's diagonal inverses and the recursive backward substitution on each rank's right-hand sides.
int lu_dist_run(Comm* c, cudaStream_t st, int n, int n_rhs, const RhsDesc* rhs, void* ws, size_t ws_bytes, int* info) {
  if (!c || !c->local) return fail_arg(1, "communicator has no segment");
  if (c->world > 1 && !c->attached) return fail_arg(1, "peers are not attached");
  if (n <= 0) return 0;
  const DistLayout L = dist_layout(n);
  if (c->bytes < L.bytes) return fail_arg(3, "segment too small for n");
  const int nblk = (n + NB - 1) / NB;
  if (nblk > COMM_MAX_BLOCKS) return fail_arg(3, "too many block columns for the flag area");
  Arena ar(ws, ws_bytes);
  LuWorkspace w;
  if (!carve(ar, 1, n, w)) return fail_arg(7, "lu_dist_run: workspace too small");
  HPS_TRY(configure_lu_kernels());
  // pivots and inverted diagonal blocks live in the symmetric segment: their owners store them there
  w.ipiv = reinterpret_cast<int*>(c->local + L.ipiv_off);
  w.Linv = reinterpret_cast<double*>(c->local + L.linv_off);
  w.Uinv = reinterpret_cast<double*>(c->local + L.uinv_off);  // stays with the factors for lu_dist_apply
  double* A = reinterpret_cast<double*>(c->local + L.A_off);
  const Mat Am{A, n, 0};
  const int rank = c->rank, world = c->world;
  const unsigned epoch = ++c->epoch;
  unsigned* flags = reinterpret_cast<unsigned*>(c->local + COMM_FLAG_OFF);
  Aux* aux = nullptr;
  HPS_TRY(aux_for_stream(st, aux));
  cudaStream_t s0 = st, s1 = aux->stream;

  HPS_CUDA(cudaMemsetAsync(info, 0, sizeof(int), s0));
  if (w.bc_Gcap) HPS_CUDA(cudaMemsetAsync(w.bc_scratch, 0, w.bc_stride, s0));
  // every rank's matrix is assembled (and nobody still uses the segment from the previous operation)
  if (world > 1) {
    comm_barrier_kernel<<<1, 32, 0, s0>>>(c->ptrs, rank, world, epoch);
    HPS_LAUNCH_CHECK("comm_barrier_kernel");
  }
  HPS_CUDA(cudaEventRecord(aux->fork, s0));
  HPS_CUDA(cudaStreamWaitEvent(s1, aux->fork, 0));

  const int speculate = g_lu_speculate.load(std::memory_order_relaxed) != 0 ? 1 : 0;  // the root D of a merge: see lu_solve
  bool structured = false;
  for (int k = 0; k < n_rhs; ++k) structured |= rhs[k].n_seg > 0;
  {
    static const bool no_struct = [] { const char* e = std::getenv("HPS_LU_STRUCT"); return e && e[0] == '0'; }();
    if (no_struct) structured = false;
  }
  // The right-hand sides ride along with the factorisation (rank-128 updates, ~28 TF/s) when that fills time the
  // caller's stream would otherwise spend waiting for the panel chain; when the trailing updates of this rank's own
  // block columns already outlast the chain (~0.4 ms per block column) — the L=4 root, n = 76 800 — the forward
  // substitution runs afterwards instead, recursively, at the large-K rate (~35 TF/s).  HPS_DIST_RIDE=0/1 forces it.
  bool ride_along = 2.0 * n * NB * ((double)n / world) / 30e12 < 1.0e-3;
  {
    static const int forced = [] { const char* e = std::getenv("HPS_DIST_RIDE"); return e ? (e[0] == '0' ? 0 : 1) : -1; }();
    if (forced >= 0) ride_along = forced != 0;
  }
  // HPS_DIST_SEND=kernel: SM stores straight into the peers' matrices; default: pack into a staging slot, one
  // contiguous copy-engine transfer per peer, a one-warp kernel raises the flags, the peer unpacks.
  static const bool send_by_kernel = [] { const char* e = std::getenv("HPS_DIST_SEND"); return e && e[0] == 'k'; }();
  cudaStream_t s2 = aux->comm_stream;
  auto send = [&](cudaStream_t s, int b, unsigned mask, unsigned* done) -> int {
    if (!mask) return 0;
    const int j = b * NB, jb = std::min(NB, n - j);
    const size_t linv_off = L.linv_off + (size_t)b * NB * NB * sizeof(double);
    prof_begin(PROF_COMM, s, (double)n * jb * 8.0 * __builtin_popcount(mask));
    if (send_by_kernel) {
      comm_bcast_blockcol_kernel<<<64, 256, 0, s>>>(c->ptrs, rank, world, mask, n, j, jb, L.ipiv_off, linv_off, L.A_off, b,
                                                    epoch, done);
    } else {
      const size_t so = L.slot_off + (size_t)(b % DIST_NSLOT) * L.slot_bytes;
      const size_t used = SLOT_PANEL_OFF + (size_t)n * jb * sizeof(double);
      for (int q = 0; q < world; ++q) {
        if (!((mask >> q) & 1u)) continue;
        HPS_CUDA(cudaMemcpyAsync(c->ptrs.peer[q] + so, c->local + so, used, cudaMemcpyDeviceToDevice, s));
      }
      comm_signal_kernel<<<1, 32, 0, s>>>(c->ptrs, world, mask, b, epoch);
    }
    prof_end(PROF_COMM, s);
    HPS_LAUNCH_CHECK("block-column send");
    return 0;
  };
  // The owner of the NEXT block column is served first, on the chain's stream; everybody else gets the column
  // from the communication stream, off the critical path.
  auto factor_and_send = [&](int b) -> int {
    const int j = b * NB, jb = std::min(NB, n - j);
    HPS_TRY(factor_block_column(s1, 1, n, Am, j, jb, w, info, speculate));
    if (world > 1) {
      if (!send_by_kernel) {
        comm_slot_kernel<true><<<128, 256, 0, s1>>>(A, n, j, jb, w.ipiv + j, w.Linv + (int64_t)b * NB * NB,
                                                    c->local + L.slot_off + (size_t)(b % DIST_NSLOT) * L.slot_bytes);
        HPS_LAUNCH_CHECK("comm_slot_kernel<pack>");
      }
      const int next = (b + 1) % world;
      const unsigned all = ((1u << world) - 1u) & ~(1u << rank);
      const unsigned first = (b + 1 < nblk && next != rank) ? (1u << next) : 0u;
      HPS_CUDA(cudaEventRecord(aux->factored, s1));
      HPS_CUDA(cudaStreamWaitEvent(s2, aux->factored, 0));
      HPS_TRY(send(s1, b, first, c->done));
      HPS_TRY(send(s2, b, all & ~first, c->done + 32));
    }
    return 0;
  };
  auto wait_block = [&](int b) -> int {
    prof_begin(PROF_WAIT, s1, 0.0);
    comm_wait_block_kernel<<<1, 1, 0, s1>>>(flags + b, epoch);
    prof_end(PROF_WAIT, s1);
    HPS_LAUNCH_CHECK("comm_wait_block_kernel");
    if (!send_by_kernel) {
      const int j = b * NB, jb = std::min(NB, n - j);
      comm_slot_kernel<false><<<128, 256, 0, s1>>>(A, n, j, jb, w.ipiv + j, w.Linv + (int64_t)b * NB * NB,
                                                   c->local + L.slot_off + (size_t)(b % DIST_NSLOT) * L.slot_bytes);
      HPS_LAUNCH_CHECK("comm_slot_kernel<unpack>");
    }
    return 0;
  };
  // owned block columns > lo as one strided batch (+ a ragged last block), then the right-hand sides
  auto apply_owned = [&](cudaStream_t s, int b, int lo, bool with_rhs) -> int {
    const int j = b * NB, jb = std::min(NB, n - j);
    int first = lo + ((rank - lo) % world + world) % world;
    int n_own = first >= nblk ? 0 : (nblk - 1 - first) / world + 1;
    if (n_own > 0) {
      const int last_c0 = (first + (n_own - 1) * world) * NB;
      const bool ragged = last_c0 + NB > n;
      const int full = ragged ? n_own - 1 : n_own;
      if (full > 0) HPS_TRY(dist_apply(s, n, Am, j, jb, w, b, A + (int64_t)first * NB, n, (int64_t)world * NB, NB, full));
      if (ragged) HPS_TRY(dist_apply(s, n, Am, j, jb, w, b, A + last_c0, n, 0, n - last_c0, 1));
    }
    if (with_rhs)
      for (int k = 0; k < n_rhs; ++k) {
        // declared leading zero rows (RhsDesc): block b leaves the columns that are still zero down to row j + jb alone
        const int nc = structured ? rhs[k].active_cols(j + jb) : rhs[k].ncols;
        HPS_TRY(dist_apply(s, n, Am, j, jb, w, b, rhs[k].ptr, rhs[k].ld, 0, nc, 1));
      }
    return 0;
  };

  if (rank == 0 % world) HPS_TRY(factor_and_send(0)); else HPS_TRY(wait_block(0));
  HPS_CUDA(cudaEventRecord(aux->panel_done[0], s1));
  for (int b = 0; b < nblk; ++b) {
    const int nb1 = b + 1;
    const bool own_next = nb1 < nblk && rank == nb1 % world;
    if (nb1 < nblk) {
      if (own_next) {
        // look-ahead: block column nb1 received blocks < b on s0; bring it up to date with block b and factor it
        if (b >= 1) HPS_CUDA(cudaStreamWaitEvent(s1, aux->update_done[(b - 1) & 1], 0));
        const int j = b * NB, jb = std::min(NB, n - j);
        const int c0 = nb1 * NB, nc = std::min(NB, n - c0);
        HPS_TRY(dist_apply(s1, n, Am, j, jb, w, b, A + c0, n, 0, nc, 1));
        HPS_TRY(factor_and_send(nb1));
      } else {
        HPS_TRY(wait_block(nb1));
      }
      HPS_CUDA(cudaEventRecord(aux->panel_done[nb1 & 1], s1));
    }
    HPS_CUDA(cudaStreamWaitEvent(s0, aux->panel_done[b & 1], 0));
    // The block column this rank factors at the NEXT step (b + 2) is brought up to date first and gets its own
    // event: the chain then never waits for the bulk of the trailing update (owned block columns + right-hand sides).
    const int nb2 = b + 2;
    int lo = own_next ? nb1 + 1 : nb1;
    if (nb2 < nblk && rank == nb2 % world) {
      const int j = b * NB, jb = std::min(NB, n - j);
      const int c0 = nb2 * NB, nc = std::min(NB, n - c0);
      HPS_TRY(dist_apply(s0, n, Am, j, jb, w, b, A + c0, n, 0, nc, 1));
      lo = nb2 + 1;
    }
    HPS_CUDA(cudaEventRecord(aux->update_done[b & 1], s0));
    HPS_TRY(apply_owned(s0, b, lo, ride_along));
  }
  if (world > 1) {  // the caller's stream also covers the sends still in flight on the communication stream
    HPS_CUDA(cudaEventRecord(aux->sent, s2));
    HPS_CUDA(cudaStreamWaitEvent(s0, aux->sent, 0));
  }
  if (structured) {  // the shortcut assumed that no row moved: verified here, reported as info = -1
    pivots_moved_info_kernel<<<64, 256, 0, s0>>>(w.ipiv, n, info);
    HPS_LAUNCH_CHECK("pivots_moved_info_kernel");
  }
  prof_begin(PROF_TRTRI, s0, (double)nblk * NB * NB * NB / 3);
  trtri_upper_kernel<<<dim3(nblk, 1), TRI_THREADS, TRTRI_UPPER_SMEM, s0>>>(A, n, 0, 0, n, w.Uinv, (int64_t)nblk * NB * NB);
  prof_end(PROF_TRTRI, s0);
  HPS_LAUNCH_CHECK("trtri_upper_kernel");
  for (int k = 0; k < n_rhs; ++k) {
    if (!ride_along) {  // forward substitution after the factorisation: recursive, large-K products
      HPS_TRY(laswp(s0, 1, rhs[k].ptr, rhs[k].ld, rhs[k].stride, 0, rhs[k].ncols, w.ipiv, n, 0, n));
      HPS_TRY(trsm_lower(s0, 1, n, Am, w, rhs[k], 0, n, structured));
    }
    HPS_TRY(trsm_upper(s0, 1, n, Am, w, rhs[k], 0, n));
  }
  return 0;
}

// rhs[k] := A^-1 rhs[k] with the factors a previous lu_dist_run left in this rank's segment (every rank holds
// all of them): the "factored root" mode of the sharded solver, where the root S = -D^-1 C is never formed and
// each solve applies D^-1 instead.  Forward substitution in the order of the factorisation (interchanges of block
// b, its inverted diagonal block, rank-jb update of the rows below), then the recursive backward substitution.
// Purely local; ws: lu_workspace_bytes(1, n).
int lu_dist_apply(Comm* c, cudaStream_t st, int n, int n_rhs, const RhsDesc* rhs, void* ws, size_t ws_bytes) {
  if (!c || !c->local) return fail_arg(1, "communicator has no segment");
  if (n <= 0 || n_rhs <= 0) return 0;
  const DistLayout L = dist_layout(n);
  if (c->bytes < L.bytes) return fail_arg(3, "segment too small for n");
  Arena ar(ws, ws_bytes);
  LuWorkspace w;
  if (!carve(ar, 1, n, w)) return fail_arg(6, "lu_dist_apply: workspace too small");
  HPS_TRY(configure_lu_kernels());
  w.ipiv = reinterpret_cast<int*>(c->local + L.ipiv_off);
  w.Linv = reinterpret_cast<double*>(c->local + L.linv_off);
  w.Uinv = reinterpret_cast<double*>(c->local + L.uinv_off);
  double* A = reinterpret_cast<double*>(c->local + L.A_off);
  const Mat Am{A, n, 0};
  const int nblk = (n + NB - 1) / NB;
  for (int k = 0; k < n_rhs; ++k) {
    for (int b = 0; b < nblk; ++b) {
      const int j = b * NB, jb = std::min(NB, n - j);
      HPS_TRY(dist_apply(st, n, Am, j, jb, w, b, rhs[k].ptr, rhs[k].ld, 0, rhs[k].ncols, 1));
    }
    HPS_TRY(trsm_upper(st, 1, n, Am, w, rhs[k], 0, n));
  }
  return 0;
}

}  // namespace hps

// Kernels that were measured and NOT adopted (developer reference; not compiled into the library).
// They refer to constants and helpers of jaxhps_b200/csrc/lu.cu (NB, ...).

// Same interchanges for a SHORT pivot range (k1 - k0 <= NB), in two parallel phases instead of k1 - k0 dependent
// swaps per thread (which made every call cost ~0.6 us x 128 whatever the column count — on the critical chain
// of the factorisation):
//   1. warp 0 composes the swaps into a move list "row dst[e] receives old row src[e]" (at most 2 (k1 - k0) rows
//      are touched: the range itself and the distinct pivot rows below it);
//   2. all threads gather the source rows of a 32-column tile into shared memory and scatter them to their
//      destinations (coalesced row segments, nmov/8 independent loads per thread).
// A CTA composes once and then walks over column tiles blockIdx.x, blockIdx.x + gridDim.x, ...
constexpr int LASWP_COLS = 32, LASWP_THREADS = 256;
constexpr size_t LASWP_SMEM = sizeof(double) * 2 * NB * LASWP_COLS;
__global__ void __launch_bounds__(LASWP_THREADS) laswp_block_kernel(double* A, int64_t lda, int64_t sA, int c0, int ncols,
                                                                    const int* ipiv, int n_ipiv, int k0, int k1) {
  extern __shared__ __align__(16) double sm[];  // [nmov][LASWP_COLS]
  __shared__ int s_piv[NB], cur_top[NB], out_row[NB], out_cur[NB], dst[2 * NB], src[2 * NB];
  __shared__ int s_nmov;
  const int nk = k1 - k0, tid = threadIdx.x;
  const int* piv = ipiv + (int64_t)blockIdx.y * n_ipiv + k0;
  for (int i = tid; i < nk; i += LASWP_THREADS) { s_piv[i] = piv[i]; cur_top[i] = k0 + i; }
  __syncthreads();
  if (tid < 32) {
    const int lane = tid;
    int nout = 0;  // warp-uniform
    for (int k = 0; k < nk; ++k) {
      const int p = s_piv[k];
      if (p == k0 + k) continue;
      if (p < k1) {
        if (lane == 0) { const int t = cur_top[k]; cur_top[k] = cur_top[p - k0]; cur_top[p - k0] = t; }
      } else {
        int found = -1;
        for (int base = 0; base < nout; base += 32) {
          const int i = base + lane;
          const unsigned m = __ballot_sync(0xffffffffu, i < nout && out_row[i] == p);
          if (m) { found = base + __ffs(m) - 1; break; }
        }
        if (found < 0) {
          found = nout++;
          if (lane == 0) { out_row[found] = p; out_cur[found] = p; }
        }
        __syncwarp();
        if (lane == 0) { const int t = cur_top[k]; cur_top[k] = out_cur[found]; out_cur[found] = t; }
      }
      __syncwarp();
    }
    int cnt = 0;
    for (int base = 0; base < nk; base += 32) {
      const int i = base + lane;
      const bool mv = i < nk && cur_top[i] != k0 + i;
      const unsigned m = __ballot_sync(0xffffffffu, mv);
      if (mv) { const int pos = cnt + __popc(m & ((1u << lane) - 1u)); dst[pos] = k0 + i; src[pos] = cur_top[i]; }
      cnt += __popc(m);
    }
    for (int base = 0; base < nout; base += 32) {
      const int i = base + lane;
      const bool mv = i < nout && out_cur[i] != out_row[i];
      const unsigned m = __ballot_sync(0xffffffffu, mv);
      if (mv) { const int pos = cnt + __popc(m & ((1u << lane) - 1u)); dst[pos] = out_row[i]; src[pos] = out_cur[i]; }
      cnt += __popc(m);
    }
    if (lane == 0) s_nmov = cnt;
  }
  __syncthreads();
  const int nmov = s_nmov;
  if (nmov == 0) return;
  const int tx = tid & 31, ty = tid >> 5;
  const int ntiles = (ncols + LASWP_COLS - 1) / LASWP_COLS;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int col = t * LASWP_COLS + tx;
    double* a = A + (int64_t)blockIdx.y * sA + c0 + col;
    if (col < ncols)
      for (int e = ty; e < nmov; e += LASWP_THREADS / 32) sm[e * LASWP_COLS + tx] = a[(int64_t)src[e] * lda];
    __syncthreads();
    if (col < ncols)
      for (int e = ty; e < nmov; e += LASWP_THREADS / 32) a[(int64_t)dst[e] * lda] = sm[e * LASWP_COLS + tx];
    __syncthreads();
  }
}


// ---- first-generation block-column kernel: SHARED = true exchanged pivot candidates through L2 with a fence and one
// polling warp (61 % of its time in the election, profiles/r02_ncu_blockcol_fenced_summary.txt); superseded by blockcol2_kernel.
// SHARED = false: one CTA per matrix, everything stays in shared memory.
// SHARED = true : G co-scheduled CTAs per matrix exchange pivot candidates through global memory (L2).
template <bool SHARED>
__global__ void __launch_bounds__(BC_THREADS, 1) blockcol_kernel(BcArgs a) {
  extern __shared__ __align__(16) double sm[];
  constexpr int LD = BC_LD;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = blockIdx.x, mat = blockIdx.y;
  const int G = a.G, jb = a.jb;
  const int rows = a.rows_cap > 0 ? min(a.rows_cap, a.n - a.j) : a.n - a.j;
  const int r0 = min(rows, g * a.rpc), r1 = min(rows, r0 + a.rpc), nr = r1 - r0;
  double* A = a.A + (int64_t)mat * a.sA + (int64_t)a.j * a.lda + a.j;
  int* ipiv = a.ipiv + (int64_t)mat * a.n + a.j;

  double* tile = sm;                                   // [rpc][LD]
  double* U = tile + (((size_t)a.rpc * LD + 1) & ~(size_t)1);  // [IB][BC_UW], 16-byte aligned for the double2 loads
  double* prow = U + IB * BC_UW;                       // [NB]
  double* red_val = prow + NB;                         // [16]
  int* red_idx = reinterpret_cast<int*>(red_val + 16); // [16]
  __shared__ int s_wg, s_wr;

  char* sc = a.scratch + (size_t)mat * a.scratch_stride;
  BcCand* cands = reinterpret_cast<BcCand*>(sc);                                       // [2][Gcap]
  double* diag = reinterpret_cast<double*>(sc + (size_t)2 * a.Gcap * sizeof(BcCand));  // [2][NB]
  double* u12g = diag + 2 * NB;                                                        // [IB][BC_UW]
  unsigned* u12_flag = reinterpret_cast<unsigned*>(u12g + IB * BC_UW);

  for (int idx = tid; idx < nr * jb; idx += BC_THREADS) {
    const int r = idx / jb, c = idx - r * jb;
    tile[r * LD + c] = A[(int64_t)(r0 + r) * a.lda + c];
  }
  __syncthreads();

  for (int c = 0; c < jb; ++c) {
    const int c0 = (c / IB) * IB, pe = min(c0 + IB, jb);
    // ---- local arg-max of |a[r][c]| over rows >= c (lowest row wins ties) ----
    double best = -1.0;
    int bidx = 0x7fffffff;
    for (int r = tid; r < nr; r += BC_THREADS) {
      if (r0 + r >= c) {
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

    int p;  // block-column-relative pivot row
    if (!SHARED) {
      p = (best >= 0.0) ? bidx : c;
      if (tid < jb) {  // each thread swaps its own column: no cross-thread hazard
        const double xp = tile[p * LD + tid], xc = tile[c * LD + tid];
        tile[p * LD + tid] = xc;
        tile[c * LD + tid] = xp;
        prow[tid] = xp;
      }
      __syncthreads();
    } else {
      const unsigned epoch = (unsigned)(a.j + c + 1);
      BcCand* mine = cands + (size_t)(c & 1) * a.Gcap + g;
      double* dg = diag + (c & 1) * NB;
      if (tid < jb) {
        if (best >= 0.0) mine->content[tid] = tile[bidx * LD + tid];
      } else if (tid >= NB && tid < NB + jb) {
        if (c >= r0 && c < r1) dg[tid - NB] = tile[(c - r0) * LD + tid - NB];
      }
      __syncthreads();
      if (tid == 0) {
        fence_acq_rel_gpu();  // the CTA's content / diag stores (ordered before by the barrier) become visible first
        st_header(mine, best, (best >= 0.0) ? r0 + bidx : -1, epoch);
      }
      if (warp == 0) {
        // every CTA elects the same winner: largest value, lowest row on ties.  All of a lane's headers are
        // requested before the first one is examined; lanes spin only on the ones still carrying an old epoch.
        double wv = -1.0; int wg = 0, wr = 0x7fffffff;
        constexpr int KMAX = (BC_MAX_G + 31) / 32;
        double hv[KMAX]; int hr[KMAX]; unsigned hf[KMAX];
#pragma unroll
        for (int i = 0; i < KMAX; ++i) {
          const int k = lane + 32 * i;
          hf[i] = epoch; hv[i] = -1.0; hr[i] = -1;
          if (k < G) ld_header(cands + (size_t)(c & 1) * a.Gcap + k, hv[i], hr[i], hf[i]);
        }
#pragma unroll
        for (int i = 0; i < KMAX; ++i) {
          const int k = lane + 32 * i;
          if (k < G) {
            while (hf[i] != epoch) ld_header(cands + (size_t)(c & 1) * a.Gcap + k, hv[i], hr[i], hf[i]);
            if (hv[i] >= 0.0 && (hv[i] > wv || (hv[i] == wv && hr[i] < wr))) { wv = hv[i]; wg = k; wr = hr[i]; }
          }
        }
        fence_acq_rel_gpu();  // acquire side: the winners' content is read after this (and after the barrier below)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double ov = __shfl_xor_sync(0xffffffffu, wv, o);
          const int og = __shfl_xor_sync(0xffffffffu, wg, o);
          const int orr = __shfl_xor_sync(0xffffffffu, wr, o);
          if (ov > wv || (ov == wv && orr < wr)) { wv = ov; wg = og; wr = orr; }
        }
        if (lane == 0) { s_wg = wg; s_wr = (wv >= 0.0) ? wr : c; }
      }
      __syncthreads();
      p = s_wr;
      if (tid < jb) {
        const double x = __ldcg(&cands[(size_t)(c & 1) * a.Gcap + s_wg].content[tid]);
        prow[tid] = x;
        if (c >= r0 && c < r1) tile[(c - r0) * LD + tid] = x;  // the pivot row moves up to row c ...
      } else if (tid >= NB && tid < NB + jb) {
        if (p != c && p >= r0 && p < r1) tile[(p - r0) * LD + tid - NB] = __ldcg(&dg[tid - NB]);  // ... and row c takes its place
      }
      __syncthreads();
    }
    if (g == 0 && tid == 0) ipiv[c] = a.j + p;
    const double piv = prow[c];
    if (piv == 0.0) {
      if (g == 0 && tid == 0 && a.info[mat] == 0) a.info[mat] = a.j + c + 1;
    } else {
      // ---- scale the column, rank-1 update of the rest of the inner panel ----
      for (int r = tid; r < nr; r += BC_THREADS) {
        if (r0 + r > c) {
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
      if (g == 0) {
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
            if (SHARED) u12g[r * BC_UW + tid] = x[r];
          }
        }
        __syncthreads();
        if (SHARED && tid == 0) st_release_u32(u12_flag, (unsigned)(a.j + pe));  // release: fence + store
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
      // 4x4 register tiles: a warp covers 16 rows x 32 columns per pass
      const int ly = lane >> 3, lx = lane & 7;
      const int nstrips = (nr + 15) >> 4, ncp = (W + 31) >> 5;
      for (int s = warp; s < nstrips; s += BC_THREADS / 32) {
        const int rb = s * 16 + ly * 4;
        if (r0 + s * 16 + 15 < pe) continue;  // whole strip above the trailing block
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

  for (int idx = tid; idx < nr * jb; idx += BC_THREADS) {
    const int r = idx / jb, c = idx - r * jb;
    A[(int64_t)(r0 + r) * a.lda + c] = tile[r * LD + c];
  }
}


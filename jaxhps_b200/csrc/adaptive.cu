// Merges and down-pass steps of NON-UNIFORM (adaptive) trees, one node per call.
//
// Reference: merge/_adaptive_3D_DtN.py:150-347 + merge/_utils_adaptive_3D_DtN.py:179-881 (oct),
// merge/_adaptive_2D_DtN.py:160-433 + merge/_utils_adaptive_2D_DtN.py:168-584 (quad),
// down_pass/_adaptive_3D_DtN.py:132-394, down_pass/_adaptive_2D_DtN.py:87-259.
//
// Every boundary vector of every node is a sequence of leaf-face PANELS of npp = q^(d-1) Gauss
// points, so all index maps are kept at panel granularity: the host (jaxhps_b200/_adaptive_plan.py)
// compiles, per node, three small integer tables and the kernels below expand them on the fly.
//   * seg table of a child: panel P of the child's "interface-ready" operator T' is either one
//     panel of T (width 1) or a run of `group` = 2^(d-1) panels that is coarsened with L_1f4/L_4f1
//     (width group); `rev` walks the run backwards (2D interfaces are traversed in opposite
//     directions by the two children that share them).
//   * interface table: for interface panel I the two owners (child, panel of T').
//   * exterior table: for exterior panel E of the parent its owner (child, panel of T').
// With those, D, -C, B, A, h_int and h_ext are produced by ONE gather kernel straight in the
// parent's boundary order (the reference assembles region-ordered blocks and permutes T twice).
// S = D^-1(-C), g~ = D^-1(-h_int) come from the pivoted LU of lu.cu; T = A + B S from the DMMA GEMM.
#include <algorithm>

#include "common.cuh"

namespace hps {

namespace {

constexpr int MAXC = 8;

struct ChildSet {
  const double* T[MAXC];
  const double* h[MAXC];
  int ld[MAXC];
};

struct ChildOut {
  double* g[MAXC];
};

__device__ __forceinline__ int seg_index(int start, int len, int rev, int j) { return rev ? start + len - 1 - j : start + j; }

// tmp[r, P*npp + k] = T[r, seg_P(k)]                       (width 1)
//                   = sum_j T[r, seg_P(j)] * Lr[j, k]      (coarsened run; Lr is (group*npp) x npp)
// grid: (output panels, row tiles); block 256
__global__ void __launch_bounds__(256) compress_cols_kernel(const double* __restrict__ T, int n, int npp, int group,
                                                            const int* __restrict__ seg, const double* __restrict__ Lr,
                                                            double* __restrict__ out, int n_out) {
  const int P = blockIdx.x;
  const int start = seg[3 * P], width = seg[3 * P + 1], rev = seg[3 * P + 2];
  const int len = width * npp;
  const int rows_per_block = max(1, 256 / npp);
  for (int r0 = blockIdx.y * rows_per_block; r0 < n; r0 += gridDim.y * rows_per_block) {
    for (int e = threadIdx.x; e < rows_per_block * npp; e += blockDim.x) {
      const int r = r0 + e / npp, k = e % npp;
      if (r >= n) continue;
      const double* row = T + (int64_t)r * n;
      double v;
      if (width == 1) {
        v = row[seg_index(start, len, rev, k)];
      } else {
        v = 0.0;
        for (int j = 0; j < len; ++j) v += row[seg_index(start, len, rev, j)] * Lr[(int64_t)j * npp + k];
      }
      out[(int64_t)r * n_out + (int64_t)P * npp + k] = v;
    }
  }
}

// out[P*npp + k, c] = in[seg_P(k), c]                       (width 1)
//                   = sum_j Lc[k, j] * in[seg_P(j), c]      (Lc is npp x (group*npp))
// One output ROW per CTA row (grid: column tiles x output rows), so that the group*npp-long dot products of
// a coarsened panel are spread over npp times more threads than a per-panel loop would give.
__global__ void __launch_bounds__(256) compress_rows_kernel(const double* __restrict__ in, int ncols, int npp, int group,
                                                            const int* __restrict__ seg, const double* __restrict__ Lc,
                                                            double* __restrict__ out) {
  const int row = blockIdx.y;
  const int P = row / npp, k = row - P * npp;
  const int start = seg[3 * P], width = seg[3 * P + 1], rev = seg[3 * P + 2];
  const int len = width * npp;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncols; c += gridDim.x * blockDim.x) {
    double v;
    if (width == 1) {
      v = in[(int64_t)seg_index(start, len, rev, k) * ncols + c];
    } else {
      const double* lrow = Lc + (int64_t)k * len;
      double v0 = 0.0, v1 = 0.0;
      int j = 0;
      for (; j + 1 < len; j += 2) {
        v0 += lrow[j] * in[(int64_t)seg_index(start, len, rev, j) * ncols + c];
        v1 += lrow[j + 1] * in[(int64_t)seg_index(start, len, rev, j + 1) * ncols + c];
      }
      if (j < len) v0 += lrow[j] * in[(int64_t)seg_index(start, len, rev, j) * ncols + c];
      v = v0 + v1;
    }
    out[(int64_t)row * ncols + c] = v;
  }
}

// The (n_int + n_ext) x (n_int + n_ext) block system of a merge, one npp x npp tile per CTA:
//   [ D   -S ]        rows: interface panels, then exterior panels (parent order)
//   [ B    A ]
// A tile couples two panels only through a child that owns both, so the two table look-ups are done
// once per tile and the body is a plain (sum of at most two) tile copy: HBM-bound, 512-byte rows.
// Exterior COLUMNS are restricted to the window [ext0, ext0 + NEloc) panels (S has leading dimension
// NEloc*npp): the whole range for an ordinary merge, one rank's share for the column-sharded root merge.
// grid: (NI + NEloc column panels, row panels); block 256
__global__ void __launch_bounds__(256) adaptive_gather_kernel(ChildSet cs, int npp, int NI, int NE,
                                                              const int* __restrict__ int_tbl,
                                                              const int* __restrict__ ext_tbl, double* __restrict__ D,
                                                              double* __restrict__ S, double* __restrict__ T_out,
                                                              double* __restrict__ B, int ext0, int NEloc) {
  const int n_int = NI * npp, n_ext = NE * npp, n_ext_loc = NEloc * npp;
  const int Pr = blockIdx.y, Pc = blockIdx.x;
  const bool irow = Pr < NI, icol = Pc < NI;
  int c0, p0, c1 = -1, p1 = 0;
  if (irow) {
    c0 = int_tbl[4 * Pr], p0 = int_tbl[4 * Pr + 1], c1 = int_tbl[4 * Pr + 2], p1 = int_tbl[4 * Pr + 3];
  } else {
    c0 = ext_tbl[2 * (Pr - NI)], p0 = ext_tbl[2 * (Pr - NI) + 1];
  }
  int d0, q0, d1 = -1, q1 = 0;
  if (icol) {
    d0 = int_tbl[4 * Pc], q0 = int_tbl[4 * Pc + 1], d1 = int_tbl[4 * Pc + 2], q1 = int_tbl[4 * Pc + 3];
  } else {
    const int E = Pc - NI + ext0;
    d0 = ext_tbl[2 * E], q0 = ext_tbl[2 * E + 1];
  }
  // up to two source tiles: (child, row panel, col panel)
  const double* srcA = nullptr;
  const double* srcB = nullptr;
  int ldA = 0, ldB = 0;
  auto add = [&](int c, int pr, int pc) {
    const double* s = cs.T[c] + (int64_t)pr * npp * cs.ld[c] + (int64_t)pc * npp;
    if (!srcA) srcA = s, ldA = cs.ld[c];
    else srcB = s, ldB = cs.ld[c];
  };
  if (d0 == c0) add(c0, p0, q0);
  else if (d0 == c1) add(c1, p1, q0);
  if (d1 >= 0) {
    if (d1 == c0) add(c0, p0, q1);
    else if (d1 == c1) add(c1, p1, q1);
  }
  double* dst;
  int64_t ldd;
  double sign = 1.0;
  if (irow && icol) dst = D + (int64_t)Pr * npp * n_int + (int64_t)Pc * npp, ldd = n_int;
  else if (irow) dst = S + (int64_t)Pr * npp * n_ext_loc + (int64_t)(Pc - NI) * npp, ldd = n_ext_loc, sign = -1.0;
  else if (icol) {
    if (!B) return;
    dst = B + (int64_t)(Pr - NI) * npp * n_int + (int64_t)Pc * npp, ldd = n_int;
  } else dst = T_out + (int64_t)(Pr - NI) * npp * n_ext + (int64_t)(Pc - NI) * npp, ldd = n_ext;
  for (int e = threadIdx.x; e < npp * npp; e += blockDim.x) {
    const int rr = e / npp, cc = e - rr * npp;
    double v = 0.0;
    if (srcA) v = srcA[(int64_t)rr * ldA + cc];
    if (srcB) v += srcB[(int64_t)rr * ldB + cc];
    dst[(int64_t)rr * ldd + cc] = sign * v;
  }
}

// right-hand sides: g~ := -(h_A + h_B) on the interface rows, h_out := h of the owner on the exterior rows
__global__ void __launch_bounds__(256) adaptive_gather_rhs_kernel(ChildSet cs, int npp, int n_src, int NI, int NE,
                                                                  const int* __restrict__ int_tbl,
                                                                  const int* __restrict__ ext_tbl,
                                                                  double* __restrict__ gt, double* __restrict__ h_out,
                                                                  int want_T) {
  const int n_int = NI * npp;
  const int64_t total = (int64_t)(want_T ? (NI + NE) : NI) * npp * n_src;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(e / n_src), k = (int)(e - (int64_t)row * n_src);
    if (row < n_int) {
      const int P = row / npp, rr = row - P * npp;
      const int c0 = int_tbl[4 * P], p0 = int_tbl[4 * P + 1], c1 = int_tbl[4 * P + 2], p1 = int_tbl[4 * P + 3];
      gt[(int64_t)row * n_src + k] = -(cs.h[c0][(int64_t)(p0 * npp + rr) * n_src + k] + cs.h[c1][(int64_t)(p1 * npp + rr) * n_src + k]);
    } else {
      const int rl = row - n_int, P = rl / npp, rr = rl - P * npp;
      const int c0 = ext_tbl[2 * P], p0 = ext_tbl[2 * P + 1];
      h_out[(int64_t)rl * n_src + k] = cs.h[c0][(int64_t)(p0 * npp + rr) * n_src + k];
    }
  }
}

// Children's boundary data from the parent's: tbl[t] = {child, source panel, start, width, rev}.
// Source panel sp < NE is exterior panel sp of g_ext, otherwise interface panel sp - NE of g_int.
// A coarsened run is re-refined with Lr ((group*npp) x npp): out[seg(j)] = sum_i Lr[j, i] g[sp*npp + i].
__global__ void __launch_bounds__(128) adaptive_down_kernel(ChildOut out, int npp, int n_src, int NE,
                                                            const int* __restrict__ tbl, const double* __restrict__ g_ext,
                                                            const double* __restrict__ g_int,
                                                            const double* __restrict__ Lr) {
  const int t = blockIdx.x;
  const int c = tbl[5 * t], sp = tbl[5 * t + 1], start = tbl[5 * t + 2], width = tbl[5 * t + 3], rev = tbl[5 * t + 4];
  const double* src = sp < NE ? g_ext + (int64_t)sp * npp * n_src : g_int + (int64_t)(sp - NE) * npp * n_src;
  const int len = width * npp;
  double* dst = out.g[c];
  for (int e = threadIdx.x; e < len * n_src; e += blockDim.x) {
    const int j = e / n_src, k = e - j * n_src;
    double v;
    if (width == 1) {
      v = src[(int64_t)j * n_src + k];
    } else {
      v = 0.0;
      for (int i = 0; i < npp; ++i) v += Lr[(int64_t)j * npp + i] * src[(int64_t)i * n_src + k];
    }
    dst[(int64_t)seg_index(start, len, rev, j) * n_src + k] = v;
  }
}

}  // namespace

size_t adaptive_compress_ws_bytes(int n, int n_out) { return align_up((size_t)n * n_out * sizeof(double), 256); }

int adaptive_compress(cudaStream_t st, int npp, int group, int n_src, int n, const double* T, const double* h,
                      int n_out_panels, const int* seg_tbl, const double* L_refine, const double* L_coarsen,
                      double* T_out, double* h_out, void* ws, size_t ws_bytes) {
  if (npp <= 0 || group <= 0 || n_src <= 0 || n <= 0 || n_out_panels <= 0) return fail_arg(2, "non-positive size");
  const int n_out = n_out_panels * npp;
  if (ws_bytes < adaptive_compress_ws_bytes(n, n_out)) return fail_arg(14, "adaptive_compress: workspace too small");
  double* tmp = static_cast<double*>(ws);
  const int rows_per_block = std::max(1, 256 / npp);
  {
    dim3 grid(n_out_panels, std::min((n + rows_per_block - 1) / rows_per_block, 512));
    prof_begin(PROF_GATHER, st, 8.0 * (double)n * n_out);
    compress_cols_kernel<<<grid, 256, 0, st>>>(T, n, npp, group, seg_tbl, L_refine, tmp, n_out);
    prof_end(PROF_GATHER, st);
    HPS_LAUNCH_CHECK("compress_cols_kernel");
  }
  {
    if (n_out > 65535) return fail_arg(8, "adaptive_compress: more than 65535 boundary points");
    dim3 grid(std::min((n_out + 255) / 256, 256), n_out);
    prof_begin(PROF_GATHER, st, 8.0 * (double)n_out * n_out);
    compress_rows_kernel<<<grid, 256, 0, st>>>(tmp, n_out, npp, group, seg_tbl, L_coarsen, T_out);
    prof_end(PROF_GATHER, st);
    HPS_LAUNCH_CHECK("compress_rows_kernel");
  }
  {
    dim3 grid(1, n_out);
    compress_rows_kernel<<<grid, 256, 0, st>>>(h, n_src, npp, group, seg_tbl, L_coarsen, h_out);
    HPS_LAUNCH_CHECK("compress_rows_kernel(h)");
  }
  return 0;
}

// B is only materialised (dense) when the caller gives no block list
size_t merge_adaptive_ws_bytes(int n_int, int n_ext, int dense_B) {
  return align_up((size_t)n_int * n_int * sizeof(double), 256) +
         (dense_B ? align_up((size_t)n_ext * n_int * sizeof(double), 256) : 0) + lu_workspace_bytes(1, n_int);
}

int merge_adaptive(cudaStream_t st, int npp, int n_src, int n_child, const double* const* T_child,
                   const double* const* h_child, const int* ld_child, int NI, const int* int_tbl, int NE,
                   const int* ext_tbl, double* S, double* gt, double* T_out, double* h_out, int want_T, int n_blocks,
                   const int* bs_tbl, int ext_panel0, int n_ext_panels_loc, void* ws, size_t ws_bytes, int* info) {
  if (npp <= 0 || n_src <= 0 || NI <= 0 || NE <= 0) return fail_arg(2, "non-positive size");
  if (ext_panel0 < 0 || n_ext_panels_loc <= 0 || ext_panel0 + n_ext_panels_loc > NE) return fail_arg(20, "exterior window out of range");
  if (want_T && (ext_panel0 != 0 || n_ext_panels_loc != NE)) return fail_arg(20, "T needs the full exterior window");
  if (n_child <= 0 || n_child > MAXC) return fail_arg(4, "n_child must be 1..8");
  const int n_int = NI * npp, n_ext = NE * npp, n_ext_loc = n_ext_panels_loc * npp;
  Arena ar(ws, ws_bytes);
  double* D = ar.take<double>((size_t)n_int * n_int);
  const bool dense_B = want_T && (n_blocks <= 0 || !bs_tbl);
  double* B = dense_B ? ar.take<double>((size_t)n_ext * n_int) : nullptr;
  if (!D || (dense_B && !B)) return fail_arg(19, "merge_adaptive: workspace too small");
  void* lu_ws = ar.base + ar.off;
  const size_t lu_ws_bytes = ar.cap - ar.off;
  ChildSet cs = {};
  for (int c = 0; c < n_child; ++c) cs.T[c] = T_child[c], cs.h[c] = h_child[c], cs.ld[c] = ld_child[c];
  {
    const int row_panels = want_T ? NI + NE : NI;
    if (row_panels > 65535) return fail_arg(8, "merge_adaptive: more than 65535 boundary panels");
    dim3 grid(NI + n_ext_panels_loc, row_panels);
    prof_begin(PROF_GATHER, st, 8.0 * (double)row_panels * npp * (n_int + n_ext_loc));
    adaptive_gather_kernel<<<grid, 256, 0, st>>>(cs, npp, NI, NE, int_tbl, ext_tbl, D, S, T_out, B, ext_panel0,
                                                 n_ext_panels_loc);
    prof_end(PROF_GATHER, st);
    HPS_LAUNCH_CHECK("adaptive_gather_kernel");
    const int64_t total = (int64_t)row_panels * npp * n_src;
    adaptive_gather_rhs_kernel<<<(unsigned)std::min<int64_t>((total + 255) / 256, 1024), 256, 0, st>>>(
        cs, npp, n_src, NI, NE, int_tbl, ext_tbl, gt, h_out, want_T);
    HPS_LAUNCH_CHECK("adaptive_gather_rhs_kernel");
  }
  RhsDesc rhs[2] = {{S, n_ext_loc, 0, n_ext_loc}, {gt, n_src, 0, n_src}};
  // interface systems of non-uniform merges: the coarsening operators can make rows change places locally, so the
  // speculative block columns keep partial pivoting inside each diagonal block (info = -2 if that was not enough)
  HPS_TRY(lu_solve(st, 1, n_int, D, n_int, 0, 2, rhs, lu_ws, lu_ws_bytes, info, LU_NO_PIVOT_EXPECTED | LU_PIVOT_IN_BLOCK));
  if (!want_T) return 0;
  // T = A + B S, h = h_ext + B g~
  if (dense_B) {
    HPS_TRY(dgemm(st, n_ext, n_ext, n_int, 1.0, B, n_int, 0, S, n_ext, 0, 1.0, T_out, n_ext, 0, 1));
    HPS_TRY(dgemm(st, n_ext, n_src, n_int, 1.0, B, n_int, 0, gt, n_src, 0, 1.0, h_out, n_src, 0, 1));
    return 0;
  }
  // only the non-zero blocks of B (an exterior face times an interface face of the SAME child), read in
  // place from the children's operators: bs_tbl[k] = {child, row0, col0, M, K, first row of S, first row of T}
  for (int k = 0; k < n_blocks; ++k) {
    const int* b = bs_tbl + 7 * k;
    const int c = b[0];
    if (c < 0 || c >= n_child) return fail_arg(18, "bs_tbl: child out of range");
    const double* Ablk = cs.T[c] + (int64_t)b[1] * cs.ld[c] + b[2];
    HPS_TRY(dgemm(st, b[3], n_ext, b[4], 1.0, Ablk, cs.ld[c], 0, S + (int64_t)b[5] * n_ext, n_ext, 0, 1.0,
                  T_out + (int64_t)b[6] * n_ext, n_ext, 0, 1));
    HPS_TRY(dgemm(st, b[3], n_src, b[4], 1.0, Ablk, cs.ld[c], 0, gt + (int64_t)b[5] * n_src, n_src, 0, 1.0,
                  h_out + (int64_t)b[6] * n_src, n_src, 0, 1));
  }
  return 0;
}

// Assembly half of merge_adaptive (D, S := -C over the exterior window, g~ := -h_int) for callers that
// factor D themselves: the multi-GPU root merge runs the distributed LU of lu.cu on it.
int merge_adaptive_assemble(cudaStream_t st, int npp, int n_src, int n_child, const double* const* T_child,
                            const double* const* h_child, const int* ld_child, int NI, const int* int_tbl, int NE,
                            const int* ext_tbl, double* D, double* S, double* gt, int ext_panel0, int n_ext_panels_loc) {
  if (npp <= 0 || n_src <= 0 || NI <= 0 || NE <= 0) return fail_arg(2, "non-positive size");
  if (ext_panel0 < 0 || n_ext_panels_loc <= 0 || ext_panel0 + n_ext_panels_loc > NE) return fail_arg(15, "exterior window out of range");
  if (n_child <= 0 || n_child > MAXC) return fail_arg(4, "n_child must be 1..8");
  if (NI > 65535) return fail_arg(8, "merge_adaptive: more than 65535 interface panels");
  ChildSet cs = {};
  for (int c = 0; c < n_child; ++c) cs.T[c] = T_child[c], cs.h[c] = h_child[c], cs.ld[c] = ld_child[c];
  dim3 grid(NI + n_ext_panels_loc, NI);
  prof_begin(PROF_GATHER, st, 8.0 * (double)NI * npp * npp * (NI + n_ext_panels_loc));
  adaptive_gather_kernel<<<grid, 256, 0, st>>>(cs, npp, NI, NE, int_tbl, ext_tbl, D, S, nullptr, nullptr, ext_panel0,
                                               n_ext_panels_loc);
  prof_end(PROF_GATHER, st);
  HPS_LAUNCH_CHECK("adaptive_gather_kernel");
  const int64_t total = (int64_t)NI * npp * n_src;
  adaptive_gather_rhs_kernel<<<(unsigned)std::min<int64_t>((total + 255) / 256, 1024), 256, 0, st>>>(
      cs, npp, n_src, NI, NE, int_tbl, ext_tbl, gt, nullptr, 0);
  HPS_LAUNCH_CHECK("adaptive_gather_rhs_kernel");
  return 0;
}

int down_adaptive(cudaStream_t st, int npp, int n_src, int n_int, int n_ext, const double* S, const double* g_ext,
                  const double* gt, int n_child, double* const* g_child, int n_tbl, const int* tbl,
                  const double* L_refine, void* ws) {
  if (npp <= 0 || n_src <= 0 || n_int <= 0 || n_ext <= 0 || n_tbl <= 0) return fail_arg(2, "non-positive size");
  if (n_child <= 0 || n_child > MAXC) return fail_arg(10, "n_child must be 1..8");
  double* g_int = static_cast<double*>(ws);
  // S == NULL: ws already holds g_int (multi-GPU root: the all-reduced sum of the ranks' partial products)
  if (S) HPS_TRY(dgemm_affine(st, n_int, n_src, n_ext, S, n_ext, 0, g_ext, n_src, 0, gt, n_src, 0, g_int, n_src, 0, 1));
  ChildOut out = {};
  for (int c = 0; c < n_child; ++c) out.g[c] = g_child[c];
  adaptive_down_kernel<<<n_tbl, 128, 0, st>>>(out, npp, n_src, n_ext / npp, tbl, g_ext, g_int, L_refine);
  HPS_LAUNCH_CHECK("adaptive_down_kernel");
  return 0;
}

}  // namespace hps

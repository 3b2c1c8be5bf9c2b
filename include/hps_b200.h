/*
 * hps_b200.h — C ABI of libhps_b200.so: the B200 (sm_100a) implementation of the HPS
 * build+solve hot path of meliao/jaxhps.
 *
 * The reference has no FFI of its own (it is pure Python on JAX); the seam this library
 * plugs into is the set of *stage functions* selected by `build_solver` / `solve`
 * (reference: src/jaxhps/_build_solver.py:102-114, src/jaxhps/_solve.py:77-83).  Each
 * entry point below replaces the device-side body of one of them; the Python shims in
 * jaxhps_b200/ keep the reference's signatures and call these through ctypes, and a
 * jax.ffi handler would wrap the same symbols one-to-one.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current CUDA device unless said otherwise;
 *   - all matrices are row-major (C-contiguous), FP64;
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it and
 *     performs no host synchronisation;
 *   - the library never allocates device memory: callers pass workspaces whose size comes
 *     from the matching *_workspace query;
 *   - return value: 0 = ok, <0 = argument error (-k: k-th argument), >0 = cudaError_t;
 *     hps_last_error_string() describes the last failure on the calling thread;
 *   - `info` (device int[batch]) follows LAPACK: 0 = ok, k>0 = exact zero pivot met at
 *     column k of that matrix.  It is written on the stream; the caller reads it back.
 */
#ifndef HPS_B200_H
#define HPS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- diagnostics ---------------------------------------------------------------- */
int hps_version(void);
const char* hps_last_error_string(void);

/* Optional device-side timing of the two dominant kernels (CUDA events on the launching
 * stream), used by bench.py for the roofline numbers.  hps_prof_enable(1) resets and starts
 * recording, hps_prof_read synchronises `stream` and returns, per category, the summed launch
 * time in ms, the summed work and the launch count, plus the number of kernel launches of any kind
 * the library issued since the reset.  Categories: 0 DMMA GEMM (work = flops), 1 LU panel,
 * 2 triangular-block inversion, 3 row interchanges, 4 inner 32x32 solves, 5 merge gather/scatter
 * (work = bytes read + written), 6 narrow-N mat-vec kernel (bytes read), 7 leaf assembly (bytes written),
 * 8 peer-to-peer block-column sends of the distributed factorisation (bytes sent), 9 waits for a peer's block column,
 * 10 in-place return of the S column panels from the solve order to the face order (bytes read + written).
 * ms/work/launches: HOST arrays of length HPS_PROF_NCAT = 11. */
#define HPS_PROF_NCAT 11
int hps_prof_enable(int on);
int hps_prof_read(void* stream, double* ms, double* work, int64_t* launches, int64_t* all_launches);
/* per-launch timeline of the recording (ms relative to the first record; stream_id numbers the streams in order of
 * first appearance; work = flops or bytes of the launch as in hps_prof_read; dims[4k..4k+3] = M, N, K, batch
 * of a GEMM launch, zeros otherwise); at most cap records, *n returned */
int hps_prof_timeline(double* t0_ms, double* t1_ms, double* work, int* cat, int* stream_id, int* dims, int64_t cap,
                      int64_t* n);

/* ---- dense building blocks (exported for tests, benches and the jax.ffi shim) ------
 * C[b] = alpha * A[b] (MxK) * B[b] (KxN) + beta * C[b], b < batch, element strides sX.
 * FP64 tensor-core (DMMA) kernel.  Replaces the XLA dot_general calls on the path
 * (e.g. merge/_schur_complement.py:222-235, local_solve/_uniform_2D_DtN.py:264-271). */
int hps_dgemm_strided_batched(void* stream, int M, int N, int K, double alpha,
                              const double* A, int64_t lda, int64_t sA,
                              const double* B, int64_t ldb, int64_t sB, double beta,
                              double* C, int64_t ldc, int64_t sC, int batch);

/* C[b] (K x N) = alpha * A[b]^T (A: M x K) * X[b] (M x N) + beta * C[b] for narrow N (bandwidth kernel, a thread per
 * column of A).  is_complex: complex128 interleaved, plain transpose (no conjugation); leading dimensions and strides
 * in elements.  The transposed mat-vecs of the adjoint passes (jaxhps_b200/adjoint.py: S^T g, Y^T w, Phi^T v), which
 * the reference obtains from jax.vjp through its einsum / matmul calls (examples/inverse_scattering_utils.py:110-171). */
int hps_gemv_t_strided_batched(void* stream, int M, int K, int N, double alpha,
                               const double* A, int64_t lda, int64_t sA,
                               const double* X, int64_t ldx, int64_t sX, double beta,
                               double* C, int64_t ldc, int64_t sC, int batch, int is_complex);

/* Batched LU with partial pivoting of A[b] (n x n) followed by the in-place solve
 * rhs_k[b] := A[b]^-1 rhs_k[b] for up to 4 right-hand-side matrices (n x ncols[k]).
 * A is overwritten (U in the upper triangle).  Replaces the `jnp.linalg.inv` + matmul
 * pairs of the reference (local_solve/_uniform_2D_DtN.py:257-259,
 * merge/_schur_complement.py:146,222,234). */
int hps_lu_solve_workspace(int batch, int n, size_t* bytes);
int hps_lu_solve(void* stream, int batch, int n, double* A, int64_t lda, int64_t sA,
                 int n_rhs, double* const* rhs, const int64_t* ld_rhs, const int64_t* s_rhs,
                 const int* ncols, void* ws, size_t ws_bytes, int* info);

/* ---- local_solve_stage (reference: local_solve/_uniform_3D_DtN.py:13-106,
 *      local_solve/_uniform_2D_DtN.py:9-102 + get_DtN :228-273) -------------------------
 * dim = 2 or 3.  n_c = p^dim, n_i = (p-2)^dim, n_b = n_c - n_i, n_g = 2*dim*q^(dim-1).
 * which[k] != 0 says coefficient k of the fixed order
 *   3D: [xx,xy,yy,xz,yz,zz,x,y,z,I]   2D: [xx,xy,yy,x,y,I]
 * is present; coeffs holds the present ones, [n_coef][n_leaves][n_c] in leaf ordering.
 * D1: [p][p] 1-D Chebyshev differentiation matrix already divided by the half side
 * length; P [n_b][n_g]; Q [n_g][n_c]; src [n_leaves][n_c][n_src].
 * Outputs: Y [n_leaves][n_c][n_g], T [n_leaves][n_g][n_g], v [n_leaves][n_c][n_src],
 * h [n_leaves][n_g][n_src]. */
int hps_local_solve_dtn_workspace(int dim, int n_leaves, int p, int q, int n_src, size_t* bytes);
int hps_local_solve_dtn(void* stream, int dim, int n_leaves, int p, int q, int n_src,
                        const uint8_t* which /* host */, const double* coeffs,
                        const double* D1, const double* P, const double* Q, const double* src,
                        double* Y, double* T, double* v, double* h,
                        void* ws, size_t ws_bytes, int* info);

/* ---- merge_stage, one tree level per call (reference: merge/_uniform_3D_DtN.py:127-234,
 *      merge/_schur_complement.py:293-774,117-147,182-237) -----------------------------
 * T_in [8*n_merges][6m][6m], h_in [8*n_merges][6m][n_src]; children in a..h order.
 * S [n_merges][12m][24m], g_tilde [n_merges][12m][n_src],
 * T_out [n_merges][24m][24m] and h_out [n_merges][24m][n_src] (face-ordered, i.e. already
 * permuted by the reference's get_rearrange_indices).  want_T = 0 skips T_out/h_out. */
int hps_merge_oct_dtn_level_workspace(int n_merges, int m, int n_src, size_t* bytes);
int hps_merge_oct_dtn_level(void* stream, int n_merges, int m, int n_src,
                            const double* T_in, const double* h_in,
                            double* S, double* g_tilde, double* T_out, double* h_out,
                            int want_T, void* ws, size_t ws_bytes, int* info);

/* Multi-GPU root merge, column-sharded (no reference counterpart: the reference is single-device;
 * this is the distributed form of the last `_uniform_oct_merge_DtN` call,
 * merge/_uniform_3D_DtN.py:95-117).  Every rank holds the 8 subtree-root T's (all-gathered) and
 * computes S[:, ext0:ext0+ncols] (row-major, leading dimension ncols) and the full g_tilde.
 * Workspace: hps_merge_oct_dtn_level_workspace(1, m, n_src). */
int hps_merge_oct_dtn_root_cols(void* stream, int m, int n_src, const double* T_in, const double* h_in,
                                int ext0, int ncols, double* S_cols, double* g_tilde,
                                void* ws, size_t ws_bytes, int* info);

/* Multi-GPU root merge sharded BY CHILD (what jaxhps_b200/_dist.py uses).  Columns of the root S that
 * belong to child X's exterior faces depend on T_X only, so each rank keeps its own C block and only
 * the children's interface blocks travel:
 *   hps_root_pack_oct: for the n_local subtree roots held by this rank (children child0.. of the root
 *     merge; T [n_local][6m][6m], h [n_local][6m][n_src]) extract Dblk [n_local][3m][3m] =
 *     T[int,int], Cblk [n_local][3m][3m] = T[int,ext], hblk [n_local][3m][n_src] (interior faces in
 *     ascending interface order, exterior faces in ascending face order);
 *   (all-gather Dblk and hblk over the ranks -> [8][3m][3m], [8][3m][n_src]);
 *   hps_root_solve_oct: assemble D (12m x 12m) from the 8 Dblk, solve for this rank's columns:
 *     S_r [12m][3m*n_local] (child-major, faces ascending) and g_tilde [12m][n_src]. */
int hps_root_pack_oct(void* stream, int n_local, int child0, int m, int n_src,
                      const double* T, const double* h, double* Dblk, double* Cblk, double* hblk);
int hps_root_solve_oct_workspace(int m, size_t* bytes);
int hps_root_solve_oct(void* stream, int m, int n_src, int child0, int n_local,
                       const double* Dblk_all, const double* hblk_all, const double* Cblk_loc,
                       double* S_r, double* g_tilde, void* ws, size_t ws_bytes, int* info);

/* The same two steps for an ARBITRARY set of exterior panels (one panel = one exterior face of one root child =
 * m columns of S), which lets the ranks trade panels so that the leading zero rows of -C_r (hps_root_cols_structure)
 * are spread evenly: jaxhps_b200/_dist.py assigns 24 / world panels to every rank, sorted by their child's first
 * interface.  Cpan [n_panels][3m][m]: panel k = Cblk[child][:, i*m:(i+1)*m] of child panel_child[k] (its i-th
 * exterior face).  S_r [12m][n_panels*m]; hps_root_panels_structure fills the segment description of such a set
 * (error if the panels are not sorted by first interface).  Workspace: hps_root_solve_oct_workspace(m). */
int hps_root_assemble_panels(void* stream, int m, int n_src, int n_panels, const int* panel_child,
                             const double* Dblk_all, const double* hblk_all, const double* Cpan,
                             double* D, double* S_r, double* g_tilde);
int hps_root_solve_panels(void* stream, int m, int n_src, int n_panels, const int* panel_child,
                          const double* Dblk_all, const double* hblk_all, const double* Cpan,
                          double* S_r, double* g_tilde, void* ws, size_t ws_bytes, int* info);
int hps_root_panels_structure(int n_panels, const int* panel_child, int m, int* n_seg, int* seg_cols,
                              int* seg_first_row);

/* Distributed factorisation of the root D (driven step by step from jaxhps_b200/_dist.py, which
 * issues the NCCL broadcasts in between).  Every rank holds the full n x n matrix (row-major, lda)
 * but keeps up to date only the 128-wide block columns it owns (b % world == rank).  Per block
 * column b: the owner calls hps_lu_dist_factor_pack (panel chain + pack of the column, its pivots and
 * the inverse of its unit-lower diagonal block into buf, hps_lu_dist_buffer_doubles(n) doubles); buf
 * is broadcast; the other ranks call hps_lu_dist_unpack; every rank calls hps_lu_dist_update for
 * the block columns first_block + i*block_stride (i < n_blocks, all > b) it owns; apply_left != 0 also
 * applies block b's interchanges to the columns on its left (exactly one call per block must do so;
 * the look-ahead driver updates the next block column on a second stream with apply_left = 0).  After the last
 * block every rank holds the complete P A = L U and hps_lu_dist_solve solves its own right-hand sides
 * (up to 4 matrices, n x ncols[k], in place).  One workspace for all calls: hps_lu_solve_workspace(1, n).
 * hps_root_assemble_oct is the assembly half of hps_root_solve_oct (D, S_r := -C_r, g_tilde := -h_int). */
int hps_root_assemble_oct(void* stream, int m, int n_src, int child0, int n_local,
                          const double* Dblk_all, const double* hblk_all, const double* Cblk_loc,
                          double* D, double* S_r, double* g_tilde);
int hps_lu_dist_buffer_doubles(int n, size_t* count);
int hps_lu_dist_factor_pack(void* stream, int n, double* A, int64_t lda, int b,
                            void* ws, size_t ws_bytes, int* info, double* buf);
int hps_lu_dist_unpack(void* stream, int n, double* A, int64_t lda, int b,
                       void* ws, size_t ws_bytes, const double* buf);
int hps_lu_dist_update(void* stream, int n, double* A, int64_t lda, int b,
                       int first_block, int n_blocks, int block_stride, int apply_left,
                       void* ws, size_t ws_bytes);
int hps_lu_dist_solve(void* stream, int n, double* A, int64_t lda, int n_rhs, double* const* rhs,
                      const int64_t* ld_rhs, const int* ncols, void* ws, size_t ws_bytes);

/* ---- Library-owned communicator (SURVEY §8(b) hps_comm_*) and the distributed root factorisation on it.
 * The reference is single-device (its only split is the serial subtree recomputation,
 * _subtree_recomp.py:310-390); this is the exchange step of the multi-GPU root merge.
 * One process per GPU.  Each rank owns one SYMMETRIC device segment (cudaMalloc) that every peer of the box
 * maps through CUDA IPC, so kernels store straight into the peers' HBM over NVLink / NVSwitch:
 *   hps_comm_create(rank, world <= 8)          handle bound to the current device
 *   hps_comm_reserve(comm, bytes, &changed)    grow the local segment; changed = 1 -> every rank must
 *                                              hps_comm_export (64-byte IPC handle), all-gather the handles
 *                                              (any transport; jaxhps_b200/_dist.py uses torch.distributed)
 *                                              and hps_comm_attach(comm, handles[world][64]); call
 *                                              hps_comm_detach on every rank before growing a mapped segment
 * hps_lu_dist_run: A (n x n, row-major, lda = n) assembled by every rank at hps_lu_dist_matrix_ptr (inside its
 * segment, hps_lu_dist_segment_bytes(n) bytes).  Block column b (128 wide) is factored by rank b % world and
 * stored — with its pivots and the inverse of its unit-lower diagonal block — into EVERY peer's segment by one
 * fused copy+signal kernel (release flag per block column, acquire-wait kernel on the consumer's stream); every
 * rank applies it to the block columns it owns and to its own right-hand sides rhs[k] (n x ncols[k], leading
 * dimension ld_rhs[k], up to 4), which ride along as extra trailing columns, so the forward substitution ends
 * with the factorisation; then U's diagonal inverses and the recursive backward substitution.  Everything is
 * enqueued on `stream` (+ one internal look-ahead stream); no host synchronisation.  info[0] (LAPACK
 * convention) is set by the owner of the offending block column only.  Workspace: hps_lu_solve_workspace(1, n). */
int hps_memcpy_d2d(void* stream, void* dst, const void* src, size_t bytes); /* stream-ordered copy into / out of a segment */
int hps_comm_create(int rank, int world, void** comm);
int hps_comm_destroy(void* comm);
int hps_comm_reserve(void* comm, size_t bytes, int* changed);
int hps_comm_detach(void* comm);
int hps_comm_export(void* comm, void* handle64);
int hps_comm_attach(void* comm, const void* handles);
int hps_lu_dist_segment_bytes(int n, size_t* bytes);
int hps_lu_dist_matrix_ptr(void* comm, int n, double** A);
int hps_lu_dist_run(void* comm, void* stream, int n, int n_rhs, double* const* rhs, const int64_t* ld_rhs,
                    const int* ncols, void* ws, size_t ws_bytes, int* info);
/* hps_lu_dist_run with declared structure of rhs[0] (the root's -C_r in child-major order): its columns are n_seg
 * segments of seg_cols columns, rows [0, seg_first_row[k]) of segment k exactly zero, seg_first_row non-decreasing
 * (hps_root_cols_structure fills these for a range of root children).  The forward substitution then skips the
 * leading zero rows (a third of its flops for an oct merge).  Valid only if the factorisation interchanges no rows
 * — true for the HPS merge matrices in practice, verified on the device: info = -1 means rows DID move and the
 * call must be repeated with hps_lu_dist_run on a freshly assembled matrix. */
int hps_lu_dist_run_structured(void* comm, void* stream, int n, int n_rhs, double* const* rhs, const int64_t* ld_rhs,
                               const int* ncols, int n_seg, int seg_cols, const int* seg_first_row, void* ws,
                               size_t ws_bytes, int* info);
/* Speculative block columns of the merge factorisations (default on; HPS_LU_SPEC=0 starts with them off).  The HPS
 * merge matrices D are never pivoted below their 128 x 128 diagonal blocks, so a block column is factored as
 * diagonal-block LU + one tensor-core product L21 = A21 U11^-1, with a device-side check that every multiplier is
 * <= 1 in magnitude (the condition under which partial pivoting makes the same choices).  When the check fails the
 * merge / root-solve / hps_lu_dist_run call reports info = -2: switch the speculation off and repeat the call. */
int hps_lu_set_speculative(int on);
/* seg_first_row[3 * n_local], *n_seg = 3 * n_local, *seg_cols = m for the root children child0 .. child0+n_local-1
 * (reference interface order 9..20, merge/_uniform_3D_DtN.py:238-380). */
int hps_root_cols_structure(int child0, int n_local, int m, int* n_seg, int* seg_cols, int* seg_first_row);
/* rhs[k] := A^-1 rhs[k] with the factors the last hps_lu_dist_run left in this rank's segment (local; every rank
 * holds all of them).  The "factored root" mode of the sharded solver: the root S = -D^-1 C — two thirds of the
 * build's flops at L >= 3 — is never formed, each solve applies D^-1 to -C g - h_int instead. */
int hps_lu_dist_apply(void* comm, void* stream, int n, int n_rhs, double* const* rhs, const int64_t* ld_rhs,
                      const int* ncols, void* ws, size_t ws_bytes);

/* 2D quad merge, DtN (reference: merge/_uniform_2D_DtN.py:206-348).
 * T_in [4*n_merges][4m][4m] (children SW,SE,NE,NW; sides S,E,N,W), S [n][4m][8m],
 * T_out [n][8m][8m]. */
int hps_merge_quad_dtn_level_workspace(int n_merges, int m, int n_src, size_t* bytes);
int hps_merge_quad_dtn_level(void* stream, int n_merges, int m, int n_src,
                             const double* T_in, const double* h_in,
                             double* S, double* g_tilde, double* T_out, double* h_out,
                             int want_T, void* ws, size_t ws_bytes, int* info);

/* ---- 2D ItI (impedance-to-impedance, complex128).  Complex arrays are interleaved (re, im)
 * doubles — the memory layout of numpy/torch complex128 — passed as double*.
 * Leaf (reference: local_solve/_uniform_2D_ItI.py:10-193): coeffs real [n_coef][n_leaves][p^2] in
 * the 2D order [xx,xy,yy,x,y,I]; P real [4(p-1)][4q]; G complex [4(p-1)][p^2]; QH complex [4q][p^2];
 * src complex [n_leaves][p^2][n_src].  Outputs complex: Y [n][p^2][4q], R [n][4q][4q],
 * v [n][p^2][n_src], h [n][4q][n_src].  Solved through the real embedding of the complex system.
 * coeffs_imag: imaginary parts of the coefficient fields, same layout as coeffs, or NULL when they
 * are real (complex fields: reference tests/test_accuracy/cases.py:212-265). */
int hps_local_solve_2d_iti_workspace(int n_leaves, int p, int q, int n_src, size_t* bytes);
int hps_local_solve_2d_iti(void* stream, int n_leaves, int p, int q, int n_src,
                           const uint8_t* which /* host */, const double* coeffs, const double* D1,
                           const double* P, const double* G, const double* QH, const double* src,
                           double* Y, double* R, double* v, double* h,
                           void* ws, size_t ws_bytes, int* info, const double* coeffs_imag);
/* One ItI quad-merge level (reference: merge/_uniform_2D_ItI.py:182-405,
 * merge/_schur_complement.py:6-41,78-114).  R_in [4n][4m][4m], h_in [4n][4m][n_src];
 * S [n][8m][8m] and g_tilde [n][8m][n_src] with rows in the reference's returned order
 * [a5,b5,b6,c6,c7,d7,d8,a8]; R_out [n][8m][8m], h_out [n][8m][n_src]. */
int hps_merge_quad_iti_level_workspace(int n_merges, int m, int n_src, size_t* bytes);
int hps_merge_quad_iti_level(void* stream, int n_merges, int m, int n_src,
                             const double* R_in, const double* h_in,
                             double* S, double* g_tilde, double* R_out, double* h_out,
                             int want_T, void* ws, size_t ws_bytes, int* info);
/* One ItI down-pass level (reference: down_pass/_uniform_2D_ItI.py:121-192).
 * ws: n_nodes * 48 m n_src doubles. */
int hps_down_quad_iti_level(void* stream, int n_nodes, int m, int n_src, const double* S,
                            const double* g_ext, const double* g_tilde, double* g_children, void* ws);
/* u = Y g + v, complex128 (reference: down_pass/_uniform_2D_ItI.py:107-116).
 * ws: n_leaves * 4 n_g n_src doubles. */
int hps_leaf_apply_complex(void* stream, int n_leaves, int n_c, int n_g, int n_src,
                           const double* Y, const double* g, const double* v, double* u, void* ws);

/* ---- source given at solve time (2D uniform): no-source build + upward pass ----------------
 * No-source merges additionally return D^-1 [n][n_int][n_int] and B D^-1 [n][n_ext][n_int]
 * (reference: merge/_nosource_uniform_2D_DtN.py:13-271, merge/_nosource_uniform_2D_ItI.py:19-344,
 * merge/_schur_complement.py:240-290).  Layout here: D^-1 rows/cols in this library's unknown
 * order (DtN: interfaces 5..8; ItI: [a5,b5,b6,c6,c7,d7,d8,a8]); B D^-1 rows in boundary order.  The
 * Python shims convert to the reference's stored layout.  scratch: 28 m (DtN) / 64 m (ItI)
 * doubles per merge.  Workspaces: the *_level_workspace queries with n_src = 1. */
int hps_merge_quad_dtn_level_nosource(void* stream, int n_merges, int m, const double* T_in,
                                      double* S, double* T_out, double* D_inv, double* BD_inv,
                                      double* scratch, void* ws, size_t ws_bytes, int* info);
int hps_merge_quad_iti_level_nosource(void* stream, int n_merges, int m, const double* R_in,
                                      double* S, double* R_out, double* D_inv, double* BD_inv,
                                      double* scratch, void* ws, size_t ws_bytes, int* info);
/* Up-pass gathers (reference: up_pass/_uniform_2D_DtN.py:110-173, up_pass/_uniform_2D_ItI.py:139-219):
 * children's outgoing data h_in [4n][4m][n_src] -> h_int [n][n_int][n_src], h_ext [n][8m][n_src];
 * ext_shift = 1 writes the exterior panels in the reference's pre-roll order.  ItI: block u of
 * h_int (the OTHER child's data for unknown u) lands at position pos8[u] (NULL = identity). */
int hps_up_gather_quad(void* stream, int n_nodes, int m, int n_src, const double* h_in,
                       double* h_int, double* h_ext, int ext_shift);
int hps_up_gather_quad_iti(void* stream, int n_nodes, int m, int n_src, const double* h_in,
                           double* h_int, double* h_ext, int ext_shift, const int* pos8 /* host */);
/* C[b] = alpha A[b] B[b] + beta C[b], complex128 interleaved, alpha/beta real; B contiguous K x N;
 * lda/ldc/strides in complex elements; ws: batch * 4 K N doubles.  One real DMMA GEMM on the
 * interleaved views with B expanded to its 2K x 2N real form. */
int hps_zgemm_strided_batched(void* stream, int M, int N, int K, double alpha,
                              const double* A, int64_t lda, int64_t sA, const double* B, int64_t sB,
                              double beta, double* C, int64_t ldc, int64_t sC, int batch, void* ws);

/* Dense complex solve X = A^-1 B (A n x n with leading dimension lda, B n x nrhs with ldb, X n x nrhs contiguous;
 * complex128 interleaved; A and B are not modified).  Replaces jnp.linalg.solve in the reference's ItI -> DtN
 * conversion of the top-level operator, T = -i eta (R - I)^-1 (R + I), and in its BIE coupling system
 * (examples/wave_scattering_utils.py:31-49, 96-242): real embedding + the pivoted FP64 LU.  info[0]: LAPACK convention
 * on the 2n x 2n embedded matrix. */
int hps_zgesv_workspace(int n, int nrhs, size_t* bytes);
int hps_zgesv(void* stream, int n, int nrhs, const double* A, int64_t lda, const double* B, int64_t ldb,
              double* X, void* ws, size_t ws_bytes, int* info);

/* ---- down_pass (reference: down_pass/_uniform_3D_DtN.py:116-246,
 *      down_pass/_uniform_2D_DtN.py:125-189) -------------------------------------------
 * One level: g_int = S g_ext + g_tilde, then the children's boundary vectors.
 * g_ext [n_nodes][24m][n_src] -> g_children [n_nodes][8][6m][n_src] (3D)
 * g_ext [n_nodes][8m][n_src]  -> g_children [n_nodes][4][4m][n_src] (2D).
 * ws: n_nodes * n_int * n_src doubles. */
int hps_down_oct_level(void* stream, int n_nodes, int m, int n_src, const double* S,
                       const double* g_ext, const double* g_tilde, double* g_children, void* ws);
/* Children's boundary vectors from an already-reduced g_int (multi-GPU root level: g_int is the
 * all-reduced sum of the ranks' partial products). */
int hps_down_oct_scatter(void* stream, int n_nodes, int m, int n_src, const double* g_ext,
                         const double* g_int, double* g_children);
int hps_down_quad_level(void* stream, int n_nodes, int m, int n_src, const double* S,
                        const double* g_ext, const double* g_tilde, double* g_children, void* ws);

/* Leaf evaluation u = Y g + v (reference: down_pass/_uniform_3D_DtN.py:104-111). */
int hps_leaf_apply(void* stream, int n_leaves, int n_c, int n_g, int n_src,
                   const double* Y, const double* g, const double* v, double* u);

/* ---- interpolation between regular grids and the HPS grid (reference _interpolation_methods.py:24-340,
 * Domain.interp_{from,to}_interior_points _domain.py:99-294).  bounds [n_leaves][2 dim] = xmin,xmax,ymin,ymax(,zmin,zmax);
 * cheb [p] Chebyshev-Lobatto points on [-1,1], left end first; index tables map between the natural (x slowest)
 * and the leaf's storage order; 2D stores y descending.
 *   hps_interp_from_hps: f [n_leaves][p^dim][n_src] -> out [n_pts][n_src] at arbitrary points pts [n_pts][dim]
 *     (owning leaf = first leaf in storage order whose closed box contains the point).
 *   hps_interp_to_hps: values [n_x][n_y]([n_z]) on a tensor grid (from_d sample points, w_d their inverse
 *     barycentric weights prod_{c != b}(x_b - x_c)) -> out [n_leaves][p^dim]. */
int hps_interp_from_hps(void* stream, int dim, int n_leaves, int p, int n_src, int n_pts, const double* bounds,
                        const double* cheb, const int* nat2leaf, const double* f, const double* pts, double* out);
int hps_interp_to_hps_workspace(int dim, int n_leaves, int p, int n_x, int n_y, int n_z, size_t* bytes);
int hps_interp_to_hps(void* stream, int dim, int n_leaves, int p, int n_x, int n_y, int n_z, const double* bounds,
                      const double* cheb, const double* from_x, const double* from_y, const double* from_z,
                      const double* w_x, const double* w_y, const double* w_z, const int* leaf2nat,
                      const double* values, double* out, void* ws, size_t ws_bytes);

/* ---- adaptive (non-uniform) trees, ONE NODE per call ---------------------------------------
 * reference: merge/_adaptive_3D_DtN.py:150-347 + merge/_utils_adaptive_3D_DtN.py:179-881,
 * merge/_adaptive_2D_DtN.py:160-433 + merge/_utils_adaptive_2D_DtN.py:168-584,
 * down_pass/_adaptive_3D_DtN.py:132-394, down_pass/_adaptive_2D_DtN.py:87-259.
 * Boundary vectors are sequences of leaf-face panels of npp = q^(dim-1) Gauss points; all index
 * tables are int32 DEVICE arrays at panel granularity, compiled on the host from the tree.
 * group = 2^(dim-1) fine panels face one coarse panel across a level jump.
 *
 * hps_adaptive_compress: T (n x n), h (n x n_src) of a child -> interface-ready T' (n' x n'), h',
 *   n' = n_out_panels * npp.  seg_tbl[P] = {start, width, rev}: panel P of T' is panel-run
 *   [start, start + width*npp) of T (width 1: copied; width group: rows coarsened with
 *   L_coarsen (npp x group*npp), columns with L_refine (group*npp x npp)); rev != 0 walks the run
 *   backwards.  Replaces _compress_rows_from_lst / _compress_cols_from_lst and the index gathers of
 *   get_{a..h}_submatrices / get_quadmerge_blocks_{a..d}. */
int hps_adaptive_compress_workspace(int n, int n_out_panels, int npp, size_t* bytes);
int hps_adaptive_compress(void* stream, int npp, int group, int n_src, int n, const double* T, const double* h,
                          int n_out_panels, const int* seg_tbl, const double* L_refine, const double* L_coarsen,
                          double* T_out, double* h_out, void* ws, size_t ws_bytes);
/* hps_merge_adaptive: merge the n_child (4 or 8) children of one node.  T_child/h_child/ld_child are
 * HOST arrays (device pointers to each child's interface-ready operator, its h, its order).
 * int_tbl[I] = {child A, panel in A, child B, panel in B} for interface panel I (n_int = NI*npp);
 * ext_tbl[E] = {child, panel} for exterior panel E in the parent's boundary order (n_ext = NE*npp).
 * Outputs: S (n_int x n_ext), g_tilde (n_int x n_src), and if want_T: T_out (n_ext x n_ext),
 * h_out (n_ext x n_src).  B S is formed either from a dense B (n_blocks = 0; workspace query with
 * dense_B = 1) or block by block from bs_tbl, a HOST array [n_blocks][7] = {child, row0, col0, M, K,
 * first row of S, first row of T_out} listing the non-zero blocks of B inside the children's operators
 * (an exterior face times an interface face of the same child: 72 blocks in 3D, 16 in 2D — a quarter of
 * the dense flops).  [ext_panel0, ext_panel0 + n_ext_panels_loc) selects the exterior COLUMNS of S that are
 * produced (S then has n_ext_panels_loc*npp columns): the full range normally (required when want_T), one
 * rank's share in the column-sharded root merge of the multi-GPU build.  Replaces _oct_merge / _adaptive_quad_merge_2D_DtN +
 * assemble_merge_outputs_DtN (merge/_schur_complement.py:117-237). */
int hps_merge_adaptive_workspace(int n_int, int n_ext, int dense_B, size_t* bytes);
int hps_merge_adaptive(void* stream, int npp, int n_src, int n_child, const double* const* T_child,
                       const double* const* h_child, const int* ld_child, int n_int_panels, const int* int_tbl,
                       int n_ext_panels, const int* ext_tbl, double* S, double* g_tilde, double* T_out, double* h_out,
                       int want_T, int n_blocks, const int* bs_tbl, int ext_panel0, int n_ext_panels_loc, void* ws,
                       size_t ws_bytes, int* info);
/* Assembly half of hps_merge_adaptive for callers that factor D themselves (the multi-GPU root merge runs
 * hps_lu_dist_* on it): D (n_int x n_int), S := -C over the exterior window, g_tilde := -h_int. */
int hps_merge_adaptive_assemble(void* stream, int npp, int n_src, int n_child, const double* const* T_child,
                                const double* const* h_child, const int* ld_child, int n_int_panels, const int* int_tbl,
                                int n_ext_panels, const int* ext_tbl, double* D, double* S, double* g_tilde, int ext_panel0,
                                int n_ext_panels_loc);
/* hps_down_adaptive: g_int = S g_ext + g_tilde, then every child's boundary vector.  g_child: HOST
 * array of n_child device pointers; tbl[t] = {child, source panel, start, width, rev}: the run
 * [start, start + width*npp) of that child's vector comes from source panel sp (sp < NE: exterior
 * panel of g_ext, else interface panel sp - NE of g_int), re-refined with L_refine when width > 1.
 * ws: n_int * n_src doubles.  S == NULL: ws already holds g_int (multi-GPU root level, where g_int is the
 * all-reduced sum of the ranks' partial products).  Replaces _propagate_down_oct / _propogate_down_quad. */
int hps_down_adaptive(void* stream, int npp, int n_src, int n_int, int n_ext, const double* S, const double* g_ext,
                      const double* g_tilde, int n_child, double* const* g_child, int n_tbl, const int* tbl,
                      const double* L_refine, void* ws);

/* Refinement check of the adaptive mesh generator for a whole queue of boxes (reference: the vmapped
 * check_current_discretization_global_{linf,l2}_norm, _adaptive_discretization_3D.py:114-165, 466-503;
 * _adaptive_discretization_2D.py:96-140).  f0 [n][n_c]: samples on each box's own Chebyshev cloud; f1 [n][n_f]: samples on
 * the clouds of its 2^d children (n_f = 2^d n_c); LT [n_c][n_f]: the transposed refinement operator (L_4f1 / L_8f1);
 * w [n][n_f]: quadrature weights of the children (NULL for the L_inf criterion).  Outputs per box: err_inf = max|L f0 - f1|,
 * err_l2 = sum w (L f0 - f1)^2 (0 without w), ref_max = max|f1|; the accept / split decision and the running global norm
 * stay with the caller (jaxhps_b200/_adaptive_discretization.py), as in the reference.  One DMMA GEMM + one reduction. */
int hps_refine_check_workspace(int n, int n_f, size_t* bytes);
int hps_refine_check(void* stream, int n, int n_c, int n_f, const double* f0, const double* f1, const double* LT,
                     const double* w, double* err_inf, double* err_l2, double* ref_max, void* ws, size_t ws_bytes);

#ifdef __cplusplus
}
#endif
#endif /* HPS_B200_H */

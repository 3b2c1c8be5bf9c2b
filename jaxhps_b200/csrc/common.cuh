// Shared helpers for libhps_b200 (sm_100a).  Internal header.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>
#include <unordered_map>

namespace hps {

// thread-local last error text, surfaced through hps_last_error_string()
std::string& last_error();
int fail_arg(int which, const char* what);
int fail_cuda(cudaError_t e, const char* where);

#define HPS_CUDA(call)                                            \
  do {                                                            \
    cudaError_t _e = (call);                                      \
    if (_e != cudaSuccess) return ::hps::fail_cuda(_e, #call);    \
  } while (0)

// every kernel launch site goes through this (or bumps g_launches itself)
extern std::atomic<long long> g_launches;

// Per-device library state.  Everything in it is either immutable after a once-only initialisation
// (kernel attributes, occupancy numbers) or keyed by the CALLER'S stream under a mutex, so calls on
// different streams / host threads / devices never share scratch streams or events.
struct Aux {  // look-ahead stream + events of one (device, caller stream) pair
  cudaStream_t stream = nullptr;
  cudaEvent_t panel_done[2] = {nullptr, nullptr};
  cudaEvent_t update_done[2] = {nullptr, nullptr};
  cudaEvent_t fork = nullptr, join = nullptr;
  cudaStream_t comm_stream = nullptr;  // peer-to-peer sends that are off the critical path
  cudaEvent_t factored = nullptr, sent = nullptr;
  int* host_flag = nullptr;  // pinned: "did the factorisation move rows?" read back by the structured solve
};
struct DeviceState {
  std::atomic<bool> gemm_configured{false};
  std::atomic<bool> lu_configured{false};
  std::atomic<int> coop_capacity{-1};   // co-resident CTAs of the cooperative panel kernel
  std::atomic<int> sm_count{0};
  std::mutex mu;
  std::unordered_map<cudaStream_t, Aux> aux;  // guarded by mu
};
int device_state(DeviceState*& out);            // state of the current device
int aux_for_stream(cudaStream_t st, Aux*& out);  // created on first use
#define HPS_LAUNCH_CHECK(name)                                    \
  do {                                                            \
    ++::hps::g_launches;                                          \
    cudaError_t _e = cudaGetLastError();                          \
    if (_e != cudaSuccess) return ::hps::fail_cuda(_e, name);     \
  } while (0)

// Optional per-category device timing (CUDA events on the launching stream), used by bench.py
// for the roofline of the dominant kernels.  Off by default: zero overhead on the product path.
enum ProfCat { PROF_GEMM = 0, PROF_PANEL = 1, PROF_TRTRI = 2, PROF_LASWP = 3, PROF_INNER = 4, PROF_GATHER = 5,
               PROF_SKINNY = 6, PROF_ASSEMBLE = 7, PROF_COMM = 8, PROF_WAIT = 9, PROF_UNSORT = 10, PROF_NCAT = 11 };
void prof_begin(int cat, cudaStream_t st, double work);
void prof_end(int cat, cudaStream_t st);
void prof_dims(int cat, int a, int b, int c, int d);  // optional shape note on the open record (GEMM: M, N, K, batch)

#define HPS_TRY(expr)          \
  do {                         \
    int _rc = (expr);          \
    if (_rc != 0) return _rc;  \
  } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// bump allocator over a caller-provided workspace
struct Arena {
  char* base;
  size_t cap;
  size_t off = 0;
  Arena(void* p, size_t n) : base(static_cast<char*>(p)), cap(n) {}
  template <typename T>
  T* take(size_t count) {
    size_t bytes = align_up(count * sizeof(T), 256);
    if (off + bytes > cap) return nullptr;
    T* r = reinterpret_cast<T*>(base + off);
    off += bytes;
    return r;
  }
};

// ---- internal dense primitives (gemm.cu / lu.cu) ---------------------------------
int dgemm(cudaStream_t st, int M, int N, int K, double alpha, const double* A, int64_t lda,
          int64_t sA, const double* B, int64_t ldb, int64_t sB, double beta, double* C,
          int64_t ldc, int64_t sC, int batch);

// C = alpha*A*B + beta*Cin for narrow N (< 16): bandwidth kernel, one warp per row of A.
int dgemm_skinny(cudaStream_t st, int M, int N, int K, double alpha, const double* A,
                 int64_t lda, int64_t sA, const double* B, int64_t ldb, int64_t sB, double beta,
                 const double* Cin, int64_t ldcin, int64_t sCin, double* C, int64_t ldc,
                 int64_t sC, int batch);

// C = A*B + Cin, C and Cin distinct (picks the narrow or the DMMA kernel by N)
int dgemm_affine(cudaStream_t st, int M, int N, int K, const double* A, int64_t lda, int64_t sA, const double* B,
                 int64_t ldb, int64_t sB, const double* Cin, int64_t ldcin, int64_t sCin, double* C, int64_t ldc,
                 int64_t sC, int batch);

// C[b] (K x N) = alpha A[b]^T X[b] + beta C[b] for narrow N; complex128 interleaved when is_complex (plain transpose;
// leading dimensions and strides in elements)
int gemv_t(cudaStream_t st, int M, int K, int N, double alpha, const double* A, int64_t lda, int64_t sA, const double* X,
           int64_t ldx, int64_t sX, double beta, double* C, int64_t ldc, int64_t sC, int batch, int is_complex);

// A block of right-hand-side columns.  Optional structure (n_seg > 0): the columns form n_seg segments of seg_cols
// columns each, and rows [0, seg_first_row[k]) of segment k are EXACTLY zero on entry, seg_first_row non-decreasing.
// (The -C of an HPS merge in the reference's region order: a child's exterior columns are non-zero only on that
// child's three interfaces.)  When the factorisation moved no rows, the forward substitution then never touches the
// leading zero rows: L Z = B with B[:r] = 0 gives Z[:r] = 0.  Results are bit-identical to the unstructured solve.
constexpr int RHS_MAX_SEG = 24;
struct RhsDesc {
  double* ptr;
  int64_t ld;
  int64_t stride;
  int ncols;
  int n_seg;
  int seg_cols;
  int seg_first_row[RHS_MAX_SEG];
  // columns whose rows [0, r_end) can hold a non-zero: a prefix, because seg_first_row is sorted
  int active_cols(int r_end) const {
    if (n_seg <= 0) return ncols;
    int k = 0;
    while (k < n_seg && seg_first_row[k] < r_end) ++k;
    return k == n_seg ? ncols : k * seg_cols;
  }
};
size_t lu_workspace_bytes(int batch, int n);
// flags: LU_NO_PIVOT_EXPECTED — the caller expects partial pivoting to stay inside the 128 x 128 diagonal blocks
// (HPS merge matrices): block columns are then factored without any cross-CTA pivot election and the assumption is
// verified on the device; info = -2 means it did not hold and the call must be repeated after lu_set_speculative(0).
constexpr int LU_NO_PIVOT_EXPECTED = 1;
constexpr int LU_PIVOT_IN_BLOCK = 2;
constexpr int LU_THRESHOLD_PIVOTING = 4;  // accept multipliers up to 4 below the diagonal block (threshold pivoting, u = 1/4)  // with LU_NO_PIVOT_EXPECTED: rows may still be interchanged INSIDE a diagonal block
int lu_solve(cudaStream_t st, int batch, int n, double* A, int64_t lda, int64_t sA, int n_rhs,
             const RhsDesc* rhs, void* ws, size_t ws_bytes, int* info, int flags = 0);
void lu_set_speculative(int on);

// step-wise distributed factorisation (lu.cu), driven from jaxhps_b200/_dist.py
size_t lu_dist_block_buffer_doubles(int n);
int lu_dist_factor_pack(cudaStream_t st, int n, double* A, int64_t lda, int b, void* ws, size_t ws_bytes, int* info,
                        double* buf);
int lu_dist_unpack(cudaStream_t st, int n, double* A, int64_t lda, int b, void* ws, size_t ws_bytes, const double* buf);
int lu_dist_update(cudaStream_t st, int n, double* A, int64_t lda, int b, int first_block, int n_blocks,
                   int block_stride, int apply_left, void* ws, size_t ws_bytes);
int lu_dist_solve(cudaStream_t st, int n, double* A, int64_t lda, int n_rhs, const RhsDesc* rhs, void* ws,
                  size_t ws_bytes);

// library-owned P2P communicator + the distributed factorisation that runs on it (lu.cu)
struct Comm;
int comm_create(int rank, int world, Comm** out);
int comm_destroy(Comm* c);
int comm_reserve(Comm* c, size_t bytes, int* changed);
int comm_detach(Comm* c);
int comm_export(Comm* c, void* handle64);
int comm_attach(Comm* c, const void* handles);
size_t lu_dist_segment_bytes(int n);
int lu_dist_matrix_ptr(Comm* c, int n, double** A);
int lu_dist_run(Comm* c, cudaStream_t st, int n, int n_rhs, const RhsDesc* rhs, void* ws, size_t ws_bytes, int* info);
int lu_dist_apply(Comm* c, cudaStream_t st, int n, int n_rhs, const RhsDesc* rhs, void* ws, size_t ws_bytes);

// ---- stages (leaf.cu / merge.cu) --------------------------------------------------------
size_t local_solve_workspace_bytes(int dim, int n_leaves, int p, int q);
int local_solve_dtn(cudaStream_t st, int dim, int n_leaves, int p, int q, int n_src, const uint8_t* which,
                    const double* coeffs, const double* D1, const double* P, const double* Q, const double* src,
                    double* Y, double* T, double* v, double* h, void* ws, size_t ws_bytes, int* info);
size_t merge_oct_ws_bytes(int n_merges, int m);
size_t merge_quad_ws_bytes(int n_merges, int m);
int merge_oct_level(cudaStream_t st, int n_merges, int m, int n_src, const double* T_in, const double* h_in, double* S,
                    double* gt, double* T_out, double* h_out, int want_T, void* ws, size_t ws_bytes, int* info);
int merge_quad_level(cudaStream_t st, int n_merges, int m, int n_src, const double* T_in, const double* h_in,
                     double* S, double* gt, double* T_out, double* h_out, int want_T, void* ws, size_t ws_bytes,
                     int* info);
int root_pack_oct(cudaStream_t st, int n_local, int child0, int m, int n_src, const double* T, const double* h,
                  double* Dblk, double* Cblk, double* hblk);
size_t root_solve_oct_ws_bytes(int m);
int root_assemble_oct(cudaStream_t st, int m, int n_src, int child0, int n_local, const double* Dblk_all,
                      const double* hblk_all, const double* Cblk_loc, double* D, double* S_r, double* gt);
void root_cols_structure(int child0, int n_local, int m, int& n_seg, int& seg_cols, int* seg_first_row);
int root_solve_oct(cudaStream_t st, int m, int n_src, int child0, int n_local, const double* Dblk_all,
                   const double* hblk_all, const double* Cblk_loc, double* S_r, double* gt, void* ws, size_t ws_bytes,
                   int* info);
int root_assemble_panels(cudaStream_t st, int m, int n_src, int n_panels, const int* panel_child, const double* Dblk_all,
                         const double* hblk_all, const double* Cpan, double* D, double* S_r, double* gt);
int root_panels_structure(int n_panels, const int* panel_child, int m, int& n_seg, int& seg_cols, int* seg_first_row);
int root_solve_panels(cudaStream_t st, int m, int n_src, int n_panels, const int* panel_child, const double* Dblk_all,
                      const double* hblk_all, const double* Cpan, double* S_r, double* gt, void* ws, size_t ws_bytes,
                      int* info);
int merge_oct_root_cols(cudaStream_t st, int m, int n_src, const double* T_in, const double* h_in, int ext0, int ncols,
                        double* S_cols, double* gt, void* ws, size_t ws_bytes, int* info);
int down_oct_scatter(cudaStream_t st, int n_nodes, int m, int n_src, const double* g_ext, const double* g_int,
                     double* g_children);
size_t local_solve_iti_workspace_bytes(int n_leaves, int p, int q, int n_src);
int local_solve_iti(cudaStream_t st, int n_leaves, int p, int q, int n_src, const uint8_t* which, const double* coeffs,
                    const double* D1, const double* P, const double* G, const double* QH, const double* src, double* Y,
                    double* R, double* v, double* h, void* ws, size_t ws_bytes, int* info, const double* coeffs_imag = nullptr);
size_t merge_quad_iti_ws_bytes(int n_merges, int m, int n_src);
int merge_quad_iti_level(cudaStream_t st, int n_merges, int m, int n_src, const double* R_in, const double* h_in,
                         double* S, double* gt, double* R_out, double* h_out, int want_T, void* ws, size_t ws_bytes,
                         int* info, double* D_inv = nullptr, double* BD_inv = nullptr);
int merge_quad_level_nosource(cudaStream_t st, int n_merges, int m, const double* T_in, double* S, double* T_out,
                              double* D_inv, double* BD_inv, double* h_zero, double* scratch_h, void* ws,
                              size_t ws_bytes, int* info);
int up_gather_quad(cudaStream_t st, int n_nodes, int m, int n_src, int is_complex, const double* h_in, double* h_int,
                   double* h_ext, int ext_shift);
int up_gather_quad_iti(cudaStream_t st, int n_nodes, int m, int n_src, const double* h_in, double* h_int, double* h_ext,
                       int ext_shift, const int* pos8);
int zgemm(cudaStream_t st, int M, int N, int K, double alpha, const double* A, int64_t lda, int64_t sA, const double* B,
          int64_t sB, double beta, double* C, int64_t ldc, int64_t sC, int batch, void* ws);
size_t zgesv_workspace_bytes(int n, int nrhs);
int zgesv(cudaStream_t st, int n, int nrhs, const double* A, int64_t lda, const double* B, int64_t ldb, double* X, void* ws,
          size_t ws_bytes, int* info);
int down_quad_iti_level(cudaStream_t st, int n_nodes, int m, int n_src, const double* S, const double* g_ext,
                        const double* gt, double* g_children, void* ws);
int leaf_apply_complex(cudaStream_t st, int n_leaves, int n_c, int n_g, int n_src, const double* Y, const double* g,
                       const double* v, double* u, void* ws);
int down_oct_level(cudaStream_t st, int n_nodes, int m, int n_src, const double* S, const double* g_ext,
                   const double* gt, double* g_children, void* ws);
int down_quad_level(cudaStream_t st, int n_nodes, int m, int n_src, const double* S, const double* g_ext,
                    const double* gt, double* g_children, void* ws);

// ---- adaptive (non-uniform) trees, one node per call (adaptive.cu) ---------------------------
size_t adaptive_compress_ws_bytes(int n, int n_out);
int adaptive_compress(cudaStream_t st, int npp, int group, int n_src, int n, const double* T, const double* h,
                      int n_out_panels, const int* seg_tbl, const double* L_refine, const double* L_coarsen,
                      double* T_out, double* h_out, void* ws, size_t ws_bytes);
size_t merge_adaptive_ws_bytes(int n_int, int n_ext, int dense_B);
int merge_adaptive(cudaStream_t st, int npp, int n_src, int n_child, const double* const* T_child,
                   const double* const* h_child, const int* ld_child, int NI, const int* int_tbl, int NE,
                   const int* ext_tbl, double* S, double* gt, double* T_out, double* h_out, int want_T, int n_blocks,
                   const int* bs_tbl, int ext_panel0, int n_ext_panels_loc, void* ws, size_t ws_bytes, int* info);
int merge_adaptive_assemble(cudaStream_t st, int npp, int n_src, int n_child, const double* const* T_child,
                            const double* const* h_child, const int* ld_child, int NI, const int* int_tbl, int NE,
                            const int* ext_tbl, double* D, double* S, double* gt, int ext_panel0, int n_ext_panels_loc);
int down_adaptive(cudaStream_t st, int npp, int n_src, int n_int, int n_ext, const double* S, const double* g_ext,
                  const double* gt, int n_child, double* const* g_child, int n_tbl, const int* tbl,
                  const double* L_refine, void* ws);

// ---- interpolation regular grid <-> HPS grid (interp.cu) ----------------------------------------
int interp_from_hps(cudaStream_t st, int dim, int n_leaves, int p, int n_src, int n_pts, const double* bounds,
                    const double* cheb, const int* nat2leaf, const double* f, const double* pts, double* out);
size_t interp_to_hps_ws_bytes(int dim, int n_leaves, int p, int n_x, int n_y, int n_z);
int interp_to_hps(cudaStream_t st, int dim, int n_leaves, int p, int n_x, int n_y, int n_z, const double* bounds,
                  const double* cheb, const double* from_x, const double* from_y, const double* from_z, const double* w_x,
                  const double* w_y, const double* w_z, const int* leaf2nat, const double* values, double* out, void* ws,
                  size_t ws_bytes);

size_t refine_check_ws_bytes(int n, int n_f);
int refine_check(cudaStream_t st, int n, int n_c, int n_f, const double* f0, const double* f1, const double* LT, const double* w,
                 double* err_inf, double* err_l2, double* ref_max, void* ws, size_t ws_bytes);

}  // namespace hps

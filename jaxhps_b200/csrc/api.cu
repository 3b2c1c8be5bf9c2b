// extern "C" surface of libhps_b200.so (declared in include/hps_b200.h).
#include "../../include/hps_b200.h"

#include <vector>

#include "common.cuh"

namespace hps {
std::string& last_error() {
  static thread_local std::string s;
  return s;
}
int fail_arg(int which, const char* what) {
  last_error() = std::string("argument error: ") + what;
  return -which;
}
int fail_cuda(cudaError_t e, const char* where) {
  last_error() = std::string(cudaGetErrorName(e)) + ": " + cudaGetErrorString(e) + " at " + where;
  return static_cast<int>(e);
}
std::atomic<long long> g_launches{0};

int device_state(DeviceState*& out) {
  static DeviceState states[64];
  int dev = 0;
  HPS_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail_arg(1, "device ordinal out of range");
  out = &states[dev];
  return 0;
}

int aux_for_stream(cudaStream_t st, Aux*& out) {
  DeviceState* ds = nullptr;
  HPS_TRY(device_state(ds));
  std::lock_guard<std::mutex> lock(ds->mu);
  Aux& a = ds->aux[st];  // node-based map: the reference stays valid
  if (!a.stream) {
    int lo = 0, hi = 0;
    HPS_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    HPS_CUDA(cudaStreamCreateWithPriority(&a.stream, cudaStreamNonBlocking, hi));
    for (int i = 0; i < 2; ++i) {
      HPS_CUDA(cudaEventCreateWithFlags(&a.panel_done[i], cudaEventDisableTiming));
      HPS_CUDA(cudaEventCreateWithFlags(&a.update_done[i], cudaEventDisableTiming));
    }
    HPS_CUDA(cudaEventCreateWithFlags(&a.fork, cudaEventDisableTiming));
    HPS_CUDA(cudaEventCreateWithFlags(&a.join, cudaEventDisableTiming));
    HPS_CUDA(cudaStreamCreateWithPriority(&a.comm_stream, cudaStreamNonBlocking, hi));
    HPS_CUDA(cudaEventCreateWithFlags(&a.factored, cudaEventDisableTiming));
    HPS_CUDA(cudaEventCreateWithFlags(&a.sent, cudaEventDisableTiming));
    HPS_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&a.host_flag), sizeof(int), cudaHostAllocDefault));
  }
  out = &a;
  return 0;
}

namespace {
struct ProfState {
  bool on = false;
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  struct Rec { int cat; size_t e0, e1; double work; cudaStream_t st; int dims[4]; };
  std::vector<Rec> recs;
  size_t open_rec[PROF_NCAT] = {};
  std::mutex mu;
  cudaEvent_t take() {
    if (used == pool.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      pool.push_back(e);
    }
    return pool[used++];
  }
} g_prof;
}  // namespace

void prof_begin(int cat, cudaStream_t st, double work) {
  if (!g_prof.on) return;
  std::lock_guard<std::mutex> lock(g_prof.mu);
  const size_t i0 = g_prof.used;
  cudaEventRecord(g_prof.take(), st);
  g_prof.open_rec[cat] = g_prof.recs.size();
  g_prof.recs.push_back({cat, i0, i0, work, st, {0, 0, 0, 0}});
}
void prof_dims(int cat, int a, int b, int c, int d) {
  if (!g_prof.on) return;
  std::lock_guard<std::mutex> lock(g_prof.mu);
  int* q = g_prof.recs[g_prof.open_rec[cat]].dims;
  q[0] = a; q[1] = b; q[2] = c; q[3] = d;
}
void prof_end(int cat, cudaStream_t st) {
  if (!g_prof.on) return;
  std::lock_guard<std::mutex> lock(g_prof.mu);
  const size_t i1 = g_prof.used;
  cudaEventRecord(g_prof.take(), st);
  g_prof.recs[g_prof.open_rec[cat]].e1 = i1;
}
}  // namespace hps

using namespace hps;

extern "C" {

int hps_version(void) { return 100; }
const char* hps_last_error_string(void) { return last_error().c_str(); }

int hps_prof_enable(int on) {
  std::lock_guard<std::mutex> lock(g_prof.mu);
  g_prof.on = on != 0;
  g_prof.used = 0;
  g_prof.recs.clear();
  g_launches = 0;
  return 0;
}

int hps_prof_read(void* stream, double* ms, double* work, int64_t* launches, int64_t* all_launches) {
  HPS_CUDA(cudaDeviceSynchronize());
  (void)stream;
  std::lock_guard<std::mutex> lock(g_prof.mu);
  for (int c = 0; c < PROF_NCAT; ++c) { ms[c] = 0.0; work[c] = 0.0; launches[c] = 0; }
  for (const auto& r : g_prof.recs) {
    float t = 0.f;
    HPS_CUDA(cudaEventElapsedTime(&t, g_prof.pool[r.e0], g_prof.pool[r.e1]));
    ms[r.cat] += t;
    work[r.cat] += r.work;
    launches[r.cat] += 1;
  }
  *all_launches = g_launches;
  return 0;
}

// Timeline of the recorded launches: start / end in ms relative to the first record, category and a small integer
// naming the stream (in order of first appearance).  Developer aid for the multi-stream drivers (tools/bench_dist_lu.py).
int hps_prof_timeline(double* t0_ms, double* t1_ms, double* work, int* cat, int* stream_id, int* dims, int64_t cap,
                      int64_t* n) {
  HPS_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lock(g_prof.mu);
  std::vector<cudaStream_t> seen;
  int64_t k = 0;
  for (const auto& r : g_prof.recs) {
    if (k >= cap) break;
    float a = 0.f, b = 0.f;
    HPS_CUDA(cudaEventElapsedTime(&a, g_prof.pool[g_prof.recs[0].e0], g_prof.pool[r.e0]));
    HPS_CUDA(cudaEventElapsedTime(&b, g_prof.pool[g_prof.recs[0].e0], g_prof.pool[r.e1]));
    size_t sid = 0;
    while (sid < seen.size() && seen[sid] != r.st) ++sid;
    if (sid == seen.size()) seen.push_back(r.st);
    t0_ms[k] = a; t1_ms[k] = b; work[k] = r.work; cat[k] = r.cat; stream_id[k] = (int)sid;
    for (int q = 0; q < 4; ++q) dims[4 * k + q] = r.dims[q];
    ++k;
  }
  *n = k;
  return 0;
}

int hps_dgemm_strided_batched(void* stream, int M, int N, int K, double alpha, const double* A, int64_t lda,
                              int64_t sA, const double* B, int64_t ldb, int64_t sB, double beta, double* C,
                              int64_t ldc, int64_t sC, int batch) {
  if (M < 0 || N < 0 || K < 0) return fail_arg(2, "negative dimension");
  return dgemm(static_cast<cudaStream_t>(stream), M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, batch);
}

int hps_gemv_t_strided_batched(void* stream, int M, int K, int N, double alpha, const double* A, int64_t lda, int64_t sA,
                               const double* X, int64_t ldx, int64_t sX, double beta, double* C, int64_t ldc, int64_t sC,
                               int batch, int is_complex) {
  if (M < 0 || N < 0 || K < 0) return fail_arg(2, "negative dimension");
  return gemv_t(static_cast<cudaStream_t>(stream), M, K, N, alpha, A, lda, sA, X, ldx, sX, beta, C, ldc, sC, batch, is_complex);
}

int hps_lu_solve_workspace(int batch, int n, size_t* bytes) {
  if (!bytes) return fail_arg(3, "null output pointer");
  *bytes = lu_workspace_bytes(batch, n);
  return 0;
}

int hps_lu_solve(void* stream, int batch, int n, double* A, int64_t lda, int64_t sA, int n_rhs, double* const* rhs,
                 const int64_t* ld_rhs, const int64_t* s_rhs, const int* ncols, void* ws, size_t ws_bytes,
                 int* info) {
  if (n_rhs < 0 || n_rhs > 4) return fail_arg(7, "n_rhs must be in [0, 4]");
  RhsDesc d[4];
  for (int k = 0; k < n_rhs; ++k) d[k] = RhsDesc{rhs[k], ld_rhs[k], s_rhs[k], ncols[k]};
  return lu_solve(static_cast<cudaStream_t>(stream), batch, n, A, lda, sA, n_rhs, d, ws, ws_bytes, info);
}

int hps_local_solve_dtn_workspace(int dim, int n_leaves, int p, int q, int n_src, size_t* bytes) {
  (void)n_src;
  if (!bytes) return fail_arg(6, "null output pointer");
  if (dim != 2 && dim != 3) return fail_arg(1, "dim must be 2 or 3");
  *bytes = local_solve_workspace_bytes(dim, n_leaves, p, q);
  return 0;
}

int hps_local_solve_dtn(void* stream, int dim, int n_leaves, int p, int q, int n_src, const uint8_t* which,
                        const double* coeffs, const double* D1, const double* P, const double* Q,
                        const double* src, double* Y, double* T, double* v, double* h, void* ws, size_t ws_bytes,
                        int* info) {
  return local_solve_dtn(static_cast<cudaStream_t>(stream), dim, n_leaves, p, q, n_src, which, coeffs, D1, P, Q, src,
                         Y, T, v, h, ws, ws_bytes, info);
}

int hps_merge_oct_dtn_level_workspace(int n_merges, int m, int n_src, size_t* bytes) {
  (void)n_src;
  if (!bytes) return fail_arg(4, "null output pointer");
  *bytes = merge_oct_ws_bytes(n_merges, m);
  return 0;
}
int hps_merge_oct_dtn_level(void* stream, int n_merges, int m, int n_src, const double* T_in, const double* h_in,
                            double* S, double* g_tilde, double* T_out, double* h_out, int want_T, void* ws,
                            size_t ws_bytes, int* info) {
  return merge_oct_level(static_cast<cudaStream_t>(stream), n_merges, m, n_src, T_in, h_in, S, g_tilde, T_out, h_out,
                         want_T, ws, ws_bytes, info);
}
int hps_merge_oct_dtn_root_cols(void* stream, int m, int n_src, const double* T_in, const double* h_in, int ext0,
                                int ncols, double* S_cols, double* g_tilde, void* ws, size_t ws_bytes, int* info) {
  return merge_oct_root_cols(static_cast<cudaStream_t>(stream), m, n_src, T_in, h_in, ext0, ncols, S_cols, g_tilde, ws,
                             ws_bytes, info);
}
int hps_root_pack_oct(void* stream, int n_local, int child0, int m, int n_src, const double* T, const double* h,
                      double* Dblk, double* Cblk, double* hblk) {
  return root_pack_oct(static_cast<cudaStream_t>(stream), n_local, child0, m, n_src, T, h, Dblk, Cblk, hblk);
}
int hps_root_solve_oct_workspace(int m, size_t* bytes) {
  if (!bytes) return fail_arg(2, "null output pointer");
  *bytes = root_solve_oct_ws_bytes(m);
  return 0;
}
int hps_root_solve_oct(void* stream, int m, int n_src, int child0, int n_local, const double* Dblk_all,
                       const double* hblk_all, const double* Cblk_loc, double* S_r, double* g_tilde, void* ws,
                       size_t ws_bytes, int* info) {
  return root_solve_oct(static_cast<cudaStream_t>(stream), m, n_src, child0, n_local, Dblk_all, hblk_all, Cblk_loc, S_r,
                        g_tilde, ws, ws_bytes, info);
}
int hps_root_assemble_oct(void* stream, int m, int n_src, int child0, int n_local, const double* Dblk_all,
                          const double* hblk_all, const double* Cblk_loc, double* D, double* S_r, double* g_tilde) {
  return root_assemble_oct(static_cast<cudaStream_t>(stream), m, n_src, child0, n_local, Dblk_all, hblk_all, Cblk_loc, D,
                           S_r, g_tilde);
}
int hps_root_assemble_panels(void* stream, int m, int n_src, int n_panels, const int* panel_child, const double* Dblk_all,
                             const double* hblk_all, const double* Cpan, double* D, double* S_r, double* g_tilde) {
  return root_assemble_panels(static_cast<cudaStream_t>(stream), m, n_src, n_panels, panel_child, Dblk_all, hblk_all, Cpan,
                              D, S_r, g_tilde);
}
int hps_root_solve_panels(void* stream, int m, int n_src, int n_panels, const int* panel_child, const double* Dblk_all,
                          const double* hblk_all, const double* Cpan, double* S_r, double* g_tilde, void* ws,
                          size_t ws_bytes, int* info) {
  return root_solve_panels(static_cast<cudaStream_t>(stream), m, n_src, n_panels, panel_child, Dblk_all, hblk_all, Cpan, S_r,
                           g_tilde, ws, ws_bytes, info);
}
int hps_root_panels_structure(int n_panels, const int* panel_child, int m, int* n_seg, int* seg_cols, int* seg_first_row) {
  if (!n_seg || !seg_cols || !seg_first_row || m <= 0) return fail_arg(3, "null output / non-positive m");
  return root_panels_structure(n_panels, panel_child, m, *n_seg, *seg_cols, seg_first_row);
}
int hps_lu_dist_buffer_doubles(int n, size_t* count) {
  if (!count) return fail_arg(2, "null output pointer");
  *count = lu_dist_block_buffer_doubles(n);
  return 0;
}
int hps_lu_dist_factor_pack(void* stream, int n, double* A, int64_t lda, int b, void* ws, size_t ws_bytes, int* info,
                            double* buf) {
  return lu_dist_factor_pack(static_cast<cudaStream_t>(stream), n, A, lda, b, ws, ws_bytes, info, buf);
}
int hps_lu_dist_unpack(void* stream, int n, double* A, int64_t lda, int b, void* ws, size_t ws_bytes, const double* buf) {
  return lu_dist_unpack(static_cast<cudaStream_t>(stream), n, A, lda, b, ws, ws_bytes, buf);
}
int hps_lu_dist_update(void* stream, int n, double* A, int64_t lda, int b, int first_block, int n_blocks,
                       int block_stride, int apply_left, void* ws, size_t ws_bytes) {
  return lu_dist_update(static_cast<cudaStream_t>(stream), n, A, lda, b, first_block, n_blocks, block_stride, apply_left, ws,
                        ws_bytes);
}
int hps_lu_dist_solve(void* stream, int n, double* A, int64_t lda, int n_rhs, double* const* rhs, const int64_t* ld_rhs,
                      const int* ncols, void* ws, size_t ws_bytes) {
  if (n_rhs < 0 || n_rhs > 4) return fail_arg(5, "n_rhs must be in [0, 4]");
  RhsDesc d[4];
  for (int k = 0; k < n_rhs; ++k) d[k] = RhsDesc{rhs[k], ld_rhs[k], 0, ncols[k]};
  return lu_dist_solve(static_cast<cudaStream_t>(stream), n, A, lda, n_rhs, d, ws, ws_bytes);
}
int hps_memcpy_d2d(void* stream, void* dst, const void* src, size_t bytes) {
  HPS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
  return 0;
}
int hps_comm_create(int rank, int world, void** comm) { return comm_create(rank, world, reinterpret_cast<Comm**>(comm)); }
int hps_comm_destroy(void* comm) { return comm_destroy(static_cast<Comm*>(comm)); }
int hps_comm_reserve(void* comm, size_t bytes, int* changed) { return comm_reserve(static_cast<Comm*>(comm), bytes, changed); }
int hps_comm_detach(void* comm) { return comm_detach(static_cast<Comm*>(comm)); }
int hps_comm_export(void* comm, void* handle64) { return comm_export(static_cast<Comm*>(comm), handle64); }
int hps_comm_attach(void* comm, const void* handles) { return comm_attach(static_cast<Comm*>(comm), handles); }
int hps_lu_dist_segment_bytes(int n, size_t* bytes) {
  if (!bytes) return fail_arg(2, "null output pointer");
  *bytes = lu_dist_segment_bytes(n);
  return 0;
}
int hps_lu_dist_matrix_ptr(void* comm, int n, double** A) { return lu_dist_matrix_ptr(static_cast<Comm*>(comm), n, A); }
int hps_lu_dist_run(void* comm, void* stream, int n, int n_rhs, double* const* rhs, const int64_t* ld_rhs, const int* ncols,
                    void* ws, size_t ws_bytes, int* info) {
  if (n_rhs < 0 || n_rhs > 4) return fail_arg(4, "n_rhs must be in [0, 4]");
  RhsDesc d[4];
  for (int k = 0; k < n_rhs; ++k) d[k] = RhsDesc{rhs[k], ld_rhs[k], 0, ncols[k]};
  return lu_dist_run(static_cast<Comm*>(comm), static_cast<cudaStream_t>(stream), n, n_rhs, d, ws, ws_bytes, info);
}
int hps_lu_dist_run_structured(void* comm, void* stream, int n, int n_rhs, double* const* rhs, const int64_t* ld_rhs,
                               const int* ncols, int n_seg, int seg_cols, const int* seg_first_row, void* ws,
                               size_t ws_bytes, int* info) {
  if (n_rhs < 1 || n_rhs > 4) return fail_arg(4, "n_rhs must be in [1, 4]");
  if (n_seg < 0 || n_seg > RHS_MAX_SEG || (n_seg > 0 && (!seg_first_row || seg_cols <= 0 || (int64_t)n_seg * seg_cols != ncols[0])))
    return fail_arg(8, "segments must tile rhs[0]: n_seg * seg_cols == ncols[0], n_seg <= 24");
  RhsDesc d[4];
  for (int k = 0; k < n_rhs; ++k) d[k] = RhsDesc{rhs[k], ld_rhs[k], 0, ncols[k]};
  d[0].n_seg = n_seg;
  d[0].seg_cols = seg_cols;
  for (int k = 0; k < n_seg; ++k) {
    if (k > 0 && seg_first_row[k] < seg_first_row[k - 1]) return fail_arg(10, "seg_first_row must be non-decreasing");
    d[0].seg_first_row[k] = seg_first_row[k];
  }
  return lu_dist_run(static_cast<Comm*>(comm), static_cast<cudaStream_t>(stream), n, n_rhs, d, ws, ws_bytes, info);
}
int hps_lu_set_speculative(int on) { lu_set_speculative(on); return 0; }
int hps_root_cols_structure(int child0, int n_local, int m, int* n_seg, int* seg_cols, int* seg_first_row) {
  if (n_local <= 0 || m <= 0 || child0 < 0 || child0 + n_local > 8 || !n_seg || !seg_cols || !seg_first_row)
    return fail_arg(1, "bad child range / null output");
  root_cols_structure(child0, n_local, m, *n_seg, *seg_cols, seg_first_row);
  return 0;
}
int hps_lu_dist_apply(void* comm, void* stream, int n, int n_rhs, double* const* rhs, const int64_t* ld_rhs, const int* ncols,
                      void* ws, size_t ws_bytes) {
  if (n_rhs < 0 || n_rhs > 4) return fail_arg(4, "n_rhs must be in [0, 4]");
  RhsDesc d[4];
  for (int k = 0; k < n_rhs; ++k) d[k] = RhsDesc{rhs[k], ld_rhs[k], 0, ncols[k]};
  return lu_dist_apply(static_cast<Comm*>(comm), static_cast<cudaStream_t>(stream), n, n_rhs, d, ws, ws_bytes);
}
int hps_down_oct_scatter(void* stream, int n_nodes, int m, int n_src, const double* g_ext, const double* g_int,
                         double* g_children) {
  return down_oct_scatter(static_cast<cudaStream_t>(stream), n_nodes, m, n_src, g_ext, g_int, g_children);
}
int hps_merge_quad_dtn_level_workspace(int n_merges, int m, int n_src, size_t* bytes) {
  (void)n_src;
  if (!bytes) return fail_arg(4, "null output pointer");
  *bytes = merge_quad_ws_bytes(n_merges, m);
  return 0;
}
int hps_merge_quad_dtn_level(void* stream, int n_merges, int m, int n_src, const double* T_in, const double* h_in,
                             double* S, double* g_tilde, double* T_out, double* h_out, int want_T, void* ws,
                             size_t ws_bytes, int* info) {
  return merge_quad_level(static_cast<cudaStream_t>(stream), n_merges, m, n_src, T_in, h_in, S, g_tilde, T_out, h_out,
                          want_T, ws, ws_bytes, info);
}

int hps_down_oct_level(void* stream, int n_nodes, int m, int n_src, const double* S, const double* g_ext,
                       const double* g_tilde, double* g_children, void* ws) {
  return down_oct_level(static_cast<cudaStream_t>(stream), n_nodes, m, n_src, S, g_ext, g_tilde, g_children, ws);
}
int hps_down_quad_level(void* stream, int n_nodes, int m, int n_src, const double* S, const double* g_ext,
                        const double* g_tilde, double* g_children, void* ws) {
  return down_quad_level(static_cast<cudaStream_t>(stream), n_nodes, m, n_src, S, g_ext, g_tilde, g_children, ws);
}

int hps_local_solve_2d_iti_workspace(int n_leaves, int p, int q, int n_src, size_t* bytes) {
  if (!bytes) return fail_arg(5, "null output pointer");
  *bytes = local_solve_iti_workspace_bytes(n_leaves, p, q, n_src);
  return 0;
}
int hps_local_solve_2d_iti(void* stream, int n_leaves, int p, int q, int n_src, const uint8_t* which,
                           const double* coeffs, const double* D1, const double* P, const double* G, const double* QH,
                           const double* src, double* Y, double* R, double* v, double* h, void* ws, size_t ws_bytes,
                           int* info, const double* coeffs_imag) {
  return local_solve_iti(static_cast<cudaStream_t>(stream), n_leaves, p, q, n_src, which, coeffs, D1, P, G, QH, src, Y, R,
                         v, h, ws, ws_bytes, info, coeffs_imag);
}
int hps_merge_quad_iti_level_workspace(int n_merges, int m, int n_src, size_t* bytes) {
  if (!bytes) return fail_arg(4, "null output pointer");
  *bytes = merge_quad_iti_ws_bytes(n_merges, m, n_src);
  return 0;
}
int hps_merge_quad_iti_level(void* stream, int n_merges, int m, int n_src, const double* R_in, const double* h_in,
                             double* S, double* g_tilde, double* R_out, double* h_out, int want_T, void* ws,
                             size_t ws_bytes, int* info) {
  return merge_quad_iti_level(static_cast<cudaStream_t>(stream), n_merges, m, n_src, R_in, h_in, S, g_tilde, R_out, h_out,
                              want_T, ws, ws_bytes, info);
}
int hps_merge_quad_dtn_level_nosource(void* stream, int n_merges, int m, const double* T_in, double* S, double* T_out,
                                      double* D_inv, double* BD_inv, double* scratch, void* ws, size_t ws_bytes,
                                      int* info) {
  // scratch: n_merges * (16m + 12m) doubles; its first 16m-per-merge part must be zero (children's h)
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  HPS_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * (size_t)n_merges * 16 * m, st));
  return merge_quad_level_nosource(st, n_merges, m, T_in, S, T_out, D_inv, BD_inv, scratch,
                                   scratch + (int64_t)n_merges * 16 * m, ws, ws_bytes, info);
}
int hps_merge_quad_iti_level_nosource(void* stream, int n_merges, int m, const double* R_in, double* S, double* R_out,
                                      double* D_inv, double* BD_inv, double* scratch, void* ws, size_t ws_bytes,
                                      int* info) {
  // scratch (complex): n_merges * (16m + 8m + 8m) complex numbers = 64 m doubles per merge
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  HPS_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * (size_t)n_merges * 32 * m, st));
  double* gt = scratch + (int64_t)n_merges * 32 * m;
  double* h_out = gt + (int64_t)n_merges * 16 * m;
  return merge_quad_iti_level(st, n_merges, m, 1, R_in, scratch, S, gt, R_out, h_out, 1, ws, ws_bytes, info, D_inv, BD_inv);
}
int hps_up_gather_quad(void* stream, int n_nodes, int m, int n_src, const double* h_in, double* h_int, double* h_ext,
                       int ext_shift) {
  return up_gather_quad(static_cast<cudaStream_t>(stream), n_nodes, m, n_src, 0, h_in, h_int, h_ext, ext_shift);
}
int hps_up_gather_quad_iti(void* stream, int n_nodes, int m, int n_src, const double* h_in, double* h_int,
                           double* h_ext, int ext_shift, const int* pos8) {
  return up_gather_quad_iti(static_cast<cudaStream_t>(stream), n_nodes, m, n_src, h_in, h_int, h_ext, ext_shift, pos8);
}
int hps_zgemm_strided_batched(void* stream, int M, int N, int K, double alpha, const double* A, int64_t lda, int64_t sA,
                              const double* B, int64_t sB, double beta, double* C, int64_t ldc, int64_t sC, int batch,
                              void* ws) {
  return zgemm(static_cast<cudaStream_t>(stream), M, N, K, alpha, A, lda, sA, B, sB, beta, C, ldc, sC, batch, ws);
}
int hps_zgesv_workspace(int n, int nrhs, size_t* bytes) {
  if (!bytes) return fail_arg(3, "null output pointer");
  *bytes = zgesv_workspace_bytes(n, nrhs);
  return 0;
}
int hps_zgesv(void* stream, int n, int nrhs, const double* A, int64_t lda, const double* B, int64_t ldb, double* X, void* ws,
              size_t ws_bytes, int* info) {
  return zgesv(static_cast<cudaStream_t>(stream), n, nrhs, A, lda, B, ldb, X, ws, ws_bytes, info);
}
int hps_down_quad_iti_level(void* stream, int n_nodes, int m, int n_src, const double* S, const double* g_ext,
                            const double* g_tilde, double* g_children, void* ws) {
  return down_quad_iti_level(static_cast<cudaStream_t>(stream), n_nodes, m, n_src, S, g_ext, g_tilde, g_children, ws);
}
int hps_leaf_apply_complex(void* stream, int n_leaves, int n_c, int n_g, int n_src, const double* Y, const double* g,
                           const double* v, double* u, void* ws) {
  return leaf_apply_complex(static_cast<cudaStream_t>(stream), n_leaves, n_c, n_g, n_src, Y, g, v, u, ws);
}

int hps_leaf_apply(void* stream, int n_leaves, int n_c, int n_g, int n_src, const double* Y, const double* g,
                   const double* v, double* u) {
  if (n_leaves <= 0 || n_c <= 0 || n_g <= 0 || n_src <= 0) return fail_arg(2, "non-positive size");
  return dgemm_affine(static_cast<cudaStream_t>(stream), n_c, n_src, n_g, Y, n_g, (int64_t)n_c * n_g, g, n_src,
                      (int64_t)n_g * n_src, v, n_src, (int64_t)n_c * n_src, u, n_src, (int64_t)n_c * n_src, n_leaves);
}

int hps_interp_from_hps(void* stream, int dim, int n_leaves, int p, int n_src, int n_pts, const double* bounds,
                        const double* cheb, const int* nat2leaf, const double* f, const double* pts, double* out) {
  return interp_from_hps(static_cast<cudaStream_t>(stream), dim, n_leaves, p, n_src, n_pts, bounds, cheb, nat2leaf, f, pts, out);
}
int hps_interp_to_hps_workspace(int dim, int n_leaves, int p, int n_x, int n_y, int n_z, size_t* bytes) {
  if (!bytes) return fail_arg(7, "null output pointer");
  *bytes = interp_to_hps_ws_bytes(dim, n_leaves, p, n_x, n_y, n_z);
  return 0;
}
int hps_interp_to_hps(void* stream, int dim, int n_leaves, int p, int n_x, int n_y, int n_z, const double* bounds,
                      const double* cheb, const double* from_x, const double* from_y, const double* from_z,
                      const double* w_x, const double* w_y, const double* w_z, const int* leaf2nat, const double* values,
                      double* out, void* ws, size_t ws_bytes) {
  return interp_to_hps(static_cast<cudaStream_t>(stream), dim, n_leaves, p, n_x, n_y, n_z, bounds, cheb, from_x, from_y, from_z,
                       w_x, w_y, w_z, leaf2nat, values, out, ws, ws_bytes);
}
int hps_adaptive_compress_workspace(int n, int n_out_panels, int npp, size_t* bytes) {
  if (!bytes) return fail_arg(4, "null output pointer");
  *bytes = adaptive_compress_ws_bytes(n, n_out_panels * npp);
  return 0;
}
int hps_adaptive_compress(void* stream, int npp, int group, int n_src, int n, const double* T, const double* h,
                          int n_out_panels, const int* seg_tbl, const double* L_refine, const double* L_coarsen,
                          double* T_out, double* h_out, void* ws, size_t ws_bytes) {
  return adaptive_compress(static_cast<cudaStream_t>(stream), npp, group, n_src, n, T, h, n_out_panels, seg_tbl, L_refine,
                           L_coarsen, T_out, h_out, ws, ws_bytes);
}
int hps_merge_adaptive_workspace(int n_int, int n_ext, int dense_B, size_t* bytes) {
  if (!bytes) return fail_arg(4, "null output pointer");
  *bytes = merge_adaptive_ws_bytes(n_int, n_ext, dense_B);
  return 0;
}
int hps_merge_adaptive(void* stream, int npp, int n_src, int n_child, const double* const* T_child,
                       const double* const* h_child, const int* ld_child, int n_int_panels, const int* int_tbl,
                       int n_ext_panels, const int* ext_tbl, double* S, double* g_tilde, double* T_out, double* h_out,
                       int want_T, int n_blocks, const int* bs_tbl, int ext_panel0, int n_ext_panels_loc, void* ws,
                       size_t ws_bytes, int* info) {
  return merge_adaptive(static_cast<cudaStream_t>(stream), npp, n_src, n_child, T_child, h_child, ld_child, n_int_panels,
                        int_tbl, n_ext_panels, ext_tbl, S, g_tilde, T_out, h_out, want_T, n_blocks, bs_tbl, ext_panel0,
                        n_ext_panels_loc, ws, ws_bytes, info);
}
int hps_merge_adaptive_assemble(void* stream, int npp, int n_src, int n_child, const double* const* T_child,
                                const double* const* h_child, const int* ld_child, int n_int_panels, const int* int_tbl,
                                int n_ext_panels, const int* ext_tbl, double* D, double* S, double* g_tilde, int ext_panel0,
                                int n_ext_panels_loc) {
  return merge_adaptive_assemble(static_cast<cudaStream_t>(stream), npp, n_src, n_child, T_child, h_child, ld_child,
                                 n_int_panels, int_tbl, n_ext_panels, ext_tbl, D, S, g_tilde, ext_panel0, n_ext_panels_loc);
}
int hps_down_adaptive(void* stream, int npp, int n_src, int n_int, int n_ext, const double* S, const double* g_ext,
                      const double* g_tilde, int n_child, double* const* g_child, int n_tbl, const int* tbl,
                      const double* L_refine, void* ws) {
  return down_adaptive(static_cast<cudaStream_t>(stream), npp, n_src, n_int, n_ext, S, g_ext, g_tilde, n_child, g_child,
                       n_tbl, tbl, L_refine, ws);
}

int hps_refine_check_workspace(int n, int n_f, size_t* bytes) {
  if (!bytes) return fail_arg(3, "null output pointer");
  *bytes = refine_check_ws_bytes(n, n_f);
  return 0;
}
int hps_refine_check(void* stream, int n, int n_c, int n_f, const double* f0, const double* f1, const double* LT,
                     const double* w, double* err_inf, double* err_l2, double* ref_max, void* ws, size_t ws_bytes) {
  return refine_check(static_cast<cudaStream_t>(stream), n, n_c, n_f, f0, f1, LT, w, err_inf, err_l2, ref_max, ws, ws_bytes);
}

}  // extern "C"

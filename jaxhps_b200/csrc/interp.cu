// Interpolation between regular grids and the HPS grid on the device — the step on either side of the
// build+solve path in every example of the reference.
//
// Reference: src/jaxhps/_interpolation_methods.py:24-93 (interp_from_hps_2D), :156-219 (interp_from_hps_3D),
// :278-308 (interp_to_hps_2D), :311-340 (interp_to_hps_3D); barycentric factors
// quadrature/_interpolation.py:118-326 (exact zeros of the distance table are nudged by machine epsilon).
// The tensor structure of the barycentric matrices is used directly: one 1-D factor per dimension and target
// point (from-HPS) or per dimension and leaf (to-HPS), never a Kronecker product.
#include <algorithm>
#include <cfloat>

#include "common.cuh"

namespace hps {

namespace {

constexpr int IP_MAXP = 32;

// ---- HPS grid -> arbitrary points: one warp per target point -----------------------------------------
// bounds [n_leaves][2*DIM]; cheb [p] Chebyshev-Lobatto points on [-1,1] (left end first); nat2leaf [p^DIM]:
// position, in the leaf's storage order, of natural index (i*p + j)*p + k (x slowest); f [n_leaves][p^DIM][n_src];
// pts [n_pts][DIM]; out [n_pts][n_src].  2D stores y descending (rev_y): the j-th natural node is cheb[p-1-j].
template <int DIM>
__global__ void __launch_bounds__(256) interp_from_hps_kernel(int n_leaves, int p, int n_src, int n_pts, int rev_y,
                                                              const double* __restrict__ bounds,
                                                              const double* __restrict__ cheb,
                                                              const int* __restrict__ nat2leaf,
                                                              const double* __restrict__ f,
                                                              const double* __restrict__ pts, double* __restrict__ out) {
  __shared__ double fac[8][3][IP_MAXP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t pt = (int64_t)blockIdx.x * 8 + warp;
  if (pt >= n_pts) return;
  double x[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int d = 0; d < DIM; ++d) x[d] = pts[pt * DIM + d];
  // first leaf (storage order) whose closed box contains the point; none -> leaf 0, like argmax over all-false
  int leaf = 0;
  for (int l0 = 0; l0 < n_leaves; l0 += 32) {
    const int l = l0 + lane;
    bool in = l < n_leaves;
    if (in) {
#pragma unroll
      for (int d = 0; d < DIM; ++d) in = in && x[d] >= bounds[(int64_t)l * 2 * DIM + 2 * d] && x[d] <= bounds[(int64_t)l * 2 * DIM + 2 * d + 1];
    }
    const unsigned hit = __ballot_sync(0xffffffffu, in);
    if (hit) { leaf = l0 + __ffs(hit) - 1; break; }
  }
  // barycentric row of each dimension: lane j owns node j
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
    const double lo = bounds[(int64_t)leaf * 2 * DIM + 2 * d], hi = bounds[(int64_t)leaf * 2 * DIM + 2 * d + 1];
    const bool rev = (DIM == 2 && d == 1 && rev_y);
    auto node = [&](int i) { return 0.5 * (hi - lo) * cheb[rev ? p - 1 - i : i] + 0.5 * (lo + hi); };
    double inv = 0.0;
    if (lane < p) {
      const double nj = node(lane);
      double w = 1.0;
      for (int i = 0; i < p; ++i)
        if (i != lane) w *= node(i) - nj;
      double dist = x[d] - nj;
      if (dist == 0.0) dist = DBL_EPSILON;
      inv = 1.0 / (w * dist);
    }
    double sum = inv;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane < p) fac[warp][d][lane] = inv / sum;
  }
  __syncwarp();
  const int np = (DIM == 3) ? p * p * p : p * p;
  const double* fl = f + (int64_t)leaf * np * n_src;
  for (int s = 0; s < n_src; ++s) {
    double acc = 0.0;
    for (int idx = lane; idx < np; idx += 32) {
      double wgt;
      if (DIM == 3) {
        const int k = idx % p, ij = idx / p;
        wgt = fac[warp][0][ij / p] * fac[warp][1][ij % p] * fac[warp][2][k];
      } else {
        wgt = fac[warp][0][idx / p] * fac[warp][1][idx % p];
      }
      acc = fma(wgt, fl[(int64_t)nat2leaf[idx] * n_src + s], acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[pt * n_src + s] = acc;
  }
}

// ---- regular grid -> HPS grid: per-leaf 1-D factor matrices, then GEMMs --------------------------------
// I[leaf][a][b] = 1 / (dist_ab * w_b * norm_a), dist_ab = node_a(leaf) - from_b (0 -> eps), norm_a = sum_b 1/(w_b dist_ab);
// transposed != 0 writes I^T ([leaf][b][a]).  One warp per (leaf, a).
__global__ void __launch_bounds__(256) interp_factor_kernel(int n_leaves, int p, int n_from, int dim2, int d, int rev,
                                                            int transposed, const double* __restrict__ bounds,
                                                            const double* __restrict__ cheb,
                                                            const double* __restrict__ from,
                                                            const double* __restrict__ w_inv, double* __restrict__ I) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + warp;
  if (row >= (int64_t)n_leaves * p) return;
  const int leaf = (int)(row / p), a = (int)(row - (int64_t)leaf * p);
  const double lo = bounds[(int64_t)leaf * dim2 + 2 * d], hi = bounds[(int64_t)leaf * dim2 + 2 * d + 1];
  const double node = 0.5 * (hi - lo) * cheb[rev ? p - 1 - a : a] + 0.5 * (lo + hi);
  double norm = 0.0;
  for (int b = lane; b < n_from; b += 32) {
    double dist = node - from[b];
    if (dist == 0.0) dist = DBL_EPSILON;
    norm += 1.0 / (w_inv[b] * dist);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) norm += __shfl_xor_sync(0xffffffffu, norm, o);
  for (int b = lane; b < n_from; b += 32) {
    double dist = node - from[b];
    if (dist == 0.0) dist = DBL_EPSILON;
    const double v = 1.0 / (dist * w_inv[b] * norm);
    if (transposed) I[((int64_t)leaf * n_from + b) * p + a] = v;
    else I[((int64_t)leaf * p + a) * n_from + b] = v;
  }
}

// out[leaf][q] = nat[leaf][leaf2nat[q]]
__global__ void interp_permute_kernel(int64_t total, int np, const int* __restrict__ leaf2nat, const double* __restrict__ nat,
                                      double* __restrict__ out) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t leaf = e / np;
    const int q = (int)(e - leaf * np);
    out[e] = nat[leaf * np + leaf2nat[q]];
  }
}

}  // namespace

int interp_from_hps(cudaStream_t st, int dim, int n_leaves, int p, int n_src, int n_pts, const double* bounds,
                    const double* cheb, const int* nat2leaf, const double* f, const double* pts, double* out) {
  if (dim != 2 && dim != 3) return fail_arg(2, "dim must be 2 or 3");
  if (p < 2 || p > IP_MAXP) return fail_arg(4, "p out of range [2, 32]");
  if (n_leaves <= 0 || n_src <= 0) return fail_arg(3, "non-positive size");
  if (n_pts <= 0) return 0;
  const unsigned grid = (unsigned)((n_pts + 7) / 8);
  if (dim == 3) interp_from_hps_kernel<3><<<grid, 256, 0, st>>>(n_leaves, p, n_src, n_pts, 0, bounds, cheb, nat2leaf, f, pts, out);
  else interp_from_hps_kernel<2><<<grid, 256, 0, st>>>(n_leaves, p, n_src, n_pts, 1, bounds, cheb, nat2leaf, f, pts, out);
  HPS_LAUNCH_CHECK("interp_from_hps_kernel");
  return 0;
}

size_t interp_to_hps_ws_bytes(int dim, int n_leaves, int p, int n_x, int n_y, int n_z) {
  const size_t L = n_leaves, P = p;
  size_t b = align_up(L * P * n_x * 8, 256) + align_up(L * P * n_y * 8, 256);
  if (dim == 3) {
    b += align_up(L * P * n_z * 8, 256) + align_up(L * P * (size_t)n_y * n_z * 8, 256) + align_up(L * P * P * n_z * 8, 256) +
         align_up(L * P * P * P * 8, 256);
  } else {
    b += align_up(L * P * n_y * 8, 256) + align_up(L * P * P * 8, 256);
  }
  return b + 1024;
}

// values: [n_x][n_y]([n_z]) samples (meshgrid "ij"); from_d / w_d: sample points and their inverse barycentric
// weights (prod_{c != b} (x_b - x_c)); leaf2nat [p^dim]; out [n_leaves][p^dim] in the leaf's storage order.
int interp_to_hps(cudaStream_t st, int dim, int n_leaves, int p, int n_x, int n_y, int n_z, const double* bounds,
                  const double* cheb, const double* from_x, const double* from_y, const double* from_z, const double* w_x,
                  const double* w_y, const double* w_z, const int* leaf2nat, const double* values, double* out, void* ws,
                  size_t ws_bytes) {
  if (dim != 2 && dim != 3) return fail_arg(2, "dim must be 2 or 3");
  if (p < 2 || p > IP_MAXP) return fail_arg(4, "p out of range [2, 32]");
  if (n_leaves <= 0 || n_x <= 0 || n_y <= 0 || (dim == 3 && n_z <= 0)) return fail_arg(3, "non-positive size");
  Arena ar(ws, ws_bytes);
  const size_t L = n_leaves, P = p;
  double* Ix = ar.take<double>(L * P * n_x);
  double* Iy = ar.take<double>(L * P * n_y);  // 2D: stored transposed (n_y x p) for the second product
  const unsigned fgrid = (unsigned)((L * P + 7) / 8);
  const int dim2 = 2 * dim;
  if (dim == 2) {
    double* W1 = ar.take<double>(L * P * n_y);
    double* W2 = ar.take<double>(L * P * P);
    if (!Ix || !Iy || !W1 || !W2) return fail_arg(19, "interp_to_hps: workspace too small");
    interp_factor_kernel<<<fgrid, 256, 0, st>>>(n_leaves, p, n_x, dim2, 0, 0, 0, bounds, cheb, from_x, w_x, Ix);
    interp_factor_kernel<<<fgrid, 256, 0, st>>>(n_leaves, p, n_y, dim2, 1, 1, 1, bounds, cheb, from_y, w_y, Iy);
    HPS_LAUNCH_CHECK("interp_factor_kernel");
    // W1 = Ix V (p x n_y), W2 = W1 Iy^T (p x p), natural order (x slow, y fast-descending)
    HPS_TRY(dgemm(st, p, n_y, n_x, 1.0, Ix, n_x, (int64_t)P * n_x, values, n_y, 0, 0.0, W1, n_y, (int64_t)P * n_y, n_leaves));
    HPS_TRY(dgemm(st, p, p, n_y, 1.0, W1, n_y, (int64_t)P * n_y, Iy, p, (int64_t)n_y * P, 0.0, W2, p, (int64_t)P * P, n_leaves));
    const int64_t total = (int64_t)L * P * P;
    interp_permute_kernel<<<(unsigned)std::min<int64_t>((total + 255) / 256, 4096), 256, 0, st>>>(total, p * p, leaf2nat, W2, out);
    HPS_LAUNCH_CHECK("interp_permute_kernel");
    return 0;
  }
  double* IzT = ar.take<double>(L * P * n_z);
  double* W1 = ar.take<double>(L * P * (size_t)n_y * n_z);
  double* W2 = ar.take<double>(L * P * P * n_z);
  double* W3 = ar.take<double>(L * P * P * P);
  if (!Ix || !Iy || !IzT || !W1 || !W2 || !W3) return fail_arg(19, "interp_to_hps: workspace too small");
  interp_factor_kernel<<<fgrid, 256, 0, st>>>(n_leaves, p, n_x, dim2, 0, 0, 0, bounds, cheb, from_x, w_x, Ix);
  interp_factor_kernel<<<fgrid, 256, 0, st>>>(n_leaves, p, n_y, dim2, 1, 0, 0, bounds, cheb, from_y, w_y, Iy);
  interp_factor_kernel<<<fgrid, 256, 0, st>>>(n_leaves, p, n_z, dim2, 2, 0, 1, bounds, cheb, from_z, w_z, IzT);
  HPS_LAUNCH_CHECK("interp_factor_kernel");
  const int64_t nyz = (int64_t)n_y * n_z;
  // W1[leaf] (p x n_y n_z) = Ix[leaf] V
  HPS_TRY(dgemm(st, p, (int)nyz, n_x, 1.0, Ix, n_x, (int64_t)P * n_x, values, nyz, 0, 0.0, W1, nyz, (int64_t)P * nyz, n_leaves));
  // W2[leaf][i] (p x n_z) = Iy[leaf] W1[leaf][i] (n_y x n_z), one batched product per i
  for (int i = 0; i < p; ++i)
    HPS_TRY(dgemm(st, p, n_z, n_y, 1.0, Iy, n_y, (int64_t)P * n_y, W1 + (int64_t)i * nyz, n_z, (int64_t)P * nyz, 0.0,
                  W2 + (int64_t)i * P * n_z, n_z, (int64_t)P * P * n_z, n_leaves));
  // W3[leaf] (p p x p) = W2[leaf] (p p x n_z) Iz[leaf]^T
  HPS_TRY(dgemm(st, p * p, p, n_z, 1.0, W2, n_z, (int64_t)P * P * n_z, IzT, p, (int64_t)n_z * P, 0.0, W3, p, (int64_t)P * P * P,
                n_leaves));
  const int64_t total = (int64_t)L * P * P * P;
  interp_permute_kernel<<<(unsigned)std::min<int64_t>((total + 255) / 256, 4096), 256, 0, st>>>(total, p * p * p, leaf2nat, W3, out);
  HPS_LAUNCH_CHECK("interp_permute_kernel");
  return 0;
}


// ---- refinement check of the adaptive mesh generator ---------------------------------------------------
// Reference: src/jaxhps/_adaptive_discretization_3D.py:114-165, 466-503 (vmapped check_current_discretization_global_*),
// _adaptive_discretization_2D.py:96-140.  Per queued box: interpolate the samples on the box's own Chebyshev cloud to
// the clouds of its 2^d children (one DMMA GEMM for the whole queue: [n][n_c] x [n_c][n_f]), compare with the samples
// taken there, and reduce: err_inf = max |interp - fine|, err_l2 = sum w (interp - fine)^2, ref_max = max |fine|.
namespace {
__global__ void __launch_bounds__(256) refine_reduce_kernel(int n_f, const double* __restrict__ interp,
                                                            const double* __restrict__ fine, const double* __restrict__ w,
                                                            double* __restrict__ err_inf, double* __restrict__ err_l2,
                                                            double* __restrict__ ref_max) {
  const int64_t box = blockIdx.x;
  const double* a = interp + box * n_f;
  const double* b = fine + box * n_f;
  const double* ww = w ? w + box * n_f : nullptr;
  double e_inf = 0.0, e_l2 = 0.0, r_max = 0.0;
  for (int i = threadIdx.x; i < n_f; i += blockDim.x) {
    const double f1 = b[i], d = a[i] - f1;
    e_inf = fmax(e_inf, fabs(d));
    r_max = fmax(r_max, fabs(f1));
    if (ww) e_l2 = fma(ww[i] * d, d, e_l2);
  }
  __shared__ double s0[8], s1[8], s2[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    e_inf = fmax(e_inf, __shfl_xor_sync(0xffffffffu, e_inf, o));
    r_max = fmax(r_max, __shfl_xor_sync(0xffffffffu, r_max, o));
    e_l2 += __shfl_xor_sync(0xffffffffu, e_l2, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s0[warp] = e_inf; s1[warp] = e_l2; s2[warp] = r_max; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) { e_inf = fmax(e_inf, s0[k]); e_l2 += s1[k]; r_max = fmax(r_max, s2[k]); }
    err_inf[box] = e_inf;
    err_l2[box] = e_l2;
    ref_max[box] = r_max;
  }
}
}  // namespace

size_t refine_check_ws_bytes(int n, int n_f) { return align_up((size_t)n * n_f * sizeof(double), 256); }

// f0 [n][n_c], f1 [n][n_f], LT [n_c][n_f] (the transposed refinement operator), w [n][n_f] or NULL
int refine_check(cudaStream_t st, int n, int n_c, int n_f, const double* f0, const double* f1, const double* LT, const double* w,
                 double* err_inf, double* err_l2, double* ref_max, void* ws, size_t ws_bytes) {
  if (n <= 0) return 0;
  if (n_c <= 0 || n_f <= 0) return fail_arg(3, "non-positive size");
  if (ws_bytes < refine_check_ws_bytes(n, n_f)) return fail_arg(12, "refine_check: workspace too small");
  double* interp = static_cast<double*>(ws);
  HPS_TRY(dgemm(st, n, n_f, n_c, 1.0, f0, n_c, 0, LT, n_f, 0, 0.0, interp, n_f, 0, 1));
  refine_reduce_kernel<<<n, 256, 0, st>>>(n_f, interp, f1, w, err_inf, err_l2, ref_max);
  HPS_LAUNCH_CHECK("refine_reduce_kernel");
  return 0;
}

}  // namespace hps

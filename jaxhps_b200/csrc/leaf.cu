// local_solve_stage on the device (reference: local_solve/_uniform_3D_DtN.py:13-106,
// local_solve/_uniform_2D_DtN.py:9-102,182-273).
//
// Per leaf:  A = sum_k diag(c_k) D_k  restricted to the interior rows, built directly from the
// scaled 1-D Chebyshev matrix (the p^d x p^d Kronecker operators are never materialised);
//   [Y_int | v_int] = A_ii^-1 [ -A_ie P | f_i ]   (batched LU with partial pivoting, lu.cu)
//   Y = [P ; Y_int],  v = [0 ; v_int],  T = Q Y,  h = Q v.
// The solve runs in place inside the caller's Y and v buffers.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace hps {

namespace {

constexpr int MAX_P = 32;

struct LeafGeom {
  int dim, p, n_c, n_i, n_b;
};

// natural (i,j[,k]) coordinates of leaf-ordered point `idx`
// 3D ordering: [x- face | x+ face | y- rest | y+ rest | z- rest | z+ rest | interior], natural
//   index = i*p*p + j*p + k (reference: _grid_creation_3D.py:421-458)
__device__ __forceinline__ void decode3(int idx, int p, int& i, int& j, int& k) {
  const int p2 = p * p, q = p - 2;
  if (idx < p2) { i = 0; j = idx / p; k = idx - j * p; return; }
  idx -= p2;
  if (idx < p2) { i = p - 1; j = idx / p; k = idx - j * p; return; }
  idx -= p2;
  if (idx < q * p) { j = 0; i = 1 + idx / p; k = idx % p; return; }
  idx -= q * p;
  if (idx < q * p) { j = p - 1; i = 1 + idx / p; k = idx % p; return; }
  idx -= q * p;
  if (idx < q * q) { k = 0; i = 1 + idx / q; j = 1 + idx % q; return; }
  idx -= q * q;
  if (idx < q * q) { k = p - 1; i = 1 + idx / q; j = 1 + idx % q; return; }
  idx -= q * q;
  i = 1 + idx / (q * q);
  const int rem = idx % (q * q);
  j = 1 + rem / q;
  k = 1 + rem % q;
}

// 2D ordering: 4(p-1) boundary points counter-clockwise from the SW corner (S,E,N,W), then
// the interior; natural index = i*p + j with j counting DOWN in y (j = p-1 is y_min)
// (reference: _grid_creation_2D.py:204-233, :62)
__device__ __forceinline__ void decode2(int idx, int p, int& i, int& j) {
  const int m = p - 1;
  if (idx < p) { i = idx; j = m; return; }                    // south, W -> E (both corners)
  if (idx < 2 * p - 1) { i = m; j = m - 1 - (idx - p); return; }      // east, S -> N
  if (idx < 3 * p - 3) { i = m - 1 - (idx - (2 * p - 1)); j = 0; return; }  // north, E -> W
  if (idx == 3 * p - 3) { i = 0; j = 0; return; }               // NW corner
  if (idx < 4 * m) { i = 0; j = 1 + (idx - (3 * p - 2)); return; }      // west, N -> S
  idx -= 4 * m;
  const int q = p - 2;
  i = 1 + idx / q;
  j = 1 + idx % q;
}

struct AssembleArgs {
  LeafGeom geo;
  int n_leaves, n_coef;
  int slot[10];            // slot[k] = position of coefficient k inside coeffs, or -1
  const double* coeffs;    // [n_coef][n_leaves][n_c]
  const double* D1;        // [p][p]
  double* Aii;             // [n_leaves][n_i][n_i]
  double* Aie;             // [n_leaves][n_i][n_b]
};

// A[row][b] of the 2D leaf operator (row, b in leaf ordering).
// D_x = kron(D, I); D_y = -kron(I, D) because y is stored descending; order xx, xy, yy, x, y, I
__device__ __forceinline__ double entry2d(const AssembleArgs& g, const double* D, const double* D2, int p, int n_c,
                                          int leaf, int row, int b) {
  auto coef = [&](int k) -> double {
    const int s = g.slot[k];
    return s < 0 ? 0.0 : g.coeffs[((int64_t)s * g.n_leaves + leaf) * n_c + row];
  };
  int i, j, i2, j2;
  decode2(row, p, i, j);
  decode2(b, p, i2, j2);
  const bool di = i == i2, dj = j == j2;
  double val = 0.0;
  if (dj) val = fma(coef(0), D2[i * p + i2], val);
  val = fma(coef(1), -(D[i * p + i2] * D[j * p + j2]), val);
  if (di) val = fma(coef(2), D2[j * p + j2], val);
  if (dj) val = fma(coef(3), D[i * p + i2], val);
  if (di) val = fma(coef(4), -D[j * p + j2], val);
  if (di && dj) val += coef(5);
  return val;
}

// blockIdx.y = leaf, blockIdx.x strides over interior rows; the (i,j,k) of every column is decoded
// once per CTA into shared memory, the row's once per row; structurally-zero entries (all but ~3p per row
// when no mixed derivative is present) skip the coefficient arithmetic.  Measured (ncu, round 2): ISSUE-bound, ~40
// instructions per structural zero, 0.44 of the HBM copy rate - the general path (2D, mixed derivatives);
// 3D operators without mixed derivatives go through assemble3_lines_kernel below.
template <int DIM>
__global__ void __launch_bounds__(256) assemble_kernel(AssembleArgs g) {
  __shared__ double D[MAX_P * MAX_P];
  __shared__ double D2[MAX_P * MAX_P];
  extern __shared__ __align__(4) unsigned char colidx[];  // [n_c][4]: i, j, k of every leaf-ordered column
  const int p = g.geo.p, n_c = g.geo.n_c, n_i = g.geo.n_i, n_b = g.geo.n_b;
  for (int t = threadIdx.x; t < p * p; t += blockDim.x) D[t] = g.D1[t];
  for (int b = threadIdx.x; b < n_c; b += blockDim.x) {
    int i, j, k = 0;
    if (DIM == 3) decode3(b, p, i, j, k); else decode2(b, p, i, j);
    colidx[4 * b] = (unsigned char)i; colidx[4 * b + 1] = (unsigned char)j; colidx[4 * b + 2] = (unsigned char)k;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < p * p; t += blockDim.x) {
    const int r = t / p, c = t - r * p;
    double s = 0.0;
    for (int u = 0; u < p; ++u) s = fma(D[r * p + u], D[u * p + c], s);
    D2[t] = s;
  }
  __syncthreads();
  const int leaf = blockIdx.y;
  const bool has_mixed = g.slot[1] >= 0 || g.slot[3] >= 0 || g.slot[4] >= 0;
  for (int a = blockIdx.x; a < n_i; a += gridDim.x) {
    const int row = n_b + a;
    const int i = colidx[4 * row], j = colidx[4 * row + 1], k = colidx[4 * row + 2];
    double cf[10];
#pragma unroll
    for (int q = 0; q < 10; ++q) {
      const int s = g.slot[q];
      cf[q] = s < 0 ? 0.0 : g.coeffs[((int64_t)s * g.n_leaves + leaf) * n_c + row];
    }
    double* out_ie = g.Aie + ((int64_t)leaf * n_i + a) * n_b;
    double* out_ii = g.Aii + ((int64_t)leaf * n_i + a) * n_i;
    for (int b = threadIdx.x; b < n_c; b += blockDim.x) {
      const uchar4 cidx = reinterpret_cast<const uchar4*>(colidx)[b];
      const int i2 = cidx.x, j2 = cidx.y, k2 = cidx.z;
      const bool di = i == i2, dj = j == j2, dk = k == k2;
      double val = 0.0;
      // an entry is structurally zero unless the two points share a grid line (two equal indices) or, when a
      // mixed-derivative coefficient is present, a grid plane: 98% of a 3D row takes the short path
      if (DIM == 3 && (int)di + (int)dj + (int)dk < (has_mixed ? 1 : 2)) {
        // zero
      } else if (DIM == 3) {
        // order of the reference's stack: xx, xy, yy, xz, yz, zz, x, y, z, I
        if (dj && dk) val = fma(cf[0], D2[i * p + i2], val);
        if (dk) val = fma(cf[1], D[i * p + i2] * D[j * p + j2], val);
        if (di && dk) val = fma(cf[2], D2[j * p + j2], val);
        if (dj) val = fma(cf[3], D[i * p + i2] * D[k * p + k2], val);
        if (di) val = fma(cf[4], D[j * p + j2] * D[k * p + k2], val);
        if (di && dj) val = fma(cf[5], D2[k * p + k2], val);
        if (dj && dk) val = fma(cf[6], D[i * p + i2], val);
        if (di && dk) val = fma(cf[7], D[j * p + j2], val);
        if (di && dj) val = fma(cf[8], D[k * p + k2], val);
        if (di && dj && dk) val += cf[9];
      } else {
        // D_x = kron(D, I); D_y = -kron(I, D) because y is stored descending; order xx, xy, yy, x, y, I
        if (dj) val = fma(cf[0], D2[i * p + i2], val);
        val = fma(cf[1], -(D[i * p + i2] * D[j * p + j2]), val);
        if (di) val = fma(cf[2], D2[j * p + j2], val);
        if (dj) val = fma(cf[3], D[i * p + i2], val);
        if (di) val = fma(cf[4], -D[j * p + j2], val);
        if (di && dj) val += cf[5];
      }
      if (b < n_b) out_ie[b] = val; else out_ii[b - n_b] = val;
    }
  }
}

// 3D operators without mixed derivatives: a row has only the 3p-2 non-zeros on the three grid lines through its point, and
// assemble_kernel spends ~40 instructions per ZERO it writes (issue-bound at 0.44 of HBM, profiles/r02_hbm_kernels_summary.txt).
// Here AL_ROWS rows are zero-filled with 16-byte stores, a barrier orders them before the scatter of the non-zeros, which
// uses the same fma sequences as assemble_kernel (bit-identical entries).  inv[] maps a natural (i,j,k) to its leaf-ordered column.
constexpr int AL_ROWS = 4;
__device__ __forceinline__ void zero_fill(double* ptr, int64_t n) {
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) == 0) {
    double2* p2 = reinterpret_cast<double2*>(ptr);
    const int64_t n2 = n >> 1;
    for (int64_t e = threadIdx.x; e < n2; e += blockDim.x) p2[e] = make_double2(0.0, 0.0);
    if ((n & 1) && threadIdx.x == 0) ptr[n - 1] = 0.0;
  } else {
    for (int64_t e = threadIdx.x; e < n; e += blockDim.x) ptr[e] = 0.0;
  }
}
__global__ void __launch_bounds__(256) assemble3_lines_kernel(AssembleArgs g) {
  __shared__ double D[MAX_P * MAX_P];
  __shared__ double D2[MAX_P * MAX_P];
  extern __shared__ __align__(4) unsigned char dyn_inv[];
  unsigned short* inv = reinterpret_cast<unsigned short*>(dyn_inv);  // [p^3]
  const int p = g.geo.p, n_c = g.geo.n_c, n_i = g.geo.n_i, n_b = g.geo.n_b, q = p - 2;
  for (int t = threadIdx.x; t < p * p; t += blockDim.x) D[t] = g.D1[t];
  for (int b = threadIdx.x; b < n_c; b += blockDim.x) {
    int i, j, k;
    decode3(b, p, i, j, k);
    inv[(i * p + j) * p + k] = (unsigned short)b;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < p * p; t += blockDim.x) {
    const int r = t / p, c = t - r * p;
    double s = 0.0;
    for (int u = 0; u < p; ++u) s = fma(D[r * p + u], D[u * p + c], s);
    D2[t] = s;
  }
  __syncthreads();
  const int leaf = blockIdx.y;
  const int64_t leaf_row0 = (int64_t)leaf * n_i;
  const int per_row = 3 * p;
  for (int a0 = blockIdx.x * AL_ROWS; a0 < n_i; a0 += gridDim.x * AL_ROWS) {
    const int nr = min(AL_ROWS, n_i - a0);
    zero_fill(g.Aie + (leaf_row0 + a0) * n_b, (int64_t)nr * n_b);  // the nr rows are contiguous
    zero_fill(g.Aii + (leaf_row0 + a0) * n_i, (int64_t)nr * n_i);
    __syncthreads();  // the zeros are ordered before the non-zeros that other threads store into the same rows
    for (int task = threadIdx.x; task < nr * per_row; task += blockDim.x) {
      const int r = task / per_row, rem = task - r * per_row, line = rem / p, idx = rem - line * p;
      const int a = a0 + r, row = n_b + a;
      const int i = 1 + a / (q * q), ar = a % (q * q), j = 1 + ar / q, k = 1 + ar % q;
      auto coef = [&](int c) -> double {
        const int sl = g.slot[c];
        return sl < 0 ? 0.0 : g.coeffs[((int64_t)sl * g.n_leaves + leaf) * n_c + row];
      };
      double val;
      int col;
      if (line == 0) {  // x line (and the diagonal entry)
        if (idx == i) {
          val = fma(coef(0), D2[i * p + i], 0.0);
          val = fma(coef(2), D2[j * p + j], val);
          val = fma(coef(5), D2[k * p + k], val);
          val = fma(coef(6), D[i * p + i], val);
          val = fma(coef(7), D[j * p + j], val);
          val = fma(coef(8), D[k * p + k], val);
          val += coef(9);
        } else {
          val = fma(coef(0), D2[i * p + idx], 0.0);
          val = fma(coef(6), D[i * p + idx], val);
        }
        col = inv[(idx * p + j) * p + k];
      } else if (line == 1) {  // y line
        if (idx == j) continue;
        val = fma(coef(2), D2[j * p + idx], 0.0);
        val = fma(coef(7), D[j * p + idx], val);
        col = inv[(i * p + idx) * p + k];
      } else {  // z line
        if (idx == k) continue;
        val = fma(coef(5), D2[k * p + idx], 0.0);
        val = fma(coef(8), D[k * p + idx], val);
        col = inv[(i * p + j) * p + idx];
      }
      if (col < n_b) g.Aie[(leaf_row0 + a) * n_b + col] = val;
      else g.Aii[(leaf_row0 + a) * n_i + (col - n_b)] = val;
    }
    // no barrier here: the next pass writes other rows
  }
}

// Y[:n_b] = P, v[:n_b] = 0, v[n_b:] = f[n_b:]
__global__ void leaf_init_kernel(int n_c, int n_b, int n_g, int n_src, const double* __restrict__ P,
                                 const double* __restrict__ src, double* __restrict__ Y, double* __restrict__ v) {
  const int leaf = blockIdx.y;
  const int64_t nY = (int64_t)n_b * n_g, nV = (int64_t)n_c * n_src;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nY + nV; e += (int64_t)gridDim.x * blockDim.x) {
    if (e < nY) {
      Y[(int64_t)leaf * n_c * n_g + e] = P[e];
    } else {
      const int64_t t = e - nY;
      const int64_t r = t / n_src;
      v[(int64_t)leaf * nV + t] = (r < n_b) ? 0.0 : src[(int64_t)leaf * nV + t];
    }
  }
}

int ipow(int b, int e) { int r = 1; while (e-- > 0) r *= b; return r; }

LeafGeom geom(int dim, int p) {
  LeafGeom g;
  g.dim = dim; g.p = p;
  g.n_c = ipow(p, dim);
  g.n_i = ipow(p - 2, dim);
  g.n_b = g.n_c - g.n_i;
  return g;
}


// =====================================================================================
// 2D ItI leaf (complex128), reference local_solve/_uniform_2D_ItI.py:120-186.
// The complex system  B X = R,  B = [G ; A_interior_rows],  is solved through its real
// embedding  [[Br, -Bi], [Bi, Br]] [Xr ; Xi] = [Rr ; Ri]  with the same pivoted FP64 LU as the
// DtN path: twice the flops of a native complex LU, no second copy of every kernel.
// =====================================================================================

struct ItiLeafArgs {
  AssembleArgs a;          // coefficient slots, D1 (Aii/Aie unused)
  const double2* G;        // [n_b][n_c] complex
  double* Be;              // [n_leaves][2 n_c][2 n_c]
  const double* coeffs_im; // imaginary parts of the coefficient fields (same slots), or null
  double* row_scale;       // [n_leaves][n_c] power-of-two row scales of B (output)
};

// One warp per row of B.  ROW EQUILIBRATION: the rows of B mix the impedance operator G (entries ~ p^2/h) with rows
// of the differential operator (~ p^4/h^2): three to four orders of magnitude apart at BASELINE config 2.  Partial
// pivoting on such a system picks poor pivots — LAPACK's result (the reference's `inv`, the oracle) is then 1e-12..1e-11
// from the exact solution of the FP64 inputs and a blocked solve with inverted diagonal blocks 1e-9..1e-8.  Scaling
// row r of B and of the right-hand sides by the power of two s_r = 2^-ilogb(max_c |B_rc|) is EXACT (no rounding), leaves
// the solution unchanged and brings the computed solution to ~1e-14 of the exact one (tools/lu_accuracy.py,
// tests/_longdouble.py).  The row is evaluated twice (max, then scaled stores) instead of being kept in registers.
__global__ void __launch_bounds__(256) iti_assemble_kernel(ItiLeafArgs g) {
  __shared__ double D[MAX_P * MAX_P];
  __shared__ double D2[MAX_P * MAX_P];
  const int p = g.a.geo.p, n_c = g.a.geo.n_c, n_b = g.a.geo.n_b;
  for (int t = threadIdx.x; t < p * p; t += blockDim.x) D[t] = g.a.D1[t];
  __syncthreads();
  for (int t = threadIdx.x; t < p * p; t += blockDim.x) {
    const int r = t / p, c = t - r * p;
    double s = 0.0;
    for (int u = 0; u < p; ++u) s = fma(D[r * p + u], D[u * p + c], s);
    D2[t] = s;
  }
  __syncthreads();
  const int leaf = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int64_t ld = 2 * n_c;
  double* Be = g.Be + (int64_t)leaf * ld * ld;
  AssembleArgs ai = g.a;  // A = sum_k diag(c_k) D_k with complex c_k: the imaginary part is the same sum over Im c_k
  ai.coeffs = g.coeffs_im;
  auto entry = [&](int r, int c, double& re, double& im) {
    if (r < n_b) {
      const double2 z = g.G[(int64_t)r * n_c + c];
      re = z.x; im = z.y;
    } else {
      re = entry2d(g.a, D, D2, p, n_c, leaf, r, c);
      im = g.coeffs_im ? entry2d(ai, D, D2, p, n_c, leaf, r, c) : 0.0;
    }
  };
  for (int r = blockIdx.x * nwarps + warp; r < n_c; r += gridDim.x * nwarps) {
    double m = 0.0;
    for (int c = lane; c < n_c; c += 32) {
      double re, im;
      entry(r, c, re, im);
      m = fmax(m, fmax(fabs(re), fabs(im)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    const double sc = (m > 0.0 && isfinite(m)) ? scalbn(1.0, -ilogb(m)) : 1.0;
    if (lane == 0) g.row_scale[(int64_t)leaf * n_c + r] = sc;
    double* top = Be + (int64_t)r * ld;
    double* bot = Be + (int64_t)(n_c + r) * ld;
    for (int c = lane; c < n_c; c += 32) {
      double re, im;
      entry(r, c, re, im);
      re *= sc; im *= sc;
      top[c] = re;
      top[n_c + c] = -im;
      bot[c] = im;
      bot[n_c + c] = re;
    }
  }
}

// stacked right-hand sides, rows scaled like B's: Ys = [[s P;0];[0;0]] (2n_c x n_g), vs = [[0;s Re f_i];[0;s Im f_i]] (2n_c x n_src)
__global__ void iti_rhs_kernel(int n_c, int n_b, int n_g, int n_src, const double* __restrict__ P,
                               const double2* __restrict__ src, const double* __restrict__ row_scale,
                               double* __restrict__ Ys, double* __restrict__ vs) {
  const int leaf = blockIdx.y;
  const double* sc = row_scale + (int64_t)leaf * n_c;
  const int64_t nY = (int64_t)2 * n_c * n_g, nV = (int64_t)2 * n_c * n_src;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nY + nV; e += (int64_t)gridDim.x * blockDim.x) {
    if (e < nY) {
      const int64_t r = e / n_g;
      Ys[(int64_t)leaf * nY + e] = (r < n_b) ? sc[r] * P[e] : 0.0;
    } else {
      const int64_t t = e - nY;
      const int64_t R = t / n_src, k = t - R * n_src;
      const int64_t r = R % n_c;
      double val = 0.0;
      if (r >= n_b) {
        const double2 z = src[((int64_t)leaf * n_c + r) * n_src + k];
        val = sc[r] * ((R < n_c) ? z.x : z.y);
      }
      vs[(int64_t)leaf * nV + t] = val;
    }
  }
}

}  // namespace

// stacked real [Xr ; Xi] (2 rows x cols) -> interleaved complex X (rows x cols) and, optionally, the
// expanded real form X2 (2 rows x 2 cols): row 2k = (re, im) pairs of row k, row 2k+1 = (-im, re) pairs,
// which turns a complex product A X into the REAL product  A_view (M x 2K) * X2.
__global__ void stacked_to_complex_kernel(int rows, int cols, const double* __restrict__ Xs, int64_t sXs,
                                          double2* __restrict__ X, int64_t sX, double* __restrict__ X2, int64_t sX2) {
  const int64_t b = blockIdx.y;
  const int64_t total = (int64_t)rows * cols;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / cols, c = e - r * cols;
    const double re = Xs[b * sXs + e], im = Xs[b * sXs + total + e];
    if (X) X[b * sX + e] = make_double2(re, im);
    if (X2) {
      double* x2 = X2 + b * sX2;
      x2[(2 * r) * 2 * cols + 2 * c] = re;
      x2[(2 * r) * 2 * cols + 2 * c + 1] = im;
      x2[(2 * r + 1) * 2 * cols + 2 * c] = -im;
      x2[(2 * r + 1) * 2 * cols + 2 * c + 1] = re;
    }
  }
}

// interleaved complex X (rows x cols) -> expanded real X2 (2 rows x 2 cols)
__global__ void complex_expand_kernel(int rows, int cols, const double2* __restrict__ X, int64_t sX,
                                      double* __restrict__ X2, int64_t sX2, double alpha) {
  const int64_t b = blockIdx.y;
  const int64_t total = (int64_t)rows * cols;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / cols, c = e - r * cols;
    double2 z = X[b * sX + e];
    z.x *= alpha; z.y *= alpha;
    double* x2 = X2 + b * sX2;
    x2[(2 * r) * 2 * cols + 2 * c] = z.x;
    x2[(2 * r) * 2 * cols + 2 * c + 1] = z.y;
    x2[(2 * r + 1) * 2 * cols + 2 * c] = -z.y;
    x2[(2 * r + 1) * 2 * cols + 2 * c + 1] = z.x;
  }
}

int stacked_to_complex(cudaStream_t st, int batch, int rows, int cols, const double* Xs, int64_t sXs, double* X,
                       int64_t sX, double* X2, int64_t sX2) {
  const int64_t total = (int64_t)rows * cols;
  if (total <= 0 || batch <= 0) return 0;
  dim3 grid((unsigned)std::min<int64_t>((total + 255) / 256, 4096), batch);
  stacked_to_complex_kernel<<<grid, 256, 0, st>>>(rows, cols, Xs, sXs, reinterpret_cast<double2*>(X), sX, X2, sX2);
  HPS_LAUNCH_CHECK("stacked_to_complex_kernel");
  return 0;
}

int complex_expand(cudaStream_t st, int batch, int rows, int cols, const double* X, int64_t sX, double* X2, int64_t sX2,
                   double alpha) {
  const int64_t total = (int64_t)rows * cols;
  if (total <= 0 || batch <= 0) return 0;
  dim3 grid((unsigned)std::min<int64_t>((total + 255) / 256, 4096), batch);
  complex_expand_kernel<<<grid, 256, 0, st>>>(rows, cols, reinterpret_cast<const double2*>(X), sX, X2, sX2, alpha);
  HPS_LAUNCH_CHECK("complex_expand_kernel");
  return 0;
}

// C[b] = alpha * A[b] * B[b] + beta * C[b], complex128 interleaved, alpha/beta real.  B must be
// contiguous (K x N); lda/ldc in complex elements.  ws: batch * 4 K N doubles.
int zgemm(cudaStream_t st, int M, int N, int K, double alpha, const double* A, int64_t lda, int64_t sA, const double* B,
          int64_t sB, double beta, double* C, int64_t ldc, int64_t sC, int batch, void* ws) {
  if (M <= 0 || N <= 0 || K <= 0 || batch <= 0) return 0;
  double* B2 = static_cast<double*>(ws);
  HPS_TRY(complex_expand(st, batch, K, N, B, sB, B2, (int64_t)4 * K * N, alpha));
  return dgemm(st, M, 2 * N, 2 * K, 1.0, A, 2 * lda, 2 * sA, B2, 2 * N, (int64_t)4 * K * N, beta, C, 2 * ldc, 2 * sC, batch);
}

// ---- dense complex solve  X = A^-1 B  (A: n x n, B: n x nrhs, complex128 interleaved, neither is modified) ----
// Used by the ItI -> DtN conversion of the top-level operator and the BIE coupling system of the scattering
// application (reference examples/wave_scattering_utils.py:31-49,96-242, where it is jnp.linalg.solve): the real
// embedding [[Ar,-Ai],[Ai,Ar]] [Xr;Xi] = [Br;Bi] goes through the same pivoted FP64 LU as every other solve.
namespace {
__global__ void complex_embed_kernel(int n, const double2* __restrict__ A, int64_t lda, double* __restrict__ Ae) {
  const int64_t total = (int64_t)n * n, ld = 2 * (int64_t)n;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / n, c = e - r * n;
    const double2 z = A[r * lda + c];
    Ae[r * ld + c] = z.x;
    Ae[r * ld + n + c] = -z.y;
    Ae[(n + r) * ld + c] = z.y;
    Ae[(n + r) * ld + n + c] = z.x;
  }
}
// power-of-two row equilibration of the embedded matrix (see iti_assemble_kernel): one warp per complex row r scales
// rows r and n + r of Ae by s_r = 2^-ilogb(max_c |Ae_rc|) and records s_r
__global__ void __launch_bounds__(256) complex_row_scale_kernel(int n, double* __restrict__ Ae, double* __restrict__ scale) {
  const int lane = threadIdx.x & 31;
  const int64_t ld = 2 * (int64_t)n;
  for (int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n; r += gridDim.x * (blockDim.x >> 5)) {
    double* top = Ae + (int64_t)r * ld;
    double* bot = Ae + (int64_t)(n + r) * ld;
    double m = 0.0;
    for (int64_t c = lane; c < ld; c += 32) m = fmax(m, fabs(top[c]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    const double sc = (m > 0.0 && isfinite(m)) ? scalbn(1.0, -ilogb(m)) : 1.0;
    if (lane == 0) scale[r] = sc;
    if (sc != 1.0)
      for (int64_t c = lane; c < ld; c += 32) { top[c] *= sc; bot[c] *= sc; }
  }
}
__global__ void complex_stack_kernel(int n, int nrhs, const double2* __restrict__ B, int64_t ldb,
                                     const double* __restrict__ scale, double* __restrict__ Bs) {
  const int64_t total = (int64_t)n * nrhs;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / nrhs, c = e - r * nrhs;
    const double2 z = B[r * ldb + c];
    const double sc = scale[r];
    Bs[e] = sc * z.x;
    Bs[total + e] = sc * z.y;
  }
}
}  // namespace

size_t zgesv_workspace_bytes(int n, int nrhs) {
  const size_t n2 = 2 * (size_t)n;
  return align_up(n2 * n2 * 8, 256) + align_up(n2 * (size_t)nrhs * 8, 256) + align_up((size_t)n * 8, 256) +
         lu_workspace_bytes(1, (int)n2) + 1024;
}

int zgesv(cudaStream_t st, int n, int nrhs, const double* A, int64_t lda, const double* B, int64_t ldb, double* X, void* ws,
          size_t ws_bytes, int* info) {
  if (n <= 0 || nrhs <= 0) return fail_arg(2, "non-positive size");
  const int n2 = 2 * n;
  Arena ar(ws, ws_bytes);
  double* Ae = ar.take<double>((size_t)n2 * n2);
  double* Bs = ar.take<double>((size_t)n2 * nrhs);
  double* scale = ar.take<double>((size_t)n);
  if (!Ae || !Bs || !scale) return fail_arg(9, "zgesv: workspace too small");
  void* lu_ws = ar.base + ar.off;
  const size_t lu_ws_bytes = ar.cap - ar.off;
  const int64_t tA = (int64_t)n * n, tB = (int64_t)n * nrhs;
  complex_embed_kernel<<<(unsigned)std::min<int64_t>((tA + 255) / 256, 4096), 256, 0, st>>>(
      n, reinterpret_cast<const double2*>(A), lda, Ae);
  HPS_LAUNCH_CHECK("complex_embed_kernel");
  complex_row_scale_kernel<<<(unsigned)std::min((n + 7) / 8, 2048), 256, 0, st>>>(n, Ae, scale);
  HPS_LAUNCH_CHECK("complex_row_scale_kernel");
  complex_stack_kernel<<<(unsigned)std::min<int64_t>((tB + 255) / 256, 4096), 256, 0, st>>>(
      n, nrhs, reinterpret_cast<const double2*>(B), ldb, scale, Bs);
  HPS_LAUNCH_CHECK("complex_stack_kernel");
  RhsDesc rhs[1] = {{Bs, nrhs, (int64_t)n2 * nrhs, nrhs}};
  HPS_TRY(lu_solve(st, 1, n2, Ae, n2, (int64_t)n2 * n2, 1, rhs, lu_ws, lu_ws_bytes, info));
  return stacked_to_complex(st, 1, n, nrhs, Bs, (int64_t)n2 * nrhs, X, (int64_t)n * nrhs, nullptr, 0);
}

size_t local_solve_iti_workspace_bytes(int n_leaves, int p, int q, int n_src) {
  const LeafGeom g = geom(2, p);
  const size_t n_g = 4 * (size_t)q, n2 = 2 * (size_t)g.n_c;
  return align_up((size_t)n_leaves * n2 * n2 * 8, 256) + align_up((size_t)n_leaves * n2 * n_g * 8, 256) +
         align_up((size_t)n_leaves * n2 * n_src * 8, 256) + align_up((size_t)n_leaves * n2 * 2 * n_g * 8, 256) +
         align_up((size_t)n_leaves * n2 * 2 * n_src * 8, 256) + align_up((size_t)n_leaves * g.n_c * 8, 256) +
         lu_workspace_bytes(n_leaves, (int)n2) + 1024;
}

// Complex outputs are interleaved (re, im) doubles: Y [n][n_c][n_g], R [n][n_g][n_g], v [n][n_c][n_src],
// h [n][n_g][n_src]; G [n_b][n_c], QH [n_g][n_c], src [n][n_c][n_src] likewise.
int local_solve_iti(cudaStream_t st, int n_leaves, int p, int q, int n_src, const uint8_t* which, const double* coeffs,
                    const double* D1, const double* P, const double* G, const double* QH, const double* src, double* Y,
                    double* R, double* v, double* h, void* ws, size_t ws_bytes, int* info, const double* coeffs_imag) {
  if (p < 3 || p > MAX_P) return fail_arg(3, "p out of range [3, 32]");
  if (n_leaves <= 0 || n_src <= 0 || q <= 0) return fail_arg(2, "non-positive size");
  if (n_leaves > 65535) return fail_arg(2, "n_leaves per call is limited to 65535; chunk the leaves");
  const LeafGeom geo = geom(2, p);
  const int n_c = geo.n_c, n_b = geo.n_b, n_g = 4 * q, n2 = 2 * n_c;
  Arena ar(ws, ws_bytes);
  double* Be = ar.take<double>((size_t)n_leaves * n2 * n2);
  double* Ys = ar.take<double>((size_t)n_leaves * n2 * n_g);
  double* vs = ar.take<double>((size_t)n_leaves * n2 * n_src);
  double* Y2 = ar.take<double>((size_t)n_leaves * n2 * 2 * n_g);
  double* v2 = ar.take<double>((size_t)n_leaves * n2 * 2 * n_src);
  double* row_scale = ar.take<double>((size_t)n_leaves * n_c);
  if (!Be || !Ys || !vs || !Y2 || !v2 || !row_scale) return fail_arg(17, "local_solve_iti: workspace too small");
  void* lu_ws = ar.base + ar.off;
  const size_t lu_ws_bytes = ar.cap - ar.off;

  ItiLeafArgs ia;
  ia.a.geo = geo; ia.a.n_leaves = n_leaves; ia.a.coeffs = coeffs; ia.a.D1 = D1; ia.a.Aii = nullptr; ia.a.Aie = nullptr;
  int n_coef = 0;
  for (int k = 0; k < 10; ++k) ia.a.slot[k] = (k < 6 && which[k]) ? n_coef++ : -1;
  ia.a.n_coef = n_coef;
  ia.G = reinterpret_cast<const double2*>(G);
  ia.Be = Be;
  ia.coeffs_im = coeffs_imag;
  ia.row_scale = row_scale;
  {
    iti_assemble_kernel<<<dim3((unsigned)std::min((n_c + 7) / 8, 32), n_leaves), 256, 0, st>>>(ia);
    HPS_LAUNCH_CHECK("iti_assemble_kernel");
    const int64_t tot2 = (int64_t)n2 * (n_g + n_src);
    iti_rhs_kernel<<<dim3((unsigned)std::min<int64_t>((tot2 + 255) / 256, 1024), n_leaves), 256, 0, st>>>(
        n_c, n_b, n_g, n_src, P, reinterpret_cast<const double2*>(src), row_scale, Ys, vs);
    HPS_LAUNCH_CHECK("iti_rhs_kernel");
  }
  RhsDesc rhs[2] = {{Ys, n_g, (int64_t)n2 * n_g, n_g}, {vs, n_src, (int64_t)n2 * n_src, n_src}};
  HPS_TRY(lu_solve(st, n_leaves, n2, Be, n2, (int64_t)n2 * n2, 2, rhs, lu_ws, lu_ws_bytes, info));
  // complex outputs + expanded copies for the products with QH
  HPS_TRY(stacked_to_complex(st, n_leaves, n_c, n_g, Ys, (int64_t)n2 * n_g, Y, (int64_t)n_c * n_g, Y2, (int64_t)n2 * 2 * n_g));
  HPS_TRY(stacked_to_complex(st, n_leaves, n_c, n_src, vs, (int64_t)n2 * n_src, v, (int64_t)n_c * n_src, v2,
                             (int64_t)n2 * 2 * n_src));
  // R = QH Y and h = QH v as real products on the interleaved views
  HPS_TRY(dgemm(st, n_g, 2 * n_g, n2, 1.0, QH, n2, 0, Y2, 2 * n_g, (int64_t)n2 * 2 * n_g, 0.0, R, 2 * n_g,
                (int64_t)n_g * 2 * n_g, n_leaves));
  HPS_TRY(dgemm(st, n_g, 2 * n_src, n2, 1.0, QH, n2, 0, v2, 2 * n_src, (int64_t)n2 * 2 * n_src, 0.0, h, 2 * n_src,
                (int64_t)n_g * 2 * n_src, n_leaves));
  return 0;
}

namespace {
}  // namespace

size_t local_solve_workspace_bytes(int dim, int n_leaves, int p, int q) {
  const LeafGeom g = geom(dim, p);
  const size_t n_g = 2 * (size_t)dim * ipow(q, dim - 1);
  return align_up((size_t)n_leaves * g.n_i * g.n_i * sizeof(double), 256) +
         align_up((size_t)n_leaves * g.n_i * g.n_b * sizeof(double), 256) + align_up(n_g * n_g * sizeof(double), 256) +
         lu_workspace_bytes(n_leaves, g.n_i) + 1024;
}

int local_solve_dtn(cudaStream_t st, int dim, int n_leaves, int p, int q, int n_src, const uint8_t* which,
                    const double* coeffs, const double* D1, const double* P, const double* Q, const double* src,
                    double* Y, double* T, double* v, double* h, void* ws, size_t ws_bytes, int* info) {
  if (dim != 2 && dim != 3) return fail_arg(2, "dim must be 2 or 3");
  if (p < 3 || p > MAX_P) return fail_arg(4, "p out of range [3, 32]");
  if (n_leaves <= 0 || n_src <= 0 || q <= 0) return fail_arg(3, "non-positive size");
  if (n_leaves > 65535) return fail_arg(3, "n_leaves per call is limited to 65535; chunk the leaves");
  const LeafGeom geo = geom(dim, p);
  const int n_g = 2 * dim * ipow(q, dim - 1);
  const int n_which = dim == 3 ? 10 : 6;

  Arena ar(ws, ws_bytes);
  double* Aii = ar.take<double>((size_t)n_leaves * geo.n_i * geo.n_i);
  double* Aie = ar.take<double>((size_t)n_leaves * geo.n_i * geo.n_b);
  double* QbP = ar.take<double>((size_t)n_g * n_g);
  if (!Aii || !Aie || !QbP) return fail_arg(17, "local_solve: workspace too small");
  void* lu_ws = ar.base + ar.off;
  const size_t lu_ws_bytes = ar.cap - ar.off;

  AssembleArgs aa;
  aa.geo = geo; aa.n_leaves = n_leaves; aa.coeffs = coeffs; aa.D1 = D1; aa.Aii = Aii; aa.Aie = Aie;
  int n_coef = 0;
  for (int k = 0; k < 10; ++k) aa.slot[k] = (k < n_which && which[k]) ? n_coef++ : -1;
  aa.n_coef = n_coef;
  {
    const int64_t total = (int64_t)geo.n_i * geo.n_c;
    // enough CTAs to fill the GPU a few times over, each handling several rows of one leaf
    const int bx = std::max(1, std::min(geo.n_i, (148 * 16 + n_leaves - 1) / n_leaves));
    const size_t idx_bytes = 4 * (size_t)geo.n_c;
    prof_begin(PROF_ASSEMBLE, st, 8.0 * n_leaves * (double)total);
    const bool mixed = aa.slot[1] >= 0 || aa.slot[3] >= 0 || aa.slot[4] >= 0;
    static const bool lines_off = std::getenv("HPS_ASSEMBLE_LINES") && std::atoi(std::getenv("HPS_ASSEMBLE_LINES")) == 0;
    if (dim == 3 && !mixed && 2 * (size_t)geo.n_c <= 30000 && !lines_off) {
      // ~10 waves of CTAs, each with a few passes of AL_ROWS rows
      const int passes = (geo.n_i + AL_ROWS - 1) / AL_ROWS;
      const int bl = std::max(1, std::min(passes, (148 * 8 * 10 + n_leaves - 1) / n_leaves));
      assemble3_lines_kernel<<<dim3(bl, n_leaves), 256, 2 * (size_t)geo.n_c, st>>>(aa);
    } else if (dim == 3) {
      assemble_kernel<3><<<dim3(bx, n_leaves), 256, idx_bytes, st>>>(aa);
    } else {
      assemble_kernel<2><<<dim3(bx, n_leaves), 256, idx_bytes, st>>>(aa);
    }
    prof_end(PROF_ASSEMBLE, st);
    HPS_LAUNCH_CHECK("assemble_kernel");
    const int64_t tot2 = (int64_t)geo.n_b * n_g + (int64_t)geo.n_c * n_src;
    leaf_init_kernel<<<dim3((int)std::min<int64_t>((tot2 + 255) / 256, 2048), n_leaves), 256, 0, st>>>(
        geo.n_c, geo.n_b, n_g, n_src, P, src, Y, v);
    HPS_LAUNCH_CHECK("leaf_init_kernel");
  }
  const int64_t sY = (int64_t)geo.n_c * n_g, sV = (int64_t)geo.n_c * n_src;
  double* Yint = Y + (int64_t)geo.n_b * n_g;
  double* vint = v + (int64_t)geo.n_b * n_src;
  // Y_int := -A_ie P
  HPS_TRY(dgemm(st, geo.n_i, n_g, geo.n_b, -1.0, Aie, geo.n_b, (int64_t)geo.n_i * geo.n_b, P, n_g, 0, 0.0, Yint, n_g, sY,
                n_leaves));
  // [Y_int | v_int] := A_ii^-1 [Y_int | v_int]
  RhsDesc rhs[2] = {{Yint, n_g, sY, n_g}, {vint, n_src, sV, n_src}};
  // A_ii of a spectral collocation operator: partial pivoting reshuffles a few rows locally, but the multipliers of the
  // block-local choice stay near 1 (measured 1.16 at p = 12) — threshold pivoting keeps the exchange-free block columns
  HPS_TRY(lu_solve(st, n_leaves, geo.n_i, Aii, geo.n_i, (int64_t)geo.n_i * geo.n_i, 2, rhs, lu_ws, lu_ws_bytes, info,
                   LU_NO_PIVOT_EXPECTED | LU_PIVOT_IN_BLOCK | LU_THRESHOLD_PIVOTING));
  // T = Q Y = (Q_b P) + Q_i Y_int : the first term is the same for every leaf and is formed once;
  // h = Q v = Q_i v_int because v vanishes on the boundary rows.
  HPS_TRY(dgemm(st, n_g, n_g, geo.n_b, 1.0, Q, geo.n_c, 0, P, n_g, 0, 0.0, QbP, n_g, 0, 1));
  HPS_TRY(dgemm_affine(st, n_g, n_g, geo.n_i, Q + geo.n_b, geo.n_c, 0, Yint, n_g, sY, QbP, n_g, 0, T, n_g,
                       (int64_t)n_g * n_g, n_leaves));
  HPS_TRY(dgemm(st, n_g, n_src, geo.n_i, 1.0, Q + geo.n_b, geo.n_c, 0, vint, n_src, sV, 0.0, h, n_src,
                (int64_t)n_g * n_src, n_leaves));
  return 0;
}

}  // namespace hps

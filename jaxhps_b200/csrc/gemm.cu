// Host side of the FP64 GEMMs: the DMMA kernel lives in gemm_kernel.cuh (templated on the tile
// configuration), the narrow-N bandwidth kernel and the dispatch are here.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include <type_traits>
#include "gemm_kernel.cuh"

namespace hps {

namespace {

// Tile configuration chosen with tools/gemm_lab.cu on B200 (profiles/r01_gemm_lab.txt):
// 128x64 CTA tile, 8 warps of 32x32, K step 16, 3-stage ring, two CTAs per SM, per-stage
// full/empty mbarriers instead of a CTA barrier per K step.
//   8192^3: 32.9 TF/s (cuBLAS 35.4);  15360^2 rank-128 update: 28.3 (cuBLAS 23.2);
//   512 x (872^2 rank-128 update): 25.9 (cuBLAS 22.2).
using Cfg = gemmk::Config<32, 32, 4, 2, 16, 3, 2>;
using gemmk::GemmArgs;
constexpr int BM = Cfg::BM, BN = Cfg::BN, THREADS = Cfg::THREADS;
constexpr size_t SMEM_BYTES = Cfg::SMEM_BYTES;
constexpr int RASTER = 12;  // 12 x 24 tiles of 128x64 = a 1536^2 block of C per wave of 296 CTAs
#define gemm_kernel gemmk::gemm_kernel_hoist<Cfg>

// ---- narrow-N kernels: one warp per output row, lanes stride over K -------------------------
// These are the HBM-bound mat-vecs of the down pass (g_int = S g_ext + g~, u = Y g + v): every
// entry of A is read exactly once, so the figure of merit is GB/s.
template <int NMAX>
__global__ void __launch_bounds__(256) skinny_kernel(int M, int N, int K, double alpha,
                                                     const double* __restrict__ A, int64_t lda, int64_t sA,
                                                     const double* __restrict__ B, int64_t ldb, int64_t sB,
                                                     double beta, const double* __restrict__ Cin, int64_t ldcin,
                                                     int64_t sCin, double* __restrict__ C, int64_t ldc, int64_t sC,
                                                     int n0) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp;
  const int64_t batch = blockIdx.y;
  if (row >= M) return;
  const double* a = A + batch * sA + (int64_t)row * lda;
  const double* b = B + batch * sB + n0;
  const int nn = min(NMAX, N - n0);
  double acc[NMAX];
#pragma unroll
  for (int n = 0; n < NMAX; ++n) acc[n] = 0.0;
  for (int k = lane; k < K; k += 32) {
    const double av = a[k];
    const double* bk = b + (int64_t)k * ldb;
#pragma unroll
    for (int n = 0; n < NMAX; ++n)
      if (n < nn) acc[n] = fma(av, bk[n], acc[n]);
  }
#pragma unroll
  for (int n = 0; n < NMAX; ++n) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], o);
  }
  if (lane == 0) {
    double* c = C + batch * sC + (int64_t)row * ldc + n0;
    const double* ci = Cin ? Cin + batch * sCin + (int64_t)row * ldcin + n0 : nullptr;
#pragma unroll
    for (int n = 0; n < NMAX; ++n)
      if (n < nn) {
        double r = alpha * acc[n];
        if (beta != 0.0 && ci) r += beta * ci[n];
        c[n] = r;
      }
  }
}

// N == 1 with 16-byte-aligned rows: the common single-source case.  Each lane streams double2
// pairs with four independent loads in flight (2 KB per warp), the vector x stays in L1/L2.
__global__ void __launch_bounds__(256) gemv_kernel(int M, int K, double alpha, const double* __restrict__ A, int64_t lda,
                                                   int64_t sA, const double* __restrict__ x, int64_t sB, double beta,
                                                   const double* __restrict__ Cin, int64_t ldcin, int64_t sCin,
                                                   double* __restrict__ C, int64_t ldc, int64_t sC) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp;
  const int64_t batch = blockIdx.y;
  if (row >= M) return;
  const double2* a = reinterpret_cast<const double2*>(A + batch * sA + (int64_t)row * lda);
  const double2* xv = reinterpret_cast<const double2*>(x + batch * sB);
  const int K2 = K >> 1;
  double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
  int k = lane;
  for (; k + 96 < K2; k += 128) {
    const double2 a0 = __ldcs(a + k), a1 = __ldcs(a + k + 32), a2 = __ldcs(a + k + 64), a3 = __ldcs(a + k + 96);
    const double2 x0 = xv[k], x1 = xv[k + 32], x2 = xv[k + 64], x3 = xv[k + 96];
    acc0 = fma(a0.x, x0.x, fma(a0.y, x0.y, acc0));
    acc1 = fma(a1.x, x1.x, fma(a1.y, x1.y, acc1));
    acc2 = fma(a2.x, x2.x, fma(a2.y, x2.y, acc2));
    acc3 = fma(a3.x, x3.x, fma(a3.y, x3.y, acc3));
  }
  for (; k < K2; k += 32) {
    const double2 a0 = __ldcs(a + k);
    const double2 x0 = xv[k];
    acc0 = fma(a0.x, x0.x, fma(a0.y, x0.y, acc0));
  }
  double acc = (acc0 + acc1) + (acc2 + acc3);
  if ((K & 1) && lane == 0) acc = fma(A[batch * sA + (int64_t)row * lda + K - 1], x[batch * sB + K - 1], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    double r = alpha * acc;
    if (beta != 0.0 && Cin) r += beta * Cin[batch * sCin + (int64_t)row * ldcin];
    C[batch * sC + (int64_t)row * ldc] = r;
  }
}

// C[b] (K x N) = alpha * A[b]^T * X[b] + beta * C[b] for narrow N (adjoint passes: S^T g, Y^T w, Phi^T v): A is M x K
// row-major, so a thread per COLUMN k reads coalesced rows; X[b] (M x N) is broadcast.  E = double or double2
// (complex128, plain transpose — no conjugation).  grid: (column tiles, batch, row splits); with more than one row
// split the partial sums are added atomically onto a C that gemv_t_scale_kernel has already scaled by beta.
template <typename E>
__device__ __forceinline__ E e_zero();
template <>
__device__ __forceinline__ double e_zero<double>() { return 0.0; }
template <>
__device__ __forceinline__ double2 e_zero<double2>() { return make_double2(0.0, 0.0); }
__device__ __forceinline__ double e_fma(double a, double x, double acc) { return fma(a, x, acc); }
__device__ __forceinline__ double2 e_fma(double2 a, double2 x, double2 acc) {
  acc.x = fma(a.x, x.x, fma(-a.y, x.y, acc.x));
  acc.y = fma(a.x, x.y, fma(a.y, x.x, acc.y));
  return acc;
}
__device__ __forceinline__ double e_scale(double s, double a) { return s * a; }
__device__ __forceinline__ double2 e_scale(double s, double2 a) { return make_double2(s * a.x, s * a.y); }
__device__ __forceinline__ double e_add(double a, double b) { return a + b; }
__device__ __forceinline__ double2 e_add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ void e_atomic_add(double* p, double v) { atomicAdd(p, v); }
__device__ __forceinline__ void e_atomic_add(double2* p, double2 v) { atomicAdd(&p->x, v.x); atomicAdd(&p->y, v.y); }

constexpr int GT_NMAX = 8;
constexpr int GT_UNROLL = 8;
constexpr int GT_TARGET_CTAS = 148 * 8 * 8;  // 8 CTAs of 128 threads are resident per SM (60 registers): ~8 waves, so the last one costs little
template <typename E>
__global__ void __launch_bounds__(128) gemv_t_kernel(int M, int K, int N, double alpha, const E* __restrict__ A, int64_t lda,
                                                     int64_t sA, const E* __restrict__ X, int64_t ldx, int64_t sX, double beta,
                                                     E* __restrict__ C, int64_t ldc, int64_t sC, int rows_per_split) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t b = blockIdx.y;
  const int m0 = blockIdx.z * rows_per_split, m1 = min(M, m0 + rows_per_split);
  if (k >= K) return;
  const E* a = A + b * sA + k;
  const E* x = X + b * sX;
  E acc[GT_NMAX];
#pragma unroll
  for (int n = 0; n < GT_NMAX; ++n) acc[n] = e_zero<E>();
  int m = m0;
  for (; m + GT_UNROLL <= m1; m += GT_UNROLL) {  // GT_UNROLL independent loads in flight per thread
    E av[GT_UNROLL];
#pragma unroll
    for (int u = 0; u < GT_UNROLL; ++u) av[u] = __ldcs(a + (int64_t)(m + u) * lda);
#pragma unroll
    for (int u = 0; u < GT_UNROLL; ++u) {
#pragma unroll
      for (int n = 0; n < GT_NMAX; ++n)
        if (n < N) acc[n] = e_fma(av[u], x[(int64_t)(m + u) * ldx + n], acc[n]);
    }
  }
  for (; m < m1; ++m) {
    const E av = a[(int64_t)m * lda];
#pragma unroll
    for (int n = 0; n < GT_NMAX; ++n)
      if (n < N) acc[n] = e_fma(av, x[(int64_t)m * ldx + n], acc[n]);
  }
  E* c = C + b * sC + (int64_t)k * ldc;
  if (gridDim.z == 1) {
#pragma unroll
    for (int n = 0; n < GT_NMAX; ++n)
      if (n < N) c[n] = (beta != 0.0) ? e_add(e_scale(alpha, acc[n]), e_scale(beta, c[n])) : e_scale(alpha, acc[n]);
  } else {
#pragma unroll
    for (int n = 0; n < GT_NMAX; ++n)
      if (n < N) e_atomic_add(&c[n], e_scale(alpha, acc[n]));
  }
}
// real case with 16-byte aligned rows: two adjacent columns per thread (a CTA row-read is 2 KB contiguous)
__global__ void __launch_bounds__(128) gemv_t_pair_kernel(int M, int K, int N, double alpha, const double* __restrict__ A,
                                                          int64_t lda, int64_t sA, const double* __restrict__ X, int64_t ldx,
                                                          int64_t sX, double beta, double* __restrict__ C, int64_t ldc,
                                                          int64_t sC, int rows_per_split) {
  const int k = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  const int64_t b = blockIdx.y;
  const int m0 = blockIdx.z * rows_per_split, m1 = min(M, m0 + rows_per_split);
  if (k >= K) return;  // K is even here
  const double* a = A + b * sA + k;
  const double* x = X + b * sX;
  double acc0[GT_NMAX], acc1[GT_NMAX];
#pragma unroll
  for (int n = 0; n < GT_NMAX; ++n) acc0[n] = acc1[n] = 0.0;
  int m = m0;
  for (; m + GT_UNROLL <= m1; m += GT_UNROLL) {
    double2 av[GT_UNROLL];
#pragma unroll
    for (int u = 0; u < GT_UNROLL; ++u) av[u] = __ldcs(reinterpret_cast<const double2*>(a + (int64_t)(m + u) * lda));
#pragma unroll
    for (int u = 0; u < GT_UNROLL; ++u) {
#pragma unroll
      for (int n = 0; n < GT_NMAX; ++n)
        if (n < N) {
          const double xv = x[(int64_t)(m + u) * ldx + n];
          acc0[n] = fma(av[u].x, xv, acc0[n]);
          acc1[n] = fma(av[u].y, xv, acc1[n]);
        }
    }
  }
  for (; m < m1; ++m) {
    const double2 av = *reinterpret_cast<const double2*>(a + (int64_t)m * lda);
#pragma unroll
    for (int n = 0; n < GT_NMAX; ++n)
      if (n < N) {
        const double xv = x[(int64_t)m * ldx + n];
        acc0[n] = fma(av.x, xv, acc0[n]);
        acc1[n] = fma(av.y, xv, acc1[n]);
      }
  }
  double* c0 = C + b * sC + (int64_t)k * ldc;
  double* c1 = c0 + ldc;
#pragma unroll
  for (int n = 0; n < GT_NMAX; ++n)
    if (n < N) {
      if (gridDim.z == 1) {
        c0[n] = (beta != 0.0) ? alpha * acc0[n] + beta * c0[n] : alpha * acc0[n];
        c1[n] = (beta != 0.0) ? alpha * acc1[n] + beta * c1[n] : alpha * acc1[n];
      } else {
        atomicAdd(&c0[n], alpha * acc0[n]);
        atomicAdd(&c1[n], alpha * acc1[n]);
      }
    }
}
template <typename E>
__global__ void gemv_t_scale_kernel(int K, int N, double beta, E* __restrict__ C, int64_t ldc, int64_t sC) {
  const int64_t total = (int64_t)K * N;
  E* c = C + (int64_t)blockIdx.y * sC;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = e / N, n = e - k * N;
    c[k * ldc + n] = (beta != 0.0) ? e_scale(beta, c[k * ldc + n]) : e_zero<E>();
  }
}

inline bool vec_ok(const void* p, int64_t ld, int64_t stride) {
  return (reinterpret_cast<uintptr_t>(p) % 16 == 0) && (ld % 2 == 0) && (stride % 2 == 0);
}

}  // namespace

int dgemm_skinny(cudaStream_t st, int M, int N, int K, double alpha, const double* A, int64_t lda,
                 int64_t sA, const double* B, int64_t ldb, int64_t sB, double beta, const double* Cin,
                 int64_t ldcin, int64_t sCin, double* C, int64_t ldc, int64_t sC, int batch) {
  if (M <= 0 || N <= 0 || batch <= 0) return 0;
  const int wpb = 8;
  prof_begin(PROF_SKINNY, st, 8.0 * M * (double)K * batch);
  const bool vec = N == 1 && ldb == 1 && (reinterpret_cast<uintptr_t>(A) % 16 == 0) && (lda % 2 == 0) && (sA % 2 == 0) &&
                   (reinterpret_cast<uintptr_t>(B) % 16 == 0) && (sB % 2 == 0);
  for (int b0 = 0; b0 < batch; b0 += 65535) {
    const int nb = min(65535, batch - b0);
    dim3 grid((M + wpb - 1) / wpb, nb);
    if (vec) {
      gemv_kernel<<<grid, wpb * 32, 0, st>>>(M, K, alpha, A + b0 * sA, lda, sA, B + b0 * sB, sB, beta,
                                             Cin ? Cin + b0 * sCin : nullptr, ldcin, sCin, C + b0 * sC, ldc, sC);
      continue;
    }
    for (int n0 = 0; n0 < N; n0 += 4) {
      skinny_kernel<4><<<grid, wpb * 32, 0, st>>>(M, N, K, alpha, A + b0 * sA, lda, sA, B + b0 * sB, ldb, sB,
                                                  beta, Cin ? Cin + b0 * sCin : nullptr, ldcin, sCin,
                                                  C + b0 * sC, ldc, sC, n0);
    }
  }
  prof_end(PROF_SKINNY, st);
  HPS_LAUNCH_CHECK("skinny_kernel");
  return 0;
}

namespace {
__global__ void copy2d_kernel(double* dst, int64_t ldd, int64_t sD, const double* src, int64_t lds, int64_t sS,
                              int rows, int cols) {
  const int64_t total = (int64_t)rows * cols;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / cols, c = idx - r * cols;
    dst[(int64_t)blockIdx.y * sD + r * ldd + c] = src[(int64_t)blockIdx.y * sS + r * lds + c];
  }
}
}  // namespace

// C = A*B + Cin (C and Cin distinct buffers): the affine maps of the down pass.
int dgemm_affine(cudaStream_t st, int M, int N, int K, const double* A, int64_t lda, int64_t sA, const double* B,
                 int64_t ldb, int64_t sB, const double* Cin, int64_t ldcin, int64_t sCin, double* C, int64_t ldc,
                 int64_t sC, int batch) {
  if (N < 16)
    return dgemm_skinny(st, M, N, K, 1.0, A, lda, sA, B, ldb, sB, 1.0, Cin, ldcin, sCin, C, ldc, sC, batch);
  for (int b0 = 0; b0 < batch; b0 += 65535) {
    const int nb = min(65535, batch - b0);
    const int64_t total = (int64_t)M * N;
    copy2d_kernel<<<dim3((unsigned)std::min<int64_t>((total + 255) / 256, 2048), nb), 256, 0, st>>>(
        C + (int64_t)b0 * sC, ldc, sC, Cin + (int64_t)b0 * sCin, ldcin, sCin, M, N);
  }
  HPS_LAUNCH_CHECK("copy2d_kernel");
  return dgemm(st, M, N, K, 1.0, A, lda, sA, B, ldb, sB, 1.0, C, ldc, sC, batch);
}

int dgemm(cudaStream_t st, int M, int N, int K, double alpha, const double* A, int64_t lda, int64_t sA,
          const double* B, int64_t ldb, int64_t sB, double beta, double* C, int64_t ldc, int64_t sC,
          int batch) {
  if (M <= 0 || N <= 0 || batch <= 0) return 0;
  if (N < 16) {
    return dgemm_skinny(st, M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, C, ldc, sC, batch);
  }
  DeviceState* ds = nullptr;
  HPS_TRY(device_state(ds));
  if (!ds->gemm_configured.load(std::memory_order_acquire)) {  // per device; idempotent, so a race is harmless
    HPS_CUDA(cudaFuncSetAttribute(gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    ds->gemm_configured.store(true, std::memory_order_release);
  }
  GemmArgs g;
  g.M = M; g.N = N; g.K = K; g.alpha = alpha; g.beta = beta;
  g.lda = lda; g.sA = sA; g.ldb = ldb; g.sB = sB; g.ldc = ldc; g.sC = sC;
  g.vecA = vec_ok(A, lda, sA); g.vecB = vec_ok(B, ldb, sB); g.vecC = vec_ok(C, ldc, sC);
  prof_begin(PROF_GEMM, st, 2.0 * M * N * (double)K * batch);
  prof_dims(PROF_GEMM, M, N, K, batch);
  for (int b0 = 0; b0 < batch; b0 += 65535) {
    const int nb = min(65535, batch - b0);
    g.A = A + (int64_t)b0 * sA; g.B = B + (int64_t)b0 * sB; g.C = C + (int64_t)b0 * sC;
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, nb);
    // large single products: launch order grouped into RASTER tile rows (profiles/r02_gemm_lab.txt)
    g.raster = (grid.y >= 2 * RASTER && grid.x >= 8) ? RASTER : 0;
    gemm_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(g);
  }
  prof_end(PROF_GEMM, st);
  HPS_LAUNCH_CHECK("gemm_kernel");
  return 0;
}


template <typename E>
static int gemv_t_impl(cudaStream_t st, int M, int K, int N, double alpha, const E* A, int64_t lda, int64_t sA, const E* X,
                       int64_t ldx, int64_t sX, double beta, E* C, int64_t ldc, int64_t sC, int batch) {
  for (int n0 = 0; n0 < N; n0 += GT_NMAX) {
    const int nn = std::min(GT_NMAX, N - n0);
    for (int b0 = 0; b0 < batch; b0 += 65535) {
      const int nb = std::min(65535, batch - b0);
      const bool pair = std::is_same<E, double>::value && K % 2 == 0 && vec_ok(A, lda, sA);
      const int tiles = pair ? (K / 2 + 127) / 128 : (K + 127) / 128;
      // few CTAs (the upper tree levels): split the rows (>= 128 per CTA) into ~8 waves of CTAs (profiles/r02_hbm_kernels_summary.txt)
      int splits = 1;
      if ((int64_t)tiles * nb < GT_TARGET_CTAS && M >= 256)
        splits = (int)std::min<int64_t>((M + 127) / 128, (GT_TARGET_CTAS + (int64_t)tiles * nb - 1) / ((int64_t)tiles * nb));
      const int rps = (M + splits - 1) / splits;
      const E* Ab = A + (int64_t)b0 * sA;
      const E* Xb = X + (int64_t)b0 * sX + n0;
      E* Cb = C + (int64_t)b0 * sC + n0;
      if (splits > 1) {
        gemv_t_scale_kernel<E><<<dim3((unsigned)std::min<int64_t>(((int64_t)K * nn + 255) / 256, 256), nb), 256, 0, st>>>(K, nn, beta, Cb, ldc, sC);
      }
      if (pair)
        gemv_t_pair_kernel<<<dim3(tiles, nb, splits), 128, 0, st>>>(
            M, K, nn, alpha, reinterpret_cast<const double*>(Ab), lda, sA, reinterpret_cast<const double*>(Xb), ldx, sX, beta,
            reinterpret_cast<double*>(Cb), ldc, sC, rps);
      else
        gemv_t_kernel<E><<<dim3(tiles, nb, splits), 128, 0, st>>>(M, K, nn, alpha, Ab, lda, sA, Xb, ldx, sX, beta, Cb, ldc, sC, rps);
    }
  }
  HPS_LAUNCH_CHECK("gemv_t_kernel");
  return 0;
}

int gemv_t(cudaStream_t st, int M, int K, int N, double alpha, const double* A, int64_t lda, int64_t sA, const double* X,
           int64_t ldx, int64_t sX, double beta, double* C, int64_t ldc, int64_t sC, int batch, int is_complex) {
  if (M <= 0 || K <= 0 || N <= 0 || batch <= 0) return 0;
  prof_begin(PROF_SKINNY, st, (is_complex ? 16.0 : 8.0) * M * (double)K * batch);
  int rc;
  if (is_complex)
    rc = gemv_t_impl<double2>(st, M, K, N, alpha, reinterpret_cast<const double2*>(A), lda, sA,
                              reinterpret_cast<const double2*>(X), ldx, sX, beta, reinterpret_cast<double2*>(C), ldc, sC, batch);
  else
    rc = gemv_t_impl<double>(st, M, K, N, alpha, A, lda, sA, X, ldx, sX, beta, C, ldc, sC, batch);
  prof_end(PROF_SKINNY, st);
  return rc;
}

}  // namespace hps

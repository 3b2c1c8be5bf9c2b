// FP64 GEMM for sm_100a built on DMMA (mma.sync.m8n8k4.f64), the only FP64 tensor shape the
// B200 executes natively (the m16n8k{4,8,16} PTX shapes lower to the same DMMA.8x8x4 SASS;
// tcgen05 has no f64 kind).  Measured on B200: DMMA peak 37.0 TFLOP/s, cuBLAS DGEMM 35.4.
//
// C[b] = alpha * A[b] (MxK) * B[b] (KxN) + beta * C[b]; all row-major.
//   CTA tile 128x128, K step 16, 4-stage cp.async (LDGSTS) ring, 16 warps each owning a
//   32x32 accumulator (4x4 DMMA tiles, 32 FP64 accumulators per thread).
//   Shared-memory leading dimensions are == 4 (mod 16) doubles, which makes both fragment
//   loads (A: lane -> (row=lane/4, k=lane%4); B: lane -> (k=lane%4, col=lane/4))
//   conflict-free per half-warp.
#include <algorithm>

#include "common.cuh"

namespace hps {

namespace {

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 4, THREADS = 512;
constexpr int LDA_S = BK + 4;   // 20 doubles
constexpr int LDB_S = BN + 4;   // 132 doubles
constexpr int A_STAGE = BM * LDA_S;
constexpr int B_STAGE = BK * LDB_S;
constexpr size_t SMEM_BYTES = sizeof(double) * STAGES * (A_STAGE + B_STAGE);

struct GemmArgs {
  int M, N, K;
  double alpha, beta;
  const double* A; int64_t lda, sA;
  const double* B; int64_t ldb, sB;
  double* C; int64_t ldc, sC;
  int vecA, vecB, vecC;  // 16-byte vector access allowed for that operand
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(THREADS, 1) gemm_kernel(GemmArgs g) {
  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = smem + STAGES * A_STAGE;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = (warp >> 2) * 32, wn = (warp & 3) * 32;
  const int bm0 = blockIdx.y * BM, bn0 = blockIdx.x * BN;
  const int64_t batch = blockIdx.z;
  const double* __restrict__ A = g.A + batch * g.sA;
  const double* B = g.B + batch * g.sB;  // may alias C (in-place M<=BM products)
  double* C = g.C + batch * g.sC;
  const int M = g.M, N = g.N, K = g.K;

  // each thread moves two 16-byte chunks of A and two of B per stage
  auto load_tile = [&](int stage, int k0) {
    double* as = As + stage * A_STAGE;
    double* bs = Bs + stage * B_STAGE;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = tid + i * THREADS;
      {  // A: 128 rows x 8 chunks
        const int r = c >> 3, kc = (c & 7) * 2;
        const int gr = bm0 + r, gk = k0 + kc;
        double* dst = as + r * LDA_S + kc;
        int valid = (gr < M) ? max(0, min(2, K - gk)) : 0;
        const double* src = valid ? (A + (int64_t)gr * g.lda + gk) : A;
        if (g.vecA) {
          cp_async16(dst, src, valid * 8);
        } else {
          cp_async8(dst, src, valid >= 1 ? 8 : 0);
          cp_async8(dst + 1, valid >= 2 ? src + 1 : A, valid >= 2 ? 8 : 0);
        }
      }
      {  // B: 16 rows x 64 chunks
        const int r = c >> 6, nc = (c & 63) * 2;
        const int gk = k0 + r, gn = bn0 + nc;
        double* dst = bs + r * LDB_S + nc;
        int valid = (gk < K) ? max(0, min(2, N - gn)) : 0;
        const double* src = valid ? (B + (int64_t)gk * g.ldb + gn) : B;
        if (g.vecB) {
          cp_async16(dst, src, valid * 8);
        } else {
          cp_async8(dst, src, valid >= 1 ? 8 : 0);
          cp_async8(dst + 1, valid >= 2 ? src + 1 : B, valid >= 2 ? 8 : 0);
        }
      }
    }
  };

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int KT = (K + BK - 1) / BK;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < KT) load_tile(s, s * BK);
    cp_async_commit();
  }

  const int a_off = (wm + (lane >> 2)) * LDA_S + (lane & 3);
  const int b_off = (lane & 3) * LDB_S + wn + (lane >> 2);

  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nk = kt + STAGES - 1;
      if (nk < KT) load_tile(nk % STAGES, nk * BK);
      cp_async_commit();
    }
    const double* as = As + (kt % STAGES) * A_STAGE + a_off;
    const double* bs = Bs + (kt % STAGES) * B_STAGE + b_off;
#pragma unroll
    for (int k4 = 0; k4 < BK / 4; ++k4) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = as[i * 8 * LDA_S + k4 * 4];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = bs[k4 * 4 * LDB_S + j * 8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  cp_async_wait<0>();

  // epilogue: lane holds C[row = lane/4][col = 2*(lane%4) + {0,1}] of every 8x8 tile
  const double alpha = g.alpha, beta = g.beta;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = bm0 + wm + i * 8 + (lane >> 2);
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = bn0 + wn + j * 8 + 2 * (lane & 3);
      if (col >= N) continue;
      double* cp = C + (int64_t)row * g.ldc + col;
      double r0 = alpha * acc[i][j][0], r1 = alpha * acc[i][j][1];
      if (col + 1 < N) {
        if (g.vecC) {
          if (beta != 0.0) {
            double2 old = *reinterpret_cast<const double2*>(cp);
            r0 += beta * old.x;
            r1 += beta * old.y;
          }
          *reinterpret_cast<double2*>(cp) = make_double2(r0, r1);
        } else {
          if (beta != 0.0) {
            r0 += beta * cp[0];
            r1 += beta * cp[1];
          }
          cp[0] = r0;
          cp[1] = r1;
        }
      } else {
        if (beta != 0.0) r0 += beta * cp[0];
        cp[0] = r0;
      }
    }
  }
}

// ---- narrow-N kernel: one warp per output row, lanes stride over K ------------------
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

inline bool vec_ok(const void* p, int64_t ld, int64_t stride) {
  return (reinterpret_cast<uintptr_t>(p) % 16 == 0) && (ld % 2 == 0) && (stride % 2 == 0);
}

}  // namespace

int dgemm_skinny(cudaStream_t st, int M, int N, int K, double alpha, const double* A, int64_t lda,
                 int64_t sA, const double* B, int64_t ldb, int64_t sB, double beta, const double* Cin,
                 int64_t ldcin, int64_t sCin, double* C, int64_t ldc, int64_t sC, int batch) {
  if (M <= 0 || N <= 0 || batch <= 0) return 0;
  const int wpb = 8;
  for (int b0 = 0; b0 < batch; b0 += 65535) {
    const int nb = min(65535, batch - b0);
    dim3 grid((M + wpb - 1) / wpb, nb);
    for (int n0 = 0; n0 < N; n0 += 4) {
      skinny_kernel<4><<<grid, wpb * 32, 0, st>>>(M, N, K, alpha, A + b0 * sA, lda, sA, B + b0 * sB, ldb, sB,
                                                  beta, Cin ? Cin + b0 * sCin : nullptr, ldcin, sCin,
                                                  C + b0 * sC, ldc, sC, n0);
    }
  }
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
  static bool configured = false;
  if (!configured) {
    HPS_CUDA(cudaFuncSetAttribute(gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    configured = true;
  }
  GemmArgs g;
  g.M = M; g.N = N; g.K = K; g.alpha = alpha; g.beta = beta;
  g.lda = lda; g.sA = sA; g.ldb = ldb; g.sB = sB; g.ldc = ldc; g.sC = sC;
  g.vecA = vec_ok(A, lda, sA); g.vecB = vec_ok(B, ldb, sB); g.vecC = vec_ok(C, ldc, sC);
  prof_begin(PROF_GEMM, st, 2.0 * M * N * (double)K * batch);
  for (int b0 = 0; b0 < batch; b0 += 65535) {
    const int nb = min(65535, batch - b0);
    g.A = A + (int64_t)b0 * sA; g.B = B + (int64_t)b0 * sB; g.C = C + (int64_t)b0 * sC;
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, nb);
    gemm_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(g);
  }
  prof_end(PROF_GEMM, st);
  HPS_LAUNCH_CHECK("gemm_kernel");
  return 0;
}

}  // namespace hps

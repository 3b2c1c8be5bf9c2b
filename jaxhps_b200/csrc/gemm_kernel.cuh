// Templated FP64 DMMA GEMM kernel for sm_100a (included by gemm.cu and tools/gemm_lab.cu).
//
// C[b] = alpha * A[b] (MxK) * B[b] (KxN) + beta * C[b]; all row-major.
//
// DMMA (mma.sync.m8n8k4.f64) is the only FP64 tensor shape the B200 executes natively (the
// m16n8k{4,8,16} PTX shapes lower to the same DMMA.8x8x4 SASS; tcgen05 has no f64 kind).
// Measured on B200: DMMA issue peak 37.0 TFLOP/s, cuBLAS DGEMM 35.4 (profiles/r01_fp64_probe.txt).
//
// Layout: CTA tile (WM*WARPS_M) x (WN*WARPS_N), K step BK, STAGES-deep cp.async (LDGSTS) ring.
// Each warp owns a WM x WN accumulator made of 8x8 DMMA tiles.  Shared-memory leading
// dimensions are == 4 (mod 16) doubles, which makes both fragment loads
// (A: lane -> (row=lane/4, k=lane%4); B: lane -> (k=lane%4, col=lane/4)) conflict-free per
// half-warp.  The epilogue stages the tile through shared memory so that C is written
// (read-modify-written when beta != 0) in full contiguous rows with all loads of a batch in
// flight before the first store.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace hps {
namespace gemmk {

struct GemmArgs {
  int M, N, K;
  double alpha, beta;
  const double* A; int64_t lda, sA;
  const double* B; int64_t ldb, sB;
  double* C; int64_t ldc, sC;
  int vecA, vecB, vecC;  // 16-byte vector access allowed for that operand
  int raster = 0;        // > 1: grouped CTA launch order with this many tile rows per group
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

template <int WM_, int WN_, int WARPS_M_, int WARPS_N_, int BK_, int STAGES_, int MIN_CTAS_>
struct Config {
  static constexpr int WM = WM_, WN = WN_, WARPS_M = WARPS_M_, WARPS_N = WARPS_N_;
  static constexpr int BK = BK_, STAGES = STAGES_, MIN_CTAS = MIN_CTAS_;
  static constexpr int BM = WM * WARPS_M, BN = WN * WARPS_N;
  static constexpr int NWARPS = WARPS_M * WARPS_N, THREADS = 32 * NWARPS;
  static constexpr int MI = WM / 8, NJ = WN / 8;
  static constexpr int LDA_S = BK + 4, LDB_S = BN + 4, LDC_S = BN + 2;
  static constexpr int A_STAGE = BM * LDA_S, B_STAGE = BK * LDB_S;
  static constexpr int PIPE_DOUBLES = STAGES * (A_STAGE + B_STAGE);
  static constexpr int EPI_DOUBLES = BM * LDC_S;
  static constexpr size_t SMEM_BYTES = sizeof(double) * (PIPE_DOUBLES > EPI_DOUBLES ? PIPE_DOUBLES : EPI_DOUBLES);
  static constexpr int A_CHUNKS = BM * BK / 2, B_CHUNKS = BK * BN / 2;
  static_assert(A_CHUNKS % THREADS == 0 && B_CHUNKS % THREADS == 0, "tile loads must divide evenly");
  static_assert(BK % 4 == 0 && WM % 8 == 0 && WN % 8 == 0 && BN % 2 == 0, "shape");
};

template <class Cfg>
__device__ __forceinline__ void epilogue(const GemmArgs& g, double* smem, double (&acc)[Cfg::MI][Cfg::NJ][2], double* C,
                                         int bm0, int bn0, int wm, int wn, int warp, int lane) {
  constexpr int BM = Cfg::BM, BN = Cfg::BN, LDC_S = Cfg::LDC_S, MI = Cfg::MI, NJ = Cfg::NJ;
  const int M = g.M, N = g.N;
  // ---- epilogue through shared memory ----
  __syncthreads();
  double* Cs = smem;
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int r = wm + i * 8 + (lane >> 2), c = wn + j * 8 + 2 * (lane & 3);
      *reinterpret_cast<double2*>(Cs + r * LDC_S + c) = make_double2(acc[i][j][0], acc[i][j][1]);
    }
  __syncthreads();
  const double alpha = g.alpha, beta = g.beta;
  const int rows_here = min(BM, M - bm0), cols_here = min(BN, N - bn0);
  if (g.vecC && cols_here == BN && (BN % 128 == 0 || BN == 64)) {
    constexpr int LANES_PER_ROW = (BN >= 128) ? 32 : 16;       // lanes covering one row pass
    constexpr int ROWS_PER_WARP = 32 / LANES_PER_ROW;          // rows a warp touches per pass
    constexpr int COL_PASSES = (BN >= 128) ? BN / 128 : 1;
    constexpr int RB = 4;                                      // row passes in flight
    const int lrow = lane / LANES_PER_ROW, lcol = (lane % LANES_PER_ROW) * 4;
    constexpr int ROW_STEP = Cfg::NWARPS * ROWS_PER_WARP;
    for (int rb = warp * ROWS_PER_WARP + lrow; rb < rows_here; rb += ROW_STEP * RB) {
#pragma unroll
      for (int cp_ = 0; cp_ < COL_PASSES; ++cp_) {
        const int c0 = cp_ * 128 + lcol;
        double2 o0[RB], o1[RB];
        if (beta != 0.0) {
#pragma unroll
          for (int t = 0; t < RB; ++t) {
            const int r = rb + t * ROW_STEP;
            if (r < rows_here) {
              const double* cp = C + (int64_t)(bm0 + r) * g.ldc + bn0 + c0;
              o0[t] = *reinterpret_cast<const double2*>(cp);
              o1[t] = *reinterpret_cast<const double2*>(cp + 2);
            }
          }
        }
#pragma unroll
        for (int t = 0; t < RB; ++t) {
          const int r = rb + t * ROW_STEP;
          if (r < rows_here) {
            const double* sp = Cs + r * LDC_S + c0;
            const double2 a0 = *reinterpret_cast<const double2*>(sp), a1 = *reinterpret_cast<const double2*>(sp + 2);
            double2 r0 = make_double2(alpha * a0.x, alpha * a0.y), r1 = make_double2(alpha * a1.x, alpha * a1.y);
            if (beta != 0.0) {
              r0.x = fma(beta, o0[t].x, r0.x); r0.y = fma(beta, o0[t].y, r0.y);
              r1.x = fma(beta, o1[t].x, r1.x); r1.y = fma(beta, o1[t].y, r1.y);
            }
            double* cp = C + (int64_t)(bm0 + r) * g.ldc + bn0 + c0;
            *reinterpret_cast<double2*>(cp) = r0;
            *reinterpret_cast<double2*>(cp + 2) = r1;
          }
        }
      }
    }
  } else {
    // ragged or unaligned tile: scalar, still row-contiguous across the warp
    for (int r = warp; r < rows_here; r += Cfg::NWARPS) {
      double* crow = C + (int64_t)(bm0 + r) * g.ldc + bn0;
      for (int c = lane; c < cols_here; c += 32) {
        double v = alpha * Cs[r * LDC_S + c];
        if (beta != 0.0) v = fma(beta, crow[c], v);
        crow[c] = v;
      }
    }
  }
}

// ---- mbarrier helpers ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_cp_async_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(a),
      "r"(parity)
      : "memory");
}

// The product kernel.  Per-stage full/empty mbarriers instead of a CTA barrier per K step: every thread
// signals full[s] through cp.async.mbarrier.arrive.noinc (fires when its copies have landed), every warp
// signals empty[s] when it has consumed the stage, so warps may drift a stage apart.  The cp.async address
// arithmetic is hoisted out of the K loop: interior tiles (no M/N edge, 16-byte aligned operands) keep ONE
// running source pointer per operand and thread — the copies of a thread differ by compile-time row
// strides — and issue full K steps without predicates; edge tiles and the ragged last K step take the
// generic path (an interior step is 239 instructions: 64 DMMA, 32 LDS, 6 LDGSTS).
// g.raster > 1 re-maps the launch order so that co-resident CTAs cover a (raster x ~296/raster) block of
// tiles instead of two full tile rows: the wave's operand footprint drops from all of B to a few panels of
// A and B (L2 / DRAM re-reads, see profiles/r02_ncu_gemm_summary.txt).
template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MIN_CTAS) gemm_kernel_hoist(GemmArgs g) {
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, STAGES = Cfg::STAGES, THREADS = Cfg::THREADS;
  constexpr int LDA_S = Cfg::LDA_S, LDB_S = Cfg::LDB_S;
  constexpr int MI = Cfg::MI, NJ = Cfg::NJ;
  constexpr int A_CH = BK / 2, B_CH = BN / 2;
  constexpr int NA = Cfg::A_CHUNKS / THREADS, NB = Cfg::B_CHUNKS / THREADS;
  constexpr int A_ROWS = THREADS / A_CH, B_ROWS = THREADS / B_CH;  // rows covered by one pass of the CTA
  static_assert(THREADS % A_CH == 0 && THREADS % B_CH == 0, "a thread keeps its column in every pass");
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  double* As = smem;
  double* Bs = smem + STAGES * Cfg::A_STAGE;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = (warp / Cfg::WARPS_N) * Cfg::WM, wn = (warp % Cfg::WARPS_N) * Cfg::WN;
  int bx = blockIdx.x, by = blockIdx.y;
  if (g.raster > 1) {  // grouped launch order: `raster` tile rows at a time, M fastest inside a group
    const int nx = gridDim.x, ny = gridDim.y;
    const int lin = by * nx + bx;
    const int grp = lin / (g.raster * nx);
    const int first = grp * g.raster;
    const int rows = min(g.raster, ny - first);
    const int in = lin - first * nx;
    by = first + in % rows;
    bx = in / rows;
  }
  const int bm0 = by * BM, bn0 = bx * BN;
  const int64_t batch = blockIdx.z;
  const double* __restrict__ A = g.A + batch * g.sA;
  const double* B = g.B + batch * g.sB;
  double* C = g.C + batch * g.sC;
  const int M = g.M, N = g.N, K = g.K;
  const bool interior = (bm0 + BM <= M) && (bn0 + BN <= N) && g.vecA && g.vecB;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], THREADS);
      mbar_init(&empty_bar[s], Cfg::NWARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  // running sources of this thread's first copy of each operand; pass i adds i * A_ROWS (B_ROWS) rows
  const int a_dst = (tid / A_CH) * LDA_S + (tid % A_CH) * 2;
  const int b_dst = (tid / B_CH) * LDB_S + (tid % B_CH) * 2;
  const double* a_src = A + (int64_t)(bm0 + tid / A_CH) * g.lda + (tid % A_CH) * 2;
  const double* b_src = B + (int64_t)(tid / B_CH) * g.ldb + bn0 + (tid % B_CH) * 2;
  const int64_t a_pass = (int64_t)A_ROWS * g.lda, b_pass = (int64_t)B_ROWS * g.ldb, b_step = (int64_t)BK * g.ldb;

  auto load_generic = [&](int stage, int k0) {
    double* as = As + stage * Cfg::A_STAGE;
    double* bs = Bs + stage * Cfg::B_STAGE;
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      const int c = tid + i * THREADS;
      const int r = c / A_CH, kc = (c % A_CH) * 2;
      const int gr = bm0 + r, gk = k0 + kc;
      double* dst = as + r * LDA_S + kc;
      const int valid = (gr < M) ? max(0, min(2, K - gk)) : 0;
      const double* src = valid ? (A + (int64_t)gr * g.lda + gk) : A;
      if (g.vecA) {
        cp_async16(dst, src, valid * 8);
      } else {
        cp_async8(dst, src, valid >= 1 ? 8 : 0);
        cp_async8(dst + 1, valid >= 2 ? src + 1 : A, valid >= 2 ? 8 : 0);
      }
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int c = tid + i * THREADS;
      const int r = c / B_CH, nc = (c % B_CH) * 2;
      const int gk = k0 + r, gn = bn0 + nc;
      double* dst = bs + r * LDB_S + nc;
      const int valid = (gk < K) ? max(0, min(2, N - gn)) : 0;
      const double* src = valid ? (B + (int64_t)gk * g.ldb + gn) : B;
      if (g.vecB) {
        cp_async16(dst, src, valid * 8);
      } else {
        cp_async8(dst, src, valid >= 1 ? 8 : 0);
        cp_async8(dst + 1, valid >= 2 ? src + 1 : B, valid >= 2 ? 8 : 0);
      }
    }
    mbar_cp_async_arrive(&full_bar[stage]);
  };
  // K steps are issued in order 0, 1, 2, ...: the fast path advances the running pointers itself
  auto load_tile = [&](int stage, int k0) {
    if (interior && k0 + BK <= K) {
      double* as = As + stage * Cfg::A_STAGE + a_dst;
      double* bs = Bs + stage * Cfg::B_STAGE + b_dst;
#pragma unroll
      for (int i = 0; i < NA; ++i) cp_async16(as + i * A_ROWS * LDA_S, a_src + i * a_pass, 16);
#pragma unroll
      for (int i = 0; i < NB; ++i) cp_async16(bs + i * B_ROWS * LDB_S, b_src + i * b_pass, 16);
      a_src += BK;
      b_src += b_step;
      mbar_cp_async_arrive(&full_bar[stage]);
    } else {
      load_generic(stage, k0);
    }
  };

  double acc[MI][NJ][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int KT = (K + BK - 1) / BK;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s)
    if (s < KT) load_tile(s, s * BK);

  const int a_off = (wm + (lane >> 2)) * LDA_S + (lane & 3);
  const int b_off = (lane & 3) * LDB_S + wn + (lane >> 2);

  for (int kt = 0; kt < KT; ++kt) {
    const int st = kt % STAGES;
    mbar_wait(&full_bar[st], (unsigned)(kt / STAGES) & 1u);
    const double* as = As + st * Cfg::A_STAGE + a_off;
    const double* bs = Bs + st * Cfg::B_STAGE + b_off;
    double a[2][MI], b[2][NJ];
#pragma unroll
    for (int i = 0; i < MI; ++i) a[0][i] = as[i * 8 * LDA_S];
#pragma unroll
    for (int j = 0; j < NJ; ++j) b[0][j] = bs[j * 8];
#pragma unroll
    for (int k4 = 0; k4 < BK / 4; ++k4) {
      const int cur = k4 & 1, nxt = cur ^ 1;
      if (k4 + 1 < BK / 4) {
#pragma unroll
        for (int i = 0; i < MI; ++i) a[nxt][i] = as[i * 8 * LDA_S + (k4 + 1) * 4];
#pragma unroll
        for (int j = 0; j < NJ; ++j) b[nxt][j] = bs[(k4 + 1) * 4 * LDB_S + j * 8];
      }
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) dmma(acc[i][j][0], acc[i][j][1], a[cur][i], b[cur][j]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[st]);
    const int nk = kt + STAGES - 1;
    if (nk < KT) {
      const int sp = nk % STAGES;
      if (nk >= STAGES) mbar_wait(&empty_bar[sp], (unsigned)(nk / STAGES - 1) & 1u);
      load_tile(sp, nk * BK);
    }
  }
  cp_async_wait<0>();
  epilogue<Cfg>(g, smem, acc, C, bm0, bn0, wm, wn, warp, lane);
}


}  // namespace gemmk
}  // namespace hps

// Lab variants of the DMMA GEMM kept for measurements only (tools/gemm_lab.cu); the product kernel is
// gemm_kernel_hoist in jaxhps_b200/csrc/gemm_kernel.cuh.
//   gemm_kernel      v1: CTA barrier per K step
//   gemm_kernel_mb   per-stage mbarriers, per-step address arithmetic
//   gemm_kernel_tma  mbarriers + cp.async.bulk row copies (TMA engine, non-tensor)
#pragma once
#include "../jaxhps_b200/csrc/gemm_kernel.cuh"

namespace hps {
namespace gemmk {

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MIN_CTAS) gemm_kernel(GemmArgs g) {
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, STAGES = Cfg::STAGES, THREADS = Cfg::THREADS;
  constexpr int LDA_S = Cfg::LDA_S, LDB_S = Cfg::LDB_S, LDC_S = Cfg::LDC_S;
  constexpr int MI = Cfg::MI, NJ = Cfg::NJ;
  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = smem + STAGES * Cfg::A_STAGE;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = (warp / Cfg::WARPS_N) * Cfg::WM, wn = (warp % Cfg::WARPS_N) * Cfg::WN;
  const int bm0 = blockIdx.y * BM, bn0 = blockIdx.x * BN;
  const int64_t batch = blockIdx.z;
  const double* __restrict__ A = g.A + batch * g.sA;
  const double* B = g.B + batch * g.sB;  // may alias C (in-place products with M <= BM)
  double* C = g.C + batch * g.sC;
  const int M = g.M, N = g.N, K = g.K;

  auto load_tile = [&](int stage, int k0) {
    double* as = As + stage * Cfg::A_STAGE;
    double* bs = Bs + stage * Cfg::B_STAGE;
    constexpr int A_CH = BK / 2, B_CH = BN / 2;
#pragma unroll
    for (int i = 0; i < Cfg::A_CHUNKS / THREADS; ++i) {
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
    for (int i = 0; i < Cfg::B_CHUNKS / THREADS; ++i) {
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
  };

  double acc[MI][NJ][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

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
    const double* as = As + (kt % STAGES) * Cfg::A_STAGE + a_off;
    const double* bs = Bs + (kt % STAGES) * Cfg::B_STAGE + b_off;
    // fragments are double-buffered in registers: the loads of step k4+1 are issued before the
    // DMMAs of step k4
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
  }
  cp_async_wait<0>();

  epilogue<Cfg>(g, smem, acc, C, bm0, bn0, wm, wn, warp, lane);
}


template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MIN_CTAS) gemm_kernel_mb(GemmArgs g) {
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, STAGES = Cfg::STAGES, THREADS = Cfg::THREADS;
  constexpr int LDA_S = Cfg::LDA_S, LDB_S = Cfg::LDB_S;
  constexpr int MI = Cfg::MI, NJ = Cfg::NJ;
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  double* As = smem;
  double* Bs = smem + STAGES * Cfg::A_STAGE;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = (warp / Cfg::WARPS_N) * Cfg::WM, wn = (warp % Cfg::WARPS_N) * Cfg::WN;
  const int bm0 = blockIdx.y * BM, bn0 = blockIdx.x * BN;
  const int64_t batch = blockIdx.z;
  const double* __restrict__ A = g.A + batch * g.sA;
  const double* B = g.B + batch * g.sB;
  double* C = g.C + batch * g.sC;
  const int M = g.M, N = g.N, K = g.K;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], THREADS);
      mbar_init(&empty_bar[s], Cfg::NWARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  auto load_tile = [&](int stage, int k0) {
    double* as = As + stage * Cfg::A_STAGE;
    double* bs = Bs + stage * Cfg::B_STAGE;
    constexpr int A_CH = BK / 2, B_CH = BN / 2;
#pragma unroll
    for (int i = 0; i < Cfg::A_CHUNKS / THREADS; ++i) {
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
    for (int i = 0; i < Cfg::B_CHUNKS / THREADS; ++i) {
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
    // this warp is done with stage st
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[st]);
    // refill the stage consumed one iteration ago with tile kt + STAGES - 1
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


// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP): one instruction moves a whole tile row ----
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"(
                   (unsigned)__cvta_generic_to_shared(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   (unsigned)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}

// Same pipeline as gemm_kernel_mb, but interior K steps of interior tiles are staged by the TMA
// engine: BM + BK threads each issue ONE bulk copy (a 128-byte row of the A tile or a 512-byte row
// of the B tile) that signals the stage's full barrier with complete_tx; ragged / unaligned steps
// fall back to the LDGSTS path.  Every thread arrives on the full barrier in both paths.
template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MIN_CTAS) gemm_kernel_tma(GemmArgs g) {
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, STAGES = Cfg::STAGES, THREADS = Cfg::THREADS;
  constexpr int LDA_S = Cfg::LDA_S, LDB_S = Cfg::LDB_S;
  constexpr int MI = Cfg::MI, NJ = Cfg::NJ;
  static_assert(BM + BK <= THREADS, "one bulk copy per thread");
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  double* As = smem;
  double* Bs = smem + STAGES * Cfg::A_STAGE;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = (warp / Cfg::WARPS_N) * Cfg::WM, wn = (warp % Cfg::WARPS_N) * Cfg::WN;
  const int bm0 = blockIdx.y * BM, bn0 = blockIdx.x * BN;
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

  auto load_tile = [&](int stage, int k0) {
    double* as = As + stage * Cfg::A_STAGE;
    double* bs = Bs + stage * Cfg::B_STAGE;
    if (interior && k0 + BK <= K) {
      if (tid < BM) {
        mbar_arrive_expect_tx(&full_bar[stage], BK * 8);
        bulk_copy_g2s(as + tid * LDA_S, A + (int64_t)(bm0 + tid) * g.lda + k0, BK * 8, &full_bar[stage]);
      } else if (tid < BM + BK) {
        const int r = tid - BM;
        mbar_arrive_expect_tx(&full_bar[stage], BN * 8);
        bulk_copy_g2s(bs + r * LDB_S, B + (int64_t)(k0 + r) * g.ldb + bn0, BN * 8, &full_bar[stage]);
      } else {
        mbar_arrive(&full_bar[stage]);
      }
      return;
    }
    constexpr int A_CH = BK / 2, B_CH = BN / 2;
#pragma unroll
    for (int i = 0; i < Cfg::A_CHUNKS / THREADS; ++i) {
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
    for (int i = 0; i < Cfg::B_CHUNKS / THREADS; ++i) {
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

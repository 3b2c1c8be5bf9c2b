// Phase timing of the multi-CTA block-column kernel (developer tool; not part of the product).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DHPS_BC_TIMING -I jaxhps_b200/csrc tools/panel_lab.cu \
//        jaxhps_b200/csrc/{gemm,leaf,merge,adaptive,interp,api}.o -o tools/panel_lab
// (includes lu.cu itself, so the library's lu.o is NOT linked.)   usage: tools/panel_lab [n] [block columns] [diagonal boost]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "../jaxhps_b200/csrc/lu.cu"
using namespace hps;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__global__ void fill(double* p, size_t n, int ld, double diag) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) { unsigned x = (unsigned)(i * 2654435761u) ^ 77u; x ^= x >> 13; x *= 0x5bd1e995; x ^= x >> 15;
    p[i] = ((x & 0xffff) / 65536.0) - 0.5 + ((i / ld) == (i % ld) ? diag : 0.0); }
}
int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 19200, nbc = argc > 2 ? atoi(argv[2]) : 3;
  const double diag = argc > 3 ? atof(argv[3]) : 40.0;  // 0: random pivots; large: the pivot is always the diagonal (HPS merges)
  double* A; CK(cudaMalloc(&A, (size_t)n * n * 8));
  fill<<<2048, 256>>>(A, (size_t)n * n, n, diag);
  const size_t wsb = lu_workspace_bytes(1, n);
  void* ws; CK(cudaMalloc(&ws, wsb)); CK(cudaMemset(ws, 0, wsb));
  int* info; CK(cudaMalloc(&info, 4)); CK(cudaMemset(info, 0, 4));
  Arena ar(ws, wsb); LuWorkspace w; if (!carve(ar, 1, n, w)) { printf("carve failed\n"); return 1; }
  if (configure_lu_kernels()) { printf("configure failed: %s\n", last_error().c_str()); return 1; }
  long long* tim; const size_t tn = (size_t)BC_MAX_G * NB * 8; CK(cudaMalloc(&tim, tn * 8));
  CK(cudaMemcpyToSymbol(g_bc_timing, &tim, sizeof(tim)));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[5] = {"argmax+reduce", "publish->hdr0 ", "all headers   ", "fetch pivot   ", "update        "};
  for (int b = 0; b < nbc; ++b) {
    const int j = b * NB; bool done = false;
    CK(cudaMemset(tim, 0, tn * 8));
    cudaEventRecord(e0, st);
    if (launch_blockcol(st, 1, n, A, n, 0, j, NB, w, info, done)) { printf("launch failed: %s\n", last_error().c_str()); return 1; }
    cudaEventRecord(e1, st); CK(cudaStreamSynchronize(st));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const int G = (n - j + BC_ROWS - 1) / BC_ROWS;
    std::vector<long long> h(tn); CK(cudaMemcpy(h.data(), tim, tn * 8, cudaMemcpyDeviceToHost));
    printf("block column %d: done=%d G=%d  %.1f us  (%.2f us per column)\n", b, (int)done, G, ms * 1e3, ms * 1e3 / NB);
    // per phase: mean over CTAs and columns, and the mean over columns of the MAX over CTAs (cycles)
    for (int ph = 0; ph < 5; ++ph) {
      double mean = 0, mx = 0, m0 = 0; long cnt = 0;
      for (int c = 0; c < NB; ++c) {
        double cmx = 0;
        for (int g = 0; g < G; ++g) {
          const long long* t = &h[((size_t)g * NB + c) * 8];
          const double d = (double)(t[ph + 1] - t[ph]);
          mean += d; ++cnt; cmx = std::max(cmx, d); if (g == 0) m0 += d;
        }
        mx += cmx;
      }
      printf("   %s mean %7.0f cyc   max-over-CTAs %7.0f cyc   CTA0 %7.0f cyc\n", names[ph], mean / cnt, mx / NB, m0 / NB);
    }
    double per = 0; for (int c = 1; c < NB; ++c) per += (double)(h[(size_t)c * 8] - h[(size_t)(c - 1) * 8]);
    printf("   CTA0 column period %7.0f cyc (incl. trailing updates every 32 columns)\n", per / (NB - 1));
    // columns inside an inner panel only
    double pin = 0; int np = 0; for (int c = 1; c < NB; ++c) if (c % 32) { pin += (double)(h[(size_t)c * 8] - h[(size_t)(c - 1) * 8]); ++np; }
    printf("   CTA0 column period inside inner panels %7.0f cyc\n", pin / np);
  }
  int hinfo; CK(cudaMemcpy(&hinfo, info, 4, cudaMemcpyDeviceToHost)); printf("info=%d\n", hinfo);
  return 0;
}

// Tile-configuration lab for the DMMA GEMM (developer tool; not part of the product).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "gemm_lab_kernels.cuh"
using namespace hps::gemmk;
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s line %d\n",cudaGetErrorString(e),__LINE__);exit(1);}}while(0)
__global__ void fill(double* p, size_t n, unsigned seed){ size_t i=blockIdx.x*(size_t)blockDim.x+threadIdx.x; size_t st=(size_t)gridDim.x*blockDim.x; for(;i<n;i+=st){ unsigned x=(unsigned)(i*2654435761u)^seed; x^=x>>13; x*=0x5bd1e995; x^=x>>15; p[i]=((x&0xffff)/65536.0)-0.5; } }
__global__ void checksum(const double* p, size_t n, double* out){ __shared__ double s[256]; double a=0; size_t i=blockIdx.x*(size_t)blockDim.x+threadIdx.x; size_t st=(size_t)gridDim.x*blockDim.x; for(;i<n;i+=st) a+=p[i]*(1+(i%7)); s[threadIdx.x]=a; __syncthreads(); for(int o=128;o>0;o>>=1){ if(threadIdx.x<o) s[threadIdx.x]+=s[threadIdx.x+o]; __syncthreads(); } if(threadIdx.x==0) atomicAdd(out,s[0]); }
template<class Cfg, int MB> auto pick(){
  if constexpr (MB == 3) return gemm_kernel_hoist<Cfg>;
  else if constexpr (MB == 2) return gemm_kernel_tma<Cfg>;
  else if constexpr (MB == 1) return gemm_kernel_mb<Cfg>;
  else return gemm_kernel<Cfg>;
}
template<class Cfg, int MB=0> double run(const char* name, GemmArgs g, int batch, double* C0, size_t csz, double* d_sum, int raster=0){
  g.raster = raster;
  auto kern = pick<Cfg, MB>();
  static bool conf=false; if(!conf){ CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,(int)Cfg::SMEM_BYTES)); conf=true; }
  dim3 grid((g.N+Cfg::BN-1)/Cfg::BN,(g.M+Cfg::BM-1)/Cfg::BM,batch);
  int occ=0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Cfg::THREADS, Cfg::SMEM_BYTES);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best=1e9;
  for(int it=0; it<4; ++it){
    CK(cudaMemcpy(g.C, C0, csz*sizeof(double), cudaMemcpyDeviceToDevice));
    cudaEventRecord(e0); kern<<<grid,Cfg::THREADS,Cfg::SMEM_BYTES>>>(g); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms,e0,e1); if(it>0 && ms<best) best=ms;
  }
  CK(cudaGetLastError());
  CK(cudaMemset(d_sum,0,8)); checksum<<<1024,256>>>(g.C,csz,d_sum); double h; CK(cudaMemcpy(&h,d_sum,8,cudaMemcpyDeviceToHost));
  double fl=2.0*g.M*g.N*(double)g.K*batch;
  printf("  %-34s occ=%d smem=%6zu  %8.3f ms  %6.2f TF/s  checksum %.10e\n", name, occ, Cfg::SMEM_BYTES, best, fl/best*1e-9, h);
  return h;
}
int main(){
  struct Shape{int M,N,K,batch; double beta;};
  std::vector<Shape> shapes={{8192,8192,8192,1,0.0},{15360,15360,128,1,1.0},{19072,2304,128,1,1.0},{872,872,128,512,1.0},{1000,600,728,512,0.0},{19200,4800,9600,1,1.0}};
  double* d_sum; CK(cudaMalloc(&d_sum,8));
  for(auto s: shapes){
    size_t asz=(size_t)s.M*s.K*s.batch, bsz=(size_t)s.K*s.N*s.batch, csz=(size_t)s.M*s.N*s.batch;
    double *A,*B,*C,*C0; CK(cudaMalloc(&A,asz*8)); CK(cudaMalloc(&B,bsz*8)); CK(cudaMalloc(&C,csz*8)); CK(cudaMalloc(&C0,csz*8));
    fill<<<2048,256>>>(A,asz,1); fill<<<2048,256>>>(B,bsz,2); fill<<<2048,256>>>(C0,csz,3); CK(cudaDeviceSynchronize());
    GemmArgs g; g.M=s.M; g.N=s.N; g.K=s.K; g.alpha=-1.0; g.beta=s.beta; g.A=A; g.lda=s.K; g.sA=(int64_t)s.M*s.K; g.B=B; g.ldb=s.N; g.sB=(int64_t)s.K*s.N; g.C=C; g.ldc=s.N; g.sC=(int64_t)s.M*s.N; g.vecA=g.vecB=g.vecC=1;
    printf("shape M=%d N=%d K=%d batch=%d beta=%g\n",s.M,s.N,s.K,s.batch,s.beta);
    //                 WM  WN  WMs WNs BK  ST  minCTA
    run<Config<32, 32, 4, 2, 16, 3, 2>, 1>("G  mb  128x64 8w(32x32) bk16 s3 x2cta", g, s.batch, C0, csz, d_sum);
    run<Config<32, 32, 4, 2, 16, 3, 2>, 3>("G  hoist 128x64 8w(32x32) bk16 s3 x2cta [product]", g, s.batch, C0, csz, d_sum);
    run<Config<32, 32, 4, 2, 16, 3, 2>, 3>("G  hoist raster 8", g, s.batch, C0, csz, d_sum, 8);
    run<Config<32, 32, 4, 2, 16, 3, 2>, 3>("G  hoist raster 12", g, s.batch, C0, csz, d_sum, 12);
    run<Config<32, 32, 4, 2, 16, 3, 2>, 3>("G  hoist raster 16", g, s.batch, C0, csz, d_sum, 16);
    run<Config<64, 32, 2, 4, 16, 3, 1>, 3>("X  hoist 128x128 8w(64x32) bk16 s3 x1cta", g, s.batch, C0, csz, d_sum);
    run<Config<64, 32, 2, 4, 16, 4, 1>, 3>("X4 hoist 128x128 8w(64x32) bk16 s4 x1cta raster 8", g, s.batch, C0, csz, d_sum, 8);
    run<Config<64, 32, 2, 2, 16, 4, 2>, 3>("W4 hoist 128x64 4w(64x32) bk16 s4 x2cta raster 12", g, s.batch, C0, csz, d_sum, 12);
    run<Config<64, 32, 2, 2, 16, 3, 3>, 3>("W3 hoist 128x64 4w(64x32) bk16 s3 x3cta", g, s.batch, C0, csz, d_sum);
    run<Config<32, 32, 4, 2, 16, 4, 2>, 3>("G4 hoist 128x64 8w bk16 s4 x2cta", g, s.batch, C0, csz, d_sum);
    run<Config<32, 32, 4, 2, 32, 2, 2>, 3>("G2 hoist 128x64 8w bk32 s2 x2cta", g, s.batch, C0, csz, d_sum);
    run<Config<32, 32, 4, 4, 16, 3, 1>, 3>("A1 hoist 128x128 16w bk16 s3", g, s.batch, C0, csz, d_sum);
    run<Config<32, 32, 4, 4, 32, 3, 1>, 3>("A  hoist 128x128 16w bk32 s3", g, s.batch, C0, csz, d_sum);
    run<Config<64, 32, 2, 2, 16, 3, 2>, 3>("W  hoist 128x64 4w(64x32) bk16 s3 x2cta", g, s.batch, C0, csz, d_sum);
    run<Config<32, 32, 4, 2, 16, 3, 2>, 2>("G  tma 128x64 8w(32x32) bk16 s3 x2cta", g, s.batch, C0, csz, d_sum);
    run<Config<32, 32, 4, 4, 32, 3, 1>, 1>("A  mb  128x128 16w bk32 s3", g, s.batch, C0, csz, d_sum);
    run<Config<32, 32, 4, 4, 32, 3, 1>, 2>("A  tma 128x128 16w bk32 s3", g, s.batch, C0, csz, d_sum);
    run<Config<32, 32, 4, 2, 32, 2, 2>, 2>("G2 tma 128x64 8w bk32 s2 x2cta", g, s.batch, C0, csz, d_sum);
    cudaFree(A); cudaFree(B); cudaFree(C); cudaFree(C0);
  }
  return 0;
}

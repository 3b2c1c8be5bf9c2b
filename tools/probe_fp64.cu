// FP64 throughput probe for B200 (sm_100a): DFMA vs DMMA shapes. Dev tool, not product.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__);return 1;}}while(0)

__global__ void k_dfma(double* out, int iters){
  double a[16]; double x = threadIdx.x*1e-9, y = 1.0000001;
  #pragma unroll
  for(int i=0;i<16;i++) a[i]=i;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<16;i++) a[i]=fma(a[i],y,x);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<16;i++) s+=a[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
// m8n8k4: A 1 reg, B 1 reg, C 2 regs
template<int NACC>
__global__ void k_dmma884(double* out, int iters){
  double c[NACC][2]; double a = threadIdx.x*1e-9, b = 1.0000001;
  #pragma unroll
  for(int i=0;i<NACC;i++){c[i][0]=i;c[i][1]=-i;}
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<NACC;i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i][0]+c[i][1];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
// m16n8k8: A 4 regs, B 2 regs, C 4 regs
template<int NACC>
__global__ void k_dmma1688(double* out, int iters){
  double c[NACC][4]; double a0=threadIdx.x*1e-9,a1=a0+1,a2=a0+2,a3=a0+3, b0=1.0000001,b1=0.999999;
  #pragma unroll
  for(int i=0;i<NACC;i++){c[i][0]=i;c[i][1]=-i;c[i][2]=i;c[i][3]=1;}
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<NACC;i++)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a0),"d"(a1),"d"(a2),"d"(a3), "d"(b0),"d"(b1));
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i][0]+c[i][1]+c[i][2]+c[i][3];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
// m16n8k16: A 8 regs, B 4 regs, C 4 regs
template<int NACC>
__global__ void k_dmma16816(double* out, int iters){
  double c[NACC][4]; double a[8], b[4];
  #pragma unroll
  for(int i=0;i<8;i++) a[i]=threadIdx.x*1e-9+i;
  #pragma unroll
  for(int i=0;i<4;i++) b[i]=1.0+1e-7*i;
  #pragma unroll
  for(int i=0;i<NACC;i++){c[i][0]=i;c[i][1]=-i;c[i][2]=i;c[i][3]=1;}
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<NACC;i++)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
        : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
        : "d"(a[0]),"d"(a[1]),"d"(a[2]),"d"(a[3]),"d"(a[4]),"d"(a[5]),"d"(a[6]),"d"(a[7]), "d"(b[0]),"d"(b[1]),"d"(b[2]),"d"(b[3]));
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i][0]+c[i][1]+c[i][2]+c[i][3];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

template<typename F> float timeit(F f){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1); return ms;
}
int main(){
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr,0));
  printf("device %s sms=%d clock=%d kHz\n", pr.name, pr.multiProcessorCount, pr.clockRate);
  int nsm=pr.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double)*nsm*8*1024));
  int iters=20000;
  for(int wpsm : {4,8,16,32}){
    int threads=256; int blocks=nsm*wpsm*32/threads;
    float ms=timeit([&]{k_dfma<<<blocks,threads>>>(out,iters);});
    double fl=2.0*16*iters*(double)blocks*threads;
    printf("DFMA        warps/SM=%2d : %8.2f TFLOP/s (%.3f ms)\n",wpsm,fl/ms*1e-9,ms);
  }
  for(int wpsm : {4,8,16,32}){
    int threads=256; int blocks=nsm*wpsm*32/threads;
    float ms=timeit([&]{k_dmma884<8><<<blocks,threads>>>(out,iters);});
    double fl=2.0*8*8*4*8*iters*(double)blocks*threads/32;
    printf("DMMA m8n8k4   warps/SM=%2d : %8.2f TFLOP/s (%.3f ms)\n",wpsm,fl/ms*1e-9,ms);
  }
  for(int wpsm : {4,8,16,32}){
    int threads=256; int blocks=nsm*wpsm*32/threads;
    float ms=timeit([&]{k_dmma1688<8><<<blocks,threads>>>(out,iters);});
    double fl=2.0*16*8*8*8*iters*(double)blocks*threads/32;
    printf("DMMA m16n8k8  warps/SM=%2d : %8.2f TFLOP/s (%.3f ms)\n",wpsm,fl/ms*1e-9,ms);
  }
  for(int wpsm : {4,8,16,32}){
    int threads=256; int blocks=nsm*wpsm*32/threads;
    float ms=timeit([&]{k_dmma16816<8><<<blocks,threads>>>(out,iters);});
    double fl=2.0*16*8*16*8*iters*(double)blocks*threads/32;
    printf("DMMA m16n8k16 warps/SM=%2d : %8.2f TFLOP/s (%.3f ms)\n",wpsm,fl/ms*1e-9,ms);
  }
  // dependent-chain latency of DMMA (1 accumulator)
  {
    float ms=timeit([&]{k_dmma884<1><<<nsm,32>>>(out,iters);});
    printf("DMMA m8n8k4 dependent latency: %.1f ns/instr\n", ms*1e6/iters);
    ms=timeit([&]{k_dmma16816<1><<<nsm,32>>>(out,iters);});
    printf("DMMA m16n8k16 dependent latency: %.1f ns/instr\n", ms*1e6/iters);
  }
  return 0;
}

// Microbenchmark: FP64 DMMA (mma.sync m8n8k4) vs DFMA issue throughput on sm_100a.
// Used once to choose the SYRK inner loop ("C2 kernel"); not part of the product path.
#include <cstdio>
#include <cuda_runtime.h>
template<int NACC>
__global__ void __launch_bounds__(256) k_dmma(double* out, const double* in, int iters){
  double a = in[threadIdx.x & 31], b = in[(threadIdx.x & 31) + 32];
  double c[NACC][2];
  #pragma unroll
  for(int j=0;j<NACC;j++){c[j][0]=0;c[j][1]=0;}
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int j=0;j<NACC;j++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
  }
  double s=0;
  #pragma unroll
  for(int j=0;j<NACC;j++) s+=c[j][0]+c[j][1];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<int NACC>
__global__ void __launch_bounds__(256) k_dfma(double* out, const double* in, int iters){
  double a = in[threadIdx.x & 31], b = in[(threadIdx.x & 31) + 32];
  double c[NACC];
  #pragma unroll
  for(int j=0;j<NACC;j++) c[j]=j;
  for(int i=0;i<iters;i++){
    #pragma unroll
    for(int j=0;j<NACC;j++) c[j]=fma(a,c[j],b);
  }
  double s=0;
  #pragma unroll
  for(int j=0;j<NACC;j++) s+=c[j];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
template<typename F> float timeit(F f){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1); return ms;
}
int main(){
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double *in,*out; cudaMalloc(&in, 4096); cudaMemset(in,0,4096); cudaMalloc(&out, sizeof(double)*sms*8*256*2);
  const int iters=20000;
  for(int bps=1;bps<=8;bps*=2){
    for (int threads=128; threads<=256; threads*=2){
    int grid=sms*bps;
    float ms=timeit([&]{k_dmma<8><<<grid,threads>>>(out,in,iters);});
    double fl=2.0*256*8*iters*(threads/32)*(double)grid;
    printf("DMMA884 nacc=8 blocks/SM=%d threads=%d: %.3f ms  %.2f TFLOP/s\n",bps,threads,ms,fl/ms*1e-9);
    ms=timeit([&]{k_dfma<16><<<grid,threads>>>(out,in,iters);});
    fl=2.0*16*iters*threads*(double)grid;
    printf("DFMA    nacc=16 blocks/SM=%d threads=%d: %.3f ms  %.2f TFLOP/s\n",bps,threads,ms,fl/ms*1e-9);
    }
  }
  float ms=timeit([&]{k_dmma<2><<<sms*4,256>>>(out,in,iters);});
  printf("DMMA884 nacc=2 4x256: %.2f TFLOP/s\n", 2.0*256*2*iters*8*(double)(sms*4)/ms*1e-9);
  ms=timeit([&]{k_dmma<1><<<sms*4,256>>>(out,in,iters);});
  printf("DMMA884 nacc=1 4x256: %.2f TFLOP/s\n", 2.0*256*1*iters*8*(double)(sms*4)/ms*1e-9);
  printf("cudaGetLastError: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

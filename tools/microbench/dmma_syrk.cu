// Microbenchmark: the reg_pass inner loop (36 lower-triangle DMMA tiles fed by 8 fragment registers) with the
// fragments (a) held in registers, (b) re-read from shared memory every k-step, to separate the DMMA issue ceiling of
// this operand pattern from memory-pipeline effects.  Not part of the product path.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE, int NTILE_LO, int NTILE_HI>
__global__ void __launch_bounds__(128) k(double* out, const double* in, int iters) {
  __shared__ double sm[64 * 68];
  for (int i = threadIdx.x; i < 64 * 68; i += blockDim.x) sm[i] = in[i % 64];
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, kq = lane & 3, warp = threadIdx.x >> 5;
  double acc[36][2];
#pragma unroll
  for (int t = 0; t < 36; ++t) acc[t][0] = acc[t][1] = 0;
  double af[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) af[j] = in[lane + j];
  for (int it = 0; it < iters; ++it) {
    if (MODE == 1) {
      const double* xr = sm + ((4 * ((it + warp) & 15) + kq) * 68 + g);
#pragma unroll
      for (int j = 0; j < 8; ++j) af[j] = xr[8 * j];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) {
        const int t = i * (i + 1) / 2 + j;
        if (t >= NTILE_LO && t < NTILE_HI) dmma(acc[t][0], acc[t][1], af[i], af[j]);
      }
  }
  double s = 0;
#pragma unroll
  for (int t = 0; t < 36; ++t) s += acc[t][0] + acc[t][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double *in, *out; cudaMalloc(&in, 8192); cudaMemset(in, 0, 8192); cudaMalloc(&out, sizeof(double) * sms * 8 * 128);
  const int iters = 4000;
  for (int bps = 1; bps <= 4; ++bps) {
    float ms = timeit([&] { k<0, 0, 36><<<sms * bps, 128>>>(out, in, iters); });
    printf("regs-only 36 tiles  blocks/SM=%d: %.2f TFLOP/s\n", bps, 2.0 * 256 * 36 * iters * 4 * (double)(sms * bps) / ms * 1e-9);
    ms = timeit([&] { k<1, 0, 36><<<sms * bps, 128>>>(out, in, iters); });
    printf("LDS frags 36 tiles  blocks/SM=%d: %.2f TFLOP/s\n", bps, 2.0 * 256 * 36 * iters * 4 * (double)(sms * bps) / ms * 1e-9);
    ms = timeit([&] { k<1, 0, 18><<<sms * bps, 128>>>(out, in, iters); });
    printf("LDS frags 18 tiles  blocks/SM=%d: %.2f TFLOP/s\n", bps, 2.0 * 256 * 18 * iters * 4 * (double)(sms * bps) / ms * 1e-9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

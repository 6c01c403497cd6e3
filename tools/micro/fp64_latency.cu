// Micro-benchmark: dependent-issue latency and throughput of DFMA / IMAD.WIDE / LDS on one SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat_dfma(double *out, long long *cyc, double a, double b) {
  double x = a;
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) x = __fma_rn(x, b, a);
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void thr_dfma(double *out, long long *cyc, double a, double b) {
  double x0 = a, x1 = a + 1, x2 = a + 2, x3 = a + 3, x4 = a + 4, x5 = a + 5, x6 = a + 6, x7 = a + 7;
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    x0 = __fma_rn(x0, b, a); x1 = __fma_rn(x1, b, a); x2 = __fma_rn(x2, b, a); x3 = __fma_rn(x3, b, a);
    x4 = __fma_rn(x4, b, a); x5 = __fma_rn(x5, b, a); x6 = __fma_rn(x6, b, a); x7 = __fma_rn(x7, b, a);
  }
  long long t1 = clock64();
  out[threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void lat_imadwide(unsigned *out, long long *cyc, unsigned a) {
  unsigned x = a, y = a + 1;
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) {
    unsigned long long p = (unsigned long long)0xD2511F53u * x;
    x = (unsigned)(p >> 32) ^ y ^ 0x9E3779B9u;
    y = (unsigned)p;
  }
  long long t1 = clock64();
  out[threadIdx.x] = x + y;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double *out; long long *cyc; unsigned *uo;
  cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8 * 64); cudaMalloc(&uo, 4 * 1024);
  long long h[8];
  for (int rep = 0; rep < 2; ++rep) {
    lat_dfma<<<1, 32>>>(out, cyc, 1.0, 0.999); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    if (rep) printf("DFMA dependent latency: %.2f cycles\n", h[0] / 256.0);
    for (int nw = 1; nw <= 8; nw *= 2) {
      thr_dfma<<<1, 32 * nw * 4>>>(out, cyc, 1.0, 0.999); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
      if (rep) printf("DFMA 8 independent chains, %d warps per scheduler: %.2f cycles per DFMA per warp (issue interval %.2f)\n", nw, h[0] / 512.0, h[0] / 512.0 / nw);
    }
    lat_imadwide<<<1, 32>>>(uo, cyc, 12345u); cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    if (rep) printf("Philox half-round (IMAD.WIDE + LOP3) dependent latency: %.2f cycles\n", h[0] / 256.0);
  }
  return 0;
}

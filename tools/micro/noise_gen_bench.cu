// Micro-benchmark: noise_gen_kernel (noise_ahead.cuh) alone on the chip -- planes of a 4096 x 4096 red-black level --
// for several (chains per warp, block size, blocks per SM) variants.  Reports us per launch and pairs of normals per ns.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I multigridmc_b200/csrc -o tools/micro/noise_gen_bench tools/micro/noise_gen_bench.cu
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#include "noise_ahead.cuh"
using namespace mgmc;

template <int NI, int THREADS, int MINB>
void run(const char *tag, NzGenP G, int nsm, int blocks_per_sm, size_t smem) {
  auto fn = noise_gen_kernel<2, NI, THREADS, MINB>;
  cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int k = 0; k < 2; ++k) fn<<<nsm * blocks_per_sm, THREADS, smem>>>(G);
  cudaEventRecord(e0);
  const int reps = 5;
  for (int k = 0; k < reps; ++k) fn<<<nsm * blocks_per_sm, THREADS, smem>>>(G);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t err = cudaGetLastError();
  const double pairs = (double)G.njobs * (G.ny - 1) * G.nz.G;
  std::printf("%-40s %8.1f us per launch  %6.2f pairs/ns  %s\n", tag, ms * 1e3 / reps, pairs / (ms * 1e6 / reps), err == cudaSuccess ? "" : cudaGetErrorString(err));
}

// where does the time of a pair of normals go?  MODE 0: Philox + Box-Muller + store, 1: no store, 2: Philox only (+ store),
// 3: Box-Muller only (+ store), 4: store only
template <int MODE>
__global__ void __launch_bounds__(512, 2) parts_kernel(NoiseP nz, double2 *buf, int total) {
  __shared__ __align__(16) double ntab[128];
  if (threadIdx.x < 128) ntab[threadIdx.x] = kNormalTabDev[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, nwarp = gridDim.x * 16, w0 = blockIdx.x * 16 + (threadIdx.x >> 5);
  double acc = 0.0;
  for (int u = w0; u < total; u += nwarp) {
    uint32_t c0 = (uint32_t)u * 32u + lane, c1 = 7u, c2 = 3u, c3 = 1u;
    double z0 = 0.0, z1 = 0.0;
    if (MODE == 0 || MODE == 1) normal_pair(nz.keys, c0, c1, c2, c3, nz.mc, ntab, z0, z1);
    if (MODE == 2) {
      philox4x32(c0, c1, c2, c3, nz.keys);
      z0 = __hiloint2double(c0, c1);
      z1 = __hiloint2double(c2, c3);
    }
    if (MODE == 3) box_muller(((uint64_t)c0 << 32) | (c0 * 2654435761u), ((uint64_t)(c0 ^ 0x9E3779B9u) << 32) | c0, nz.mc, ntab, z0, z1);
    if (MODE == 4) z0 = (double)c0;
    if (MODE == 1) acc += z0 + z1;
    else __stcs(buf + (size_t)u * 32 + lane, make_double2(z0, z1));
  }
  if (MODE == 1 && acc == 12345.678) buf[0] = make_double2(acc, acc);
}
template <int MODE>
void run_parts(const char *tag, NoiseP nz, double2 *buf, int total, int nsm, int bps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  parts_kernel<MODE><<<nsm * bps, 512>>>(nz, buf, total);
  cudaEventRecord(e0);
  for (int k = 0; k < 5; ++k) parts_kernel<MODE><<<nsm * bps, 512>>>(nz, buf, total);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  std::printf("%-40s %8.1f us per launch (%d CTAs/SM)\n", tag, ms * 1e3 / 5, bps);
}

int main() {
  int nsm = 0;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  const int n = 4096, pitch = ((16 + n + 1 + 2 + 15) / 16) * 16, gp = pitch / 4;
  NzGenP G;
  std::memset(&G, 0, sizeof(G));
  G.nz.keys = philox_round_keys(1234);
  G.nz.mc = kNormalConstsHost;
  uint32_t *d_sample;
  cudaMalloc(&d_sample, 4);
  cudaMemset(d_sample, 0, 4);
  G.nz.sample = d_sample;
  G.nz.G = n / 4 + 1;
  G.nx = G.ny = n;
  G.gp = gp;
  G.njobs = 3;
  for (int k = 0; k < 3; ++k) {
    cudaMalloc(&G.job[k].buf, (size_t)(n + 1) * gp * sizeof(double2));
    G.job[k].colour = k & 1;
    G.job[k].c1 = k;
  }
  run<1, 512, 2>("NI=1 512thr 1/SM (116KB claim)", G, nsm, 1, 116 * 1024);
  run<2, 512, 2>("NI=2 512thr 1/SM (116KB claim)", G, nsm, 1, 116 * 1024);
  run<3, 512, 2>("NI=3 512thr 1/SM (116KB claim)", G, nsm, 1, 116 * 1024);
  run<4, 512, 2>("NI=4 512thr 1/SM (116KB claim)", G, nsm, 1, 116 * 1024);
  run<2, 512, 1>("NI=2 512thr<=128reg 1/SM", G, nsm, 1, 116 * 1024);
  run<4, 512, 1>("NI=4 512thr<=128reg 1/SM", G, nsm, 1, 116 * 1024);
  run<2, 512, 2>("NI=2 512thr 2/SM", G, nsm, 2, 0);
  run<2, 512, 2>("NI=2 512thr 4/SM (full chip)", G, nsm, 4, 0);
  run<1, 512, 2>("NI=1 512thr 4/SM (full chip)", G, nsm, 4, 0);
  run<2, 1024, 1>("NI=2 1024thr 1/SM (116KB claim)", G, nsm, 1, 116 * 1024);
  run<1, 1024, 1>("NI=1 1024thr 1/SM (116KB claim)", G, nsm, 1, 116 * 1024);
  run<2, 1024, 2>("NI=2 1024thr<=32reg 1/SM", G, nsm, 1, 116 * 1024);
  const int total = 3 * 4095 * 33;
  for (int bps : {1, 4}) {
    run_parts<0>("parts: philox + box-muller + store", G.nz, G.job[0].buf, total / 3, nsm, bps);
    run_parts<1>("parts: philox + box-muller, no store", G.nz, G.job[0].buf, total / 3, nsm, bps);
    run_parts<2>("parts: philox + store", G.nz, G.job[0].buf, total / 3, nsm, bps);
    run_parts<3>("parts: box-muller + store", G.nz, G.job[0].buf, total / 3, nsm, bps);
    run_parts<4>("parts: store only", G.nz, G.job[0].buf, total / 3, nsm, bps);
  }
  return 0;
}

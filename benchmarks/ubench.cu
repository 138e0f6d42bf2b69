// Micro-benchmark: SM-wide throughput of warp-level primitives used for radix ranking.
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 256
template <int MODE>
__global__ void k(unsigned* out, const unsigned* in, long long* cyc) {
  __shared__ unsigned sh[8 * 256];
  for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  unsigned v = in[blockIdx.x * blockDim.x + threadIdx.x];
  unsigned acc = 0;
  unsigned* wc = sh + (threadIdx.x >> 5) * 256;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < ITER; ++i) {
    unsigned d = (v >> (i & 7)) & 255u;
    if (MODE == 0) acc += __match_any_sync(0xffffffffu, d);
    if (MODE == 1) acc += __ballot_sync(0xffffffffu, d & 1);
    if (MODE == 2) acc += __shfl_sync(0xffffffffu, d, (i * 7) & 31);
    if (MODE == 3) acc += atomicAdd(&wc[d], 1u);
    if (MODE == 4) { // 8-ballot match
      unsigned m = 0xffffffffu;
#pragma unroll
      for (int b = 0; b < 8; ++b) { unsigned bal = __ballot_sync(0xffffffffu, (d >> b) & 1); m &= ((d >> b) & 1) ? bal : ~bal; }
      acc += m;
    }
    if (MODE == 5) acc += __reduce_add_sync(0xffffffffu, d);
    if (MODE == 6) { atomicAdd(&wc[d], 1u); }  // no return
    v = v * 1664525u + 1013904223u;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + sh[threadIdx.x];
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, unsigned* out, unsigned* in, long long* cyc, int warps) {
  int blocks = 148;
  k<MODE><<<blocks, warps * 32>>>(out, in, cyc);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  k<MODE><<<blocks, warps * 32>>>(out, in, cyc);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  printf("%-14s warps/SM=%2d  cycles/SM per warp-instr = %.2f   (%.1f us)\n", name, warps, c / (double)(ITER * warps), ms * 1e3);
}
int main() {
  unsigned *in, *out; long long* cyc;
  cudaMalloc(&in, 148 * 1024 * 4); cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  unsigned* h = new unsigned[148 * 1024];
  for (int i = 0; i < 148 * 1024; ++i) h[i] = (unsigned)rand() * 2654435761u;
  cudaMemcpy(in, h, 148 * 1024 * 4, cudaMemcpyHostToDevice);
  for (int warps : {8, 16, 32}) {
    run<0>("match_any", out, in, cyc, warps);
    run<1>("ballot", out, in, cyc, warps);
    run<2>("shfl", out, in, cyc, warps);
    run<3>("atoms_ret", out, in, cyc, warps);
    run<6>("atoms_noret", out, in, cyc, warps);
    run<4>("ballot8_match", out, in, cyc, warps);
    run<5>("redux", out, in, cyc, warps);
  }
  return 0;
}

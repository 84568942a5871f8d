// Does ONE warp pipeline independent FP64 instructions?  K independent chains per thread (K = 1, 2, 4, 8), W warps per SM
// sub-partition (W = 1, 2, 3, 6): cycles per warp instruction.  If a lone warp with 4 independent DADD chains issues one
// every ~2 cycles the FP64 pipe is pipelined per warp; if it stays near the 8-cycle latency, only more warps raise the rate.
// Also the ESIM-like mix: per chain DADD, DADD, DFMA, DSETP+SEL, DFMA, DFMA (the potential's loop-carried chain).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int ITERS = 2048;
template <int K, int MIX>
__global__ void k(double* out, long long* cyc, double seed) {
  double a[K], b = seed * 1e-3, c = seed * 0.25;
#pragma unroll
  for (int j = 0; j < K; ++j) a[j] = seed + threadIdx.x + j;
  long long t0 = clock64();
#pragma unroll 4
  for (int it = 0; it < ITERS; ++it) {
    if (MIX == 0) {
#pragma unroll
      for (int j = 0; j < K; ++j) a[j] = __dadd_rn(a[j], b);
    } else {
      double x[K];
      unsigned h[K];
#pragma unroll
      for (int j = 0; j < K; ++j) x[j] = __dadd_rn(a[j], b);
#pragma unroll
      for (int j = 0; j < K; ++j) x[j] = __fma_rn(c, b, x[j]);
#pragma unroll
      for (int j = 0; j < K; ++j) asm volatile("{.reg .pred p; setp.ge.f64 p, %1, %2; selp.u32 %0, 0x3ff00000, 0, p;}" : "=r"(h[j]) : "d"(x[j]), "d"(c));
#pragma unroll
      for (int j = 0; j < K; ++j) x[j] = __fma_rn(-c, __hiloint2double(h[j], 0), x[j]);
#pragma unroll
      for (int j = 0; j < K; ++j) a[j] = __fma_rn(b, __hiloint2double(h[j], 0), x[j]);
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int j = 0; j < K; ++j) s += a[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int K, int MIX>
void run(int warps_per_smsp) {
  const int threads = 128 * warps_per_smsp;
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(double)); cudaMalloc(&cyc, 148 * sizeof(long long));
  k<K, MIX><<<148, threads>>>(out, cyc, 1.000001); k<K, MIX><<<148, threads>>>(out, cyc, 1.000001);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  const int per_it = MIX == 0 ? K : 6 * K;        // warp instructions per iteration and warp (mix: 4 FP64-pipe + setp + selp per chain)
  printf("%s K=%d chains, %d warps/SMSP: %6.2f cycles per iteration and warp, %5.2f cycles per instruction and warp, %5.2f per instruction and SMSP (%s)\n",
         MIX ? "ESIM-like chain" : "DADD          ", K, warps_per_smsp, avg / ITERS, avg / ITERS / per_it, avg / ITERS / per_it / warps_per_smsp,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {1, 2, 3, 6, 8}) { run<1, 0>(w); run<2, 0>(w); run<4, 0>(w); run<8, 0>(w); }
  for (int w : {1, 2, 3, 6, 8}) { run<1, 1>(w); run<4, 1>(w); }
  return 0;
}

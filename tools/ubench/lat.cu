// Dependent-chain latency of the FP64 / XU / ALU instructions of the ESIM kernel on B200, alone (1 warp per SM
// sub-partition) and under load (8 warps per sub-partition, each a dependent chain).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int ITERS = 4096;
template <int OP>
__global__ void k(double* out, long long* cyc, double seed) {
  double a = seed + threadIdx.x, b = seed, c = seed * 0.5;
  float f = 1.0f + threadIdx.x;
  uint32_t u = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int it = 0; it < ITERS; ++it) {
    if (OP == 0) a = __fma_rn(a, b, c);
    if (OP == 1) a = __dadd_rn(a, b);
    if (OP == 2) { unsigned r; asm volatile("{.reg .pred p; setp.ge.f64 p, %1, %2; selp.u32 %0, 0x3ff00000, 0, p;}" : "=r"(r) : "d"(a), "d"(b)); a = __hiloint2double(r, __double2loint(a)); }
    if (OP == 3) { asm volatile("cvt.f64.f32 %0, %1;" : "=d"(a) : "f"(f)); f = __int_as_float(__double2hiint(a)); }
    if (OP == 4) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f));
    if (OP == 5) f = fmaf(f, 1.0001f, 0.5f);
    if (OP == 6) u = (u & 0x1fffffu) ^ (u >> 3);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + f + u;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP>
void run(const char* name, int threads, int per_it) {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(double)); cudaMalloc(&cyc, 148 * sizeof(long long));
  k<OP><<<148, threads>>>(out, cyc, 1.000001); k<OP><<<148, threads>>>(out, cyc, 1.000001);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  printf("%-34s %4d threads/SM: %7.2f cycles per dependent step (%s)\n", name, threads, avg / (ITERS * per_it), cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int th : {32, 128, 1024}) {
    if (th == 32) { run<0>("DFMA", 32, 1); run<1>("DADD", 32, 1); run<2>("DSETP+SEL(+mov)", 32, 1); run<3>("F2F.F64.F32 (+mov)", 32, 1); run<4>("MUFU.LG2", 32, 1); run<5>("FFMA", 32, 1); run<6>("SHF+LOP3", 32, 1); }
    if (th == 128) { run<0>("DFMA", 128, 1); run<1>("DADD", 128, 1); run<2>("DSETP+SEL(+mov)", 128, 1); run<4>("MUFU.LG2", 128, 1); }
    if (th == 1024) { run<0>("DFMA", 1024, 1); run<1>("DADD", 1024, 1); run<2>("DSETP+SEL(+mov)", 1024, 1); run<3>("F2F.F64.F32 (+mov)", 1024, 1); run<4>("MUFU.LG2", 1024, 1); }
  }
  return 0;
}

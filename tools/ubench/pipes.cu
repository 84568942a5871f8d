// Pipe-rate microbenchmark for B200 (sm_100a): cycles per warp-instruction per SM sub-partition for the instruction
// classes the ESIM kernel is made of.  One CTA of 1024 threads per SM (8 warps per sub-partition), ILP independent chains
// per thread, clock64() around an unrolled loop.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;
constexpr int ILP = 8;

template <int OP>
__global__ void __launch_bounds__(1024, 1) k(double* out, long long* cyc, double seed, int zero) {
  double a[ILP], b = seed, c = seed * 0.5;
  float f[ILP];
  uint32_t u[ILP];
  __shared__ __align__(16) double sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
  __syncthreads();
  const uint32_t sb = static_cast<uint32_t>(__cvta_generic_to_shared(sm));
#pragma unroll
  for (int i = 0; i < ILP; ++i) { a[i] = seed + i + threadIdx.x; f[i] = 1.0f + i + threadIdx.x; u[i] = i * 977 + threadIdx.x; }
  unsigned pred = 0;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (OP == 0) a[i] = __fma_rn(a[i], b, c);
      if (OP == 1) a[i] = __dadd_rn(a[i], b);
      if (OP == 2) { unsigned r; asm volatile("{.reg .pred p; setp.ge.f64 p, %1, %2; selp.u32 %0, 1, 0, p;}" : "=r"(r) : "d"(a[i]), "d"(b)); pred += r; }
      if (OP == 3) { asm volatile("cvt.f64.f32 %0, %1;" : "=d"(a[i]) : "f"(f[i])); f[i] = __int_as_float(__double2hiint(a[i])); }
      if (OP == 14) { a[i] = __fma_rn(a[i], b, c); asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f[i])); }      // DFMA + MUFU
      if (OP == 15) { f[i] = fmaf(f[i], 1.0001f, 0.5f); u[i] = (u[i] & 0x1fffffu) ^ (u[i] >> 3); }               // FFMA + 2 ALU
      if (OP == 16) { a[i] = __fma_rn(a[i], b, c); f[i] = fmaf(f[i], 1.0001f, 0.5f); u[i] = (u[i] & 0x1fffffu) ^ (u[i] >> 3); }   // DFMA + FFMA + 2 ALU
      if (OP == 17) { a[i] = __fma_rn(a[i], b, c); a[i] = __dadd_rn(a[i], c); u[i] = (u[i] & 0x1fffffu) ^ (u[i] >> 3); u[i] = (u[i] | 0x11u) + (u[i] >> 5); }   // 2 FP64 + 4 ALU
      if (OP == 18) { uint64_t p = (uint64_t)u[i] * 0xf9b25d65u + 12345ull; u[i] = (uint32_t)(p >> 32); }
      if (OP == 20) a[i] = __fma_rn(a[i], a[(i + 3) % ILP], a[(i + 5) % ILP]);                 // DFMA, three varying operands
      if (OP == 21) a[i] = __dadd_rn(a[i], a[(i + 3) % ILP]);                                   // DADD, two varying operands
      if (OP == 22) { unsigned r; asm volatile("{.reg .pred p; setp.ge.f64 p, %1, %2; selp.u32 %0, 1, 0, p;}" : "=r"(r) : "d"(a[i]), "d"(a[(i + 3) % ILP])); pred += r; }
      if (OP == 23) { float2 v = make_float2(f[i], f[(i + 1) % ILP]); v = __ffma2_rn(v, v, make_float2(0.5f, 0.25f)); f[i] = v.x; f[(i + 1) % ILP] = v.y; }
      if (OP == 24) { float2 v = make_float2(f[i], f[(i + 1) % ILP]); v = __fadd2_rn(v, make_float2(0.5f, 0.25f)); f[i] = v.x; f[(i + 1) % ILP] = v.y; }
      if (OP == 25) { a[i] = __fma_rn(a[i], a[(i + 3) % ILP], a[(i + 5) % ILP]); asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f[i])); u[i] = (u[i] & 0x1fffffu) ^ (u[i] >> 3); f[(i + 1) % ILP] = fmaf(f[(i + 1) % ILP], 1.0001f, 0.5f); }
      if (OP == 19) { uint32_t r; asm volatile("prmt.b32 %0, %1, %2, 0x7614;" : "=r"(r) : "r"(u[i]), "r"(threadIdx.x)); u[i] = r + 1; }
      if (OP == 4) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      if (OP == 5) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      if (OP == 6) f[i] = fmaf(f[i], 1.0001f, 0.5f);
      if (OP == 7) u[i] = (u[i] & 0x1fffffu) ^ (u[i] >> 3);          // LOP3 + SHF
      if (OP == 8) { uint64_t p = (uint64_t)u[i] * 0xf9b25d65u + 12345ull; u[i] = (uint32_t)(p >> 32) ^ (uint32_t)p; }
      if (OP == 9) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(sb + ((u[i] & 0x7f00u) | ((threadIdx.x & 31) * 8)))); a[i] = v; u[i] += 0x100; }
      if (OP == 10) { double v, w; asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v), "=d"(w) : "r"(sb + ((u[i] & 0x7f80u) | ((threadIdx.x & 7) * 16)))); a[i] = v + w; u[i] += 0x80; }
      if (OP == 11) { a[i] = __fma_rn(a[i], b, c); u[i] = (u[i] & 0x1fffffu) ^ (u[i] >> 3); }      // DFMA + 2 ALU
      if (OP == 12) { a[i] = __fma_rn(a[i], b, c); f[i] = fmaf(f[i], 1.0001f, 0.5f); }               // DFMA + FFMA
      if (OP == 13) { unsigned r; asm volatile("{.reg .pred p; setp.ge.f64 p, %1, %2; selp.u32 %0, 0x3ff00000, 0, p;}" : "=r"(r) : "d"(a[i]), "d"(b)); a[i] = __fma_rn(-b, __hiloint2double(r, zero), a[i]); }
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i] + f[i] + u[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + pred;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int instr_per_it) {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(double));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  k<OP><<<148, 1024>>>(out, cyc, 1.000001, 0);
  k<OP><<<148, 1024>>>(out, cyc, 1.000001, 0);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
  // 8 warps per sub-partition, ILP*ITERS*instr_per_it warp-instructions each
  printf("%-28s %7.2f cycles per warp-instruction per sub-partition (%s)\n", name, avg / (8.0 * ILP * ITERS * instr_per_it),
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("DFMA", 1);
  run<1>("DADD", 1);
  run<2>("DSETP+SEL+IADD", 1);
  run<3>("F2F.F64.F32", 1);
  run<4>("MUFU.LG2", 1);
  run<5>("MUFU.SQRT", 1);
  run<6>("FFMA", 1);
  run<7>("LOP3+SHF+LOP3 (3 ALU)", 1);
  run<8>("IMAD.WIDE+LOP3", 1);
  run<9>("LDS.64 conflict-free", 1);
  run<10>("LDS.128 quarter-warp-free", 1);
  run<11>("DFMA + 3 ALU", 1);
  run<12>("DFMA + FFMA", 1);
  run<13>("DSETP+SEL+MOV+DFMA", 1);
  run<14>("DFMA + MUFU", 1);
  run<15>("FFMA + 2 ALU", 1);
  run<16>("DFMA + FFMA + 2 ALU", 1);
  run<17>("2 FP64 + 4 ALU", 1);
  run<18>("IMAD.WIDE", 1);
  run<19>("PRMT + IADD", 1);
  run<20>("DFMA 3 varying operands", 1);
  run<21>("DADD 2 varying operands", 1);
  run<22>("DSETP 2 varying + SEL", 1);
  run<23>("FFMA2", 1);
  run<24>("FADD2", 1);
  run<25>("DFMA3 + MUFU + 2 ALU + FFMA", 1);
  return 0;
}

// Shared device/host helpers for the v2v_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/v2v_b200.h"

namespace v2v {

// ---- host-side error plumbing ---------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void count_launch(int n = 1);

#define V2V_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      ::v2v::set_error(__VA_ARGS__);  \
      return (code);                  \
    }                                 \
  } while (0)

#define V2V_CUDA(expr)                                            \
  do {                                                            \
    cudaError_t _e = (expr);                                      \
    if (_e != cudaSuccess) return ::v2v::cuda_fail(_e, #expr);    \
  } while (0)

static inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }
__device__ __forceinline__ bool aligned_dev(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// ---- streaming loads / stores ----------------------------------------------
// Inputs are read exactly once and outputs written exactly once: keep them out
// of L1 and mark them evict-first in L2.
__device__ __forceinline__ uint32_t ld_stream_u32(const void* p) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint8_t ld_stream_u8(const void* p) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(v) : "l"(p));
  return static_cast<uint8_t>(v);
}
__device__ __forceinline__ void st_stream_f32x4(float* p, float a, float b, float c, float d) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void st_stream_f32(float* p, float a) {
  asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}
__device__ __forceinline__ double2 ld_stream_f64x2(const double* p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

// ---- exact floor division ---------------------------------------------------
// floor(a/b) as a mathematical quantity for a >= 0, b > 0 (np.floor_divide on
// float64 is fmod-based and equals it; see oracle/ and SURVEY §7).  `rb` is
// RN(1/b).  The candidate floor(a*rb) is off by at most one; the sign of a
// single-rounded FMA residual is exact, which fixes it.
__device__ __forceinline__ double floor_div_exact(double a, double b, double rb) {
  double q = floor(__dmul_rn(a, rb));
  double r = __fma_rn(-q, b, a);
  if (r < 0.0) {
    q -= 1.0;
  } else if (__fma_rn(-(q + 1.0), b, a) >= 0.0) {
    q += 1.0;
  }
  return q;
}

// ---- Philox4x32-10 -----------------------------------------------------------
struct Philox {
  static constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  __host__ __device__ static inline uint4 run(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
      uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
      uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
#else
      uint64_t p0 = static_cast<uint64_t>(M0) * c.x, p1 = static_cast<uint64_t>(M1) * c.z;
      uint32_t hi0 = static_cast<uint32_t>(p0 >> 32), lo0 = static_cast<uint32_t>(p0);
      uint32_t hi1 = static_cast<uint32_t>(p1 >> 32), lo1 = static_cast<uint32_t>(p1);
#endif
      c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
      k.x += W0;
      k.y += W1;
    }
    return c;
  }
  // Same function with the 10 round keys precomputed on the host (kernel parameters live in the constant bank,
  // so the key schedule costs no instructions on the device).
  __host__ static inline void round_keys(uint64_t seed, uint32_t (&rk)[20]) {
    uint32_t kx = static_cast<uint32_t>(seed), ky = static_cast<uint32_t>(seed >> 32);
    for (int r = 0; r < 10; ++r) {
      rk[2 * r] = kx;
      rk[2 * r + 1] = ky;
      kx += W0;
      ky += W1;
    }
  }
  __device__ static __forceinline__ uint4 run_rk(uint4 c, const uint32_t (&rk)[20]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {                      // per round: two 32x32->64 multiplies and two 3-input XORs
      const uint64_t p0 = static_cast<uint64_t>(M0) * c.x, p1 = static_cast<uint64_t>(M1) * c.z;
      c = make_uint4(static_cast<uint32_t>(p1 >> 32) ^ c.y ^ rk[2 * r], static_cast<uint32_t>(p1),
                     static_cast<uint32_t>(p0 >> 32) ^ c.w ^ rk[2 * r + 1], static_cast<uint32_t>(p0));
    }
    return c;
  }
};

// ---- xoshiro128++ (Blackman & Vigna): the sequential generator behind the per-pixel-group noise streams ----------
// Seeded from Philox4x32-10 (one counter per stream), so streams of different groups / clips are independent and the
// values do not depend on launch geometry; 9 instructions per 32-bit word.
struct GroupStream {
  uint32_t s0, s1, s2, s3;
};

__device__ __forceinline__ uint32_t group_stream_next(GroupStream& s) {      // xoshiro128++
  const uint32_t r = __funnelshift_l(s.s0 + s.s3, s.s0 + s.s3, 7) + s.s0;
  const uint32_t t = s.s1 << 9;
  const uint32_t n1 = s.s1 ^ s.s2 ^ s.s0, n0 = s.s0 ^ s.s3 ^ s.s1, n2 = s.s2 ^ s.s0 ^ t, x3 = s.s3 ^ s.s1;
  s.s0 = n0;
  s.s1 = n1;
  s.s2 = n2;
  s.s3 = __funnelshift_l(x3, x3, 11);
  return r;
}

// Two scaled normals from ONE 32-bit word (Box-Muller): the low 20 bits give the radius (lg2 + sqrt on the SFU,
// tail cut at sqrt(2 ln 2^20) = 5.27 sigma), the high 12 bits pick one of 4096 directions from a (cos, sin) table in
// shared memory (bin centres, filled once per CTA), so one Philox4x32 call yields 8 normals with 2 SFU ops each.
// `c2` = -2 ln2 * scale^2 folds the noise standard deviation into the radius.  The in-kernel generator is a statistical
// stand-in for the reference's MT19937 + polar method, which cannot be reproduced on a GPU; bit-parity runs use
// EXPLICIT noise fields instead, and the audit hooks dump exactly what this function produced.
constexpr int kTrigEntries = 2048;

__device__ __forceinline__ void fill_trig_table(float2* tab) {      // call with the whole CTA, then __syncthreads()
  for (int k = threadIdx.x; k < kTrigEntries; k += blockDim.x) {
    float s, c;      // SFU approximations (abs error ~1e-6): directions of statistical noise, cheap enough to refill per CTA
    __sincosf(static_cast<float>(2 * k + 1) * (3.14159265358979323846f / kTrigEntries), &s, &c);
    tab[k] = make_float2(c, s);
  }
}

__device__ __forceinline__ float2 box_muller16(uint32_t w, float c2, const float2* trig) {
  // mantissa trick: a float in [1,2) straight from the random bits, no int->float conversion
  const float f1 = __uint_as_float(((w << 2) & 0x007ffffcu) | 0x3f800000u);
  const float u1 = 2.0f - f1;                                  // (0,1], 20 bits
  float l, r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(u1));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(l * c2));  // scale * sqrt(-2 ln u1)
  const float2 cs = trig[w >> 21];
  return make_float2(r * cs.x, r * cs.y);
}

// Full-resolution pair from two words (used once per pixel for the hot-pixel amplitude).
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  const float f1 = __uint_as_float((a & 0x007fffffu) | 0x3f800000u);
  const float f2 = __uint_as_float((b & 0x007fffffu) | 0x3f800000u);
  const float u1 = 2.0f - f1;
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__log2f(u1) * -1.3862943611198906f));
  const float th = fmaf(f2, 6.28318530717958647692f, -6.28318530717958647692f);
  return make_float2(r * __cosf(th), r * __sinf(th));
}

// Uniform double in [0,1) with 53 random bits (same construction as numpy's
// legacy random_sample: (a>>5, b>>6) -> (a*2^26+b)/2^53).
__device__ __forceinline__ double uniform53(uint32_t a, uint32_t b) {
  return (static_cast<double>(a >> 5) * 67108864.0 + static_cast<double>(b >> 6)) * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ long long warp_sum(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace v2v

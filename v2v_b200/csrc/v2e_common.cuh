// Shared pieces of the v2e kernels: launch arguments, the counter-based random draws and the exact count helper.
#pragma once
#include "common.cuh"

namespace v2v {

struct V2eArgs {
  v2v_v2e_desc d;
  int64_t HW;
  int32_t T, G;
  double tau;          // 1/(2*pi*cutoff)
  float leak_hz_f32;   // leak_rate_hz as the float32 it becomes in `leak_rate_hz*noise_rate_array`
  uint32_t rk[20];     // Philox round keys of d.seed (host-precomputed)
  int32_t Mraw;        // frames per clip in d.frames (N, or raw_frames_per_clip with frame_index)
};

// frame n of clip b = raw frame frame_index[b][n] (clamped), or n itself (data/v2v_datasets.py:285-311)
__device__ __forceinline__ int v2e_frame_number(const V2eArgs& a, int b, int n) {
  const int32_t* fi = a.d.frame_index;
  return fi ? min(max(fi[static_cast<int64_t>(b) * a.d.N + n], 0), a.Mraw - 1) : n;
}
// pixel value after the optional per-clip value map (HDR/LDR degrade, :473-483)
__device__ __forceinline__ int v2e_mapped(const V2eArgs& a, int b, int v) {
  return a.d.value_map ? a.d.value_map[static_cast<int64_t>(b) * 256 + v] : v;
}
// rescale_intensity_frame (:190) of the mapped pixel value: (v+20)/275 in float64, or in the caller's uint8 arithmetic
// (v+20 wraps for v >= 236) when the video handed to the reference was a uint8 array
__device__ __forceinline__ double v2e_inten01(const V2eArgs& a, int mv) {
  const int w = (a.d.kernel_flags & V2V_V2E_FLAG_U8_INTENSITY) ? ((mv + 20) & 255) : (mv + 20);
  return __ddiv_rn(static_cast<double>(w), 275.0);
}

// np.floor_divide(max(diff,0), thr) for a >= thr > 0 (the caller filters a < thr); the reciprocal is only
// needed on the multi-threshold path
__device__ __forceinline__ double count_floor(double a, double thr) {
  if (a < __dadd_rn(thr, thr)) return 1.0;
  return floor_div_exact(a, thr, __drcp_rn(thr));
}

// ---- random draws of the v2e model ----------------------------------------------------------------------------
// Every kernel (generic, fast, field dump) draws the same values for the same (seed, clip, pixel, interval),
// independent of launch geometry: one xoshiro128++ stream per aligned group of 4 pixels and clip, seeded with the
// Philox4x32-10 output of counter (group lo32, 0, clip lo32, tag2|group hi|clip hi16).  Interval j (0-based) consumes,
// in this order:
//   leak jitter (only if leak_rate_hz > 0, only when j is even): four words, word k -> pixel k: one Box-Muller pair,
//                .x for interval j, .y for interval j+1;
//   shot noise  (only if shot_noise_rate_hz > 0): four words, word k -> pixel k: low 16 bits = uniform of the ON draw,
//                high 16 bits = OFF draw (bin centres).
__device__ __forceinline__ GroupStream v2e_stream_init(uint64_t g4, uint64_t clip_id, const uint32_t (&rk)[20]) {
  const uint32_t hi = (static_cast<uint32_t>(g4 >> 32) & 0x3fffu) << 16 | static_cast<uint32_t>((clip_id >> 32) & 0xffffu);
  const uint4 r = Philox::run_rk(make_uint4(static_cast<uint32_t>(g4), 0u, static_cast<uint32_t>(clip_id), 0x80000000u | hi), rk);
  GroupStream s{r.x, r.y, r.z, r.w};
  if ((s.s0 | s.s1 | s.s2 | s.s3) == 0u) s.s0 = 0x9E3779B9u;
  return s;
}

// (h + 0.5) / 65536 for a 16-bit h, through the mantissa (no int->float conversion)
__device__ __forceinline__ float v2e_u16(uint32_t h) {
  return __uint_as_float(0x3f800000u | (h << 7)) - 0.99999237060546875f;
}

__device__ __forceinline__ void v2e_shot_uniforms(GroupStream& s, float (&up)[4], float (&un)[4]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t w = group_stream_next(s);
    up[k] = v2e_u16(w & 0xffffu);
    un[k] = v2e_u16(w >> 16);
  }
}

__device__ __forceinline__ void v2e_leak_normals(GroupStream& s, const float2* trig, float (&even)[4], float (&odd)[4]) {
  const float c2 = -1.3862943611198906f;      // -2 ln 2: unit variance
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 p = box_muller16(group_stream_next(s), c2, trig);
    even[k] = p.x;
    odd[k] = p.y;
  }
}

// shot-noise Poisson rate of one pixel (float32, statistical mode): fac(v) * nominal/thres * per-frame scale (:90-99).
// The two per-pixel / per-frame factors are clamped at 0 where they are formed (a negative or NaN factor draws no
// events in every kernel), fac(v) is in [0.25, 1): the product needs no clamp of its own.
__device__ __forceinline__ float v2e_pre_prob(double nominal, double thr) { return fmaxf(static_cast<float>(__ddiv_rn(nominal, thr)), 0.f); }   // :396-399
__device__ __forceinline__ float v2e_scale_f32(double s) { return fmaxf(static_cast<float>(s), 0.f); }
__device__ __forceinline__ float v2e_shot_lambda(float facf, float pre_prob, float scale) { return __fmul_rn(__fmul_rn(facf, pre_prob), scale); }

// Poisson(lam) by inversion from one uniform.  lam is a fraction of an event per frame in every shipped preset:
// the first three CDF steps are straight-line code (k <= 2 covers all but ~lam^3/6 of the draws), the tail is a loop.
// float32: statistical mode only (the audit hook dumps exactly what this function returns).
struct PoissonCdf {
  float c0, c1, c2, p2;
};
__device__ __forceinline__ PoissonCdf poisson_cdf(float lam) {
  PoissonCdf c;
  float p0;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(__fmul_rn(lam, -1.4426950408889634f)));
  const float t = __fmul_rn(p0, lam);
  c.c0 = p0;
  c.c1 = __fadd_rn(p0, t);
  c.p2 = __fmul_rn(t, __fmul_rn(0.5f, lam));
  c.c2 = __fadd_rn(c.c1, c.p2);
  return c;
}
static __device__ __noinline__ int poisson_tail(float lam, float u, float p2, float c2) {      // u >= c2: k >= 3 (rare)
  int k = 2;
  float p = p2, cdf = c2;
  while (u >= cdf && k < 64) {
    ++k;
    p = __fmul_rn(p, __fdiv_rn(lam, static_cast<float>(k)));
    cdf = __fadd_rn(cdf, p);
  }
  return k;
}
__device__ __forceinline__ int poisson_small(float lam, float u) {
  if (!(lam > 0.f)) return 0;
  const PoissonCdf c = poisson_cdf(lam);
  if (u >= c.c2) return poisson_tail(lam, u, c.p2, c.c2);
  return (u >= c.c0 ? 1 : 0) + (u >= c.c1 ? 1 : 0);
}

int launch_v2e_fast(const V2eArgs& a, cudaStream_t s);      // v2e_fast.cu
bool v2e_fast_eligible(const V2eArgs& a);

}  // namespace v2v

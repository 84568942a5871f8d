// Shared pieces of the ESIM kernels: launch arguments and the counter-based noise.
#pragma once
#include "common.cuh"

namespace v2v {

struct EsimArgs {
  v2v_esim_desc d;
  int64_t HW;
  int32_t T;          // voxels per clip
  int32_t G;          // bins * fpb
  int64_t row_stride, plane_stride;
  int32_t padded;     // voxel rows are strided (row_stride != W)
  int32_t Tf;         // frames written per clip in frame_out
  uint32_t rk[20];    // Philox round keys of d.seed (host-precomputed)
};

constexpr int kEsimThreads = 256;

// ---- Philox noise fields ------------------------------------------------------
// All kernels (generic, fast, field dump) draw the same values for the same
// (seed, clip, pixel, interval), independent of launch geometry:
//   base noise : one Philox call per aligned group of 4 pixels and PAIR of intervals,
//                counter (group lo32, interval/2, clip lo32, tag0|group hi|clip hi16)
//                -> 8 normals (Box-Muller, 20-bit radius, 4096 tabulated directions); bn = double(std*r*cos|sin)
//   init fields: counter (pixel lo32, pixel hi32, clip lo32, tag2|clip hi16)
//                -> u0 (53 bit), hot-mask uniform (53 bit)
//   hot normal : same counter with tag3 -> z; hot = double(float(hot_pixel_std) * z)
struct NoiseKey {
  uint32_t clip_lo, clip_hi16;
};

__device__ __forceinline__ NoiseKey make_noise_key(uint64_t clip_id) {
  NoiseKey k;
  k.clip_lo = static_cast<uint32_t>(clip_id);
  k.clip_hi16 = static_cast<uint32_t>((clip_id >> 32) & 0xffffu);
  return k;
}

// Base noise of the aligned 4-pixel group g4 for the interval pair (2*pair, 2*pair+1), already multiplied
// by float(base_noise_std): even[k] belongs to pixel 4*g4+k at interval 2*pair, odd[k] at 2*pair+1.
__device__ __forceinline__ float noise_c2(float scale) { return -1.3862943611198906f * scale * scale; }

__device__ __forceinline__ void philox_noise8(uint64_t g4, uint32_t pair, const NoiseKey& nk, const uint32_t (&rk)[20], float c2,
                                              const float2* trig, float (&even)[4], float (&odd)[4]) {
  const uint4 r = Philox::run_rk(make_uint4(static_cast<uint32_t>(g4), pair, nk.clip_lo,
                                            (static_cast<uint32_t>(g4 >> 32) & 0x3fffu) << 16 | nk.clip_hi16), rk);
  const float2 p0 = box_muller16(r.x, c2, trig), p1 = box_muller16(r.y, c2, trig), p2 = box_muller16(r.z, c2, trig),
               p3 = box_muller16(r.w, c2, trig);
  even[0] = p0.x; odd[0] = p0.y;
  even[1] = p1.x; odd[1] = p1.y;
  even[2] = p2.x; odd[2] = p2.y;
  even[3] = p3.x; odd[3] = p3.y;
}

// Same values for one pixel and one interval (generic kernel, field dump).
__device__ __forceinline__ float philox_noise1(uint64_t px, uint32_t interval, const NoiseKey& nk, const uint32_t (&rk)[20], float c2,
                                               const float2* trig) {
  float ev[4], od[4];
  philox_noise8(px >> 2, interval >> 1, nk, rk, c2, trig, ev, od);
  const int k = static_cast<int>(px & 3);
  const float e = k == 0 ? ev[0] : k == 1 ? ev[1] : k == 2 ? ev[2] : ev[3];
  const float o = k == 0 ? od[0] : k == 1 ? od[1] : k == 2 ? od[2] : od[3];
  return (interval & 1u) ? o : e;
}

__device__ __forceinline__ void philox_init_pixel(uint64_t px, const NoiseKey& nk, const uint32_t (&rk)[20], double hot_fraction,
                                                  float hot_std, double* u0, double* hot) {
  const uint4 r = Philox::run_rk(make_uint4(static_cast<uint32_t>(px), static_cast<uint32_t>(px >> 32), nk.clip_lo,
                                            0x80000000u | nk.clip_hi16), rk);
  *u0 = uniform53(r.x, r.y);
  *hot = 0.0;
  if (uniform53(r.z, r.w) < hot_fraction) {
    const uint4 r2 = Philox::run_rk(make_uint4(static_cast<uint32_t>(px), static_cast<uint32_t>(px >> 32), nk.clip_lo,
                                               0xC0000000u | nk.clip_hi16), rk);
    *hot = static_cast<double>(__fmul_rn(hot_std, box_muller(r2.x, r2.y).x));
  }
}

int launch_esim_fast(const EsimArgs& a, cudaStream_t s);   // esim_fast.cu
bool esim_fast_eligible(const EsimArgs& a);

}  // namespace v2v

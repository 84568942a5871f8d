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
  int32_t Mraw;       // frames per clip in d.frames (N, or raw_frames_per_clip with frame_index)
  uint32_t rk[20];    // Philox round keys of d.seed (host-precomputed)
};

constexpr int kEsimThreads = 256;

// ---- noise fields ----------------------------------------------------------------
// All kernels (generic, fast, field dump) draw the same values for the same
// (seed, clip, pixel, interval), independent of launch geometry:
//   init fields: Philox4x32-10, counter (pixel lo32, pixel hi32, clip lo32, tag2|clip hi16)
//                -> u0 (53 bit), hot-mask uniform (53 bit)
//   hot normal : same counter with tag3 -> z; hot = double(float(hot_pixel_std) * z)
//   base noise : one stream per aligned group of 4 pixels and clip: xoshiro128++ (Blackman & Vigna) seeded with the
//                Philox4x32-10 output of counter (group lo32, 0, clip lo32, tag1|group hi|clip hi16); every PAIR of
//                intervals consumes four 32-bit outputs, word k -> pixel k -> one Box-Muller pair (20-bit radius,
//                2048 tabulated directions): .x for the even interval, .y for the odd one; bn = double(std*r*cos|sin).
//                The counter-based generator pays for itself once per pixel group (and for the init fields); the
//                per-interval draws cost 9 instructions per word instead of Philox's 19.
struct NoiseKey {
  uint32_t clip_lo, clip_hi16;
};

__device__ __forceinline__ NoiseKey make_noise_key(uint64_t clip_id) {
  NoiseKey k;
  k.clip_lo = static_cast<uint32_t>(clip_id);
  k.clip_hi16 = static_cast<uint32_t>((clip_id >> 32) & 0xffffu);
  return k;
}

__device__ __forceinline__ float noise_c2(float scale) { return -1.3862943611198906f * scale * scale; }

__device__ __forceinline__ GroupStream group_stream_init(uint64_t g4, const NoiseKey& nk, const uint32_t (&rk)[20]) {
  const uint4 r = Philox::run_rk(make_uint4(static_cast<uint32_t>(g4), 0u, nk.clip_lo,
                                            0x40000000u | (static_cast<uint32_t>(g4 >> 32) & 0x3fffu) << 16 | nk.clip_hi16), rk);
  GroupStream s{r.x, r.y, r.z, r.w};
  if ((s.s0 | s.s1 | s.s2 | s.s3) == 0u) s.s0 = 0x9E3779B9u;      // the all-zero state is the generator's only fixed point
  return s;
}

// Base noise of one aligned 4-pixel group for the NEXT pair of intervals, already multiplied by float(base_noise_std):
// even[k] belongs to pixel 4*g4+k at the even interval of the pair, odd[k] at the odd one.  Advances the stream.
__device__ __forceinline__ void stream_noise8(GroupStream& s, float c2, const float2* trig, float (&even)[4], float (&odd)[4]) {
  const uint32_t w0 = group_stream_next(s), w1 = group_stream_next(s), w2 = group_stream_next(s), w3 = group_stream_next(s);
  const float2 p0 = box_muller16(w0, c2, trig), p1 = box_muller16(w1, c2, trig), p2 = box_muller16(w2, c2, trig),
               p3 = box_muller16(w3, c2, trig);
  even[0] = p0.x; odd[0] = p0.y;
  even[1] = p1.x; odd[1] = p1.y;
  even[2] = p2.x; odd[2] = p2.y;
  even[3] = p3.x; odd[3] = p3.y;
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

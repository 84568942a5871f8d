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
// All kernels (throughput, generic, field dump) draw the same values for the same
// (seed, clip, pixel, interval), independent of launch geometry:
//   init fields: Philox4x32-10, counter (pixel lo32, pixel hi32, clip lo32, tag2|clip hi16)
//                -> u0 (53 bit), hot-mask uniform (53 bit)
//   hot normal : same counter with tag3 -> z; hot = double(float(hot_pixel_std) * z)
//   base noise : one stream per aligned group of 4 pixels and clip.  The counter-based generator pays once per stream:
//                the Philox4x32-10 block of counter (group lo32, 0, clip lo32, tag1|group hi|clip hi16) seeds a 64-bit
//                linear congruential generator s' = s*0xf9b25d65 + c (mod 2^64; 32-bit multiplier from Steele & Vigna
//                2021, spectral figures 0.91..0.76 in dimensions 2..8, tools/mwc_spectral.py; state = Philox words 0,1,
//                per-stream odd increment c = Philox words 2,3 | 1) whose HIGH word is the output: two integer
//                multiply-adds per 32-bit word.  Every interval consumes two outputs;
//                each gives one Box-Muller pair for two neighbouring pixels (word 0 -> pixels 0,1; word 1 -> pixels
//                2,3; .cos for the even pixel, .sin for the odd one): the low 21 bits are the radius (lg2 + sqrt on the
//                SFU, tail cut at sqrt(2 ln 2^21) = 5.40 sigma), the high 8 bits together with the low 3 bits of the
//                GROUP index pick one of 2048 tabulated directions (each group draws from 256 equally spaced directions
//                with a group-dependent offset of k*2pi/2048; in the throughput kernel the 8 lanes of a quarter warp
//                therefore never collide on a shared-memory bank when they fetch their 16-byte table entries).
//   The noise value is the float64 product  double(float radius) * double(direction, 21 significant bits), which is
//   EXACT, so `x = fma(radius, direction, x)` in the throughput kernel and `x += radius*direction` in the generic
//   kernel, the dump hook and the CPU oracle's replay of the dumped field are the same single rounding — the
//   reference's `potential += base_noise` (data/v2v_core_esim.py:47) applied to that value.  No FP32 multiply and no
//   float->double conversion per normal (one per pair, for the radius).
struct NoiseKey {
  uint32_t clip_lo, clip_hi16;
};

__device__ __forceinline__ NoiseKey make_noise_key(uint64_t clip_id) {
  NoiseKey k;
  k.clip_lo = static_cast<uint32_t>(clip_id);
  k.clip_hi16 = static_cast<uint32_t>((clip_id >> 32) & 0xffffu);
  return k;
}

constexpr int kDirEntries = 2048;
constexpr uint32_t kLcgMul = 0xf9b25d65u;

// {hi32(cos), hi32(sin)} of the 2048 directions (tools/gen_dir_table.py; the low words are zero).
static __device__ const uint2 g_dir_table[kDirEntries] = {
#include "dir_table.inc"
};

struct NoiseStream {
  uint32_t lo, hi;
  uint64_t inc;      // odd
};

__device__ __forceinline__ uint32_t noise_stream_next(NoiseStream& s) {
  uint64_t p;
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(s.lo), "r"(kLcgMul));
  const uint64_t t = p + s.inc;                                           // fused: IMAD.WIDE.U32
  s.hi = s.hi * kLcgMul + static_cast<uint32_t>(t >> 32);                 // IMAD
  s.lo = static_cast<uint32_t>(t);
  return s.hi;
}

__device__ __forceinline__ NoiseStream noise_stream_init(uint64_t g4, const NoiseKey& nk, const uint32_t (&rk)[20]) {
  const uint4 r = Philox::run_rk(make_uint4(static_cast<uint32_t>(g4), 0u, nk.clip_lo,
                                            0x40000000u | (static_cast<uint32_t>(g4 >> 32) & 0x3fffu) << 16 | nk.clip_hi16), rk);
  return NoiseStream{r.x, r.y, (static_cast<uint64_t>(r.w) << 32) | r.z | 1u};
}

// Per-clip constants of the radius: c2 = -2 ln2 * std^2 (so r = sqrt(c2 * lg2 u)) and 2*c2.
struct NoiseScale {
  float c2, c2x2;
};

__device__ __forceinline__ NoiseScale make_noise_scale(float std) {
  NoiseScale n;
  n.c2 = -1.3862943611198906f * std * std;
  n.c2x2 = n.c2 + n.c2;
  return n;
}

// Radius of one stream word: u = (2^21 - (w & 0x1fffff)) / 2^21 in (0,1] is formed as 4*(1.25 - f), f in [1,1.25)
// straight from the mantissa bits (a zero std gives an exactly zero field).
__device__ __forceinline__ float noise_word_radius(uint32_t w, const NoiseScale& n) {
  const float f = __uint_as_float((w & 0x001fffffu) | 0x3f800000u);
  float l, r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(1.25f - f));
  // (|.|: the SFU's lg2 of a value just below 0.25 may come out a hair above -2, which would make the product negative)
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fabsf(fmaf(l, n.c2, n.c2x2))));
  return r;
}

// Direction index of a stream word for pixel group g4: the word's top 8 bits, then the group's low 3 bits.
__device__ __forceinline__ uint32_t noise_word_dir(uint32_t w, uint32_t g4_low3) { return ((w >> 21) & 0x7f8u) | g4_low3; }

// Base noise of one aligned 4-pixel group for the NEXT interval as float64 values (generic kernel, dump hook):
// bn[k] belongs to pixel 4*g4+k.  Advances the stream by two words.
__device__ __forceinline__ void stream_noise4(NoiseStream& s, const NoiseScale& n, const uint2* dir, uint32_t g4_low3, double (&bn)[4]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const uint32_t w = noise_stream_next(s);
    const double r = static_cast<double>(noise_word_radius(w, n));
    const uint2 cs = dir[noise_word_dir(w, g4_low3)];
    bn[2 * h] = __dmul_rn(r, __hiloint2double(static_cast<int>(cs.x), 0));     // exact products
    bn[2 * h + 1] = __dmul_rn(r, __hiloint2double(static_cast<int>(cs.y), 0));
  }
}

__device__ __forceinline__ void fill_dir_table(uint2* tab) {      // call with the whole CTA, then __syncthreads()
  for (int k = threadIdx.x; k < kDirEntries; k += blockDim.x) tab[k] = g_dir_table[k];
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

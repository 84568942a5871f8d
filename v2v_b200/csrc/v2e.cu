// v2e-style frames -> voxel (sm_100a).
//
// Replaces video_to_voxel / EventEmulator.generate_events of the reference's
// deprecated richer sensor model (data/v2v_core_v2e.py:401-581): float32 log
// LUT, intensity-dependent IIR low-pass, leak with jitter, per-pixel ON/OFF
// threshold maps, Poisson shot noise.  Same streaming structure as the ESIM
// kernel: one thread owns P pixels for the whole clip, state in registers,
// frames through a register ring, voxels out as 128-bit stores.
//
// Dtype rules follow the reference exactly: the memorised
// brightness `base` and `diff` are float32 only when neither the low-pass nor
// the leak runs (state_f32), float64 otherwise; every float64 step is a single
// correctly rounded operation (no FMA contraction).
#include "v2e_common.cuh"

#include <algorithm>
#include <cstdlib>

namespace v2v {
namespace {

constexpr int kV2eThreads = 256;
constexpr int kPF = 4;

__device__ __forceinline__ uint32_t load_pix4(const uint8_t* p) { return ld_stream_u32(p); }

struct V2eLuts {
  float logv[256];     // float32(log(v/255+0.01)), from the host          (:120-137)
  double inten[256];   // (v+20)/275                                       (:190)
  float facf[256];     // 1 - 0.75*inten as float (shot-noise rate factor) (:90)
};

template <int P, bool F32STATE, bool CUTOFF, bool LEAK, bool SHOT>
__global__ void __launch_bounds__(kV2eThreads) v2e_kernel(const V2eArgs a) {
  __shared__ V2eLuts L;
  __shared__ float2 trig_s[LEAK ? kTrigEntries : 1];
  const v2v_v2e_desc& d = a.d;
  if (LEAK && d.noise_mode == V2V_NOISE_PHILOX) fill_trig_table(trig_s);
  for (int i = threadIdx.x; i < 256; i += kV2eThreads) {
    const int mv = v2e_mapped(a, blockIdx.y, i);                  // the degrade is a function of the pixel value: folded into the LUTs
    L.logv[i] = d.lut[mv];
    const double it = v2e_inten01(a, mv);
    L.inten[i] = it;
    L.facf[i] = static_cast<float>(__dsub_rn(1.0, __dmul_rn(0.75, it)));
  }
  __syncthreads();
  const int b = blockIdx.y;
  const int64_t pix0 = (static_cast<int64_t>(blockIdx.x) * kV2eThreads + threadIdx.x) * P;
  const int64_t HW = a.HW;
  if (pix0 >= HW) return;
  const int N = d.N;
  const int64_t mp = static_cast<int64_t>(b) * HW + pix0;
  const uint64_t clip_id = d.clip_index_base + static_cast<uint64_t>(b);
  const bool leak_on = d.leak_rate_hz > 0.0;
  GroupStream gs{0u, 0u, 0u, 0u};
  float lodd[4] = {0.f, 0.f, 0.f, 0.f};
  if (d.noise_mode == V2V_NOISE_PHILOX) gs = v2e_stream_init(static_cast<uint64_t>(pix0) >> 2, clip_id, a.rk);
  const bool philox = d.noise_mode == V2V_NOISE_PHILOX, explicit_noise = d.noise_mode == V2V_NOISE_EXPLICIT;

  double pth[P], nth[P], lp[P], base[P];
  float nrate[LEAK ? P : 1], ppf[SHOT ? P : 1], npf[SHOT ? P : 1];
  const bool thr_pi = d.thres_per_interval != 0;      // maps re-drawn on every frame (:417-421): reloaded per interval below
#pragma unroll
  for (int k = 0; k < P; ++k) {
    pth[k] = thr_pi ? 1.0 : d.pos_thres[mp + k];
    nth[k] = thr_pi ? 1.0 : d.neg_thres[mp + k];
    if (LEAK) nrate[LEAK ? k : 0] = d.noise_rate ? d.noise_rate[mp + k] : 1.0f;
    if (SHOT) {                                                                    // :396-399
      ppf[SHOT ? k : 0] = v2e_pre_prob(d.pos_thres_nominal, pth[k]);
      npf[SHOT ? k : 0] = v2e_pre_prob(d.neg_thres_nominal, nth[k]);
    }
  }

  const uint8_t* fr = d.frames + static_cast<int64_t>(b) * a.Mraw * HW + pix0;
  auto load = [&](int i) -> uint32_t {
    const uint8_t* p = fr + static_cast<int64_t>(v2e_frame_number(a, b, i)) * HW;
    return P == 4 ? load_pix4(p) : static_cast<uint32_t>(ld_stream_u8(p));
  };
  {
    // first frame: lp = log_new; the filter runs with dt = 0 (eps = 0); base = lp   (:463-478)
    const uint32_t w0 = load(0);
#pragma unroll
    for (int k = 0; k < P; ++k) {
      const double l0 = static_cast<double>(L.logv[(w0 >> (8 * k)) & 0xffu]);
      lp[k] = CUTOFF ? __dadd_rn(__dmul_rn(1.0, l0), __dmul_rn(0.0, l0)) : l0;
      base[k] = lp[k];
    }
  }

  float* vox = d.voxel + static_cast<int64_t>(b) * a.T * d.num_bins * HW + pix0;
  int acc[P];
#pragma unroll
  for (int k = 0; k < P; ++k) acc[k] = 0;
  unsigned int npos = 0, nneg = 0;
  int sub = 0;
  double t_prev = 0.0;

  uint32_t ring[kPF];
#pragma unroll
  for (int u = 0; u < kPF; ++u) ring[u] = (1 + u < N) ? load(1 + u) : 0u;

  for (int i0 = 1; i0 < N; i0 += kPF) {
#pragma unroll
    for (int u = 0; u < kPF; ++u) {
      const int i = i0 + u;
      if (i >= N) break;
      const uint32_t w = ring[u];
      if (i + kPF < N) ring[u] = load(i + kPF);
      // per-interval quantities shared by all pixels
      const double t_k = __ddiv_rn(static_cast<double>(i), d.fps);     // :577
      const double dt = __dsub_rn(t_k, t_prev);                         // :442
      t_prev = t_k;
      const double qdt = CUTOFF ? __ddiv_rn(dt, a.tau) : 0.0;           // :167
      const int64_t fo = (static_cast<int64_t>(b) * (N - 1) + (i - 1)) * HW + pix0;
      if (thr_pi) {
#pragma unroll
        for (int k = 0; k < P; ++k) {
          pth[k] = d.pos_thres[fo + k];
          nth[k] = d.neg_thres[fo + k];
        }
      }
      float sps = 0.f, sns = 0.f;
      if (SHOT && philox) {
        const int64_t si = static_cast<int64_t>(b) * (N - 1) + (i - 1);
        sps = v2e_scale_f32(d.shot_pos_scale[si]);
        sns = v2e_scale_f32(d.shot_neg_scale[si]);
      }

      // per-interval random fields
      double lr[P];
      int sp[P], sn[P];
#pragma unroll
      for (int k = 0; k < P; ++k) { lr[k] = 0.0; sp[k] = 0; sn[k] = 0; }
      if (explicit_noise) {
        if (LEAK && d.leak_randn) {
#pragma unroll
          for (int k = 0; k < P; ++k) lr[k] = d.leak_randn[fo + k];
        }
        if (SHOT && d.pos_shot && d.neg_shot) {
#pragma unroll
          for (int k = 0; k < P; ++k) { sp[k] = d.pos_shot[fo + k]; sn[k] = d.neg_shot[fo + k]; }
        }
      } else if (philox && (LEAK || SHOT)) {
        float lz[4] = {0.f, 0.f, 0.f, 0.f}, up[4] = {1.f, 1.f, 1.f, 1.f}, un[4] = {1.f, 1.f, 1.f, 1.f};
        if (LEAK && leak_on) {                       // stream order: leak pair (even intervals), then shot
          if (((i - 1) & 1) == 0) v2e_leak_normals(gs, trig_s, lz, lodd);
          else {
#pragma unroll
            for (int k = 0; k < 4; ++k) lz[k] = lodd[k];
          }
        }
        if (SHOT) v2e_shot_uniforms(gs, up, un);
#pragma unroll
        for (int k = 0; k < P; ++k) {
          const int j = P == 4 ? k : static_cast<int>(pix0 & 3);
          const float lzj = j == 0 ? lz[0] : j == 1 ? lz[1] : j == 2 ? lz[2] : lz[3];
          const float upj = j == 0 ? up[0] : j == 1 ? up[1] : j == 2 ? up[2] : up[3];
          const float unj = j == 0 ? un[0] : j == 1 ? un[1] : j == 2 ? un[2] : un[3];
          if (LEAK) lr[k] = static_cast<double>(lzj);
          if (SHOT) {
            const float fac = L.facf[(w >> (8 * k)) & 0xffu];
            sp[k] = poisson_small(v2e_shot_lambda(fac, ppf[SHOT ? k : 0], sps), upj);
            sn[k] = poisson_small(v2e_shot_lambda(fac, npf[SHOT ? k : 0], sns), unj);
          }
        }
      }

      float outv[P];
#pragma unroll
      for (int k = 0; k < P; ++k) {
        const uint32_t v = (w >> (8 * k)) & 0xffu;
        const float lognew = L.logv[v];                                             // :447
        if (CUTOFF) {                                                               // :157-173
          double eps = __dmul_rn(L.inten[v], qdt);
          eps = fmin(eps, 1.0);
          lp[k] = __dadd_rn(__dmul_rn(__dsub_rn(1.0, eps), lp[k]), __dmul_rn(eps, static_cast<double>(lognew)));
        } else {
          lp[k] = static_cast<double>(lognew);
        }
        if (LEAK) {                                                                 // :192-211
          const float r32 = __fmul_rn(a.leak_hz_f32, nrate[LEAK ? k : 0]);
          const double rate = __dmul_rn(static_cast<double>(r32), __dsub_rn(1.0, __dmul_rn(d.leak_jitter_fraction, lr[k])));
          base[k] = __dsub_rn(base[k], __dmul_rn(__dmul_rn(dt, rate), pth[k]));
        }
        double diff;                                                                // :503
        if (F32STATE) diff = static_cast<double>(__fsub_rn(static_cast<float>(lp[k]), static_cast<float>(base[k])));
        else diff = __dsub_rn(lp[k], base[k]);
        double pe = 0.0, ne = 0.0;                                                  // :55-60
        if (diff >= pth[k]) pe = count_floor(diff, pth[k]);
        else if (-diff >= nth[k]) ne = count_floor(-diff, nth[k]);
        if (SHOT) {                                                                 // :530-531
          pe += static_cast<double>(sp[k]);
          ne += static_cast<double>(sn[k]);
        }
        if (pe != 0.0 || ne != 0.0) {
          // :547-548 — in place: the float64 sum is cast back to the state dtype after each line
          double nb = __dadd_rn(base[k], __dmul_rn(pe, pth[k]));
          if (F32STATE) nb = static_cast<double>(__double2float_rn(nb));
          nb = __dsub_rn(nb, __dmul_rn(ne, nth[k]));
          if (F32STATE) nb = static_cast<double>(__double2float_rn(nb));
          base[k] = nb;
          npos += static_cast<unsigned int>(pe);
          nneg += static_cast<unsigned int>(ne);
          acc[k] += static_cast<int>(pe) - static_cast<int>(ne);                    // :579-580
        }
        outv[k] = static_cast<float>(acc[k]);
      }
      if (++sub == d.frames_per_bin) {
        sub = 0;
        if (P == 4) st_stream_f32x4(vox, outv[0], outv[1 % P], outv[2 % P], outv[3 % P]);
        else st_stream_f32(vox, outv[0]);
        vox += HW;
#pragma unroll
        for (int k = 0; k < P; ++k) acc[k] = 0;
      }
    }
  }
  if (d.stats) {
    unsigned long long* st = reinterpret_cast<unsigned long long*>(d.stats + 2 * b);
    if (npos) atomicAdd(st, static_cast<unsigned long long>(npos));
    if (nneg) atomicAdd(st + 1, static_cast<unsigned long long>(nneg));
  }
}

// Full-frame means of generate_shot_noise (:90-96) -> per-frame Poisson scales.
// One pass over the clip: a thread keeps nominal/thres of its 4 pixels in registers, walks the frames and adds
// fac(v)*pre_prob to per-frame sums.  The sums are accumulated as 2^-36 fixed point in int64 (exact and order
// independent, so the scales - and with them the Poisson draws - are run-to-run deterministic): per lane the four
// products are summed in float64 in a fixed order and converted once, a warp adds its 32 integers with two REDUX
// halves, every warp of the CTA keeps its own shared-memory accumulator row per frame (plain adds by lane 0), and the CTA
// flushes the row sums to the output arrays at the end (one global atomic per frame, polarity and CTA); a second tiny kernel converts in place to (rate/2*dt) / mean.
constexpr double kShotFix = 68719476736.0;   // 2^36
constexpr int kShotChunk = 256;              // frames per shared-memory pass (8 warps x 2 x 2 KB of accumulators)

__global__ void __launch_bounds__(256) v2e_shot_accum_kernel(const V2eArgs a, long long* pos_acc, long long* neg_acc, int tiles_per_cta) {
  __shared__ double fac_s[256];
  __shared__ unsigned long long sacc[8][2][kShotChunk];      // one row per warp: lane 0 adds without atomics
  const v2v_v2e_desc& d = a.d;
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < 256; i += 256) fac_s[i] = __dsub_rn(1.0, __dmul_rn(0.75, v2e_inten01(a, v2e_mapped(a, b, i))));
  const bool vec = (a.HW % 4 == 0) && aligned_dev(d.frames, 4);
  const int M = d.N - 1;
  for (int c0 = 0; c0 < M; c0 += kShotChunk) {
    const int cn = min(kShotChunk, M - c0);
    for (int i = threadIdx.x; i < 8 * 2 * kShotChunk; i += 256) (&sacc[0][0][0])[i] = 0ull;
    __syncthreads();
    for (int tile = 0; tile < tiles_per_cta; ++tile) {
      const int64_t pix0 = ((static_cast<int64_t>(blockIdx.x) * tiles_per_cta + tile) * 256 + threadIdx.x) * 4;
      if (pix0 - threadIdx.x * 4 >= a.HW) break;                   // whole tile past the frame (CTA-uniform)
      double pp[4], np_[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool ok = pix0 + k < a.HW;
        pp[k] = ok ? (d.pos_thres_nominal / d.pos_thres[static_cast<int64_t>(b) * a.HW + pix0 + k]) * kShotFix : 0.0;
        np_[k] = ok ? (d.neg_thres_nominal / d.neg_thres[static_cast<int64_t>(b) * a.HW + pix0 + k]) * kShotFix : 0.0;
      }
      bool lane_small = true;
#pragma unroll
      for (int k = 0; k < 4; ++k) lane_small = lane_small && pp[k] >= 0.0 && np_[k] >= 0.0 && pp[k] < 256.0 * kShotFix && np_[k] < 256.0 * kShotFix;
      const bool small = __all_sync(0xffffffffu, lane_small);      // warp-uniform: the lane sums stay below 2^46
      const uint8_t* fr0 = d.frames + static_cast<int64_t>(b) * a.Mraw * a.HW + pix0;
      auto load = [&](int j) -> uint32_t {
        uint32_t w = 0;
        if (j < cn && pix0 < a.HW) {
          const uint8_t* fr = fr0 + static_cast<int64_t>(v2e_frame_number(a, b, c0 + 1 + j)) * a.HW;
          if (vec) w = ld_stream_u32(fr);
          else {
#pragma unroll
            for (int k = 0; k < 4; ++k) if (pix0 + k < a.HW) w |= static_cast<uint32_t>(fr[k]) << (8 * k);
          }
        }
        return w;
      };
      constexpr int kAhead = 8;                                     // frames in flight per lane (the pass is latency bound otherwise)
      uint32_t ring[kAhead];
#pragma unroll
      for (int u = 0; u < kAhead; ++u) ring[u] = load(u);
      for (int j0 = 0; j0 < cn; j0 += kAhead) {
#pragma unroll
        for (int u = 0; u < kAhead; ++u) {
          const int j = j0 + u;
          if (j >= cn) break;                                        // CTA-uniform
          const uint32_t w = ring[u];
          ring[u] = load(j + kAhead);
          double sp = 0.0, sn = 0.0;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const double fc = fac_s[(w >> (8 * k)) & 0xffu];
            sp = __dadd_rn(sp, __dmul_rn(fc, pp[k]));
            sn = __dadd_rn(sn, __dmul_rn(fc, np_[k]));
          }
          unsigned long long ps, ns;
          unsigned int plo, nlo;
          if (small) {
            // RN(sp) as an integer without the conversion unit: adding 2^52+2^51 leaves it in the low mantissa bits
            // (sp < 2^46 here); 64-bit warp sums from two 23-bit halves (32 of them stay below 2^28), one REDUX each
            const double tp = __dadd_rn(sp, 6755399441055744.0), tn = __dadd_rn(sn, 6755399441055744.0);
            const unsigned int pl = static_cast<unsigned int>(__double2loint(tp)), ph = static_cast<unsigned int>(__double2hiint(tp));
            const unsigned int nl = static_cast<unsigned int>(__double2loint(tn)), nh = static_cast<unsigned int>(__double2hiint(tn));
            plo = __reduce_add_sync(0xffffffffu, pl & 0x7fffffu);
            nlo = __reduce_add_sync(0xffffffffu, nl & 0x7fffffu);
            ps = static_cast<unsigned long long>(__reduce_add_sync(0xffffffffu, __funnelshift_r(pl, ph, 23) & 0x7fffffu)) << 23;
            ns = static_cast<unsigned long long>(__reduce_add_sync(0xffffffffu, __funnelshift_r(nl, nh, 23) & 0x7fffffu)) << 23;
          } else {                                                   // nominal/thres beyond 256: plain 64-bit shuffles
            ps = static_cast<unsigned long long>(warp_sum(__double2ll_rn(sp)));
            ns = static_cast<unsigned long long>(warp_sum(__double2ll_rn(sn)));
            plo = nlo = 0u;
          }
          if ((threadIdx.x & 31) == 0) {
            sacc[threadIdx.x >> 5][0][j] += ps + plo;
            sacc[threadIdx.x >> 5][1][j] += ns + nlo;
          }
        }
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * cn; i += 256) {
      const int pol = i >= cn, j = pol ? i - cn : i;
      unsigned long long v = 0ull;
#pragma unroll
      for (int wv = 0; wv < 8; ++wv) v += sacc[wv][pol][j];
      if (v) atomicAdd(reinterpret_cast<unsigned long long*>((pol ? neg_acc : pos_acc) + static_cast<int64_t>(b) * M + c0 + j), v);
    }
    __syncthreads();
  }
}

__global__ void v2e_shot_finalize_kernel(const V2eArgs a, double* pos_scale, double* neg_scale) {
  const v2v_v2e_desc& d = a.d;
  const int64_t n = static_cast<int64_t>(d.B) * (d.N - 1);
  const int64_t o = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (o >= n) return;
  const int k = static_cast<int>(o % (d.N - 1)) + 1;
  const double dt = static_cast<double>(k) / d.fps - static_cast<double>(k - 1) / d.fps;
  const double sf = (d.shot_noise_rate_hz / 2) * dt;
  const double tp = static_cast<double>(reinterpret_cast<long long*>(pos_scale)[o]) / kShotFix;
  const double tn = static_cast<double>(reinterpret_cast<long long*>(neg_scale)[o]) / kShotFix;
  pos_scale[o] = sf / (tp / static_cast<double>(a.HW));
  neg_scale[o] = sf / (tn / static_cast<double>(a.HW));
}

// Audit hook: the random fields a PHILOX run draws, for explicit replay / oracle checks.
__global__ void v2e_philox_fields_kernel(const V2eArgs a, double* leak_randn, int32_t* pos_shot, int32_t* neg_shot) {
  __shared__ float2 trig_s[kTrigEntries];
  fill_trig_table(trig_s);
  __syncthreads();
  const v2v_v2e_desc& d = a.d;
  const int b = blockIdx.y;
  const int64_t pix = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (pix >= a.HW) return;
  const uint64_t clip_id = d.clip_index_base + static_cast<uint64_t>(b);
  const bool shot = d.shot_noise_rate_hz > 0.0, leak = d.leak_rate_hz > 0.0;
  const int64_t mp = static_cast<int64_t>(b) * a.HW + pix;
  const float ppf = v2e_pre_prob(d.pos_thres_nominal, d.pos_thres[mp]);
  const float npf = v2e_pre_prob(d.neg_thres_nominal, d.neg_thres[mp]);
  const int j = static_cast<int>(pix & 3);
  GroupStream gs = v2e_stream_init(static_cast<uint64_t>(pix) >> 2, clip_id, a.rk);
  float lodd[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = 1; i < d.N; ++i) {
    float lz[4] = {0.f, 0.f, 0.f, 0.f}, up[4] = {1.f, 1.f, 1.f, 1.f}, un[4] = {1.f, 1.f, 1.f, 1.f};
    if (leak) {
      if (((i - 1) & 1) == 0) v2e_leak_normals(gs, trig_s, lz, lodd);
      else {
        for (int k = 0; k < 4; ++k) lz[k] = lodd[k];
      }
    }
    if (shot) v2e_shot_uniforms(gs, up, un);
    const int64_t o = (static_cast<int64_t>(b) * (d.N - 1) + (i - 1)) * a.HW + pix;
    if (leak_randn) leak_randn[o] = static_cast<double>(lz[j]);
    if (shot && pos_shot && neg_shot) {
      const uint32_t v = v2e_mapped(a, b, d.frames[(static_cast<int64_t>(b) * a.Mraw + v2e_frame_number(a, b, i)) * a.HW + pix]);
      const double it = v2e_inten01(a, static_cast<int>(v));
      const float fac = static_cast<float>(__dsub_rn(1.0, __dmul_rn(0.75, it)));
      const int64_t si = static_cast<int64_t>(b) * (d.N - 1) + (i - 1);
      pos_shot[o] = poisson_small(v2e_shot_lambda(fac, ppf, v2e_scale_f32(d.shot_pos_scale[si])), up[j]);
      neg_shot[o] = poisson_small(v2e_shot_lambda(fac, npf, v2e_scale_f32(d.shot_neg_scale[si])), un[j]);
    }
  }
}

int validate(const v2v_v2e_desc& d, V2eArgs* a) {
  V2V_REQUIRE(d.B >= 0 && d.N >= 1 && d.H >= 0 && d.W >= 0, V2V_ERR_INVALID_ARG, "bad shape");
  V2V_REQUIRE(d.num_bins >= 1 && d.frames_per_bin >= 1, V2V_ERR_INVALID_ARG, "num_bins and frames_per_bin must be >= 1");
  a->G = d.num_bins * d.frames_per_bin;
  V2V_REQUIRE((d.N - 1) % a->G == 0, V2V_ERR_SHAPE, "(N-1)=%d is not a multiple of num_bins*frames_per_bin=%d", d.N - 1, a->G);
  V2V_REQUIRE(d.B <= 65535, V2V_ERR_UNSUPPORTED, "B > 65535");
  V2V_REQUIRE(d.fps > 0.0, V2V_ERR_INVALID_ARG, "fps must be > 0");
  V2V_REQUIRE(d.noise_mode >= 0 && d.noise_mode <= 2, V2V_ERR_INVALID_ARG, "bad noise_mode");
  a->d = d;
  a->HW = static_cast<int64_t>(d.H) * d.W;
  a->T = (d.N - 1) / a->G;
  a->tau = d.cutoff_hz > 0.0 ? 1.0 / (3.141592653589793 * 2 * d.cutoff_hz) : 0.0;    // :162
  a->leak_hz_f32 = static_cast<float>(d.leak_rate_hz);
  Philox::round_keys(d.seed, a->rk);
  V2V_REQUIRE(d.raw_frames_per_clip >= 0 && (d.frame_index || d.raw_frames_per_clip == 0 || d.raw_frames_per_clip == d.N),
              V2V_ERR_INVALID_ARG, "raw_frames_per_clip=%d needs frame_index", d.raw_frames_per_clip);
  a->Mraw = (d.frame_index && d.raw_frames_per_clip > 0) ? d.raw_frames_per_clip : d.N;
  return V2V_OK;
}

}  // namespace
}  // namespace v2v

extern "C" int v2v_v2e_frames_to_voxel(const v2v_v2e_desc* desc, void* stream) {
  using namespace v2v;
  V2V_REQUIRE(desc != nullptr, V2V_ERR_INVALID_ARG, "desc is NULL");
  V2eArgs a;
  int rc = validate(*desc, &a);
  if (rc != V2V_OK) return rc;
  const v2v_v2e_desc& d = *desc;
  if (d.B == 0 || a.HW == 0 || d.N == 1) return V2V_OK;
  V2V_REQUIRE(d.frames && d.lut && d.pos_thres && d.neg_thres && d.voxel, V2V_ERR_INVALID_ARG,
              "frames, lut, pos_thres, neg_thres and voxel must be non-NULL");
  V2V_REQUIRE(!d.thres_per_interval || d.noise_mode != V2V_NOISE_PHILOX, V2V_ERR_UNSUPPORTED,
              "per-interval threshold maps come with host-drawn fields (noise_mode NONE or EXPLICIT)");
  V2V_REQUIRE(!(d.state_f32 && (d.cutoff_hz > 0.0 || d.leak_rate_hz > 0.0)), V2V_ERR_INVALID_ARG,
              "state_f32 is only meaningful with cutoff_hz<=0 and leak_rate_hz<=0");
  if (d.noise_mode == V2V_NOISE_PHILOX && d.shot_noise_rate_hz > 0.0)
    V2V_REQUIRE(d.shot_pos_scale && d.shot_neg_scale, V2V_ERR_INVALID_ARG, "PHILOX shot noise needs the scales from v2v_v2e_shot_scales");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (v2e_fast_eligible(a) && !(d.kernel_flags & V2V_V2E_FLAG_GENERIC)) return launch_v2e_fast(a, s);
  const bool vec4 = (a.HW % 4 == 0) && aligned(d.frames, 4) && aligned(d.voxel, 16) &&
                    static_cast<int64_t>(d.B) * a.HW >= 148LL * 2048;
  const int P = vec4 ? 4 : 1;
  dim3 grid(static_cast<unsigned int>(((a.HW + P - 1) / P + kV2eThreads - 1) / kV2eThreads), static_cast<unsigned int>(d.B));
  const bool cut = d.cutoff_hz > 0.0, lk = d.leak_rate_hz > 0.0, sh = d.shot_noise_rate_hz > 0.0 && d.noise_mode != V2V_NOISE_NONE;
#define V2V_V2E(PP, F32, CU, LK, SH) v2e_kernel<PP, F32, CU, LK, SH><<<grid, kV2eThreads, 0, s>>>(a)
#define V2V_V2E_P(PP)                                                                   \
  do {                                                                                  \
    if (d.state_f32) { if (sh) V2V_V2E(PP, true, false, false, true); else V2V_V2E(PP, true, false, false, false); } \
    else if (cut && lk) { if (sh) V2V_V2E(PP, false, true, true, true); else V2V_V2E(PP, false, true, true, false); } \
    else if (cut) { if (sh) V2V_V2E(PP, false, true, false, true); else V2V_V2E(PP, false, true, false, false); }     \
    else { if (sh) V2V_V2E(PP, false, false, true, true); else V2V_V2E(PP, false, false, true, false); }              \
  } while (0)
  if (vec4) V2V_V2E_P(4); else V2V_V2E_P(1);
#undef V2V_V2E_P
#undef V2V_V2E
  count_launch();
  V2V_CUDA(cudaGetLastError());
  return V2V_OK;
}

extern "C" int v2v_v2e_shot_scales(const v2v_v2e_desc* desc, double* shot_pos_scale, double* shot_neg_scale, void* stream) {
  using namespace v2v;
  V2V_REQUIRE(desc != nullptr && shot_pos_scale && shot_neg_scale, V2V_ERR_INVALID_ARG, "NULL argument");
  V2eArgs a;
  int rc = validate(*desc, &a);
  if (rc != V2V_OK) return rc;
  const v2v_v2e_desc& d = *desc;
  if (d.B == 0 || a.HW == 0 || d.N == 1) return V2V_OK;
  V2V_REQUIRE(d.frames && d.pos_thres && d.neg_thres, V2V_ERR_INVALID_ARG, "frames and threshold maps must be non-NULL");
  V2V_REQUIRE(aligned(shot_pos_scale, 8) && aligned(shot_neg_scale, 8), V2V_ERR_ALIGNMENT, "misaligned scale arrays");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t nbytes = static_cast<size_t>(d.B) * (d.N - 1) * sizeof(double);
  V2V_CUDA(cudaMemsetAsync(shot_pos_scale, 0, nbytes, s));
  V2V_CUDA(cudaMemsetAsync(shot_neg_scale, 0, nbytes, s));
  const int64_t tiles = (a.HW + 1023) / 1024;                       // 256 lanes x 4 pixels
  // several tiles per CTA once the grid is a few waves deep: fewer per-frame flushes to the global accumulators
  const int tpc = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(8, tiles * d.B / (148 * 8))));
  dim3 grid(static_cast<unsigned int>((tiles + tpc - 1) / tpc), static_cast<unsigned int>(d.B));
  v2e_shot_accum_kernel<<<grid, 256, 0, s>>>(a, reinterpret_cast<long long*>(shot_pos_scale), reinterpret_cast<long long*>(shot_neg_scale), tpc);
  const int64_t n = static_cast<int64_t>(d.B) * (d.N - 1);
  v2e_shot_finalize_kernel<<<static_cast<unsigned int>((n + 255) / 256), 256, 0, s>>>(a, shot_pos_scale, shot_neg_scale);
  count_launch(2);
  V2V_CUDA(cudaGetLastError());
  return V2V_OK;
}

extern "C" int v2v_v2e_philox_fields(const v2v_v2e_desc* desc, double* leak_randn, int32_t* pos_shot, int32_t* neg_shot, void* stream) {
  using namespace v2v;
  V2V_REQUIRE(desc != nullptr, V2V_ERR_INVALID_ARG, "desc is NULL");
  V2eArgs a;
  int rc = validate(*desc, &a);
  if (rc != V2V_OK) return rc;
  const v2v_v2e_desc& d = *desc;
  if (d.B == 0 || a.HW == 0 || d.N == 1) return V2V_OK;
  V2V_REQUIRE(d.frames && d.pos_thres && d.neg_thres, V2V_ERR_INVALID_ARG, "frames and threshold maps must be non-NULL");
  V2V_REQUIRE(!(d.shot_noise_rate_hz > 0.0 && pos_shot) || (d.shot_pos_scale && d.shot_neg_scale), V2V_ERR_INVALID_ARG,
              "shot fields need the scales from v2v_v2e_shot_scales");
  dim3 grid(static_cast<unsigned int>((a.HW + 255) / 256), static_cast<unsigned int>(d.B));
  v2e_philox_fields_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, leak_randn, pos_shot, neg_shot);
  count_launch();
  V2V_CUDA(cudaGetLastError());
  return V2V_OK;
}

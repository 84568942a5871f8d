// v2e-style frames -> voxel: the throughput kernel (noise none or in-kernel Philox, 4-pixel groups).
//
// Same arithmetic as the generic kernel in v2e.cu (reference data/v2v_core_v2e.py:401-581), restructured the way
// esim_fast.cu is: 128-thread CTAs, 4 pixels per lane, everything that is the same for all pixels of a frame
// (t = k/fps, dt, dt/tau, the shot-noise scales) computed once per CTA into a shared-memory table instead of
// two float64 divisions per lane and frame, Philox with host-precomputed round keys, three generator calls per
// two intervals (shot uniforms per interval, leak normals per interval pair), Poisson inversion as straight-line
// code for k <= 2, and the single-threshold crossing handled without branches:
//   pe, ne in {0,1} as float64 built from the compare predicates (high-word selects), shot counts added as float64,
//   base += pe*pth; base -= ne*nth as separately rounded multiply and add (exact for any count, :547-548).
// Multi-threshold crossings take the exact floor-division path (per lane, only where |diff| >= 2*min threshold).
#include "v2e_common.cuh"

#include <cstdlib>

namespace v2v {
namespace {

constexpr int kThreads = 128;
constexpr int kPF = 4;             // frames per loop trip (two interval pairs)
constexpr int kMaxIntervals = 1024;

struct __align__(16) IntervalRow {  // one row per frame interval, identical for every pixel of the clip
  double dt;                        // t_k - t_{k-1}, t_k = k/fps                 (:442,577)
  double qdt;                       // dt / tau                                   (:167)
  float sps, sns;                   // float32 shot-noise scales of this frame    (:98-99)
  float pad[2];
};

constexpr int kLutBytes = 256 * 16, kFacBytes = 256 * 4, kLogfBytes = 256 * 4, kTrigBytes = kTrigEntries * 8;
constexpr int kRcpBytes = 4 * 2 * kThreads * 8;

__device__ __forceinline__ double hi_double(int hi) { return __hiloint2double(hi, 0); }
// x with the sign of s multiplied in (x >= 0): -x where s is negative — one logic op on the high word instead of a select pair
__device__ __forceinline__ double with_sign_of(double x, double s) {
  return __hiloint2double(__double2hiint(x) ^ (__double2hiint(s) & static_cast<int>(0x80000000u)), __double2loint(x));
}
__device__ __forceinline__ double with_sign_of(double x, float s) {
  return __hiloint2double(__double2hiint(x) ^ (__float_as_int(s) & static_cast<int>(0x80000000u)), __double2loint(x));
}
// float32 of an integer-valued float64 with |n| < 2^22, without the conversion unit: 2^52+2^51 + n leaves n as a two's
// complement integer in the low word; 1.5*2^23 + n is exact in float32 and its bit pattern is 0x4b400000 + n.
// `tiny` lanes (a threshold below 1e-5: counts could reach 2^22) take the plain conversion instead.
template <bool MAGIC>
__device__ __forceinline__ float small_int_to_float(double n, bool tiny) {
  if (!MAGIC) return static_cast<float>(n);
  if (tiny) {          // volatile: keeps this a branch, so the conversion is not issued speculatively on every step
    float r;
    asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(r) : "d"(n));
    return r;
  }
  const int i = __double2loint(__dadd_rn(n, 6755399441055744.0));
  return __fsub_rn(__int_as_float(0x4b400000 + i), 12582912.0f);
}

// floor(a/b) for a >= 0, b > 0 without branches (same candidate-and-correct scheme as floor_div_exact)
// The candidate is a*RN(1/b) rounded to the NEAREST integer with the 2^52 trick (two FP64 adds instead of a trip through
// the conversion unit): it is floor or floor+1 of a value within 1 ulp of the true quotient, still inside the +-1 window
// the two residual tests correct.
// (MAGIC: for the variant that is bound by the conversion unit; the others keep floor().)
template <bool MAGIC>
__device__ __forceinline__ double floor_div_bf(double a, double b, double rb) {
  const double p = __dmul_rn(a, rb);
  const double q = MAGIC ? __dsub_rn(__dadd_rn(p, 4503599627370496.0), 4503599627370496.0) : floor(p);
  const double qp = __dadd_rn(q, 1.0), qm = __dadd_rn(q, -1.0);
  const double r1 = __fma_rn(-q, b, a), r2 = __fma_rn(-qp, b, a);
  double out = (r2 >= 0.0) ? qp : q;
  out = (r1 < 0.0) ? qm : out;
  return out;
}

// BF: every pixel takes the exact floor division (no trigger, no divergent block) — for footage where several
// thresholds are crossed per frame in most warps (HDR-degraded clips); !BF: single-crossing fast path + divergent exact path.
template <bool F32STATE, bool CUTOFF, bool LEAK, bool SHOT, bool PHILOX, bool BF>
__global__ void __launch_bounds__(kThreads, 4) v2e_fast_kernel(const V2eArgs a) {
  extern __shared__ __align__(16) unsigned char dyn[];
  double2* lut2 = reinterpret_cast<double2*>(dyn);                                   // {double(logv), (v+20)/275}
  float* facf_s = reinterpret_cast<float*>(dyn + kLutBytes);                        // 1 - 0.75*inten as float
  float* logf_s = reinterpret_cast<float*>(dyn + kLutBytes + kFacBytes);            // float32 log LUT (float32-state path)
  float2* trig_s = reinterpret_cast<float2*>(dyn + kLutBytes + kFacBytes + kLogfBytes);
  double* rcp_s = reinterpret_cast<double*>(dyn + kLutBytes + kFacBytes + kLogfBytes + (LEAK && PHILOX ? kTrigBytes : 0));   // [4][2][kThreads] 1/thr
  IntervalRow* itab = reinterpret_cast<IntervalRow*>(reinterpret_cast<unsigned char*>(rcp_s) + (BF ? kRcpBytes : 0));
  int* fnum_s = reinterpret_cast<int*>(itab + (a.d.N - 1));                          // raw frame number of every frame of the clip

  const v2v_v2e_desc& d = a.d;
  // float32-state variant: conversion unit 71 % busy in ncu -> integer rounding and int->float by FP64/FP32 adds instead;
  // the float64-state variants are shorter of FP64 issue slots and keep floor() / cvt (same-box A/B: 0.66 vs 0.70 ms)
  constexpr bool kMagic = F32STATE;
  const int N = d.N, b = blockIdx.y;
  if (LEAK && PHILOX) fill_trig_table(trig_s);
  for (int i = threadIdx.x; i < 256; i += kThreads) {
    const int mv = v2e_mapped(a, b, i);                                                // degrade folded into the LUTs
    const double it = v2e_inten01(a, mv);                                              // :190
    lut2[i] = make_double2(static_cast<double>(d.lut[mv]), it);
    facf_s[i] = static_cast<float>(__dsub_rn(1.0, __dmul_rn(0.75, it)));              // :90
    logf_s[i] = d.lut[mv];
  }
  for (int n = threadIdx.x; n < N; n += kThreads) fnum_s[n] = v2e_frame_number(a, b, n);
  for (int i = 1 + threadIdx.x; i < N; i += kThreads) {
    IntervalRow r;
    const double t_k = __ddiv_rn(static_cast<double>(i), d.fps);                       // :577
    const double t_p = i == 1 ? 0.0 : __ddiv_rn(static_cast<double>(i - 1), d.fps);
    r.dt = __dsub_rn(t_k, t_p);                                                        // :442
    r.qdt = CUTOFF ? __ddiv_rn(r.dt, a.tau) : 0.0;                                     // :167
    r.sps = r.sns = r.pad[0] = r.pad[1] = 0.f;
    if (SHOT) {
      const int64_t si = static_cast<int64_t>(b) * (N - 1) + (i - 1);
      r.sps = v2e_scale_f32(d.shot_pos_scale[si]);
      r.sns = v2e_scale_f32(d.shot_neg_scale[si]);
    }
    itab[i - 1] = r;
  }
  __syncthreads();

  const int64_t pix0 = (static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x) * 4;
  const int64_t HW = a.HW;
  const bool want_stats = d.stats != nullptr;
  if (pix0 < HW) {
  const int64_t mp = static_cast<int64_t>(b) * HW + pix0;
  const uint64_t clip_id = d.clip_index_base + static_cast<uint64_t>(b);
  GroupStream gs{0u, 0u, 0u, 0u};
  if (PHILOX) gs = v2e_stream_init(static_cast<uint64_t>(pix0) >> 2, clip_id, a.rk);
  const bool leak_on = d.leak_rate_hz > 0.0;       // the leak variant also serves a float64 state with the leak switched off

  double pth[4], nth[4], lp[4], base[4];
  float basef[F32STATE ? 4 : 1], pthf[F32STATE ? 4 : 1], nthf[F32STATE ? 4 : 1];
  double r32d[LEAK ? 4 : 1];
  float ppf[SHOT ? 4 : 1], npf[SHOT ? 4 : 1];
  double thr2 = 1e300;
  {
    const double2 p01 = ld_stream_f64x2(d.pos_thres + mp), p23 = ld_stream_f64x2(d.pos_thres + mp + 2);
    const double2 n01 = ld_stream_f64x2(d.neg_thres + mp), n23 = ld_stream_f64x2(d.neg_thres + mp + 2);
    pth[0] = p01.x, pth[1] = p01.y, pth[2] = p23.x, pth[3] = p23.y;
    nth[0] = n01.x, nth[1] = n01.y, nth[2] = n23.x, nth[3] = n23.y;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    thr2 = fmin(thr2, fmin(pth[k], nth[k]));
    if (LEAK) r32d[LEAK ? k : 0] = static_cast<double>(__fmul_rn(a.leak_hz_f32, d.noise_rate ? d.noise_rate[mp + k] : 1.0f));   // :205
    if (SHOT) {                                                                      // :396-399
      ppf[SHOT ? k : 0] = v2e_pre_prob(d.pos_thres_nominal, pth[k]);
      npf[SHOT ? k : 0] = v2e_pre_prob(d.neg_thres_nominal, nth[k]);
    }
    if (F32STATE) {      // float32 diff compared with float64 thresholds: diff >= thr  <=>  diff >= RU_f32(thr)
      pthf[F32STATE ? k : 0] = __double2float_ru(pth[k]);
      nthf[F32STATE ? k : 0] = __double2float_ru(nth[k]);
    }
    if (BF) {            // own slots, read back by this lane only: no barrier
      rcp_s[(2 * k) * kThreads + threadIdx.x] = __drcp_rn(pth[k]);
      rcp_s[(2 * k + 1) * kThreads + threadIdx.x] = __drcp_rn(nth[k]);
    }
  }
  const bool tiny = !(thr2 >= 1e-5);  // absurdly small (or NaN) threshold on this lane: counts may not fit the small-integer shortcuts
  thr2 = __dadd_rn(thr2, thr2);       // below 2*min(threshold) of this lane's pixels at most one threshold is crossed
  const float thr2f = __double2float_rd(thr2);

  const uint8_t* fr = d.frames + static_cast<int64_t>(b) * a.Mraw * HW + pix0;
  auto frame_ptr = [&](int n) -> const uint8_t* { return fr + static_cast<int64_t>(fnum_s[n]) * HW; };
  {
    // first frame: lp = log_new; the filter runs with dt = 0 (eps = 0); base = lp   (:463-478)
    const uint32_t w0 = ld_stream_u32(frame_ptr(0));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double l0 = lut2[(w0 >> (8 * k)) & 0xffu].x;
      lp[k] = CUTOFF ? __dadd_rn(__dmul_rn(1.0, l0), __dmul_rn(0.0, l0)) : l0;
      base[k] = lp[k];
      if (F32STATE) basef[F32STATE ? k : 0] = static_cast<float>(l0);
    }
  }

  float* vox = d.voxel + static_cast<int64_t>(b) * a.T * d.num_bins * HW + pix0;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  double dpos = 0.0, dneg = 0.0;      // event totals of this lane (exact integers in float64)
  int sub = 0;
  const int fpb = d.frames_per_bin;
  const double jit = d.leak_jitter_fraction;

  auto byte_of = [](uint32_t w, int k) -> uint32_t {
    uint32_t v;
    asm("prmt.b32 %0, %1, 0, %2;" : "=r"(v) : "r"(w), "r"(0x4440u | static_cast<uint32_t>(k)));
    return v;
  };

  // one frame interval for the lane's 4 pixels; lz = leak jitter normals, (up, un) = shot uniforms
  auto step = [&](const uint32_t w, const int iv, const float (&lz)[4], const float (&up)[4], const float (&un)[4]) {
    const IntervalRow row = itab[iv];
    float outv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t v = byte_of(w, k);
      double pe, ne;          // threshold-crossing counts as float64 (:55-60)
      float of;               // pe - ne as float32 (the voxel contribution), kept beside them to stay off the conversion unit
      if (F32STATE) {
        const float lognew = logf_s[v];                                               // :447 (lp == log_new)
        const float df = __fsub_rn(lognew, basef[F32STATE ? k : 0]);                    // :503 in float32
        if (BF) {
          const bool dn = df < 0.f;
          const double thr = dn ? nth[k] : pth[k];
          const double q = floor_div_bf<kMagic>(static_cast<double>(fabsf(df)), thr, rcp_s[(2 * k + (dn ? 1 : 0)) * kThreads + threadIdx.x]);
          const double t = __dmul_rn(q, thr);                                          // pe*pth or ne*nth (the other product is 0)
          const double sq = with_sign_of(q, df), st = with_sign_of(t, df);
          basef[F32STATE ? k : 0] = __double2float_rn(__dadd_rn(static_cast<double>(basef[F32STATE ? k : 0]), st));   // :547-548
          if (want_stats) {
            dpos = __dadd_rn(dpos, q);           // here: dpos = #pos + #neg, dneg = #pos - #neg (resolved after the loop)
            dneg = __dadd_rn(dneg, sq);
          }
          acc[k] += small_int_to_float<kMagic>(sq, tiny);
          outv[k] = acc[k];
          continue;
        }
        const bool upc = df >= pthf[F32STATE ? k : 0], dnc = -df >= nthf[F32STATE ? k : 0];
        pe = hi_double(upc ? 0x3ff00000 : 0);
        ne = hi_double(dnc ? 0x3ff00000 : 0);
        of = upc ? 1.0f : 0.0f;
        of = dnc ? -1.0f : of;
        if (fabsf(df) >= thr2f) {                                                      // possibly several thresholds: exact path
          const double diff = static_cast<double>(df);
          if (diff >= pth[k]) pe = count_floor(diff, pth[k]);
          else if (-diff >= nth[k]) ne = count_floor(-diff, nth[k]);
          of = static_cast<float>(__dsub_rn(pe, ne));
        }
        // :547-548 — in place on a float32 array, the float64 sum cast back after each line.  Without shot noise at
        // most one of pe, ne is non-zero, the other line is the identity: one add of (pe*pth - ne*nth), one cast.
        const double s = __dsub_rn(__dmul_rn(pe, pth[k]), __dmul_rn(ne, nth[k]));
        basef[F32STATE ? k : 0] = __double2float_rn(__dadd_rn(static_cast<double>(basef[F32STATE ? k : 0]), s));
      } else {
        const double2 lv = lut2[v];                                                    // {log_new, inten01}
        if (CUTOFF) {                                                                  // :157-173
          double eps = __dmul_rn(lv.y, row.qdt);
          eps = fmin(eps, 1.0);
          lp[k] = __dadd_rn(__dmul_rn(__dsub_rn(1.0, eps), lp[k]), __dmul_rn(eps, lv.x));
        } else {
          lp[k] = lv.x;
        }
        if (LEAK) {                                                                    // :192-211
          const double lr = PHILOX ? static_cast<double>(lz[k]) : 0.0;
          const double rate = __dmul_rn(r32d[LEAK ? k : 0], __dsub_rn(1.0, __dmul_rn(jit, lr)));
          base[k] = __dsub_rn(base[k], __dmul_rn(__dmul_rn(row.dt, rate), pth[k]));
        }
        const double diff = __dsub_rn(lp[k], base[k]);                                 // :503
        if (BF) {
          const bool dn = diff < 0.0;
          const double thr = dn ? nth[k] : pth[k];
          const double q = floor_div_bf<kMagic>(fabs(diff), thr, rcp_s[(2 * k + (dn ? 1 : 0)) * kThreads + threadIdx.x]);    // :55-60
          if (!SHOT) {
            const double t = __dmul_rn(q, thr);                                        // pe*pth or ne*nth (the other product is 0)
            base[k] = __dadd_rn(base[k], with_sign_of(t, diff));                       // :547-548
            const double sq = with_sign_of(q, diff);
            if (want_stats) {
              dpos = __dadd_rn(dpos, q);         // here: dpos = #pos + #neg, dneg = #pos - #neg (resolved after the loop)
              dneg = __dadd_rn(dneg, sq);
            }
            acc[k] += small_int_to_float<kMagic>(sq, tiny);
            outv[k] = acc[k];
            continue;
          }
          pe = dn ? 0.0 : q;
          ne = dn ? q : 0.0;
          of = small_int_to_float<kMagic>(with_sign_of(q, diff), tiny);
        } else {
        const bool upc = diff >= pth[k], dnc = -diff >= nth[k];
        pe = hi_double(upc ? 0x3ff00000 : 0);
        ne = hi_double(dnc ? 0x3ff00000 : 0);
        of = upc ? 1.0f : 0.0f;
        of = dnc ? -1.0f : of;
        if (fabs(diff) >= thr2) {                                                      // possibly several thresholds: exact path
          if (upc) pe = count_floor(diff, pth[k]);
          else if (dnc) ne = count_floor(-diff, nth[k]);
          of = static_cast<float>(__dsub_rn(pe, ne));
        }
        }
      }
      if (SHOT) {                                                                      // :90-103, 530-531
        const float fac = facf_s[v];
        const float lam_p = v2e_shot_lambda(fac, ppf[SHOT ? k : 0], row.sps), lam_n = v2e_shot_lambda(fac, npf[SHOT ? k : 0], row.sns);
        const PoissonCdf cp = poisson_cdf(lam_p), cn = poisson_cdf(lam_n);
        // k in {0,1,2} from two compares, as the high word of a float64 (0.0, 1.0, 2.0) whose exponent bits are also the
        // float32 of the same value; k >= 3 (about lam^3/6 of the draws) takes the loop
        const int hp = up[k] >= cp.c1 ? 0x40000000 : (up[k] >= cp.c0 ? 0x3ff00000 : 0);
        const int hn = un[k] >= cn.c1 ? 0x40000000 : (un[k] >= cn.c0 ? 0x3ff00000 : 0);
        double kp = hi_double(hp), kn = hi_double(hn);
        float kpf = __int_as_float(hp & 0x7f800000), knf = __int_as_float(hn & 0x7f800000);
        if (up[k] >= cp.c2) {
          const int kk = poisson_tail(lam_p, up[k], cp.p2, cp.c2);
          kp = static_cast<double>(kk);
          kpf = static_cast<float>(kk);
        }
        if (un[k] >= cn.c2) {
          const int kk = poisson_tail(lam_n, un[k], cn.p2, cn.c2);
          kn = static_cast<double>(kk);
          knf = static_cast<float>(kk);
        }
        pe = __dadd_rn(pe, kp);
        ne = __dadd_rn(ne, kn);
        of = __fadd_rn(of, __fsub_rn(kpf, knf));
      }
      if (!F32STATE) {
        // :547-548, unconditionally: a zero count adds 0*thr = 0 and leaves base untouched
        base[k] = __dadd_rn(base[k], __dmul_rn(pe, pth[k]));
        base[k] = __dsub_rn(base[k], __dmul_rn(ne, nth[k]));
      }
      if (want_stats) {
        dpos = __dadd_rn(dpos, pe);
        dneg = __dadd_rn(dneg, ne);
      }
      acc[k] += of;                                                                     // :579-580 (integers: exact in float32)
      outv[k] = acc[k];
    }
    if (++sub == fpb) {
      sub = 0;
      st_stream_f32x4(vox, outv[0], outv[1], outv[2], outv[3]);
      vox += HW;
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] = 0.f;
    }
  };

  const int M = N - 1;
  if (SHOT) {
    // ---- shot-noise variants: the step is ~600 instructions, so the loop stays rolled (one copy of the step: the unrolled
    //      form overflows the instruction cache); three frames in flight through a rotating register window ----
    uint32_t w0 = ld_stream_u32(frame_ptr(1)), w1 = M > 1 ? ld_stream_u32(frame_ptr(2)) : 0u, w2 = M > 2 ? ld_stream_u32(frame_ptr(3)) : 0u;
    float le[4] = {0.f, 0.f, 0.f, 0.f}, lo[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int j = 0; j < M; ++j) {                // interval j = frames (j, j+1)
      const uint32_t w = w0;
      w0 = w1;
      w1 = w2;
      if (j + 4 < N) w2 = ld_stream_u32(frame_ptr(j + 4));
      float lz[4] = {0.f, 0.f, 0.f, 0.f};
      if (LEAK && PHILOX && leak_on) {
        if ((j & 1) == 0) v2e_leak_normals(gs, trig_s, le, lo);
#pragma unroll
        for (int k = 0; k < 4; ++k) lz[k] = (j & 1) ? lo[k] : le[k];
      }
      float up[4], un[4];
      v2e_shot_uniforms(gs, up, un);
      step(w, j, lz, up, un);
    }
  } else {
  // ---- main loop: kPF frames per trip, the next trip's words already in flight ----
  const int trips = M / kPF;
  uint32_t cur[kPF], nxt[kPF];
  if (trips > 0) {
#pragma unroll
    for (int u = 0; u < kPF; ++u) cur[u] = ld_stream_u32(frame_ptr(1 + u));
  }
  const float one4[4] = {1.f, 1.f, 1.f, 1.f};
  int i = 1;
  for (int t = 0; t < trips; ++t) {
    if (t + 1 < trips) {
#pragma unroll
      for (int u = 0; u < kPF; ++u) nxt[u] = ld_stream_u32(frame_ptr(i + kPF + u));
    }
#pragma unroll
    for (int h = 0; h < kPF / 2; ++h) {           // interval pairs (i-1+2h, i+2h)
      float le[4] = {0.f, 0.f, 0.f, 0.f}, lo[4] = {0.f, 0.f, 0.f, 0.f};
      if (LEAK && PHILOX && leak_on) v2e_leak_normals(gs, trig_s, le, lo);
      step(cur[2 * h], i - 1 + 2 * h, le, one4, one4);
      step(cur[2 * h + 1], i + 2 * h, lo, one4, one4);
    }
    i += kPF;
#pragma unroll
    for (int u = 0; u < kPF; ++u) cur[u] = nxt[u];
  }
  {
    float le[4] = {0.f, 0.f, 0.f, 0.f}, lo[4] = {0.f, 0.f, 0.f, 0.f}, lz[4];
    for (; i < N; ++i) {                                                 // ragged tail (< kPF intervals; starts at an even interval)
      if (LEAK && PHILOX && leak_on && ((i - 1) & 1) == 0) v2e_leak_normals(gs, trig_s, le, lo);
#pragma unroll
      for (int k = 0; k < 4; ++k) lz[k] = ((i - 1) & 1) ? lo[k] : le[k];
      step(ld_stream_u32(frame_ptr(i)), i - 1, lz, one4, one4);
    }
  }
  }

  if (want_stats) {
    const unsigned int m = __activemask();
    if (BF && !SHOT) {                 // (total, net) -> (#pos, #neg); exact: both are integers below 2^53
      const double tot = dpos, net = dneg;
      dpos = __dmul_rn(__dadd_rn(tot, net), 0.5);
      dneg = __dmul_rn(__dsub_rn(tot, net), 0.5);
    }
    unsigned long long sp = static_cast<unsigned long long>(dpos), sn = static_cast<unsigned long long>(dneg);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long op = __shfl_xor_sync(m, sp, o), on = __shfl_xor_sync(m, sn, o);
      const bool ok = ((m >> ((threadIdx.x & 31) ^ o)) & 1u) != 0;      // partner lane is active
      sp += ok ? op : 0ull;
      sn += ok ? on : 0ull;
    }
    if ((threadIdx.x & 31) == (__ffs(m) - 1)) {      // one pair of global reductions per warp: no CTA barrier, no shared stage
      if (sp) atomicAdd(reinterpret_cast<unsigned long long*>(d.stats + 2 * b), sp);
      if (sn) atomicAdd(reinterpret_cast<unsigned long long*>(d.stats + 2 * b) + 1, sn);
    }
  }
  }  // valid
}

size_t fast_smem_bytes(const V2eArgs& a, bool trig, bool bf) {
  return kLutBytes + kFacBytes + kLogfBytes + (trig ? kTrigBytes : 0) + (bf ? kRcpBytes : 0) + static_cast<size_t>(a.d.N - 1) * sizeof(IntervalRow) +
         static_cast<size_t>(a.d.N) * sizeof(int);
}

}  // namespace

bool v2e_fast_eligible(const V2eArgs& a) {
  const v2v_v2e_desc& d = a.d;
  if (d.noise_mode == V2V_NOISE_EXPLICIT || d.thres_per_interval) return false;
  if (a.HW % 4 != 0 || !aligned(d.frames, 4) || !aligned(d.voxel, 16) || !aligned(d.pos_thres, 16) || !aligned(d.neg_thres, 16)) return false;
  if (d.N - 1 > kMaxIntervals) return false;
  const bool sh = d.shot_noise_rate_hz > 0.0 && d.noise_mode != V2V_NOISE_NONE;
  if (d.state_f32 && sh) return false;          // float32 state with shot noise: two float32 roundings per update, generic kernel
  if (d.kernel_flags & V2V_V2E_FLAG_FAST) return true;      // tests: force the fast kernel on small shapes
  return static_cast<int64_t>(d.B) * a.HW >= 148LL * 2048;
}

int launch_v2e_fast(const V2eArgs& a, cudaStream_t s) {
  const v2v_v2e_desc& d = a.d;
  const int64_t groups = a.HW / 4;
  dim3 grid(static_cast<unsigned int>((groups + kThreads - 1) / kThreads), static_cast<unsigned int>(d.B));
  const bool ph = d.noise_mode == V2V_NOISE_PHILOX;
  const bool cut = d.cutoff_hz > 0.0, lk = d.leak_rate_hz > 0.0, sh = d.shot_noise_rate_hz > 0.0 && ph;
  const bool leak_variant = !d.state_f32 && (lk || !cut);      // the template's LEAK (a float64 state without cutoff runs the leak variant)
  // exact division for every pixel (default) or single-crossing fast path + divergent exact path (V2V_V2E_BF=0)
  const bool bf = !(d.kernel_flags & V2V_V2E_FLAG_DIVERGENT_DIV);
  const size_t smem = fast_smem_bytes(a, leak_variant && ph, bf);
#define V2V_KB(F32, CU, LK, SH, PH, BFV)                                                                                    \
  do {                                                                                                                      \
    V2V_CUDA(cudaFuncSetAttribute(v2e_fast_kernel<F32, CU, LK, SH, PH, BFV>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))); \
    v2e_fast_kernel<F32, CU, LK, SH, PH, BFV><<<grid, kThreads, smem, s>>>(a);                                              \
  } while (0)
#define V2V_K(F32, CU, LK, SH, PH)                                          \
  do {                                                                      \
    if (bf) V2V_KB(F32, CU, LK, SH, PH, true); else V2V_KB(F32, CU, LK, SH, PH, false); \
  } while (0)
#define V2V_PHX(F32, CU, LK)                                                     \
  do {                                                                           \
    if (ph) { if (sh) V2V_K(F32, CU, LK, true, true); else V2V_K(F32, CU, LK, false, true); } \
    else V2V_K(F32, CU, LK, false, false);                                       \
  } while (0)
  if (d.state_f32) V2V_K(true, false, false, false, false);
  else if (cut && lk) V2V_PHX(false, true, true);
  else if (cut) V2V_PHX(false, true, false);
  else V2V_PHX(false, false, true);
#undef V2V_PHX
#undef V2V_K
#undef V2V_KB
  count_launch();
  V2V_CUDA(cudaGetLastError());
  return V2V_OK;
}

}  // namespace v2v

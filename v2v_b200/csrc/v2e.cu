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
#include "common.cuh"

namespace v2v {
namespace {

constexpr int kV2eThreads = 256;
constexpr int kPF = 4;

struct V2eArgs {
  v2v_v2e_desc d;
  int64_t HW;
  int32_t T, G;
  double tau;          // 1/(2*pi*cutoff)
  float leak_hz_f32;   // leak_rate_hz as the float32 it becomes in `leak_rate_hz*noise_rate_array`
};

__device__ __forceinline__ double count_floor(double a, double thr, double rthr) {
  // np.floor_divide(max(diff,0), thr) for a >= 0, thr > 0
  if (a < thr) return 0.0;
  if (a < __dadd_rn(thr, thr)) return 1.0;
  return floor_div_exact(a, thr, rthr);
}

__device__ __forceinline__ int poisson_small(double lam, double u) {
  // inversion; lam is a fraction of an event per frame in every shipped preset
  if (!(lam > 0.0)) return 0;
  double p = exp(-lam), cdf = p;
  int k = 0;
  while (u > cdf && k < 64) {
    ++k;
    p *= lam / k;
    cdf += p;
  }
  return k;
}


// Per pixel / interval Philox draws of the v2e model: leak jitter normal and the two Poisson uniforms.
__device__ __forceinline__ void v2e_philox_draw(uint64_t px, uint32_t interval, uint64_t clip_id, uint2 key,
                                                float* leak_z, double* u_pos, double* u_neg) {
  const uint4 r = Philox::run(make_uint4(static_cast<uint32_t>(px), interval, static_cast<uint32_t>(clip_id),
                                         0x40000000u | (static_cast<uint32_t>(px >> 32) & 0x3fffu) << 16 |
                                             static_cast<uint32_t>((clip_id >> 32) & 0xffffu)),
                              key);
  *leak_z = box_muller(r.x, r.y).x;
  *u_pos = static_cast<double>(r.z) * (1.0 / 4294967296.0);
  *u_neg = static_cast<double>(r.w) * (1.0 / 4294967296.0);
}

__device__ __forceinline__ double v2e_shot_lambda(uint32_t v, double pre_prob, double scale) {
  const double inten = __ddiv_rn(__dadd_rn(static_cast<double>(v), 20.0), 275.0);     // :190
  const double fac = __dsub_rn(1.0, __dmul_rn(0.75, inten));                          // :90
  return fac * pre_prob * scale;                                                      // :91-99
}

__device__ __forceinline__ uint32_t load_pix4(const uint8_t* p) { return ld_stream_u32(p); }

template <int P, bool F32STATE>
__global__ void __launch_bounds__(kV2eThreads) v2e_kernel(const V2eArgs a) {
  __shared__ float lut_s[256];
  const v2v_v2e_desc& d = a.d;
  for (int i = threadIdx.x; i < 256; i += kV2eThreads) lut_s[i] = d.lut[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int64_t pix0 = (static_cast<int64_t>(blockIdx.x) * kV2eThreads + threadIdx.x) * P;
  const int64_t HW = a.HW;
  if (pix0 >= HW) return;
  const int N = d.N;
  const int64_t mp = static_cast<int64_t>(b) * HW + pix0;
  const bool cutoff = d.cutoff_hz > 0.0, leak = d.leak_rate_hz > 0.0, shot = d.shot_noise_rate_hz > 0.0;
  const bool need_inten = cutoff || shot;
  const uint64_t clip_id = d.clip_index_base + static_cast<uint64_t>(b);
  const uint2 key = make_uint2(static_cast<uint32_t>(d.seed), static_cast<uint32_t>(d.seed >> 32));

  double pth[P], nth[P], rp[P], rn[P], lp[P], base[P], ppp[P], npp[P];
  float nrate[P];
#pragma unroll
  for (int k = 0; k < P; ++k) {
    pth[k] = d.pos_thres[mp + k];
    nth[k] = d.neg_thres[mp + k];
    rp[k] = __drcp_rn(pth[k]);
    rn[k] = __drcp_rn(nth[k]);
    nrate[k] = (leak && d.noise_rate) ? d.noise_rate[mp + k] : 1.0f;
    ppp[k] = __ddiv_rn(d.pos_thres_nominal, pth[k]);     // :396-399
    npp[k] = __ddiv_rn(d.neg_thres_nominal, nth[k]);
  }

  const uint8_t* fr = d.frames + static_cast<int64_t>(b) * N * HW + pix0;
  auto load = [&](int i) -> uint32_t { return P == 4 ? load_pix4(fr + static_cast<int64_t>(i) * HW)
                                                     : static_cast<uint32_t>(ld_stream_u8(fr + static_cast<int64_t>(i) * HW)); };
  {
    // first frame: lp = log_new; the filter runs with dt = 0 (eps = 0); base = lp   (:463-478)
    const uint32_t w0 = load(0);
#pragma unroll
    for (int k = 0; k < P; ++k) {
      const double l0 = static_cast<double>(lut_s[(w0 >> (8 * k)) & 0xffu]);
      lp[k] = cutoff ? __dadd_rn(__dmul_rn(1.0, l0), __dmul_rn(0.0, l0)) : l0;
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
      const double t_k = __ddiv_rn(static_cast<double>(i), d.fps);     // :577
      const double dt = __dsub_rn(t_k, t_prev);                         // :442
      t_prev = t_k;
      const int64_t fo = (static_cast<int64_t>(b) * (N - 1) + (i - 1)) * HW + pix0;

      // per-interval random fields
      double lr[P];
      int sp[P], sn[P];
#pragma unroll
      for (int k = 0; k < P; ++k) { lr[k] = 0.0; sp[k] = 0; sn[k] = 0; }
      if (d.noise_mode == V2V_NOISE_EXPLICIT) {
        if (leak && d.leak_randn) {
#pragma unroll
          for (int k = 0; k < P; ++k) lr[k] = d.leak_randn[fo + k];
        }
        if (shot && d.pos_shot && d.neg_shot) {
#pragma unroll
          for (int k = 0; k < P; ++k) { sp[k] = d.pos_shot[fo + k]; sn[k] = d.neg_shot[fo + k]; }
        }
      }

      float outv[P];
#pragma unroll
      for (int k = 0; k < P; ++k) {
        const uint32_t v = (w >> (8 * k)) & 0xffu;
        const float lognew = lut_s[v];                                              // :447
        double inten = 0.0;
        if (need_inten) inten = __ddiv_rn(__dadd_rn(static_cast<double>(v), 20.0), 275.0);   // :190
        if (cutoff) {                                                               // :157-173
          double eps = __dmul_rn(inten, __ddiv_rn(dt, a.tau));
          eps = fmin(eps, 1.0);
          lp[k] = __dadd_rn(__dmul_rn(__dsub_rn(1.0, eps), lp[k]), __dmul_rn(eps, static_cast<double>(lognew)));
        } else {
          lp[k] = static_cast<double>(lognew);
        }
        if (d.noise_mode == V2V_NOISE_PHILOX && (leak || shot)) {
          float lz;
          double up, un;
          v2e_philox_draw(static_cast<uint64_t>(pix0 + k), static_cast<uint32_t>(i - 1), clip_id, key, &lz, &up, &un);
          if (leak) lr[k] = static_cast<double>(lz);
          if (shot) {
            const int64_t si = static_cast<int64_t>(b) * (N - 1) + (i - 1);
            sp[k] = poisson_small(v2e_shot_lambda(v, ppp[k], d.shot_pos_scale[si]), up);
            sn[k] = poisson_small(v2e_shot_lambda(v, npp[k], d.shot_neg_scale[si]), un);
          }
        }
        if (leak) {                                                                 // :192-211
          const float r32 = __fmul_rn(a.leak_hz_f32, nrate[k]);
          const double rate = __dmul_rn(static_cast<double>(r32), __dsub_rn(1.0, __dmul_rn(d.leak_jitter_fraction, lr[k])));
          base[k] = __dsub_rn(base[k], __dmul_rn(__dmul_rn(dt, rate), pth[k]));
        }
        double diff;                                                                // :503
        if (F32STATE) diff = static_cast<double>(__fsub_rn(static_cast<float>(lp[k]), static_cast<float>(base[k])));
        else diff = __dsub_rn(lp[k], base[k]);
        double pe = 0.0, ne = 0.0;                                                  // :55-60
        if (diff > 0.0) pe = count_floor(diff, pth[k], rp[k]);
        else if (diff < 0.0) ne = count_floor(-diff, nth[k], rn[k]);
        pe += static_cast<double>(sp[k]);                                           // :530-531
        ne += static_cast<double>(sn[k]);
        // :547-548 — in place: the float64 sum is cast back to the state dtype after each line
        double nb = __dadd_rn(base[k], __dmul_rn(pe, pth[k]));
        if (F32STATE) nb = static_cast<double>(__double2float_rn(nb));
        nb = __dsub_rn(nb, __dmul_rn(ne, nth[k]));
        if (F32STATE) nb = static_cast<double>(__double2float_rn(nb));
        base[k] = nb;
        npos += static_cast<unsigned int>(pe);
        nneg += static_cast<unsigned int>(ne);
        acc[k] += static_cast<int>(pe) - static_cast<int>(ne);                      // :579-580
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
__global__ void __launch_bounds__(256) v2e_shot_scale_kernel(const V2eArgs a, double* pos_scale, double* neg_scale) {
  const v2v_v2e_desc& d = a.d;
  const int k = blockIdx.x + 1, b = blockIdx.y;
  const uint8_t* fr = d.frames + (static_cast<int64_t>(b) * d.N + k) * a.HW;
  const double* pt = d.pos_thres + static_cast<int64_t>(b) * a.HW;
  const double* nt = d.neg_thres + static_cast<int64_t>(b) * a.HW;
  double sp = 0.0, sn = 0.0;
  for (int64_t i = threadIdx.x; i < a.HW; i += 256) {
    const double inten = (static_cast<double>(fr[i]) + 20.0) / 275.0;
    const double fac = 1.0 - 0.75 * inten;
    sp += fac * (d.pos_thres_nominal / pt[i]);
    sn += fac * (d.neg_thres_nominal / nt[i]);
  }
  __shared__ double red[2][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sp += __shfl_xor_sync(0xffffffffu, sp, o);
    sn += __shfl_xor_sync(0xffffffffu, sn, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sp; red[1][threadIdx.x >> 5] = sn; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tp = 0.0, tn = 0.0;
    for (int i = 0; i < 8; ++i) { tp += red[0][i]; tn += red[1][i]; }
    const double dt = static_cast<double>(k) / d.fps - static_cast<double>(k - 1) / d.fps;
    const double sf = (d.shot_noise_rate_hz / 2) * dt;
    const int64_t o = static_cast<int64_t>(b) * (d.N - 1) + (k - 1);
    pos_scale[o] = sf / (tp / static_cast<double>(a.HW));
    neg_scale[o] = sf / (tn / static_cast<double>(a.HW));
  }
}


// Audit hook: the random fields a PHILOX run draws, for explicit replay / oracle checks.
__global__ void v2e_philox_fields_kernel(const V2eArgs a, double* leak_randn, int32_t* pos_shot, int32_t* neg_shot) {
  const v2v_v2e_desc& d = a.d;
  const int b = blockIdx.y;
  const int64_t pix = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (pix >= a.HW) return;
  const uint64_t clip_id = d.clip_index_base + static_cast<uint64_t>(b);
  const uint2 key = make_uint2(static_cast<uint32_t>(d.seed), static_cast<uint32_t>(d.seed >> 32));
  const bool shot = d.shot_noise_rate_hz > 0.0;
  const int64_t mp = static_cast<int64_t>(b) * a.HW + pix;
  const double ppp = __ddiv_rn(d.pos_thres_nominal, d.pos_thres[mp]), npp = __ddiv_rn(d.neg_thres_nominal, d.neg_thres[mp]);
  for (int i = 1; i < d.N; ++i) {
    float lz;
    double up, un;
    v2e_philox_draw(static_cast<uint64_t>(pix), static_cast<uint32_t>(i - 1), clip_id, key, &lz, &up, &un);
    const int64_t o = (static_cast<int64_t>(b) * (d.N - 1) + (i - 1)) * a.HW + pix;
    if (leak_randn) leak_randn[o] = static_cast<double>(lz);
    if (shot && pos_shot && neg_shot) {
      const uint32_t v = d.frames[(static_cast<int64_t>(b) * d.N + i) * a.HW + pix];
      const int64_t si = static_cast<int64_t>(b) * (d.N - 1) + (i - 1);
      pos_shot[o] = poisson_small(v2e_shot_lambda(v, ppp, d.shot_pos_scale[si]), up);
      neg_shot[o] = poisson_small(v2e_shot_lambda(v, npp, d.shot_neg_scale[si]), un);
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
  V2V_REQUIRE(!(d.state_f32 && (d.cutoff_hz > 0.0 || d.leak_rate_hz > 0.0)), V2V_ERR_INVALID_ARG,
              "state_f32 is only meaningful with cutoff_hz<=0 and leak_rate_hz<=0");
  if (d.noise_mode == V2V_NOISE_PHILOX && d.shot_noise_rate_hz > 0.0)
    V2V_REQUIRE(d.shot_pos_scale && d.shot_neg_scale, V2V_ERR_INVALID_ARG, "PHILOX shot noise needs the scales from v2v_v2e_shot_scales");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool vec4 = (a.HW % 4 == 0) && aligned(d.frames, 4) && aligned(d.voxel, 16) &&
                    static_cast<int64_t>(d.B) * a.HW >= 148LL * 2048;
  const int P = vec4 ? 4 : 1;
  dim3 grid(static_cast<unsigned int>(((a.HW + P - 1) / P + kV2eThreads - 1) / kV2eThreads), static_cast<unsigned int>(d.B));
  if (vec4) {
    if (d.state_f32) v2e_kernel<4, true><<<grid, kV2eThreads, 0, s>>>(a);
    else v2e_kernel<4, false><<<grid, kV2eThreads, 0, s>>>(a);
  } else {
    if (d.state_f32) v2e_kernel<1, true><<<grid, kV2eThreads, 0, s>>>(a);
    else v2e_kernel<1, false><<<grid, kV2eThreads, 0, s>>>(a);
  }
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
  dim3 grid(static_cast<unsigned int>(d.N - 1), static_cast<unsigned int>(d.B));
  v2e_shot_scale_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, shot_pos_scale, shot_neg_scale);
  count_launch();
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

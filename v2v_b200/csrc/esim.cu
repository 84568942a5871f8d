// ESIM-style frames -> voxel, fused single pass (sm_100a).
//
// Replaces EventEmulator.video_to_voxel (reference data/v2v_core_esim.py:26-69),
// the bin accumulation of data/v2v_datasets.py:399-400 and the float32 packing
// of :328-356.  One thread owns P consecutive pixels of one clip for the whole
// frame sequence: the float64 potential, the previous log intensity and the
// thresholds live in registers; frames stream in as coalesced 32-bit words with
// a register ring of PF frames in flight, voxels stream out as 128-bit stores.
// Log intensities come from the 256-entry float64 LUT the host built with the
// reference's NumPy expression (never recomputed here: bit parity).  HBM bound:
// 1 byte in + 4/frames_per_bin bytes out per pixel-interval.
#include "esim_common.cuh"


namespace v2v {
namespace {

constexpr int kThreads = kEsimThreads;

template <int P>
struct PixWord;
template <>
struct PixWord<4> {
  using type = uint32_t;
  static __device__ __forceinline__ uint32_t load(const uint8_t* p) { return ld_stream_u32(p); }
};
template <>
struct PixWord<1> {
  using type = uint32_t;
  static __device__ __forceinline__ uint32_t load(const uint8_t* p) { return ld_stream_u8(p); }
};

template <int P, int NOISE, bool EXTERNAL, bool PERPIXEL, int PF, int LUTC>
__global__ void __launch_bounds__(kThreads) esim_kernel(const EsimArgs a) {
  __shared__ double lut_s[256 * LUTC];
  __shared__ uint2 dir_s[NOISE == V2V_NOISE_PHILOX ? kDirEntries : 1];
  if (NOISE == V2V_NOISE_PHILOX) fill_dir_table(dir_s);
  const v2v_esim_desc& d = a.d;
  __shared__ float f255_s[256];                                    // (mapped value)/255 of the ground-truth frame output
  {
    const uint8_t* vmap = d.value_map ? d.value_map + static_cast<int64_t>(blockIdx.y) * 256 : nullptr;   // degrade folded into the LUTs
    for (int i = threadIdx.x; i < 256 * LUTC; i += kThreads) lut_s[i] = d.lut[vmap ? vmap[i / LUTC] : i / LUTC];
    for (int i = threadIdx.x; i < 256; i += kThreads) f255_s[i] = __fdiv_rn(static_cast<float>(vmap ? vmap[i] : i), 255.0f);
  }
  __syncthreads();

  const int b = blockIdx.y;
  const int64_t grp = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
  const int64_t pix0 = grp * P;
  const int64_t HW = a.HW;
  if (pix0 >= HW) return;
  const int lutc = threadIdx.x & (LUTC - 1);
  const int N = d.N;
  const int64_t clip_pix = static_cast<int64_t>(b) * HW + pix0;   // offset in [B,H,W] maps
  const uint64_t clip_id = d.clip_index_base + static_cast<uint64_t>(b);
  const NoiseKey nkey = make_noise_key(clip_id);
  const NoiseScale nsc = make_noise_scale(static_cast<float>((NOISE != V2V_NOISE_NONE && d.base_noise_std) ? d.base_noise_std[b] : 0.0));
  NoiseStream gs{0u, 0u, 1u};                 // base-noise stream of this lane's 4-pixel group (every lane of a group walks the same stream)
  if (NOISE == V2V_NOISE_PHILOX) gs = noise_stream_init(static_cast<uint64_t>(pix0) >> 2, nkey, a.rk);

  // ---- per-pixel / per-clip constants ----
  double pos[PERPIXEL ? P : 1], neg[PERPIXEL ? P : 1], rpos[PERPIXEL ? P : 1], rneg[PERPIXEL ? P : 1];
  if (PERPIXEL) {
#pragma unroll
    for (int k = 0; k < (PERPIXEL ? P : 1); ++k) {
      pos[k] = d.pos_thres[clip_pix + k];
      neg[k] = d.neg_thres[clip_pix + k];
    }
  } else {
    pos[0] = d.pos_thres[b];
    neg[0] = d.neg_thres[b];
  }
#pragma unroll
  for (int k = 0; k < (PERPIXEL ? P : 1); ++k) {
    rpos[k] = __drcp_rn(pos[k]);
    rneg[k] = __drcp_rn(neg[k]);
  }
  const double nstd = (NOISE != V2V_NOISE_NONE && d.base_noise_std) ? d.base_noise_std[b] : 0.0;

  // ---- initial state ----
  double pot[P], lprev[P], hot[P];
  // frame n of the clip = raw frame frame_index[b][n] (pause gather, data/v2v_datasets.py:285-311), or n itself
  const int Mraw = a.Mraw;
  const int32_t* fidx = d.frame_index ? d.frame_index + static_cast<int64_t>(b) * N : nullptr;
  const uint8_t* fr = d.frames + (static_cast<int64_t>(b) * Mraw) * HW + pix0;
  auto frame_ptr = [&](int n) -> const uint8_t* {
    const int r = fidx ? min(max(fidx[n], 0), Mraw - 1) : n;
    return fr + static_cast<int64_t>(r) * HW;
  };
  {
    uint32_t w0 = PixWord<P>::load(frame_ptr(0));
#pragma unroll
    for (int k = 0; k < P; ++k) lprev[k] = lut_s[((w0 >> (8 * k)) & 0xffu) * LUTC + lutc];
  }
#pragma unroll
  for (int k = 0; k < P; ++k) {
    const double pk = pos[PERPIXEL ? k : 0], nk = neg[PERPIXEL ? k : 0];
    hot[k] = 0.0;
    double u = -1.0;
    if (NOISE == V2V_NOISE_PHILOX) {
      philox_init_pixel(static_cast<uint64_t>(pix0 + k), nkey, a.rk, d.hot_pixel_fraction[b],
                        static_cast<float>(d.hot_pixel_std[b]), &u, &hot[k]);
    } else if (NOISE == V2V_NOISE_EXPLICIT) {
      if (d.hot_noise) hot[k] = d.hot_noise[clip_pix + k];
    }
    if (d.u0) u = d.u0[clip_pix + k];
    if (d.potential_in) {
      pot[k] = d.potential_in[clip_pix + k];
    } else if (u >= 0.0) {
      // data/v2v_core_esim.py:29: rand*(pos+neg) - neg, three separately rounded ops
      pot[k] = __dsub_rn(__dmul_rn(u, __dadd_rn(pk, nk)), nk);
    } else {
      pot[k] = 0.0;
    }
  }

  // ---- output addressing (constant across frames) ----
  int64_t out_off;
  if (a.padded) {
    const int64_t row = pix0 / d.W, col = pix0 - row * d.W;
    out_off = row * a.row_stride + col;
  } else {
    out_off = pix0;
  }
  float* vox = d.voxel + static_cast<int64_t>(b) * a.T * d.num_bins * a.plane_stride + out_off;
  float* fout = d.frame_out ? d.frame_out + static_cast<int64_t>(b) * a.Tf * HW + pix0 : nullptr;
  const double* gauss = (NOISE == V2V_NOISE_EXPLICIT && d.base_gauss)
                            ? d.base_gauss + static_cast<int64_t>(b) * (N - 1) * HW + pix0
                            : nullptr;

  if (fout && d.frame_out_mode == 2) {   // frame 0 is an output frame (output_additional_frame)
    const uint32_t w0 = PixWord<P>::load(frame_ptr(0));
#pragma unroll
    for (int k = 0; k < P; ++k) fout[k] = f255_s[(w0 >> (8 * k)) & 0xffu];
  }

  // accumulators over frames_per_bin
  double accd[EXTERNAL ? P : 1];
  int acci[EXTERNAL ? 1 : P];
#pragma unroll
  for (int k = 0; k < P; ++k) {
    if (EXTERNAL) accd[k] = 0.0; else acci[k] = 0;
  }
  unsigned int npos = 0, nneg = 0;
  const int fpb = d.frames_per_bin;
  int sub = 0;       // intervals accumulated in the current bin
  int gsub = 0;      // intervals since the last output frame
  int tframe = (d.frame_out_mode == 2) ? 1 : 0;

  // ---- register ring of PF frames in flight ----
  uint32_t ring[PF];
#pragma unroll
  for (int u = 0; u < PF; ++u) ring[u] = (1 + u < N) ? PixWord<P>::load(frame_ptr(1 + u)) : 0u;

  for (int i0 = 1; i0 < N; i0 += PF) {
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int i = i0 + u;
      if (i < N) {
        const uint32_t w = ring[u];
        if (i + PF < N) ring[u] = PixWord<P>::load(frame_ptr(i + PF));

        // noise for this interval
        double bn[P];
        if (NOISE == V2V_NOISE_EXPLICIT) {
          if (gauss) {
            const double* gp = gauss + static_cast<int64_t>(i - 1) * HW;
            if (P == 4) {
              double2 g01 = ld_stream_f64x2(gp), g23 = ld_stream_f64x2(gp + 2);
              bn[0] = __dmul_rn(nstd, g01.x);
              bn[1 % P] = __dmul_rn(nstd, g01.y);
              bn[2 % P] = __dmul_rn(nstd, g23.x);
              bn[3 % P] = __dmul_rn(nstd, g23.y);
            } else {
#pragma unroll
              for (int k = 0; k < P; ++k) bn[k] = __dmul_rn(nstd, gp[k]);
            }
          } else {
#pragma unroll
            for (int k = 0; k < P; ++k) bn[k] = 0.0;
          }
        } else if (NOISE == V2V_NOISE_PHILOX) {
          double b4[4];
          stream_noise4(gs, nsc, dir_s, static_cast<uint32_t>(pix0 >> 2) & 7u, b4);   // intervals are walked in order: one draw each
#pragma unroll
          for (int k = 0; k < P; ++k) {
            const int j = P == 4 ? k : static_cast<int>(pix0 & 3);
            bn[k] = j == 0 ? b4[0] : j == 1 ? b4[1] : j == 2 ? b4[2] : b4[3];
          }
        }

        float outv[P];
#pragma unroll
        for (int k = 0; k < P; ++k) {
          const double pk = pos[PERPIXEL ? k : 0], nk = neg[PERPIXEL ? k : 0];
          const uint32_t v = (w >> (8 * k)) & 0xffu;
          const double L = lut_s[v * LUTC + lutc];
          double x = __dadd_rn(pot[k], __dsub_rn(L, lprev[k]));     // :42-43
          lprev[k] = L;
          if (NOISE != V2V_NOISE_NONE && !EXTERNAL) {               // :46-49
            x = __dadd_rn(x, bn[k]);
            x = __dadd_rn(x, hot[k]);
          }
          int cnt = 0;
          if (x >= pk) {                                            // :51-52,57
            const double q = (x < __dadd_rn(pk, pk)) ? 1.0 : floor_div_exact(x, pk, rpos[PERPIXEL ? k : 0]);
            x = __dsub_rn(x, __dmul_rn(q, pk));
            cnt = static_cast<int>(q);
            npos += static_cast<unsigned int>(cnt);
          } else if (x <= -nk) {                                    // :54-55,58
            const double ax = -x;
            const double q = (ax < __dadd_rn(nk, nk)) ? 1.0 : floor_div_exact(ax, nk, rneg[PERPIXEL ? k : 0]);
            x = __dadd_rn(x, __dmul_rn(q, nk));
            cnt = -static_cast<int>(q);
            nneg += static_cast<unsigned int>(-cnt);
          }
          pot[k] = x;
          if (EXTERNAL) {                                           // :60-65
            double vv = static_cast<double>(cnt);
            vv = __dadd_rn(vv, bn[k]);
            vv = __dadd_rn(vv, hot[k]);
            accd[EXTERNAL ? k : 0] = __dadd_rn(accd[EXTERNAL ? k : 0], vv);
            outv[k] = __double2float_rn(accd[EXTERNAL ? k : 0]);
          } else {
            acci[EXTERNAL ? 0 : k] += cnt;
            outv[k] = static_cast<float>(acci[EXTERNAL ? 0 : k]);
          }
        }

        if (++sub == fpb) {                                         // data/v2v_datasets.py:399-400
          sub = 0;
          if (P == 4) {
            st_stream_f32x4(vox, outv[0], outv[1 % P], outv[2 % P], outv[3 % P]);
          } else {
#pragma unroll
            for (int k = 0; k < P; ++k) st_stream_f32(vox + k, outv[k]);
          }
          vox += a.plane_stride;
#pragma unroll
          for (int k = 0; k < P; ++k) {
            if (EXTERNAL) accd[EXTERNAL ? k : 0] = 0.0; else acci[EXTERNAL ? 0 : k] = 0;
          }
        }

        if (fout && ++gsub == a.G) {                                // data/v2v_datasets.py:329-338,352
          gsub = 0;
          float* fo = fout + static_cast<int64_t>(tframe) * HW;
          ++tframe;
          if (P == 4) {
            st_stream_f32x4(fo, f255_s[w & 0xffu], f255_s[(w >> 8) & 0xffu], f255_s[(w >> 16) & 0xffu], f255_s[(w >> 24) & 0xffu]);
          } else {
            st_stream_f32(fo, f255_s[w & 0xffu]);
          }
        }
      }
    }
  }

  if (d.potential_out) {
#pragma unroll
    for (int k = 0; k < P; ++k) d.potential_out[clip_pix + k] = pot[k];
  }
  if (d.stats) {
    // the whole CTA belongs to clip b; lanes past the plane end exited above
    unsigned long long* st = reinterpret_cast<unsigned long long*>(d.stats + 2 * b);
    if (__activemask() == 0xffffffffu) {
      const long long sp = warp_sum(static_cast<long long>(npos)), sn = warp_sum(static_cast<long long>(nneg));
      if ((threadIdx.x & 31) == 0) {
        if (sp) atomicAdd(st, static_cast<unsigned long long>(sp));
        if (sn) atomicAdd(st + 1, static_cast<unsigned long long>(sn));
      }
    } else {   // ragged last warp of the plane
      if (npos) atomicAdd(st, static_cast<unsigned long long>(npos));
      if (nneg) atomicAdd(st + 1, static_cast<unsigned long long>(nneg));
    }
  }
}

template <int P, int NOISE, bool EXTERNAL, bool PERPIXEL, int PF, int LUTC>
int launch(const EsimArgs& a, cudaStream_t s) {
  const int64_t groups = (a.HW + P - 1) / P;
  dim3 grid(static_cast<unsigned int>((groups + kThreads - 1) / kThreads), static_cast<unsigned int>(a.d.B));
  esim_kernel<P, NOISE, EXTERNAL, PERPIXEL, PF, LUTC><<<grid, kThreads, 0, s>>>(a);
  count_launch();
  V2V_CUDA(cudaGetLastError());
  return V2V_OK;
}

template <int P, int PF, int LUTC>
int dispatch_mode(const EsimArgs& a, cudaStream_t s) {
  const bool ext = a.d.put_noise_external != 0 && a.d.noise_mode != V2V_NOISE_NONE;
  const bool pp = a.d.threshold_mode == V2V_THRES_PER_PIXEL;
#define V2V_GO(NM, EX, PPX) return launch<P, NM, EX, PPX, PF, LUTC>(a, s)
  switch (a.d.noise_mode) {
    case V2V_NOISE_NONE:
      if (pp) V2V_GO(V2V_NOISE_NONE, false, true); else V2V_GO(V2V_NOISE_NONE, false, false);
    case V2V_NOISE_EXPLICIT:
      if (ext) { if (pp) V2V_GO(V2V_NOISE_EXPLICIT, true, true); else V2V_GO(V2V_NOISE_EXPLICIT, true, false); }
      else     { if (pp) V2V_GO(V2V_NOISE_EXPLICIT, false, true); else V2V_GO(V2V_NOISE_EXPLICIT, false, false); }
    case V2V_NOISE_PHILOX:
      if (ext) { if (pp) V2V_GO(V2V_NOISE_PHILOX, true, true); else V2V_GO(V2V_NOISE_PHILOX, true, false); }
      else     { if (pp) V2V_GO(V2V_NOISE_PHILOX, false, true); else V2V_GO(V2V_NOISE_PHILOX, false, false); }
  }
#undef V2V_GO
  set_error("bad noise_mode %d", a.d.noise_mode);
  return V2V_ERR_INVALID_ARG;
}

// Materialise the Philox noise fields exactly as the simulation kernels draw them (test / audit hook):
// feeding them back through V2V_NOISE_EXPLICIT (with base_noise_std = 1) must reproduce a PHILOX run bit for bit.
__global__ void esim_philox_fields_kernel(const EsimArgs a, double* u0, double* hot, double* bn) {
  __shared__ uint2 dir_s[kDirEntries];
  fill_dir_table(dir_s);
  __syncthreads();
  const v2v_esim_desc& d = a.d;
  const int b = blockIdx.y;
  const int64_t pix = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (pix >= a.HW) return;
  const NoiseKey nkey = make_noise_key(d.clip_index_base + static_cast<uint64_t>(b));
  const int64_t o = static_cast<int64_t>(b) * a.HW + pix;
  double u, h;
  philox_init_pixel(static_cast<uint64_t>(pix), nkey, a.rk, d.hot_pixel_fraction[b], static_cast<float>(d.hot_pixel_std[b]), &u, &h);
  if (u0) u0[o] = u;
  if (hot) hot[o] = h;
  if (bn) {
    const NoiseScale nsc = make_noise_scale(static_cast<float>(d.base_noise_std[b]));
    NoiseStream gs = noise_stream_init(static_cast<uint64_t>(pix) >> 2, nkey, a.rk);
    const int j = static_cast<int>(pix & 3);
    for (int i = 0; i < d.N - 1; ++i) {
      double b4[4];
      stream_noise4(gs, nsc, dir_s, static_cast<uint32_t>(pix >> 2) & 7u, b4);
      bn[(static_cast<int64_t>(b) * (d.N - 1) + i) * a.HW + pix] = j == 0 ? b4[0] : j == 1 ? b4[1] : j == 2 ? b4[2] : b4[3];
    }
  }
}

}  // namespace
}  // namespace v2v

namespace v2v {
namespace {
__global__ void rng_words_kernel(uint4 ctr, uint2 key, uint32_t* philox_out, EsimArgs a, uint64_t clip_index, uint64_t group, int n,
                                 uint32_t* words_out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (philox_out) {
    const uint4 r = Philox::run(ctr, key);
    philox_out[0] = r.x, philox_out[1] = r.y, philox_out[2] = r.z, philox_out[3] = r.w;
  }
  if (words_out) {
    NoiseStream gs = noise_stream_init(group, make_noise_key(clip_index), a.rk);
    for (int i = 0; i < n; ++i) words_out[i] = noise_stream_next(gs);
  }
}
}  // namespace
}  // namespace v2v

extern "C" int v2v_noise_direction_table(uint32_t* table_host) {
  using namespace v2v;
  V2V_REQUIRE(table_host != nullptr, V2V_ERR_INVALID_ARG, "table_host is NULL");
  static const uint2 h_dir_table[kDirEntries] = {
#include "dir_table.inc"
  };
  for (int k = 0; k < kDirEntries; ++k) {
    table_host[2 * k] = h_dir_table[k].x;
    table_host[2 * k + 1] = h_dir_table[k].y;
  }
  return V2V_OK;
}

extern "C" int v2v_rng_words(const uint32_t counter[4], const uint32_t key[2], uint32_t* philox_out, uint64_t seed, uint64_t clip_index,
                             uint64_t pixel_group, int32_t n_words, uint32_t* words_out, void* stream) {
  using namespace v2v;
  V2V_REQUIRE(n_words >= 0 && (philox_out == nullptr || (counter && key)), V2V_ERR_INVALID_ARG, "bad arguments");
  EsimArgs a{};
  Philox::round_keys(seed, a.rk);
  const uint4 c = counter ? make_uint4(counter[0], counter[1], counter[2], counter[3]) : make_uint4(0, 0, 0, 0);
  const uint2 k = key ? make_uint2(key[0], key[1]) : make_uint2(0, 0);
  rng_words_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(c, k, philox_out, a, clip_index, pixel_group, n_words, words_out);
  count_launch();
  V2V_CUDA(cudaGetLastError());
  return V2V_OK;
}

extern "C" int v2v_esim_frames_to_voxel(const v2v_esim_desc* desc, void* stream) {
  using namespace v2v;
  V2V_REQUIRE(desc != nullptr, V2V_ERR_INVALID_ARG, "desc is NULL");
  const v2v_esim_desc& d = *desc;
  V2V_REQUIRE(d.B >= 0 && d.N >= 1 && d.H >= 0 && d.W >= 0, V2V_ERR_INVALID_ARG, "bad shape B=%d N=%d H=%d W=%d", d.B, d.N, d.H, d.W);
  V2V_REQUIRE(d.num_bins >= 1 && d.frames_per_bin >= 1, V2V_ERR_INVALID_ARG, "num_bins and frames_per_bin must be >= 1");
  const int G = d.num_bins * d.frames_per_bin;
  // data/v2v_datasets.py:365
  V2V_REQUIRE((d.N - 1) % G == 0, V2V_ERR_SHAPE, "(N-1)=%d is not a multiple of num_bins*frames_per_bin=%d", d.N - 1, G);
  V2V_REQUIRE(d.noise_mode >= 0 && d.noise_mode <= 2, V2V_ERR_INVALID_ARG, "bad noise_mode %d", d.noise_mode);
  V2V_REQUIRE(d.threshold_mode == 0 || d.threshold_mode == 1, V2V_ERR_INVALID_ARG, "bad threshold_mode %d", d.threshold_mode);
  V2V_REQUIRE(d.frame_out_mode >= 0 && d.frame_out_mode <= 2, V2V_ERR_INVALID_ARG, "bad frame_out_mode %d", d.frame_out_mode);
  V2V_REQUIRE(d.B <= 65535, V2V_ERR_UNSUPPORTED, "B=%d > 65535 clips per call", d.B);
  const int64_t HW = static_cast<int64_t>(d.H) * d.W;
  if (d.B == 0 || HW == 0 || d.N == 1) return V2V_OK;   // empty input: nothing to write
  V2V_REQUIRE(d.frames && d.lut && d.pos_thres && d.neg_thres && d.voxel, V2V_ERR_INVALID_ARG,
              "frames, lut, pos_thres, neg_thres and voxel must be non-NULL");
  V2V_REQUIRE(d.noise_mode != V2V_NOISE_PHILOX || (d.base_noise_std && d.hot_pixel_fraction && d.hot_pixel_std),
              V2V_ERR_INVALID_ARG, "PHILOX noise needs base_noise_std, hot_pixel_fraction and hot_pixel_std");
  V2V_REQUIRE(d.noise_mode != V2V_NOISE_EXPLICIT || d.base_noise_std || !d.base_gauss, V2V_ERR_INVALID_ARG,
              "EXPLICIT noise with base_gauss needs base_noise_std");
  V2V_REQUIRE(d.frame_out_mode == 0 || d.frame_out, V2V_ERR_INVALID_ARG, "frame_out_mode set but frame_out is NULL");
  V2V_REQUIRE(d.raw_frames_per_clip >= 0 && (d.frame_index || d.raw_frames_per_clip == 0 || d.raw_frames_per_clip == d.N),
              V2V_ERR_INVALID_ARG, "raw_frames_per_clip=%d needs frame_index (frames is [B,N,H,W] without it)", d.raw_frames_per_clip);
  V2V_REQUIRE(!d.frame_index || aligned(d.frame_index, 4), V2V_ERR_ALIGNMENT, "frame_index must be 4-byte aligned");
  V2V_REQUIRE(aligned(d.lut, 8) && aligned(d.pos_thres, 8) && aligned(d.neg_thres, 8) && aligned(d.voxel, 4),
              V2V_ERR_ALIGNMENT, "misaligned float64/float32 pointer");

  EsimArgs a;
  a.d = d;
  a.HW = HW;
  a.G = G;
  a.T = (d.N - 1) / G;
  a.row_stride = d.voxel_row_stride ? d.voxel_row_stride : d.W;
  a.plane_stride = d.voxel_plane_stride ? d.voxel_plane_stride : HW;
  V2V_REQUIRE(a.row_stride >= d.W && a.plane_stride >= a.row_stride * (d.H - 1) + d.W, V2V_ERR_SHAPE,
              "voxel strides too small");
  a.padded = a.row_stride != d.W;
  Philox::round_keys(d.seed, a.rk);
  a.Tf = d.frame_out_mode == 2 ? a.T + 1 : a.T;
  a.Mraw = (d.frame_index && d.raw_frames_per_clip > 0) ? d.raw_frames_per_clip : d.N;
  if (!d.frame_out) a.d.frame_out_mode = 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);

  // 4 pixels per thread needs 4-byte aligned frame words and 16-byte aligned voxel quads
  const bool vec4 = (HW % 4 == 0) && aligned(d.frames, 4) && aligned(d.voxel, 16) && (a.plane_stride % 4 == 0) &&
                    (!a.padded || (d.W % 4 == 0 && a.row_stride % 4 == 0)) && (!d.frame_out || aligned(d.frame_out, 16)) &&
                    (!d.base_gauss || aligned(d.base_gauss, 16));
  // small launches: one pixel per thread spreads the serial recurrence over more SMs (noise-free only: with
  // Philox the 4-pixel kernel shares one generator call between its pixels and wins at every size)
  const bool big = static_cast<int64_t>(d.B) * HW >= 148LL * 2048;
  const bool generic_only = (d.kernel_flags & V2V_ESIM_FLAG_GENERIC) != 0;
  bool small_fast = d.noise_mode == V2V_NOISE_PHILOX;
  if (d.kernel_flags & V2V_ESIM_FLAG_SMALL_FAST) small_fast = true;
  if (d.kernel_flags & V2V_ESIM_FLAG_SMALL_P1) small_fast = false;
  if (vec4 && !generic_only && esim_fast_eligible(a) && (big || small_fast)) return launch_esim_fast(a, s);
  if (!vec4 || !big) return dispatch_mode<1, 4, 1>(a, s);
  return dispatch_mode<4, 4, 1>(a, s);
}

extern "C" int v2v_esim_philox_fields(const v2v_esim_desc* desc, double* u0, double* hot_noise, double* base_noise, void* stream) {
  using namespace v2v;
  V2V_REQUIRE(desc != nullptr, V2V_ERR_INVALID_ARG, "desc is NULL");
  const v2v_esim_desc& d = *desc;
  V2V_REQUIRE(d.B >= 0 && d.N >= 1 && d.H >= 0 && d.W >= 0 && d.B <= 65535, V2V_ERR_INVALID_ARG, "bad shape");
  V2V_REQUIRE(d.base_noise_std && d.hot_pixel_fraction && d.hot_pixel_std, V2V_ERR_INVALID_ARG,
              "base_noise_std, hot_pixel_fraction and hot_pixel_std must be non-NULL");
  EsimArgs a;
  a.d = d;
  a.HW = static_cast<int64_t>(d.H) * d.W;
  Philox::round_keys(d.seed, a.rk);
  if (d.B == 0 || a.HW == 0) return V2V_OK;
  dim3 grid(static_cast<unsigned int>((a.HW + 255) / 256), static_cast<unsigned int>(d.B));
  esim_philox_fields_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, u0, hot_noise, base_noise);
  count_launch();
  V2V_CUDA(cudaGetLastError());
  return V2V_OK;
}

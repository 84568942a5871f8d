// ESIM frames -> voxel: the throughput kernel for the shipped configuration
// (per-clip thresholds, frames_per_bin == 1, noise none or generated in the
// kernel and applied to the potential; reference data/v2v_core_esim.py:26-69
// with put_noise_external=False, data/v2v_datasets.py:399-400 with fpb=1).
//
// Same arithmetic as the generic kernel in esim.cu, restructured so that the
// warp issues as few instructions per pixel-interval as possible (the generic
// kernel is instruction-issue bound at ~40 % of HBM bandwidth):
//   * 4 pixels per lane, whole clip in registers, frames through a register
//     ring of 2*PF words, voxels as one 128-bit streaming store per frame;
//   * the LUT is replicated 8x in shared memory ([value][lane&7]): a half-warp's
//     64-bit lookups collide only between lanes l and l+8 on different values;
//   * a crossing by exactly one threshold (the common case) is branch-free
//     (select + add, or two FMAs with a 0/1 factor built from the predicate);
//     only multi-threshold crossings take the divergent exact floor-division
//     path, triggered by four FP64 compares chained through one predicate;
//   * noise from one xoshiro128++ stream per 4-pixel group (seeded by Philox),
//     8 normals per draw for a pair of intervals (esim_common.cuh);
//   * statistics in registers, one pair of global reductions per warp at the
//     end: no shared-memory stage and no CTA barrier after the loop;
//   * optional pause gather / degrade of the dataset fused into the frame loads
//     and the LUTs (frame_index, value_map).
#include "esim_common.cuh"

#include <cstdlib>

namespace v2v {
namespace {

constexpr int kPF = 4;        // frames per loop trip; 2*kPF frames in flight
constexpr int kLutCopies = 8;
constexpr int kFastThreads = 128;
constexpr int kFlushTrips = 8;  // statistics: float partial sums cover 4*kPF*kFlushTrips pixel-intervals (exact for counts < 2^17 each)

// Exact multi-threshold crossing on the ORIGINAL potential x (data/v2v_core_esim.py:51-58).
// Also correct for |x| below the threshold (count 0), so the caller's trigger may be conservative.
__device__ __forceinline__ double multi_cross(double x, double pos, double neg, double rpos, double rneg, int* cnt) {
  const bool down = x < 0.0;
  const double a = fabs(x), thr = down ? neg : pos, rthr = down ? rneg : rpos;
  if (a < thr) {
    *cnt = 0;
    return x;
  }
  const double q = floor_div_exact(a, thr, rthr);
  const double an = __dsub_rn(a, __dmul_rn(q, thr));
  *cnt = down ? -static_cast<int>(q) : static_cast<int>(q);
  return down ? -an : an;
}

// The common case, branch-free: at most one threshold is crossed, so q*thr == thr exactly and
// x - q*pos is one rounded subtraction (same value as the reference's separately rounded product
// and difference because the product is exact).  Written as two predicated FP64 adds: the FP64 pipe
// has slack, the ALU pipe (selects) does not.
__device__ __forceinline__ void single_cross(double& x, float& ov, double pos, double mneg) {
  const bool up = x >= pos, dn = x <= mneg;                                 // :52,55
  // qu in {0.0, 1.0}, qd in {0.0, -1.0}: q*thr is exact, so fma(-thr, q, x) rounds once exactly like x -/+ q*thr (:57-58),
  // and a zero q adds -0.0, which leaves every x untouched.  Only the high words are selected (the low words are 0).
  const int hu = up ? 0x3ff00000 : 0, hd = dn ? static_cast<int>(0xbff00000u) : 0;
  x = __fma_rn(-pos, __hiloint2double(hu, 0), x);
  x = __fma_rn(mneg, __hiloint2double(hd, 0), x);
  ov = __int_as_float((hu | hd) & static_cast<int>(0xbf800000u));           // +1.0f, -1.0f or 0.0f from the same words
}

// Select form of the same update (one FP64 add of {-pos, +neg, 0}): fewer registers; used by the noise-free
// variants, which run at 64 registers / 8 CTAs per SM (the FMA form costs them registers: 1.20 vs 1.15 ms per 32 clips).
__device__ __forceinline__ void single_cross_sel(double& x, float& ov, double pos, double mneg, double neg) {
  const bool up = x >= pos, dn = x <= mneg;
  double sel = up ? -pos : 0.0;
  sel = dn ? neg : sel;
  x = __dadd_rn(x, sel);
  ov = up ? 1.0f : 0.0f;
  ov = dn ? -1.0f : ov;
}

template <int NOISE, bool FRAMES, bool STATS, int CTAS>
__global__ void __launch_bounds__(kFastThreads, CTAS) esim_fast_kernel(const EsimArgs a) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];     // [LUT copies 32 KB][trig table 32 KB, Philox only]
  double* lut_s = reinterpret_cast<double*>(dyn_smem);
  float2* trig_s = reinterpret_cast<float2*>(dyn_smem + 256 * kLutCopies * sizeof(double));
  __shared__ float f255_s[FRAMES ? 256 : 1];                       // (mapped value)/255 of the ground-truth frame output
  __shared__ double cta_rcp[2];                                    // 1/pos, 1/neg of this CTA's clip: only the rare path reads them
  __shared__ float4 hot_s[NOISE == V2V_NOISE_PHILOX ? kFastThreads : 1];   // per-lane hot-pixel noise: read by the ~6 % of warps that own one
  int* fnum_s = reinterpret_cast<int*>(dyn_smem + 256 * kLutCopies * sizeof(double) + (NOISE == V2V_NOISE_PHILOX ? kTrigEntries * sizeof(float2) : 0));
  if (NOISE == V2V_NOISE_PHILOX) fill_trig_table(trig_s);
  const v2v_esim_desc& d = a.d;
  {
    const int32_t* fidx = d.frame_index ? d.frame_index + static_cast<int64_t>(blockIdx.y) * d.N : nullptr;
    for (int n = threadIdx.x; n < d.N; n += kFastThreads) fnum_s[n] = fidx ? min(max(fidx[n], 0), a.Mraw - 1) : n;
  }
  if (threadIdx.x < 2) cta_rcp[threadIdx.x] = __drcp_rn(threadIdx.x ? d.neg_thres[blockIdx.y] : d.pos_thres[blockIdx.y]);
  {
    const uint8_t* vmap = d.value_map ? d.value_map + static_cast<int64_t>(blockIdx.y) * 256 : nullptr;   // degrade folded into the LUTs
    for (int e = threadIdx.x; e < 256; e += kFastThreads) {
      const int ev = vmap ? vmap[e] : e;
      const double v = d.lut[ev];
      if (FRAMES) f255_s[e] = __fdiv_rn(static_cast<float>(ev), 255.0f);
#pragma unroll
      for (int c = 0; c < kLutCopies; ++c) lut_s[e * kLutCopies + c] = v;
    }
  }
  __syncthreads();

  const int b = blockIdx.y;
  const int64_t pix0 = (static_cast<int64_t>(blockIdx.x) * kFastThreads + threadIdx.x) * 4;
  const int64_t HW = a.HW;
  if (pix0 < HW) {
  const int N = d.N;
  const int64_t clip_pix = static_cast<int64_t>(b) * HW + pix0;
  const NoiseKey nkey = make_noise_key(d.clip_index_base + static_cast<uint64_t>(b));
  GroupStream gs{0u, 0u, 0u, 0u};
  if (NOISE == V2V_NOISE_PHILOX) gs = group_stream_init(static_cast<uint64_t>(pix0) >> 2, nkey, a.rk);

  const double pos = d.pos_thres[b], neg = d.neg_thres[b];
  const double mneg = -neg;
  const double thr2 = __dadd_rn(fmin(pos, neg), fmin(pos, neg));   // below 2*min(pos,neg) at most one threshold is crossed
  constexpr bool kFmaCross = NOISE == V2V_NOISE_PHILOX;            // which form of the single crossing (see single_cross*)
  const float nc2 = NOISE == V2V_NOISE_PHILOX ? noise_c2(static_cast<float>(d.base_noise_std[b])) : 0.f;
  // byte offset of this lane's LUT copy
  const uint32_t lut_base = static_cast<uint32_t>(__cvta_generic_to_shared(dyn_smem)) + (threadIdx.x & (kLutCopies - 1)) * 8u;
  auto lut_at = [&](uint32_t v) -> double {
    double r;
    asm("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"(lut_base + v * (kLutCopies * 8u)));
    return r;
  };
  auto byte_of = [](uint32_t w, int k) -> uint32_t {      // one PRMT instead of shift+mask
    uint32_t v;
    asm("prmt.b32 %0, %1, 0, %2;" : "=r"(v) : "r"(w), "r"(0x4440u | static_cast<uint32_t>(k)));
    return v;
  };

  // frame n of the clip = raw frame frame_index[b][n] (pause gather, data/v2v_datasets.py:285-311), or n itself: the
  // per-clip table of raw frame numbers sits in shared memory (one broadcast LDS per load, no extra live registers)
  const uint8_t* fr = d.frames + (static_cast<int64_t>(b) * a.Mraw) * HW + pix0;
  auto frame_ptr = [&](int n) -> const uint8_t* { return fr + static_cast<int64_t>(fnum_s[n]) * HW; };
  double pot[4], lprev[4];
  float hotf[4];            // Philox hot-pixel noise is double(float) by construction: keep the float (parked in smem)
  const uint32_t w0 = ld_stream_u32(frame_ptr(0));
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    lprev[k] = lut_at(byte_of(w0, k));
    hotf[k] = 0.f;
    double u = -1.0;
    if (NOISE == V2V_NOISE_PHILOX) {
      double hk;
      philox_init_pixel(static_cast<uint64_t>(pix0 + k), nkey, a.rk, d.hot_pixel_fraction[b], static_cast<float>(d.hot_pixel_std[b]), &u, &hk);
      hotf[k] = static_cast<float>(hk);     // exact: hk was produced from a float
    }
    if (d.u0) u = d.u0[clip_pix + k];
    if (d.potential_in) pot[k] = d.potential_in[clip_pix + k];
    else if (u >= 0.0) pot[k] = __dsub_rn(__dmul_rn(u, __dadd_rn(pos, neg)), neg);
    else pot[k] = 0.0;
  }

  int64_t out_off = pix0;
  if (a.padded) {
    const int64_t row = pix0 / d.W;
    out_off = row * a.row_stride + (pix0 - row * d.W);
  }
  float* vox = d.voxel + static_cast<int64_t>(b) * a.T * d.num_bins * a.plane_stride + out_off;
  float* fout = FRAMES ? d.frame_out + static_cast<int64_t>(b) * a.Tf * HW + pix0 : nullptr;
  int gsub = 0;
  if (FRAMES && d.frame_out_mode == 2) {
    st_stream_f32x4(fout, f255_s[w0 & 0xffu], f255_s[(w0 >> 8) & 0xffu], f255_s[(w0 >> 16) & 0xffu], f255_s[w0 >> 24]);
    fout += HW;
  }
  float net = 0.f, tot = 0.f;             // since the last flush: net = #pos - #neg, tot = #pos + #neg
  unsigned int wpos = 0u, wneg = 0u;      // this warp's event totals (the same value in every lane)
  auto flush_stats = [&]() {              // float sums -> integers, added across the warp with one REDUX each; registers only
    const unsigned int m = __activemask();       // (all active lanes of a warp run the same trip count)
    wpos += __reduce_add_sync(m, static_cast<unsigned int>((tot + net) * 0.5f));
    wneg += __reduce_add_sync(m, static_cast<unsigned int>((tot - net) * 0.5f));
    net = tot = 0.f;
  };

  bool lane_hot = false;
#pragma unroll
  for (int k = 0; k < 4; ++k) lane_hot = lane_hot || hotf[k] != 0.f;
  // warp-uniform: a real branch that 94 % of the warps never take (hot_pixel_fraction <= 1e-3)
  const bool any_hot = __any_sync(__activemask(), lane_hot);
  if (NOISE == V2V_NOISE_PHILOX) hot_s[threadIdx.x] = make_float4(hotf[0], hotf[1], hotf[2], hotf[3]);   // own slot: no barrier needed

  auto step = [&](const uint32_t w, const float (&bnf)[4]) {
    float o[4];
    double x0[4];
    bool rare = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double L = lut_at(byte_of(w, k));
      double x = __dadd_rn(pot[k], __dsub_rn(L, lprev[k]));          // :42-43
      lprev[k] = L;
      if (NOISE == V2V_NOISE_PHILOX) x = __dadd_rn(x, static_cast<double>(bnf[k]));   // :46-48
      x0[k] = x;
    }
    if (NOISE == V2V_NOISE_PHILOX && any_hot) {                       // :49; x + 0.0 == x: skipped by the warps without a hot pixel
      asm volatile("" ::: "memory");                                  // keep this a (warp-uniform) branch
      const float4 h = hot_s[threadIdx.x];
      x0[0] = __dadd_rn(x0[0], static_cast<double>(h.x));
      x0[1] = __dadd_rn(x0[1], static_cast<double>(h.y));
      x0[2] = __dadd_rn(x0[2], static_cast<double>(h.z));
      x0[3] = __dadd_rn(x0[3], static_cast<double>(h.w));
    }
    {   // trigger of the exact multi-threshold path: four FP64 compares chained through one predicate (no ALU-pipe work)
      unsigned int r;
      asm("{\n"
          " .reg .pred p;\n"
          " .reg .f64 t;\n"
          " abs.f64 t, %1;\n setp.ge.f64 p, t, %5;\n"
          " abs.f64 t, %2;\n setp.ge.or.f64 p, t, %5, p;\n"
          " abs.f64 t, %3;\n setp.ge.or.f64 p, t, %5, p;\n"
          " abs.f64 t, %4;\n setp.ge.or.f64 p, t, %5, p;\n"
          " selp.u32 %0, 1, 0, p;\n"
          "}"
          : "=r"(r)
          : "d"(x0[0]), "d"(x0[1]), "d"(x0[2]), "d"(x0[3]), "d"(thr2));
      rare = r != 0;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      double x = x0[k];
      if (kFmaCross) single_cross(x, o[k], pos, mneg);
      else single_cross_sel(x, o[k], pos, mneg, neg);                        // :51-58 with q in {0,1}
      pot[k] = x;
    }
    if (rare) {                                                       // a few % of warp-steps on natural video
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (fabs(x0[k]) >= thr2) {
          int cnt;
          pot[k] = multi_cross(x0[k], pos, neg, cta_rcp[0], cta_rcp[1], &cnt);   // conservative trigger: correct for any x
          o[k] = static_cast<float>(cnt);
        }
      }
    }
    if (STATS) {            // from the final counts; float sums are exact while below 2^24 (flushed every kFlushTrips trips)
      net += (o[0] + o[1]) + (o[2] + o[3]);
      tot += (fabsf(o[0]) + fabsf(o[1])) + (fabsf(o[2]) + fabsf(o[3]));
    }
    st_stream_f32x4(vox, o[0], o[1], o[2], o[3]);
    vox += a.plane_stride;
    if (FRAMES) {                                                     // data/v2v_datasets.py:329-338,352
      if (++gsub == a.G) {
        gsub = 0;
        st_stream_f32x4(fout, f255_s[w & 0xffu], f255_s[(w >> 8) & 0xffu], f255_s[(w >> 16) & 0xffu], f255_s[w >> 24]);
        fout += HW;
      }
    }
  };

  // ---- main loop: kPF frames per trip, the next trip's words already in flight ----
  const int M = N - 1;                       // intervals
  const int trips = M / kPF;
  uint32_t cur[kPF], nxt[kPF];
  if (trips > 0) {
#pragma unroll
    for (int u = 0; u < kPF; ++u) cur[u] = ld_stream_u32(frame_ptr(1 + u));
  }
  int i = 1;
  const float zero4[4] = {0.f, 0.f, 0.f, 0.f};
  for (int t = 0; t < trips; ++t) {
    if (t + 1 < trips) {
#pragma unroll
      for (int u = 0; u < kPF; ++u) nxt[u] = ld_stream_u32(frame_ptr(i + kPF + u));
    }
    if (NOISE == V2V_NOISE_PHILOX) {
      // intervals i-1 .. i+2 = 4t .. 4t+3: two draws of the group's stream, 8 normals each
      float e0[4], o0[4], e1[4], o1[4];
      stream_noise8(gs, nc2, trig_s, e0, o0);
      step(cur[0], e0);
      step(cur[1], o0);
      stream_noise8(gs, nc2, trig_s, e1, o1);
      step(cur[2], e1);
      step(cur[3], o1);
    } else {
#pragma unroll
      for (int u = 0; u < kPF; ++u) step(cur[u], zero4);
    }
    i += kPF;
#pragma unroll
    for (int u = 0; u < kPF; ++u) cur[u] = nxt[u];
    if (STATS && (t & (kFlushTrips - 1)) == kFlushTrips - 1) flush_stats();
  }
  {
    float tev[4] = {0.f, 0.f, 0.f, 0.f}, tod[4] = {0.f, 0.f, 0.f, 0.f};
    for (; i < N; ++i) {                                              // ragged tail (< kPF intervals; starts at an even interval)
      float bn1[4] = {0.f, 0.f, 0.f, 0.f};
      if (NOISE == V2V_NOISE_PHILOX) {
        if (((i - 1) & 1) == 0) stream_noise8(gs, nc2, trig_s, tev, tod);
#pragma unroll
        for (int k = 0; k < 4; ++k) bn1[k] = ((i - 1) & 1) ? tod[k] : tev[k];
      }
      step(ld_stream_u32(frame_ptr(i)), bn1);
    }
  }

  if (d.potential_out) {
#pragma unroll
    for (int k = 0; k < 4; ++k) d.potential_out[clip_pix + k] = pot[k];
  }
  if (STATS) {
    // One pair of global reductions per warp, no CTA barrier and no shared-memory stage: a warp that is done leaves, and the
    // kernel needs 94 instead of 114 registers (5 resident CTAs per SM instead of 4: -9 % on the Philox variant).  153 k
    // fire-and-forget REDs per 32-clip launch on 64 addresses are invisible next to 2 ms of work.
    flush_stats();
    if ((threadIdx.x & 31) == (__ffs(__activemask()) - 1)) {
      if (wpos) atomicAdd(reinterpret_cast<unsigned long long*>(d.stats + 2 * blockIdx.y), static_cast<unsigned long long>(wpos));
      if (wneg) atomicAdd(reinterpret_cast<unsigned long long*>(d.stats + 2 * blockIdx.y) + 1, static_cast<unsigned long long>(wneg));
    }
  }
  }  // valid
}

}  // namespace

bool esim_fast_eligible(const EsimArgs& a) {
  const v2v_esim_desc& d = a.d;
  if (d.frames_per_bin != 1 || d.threshold_mode != V2V_THRES_PER_CLIP) return false;
  if (d.noise_mode == V2V_NOISE_EXPLICIT) return false;
  if (d.noise_mode == V2V_NOISE_PHILOX && d.put_noise_external) return false;
  if (d.N > 16384) return false;          // the per-clip frame-number table lives in shared memory
  return true;
}

int launch_esim_fast(const EsimArgs& a, cudaStream_t s) {
  const int64_t groups = a.HW / 4;
  dim3 grid(static_cast<unsigned int>((groups + kFastThreads - 1) / kFastThreads), static_cast<unsigned int>(a.d.B));
  const bool ph = a.d.noise_mode == V2V_NOISE_PHILOX, fr = a.d.frame_out_mode != 0, st = a.d.stats != nullptr;
  // resident CTAs per SM the register allocator must allow (4 -> 128 regs, 6 -> 80, 8 -> 64), chosen per
  // variant from same-box sweeps on B200 (profiles/r01_esim_minb_sweep.txt); V2V_ESIM_CTAS overrides for tuning.
  // (Philox without statistics needs 96 registers: 5 CTAs are resident under the 4-CTA bound.)
  int ctas = ph ? (st ? 6 : 4) : 8;
  if (const char* e = getenv("V2V_ESIM_CTAS")) ctas = atoi(e);
  const size_t smem = 256 * kLutCopies * sizeof(double) + (ph ? kTrigEntries * sizeof(float2) : 0) + static_cast<size_t>(a.d.N) * sizeof(int);
#define V2V_G(NM, FR, ST, CT)                                                                                     \
  do {                                                                                                            \
    V2V_CUDA(cudaFuncSetAttribute(esim_fast_kernel<NM, FR, ST, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))); \
    esim_fast_kernel<NM, FR, ST, CT><<<grid, kFastThreads, smem, s>>>(a);                                        \
  } while (0)
#define V2V_F(NM, FR, ST)                       \
  do {                                          \
    if (ctas <= 4) V2V_G(NM, FR, ST, 4);        \
    else if (ctas <= 7) V2V_G(NM, FR, ST, 6);   \
    else V2V_G(NM, FR, ST, 8);                  \
  } while (0)
  if (ph) {
    if (fr) { if (st) V2V_F(V2V_NOISE_PHILOX, true, true); else V2V_F(V2V_NOISE_PHILOX, true, false); }
    else    { if (st) V2V_F(V2V_NOISE_PHILOX, false, true); else V2V_F(V2V_NOISE_PHILOX, false, false); }
  } else {
    if (fr) { if (st) V2V_F(V2V_NOISE_NONE, true, true); else V2V_F(V2V_NOISE_NONE, true, false); }
    else    { if (st) V2V_F(V2V_NOISE_NONE, false, true); else V2V_F(V2V_NOISE_NONE, false, false); }
  }
#undef V2V_F
#undef V2V_G
  count_launch();
  V2V_CUDA(cudaGetLastError());
  return V2V_OK;
}

}  // namespace v2v

// ESIM frames -> voxel: the throughput kernel for the shipped configuration
// (per-clip thresholds, frames_per_bin == 1, noise none or generated in the
// kernel and applied to the potential; reference data/v2v_core_esim.py:26-69
// with put_noise_external=False, data/v2v_datasets.py:399-400 with fpb=1).
//
// Same arithmetic as the generic kernel in esim.cu.  The kernel is bound by
// instruction issue, not by bytes, so everything here is about instructions
// (and shared-memory wavefronts) per (4 pixels x interval):
//   * 4 pixels per lane, whole clip in registers, frames through a register
//     ring of 2*kPF words, voxels as one 128-bit streaming store per frame;
//   * the log LUT sits in shared memory as [value][lane] (32 copies, 64 KB per
//     CTA): every 64-bit lookup of a warp is conflict-free, and because an entry
//     row is 256 bytes the address is ONE byte-permute of the frame word into
//     the lane's offset (PRMT) added to a uniform base by the load itself;
//   * a crossing by exactly one threshold (the common case) is branch-free
//     (two FMAs with a 0/1 factor built from the predicate, or select + add);
//     only multi-threshold crossings take the divergent exact floor-division
//     path, triggered by four FP64 compares chained through one predicate;
//   * noise: one 64-bit LCG stream per 4-pixel group (seeded by Philox), one
//     32-bit word per pair of normals (two neighbouring pixels), radius from
//     the SFU, direction from a 2048-entry float64 table whose 16-byte entries
//     a quarter warp fetches without bank conflicts; the noise enters the
//     potential as x = fma(radius, direction, x) — the product is exact, so this
//     is the reference's `potential += noise` without an FP32 multiply;
//   * statistics with packed f32x2 adds, one pair of global reductions per warp
//     at the end: no shared-memory stage and no CTA barrier after the loop;
//   * optional pause gather / degrade of the dataset fused into the frame loads
//     and the LUTs (frame_index, value_map).
#include "esim_common.cuh"

#include <atomic>

namespace v2v {
namespace {

#ifndef V2V_KPF
#define V2V_KPF 4
#endif
constexpr int kPF = V2V_KPF;    // frames per loop trip; 2*kPF frames in flight
constexpr int kLutCopies = 32;  // one copy per lane
constexpr int kLutBytes = 256 * kLutCopies * 8;
// Shared-window address of the dynamic shared memory of a non-cluster CTA without static shared memory on sm_100a (the
// first KB of the window is reserved; ptxas itself emits this constant when it forms the address).  With the base known at
// compile time the table lookups below are `LDS [reg + immediate]`: no address add per lookup.  The kernel checks the
// assumption once per CTA and traps if it ever does not hold.
constexpr uint32_t kSmemWindowBase = 0x400u;
constexpr int kFlushTrips = 32 / kPF;  // statistics: the packed counters are unpacked every 32 intervals

// Exact multi-threshold crossing (data/v2v_core_esim.py:51-58) of a magnitude a >= thr: q = floor(a/thr) as a
// mathematical quantity (np.floor_divide), remainder a - RN(q*thr).  rthr = RN(1/thr).
__device__ __forceinline__ double multi_cross(double a, double thr, double rthr, double* q_out) {
  double q = floor(__dmul_rn(a, rthr));          // off by at most one
  const double r = __fma_rn(-q, thr, a);         // sign and size of the single-rounded residual are exact
  q = r < 0.0 ? q - 1.0 : (r >= thr ? q + 1.0 : q);
  *q_out = q;
  return __dsub_rn(a, __dmul_rn(q, thr));
}

// The common case, branch-free: at most one threshold is crossed, so q*thr == thr exactly and x - q*pos is one rounded
// subtraction.  qu in {0.0, 1.0}, qd in {0.0, -1.0}: fma(-thr, q, x) rounds once exactly like x -/+ q*thr (:57-58), and
// a zero q adds -0.0, which leaves every x untouched.  Only the high words are selected (the low words are 0).
__device__ __forceinline__ void single_cross(double& x, float& ov, uint32_t& hu_out, double pos, double mneg) {
  const bool up = x >= pos, dn = x <= mneg;                                 // :52,55
  const int hu = up ? 0x3ff00000 : 0, hd = dn ? static_cast<int>(0xbff00000u) : 0;
  hu_out = static_cast<uint32_t>(hu);
  x = __fma_rn(-pos, __hiloint2double(hu, 0), x);
  x = __fma_rn(mneg, __hiloint2double(hd, 0), x);
  ov = __int_as_float((hu | hd) & static_cast<int>(0xbf800000u));           // +1.0f, -1.0f or 0.0f from the same words
}

// Select form of the same update (one FP64 add of {-pos, +neg, 0}): fewer registers; used by the noise-free variants.
__device__ __forceinline__ void single_cross_sel(double& x, float& ov, uint32_t& hu_out, double pos, double mneg, double neg) {
  const bool up = x >= pos, dn = x <= mneg;
  hu_out = up ? 0x3ff00000u : 0u;
  double sel = up ? -pos : 0.0;
  sel = dn ? neg : sel;
  x = __dadd_rn(x, sel);
  ov = up ? 1.0f : 0.0f;
  ov = dn ? -1.0f : ov;
}

// ---- bulk-copy frame ring (STAGED variants): cp.async.bulk + mbarrier, SASS UBLKCP / SYNCS ----------------------------
constexpr int kRingDepth = 8;   // frame tiles in flight per CTA
constexpr int kRingLag = 4;     // a slot is refilled kRingLag intervals after it was read (no warp waits for the slowest one)

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, int bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

template <int NOISE, bool FRAMES, bool GATHER, int THREADS>
struct FastSmem {
  static constexpr bool kPh = NOISE == V2V_NOISE_PHILOX;
  static constexpr int lut = 0;
  static constexpr int dir = lut + kLutBytes;
  static constexpr int hot = dir + (kPh ? kDirEntries * 16 : 0);
  static constexpr int f255 = hot + (kPh ? THREADS * 16 : 0);
  static constexpr int rcp = f255 + (FRAMES ? 256 * 4 : 0);
  static constexpr int xst = rcp + 16;                        // the warps' running event totals (variants that park them): uint2 per warp
  static constexpr int fnum = xst + (THREADS / 32) * 8;
  static __host__ __device__ size_t bytes(int N) { return fnum + (GATHER ? static_cast<size_t>(N) * 4 : 0); }
  static __host__ __device__ size_t ring(int N) { return (bytes(N) + 127) / 128 * 128; }                       // offset of the staged frame ring
  static size_t bytes_staged(int N) { return ring(N) + static_cast<size_t>(kRingDepth) * THREADS * 4 + 16 * kRingDepth; }
};

template <int NOISE, bool FRAMES, bool STATS, bool GATHER, int THREADS, int CTAS, bool STAGED = false>
__global__ void __launch_bounds__(THREADS, CTAS) esim_fast_kernel(const EsimArgs a) {
  using L = FastSmem<NOISE, FRAMES, GATHER, THREADS>;
  constexpr bool kPh = NOISE == V2V_NOISE_PHILOX;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  float4* hot_s = reinterpret_cast<float4*>(dyn_smem + L::hot);        // per-lane hot-pixel noise: read by the ~6 % of warps that own one
  float* f255_s = reinterpret_cast<float*>(dyn_smem + L::f255);        // (mapped value)/255 of the ground-truth frame output
  double* cta_rcp = reinterpret_cast<double*>(dyn_smem + L::rcp);      // 1/pos, 1/neg of this CTA's clip: only the rare path reads them
  uint32_t* fnum_s = reinterpret_cast<uint32_t*>(dyn_smem + L::fnum);  // raw frame number of frame n (pause gather)
  const v2v_esim_desc& d = a.d;
  const uint32_t smem_base = static_cast<uint32_t>(__cvta_generic_to_shared(dyn_smem));
  if (smem_base != kSmemWindowBase) __trap();
  if (kPh) {                       // {double(cos), double(sin)} per direction: the table's high words over zero low words
    uint4* dir_s = reinterpret_cast<uint4*>(dyn_smem + L::dir);
    for (int k = threadIdx.x; k < kDirEntries; k += THREADS) dir_s[k] = make_uint4(0u, g_dir_table[k].x, 0u, g_dir_table[k].y);
  }
  if (GATHER) {
    const int32_t* fidx = d.frame_index + static_cast<int64_t>(blockIdx.y) * d.N;
    for (int n = threadIdx.x; n < d.N; n += THREADS) fnum_s[n] = static_cast<uint32_t>(min(max(fidx[n], 0), a.Mraw - 1));
  }
  if (threadIdx.x < 2) cta_rcp[threadIdx.x] = __drcp_rn(threadIdx.x ? d.neg_thres[blockIdx.y] : d.pos_thres[blockIdx.y]);
  const uint32_t ring_a = smem_base + static_cast<uint32_t>(L::ring(d.N));                 // STAGED: frame ring, then full / empty barriers
  const uint32_t full_a = ring_a + kRingDepth * THREADS * 4, empty_a = full_a + 8 * kRingDepth;
  if (STAGED && threadIdx.x == 0) {
    const int64_t left = (a.HW - static_cast<int64_t>(blockIdx.x) * THREADS * 4 + 3) / 4;          // lanes of this tile that own pixels
    const int warps = static_cast<int>((min(left, static_cast<int64_t>(THREADS)) + 31) / 32);
    for (int q = 0; q < kRingDepth; ++q) {
      mbar_init(full_a + 8 * q, 1);
      mbar_init(empty_a + 8 * q, warps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    const uint8_t* vmap = d.value_map ? d.value_map + static_cast<int64_t>(blockIdx.y) * 256 : nullptr;   // degrade folded into the LUTs
    for (int j = threadIdx.x; j < 256 * 4; j += THREADS) {     // a quarter row (8 copies) per task, two copies per 128-bit store
      const int v = j >> 2;
      const int ev = vmap ? vmap[v] : v;
      const double val = d.lut[ev];
      const uint32_t at = smem_base + L::lut + j * 64;
#pragma unroll
      for (int q = 0; q < 4; ++q) asm volatile("st.shared.v2.f64 [%0], {%1, %1};" ::"r"(at + q * 16), "d"(val) : "memory");
      if (FRAMES && (j & 3) == 0) f255_s[v] = __fdiv_rn(static_cast<float>(ev), 255.0f);
    }
  }
  __syncthreads();

  const int b = blockIdx.y;
  const int64_t pix0 = (static_cast<int64_t>(blockIdx.x) * THREADS + threadIdx.x) * 4;
  const int64_t HW = a.HW;
  const unsigned int full = __ballot_sync(0xffffffffu, pix0 < HW);   // the lanes that work: fixed before any divergence
  if (pix0 < HW) {
  const int N = d.N;
  const uint32_t hw32 = static_cast<uint32_t>(HW);
  const int64_t clip_pix = static_cast<int64_t>(b) * HW + pix0;
  const NoiseKey nkey = make_noise_key(d.clip_index_base + static_cast<uint64_t>(b));
  NoiseStream gs{0u, 0u, 1u};
  if (kPh) gs = noise_stream_init(static_cast<uint64_t>(pix0) >> 2, nkey, a.rk);

  const double pos = d.pos_thres[b], neg = d.neg_thres[b];
  const double mneg = -neg;
  const NoiseScale nsc = make_noise_scale(kPh ? static_cast<float>(d.base_noise_std[b]) : 0.f);
  // LUT address = uniform base + (value << 8 | lane*8): the PRMT drops byte k of the frame word into byte 1 of the lane offset
  const uint32_t lane8 = (threadIdx.x & 31u) * 8u;
  auto lut_at = [&](uint32_t w, int k) -> double {
    uint32_t off;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(off) : "r"(w), "r"(lane8), "r"(0x7604u | (static_cast<uint32_t>(k) << 4)));
    double r;
    asm("ld.shared.f64 %0, [%1+%2];" : "=d"(r) : "r"(off), "n"(kSmemWindowBase + L::lut));
    return r;
  };

  // frame n of the clip = raw frame frame_index[b][n] (pause gather, data/v2v_datasets.py:285-311), or n itself
  // (without the gather the loads walk a loop-carried pointer: ptxas would otherwise rebuild the address from the
  // block and thread indices on every trip instead of holding two registers)
  const uint8_t* fr = d.frames + (static_cast<int64_t>(b) * a.Mraw) * HW + pix0;
  const uint8_t* fwalk = fr;
  auto frame_ptr = [&](int n) -> const uint8_t* {
    if (GATHER) return fr + static_cast<uint64_t>(fnum_s[n]) * hw32;            // one IMAD.WIDE.U32
    const uint8_t* p = fwalk;                                                   // frames are requested in order
#ifndef V2V_ABL_MEM
    fwalk += hw32;
#endif
    return p;
  };
  double pot[4], lprev[4];
  float hotf[4];            // Philox hot-pixel noise is double(float) by construction: keep the float (parked in smem)
  const uint32_t w0 = ld_stream_u32(frame_ptr(0));
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    lprev[k] = lut_at(w0, k);
    hotf[k] = 0.f;
    double u = -1.0;
    if (kPh) {
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
  // Event statistics without floating point: the crossing already holds, per pixel, hu in {0, 0x3ff00000} (up) and the
  // output word ov in {0, 0x3f800000, 0xbf800000}.  Summed as plain integers (two 3-input adds each per interval),
  // su = 1023*2^20 * #up and sv = 2^23 * (127*#up + 383*#down) modulo 2^32; 1023 and 383 are odd, so both counts come
  // back exactly while #up < 4096 and #down < 512 per lane — they are unpacked every 32 intervals (<= 128 each).  A
  // multi-threshold crossing has already contributed one event here; the exact path adds the other q-1.
  uint32_t su = 0u, sv = 0u;
  unsigned int xpos = 0u, xneg = 0u;      // extra events of the multi-threshold crossings since the last flush
  // This warp's event totals.  The noise variants without frame output, which are short of registers in the loop, park them
  // in shared memory (one slot per warp, written by its first working lane: same-box -3.5 % on config 2); with the frame
  // output (config-5 shape, 25 intervals per clip) the parked form is 3.4 % slower, so those keep them in registers.
  constexpr bool kPark = STATS && kPh && !FRAMES;
  unsigned int wpos = 0u, wneg = 0u;
  uint2* wst_s = reinterpret_cast<uint2*>(dyn_smem + L::xst);
  if (kPark) {
    if ((threadIdx.x & 31) == 0) wst_s[threadIdx.x >> 5] = make_uint2(0u, 0u);
    __syncwarp();
  }
  auto flush_stats = [&]() {              // unpack, add across the warp with one REDUX each; registers only
    const uint32_t nu = ((su >> 20) * 3071u) & 4095u;                     // 1023 * 3071 = 1 (mod 4096)
    const uint32_t nd = ((((sv >> 23) - 127u * nu) & 511u) * 127u) & 511u;   // 383 * 127 = 1 (mod 512)
    const unsigned int rp = __reduce_add_sync(full, nu + xpos), rn = __reduce_add_sync(full, nd + xneg);
    xpos = xneg = 0u;
    if (kPark) {
      if ((threadIdx.x & 31) == (__ffs(full) - 1)) {
        uint2 w = wst_s[threadIdx.x >> 5];
        w.x += rp, w.y += rn;
        wst_s[threadIdx.x >> 5] = w;
      }
    } else {
      wpos += rp, wneg += rn;
    }
    su = sv = 0u;
  };

  bool lane_hot = false;
#pragma unroll
  for (int k = 0; k < 4; ++k) lane_hot = lane_hot || hotf[k] != 0.f;
  // warp-uniform: a real branch that 94 % of the warps never take (hot_pixel_fraction <= 1e-3)
  // warp-uniform: the ~6 % of the warps that own a hot pixel (hot_pixel_fraction <= 1e-3) send EVERY interval through
  // the exact path below, which adds the hot noise there (:49; x + 0.0 == x for everyone else): no hot-pixel code and no
  // extra branch in the common path.  Below 2*min(pos,neg) at most one threshold is crossed.
  const bool any_hot = kPh && __any_sync(full, lane_hot);
  if (kPh) hot_s[threadIdx.x] = make_float4(hotf[0], hotf[1], hotf[2], hotf[3]);   // own slot: no barrier needed
  const double thr2 = any_hot ? -1.0 : __dadd_rn(fmin(pos, neg), fmin(pos, neg));

  const uint32_t dir_lane = (threadIdx.x & 7u) * 16u;     // this lane's (= this group's) direction sub-table
  const double pos2 = __dadd_rn(pos, pos), mneg2 = -__dadd_rn(neg, neg);
  // Software pipeline of the generator: the radius (lg2, sqrt on the SFU) and the table offset of interval i+1 are
  // produced while interval i is integrated, so their latency never sits in front of the potential's FP64 chain.
  // The stream is consumed in the same order as in the generic kernel (two words per interval).
  float nrad[2] = {0.f, 0.f};
  uint32_t ndir[2] = {0u, 0u};
  auto draw_next = [&]() {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint32_t nw = noise_stream_next(gs);
      nrad[h] = noise_word_radius(nw, nsc);
      ndir[h] = ((nw >> 17) & 0x7f80u) | dir_lane;
    }
  };
#ifndef V2V_EXP_NOPIPE
  if (kPh) draw_next();
#endif
  auto step = [&](const uint32_t w) {
    float o[4];
    uint32_t hu[4];
    double x0[4];
    bool rare = false;
    double rad[2], dirv[4];
#ifdef V2V_ABL_NOISE
    if (kPh) { rad[0] = rad[1] = 0.01; dirv[0] = dirv[2] = 0.5; dirv[1] = dirv[3] = -0.25; }
    if (false) {
#else
    if (kPh) {      // the four noise values of this interval from the two words drawn one interval ago, then the next draw
#endif
#pragma unroll
#ifdef V2V_EXP_NOPIPE
      draw_next();
#endif
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        rad[h] = static_cast<double>(nrad[h]);
        asm("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(dirv[2 * h]), "=d"(dirv[2 * h + 1]) : "r"(ndir[h]), "n"(kSmemWindowBase + L::dir));
      }
#ifndef V2V_EXP_NOPIPE
      draw_next();
#endif
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double Lk = lut_at(w, k);
      double x = __dadd_rn(pot[k], __dsub_rn(Lk, lprev[k]));          // :42-43
      lprev[k] = Lk;
      if (kPh) x = __fma_rn(rad[k >> 1], dirv[k], x);                 // :46-48: the product is exact, one rounding
      x0[k] = x;
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
#ifdef V2V_ABL_TRIGGER
      rare = false;
#endif
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      double x = x0[k];
#ifdef V2V_ABL_CROSS
      o[k] = __int_as_float(__double2hiint(x) & 0x3f800000);
      pot[k] = x;
      continue;
#endif
#ifdef V2V_SEL_CROSS
      if (false) single_cross(x, o[k], hu[k], pos, mneg);
#else
      if (kPh) single_cross(x, o[k], hu[k], pos, mneg);
#endif
      else single_cross_sel(x, o[k], hu[k], pos, mneg, neg);          // :51-58 with q in {0,1}
      pot[k] = x;
    }
#ifdef V2V_ABL_STATS
    if (false) {
#else
    if (STATS) {            // from the single-crossing words (the exact path below only adds its surplus)
#endif
      su += hu[0] + hu[1];
      su += hu[2] + hu[3];
      sv += __float_as_uint(o[0]) + __float_as_uint(o[1]);
      sv += __float_as_uint(o[2]) + __float_as_uint(o[3]);
    }
    if (rare) {                 // a multi-threshold crossing somewhere in the warp, or a warp that owns a hot pixel
      float hk[4] = {0.f, 0.f, 0.f, 0.f};
      if (any_hot) {
        const float4 h = hot_s[threadIdx.x];
        hk[0] = h.x, hk[1] = h.y, hk[2] = h.z, hk[3] = h.w;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {                                   // the exact conditions, only evaluated in here
        const double xr = __dadd_rn(x0[k], static_cast<double>(hk[k]));                 // :49
        if (xr >= pos2 || xr <= mneg2 || hk[k] != 0.f) {              // redo this pixel from x: exact for any x
          const bool down = xr < 0.0;
          double q;
          const double an = multi_cross(fabs(xr), down ? neg : pos, cta_rcp[down ? 1 : 0], &q);
          pot[k] = down ? -an : an;
          // Replace what the common path counted for this pixel (from x without the hot noise).  Three forms, chosen per
          // variant by same-box A/Bs (profiles/r02_esim_experiments.md section 7).  Asking x0 >= pos again in here makes the
          // compiler carry the common path's predicates across the branch as a bit mask (8 ALU-pipe LOP3 per interval: with
          // statistics the noise kernel ran 15 % slower than without).  Noise variants: take the very words the common
          // path added (hu[k], the bits of o[k]) back out of the packed sums (-5.7 %).  Noise-free variants: read the count
          // back from the sign of the output word (-6 %; the subtract form costs them 2 %).
          const uint32_t counted_v = __float_as_uint(o[k]);
          o[k] = static_cast<float>(down ? -q : q);
          if (STATS) {
            const int qi = static_cast<int>(q);
            if (kPh && !FRAMES) {
              su -= hu[k];
              sv -= counted_v;
              xpos += static_cast<unsigned int>(down ? 0 : qi);
              xneg += static_cast<unsigned int>(down ? qi : 0);
            } else if (kPh) {       // (with the frame output in the loop the plain compare form is 4 % faster: config-5 shape)
              xpos += static_cast<unsigned int>((down ? 0 : qi) - (x0[k] >= pos ? 1 : 0));
              xneg += static_cast<unsigned int>((down ? qi : 0) - (x0[k] <= mneg ? 1 : 0));
            } else {
              const int counted = static_cast<int>(counted_v);
              xpos += static_cast<unsigned int>((down ? 0 : qi) - (counted > 0 ? 1 : 0));
              xneg += static_cast<unsigned int>((down ? qi : 0) - (counted < 0 ? 1 : 0));
            }
          }
        }
      }
    }
#ifdef V2V_ABL_MEM
    if (a.T < 0)
#endif
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

  const int M = N - 1;                       // intervals
  if (STAGED) {
    // ---- main loop, frames staged through shared memory: thread 0 keeps kRingDepth tiles of THREADS*4 bytes in flight
    // with bulk async copies that complete on the slot's "full" mbarrier; a warp waits for the slot, reads its words with
    // one LDS.32 per lane and releases the slot on its "empty" mbarrier; thread 0 refills a slot kRingLag intervals after
    // it was read, when every warp of the tile has released it.  No frame word is held in registers across intervals and
    // no lane computes a global address.
    const int64_t tile0 = static_cast<int64_t>(blockIdx.x) * THREADS * 4;
    const uint8_t* tile_src = d.frames + (static_cast<int64_t>(b) * a.Mraw) * HW + tile0;
    const int tile_bytes = static_cast<int>(min(static_cast<int64_t>(THREADS) * 4, HW - tile0));
    if (threadIdx.x == 0) {
      for (int q = 0; q < kRingDepth && q < M; ++q) {
        mbar_expect_tx(full_a + 8 * q, tile_bytes);
        bulk_g2s(ring_a + q * THREADS * 4, tile_src + static_cast<int64_t>(1 + q) * HW, tile_bytes, full_a + 8 * q);
      }
    }
#pragma unroll 4
    for (int i = 0; i < M; ++i) {
      const int q = i % kRingDepth;
      mbar_wait(full_a + 8 * q, (i / kRingDepth) & 1);
      uint32_t w;
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(ring_a + q * THREADS * 4 + threadIdx.x * 4));
      step(w);
      if (STATS && (i & 31) == 31) flush_stats();
      __syncwarp(full);
      if ((threadIdx.x & 31) == 0) mbar_arrive(empty_a + 8 * q);
      if (threadIdx.x == 0 && i >= kRingLag && i - kRingLag + kRingDepth < M) {
        const int j = i - kRingLag, qj = j % kRingDepth;
        mbar_wait(empty_a + 8 * qj, (j / kRingDepth) & 1);
        mbar_expect_tx(full_a + 8 * qj, tile_bytes);
        bulk_g2s(ring_a + qj * THREADS * 4, tile_src + static_cast<int64_t>(1 + j + kRingDepth) * HW, tile_bytes, full_a + 8 * qj);
      }
    }
  } else {
  // ---- main loop: kPF frames per trip, the next trip's words already in flight ----
  const int trips = M / kPF;
  uint32_t cur[kPF], nxt[kPF];
  if (trips > 0) {
#pragma unroll
    for (int u = 0; u < kPF; ++u) cur[u] = ld_stream_u32(frame_ptr(1 + u));
  }
  int i = 1;
  for (int t = 0; t < trips; ++t) {
    if (t + 1 < trips) {
#pragma unroll
      for (int u = 0; u < kPF; ++u) nxt[u] = ld_stream_u32(frame_ptr(i + kPF + u));
    }
#pragma unroll
    for (int u = 0; u < kPF; ++u) step(cur[u]);
    i += kPF;
#pragma unroll
    for (int u = 0; u < kPF; ++u) cur[u] = nxt[u];
    if (STATS && (t & (kFlushTrips - 1)) == kFlushTrips - 1) flush_stats();
  }
  for (; i < N; ++i) step(ld_stream_u32(frame_ptr(i)));               // ragged tail (< kPF intervals)
  }

  if (d.potential_out) {
#pragma unroll
    for (int k = 0; k < 4; ++k) d.potential_out[clip_pix + k] = pot[k];
  }
  if (STATS) {
    // One pair of global reductions per warp, no CTA barrier and no shared-memory stage: a warp that is done leaves.
    flush_stats();
    if ((threadIdx.x & 31) == (__ffs(full) - 1)) {
      if (kPark) wpos = wst_s[threadIdx.x >> 5].x, wneg = wst_s[threadIdx.x >> 5].y;
      if (wpos) atomicAdd(reinterpret_cast<unsigned long long*>(d.stats + 2 * blockIdx.y), static_cast<unsigned long long>(wpos));
      if (wneg) atomicAdd(reinterpret_cast<unsigned long long*>(d.stats + 2 * blockIdx.y) + 1, static_cast<unsigned long long>(wneg));
    }
  }
  }  // valid
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute of each instantiation: set it on the first
// launch on a device, not on every launch.
template <int NOISE, bool FRAMES, bool STATS, bool GATHER, int THREADS, int CTAS, bool STAGED = false>
int launch_variant(const EsimArgs& a, cudaStream_t s) {
  using L = FastSmem<NOISE, FRAMES, GATHER, THREADS>;
  static std::atomic<uint64_t> configured{0};
  int dev = 0;
  V2V_CUDA(cudaGetDevice(&dev));
  const uint64_t bit = 1ull << (dev & 63);
  const size_t smem = STAGED ? L::bytes_staged(a.d.N) : L::bytes(a.d.N);
  if (!(configured.load(std::memory_order_acquire) & bit)) {
    V2V_CUDA(cudaFuncSetAttribute(esim_fast_kernel<NOISE, FRAMES, STATS, GATHER, THREADS, CTAS, STAGED>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured.fetch_or(bit, std::memory_order_release);
  }
  const int64_t groups = a.HW / 4;
  dim3 grid(static_cast<unsigned int>((groups + THREADS - 1) / THREADS), static_cast<unsigned int>(a.d.B));
  esim_fast_kernel<NOISE, FRAMES, STATS, GATHER, THREADS, CTAS, STAGED><<<grid, THREADS, smem, s>>>(a);
  count_launch();
  V2V_CUDA(cudaGetLastError());
  return V2V_OK;
}

template <int NOISE, int THREADS, int CTAS>
int launch_geom(const EsimArgs& a, cudaStream_t s) {
  const bool fr = a.d.frame_out_mode != 0, st = a.d.stats != nullptr, ga = a.d.frame_index != nullptr;
#define V2V_V(FR, ST, GA) return launch_variant<NOISE, FR, ST, GA, THREADS, CTAS>(a, s)
  if (fr) {
    if (st) { if (ga) V2V_V(true, true, true); else V2V_V(true, true, false); }
    else    { if (ga) V2V_V(true, false, true); else V2V_V(true, false, false); }
  } else {
    if (st) { if (ga) V2V_V(false, true, true); else V2V_V(false, true, false); }
    else    { if (ga) V2V_V(false, false, true); else V2V_V(false, false, false); }
  }
#undef V2V_V
}

}  // namespace

bool esim_fast_eligible(const EsimArgs& a) {
  const v2v_esim_desc& d = a.d;
  if (d.frames_per_bin != 1 || d.threshold_mode != V2V_THRES_PER_CLIP) return false;
  if (d.noise_mode == V2V_NOISE_EXPLICIT) return false;
  if (d.noise_mode == V2V_NOISE_PHILOX && d.put_noise_external) return false;
  if (d.N > 16384) return false;               // the per-clip frame-number table lives in shared memory
  if (a.HW >= (1ll << 32)) return false;       // 32-bit plane size in the frame address multiply
  return true;
}

// Geometry (threads per CTA, CTAs per SM) per variant from same-box sweeps on B200 (profiles/r02_esim_experiments.md section 5);
// bits 8..11 of v2v_esim_desc.kernel_flags select another one for tuning.
int launch_esim_fast(const EsimArgs& a, cudaStream_t s) {
  const int geom = (a.d.kernel_flags >> 8) & 0xf;
  if (a.d.noise_mode == V2V_NOISE_PHILOX) {
    switch (geom) {
      case 1: return launch_geom<V2V_NOISE_PHILOX, 512, 2>(a, s);
      case 2: return launch_geom<V2V_NOISE_PHILOX, 320, 2>(a, s);
      case 3: return launch_geom<V2V_NOISE_PHILOX, 768, 1>(a, s);
      default: return launch_geom<V2V_NOISE_PHILOX, 384, 2>(a, s);
    }
  }
  // frames through a bulk-copy ring in shared memory (cp.async.bulk + mbarrier) instead of per-lane LDG: A/B in
  // profiles/r02_esim_experiments.md; needs 16-byte aligned tiles
  if ((a.d.kernel_flags & V2V_ESIM_FLAG_STAGED) && geom == 0 && !a.d.frame_index && a.HW % 16 == 0 && aligned(a.d.frames, 16)) {
    const bool fr = a.d.frame_out_mode != 0, st = a.d.stats != nullptr;
    if (fr) return st ? launch_variant<V2V_NOISE_NONE, true, true, false, 512, 2, true>(a, s) : launch_variant<V2V_NOISE_NONE, true, false, false, 512, 2, true>(a, s);
    return st ? launch_variant<V2V_NOISE_NONE, false, true, false, 512, 2, true>(a, s) : launch_variant<V2V_NOISE_NONE, false, false, false, 512, 2, true>(a, s);
  }
  switch (geom) {
    case 1: return launch_geom<V2V_NOISE_NONE, 320, 3>(a, s);
    case 2: return launch_geom<V2V_NOISE_NONE, 256, 3>(a, s);
    case 3: return launch_geom<V2V_NOISE_NONE, 384, 2>(a, s);
    default: return launch_geom<V2V_NOISE_NONE, 512, 2>(a, s);
  }
}

}  // namespace v2v

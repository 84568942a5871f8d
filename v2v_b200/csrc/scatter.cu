// Event stream -> voxel / image scatter (sm_100a).
//
// Replaces TestH5Dataset.make_voxel (reference data/testh5.py:60-90),
// events_to_voxel_torch / events_to_neg_pos_voxel_torch
// (utils/event_utils.py:466-541) and events_to_image_torch (:330-376).
//
// Voxel kernel: shared-memory privatised, write-once.  A work item is
// (window, bin, strip of rows).  Timestamps are non-decreasing inside a window,
// so every bin owns a contiguous range of the window's events: warp 0 finds it
// with a 32-ary search (same bin arithmetic as the scatter itself) while the
// other warps zero a [rows, W] tile of accumulators in shared memory; the CTA
// then scans only that range (coalesced reads of ys, 4 loads in flight per
// thread; xs/ps only for events inside the strip), adds with shared-memory
// integer atomics and streams the finished strip to HBM once.  The output is
// never zero-filled or read back: HBM traffic is the event stream (first strip
// of a bin; the others hit L2) plus 4 B per voxel cell.
//
// Accumulation is exact and order independent where the reference's is not:
//   h5 discrete  : int32 counts (reference: float64 adds of +-1, exact too);
//   h5 interp    : fixed point, round(w*2^30) split in two int32 words
//                  (reference: sequential float64) -> |err| <= n*2^-31 per cell;
//   torch modes  : float32 shared atomics (reference: sequential float32).
#include "scatter_common.cuh"


namespace v2v {
namespace {

constexpr int kScatterThreads = 512;
constexpr int kSmemBudget = 100 * 1024;   // two CTAs per SM

struct ScatterArgs {
  v2v_scatter_desc d;
  int rows_per_strip, num_strips;
  int packed16;        // h5 discrete: strips are sized for 2-byte cells
  int tile_bytes;      // bytes of the accumulator tile in dynamic shared memory (the two-tap hit list follows it)
  int generic_scan;    // V2V_SCATTER_GENERIC=1: never take the 16-bit coordinate scan (A/B and tests)
  int num_splits;      // >1: each (window, bin, strip) is shared by this many CTAs, each scanning a slice of the bin's events
                       //     into a private tile and adding it to the (pre-zeroed) output with global atomics
  int64_t* bounds;     // [Wn, bins+2]: first event with bin_floor >= k for k = -1 .. bins (from the pre-pass), or NULL
  struct WinConst* wcs; // [Wn] window constants from the pre-pass (valid when bounds != NULL)
};

// first event in [lo,hi) whose bin_floor is >= target, or hi (32-ary search by one warp)
template <int MODE>
__device__ __forceinline__ int64_t first_with_bin_ge(const v2v_scatter_desc& d, const WinConst& c, int64_t lo, int64_t hi,
                                                      double target, int lane) {
  while (hi - lo > 0) {
    const int64_t n = hi - lo;
    const int64_t step = (n + 31) / 32;
    const int64_t probe = lo + static_cast<int64_t>(lane) * step;     // probes lo, lo+step, ...
    bool ge = true;                                                   // positions >= hi count as "true"
    if (probe < hi) {
      double co;
      ge = bin_floor<MODE>(d, c, probe, &co) >= target;
    }
    const unsigned int m = __ballot_sync(0xffffffffu, ge);
    if (m == 0u) {                                                    // false at all 32 probes: answer lies after the last one
      lo = lo + 31 * step + 1;
      continue;
    }
    const int first = __ffs(m) - 1;                                   // first probing lane whose predicate holds
    if (first == 0) return lo;                                        // predicate already true at lo
    const int64_t nlo = lo + static_cast<int64_t>(first - 1) * step + 1;
    int64_t nhi = lo + static_cast<int64_t>(first) * step;
    if (nhi > hi) nhi = hi;
    if (step == 1) return nhi;
    lo = nlo;
    hi = nhi;
  }
  return lo;
}

// Pre-pass: all bin boundaries of all windows, one warp per (window, k).  Takes the dependent search
// round-trips out of the scatter kernel's critical path.
template <int MODE>
__global__ void __launch_bounds__(256) scatter_bounds_kernel(const ScatterArgs a) {
  const v2v_scatter_desc& d = a.d;
  const int B = d.num_bins;
  const int64_t total = static_cast<int64_t>(d.num_windows) * (B + 2);
  const int64_t wid = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (wid >= total) return;
  const int win = static_cast<int>(wid / (B + 2)), k = static_cast<int>(wid - static_cast<int64_t>(win) * (B + 2)) - 1;
  const int64_t e0 = d.window_offsets[win], e1 = d.window_offsets[win + 1];
  int64_t r = e0;
  WinConst wc;
  wc.e0 = e0;
  if (e1 > e0) {
    wc = window_constants<MODE>(d, e0, e1);
    r = first_with_bin_ge<MODE>(d, wc, e0, e1, static_cast<double>(k), threadIdx.x & 31);
  }
  if ((threadIdx.x & 31) == 0) {
    a.bounds[wid] = r;
    if (k == -1) a.wcs[win] = wc;
  }
}

template <int MODE>
__global__ void __launch_bounds__(kScatterThreads, 2) scatter_kernel(const ScatterArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int64_t s_range[2];
  __shared__ WinConst s_wc;
  __shared__ int s_hits;
  const v2v_scatter_desc& d = a.d;
  const int W = d.W, H = d.H, B = d.num_bins;
  const int R = a.rows_per_strip;
  constexpr bool kInterp = MODE == V2V_SCATTER_H5_INTERP;
  constexpr bool kTorch = MODE == V2V_SCATTER_TORCH_DISCRETE || MODE == V2V_SCATTER_TORCH_BILINEAR;
  constexpr bool kTwoTap = MODE == V2V_SCATTER_H5_INTERP || MODE == V2V_SCATTER_TORCH_BILINEAR;
  constexpr bool kH5 = !kTorch;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  int* acc_i = reinterpret_cast<int*>(smem_raw);
  float* acc_f = reinterpret_cast<float*>(smem_raw);
  int* acc_lo = acc_i + R * W;                  // second word of the fixed-point pair (h5 interp only)
  // two-tap modes: hit list of one scan trip ((event offset in the trip) << 16 | cell), behind the tile
  int* hit_list = reinterpret_cast<int*>(smem_raw + a.tile_bytes);

  // work item = (window, bin, strip of rows); items of one window are adjacent so its events stay in L2
  const int K = a.num_splits;
  const int64_t per_win = static_cast<int64_t>(B) * a.num_strips * K;
  // V2V_POL_SPLIT: every window is done twice, as output slot 2w with the positive-only weights and as slot 2w+1 with
  // the negative-only weights (events_to_neg_pos_voxel_torch in one launch)
  const bool psplit = d.polarity_mode == V2V_POL_SPLIT;
  const int64_t items = static_cast<int64_t>(d.num_windows) * per_win * (psplit ? 2 : 1);
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int vwin = static_cast<int>(item / per_win);                  // output slot
    const int win = psplit ? vwin >> 1 : vwin;                          // window of the event stream
    const int polm = psplit ? ((vwin & 1) ? V2V_POL_NEG_ONLY : V2V_POL_POS_ONLY) : d.polarity_mode;
    int rem = static_cast<int>(item - static_cast<int64_t>(vwin) * per_win);
    const int split = rem % K;
    rem /= K;
    const int bin = rem / a.num_strips, strip = rem - bin * a.num_strips;
    const int r0 = strip * R;
    const int rows = min(R, H - r0);
    const int64_t e0 = d.window_offsets[win], e1 = d.window_offsets[win + 1];

    // event range of this bin: one round trip to the pre-pass table, or searched here by warp 0
    if (warp == 0) {
      int64_t lo = e0, hi = e0;
      if (e1 > e0) {
        WinConst wc;
        // discrete: events with bin == b;  two-tap: events with floor(coord) in {b-1, b}
        if (a.bounds) {
          const int64_t* bw = a.bounds + static_cast<int64_t>(win) * (B + 2) + 1;      // bw[k], k = -1 .. B
          lo = bw[kTwoTap ? bin - 1 : bin];
          hi = bw[bin + 1];
          wc = a.wcs[win];
        } else {
          wc = window_constants<MODE>(d, e0, e1);
          lo = first_with_bin_ge<MODE>(d, wc, e0, e1, static_cast<double>(kTwoTap ? bin - 1 : bin), lane);
          hi = first_with_bin_ge<MODE>(d, wc, lo, e1, static_cast<double>(bin + 1), lane);
        }
        if (lane == 0) s_wc = wc;
        // events whose bin falls outside [0, B) are dropped: counted once per window
        if (lane == 0 && strip == 0 && split == 0 && d.dropped && !(psplit && (vwin & 1))) {
          long long nd = 0;
          if (bin == 0 && !kTwoTap) nd += lo - e0;                    // bin < 0 (unsorted / negative timestamps)
          if (bin == B - 1) nd += e1 - hi;                            // bin >= B
          if (nd) atomicAdd(reinterpret_cast<unsigned long long*>(d.dropped), static_cast<unsigned long long>(nd));
        }
      }
      if (K > 1) {                                   // this CTA's slice of the bin's events
        const int64_t n = hi - lo, per = (n + K - 1) / K;
        const int64_t slo = lo + split * per;
        hi = min(hi, slo + per);
        lo = min(slo, hi);
      }
      if (lane == 0) { s_range[0] = lo; s_range[1] = hi; }
    }
    __syncthreads();
    const int64_t lo = s_range[0], hi = s_range[1];
    const WinConst wc = s_wc;
    // h5 discrete: packed 16-bit counters (two cells per word, biased by 0x8000) are exact whenever the bin holds
    // at most 32767 events; otherwise the strip is done as two half-height passes with 32-bit counters
    const bool packed = MODE == V2V_SCATTER_H5_DISCRETE && a.packed16 && (hi - lo) <= 32767;
    // h5 interpolated: the two 32-bit words of a cell hold up to 65535 same-sign unit weights; a range that could exceed
    // that (a hot pixel in a multi-million-event window) accumulates in ONE 64-bit word per cell instead (same footprint)
    const bool wide = MODE == V2V_SCATTER_H5_INTERP && (hi - lo) > 65535;
    const int passes = (MODE == V2V_SCATTER_H5_DISCRETE && a.packed16 && !packed) ? 2 : 1;
    const int strip_r0 = r0, strip_rows = rows;
    for (int pass = 0; pass < passes; ++pass) {
    const int prow = passes == 2 ? (R + 1) / 2 : R;
    const int r0 = strip_r0 + pass * prow;
    const int rows = max(0, min(prow, strip_r0 + strip_rows - r0));
    if (rows == 0) break;
    const int tile_words_now = packed ? (rows * W + 1) / 2 : (kInterp ? R * W + rows * W : rows * W);   // interp: hi words at 0, lo words at R*W
    {
      int4* z = reinterpret_cast<int4*>(smem_raw);
      const int n4 = (tile_words_now + 3) / 4;
      const int zv = packed ? static_cast<int>(0x80008000u) : 0;
      for (int i = threadIdx.x; i < n4; i += kScatterThreads) z[i] = make_int4(zv, zv, zv, zv);
    }
    __syncthreads();

    long long ndrop = 0;
    constexpr int kU = 8;                                  // events per thread in flight
    const bool fast16 = (d.xs_dtype == V2V_U16 || d.xs_dtype == V2V_I16) && (d.ys_dtype == V2V_U16 || d.ys_dtype == V2V_I16) &&
                        H <= 32767 && W <= 32767 && !a.generic_scan;
    for (int64_t eb = lo; eb < hi; eb += kU * kScatterThreads) {
      // all loads of kU events are issued before any dependent work (one memory round-trip per trip)
      long long yv[kU], xv[kU];
      float pv[kU];
      bool okv[kU];
      if (kTwoTap) {
        if (threadIdx.x == 0) s_hits = 0;
        __syncthreads();
      }
      if (fast16) {
        // 16-bit coordinates (every h5 converter of the reference writes uint16 / int16): no dtype dispatch, 32-bit index
        // math, negative int16 values read as >= 32768 and fall out of the sensor like any other out-of-range coordinate;
        // two-tap modes append to the hit list once per warp (ballot + one shared atomic) instead of once per lane
        const uint16_t* y16 = static_cast<const uint16_t*>(d.ys);
        const uint16_t* x16 = static_cast<const uint16_t*>(d.xs);
        const int rem = static_cast<int>(min(hi - eb, static_cast<int64_t>(kU * kScatterThreads)));   // events of this trip
        const bool dropcheck = strip == 0 && pass == 0;
        const int tid = static_cast<int>(threadIdx.x);
        int yq[kU], xq[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const int o = u * kScatterThreads + tid;
          okv[u] = o < rem;
          yq[u] = okv[u] ? static_cast<int>(y16[eb + o]) : 0x7fffffff;
          xq[u] = okv[u] ? static_cast<int>(x16[eb + o]) : 0x7fffffff;
          pv[u] = 0.f;
        }
        if (!kTwoTap) {                            // polarity: the common dtypes without the per-load dispatch (an indirect branch each)
          if (d.ps_dtype == V2V_U8) {
            const uint8_t* p8 = static_cast<const uint8_t*>(d.ps);
#pragma unroll
            for (int u = 0; u < kU; ++u) if (okv[u]) pv[u] = static_cast<float>(p8[eb + u * kScatterThreads + tid]);
          } else if (d.ps_dtype == V2V_F32) {
            const float* pf = static_cast<const float*>(d.ps);
#pragma unroll
            for (int u = 0; u < kU; ++u) if (okv[u]) pv[u] = pf[eb + u * kScatterThreads + tid];
          } else {
#pragma unroll
            for (int u = 0; u < kU; ++u) if (okv[u]) pv[u] = load_f32(d.ps, d.ps_dtype, eb + u * kScatterThreads + tid);
          }
        }
        // hot loop: range tests only; out-of-sensor events (rare) are collected in `bad` and counted after the loop
        bool bad = false;
        unsigned int hm[kTwoTap ? kU : 1];          // two-tap: ballots of the hits, one hit-list reservation per warp and trip
        int tot = 0;
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const int ry = yq[u] - r0;
          const bool inrow = static_cast<unsigned int>(ry) < static_cast<unsigned int>(rows);
          const bool inx = static_cast<unsigned int>(xq[u]) < static_cast<unsigned int>(W);
          const bool hit = inrow && inx;
          bad = bad || (okv[u] && ((dropcheck && !inrow && yq[u] >= H) || (inrow && !inx)));
          if (kTwoTap) {
            hm[kTwoTap ? u : 0] = __ballot_sync(0xffffffffu, hit);
            tot += __popc(hm[kTwoTap ? u : 0]);
            continue;
          }
          if (!hit) continue;
          const int cell = ry * W + xq[u];
          float pw;                                                                     // polarity -> weight
          {
            const float p = pv[u];
            if (polm == V2V_POL_POS_ONLY) pw = p > 0.f ? 1.f : 0.f;          // event_utils.py:533
            else if (polm == V2V_POL_NEG_ONLY) pw = p <= 0.f ? 1.f : 0.f;    // :534
            else pw = kH5 ? (2.f * p - 1.f) : p;                                        // testh5.py:67
          }
          if (MODE == V2V_SCATTER_H5_DISCRETE) {                                        // testh5.py:73
            if (packed) atomicAdd(&acc_i[cell >> 1], static_cast<int>(pw) * ((cell & 1) ? 65536 : 1));
            else atomicAdd(&acc_i[cell], static_cast<int>(pw));
          } else {
            atomicAdd(&acc_f[cell], pw);                                                // event_utils.py:505
          }
        }
        if (kTwoTap && tot) {                        // warp-uniform
          int base = 0;
          if (lane == 0)                             // (inline PTX: the compiler would wrap its own warp aggregation around atomicAdd)
            asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(base) : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(&s_hits))), "r"(tot) : "memory");
          base = __shfl_sync(0xffffffffu, base, 0);
          const unsigned int lt = (1u << lane) - 1u;
#pragma unroll
          for (int u = 0; u < kU; ++u) {
            const unsigned int m = hm[kTwoTap ? u : 0];
            if ((m >> lane) & 1u) hit_list[base + __popc(m & lt)] = ((u * kScatterThreads + tid) << 16) | ((yq[u] - r0) * W + xq[u]);
            base += __popc(m);
          }
        }
        if (bad) {
          // out-of-sensor events are reported once: rows by strip 0, columns by the strip that owns the row; for
          // two-tap modes by the event's primary bin only
#pragma unroll 1
          for (int u = 0; u < kU; ++u) {             // rolled, re-reading the coordinates: the register arrays stay statically indexed
            const int o = u * kScatterThreads + tid;
            if (o >= rem) break;
            const int yy = static_cast<int>(y16[eb + o]), xx = static_cast<int>(x16[eb + o]);
            const int ry = yy - r0;
            const bool inrow = static_cast<unsigned int>(ry) < static_cast<unsigned int>(rows);
            const bool inx = static_cast<unsigned int>(xx) < static_cast<unsigned int>(W);
            if ((dropcheck && !inrow && yy >= H) || (inrow && !inx)) {
              bool primary = true;
              if (kTwoTap) { double co; primary = bin_floor<MODE>(d, wc, eb + o, &co) == static_cast<double>(bin); }
              if (primary) ++ndrop;
            }
          }
        }
      } else {
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int64_t e = eb + u * kScatterThreads + threadIdx.x;
        okv[u] = e < hi;
        bool ok = true;
        yv[u] = okv[u] ? load_int(d.ys, d.ys_dtype, e, &ok) : -1;
        xv[u] = okv[u] ? load_int(d.xs, d.xs_dtype, e, &ok) : -1;
        pv[u] = (okv[u] && !kTwoTap) ? load_f32(d.ps, d.ps_dtype, e) : 0.f;
        if (!ok) yv[u] = xv[u] = -1;
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        if (!okv[u]) continue;
        const int64_t e = eb + u * kScatterThreads + threadIdx.x;
        const long long y = yv[u], ry = y - r0;
        if (ry < 0 || ry >= rows) {
          // out-of-sensor rows are reported once: by strip 0, and for two-tap modes by the event's primary bin only
          if (strip == 0 && pass == 0 && (y < 0 || y >= H)) {
            bool primary = true;
            if (kTwoTap) { double co; primary = bin_floor<MODE>(d, wc, e, &co) == static_cast<double>(bin); }
            if (primary) ++ndrop;
          }
          continue;
        }
        const long long x = xv[u];
        if (x < 0 || x >= W) {
          bool primary = true;
          if (kTwoTap) { double co; primary = bin_floor<MODE>(d, wc, e, &co) == static_cast<double>(bin); }
          if (primary) ++ndrop;
          continue;
        }
        const int cell = static_cast<int>(ry) * W + static_cast<int>(x);
        if (kTwoTap) {
          // the weight needs double-precision division: do it densely in a second phase instead of in this
          // sparsely active branch (only rows/H of the lanes get here)
          hit_list[atomicAdd(&s_hits, 1)] = (static_cast<int>(e - eb) << 16) | cell;
          continue;
        }
        float pw;                                                                     // polarity -> weight
        {
          const float p = pv[u];
          if (polm == V2V_POL_POS_ONLY) pw = p > 0.f ? 1.f : 0.f;          // event_utils.py:533
          else if (polm == V2V_POL_NEG_ONLY) pw = p <= 0.f ? 1.f : 0.f;    // :534
          else pw = kH5 ? (2.f * p - 1.f) : p;                                        // testh5.py:67
        }
        if (MODE == V2V_SCATTER_H5_DISCRETE) {                                        // testh5.py:73
          if (packed) atomicAdd(&acc_i[cell >> 1], static_cast<int>(pw) * ((cell & 1) ? 65536 : 1));
          else atomicAdd(&acc_i[cell], static_cast<int>(pw));
        } else {
          atomicAdd(&acc_f[cell], pw);                                                // event_utils.py:505
        }
      }
      }
      if (kTwoTap) {
        __syncthreads();
        const int nh = s_hits;
        for (int h = threadIdx.x; h < nh; h += kScatterThreads) {
          const int ent = hit_list[h];
          const int cell = ent & 0xffff;
          const int64_t e = eb + (ent >> 16);
          float pw;
          {
            const float p = d.ps_dtype == V2V_U8 ? static_cast<float>(static_cast<const uint8_t*>(d.ps)[e])
                          : d.ps_dtype == V2V_F32 ? static_cast<const float*>(d.ps)[e] : load_f32(d.ps, d.ps_dtype, e);
            if (polm == V2V_POL_POS_ONLY) pw = p > 0.f ? 1.f : 0.f;
            else if (polm == V2V_POL_NEG_ONLY) pw = p <= 0.f ? 1.f : 0.f;
            else pw = kH5 ? (2.f * p - 1.f) : p;
          }
          double tnd;
          bin_floor<MODE>(d, wc, e, &tnd);
          if (MODE == V2V_SCATTER_H5_INTERP) {
            const double wgt = fmax(0.0, __dsub_rn(1.0, fabs(__dsub_rn(tnd, static_cast<double>(bin)))));   // testh5.py:79
            const long long fx = __double2ll_rn(__dmul_rn(wgt * static_cast<double>(pw), 1073741824.0));
            if (wide) {
              if (fx != 0) atomicAdd(reinterpret_cast<unsigned long long*>(acc_i) + cell, static_cast<unsigned long long>(fx));
            } else if (fx != 0) {
              const int hiw = static_cast<int>(fx >> kLoBits);
              const unsigned int low = static_cast<unsigned int>(fx & ((1 << kLoBits) - 1));
              if (hiw) atomicAdd(&acc_i[cell], hiw);
              if (low) atomicAdd(reinterpret_cast<unsigned int*>(&acc_lo[cell]), low);
            }
          } else {   // TORCH_BILINEAR
            const float tn = static_cast<float>(tnd);
            const float wgt = fmaxf(0.f, __fsub_rn(1.0f, fabsf(__fsub_rn(tn, static_cast<float>(bin)))));  // event_utils.py:494
            const float v = __fmul_rn(pw, wgt);                                                            // :495
            if (v != 0.f) atomicAdd(&acc_f[cell], v);
          }
        }
        __syncthreads();
      }
    }
    if (d.dropped && ndrop) atomicAdd(reinterpret_cast<unsigned long long*>(d.dropped), static_cast<unsigned long long>(ndrop));
    __syncthreads();

    // stream the strip out once: [win, bin, r0 + r, x]
    {
      const int cells = rows * W;
      const int64_t out_base = ((static_cast<int64_t>(vwin) * B + bin) * H + r0) * W;
      auto value = [&](int i) -> double {          // float64 outputs (drop-in make_voxel)
        if (kInterp) {
          const long long tot = wide ? reinterpret_cast<const long long*>(acc_i)[i]
                                     : static_cast<long long>(acc_i[i]) * (1 << kLoBits) +
                                       static_cast<long long>(reinterpret_cast<unsigned int*>(acc_lo)[i]);
          return static_cast<double>(tot) * (1.0 / 1073741824.0);
        }
        if (packed) return static_cast<double>(static_cast<int>((static_cast<unsigned int>(acc_i[i >> 1]) >> ((i & 1) * 16)) & 0xffffu) - 32768);
        return kTorch ? static_cast<double>(acc_f[i]) : static_cast<double>(acc_i[i]);
      };
      auto valuef = [&](int i) -> float {          // float32 outputs: stay off the conversion unit where possible
        if (kInterp) {
          if (wide) return static_cast<float>(static_cast<double>(reinterpret_cast<const long long*>(acc_i)[i]) * (1.0 / 1073741824.0));
          const int h = acc_i[i];
          const unsigned int l = reinterpret_cast<unsigned int*>(acc_lo)[i];
          if ((static_cast<unsigned int>(h) | l) == 0u) return 0.f;                   // most cells of a strip hold no event
          // h*2^-15 + l*2^-30 is exact in float64 (46 significant bits at most): one rounding, to float32
          return static_cast<float>(__fma_rn(static_cast<double>(h), 1.0 / 32768.0, __dmul_rn(static_cast<double>(l), 1.0 / 1073741824.0)));
        }
        if (packed) {                               // count in [-32768, 32767]: 1.5*2^23 + n is exact and its bits are 0x4b400000 + n
          const int n = static_cast<int>((static_cast<unsigned int>(acc_i[i >> 1]) >> ((i & 1) * 16)) & 0xffffu) - 32768;
          return __fsub_rn(__int_as_float(0x4b400000 + n), 12582912.0f);
        }
        return kTorch ? acc_f[i] : static_cast<float>(acc_i[i]);
      };
      if (K > 1) {                                   // partial tile: add the non-zero cells to the pre-zeroed output
        if (hi > lo) {
          for (int i = threadIdx.x; i < cells; i += kScatterThreads) {
            const double v = value(i);
            if (v != 0.0) {
              if (d.out_dtype == V2V_F64) atomicAdd(static_cast<double*>(d.voxel) + out_base + i, v);
              else atomicAdd(static_cast<float*>(d.voxel) + out_base + i, static_cast<float>(v));
            }
          }
        }
      } else if (d.out_dtype == V2V_F64) {
        double* o = static_cast<double*>(d.voxel) + out_base;
        for (int i = threadIdx.x; i < cells; i += kScatterThreads) o[i] = value(i);
      } else {
        float* o = static_cast<float*>(d.voxel) + out_base;
        // 128-bit stores where the strip start is 16-byte aligned
        const int head = static_cast<int>((4 - (out_base & 3)) & 3);
        for (int i = threadIdx.x; i < min(head, cells); i += kScatterThreads) st_stream_f32(o + i, valuef(i));
        const int n4 = (cells - min(head, cells)) / 4;
        if (packed) {
          // Packed 16-bit counters: four cells are two words (three when the strip starts on an odd cell).  One PRMT drops a
          // counter c under the high half 0x4b3f: the float with those bits is 12517376 + c, exactly, and the count is
          // c - 32768, so one FADD finishes it (was a variable shift, a mask, an integer add and an FADD per cell, behind
          // one LDS per cell: 49 instructions per four cells, half of this kernel's instructions).
          const unsigned int* wq = reinterpret_cast<const unsigned int*>(acc_i);
          const bool odd = (head & 1) != 0;
          auto cnt = [](unsigned int w, unsigned int sel) { return __fadd_rn(__uint_as_float(__byte_perm(w, 0x4b3fu, sel)), -12550144.0f); };
          for (int q = threadIdx.x; q < n4; q += kScatterThreads) {
            const int i = head + 4 * q, wi = i >> 1;
            unsigned int a, b;
            if (!odd) {
              a = wq[wi], b = wq[wi + 1];
            } else {
              const unsigned int w0 = wq[wi], w1 = wq[wi + 1], w2 = wq[wi + 2];
              a = __funnelshift_r(w0, w1, 16), b = __funnelshift_r(w1, w2, 16);
            }
            st_stream_f32x4(o + i, cnt(a, 0x5410u), cnt(a, 0x5432u), cnt(b, 0x5410u), cnt(b, 0x5432u));
          }
        } else
        for (int q = threadIdx.x; q < n4; q += kScatterThreads) {
          const int i = head + 4 * q;
          st_stream_f32x4(o + i, valuef(i), valuef(i + 1), valuef(i + 2),
                          valuef(i + 3));
        }
        for (int i = head + 4 * n4 + threadIdx.x; i < cells; i += kScatterThreads) st_stream_f32(o + i, valuef(i));
      }
    }
    __syncthreads();
    }  // pass
  }
}

// ---- event image: zero + global atomics (legacy path, single image) ---------
struct ImageArgs {
  v2v_image_desc d;
  int Ho, Wo;
};

__global__ void image_kernel(const ImageArgs a) {
  const v2v_image_desc& d = a.d;
  long long ndrop = 0;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < d.num_events;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float w = d.ps ? load_f32(d.ps, d.ps_dtype, e) : 1.f;
    if (d.bilinear) {
      // utils/event_utils.py:352-369,176-184
      const float xf = load_f32(d.xs, d.xs_dtype, e), yf = load_f32(d.ys, d.ys_dtype, e);
      float mask = 1.f;
      if (d.clip_out_of_range) mask = (xf >= static_cast<float>(a.Wo - 1) ? 0.f : 1.f) * (yf >= static_cast<float>(a.Ho - 1) ? 0.f : 1.f);
      const float px = floorf(xf), py = floorf(yf);
      const float dx = __fsub_rn(xf, px), dy = __fsub_rn(yf, py);
      const long long ix = static_cast<long long>(__fmul_rn(px, mask)), iy = static_cast<long long>(__fmul_rn(py, mask));
      const float mw = __fmul_rn(w, mask);
      if (ix < 0 || iy < 0 || ix + 1 >= a.Wo || iy + 1 >= a.Ho) { ++ndrop; continue; }
      float* img = static_cast<float*>(d.image);
      const float omx = __fsub_rn(1.0f, dx), omy = __fsub_rn(1.0f, dy);
      atomicAdd(img + iy * a.Wo + ix, __fmul_rn(__fmul_rn(mw, omx), omy));
      atomicAdd(img + iy * a.Wo + ix + 1, __fmul_rn(__fmul_rn(mw, dx), omy));
      atomicAdd(img + (iy + 1) * a.Wo + ix, __fmul_rn(__fmul_rn(mw, omx), dy));
      atomicAdd(img + (iy + 1) * a.Wo + ix + 1, __fmul_rn(__fmul_rn(mw, dx), dy));
    } else {
      bool ok = true;
      const long long x = load_int(d.xs, d.xs_dtype, e, &ok), y = load_int(d.ys, d.ys_dtype, e, &ok);
      if (!ok || x < 0 || y < 0 || x >= a.Wo || y >= a.Ho) { ++ndrop; continue; }
      const int64_t o = y * a.Wo + x;
      if (d.out_dtype == V2V_I64) atomicAdd(static_cast<unsigned long long*>(d.image) + o, 1ULL);
      else if (d.out_dtype == V2V_F64) atomicAdd(static_cast<double*>(d.image) + o, (d.ps && d.ps_dtype == V2V_F64) ? static_cast<const double*>(d.ps)[e] : static_cast<double>(w));
      else atomicAdd(static_cast<float*>(d.image) + o, w);
    }
  }
  if (d.dropped && ndrop) atomicAdd(reinterpret_cast<unsigned long long*>(d.dropped), static_cast<unsigned long long>(ndrop));
}

// Contiguous-range modes need non-decreasing timestamps inside every window.  One pass over the timestamps (8 bytes per
// event, ~1 % of the scatter's traffic): status[0] += number of positions where ts[e+1] < ts[e] inside a window.
__global__ void __launch_bounds__(256) sorted_check_kernel(const v2v_scatter_desc d, long long* status) {
  long long bad = 0;
  const int64_t first = d.num_windows > 0 ? d.window_offsets[0] : 0, last = d.num_windows > 0 ? d.window_offsets[d.num_windows] : 0;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  auto is_window_start = [&](int64_t e) {            // rare: a decrease onto the first event of a window is legitimate
    int lo = 0, hi = d.num_windows;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (d.window_offsets[mid] <= e) lo = mid; else hi = mid;
    }
    return d.window_offsets[lo] == e;
  };
  if (d.ts_dtype == V2V_F64) {
    const double* t = static_cast<const double*>(d.ts);
    for (int64_t e0 = first + (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; e0 + 1 < last; e0 += stride * 4) {
      double v[5];                                   // four comparisons per trip, the five loads in flight together
#pragma unroll
      for (int k = 0; k < 5; ++k) v[k] = e0 + k < last ? t[e0 + k] : 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (e0 + k + 1 < last && v[k + 1] < v[k] && !is_window_start(e0 + k + 1)) ++bad;
    }
  } else {
    for (int64_t e = first + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e + 1 < last; e += stride)
      if (load_f32(d.ts, d.ts_dtype, e + 1) < load_f32(d.ts, d.ts_dtype, e) && !is_window_start(e + 1)) ++bad;
  }
  if (bad) atomicAdd(reinterpret_cast<unsigned long long*>(status), static_cast<unsigned long long>(bad));
}

bool coord_dtype_ok(int t) { return t == V2V_U16 || t == V2V_I16 || t == V2V_I32 || t == V2V_I64 || t == V2V_F32 || t == V2V_F64 || t == V2V_U8; }

}  // namespace
}  // namespace v2v

extern "C" int v2v_events_to_voxel(const v2v_scatter_desc* desc, void* stream) {
  using namespace v2v;
  V2V_REQUIRE(desc != nullptr, V2V_ERR_INVALID_ARG, "desc is NULL");
  const v2v_scatter_desc& d = *desc;
  V2V_REQUIRE(d.num_events >= 0 && d.num_windows >= 0 && d.num_bins >= 1 && d.H >= 0 && d.W >= 0, V2V_ERR_INVALID_ARG,
              "bad sizes Ne=%lld Wn=%d bins=%d H=%d W=%d", static_cast<long long>(d.num_events), d.num_windows, d.num_bins, d.H, d.W);
  V2V_REQUIRE(d.mode >= 0 && d.mode <= 3, V2V_ERR_INVALID_ARG, "bad mode %d", d.mode);
  V2V_REQUIRE(d.polarity_mode >= 0 && d.polarity_mode <= 3, V2V_ERR_INVALID_ARG, "bad polarity_mode %d", d.polarity_mode);
  V2V_REQUIRE(d.out_dtype == V2V_F32 || d.out_dtype == V2V_F64, V2V_ERR_INVALID_ARG, "out_dtype must be F32 or F64");
  if (d.num_windows == 0 || d.H == 0 || d.W == 0) return V2V_OK;
  V2V_REQUIRE(d.window_offsets && d.voxel, V2V_ERR_INVALID_ARG, "window_offsets and voxel must be non-NULL");
  V2V_REQUIRE(d.num_events == 0 || (d.xs && d.ys && d.ts && d.ps), V2V_ERR_INVALID_ARG, "event arrays must be non-NULL");
  V2V_REQUIRE(coord_dtype_ok(d.xs_dtype) && coord_dtype_ok(d.ys_dtype), V2V_ERR_INVALID_ARG, "bad coordinate dtype");
  const bool h5 = d.mode == V2V_SCATTER_H5_DISCRETE || d.mode == V2V_SCATTER_H5_INTERP;
  V2V_REQUIRE(!h5 || d.ts_dtype == V2V_F64 || d.ts_dtype == V2V_F32, V2V_ERR_INVALID_ARG, "h5 modes need F64 or F32 timestamps");
  V2V_REQUIRE(h5 || d.ts_dtype == V2V_F32 || d.ts_dtype == V2V_F64, V2V_ERR_INVALID_ARG, "torch modes need float timestamps");
  V2V_REQUIRE(d.ps_dtype == V2V_U8 || d.ps_dtype == V2V_I8 || d.ps_dtype == V2V_F32, V2V_ERR_INVALID_ARG, "bad polarity dtype");

  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // h5 interpolated mode with enough workspace: counting sort by strip, one visit per event, any event order
  if (scatter_sorted_eligible(d) && !(d.kernel_flags & V2V_SCATTER_FLAG_RANGES)) return launch_scatter_sorted(d, s);
  if (d.unsorted && d.num_events > 1) {          // the kernels below assume sorted windows: report violations
    sorted_check_kernel<<<148 * 8, 256, 0, s>>>(d, d.unsorted);
    count_launch();
  }

  ScatterArgs a;
  a.d = d;
  // one tile = one bin x a strip of rows; sized so that two CTAs share an SM
  a.packed16 = d.mode == V2V_SCATTER_H5_DISCRETE && !(d.kernel_flags & V2V_SCATTER_FLAG_NO_PACKED16);
  const int cell_bytes = d.mode == V2V_SCATTER_H5_INTERP ? 8 : (a.packed16 ? 2 : 4);
  const int64_t row_bytes = static_cast<int64_t>(d.W) * cell_bytes;
  int budget = d.tuning_smem_kb > 0 ? d.tuning_smem_kb * 1024 : kSmemBudget;
  const bool two_tap = d.mode == V2V_SCATTER_H5_INTERP || d.mode == V2V_SCATTER_TORCH_BILINEAR;
  const int list_bytes = two_tap ? 8 * kScatterThreads * 4 : 0;        // one scan trip: 8 events per thread, 4 bytes each
  budget -= list_bytes;
  if (two_tap && budget > 65535 * cell_bytes) budget = 65535 * cell_bytes;   // hit-list entries hold 16-bit cell indices
  V2V_REQUIRE(row_bytes <= budget, V2V_ERR_UNSUPPORTED, "W=%d does not fit one shared-memory row tile", d.W);
  a.rows_per_strip = static_cast<int>(budget / row_bytes);
  if (a.rows_per_strip > d.H) a.rows_per_strip = d.H;
  a.num_strips = (d.H + a.rows_per_strip - 1) / a.rows_per_strip;
  a.rows_per_strip = (d.H + a.num_strips - 1) / a.num_strips;                 // balance the strips
  // (the 32-bit fallback of a packed strip uses two passes of ceil(R/2) rows: one extra row of slack)
  a.tile_bytes = static_cast<int>((static_cast<size_t>(a.rows_per_strip + (a.packed16 ? 1 : 0)) * row_bytes + 31) / 16 * 16);
  const size_t smem = static_cast<size_t>(a.tile_bytes) + list_bytes;
  const int vw = d.polarity_mode == V2V_POL_SPLIT ? 2 : 1;          // output slots per window
  int64_t items = static_cast<int64_t>(d.num_windows) * vw * d.num_bins * a.num_strips;
  int dev = 0, sms = 148;
  V2V_CUDA(cudaGetDevice(&dev));
  V2V_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  // few, large windows (e.g. the offline cache builder: one window of millions of events) cannot fill the GPU with
  // one CTA per (window, bin, strip): split every bin's event range over several CTAs
  a.num_splits = 1;
  a.generic_scan = (d.kernel_flags & V2V_SCATTER_FLAG_GENERIC_SCAN) != 0;
  const int64_t ev_per_item = d.num_events / (static_cast<int64_t>(d.num_windows) * d.num_bins > 0 ? static_cast<int64_t>(d.num_windows) * d.num_bins : 1);
  if (items < 2LL * sms && ev_per_item > 16384) {
    int64_t k = (4LL * sms + items - 1) / items;
    const int64_t kmax = ev_per_item / 8192;
    if (k > kmax) k = kmax;
    if (k > 64) k = 64;
    if (k > 1) a.num_splits = static_cast<int>(k);
  }
  if (d.tuning_splits > 0) a.num_splits = d.tuning_splits;
  items *= a.num_splits;
  const int grid = static_cast<int>(items < 4LL * sms ? items : 4LL * sms);
  a.bounds = nullptr;
  a.wcs = nullptr;
  const size_t nb = static_cast<size_t>(d.num_windows) * (d.num_bins + 2) * sizeof(int64_t);
  const size_t need = nb + static_cast<size_t>(d.num_windows) * sizeof(WinConst);
  if (d.workspace && static_cast<size_t>(d.workspace_bytes) >= need && aligned(d.workspace, 8)) {
    a.bounds = static_cast<int64_t*>(d.workspace);
    a.wcs = reinterpret_cast<WinConst*>(static_cast<char*>(d.workspace) + nb);
  }
  if (a.num_splits > 1)
    V2V_CUDA(cudaMemsetAsync(d.voxel, 0, static_cast<size_t>(d.num_windows) * vw * d.num_bins * d.H * d.W * (d.out_dtype == V2V_F64 ? 8 : 4), s));
#define V2V_LAUNCH(M)                                                                                      \
  do {                                                                                                     \
    if (a.bounds) {                                                                                        \
      const int64_t warps = static_cast<int64_t>(d.num_windows) * (d.num_bins + 2);                        \
      scatter_bounds_kernel<M><<<static_cast<int>((warps * 32 + 255) / 256), 256, 0, s>>>(a);              \
      count_launch();                                                                                      \
    }                                                                                                      \
    V2V_CUDA(cudaFuncSetAttribute(scatter_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))); \
    scatter_kernel<M><<<grid, kScatterThreads, smem, s>>>(a);                                              \
  } while (0)
  switch (d.mode) {
    case V2V_SCATTER_H5_DISCRETE: V2V_LAUNCH(V2V_SCATTER_H5_DISCRETE); break;
    case V2V_SCATTER_H5_INTERP: V2V_LAUNCH(V2V_SCATTER_H5_INTERP); break;
    case V2V_SCATTER_TORCH_DISCRETE: V2V_LAUNCH(V2V_SCATTER_TORCH_DISCRETE); break;
    default: V2V_LAUNCH(V2V_SCATTER_TORCH_BILINEAR); break;
  }
#undef V2V_LAUNCH
  count_launch();
  V2V_CUDA(cudaGetLastError());
  return V2V_OK;
}

extern "C" int64_t v2v_scatter_workspace_bytes(const v2v_scatter_desc* desc) {
  using namespace v2v;
  if (!desc || desc->num_windows < 0 || desc->num_bins < 1) return 0;
  int R = 0, S = 0;
  const size_t ranges = static_cast<size_t>(desc->num_windows) * ((desc->num_bins + 2) * 8 + sizeof(WinConst)) + 64;
  const size_t sorted = desc->mode == V2V_SCATTER_H5_INTERP ? scatter_sorted_workspace_bytes(*desc, &R, &S) : 0;
  return static_cast<int64_t>(sorted > ranges ? sorted : ranges);
}

extern "C" int v2v_events_to_image(const v2v_image_desc* desc, void* stream) {
  using namespace v2v;
  V2V_REQUIRE(desc != nullptr, V2V_ERR_INVALID_ARG, "desc is NULL");
  const v2v_image_desc& d = *desc;
  V2V_REQUIRE(d.num_events >= 0 && d.H >= 0 && d.W >= 0, V2V_ERR_INVALID_ARG, "bad sizes");
  V2V_REQUIRE(d.out_dtype == V2V_F32 || d.out_dtype == V2V_F64 || d.out_dtype == V2V_I64, V2V_ERR_INVALID_ARG, "bad out_dtype");
  V2V_REQUIRE(!d.bilinear || d.out_dtype == V2V_F32, V2V_ERR_UNSUPPORTED, "bilinear images are float32");
  V2V_REQUIRE(d.out_dtype != V2V_I64 || d.ps == nullptr, V2V_ERR_UNSUPPORTED, "count maps take no weights");
  ImageArgs a;
  a.d = d;
  a.Ho = d.H + ((d.bilinear && d.padding) ? 1 : 0);
  a.Wo = d.W + ((d.bilinear && d.padding) ? 1 : 0);
  if (a.Ho == 0 || a.Wo == 0) return V2V_OK;
  V2V_REQUIRE(d.image, V2V_ERR_INVALID_ARG, "image is NULL");
  V2V_REQUIRE(d.num_events == 0 || (d.xs && d.ys), V2V_ERR_INVALID_ARG, "xs/ys NULL");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t esz = d.out_dtype == V2V_F32 ? 4 : 8;
  V2V_CUDA(cudaMemsetAsync(d.image, 0, esz * a.Ho * a.Wo, s));
  if (d.num_events > 0) {
    const int threads = 256;
    int64_t blocks = (d.num_events + threads - 1) / threads;
    if (blocks > 148 * 16) blocks = 148 * 16;
    image_kernel<<<static_cast<int>(blocks), threads, 0, s>>>(a);
    count_launch();
    V2V_CUDA(cudaGetLastError());
  }
  return V2V_OK;
}

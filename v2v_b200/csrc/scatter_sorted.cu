// Event stream -> voxel, h5 interpolated mode (TestH5Dataset.make_voxel with interpolate_bins, reference
// data/testh5.py:60-69,74-80), one visit per event and no ordering precondition.
//
// The contiguous-range kernel in scatter.cu visits an event once per (tap, strip) item whose bin range contains it —
// 16 times for 5 bins of 260 rows — because a bin's events are not ordered by row.  Here a counting sort by
// (window, strip of rows) comes first:
//   1. window constants (one thread per window: tau_last, the denominator of :76-77);
//   2. count   : every event -> its window (contiguous chunks, a step search from the chunk's first window) and its strip
//                y / R (only the row coordinate is read); chunk-local counters in shared memory, then one global
//                reduction per touched (window, strip) and CTA;
//   3. scan    : exclusive prefix sum of the counts (one CTA; windows x strips is a few thousand entries);
//   4. fill    : the same traversal builds an 8-byte record per event {cell in the strip, floor(t_norm)+1, polarity sign,
//                frac(t_norm) as 2^-30 fixed point}, reserves one range per touched (window, strip) and CTA in the
//                segment cursors and writes the records there;
//   5. scatter : one work item per (window, strip) holds ALL bins of its rows in shared memory as exact fixed-point
//                pairs, reads only its own records (coalesced 8-byte loads), adds both temporal taps of an event in
//                the same visit (max(0, 1-|t_norm-b|) is 1-frac for b = floor and frac for b = floor+1) and streams
//                the finished planes out once with 128-bit stores.
// Accumulation is integer, so the result does not depend on the (atomic) record order: deterministic; exact to
// n * 2^-31 per cell like the contiguous-range kernel for items of more than 255 records, n * 2^-24 for the others
// (one 32-bit word per cell, half the shared atomics).  Timestamps need not be sorted: every event's bin comes from its
// own t_norm, as in the reference (events whose taps fall outside [0, bins) are dropped and counted).
// Extra HBM traffic: three reads of the 13-byte events (the 2nd and 3rd mostly from L2) + 8 bytes written and read per
// event, against 4 bytes per voxel cell written once.
#include "scatter_common.cuh"

namespace v2v {
namespace {

constexpr int kChunk = 2048;          // events per CTA in the count / fill passes
constexpr int kSortThreads = 256;
constexpr int kItemThreads = 256;
constexpr int kTileBudget = 28 * 1024;    // eight CTAs per SM: an item is short (~200 records), its latency is hidden across CTAs

struct SortedArgs {
  v2v_scatter_desc d;
  int R, S;               // rows per strip, strips per window
  uint32_t r_magic;       // y / R == (y * r_magic) >> 32 for y < 65536
  int64_t items;          // Wn * S
  WinConst* wcs;          // [Wn]
  uint32_t* counts;       // [items]
  uint32_t* cursor;       // [items + 1]: exclusive prefix (segment starts), advanced by the fill pass
  uint32_t* starts;       // [items + 1]: exclusive prefix, kept
  uint2* records;         // [Ne]
};

__global__ void window_constants_kernel(const SortedArgs a) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= a.d.num_windows) return;
  const int64_t e0 = a.d.window_offsets[w], e1 = a.d.window_offsets[w + 1];
  WinConst c;
  c.e0 = e0;
  c.h5_tpb = c.h5_den = 0.0;
  c.t_first = c.t_span = c.t_tpb = 0.f;
  if (e1 > e0) {
    c = window_constants<V2V_SCATTER_H5_INTERP>(a.d, e0, e1);
    c.h5_den = __ddiv_rn(static_cast<double>(a.d.num_bins - 1), c.h5_den);       // the fill pass multiplies
    if (a.d.ts_dtype == V2V_F64) c.h5_tpb = static_cast<const double*>(a.d.ts)[e0];   // first timestamp of the window
    else c.t_first = static_cast<const float*>(a.d.ts)[e0];
  }
  a.wcs[w] = c;
}

// Both passes walk the stream in chunks of kChunk consecutive events per CTA.  A chunk rarely spans more than a couple
// of windows, so the (window, strip) counters of the chunk live in shared memory: one shared atomic per event, and
// one global atomic per touched (window, strip) per CTA (a chunk that spans more than kMaxWin windows falls back to one
// global atomic per event).
constexpr int kMaxWin = 4;

__device__ __forceinline__ int first_window_of(const v2v_scatter_desc& d, int64_t e) {     // last window with start <= e
  int lo = 0, hi = d.num_windows;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (d.window_offsets[mid] <= e) lo = mid; else hi = mid;
  }
  return lo;
}

// Count pass: only the row coordinate is read (2 bytes per event).  It over-counts (events that the fill pass drops for
// their column, bin or zero weight keep their slot): segments are sized by it, the fill pass records how much of each
// segment is used.
__global__ void __launch_bounds__(kSortThreads) count_events_kernel(const SortedArgs a) {
  const v2v_scatter_desc& d = a.d;
  extern __shared__ uint32_t h_s[];                 // [kMaxWin * S]
  __shared__ int s_w0, s_w1;
  const int64_t c0 = static_cast<int64_t>(blockIdx.x) * kChunk;
  const int64_t c1 = min(d.num_events, c0 + kChunk);
  if (threadIdx.x == 0) {
    s_w0 = first_window_of(d, c0);
    s_w1 = first_window_of(d, c1 - 1);
  }
  for (int i = threadIdx.x; i < kMaxWin * a.S; i += kSortThreads) h_s[i] = 0u;
  __syncthreads();
  const int w0 = s_w0;
  const bool local = s_w1 - w0 < kMaxWin;
  const bool c16 = d.ys_dtype == V2V_U16 || d.ys_dtype == V2V_I16;
  const int64_t first = d.window_offsets[0], last = d.window_offsets[d.num_windows];
  int w = w0;
  for (int64_t e = c0 + threadIdx.x; e < c1; e += kSortThreads) {
    if (e < first || e >= last) continue;
    while (e >= d.window_offsets[w + 1]) ++w;
    long long y;
    if (c16) {
      y = static_cast<const uint16_t*>(d.ys)[e];
    } else {
      bool ok = true;
      y = load_int(d.ys, d.ys_dtype, e, &ok);
      if (!ok) y = -1;
    }
    if (y < 0 || y >= d.H) continue;
    const uint32_t strip = static_cast<uint32_t>((static_cast<uint64_t>(y) * a.r_magic) >> 32);
    if (local) atomicAdd(&h_s[(w - w0) * a.S + strip], 1u);
    else atomicAdd(a.counts + static_cast<int64_t>(w) * a.S + strip, 1u);
  }
  __syncthreads();
  if (local) {
    const int n = min(kMaxWin, d.num_windows - w0) * a.S;
    for (int i = threadIdx.x; i < n; i += kSortThreads)
      if (h_s[i]) atomicAdd(a.counts + static_cast<int64_t>(w0) * a.S + i, h_s[i]);
  }
}

// Fill pass: every thread keeps the records of its kChunk / kSortThreads events in registers, takes a chunk-local slot per
// event from the shared counters, the CTA reserves one range per touched (window, strip) in the global cursors, and the
// records go to their slots.
__global__ void __launch_bounds__(kSortThreads) fill_events_kernel(const SortedArgs a) {
  const v2v_scatter_desc& d = a.d;
  extern __shared__ uint32_t h_s[];                 // [kMaxWin * S] counters, then [kMaxWin * S] reserved bases
  uint32_t* base_s = h_s + kMaxWin * a.S;
  __shared__ int s_w0, s_w1;
  const int64_t c0 = static_cast<int64_t>(blockIdx.x) * kChunk;
  const int64_t c1 = min(d.num_events, c0 + kChunk);
  if (threadIdx.x == 0) {
    s_w0 = first_window_of(d, c0);
    s_w1 = first_window_of(d, c1 - 1);
  }
  for (int i = threadIdx.x; i < kMaxWin * a.S; i += kSortThreads) h_s[i] = 0u;
  __syncthreads();
  const int w0 = s_w0;
  const bool local = s_w1 - w0 < kMaxWin;
  const int B = d.num_bins, H = d.H, W = d.W;
  const bool c16 = (d.xs_dtype == V2V_U16 || d.xs_dtype == V2V_I16) && (d.ys_dtype == V2V_U16 || d.ys_dtype == V2V_I16);
  const int64_t first = d.window_offsets[0], last = d.window_offsets[d.num_windows];
  constexpr int kPer = kChunk / kSortThreads;
  uint2 rec[kPer];
  int where[kPer];                                  // index of the event's (window, strip) counter, or -1
  uint32_t slot[kPer];
  long long ndrop = 0;
  // phase 1: every load of the thread's events is issued before any dependent arithmetic
  int wv[kPer];
  long long yv[kPer], xv[kPer];
  float pv[kPer];
  double tv[kPer];
  {
    int w = w0;
#pragma unroll
    for (int u = 0; u < kPer; ++u) {
      const int64_t e = c0 + u * kSortThreads + threadIdx.x;
      wv[u] = -1;
      yv[u] = xv[u] = -1;
      pv[u] = 0.f;
      tv[u] = 0.0;
      if (e >= c1 || e < first || e >= last) continue;
      while (e >= d.window_offsets[w + 1]) ++w;
      wv[u] = w;
      if (c16) {
        yv[u] = static_cast<const uint16_t*>(d.ys)[e];                      // negative int16 read as >= 32768: out of the sensor
        xv[u] = static_cast<const uint16_t*>(d.xs)[e];
      } else {
        bool ok = true;
        yv[u] = load_int(d.ys, d.ys_dtype, e, &ok);
        xv[u] = load_int(d.xs, d.xs_dtype, e, &ok);
        if (!ok) yv[u] = -1;
      }
      pv[u] = d.ps_dtype == V2V_U8 ? static_cast<float>(static_cast<const uint8_t*>(d.ps)[e]) : load_f32(d.ps, d.ps_dtype, e);
      tv[u] = d.ts_dtype == V2V_F64 ? static_cast<const double*>(d.ts)[e] : static_cast<double>(static_cast<const float*>(d.ts)[e]);
    }
  }
#pragma unroll
  for (int u = 0; u < kPer; ++u) {
    where[u] = -1;
    if (wv[u] < 0) continue;
    const int w = wv[u];
    const long long y = yv[u], x = xv[u];
    float pw;
    {
      const float p = pv[u];
      if (d.polarity_mode == V2V_POL_POS_ONLY) pw = p > 0.f ? 1.f : 0.f;
      else if (d.polarity_mode == V2V_POL_NEG_ONLY) pw = p <= 0.f ? 1.f : 0.f;
      else pw = 2.f * p - 1.f;                                              // testh5.py:67
    }
    const WinConst wc = a.wcs[w];
    // t_norm = tau / (tau_last + 1e-4) * (bins - 1) (:68,77) with the two constant factors folded into one per window
    // (h5_den holds (bins-1)/(tau_last+1e-4) here, t_first/h5_tpb the window's first timestamp): at most 2 ulp from the
    // reference's two roundings, far below the 2^-30 the weights are rounded to.  tau = trunc((ts - ts0)*1e6) stays in
    // floating point (no 64-bit integer round trip); float32 timestamps (EVAID) subtract and scale in float32 (:68).
    double tau;
    if (d.ts_dtype == V2V_F64) tau = trunc(__dmul_rn(__dsub_rn(tv[u], wc.h5_tpb), 1e6));
    else tau = static_cast<double>(truncf(__fmul_rn(__fsub_rn(static_cast<float>(tv[u]), wc.t_first), 1e6f)));
    const double tn = __dmul_rn(tau, wc.h5_den);
    const double fl = floor(tn);
    const bool in_sensor = y >= 0 && y < H && x >= 0 && x < W;
    const bool in_bins = fl >= -1.0 && fl < static_cast<double>(B);        // at least one tap inside [0, bins)
    if (!in_sensor || !in_bins) {
      ++ndrop;
      continue;
    }
    if (pw == 0.f) continue;                                               // one-polarity modes: weight 0 contributes nothing
    const uint32_t strip = static_cast<uint32_t>((static_cast<uint64_t>(y) * a.r_magic) >> 32);
    const uint32_t cell = static_cast<uint32_t>((y - static_cast<long long>(strip) * a.R) * W + x);
    const uint32_t frac = static_cast<uint32_t>(__double2int_rn(__dmul_rn(__dsub_rn(tn, fl), 1073741824.0)));   // in [0, 2^30]
    rec[u] = make_uint2(cell | (static_cast<uint32_t>(static_cast<int>(fl) + 1) << 16) | (pw < 0.f ? 0x80000000u : 0u), frac);
    if (local) {
      where[u] = (w - w0) * a.S + static_cast<int>(strip);
      slot[u] = atomicAdd(&h_s[where[u]], 1u);
    } else {
      a.records[atomicAdd(a.cursor + static_cast<int64_t>(w) * a.S + strip, 1u)] = rec[u];
    }
  }
  __syncthreads();
  if (local) {
    const int n = min(kMaxWin, d.num_windows - w0) * a.S;
    for (int i = threadIdx.x; i < n; i += kSortThreads)
      base_s[i] = h_s[i] ? atomicAdd(a.cursor + static_cast<int64_t>(w0) * a.S + i, h_s[i]) : 0u;
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kPer; ++u)
      if (where[u] >= 0) a.records[base_s[where[u]] + slot[u]] = rec[u];
  }
  if (d.dropped && ndrop) atomicAdd(reinterpret_cast<unsigned long long*>(d.dropped), static_cast<unsigned long long>(ndrop));
}

// exclusive prefix sum of counts[0..items) into cursor and starts (one CTA of 1024 threads, 8 consecutive items per thread)
__global__ void __launch_bounds__(1024) scan_counts_kernel(const SortedArgs a) {
  __shared__ uint32_t warp_tot[32];
  __shared__ uint32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kV = 8;
  for (int64_t base = 0; base < a.items; base += 1024 * kV) {
    const int64_t i0 = base + static_cast<int64_t>(threadIdx.x) * kV;
    uint32_t v[kV], tsum = 0u;
    if (i0 + kV <= a.items) {               // two 128-bit loads per thread: the warp reads 1 KB contiguously
      const uint4 q0 = reinterpret_cast<const uint4*>(a.counts + i0)[0], q1 = reinterpret_cast<const uint4*>(a.counts + i0)[1];
      v[0] = q0.x, v[1] = q0.y, v[2] = q0.z, v[3] = q0.w, v[4] = q1.x, v[5] = q1.y, v[6] = q1.z, v[7] = q1.w;
    } else {
#pragma unroll
      for (int k = 0; k < kV; ++k) v[k] = i0 + k < a.items ? a.counts[i0 + k] : 0u;
    }
#pragma unroll
    for (int k = 0; k < kV; ++k) tsum += v[k];
    uint32_t inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      uint32_t t = warp_tot[lane], s = t;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += u;
      }
      warp_tot[lane] = s - t;              // exclusive prefix of the warp totals
    }
    __syncthreads();
    uint32_t excl = carry_s + warp_tot[warp] + inc - tsum;
    uint32_t ex[kV];
#pragma unroll
    for (int k = 0; k < kV; ++k) {
      ex[k] = excl;
      excl += v[k];
    }
    if (i0 + kV <= a.items) {
      const uint4 q0 = make_uint4(ex[0], ex[1], ex[2], ex[3]), q1 = make_uint4(ex[4], ex[5], ex[6], ex[7]);
      reinterpret_cast<uint4*>(a.cursor + i0)[0] = q0, reinterpret_cast<uint4*>(a.cursor + i0)[1] = q1;
      reinterpret_cast<uint4*>(a.starts + i0)[0] = q0, reinterpret_cast<uint4*>(a.starts + i0)[1] = q1;
    } else {
#pragma unroll
      for (int k = 0; k < kV; ++k)
        if (i0 + k < a.items) a.cursor[i0 + k] = a.starts[i0 + k] = ex[k];
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = excl;
    __syncthreads();
  }
  if (threadIdx.x == 0) a.starts[a.items] = carry_s;
}

// work item = (window, strip): all bins of rows [r0, r0+rows) as fixed-point pairs in shared memory
__global__ void __launch_bounds__(kItemThreads, 8) scatter_sorted_kernel(const SortedArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const v2v_scatter_desc& d = a.d;
  const int W = d.W, H = d.H, B = d.num_bins, R = a.R;
  int* acc_hi = reinterpret_cast<int*>(smem_raw);
  for (int64_t item = blockIdx.x; item < a.items; item += gridDim.x) {
    const int win = static_cast<int>(item / a.S), strip = static_cast<int>(item - static_cast<int64_t>(win) * a.S);
    const int r0 = strip * R, rows = min(R, H - r0);
    const int cells = rows * W;                       // per bin
    const int plane = R * W;                          // tile layout: [bin][R*W] high words, then the same for low words
    unsigned int* acc_lo = reinterpret_cast<unsigned int*>(acc_hi + B * plane);
    const uint32_t s0 = a.starts[item], s1 = a.cursor[item];            // the used part of the segment (the count pass over-counts)
    const int64_t out_base = (static_cast<int64_t>(win) * B * H + r0) * W;      // + bin * H * W
    if (s1 == s0) {                                   // no event in these rows: zeros straight to HBM
      for (int b = 0; b < B; ++b) {
        const int64_t ob = out_base + static_cast<int64_t>(b) * H * W;
        if (d.out_dtype == V2V_F64) {
          double* o = static_cast<double*>(d.voxel) + ob;
          for (int i = threadIdx.x; i < cells; i += kItemThreads) o[i] = 0.0;
        } else {
          float* o = static_cast<float*>(d.voxel) + ob;
          for (int i = threadIdx.x; i < cells; i += kItemThreads) st_stream_f32(o + i, 0.f);
        }
      }
      continue;
    }
    // An item with at most 255 records cannot overflow ONE 32-bit word per cell at 2^-23 per unit weight (|sum| < 255 *
    // 2^23 < 2^31; error <= n_cell * 2^-24 per cell): two shared atomics per event instead of four and half the tile to
    // zero and read.  Larger items (hot rows, long windows) keep the exact two-word form.
    const bool one_word = s1 - s0 <= 255u;
    // ... and an item with more than 65535 records could overflow the two-word form (65536 same-sign unit weights on one
    // cell): it accumulates in ONE 64-bit word per cell instead, same footprint, exact for any count
    const bool wide = s1 - s0 > 65535u;
    unsigned long long* acc64 = reinterpret_cast<unsigned long long*>(smem_raw);
    {
      int4* z = reinterpret_cast<int4*>(smem_raw);
      const int n4 = ((one_word ? 1 : 2) * B * plane + 3) / 4;
      for (int i = threadIdx.x; i < n4; i += kItemThreads) z[i] = make_int4(0, 0, 0, 0);
    }
    __syncthreads();
    for (uint32_t r = s0 + threadIdx.x; r < s1; r += kItemThreads) {
      const uint2 rec = a.records[r];
      const int cell = static_cast<int>(rec.x & 0xffffu);
      const int b0 = static_cast<int>((rec.x >> 16) & 0xffu) - 1;          // floor(t_norm)
      const bool negp = (rec.x >> 31) != 0u;
      if (one_word) {
        const int f1 = static_cast<int>((rec.y + 64u) >> 7), f0 = 8388608 - f1;    // weights of bins b0+1 and b0 (:79), 2^-23 units
        if (b0 >= 0 && f0 != 0) atomicAdd(&acc_hi[b0 * plane + cell], negp ? -f0 : f0);
        if (b0 + 1 < B && f1 != 0) atomicAdd(&acc_hi[(b0 + 1) * plane + cell], negp ? -f1 : f1);
        continue;
      }
      const long long f1 = static_cast<long long>(rec.y), f0 = 1073741824ll - f1;   // 2^-30 units
      if (wide) {
        if (b0 >= 0 && f0 != 0) atomicAdd(&acc64[b0 * plane + cell], static_cast<unsigned long long>(negp ? -f0 : f0));
        if (b0 + 1 < B && f1 != 0) atomicAdd(&acc64[(b0 + 1) * plane + cell], static_cast<unsigned long long>(negp ? -f1 : f1));
        continue;
      }
      if (b0 >= 0 && f0 != 0) {
        const long long fx = negp ? -f0 : f0;
        const int hiw = static_cast<int>(fx >> kLoBits);
        const unsigned int low = static_cast<unsigned int>(fx & ((1 << kLoBits) - 1));
        if (hiw) atomicAdd(&acc_hi[b0 * plane + cell], hiw);
        if (low) atomicAdd(&acc_lo[b0 * plane + cell], low);
      }
      if (b0 + 1 < B && f1 != 0) {
        const long long fx = negp ? -f1 : f1;
        const int hiw = static_cast<int>(fx >> kLoBits);
        const unsigned int low = static_cast<unsigned int>(fx & ((1 << kLoBits) - 1));
        if (hiw) atomicAdd(&acc_hi[(b0 + 1) * plane + cell], hiw);
        if (low) atomicAdd(&acc_lo[(b0 + 1) * plane + cell], low);
      }
    }
    __syncthreads();
    for (int b = 0; b < B; ++b) {
      const int64_t ob = out_base + static_cast<int64_t>(b) * H * W;
      const int* h = acc_hi + b * plane;
      const unsigned int* l = acc_lo + b * plane;
      auto valuef = [&](int i) -> float {
        if (wide) return static_cast<float>(static_cast<double>(static_cast<long long>(acc64[b * plane + i])) * (1.0 / 1073741824.0));
        const int hv = h[i];
        if (one_word) return hv == 0 ? 0.f : __fmul_rn(static_cast<float>(hv), 1.0f / 8388608.0f);
        const unsigned int lv = l[i];
        if ((static_cast<unsigned int>(hv) | lv) == 0u) return 0.f;         // most cells hold no event
        // hv*2^-15 + lv*2^-30 is exact in float64 (46 significant bits at most): one rounding, to float32
        return static_cast<float>(__fma_rn(static_cast<double>(hv), 1.0 / 32768.0, __dmul_rn(static_cast<double>(lv), 1.0 / 1073741824.0)));
      };
      if (d.out_dtype == V2V_F64) {
        double* o = static_cast<double*>(d.voxel) + ob;
        for (int i = threadIdx.x; i < cells; i += kItemThreads) {
          if (wide) {
            o[i] = static_cast<double>(static_cast<long long>(acc64[b * plane + i])) * (1.0 / 1073741824.0);
          } else if (one_word) {
            o[i] = static_cast<double>(h[i]) * (1.0 / 8388608.0);
          } else {
            const long long tot = static_cast<long long>(h[i]) * (1 << kLoBits) + static_cast<long long>(l[i]);
            o[i] = static_cast<double>(tot) * (1.0 / 1073741824.0);
          }
        }
      } else {
        float* o = static_cast<float*>(d.voxel) + ob;
        const int head = min(static_cast<int>((4 - (ob & 3)) & 3), cells);
        for (int i = threadIdx.x; i < head; i += kItemThreads) st_stream_f32(o + i, valuef(i));
        const int n4 = (cells - head) / 4;
        for (int q = threadIdx.x; q < n4; q += kItemThreads) {
          const int i = head + 4 * q;
          st_stream_f32x4(o + i, valuef(i), valuef(i + 1), valuef(i + 2), valuef(i + 3));
        }
        for (int i = head + 4 * n4 + threadIdx.x; i < cells; i += kItemThreads) st_stream_f32(o + i, valuef(i));
      }
    }
    __syncthreads();
  }
}

}  // namespace

size_t scatter_sorted_workspace_bytes(const v2v_scatter_desc& d, int* rows_per_strip, int* strips) {
  if (d.H <= 0 || d.W <= 0) return 0;
  const int64_t per_row = static_cast<int64_t>(d.num_bins) * d.W * 8;
  int R = static_cast<int>(kTileBudget / per_row);
  if (R < 1) return 0;                                            // one row of all bins does not fit: not eligible
  if (static_cast<int64_t>(R) * d.W > 65535) R = 65535 / d.W;     // 16-bit cell index in the record
  if (R > d.H) R = d.H;
  if (R < 1) return 0;
  int S = (d.H + R - 1) / R;
  R = (d.H + S - 1) / S;                                          // balance the strips
  S = (d.H + R - 1) / R;
  if (S > 1024) return 0;                                         // the chunk-local counters of the sort live in shared memory
  *rows_per_strip = R;
  *strips = S;
  const int64_t items = static_cast<int64_t>(d.num_windows) * S;
  return (static_cast<size_t>(d.num_events) * 8 + 15) / 16 * 16 + 3 * ((static_cast<size_t>(items + 1) * 4 + 15) / 16 * 16) +
         static_cast<size_t>(d.num_windows) * sizeof(WinConst) + 64;
}

bool scatter_sorted_eligible(const v2v_scatter_desc& d) {
  int R, S;
  if (d.mode != V2V_SCATTER_H5_INTERP || d.polarity_mode == V2V_POL_SPLIT || d.num_bins > 254 || d.H > 65535 || d.num_events >= (1ll << 32)) return false;
  const size_t need = scatter_sorted_workspace_bytes(d, &R, &S);
  return need != 0 && d.workspace && static_cast<size_t>(d.workspace_bytes) >= need && aligned(d.workspace, 16);
}

int launch_scatter_sorted(const v2v_scatter_desc& d, cudaStream_t s) {
  SortedArgs a;
  a.d = d;
  scatter_sorted_workspace_bytes(d, &a.R, &a.S);
  a.r_magic = static_cast<uint32_t>(((1ull << 32) + a.R - 1) / a.R);          // exact for y < 65536 (y * (R-1) < 2^32)
  a.items = static_cast<int64_t>(d.num_windows) * a.S;
  char* p = static_cast<char*>(d.workspace);
  a.records = reinterpret_cast<uint2*>(p);
  p += (static_cast<size_t>(d.num_events) * 8 + 15) / 16 * 16;
  const size_t arr = (static_cast<size_t>(a.items + 1) * 4 + 15) / 16 * 16;     // 16-byte aligned arrays (128-bit loads in the scan)
  a.counts = reinterpret_cast<uint32_t*>(p);
  p += arr;
  a.cursor = reinterpret_cast<uint32_t*>(p);
  p += arr;
  a.starts = reinterpret_cast<uint32_t*>(p);
  p += arr;
  p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(p) + 15) & ~static_cast<uintptr_t>(15));
  a.wcs = reinterpret_cast<WinConst*>(p);
  V2V_CUDA(cudaMemsetAsync(a.counts, 0, static_cast<size_t>(a.items + 1) * 4, s));
  window_constants_kernel<<<(d.num_windows + 127) / 128, 128, 0, s>>>(a);
  const int chunks = static_cast<int>((d.num_events + kChunk - 1) / kChunk);
  const size_t hsm = static_cast<size_t>(kMaxWin) * a.S * sizeof(uint32_t);
  if (chunks > 0) count_events_kernel<<<chunks, kSortThreads, hsm, s>>>(a);
  scan_counts_kernel<<<1, 1024, 0, s>>>(a);
  if (chunks > 0) fill_events_kernel<<<chunks, kSortThreads, 2 * hsm, s>>>(a);
  const size_t smem = static_cast<size_t>(2) * d.num_bins * a.R * d.W * 4 + 16;
  static std::atomic<uint64_t> configured{0};
  int dev = 0, sms = 148;
  V2V_CUDA(cudaGetDevice(&dev));
  const uint64_t bit = 1ull << (dev & 63);
  if (!(configured.load(std::memory_order_acquire) & bit)) {
    V2V_CUDA(cudaFuncSetAttribute(scatter_sorted_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 30 * 1024));
    configured.fetch_or(bit, std::memory_order_release);
  }
  V2V_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = static_cast<int>(a.items < 32ll * sms ? a.items : 32ll * sms);
  scatter_sorted_kernel<<<grid, kItemThreads, smem, s>>>(a);
  count_launch(5);
  V2V_CUDA(cudaGetLastError());
  return V2V_OK;
}

}  // namespace v2v

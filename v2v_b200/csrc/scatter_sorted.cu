// Event stream -> voxel, h5 interpolated mode (TestH5Dataset.make_voxel with interpolate_bins, reference
// data/testh5.py:60-69,74-80), one visit per event and no ordering precondition.
//
// The contiguous-range kernel in scatter.cu visits an event once per (tap, strip) item whose bin range contains it —
// 16 times for 5 bins of 260 rows — because a bin's events are not ordered by row.  Here every event is read ONCE:
//   1. window constants (one thread per window: tau_last, the folded factor of :76-77, and — without a search — the
//      window of the first event of every chunk that starts inside this window);
//   2. sort    : one CTA per chunk of kChunk consecutive events builds an 8-byte record per kept event {cell in its strip,
//                floor(t_norm)+1, polarity sign, frac(t_norm) as 2^-30 fixed point}, counts per (window, strip) in shared
//                memory, turns the counters into offsets, stages the records in (window, strip) order and writes them back
//                over the chunk's own slice of the record array with coalesced stores, next to a 16-bit table of run
//                offsets (no global counters, no second traversal);
//   3. scatter : one work item per (window, strip) holds ALL bins of its rows in shared memory as fixed-point words, reads
//                only its own runs (one per chunk that overlaps the window; coalesced 8-byte loads), adds both temporal
//                taps of an event in the same visit (max(0, 1-|t_norm-b|) is 1-frac for b = floor and frac for
//                b = floor+1) and streams the finished planes out once with 128-bit stores.
// Accumulation is integer, so the result does not depend on the (atomic) record order: deterministic; exact to
// n * 2^-31 per cell like the contiguous-range kernel for items of more than 255 records, n * 2^-24 for the others
// (one 32-bit word per cell, half the shared atomics); items beyond 65535 records take one 64-bit word per cell.
// Timestamps need not be sorted: every event's bin comes from its own t_norm, as in the reference (events whose taps
// fall outside [0, bins) are dropped and counted).
// Extra HBM traffic: 8 bytes written and read per event, against 13 bytes of event read once and 4 bytes per voxel cell
// written once.
#include "scatter_common.cuh"

#include <type_traits>

namespace v2v {
namespace {

constexpr int kSortThreads = 512;
constexpr int kPer = 8;                        // events per thread of the sort pass (all loads issued before any arithmetic)
constexpr int kChunk = kSortThreads * kPer;    // events per CTA of the sort pass
constexpr int kMaxCounters = 2048;             // (window, strip) counters of one pass over a chunk (8 KB of shared memory)
constexpr int kItemThreads = 256;
constexpr int kTileBudget = 28 * 1024;    // eight CTAs per SM: an item is short (~200 records), its latency is hidden across CTAs
constexpr int kTileBudgetLarge = 56 * 1024;   // many bins (one row of 15 x 346 cells is 41 KB): four CTAs per SM

struct SortedArgs {
  v2v_scatter_desc d;
  int R, S;               // rows per strip, strips per window
  uint64_t r_magic;       // y / R == (y * r_magic) >> 32 for y < 65536 (2^32 for R = 1: 33 bits)
  int64_t items;          // Wn * S
  WinConst* wcs;          // [Wn]
  int32_t* chunk_w0;      // [chunks + 1]: window of the chunk's first event (written per window, no search); [chunks] = Wn - 1
  int32_t* win_w0f;       // [Wn]: chunk_w0 of the chunk that holds the window's first event
  int64_t chunks;
  uint16_t* tab;          // per chunk c, at (c + w0[c]) * S + c: exclusive offsets of its (window - w0, strip) runs, then the total
  uint2* records;         // [Ne]: chunk c's records, sorted by (window, strip), at [c * kChunk, c * kChunk + total)
};

__device__ __forceinline__ int first_window_of(const v2v_scatter_desc& d, int64_t e) {     // last window with start <= e
  int lo = 0, hi = d.num_windows;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (d.window_offsets[mid] <= e) lo = mid; else hi = mid;
  }
  return lo;
}

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
    c.h5_den = __ddiv_rn(static_cast<double>(a.d.num_bins - 1), c.h5_den);       // the sort pass multiplies
    if (a.d.ts_dtype == V2V_F64) c.h5_tpb = static_cast<const double*>(a.d.ts)[e0];   // first timestamp of the window
    else c.t_first = static_cast<const float*>(a.d.ts)[e0];
  }
  a.wcs[w] = c;
  // the chunks whose FIRST event lies in this window get their first window from here (no search in the sort pass);
  // window 0 also takes the chunks before the first offset, the last window everything after the last one
  const int64_t lo = w == 0 ? 0 : (e0 + kChunk - 1) / kChunk;
  const int64_t hi = w == a.d.num_windows - 1 ? a.chunks + 1 : min(a.chunks + 1, (e1 + kChunk - 1) / kChunk);
  for (int64_t ch = lo; ch < hi; ++ch) a.chunk_w0[ch] = w;
  a.win_w0f[w] = first_window_of(a.d, e0 / kChunk * kChunk);
}

__device__ __forceinline__ int64_t table_base(const SortedArgs& a, int64_t chunk, int w0) {
  // chunk c owns (w1 - w0 + 1) * S + 1 entries; consecutive chunks share at most one window, so these bases never overlap
  return (chunk + w0) * a.S + chunk;
}

// In-place exclusive prefix sum of h[0..n) by the whole CTA (n <= kMaxCounters); returns the total.
__device__ __forceinline__ uint32_t cta_exclusive_scan(uint32_t* h, int n, uint32_t* warp_tot) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per = (n + kSortThreads - 1) / kSortThreads;
  const int i0 = threadIdx.x * per;
  uint32_t tsum = 0u;
  for (int k = 0; k < per; ++k)
    if (i0 + k < n) tsum += h[i0 + k];
  uint32_t inc = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const uint32_t t = lane < kSortThreads / 32 ? warp_tot[lane] : 0u;
    uint32_t sc = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xffffffffu, sc, o);
      if (lane >= o) sc += u;
    }
    warp_tot[lane] = sc - t;                       // exclusive prefix of the warp totals
    if (lane == 31) warp_tot[32] = sc;             // the total
  }
  __syncthreads();
  uint32_t excl = warp_tot[warp] + inc - tsum;
  for (int k = 0; k < per; ++k)
    if (i0 + k < n) {
      const uint32_t v = h[i0 + k];
      h[i0 + k] = excl;
      excl += v;
    }
  const uint32_t total = warp_tot[32];
  __syncthreads();
  return total;
}

// Sort pass: ONE read of the events.  A CTA takes kChunk consecutive events, builds the 8-byte record of every event it
// keeps, counts them per (window, strip) in shared memory (the slot of a record is the value its shared atomic returns),
// turns the counters into offsets, stages the records in (window, strip) order in shared memory and writes them back over
// the chunk's own slice of the record array with coalesced stores, next to a small table of the run offsets.  No global
// counters, no second traversal: what used to be count + scan + fill.  A chunk whose windows x strips exceed the shared
// counters (thousands of tiny or empty windows) is done in several rounds over groups of windows.
__global__ void __launch_bounds__(kSortThreads, 2) sort_chunks_kernel(const SortedArgs a) {
  const v2v_scatter_desc& d = a.d;
  __shared__ uint32_t h_s[kMaxCounters];
  __shared__ uint2 stage[kChunk];
  __shared__ uint32_t warp_tot[33];
  const int64_t c0 = static_cast<int64_t>(blockIdx.x) * kChunk;
  const int64_t c1 = min(d.num_events, c0 + kChunk);
  // the chunk's windows: from the window of its first event to the window of the next chunk's first event (which may hold
  // none of this chunk's events: its runs are then empty; the table layout counts on exactly this span)
  const int w0 = a.chunk_w0[blockIdx.x], w1 = a.chunk_w0[blockIdx.x + 1];
  const int B = d.num_bins, H = d.H, W = d.W, S = a.S;
  const bool c16 = (d.xs_dtype == V2V_U16 || d.xs_dtype == V2V_I16) && (d.ys_dtype == V2V_U16 || d.ys_dtype == V2V_I16);
  const int64_t first = d.window_offsets[0], last = d.window_offsets[d.num_windows];
  uint2 rec[kPer];
  int where[kPer];                                  // (window - w0) * S + strip of a kept event, or -1
  long long ndrop = 0;
  // phase 1: every load of the thread's events is issued before anything depends on one (the window search comes after).
  // h5 streams (16-bit coordinates, uint8 polarity, float64 timestamps, 16-byte aligned, full chunk): a thread takes kPer
  // CONSECUTIVE events with seven 128/64-bit loads (a warp request covers 512 B - 2 KB) instead of kPer strided ones with
  // 32 small loads; the order of the records inside a run does not matter.
  int wv[kPer];
  long long yv[kPer], xv[kPer];
  float pv[kPer];
  double tv[kPer];
  static_assert(kPer == 8, "the vector form below loads eight events per thread");
  auto al = [](const void* q, uintptr_t n) { return (reinterpret_cast<uintptr_t>(q) & (n - 1)) == 0; };
  const bool vec = c16 && d.ps_dtype == V2V_U8 && d.ts_dtype == V2V_F64 && c1 - c0 == kChunk && al(d.ys, 16) && al(d.xs, 16) &&
                   al(d.ps, 8) && al(d.ts, 16);
  auto event_of = [&](int u) -> int64_t { return vec ? c0 + static_cast<int64_t>(threadIdx.x) * kPer + u : c0 + u * kSortThreads + threadIdx.x; };
  if (vec) {
    const int64_t e = c0 + static_cast<int64_t>(threadIdx.x) * kPer;
    const uint4 y8 = *reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(d.ys) + e);
    const uint4 x8 = *reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(d.xs) + e);
    const uint2 p8 = *reinterpret_cast<const uint2*>(static_cast<const uint8_t*>(d.ps) + e);
    const double2* t2 = reinterpret_cast<const double2*>(static_cast<const double*>(d.ts) + e);
    const double2 ta = t2[0], tb = t2[1], tc = t2[2], td = t2[3];
    const uint32_t yw[4] = {y8.x, y8.y, y8.z, y8.w}, xw[4] = {x8.x, x8.y, x8.z, x8.w}, pw2[2] = {p8.x, p8.y};
    const double tt[8] = {ta.x, ta.y, tb.x, tb.y, tc.x, tc.y, td.x, td.y};
#pragma unroll
    for (int u = 0; u < kPer; ++u) {
      yv[u] = (yw[u >> 1] >> ((u & 1) * 16)) & 0xffffu;
      xv[u] = (xw[u >> 1] >> ((u & 1) * 16)) & 0xffffu;
      pv[u] = static_cast<float>((pw2[u >> 2] >> ((u & 3) * 8)) & 0xffu);
      tv[u] = tt[u];
    }
  } else {
#pragma unroll
  for (int u = 0; u < kPer; ++u) {
    const int64_t e = event_of(u);
    yv[u] = xv[u] = -1;
    pv[u] = 0.f;
    tv[u] = 0.0;
    if (e >= c1) continue;
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
  {
    int w = w0;
#pragma unroll
    for (int u = 0; u < kPer; ++u) {
      const int64_t e = event_of(u);
      wv[u] = -1;
      if (e >= c1 || e < first || e >= last) continue;
      if (w0 != w1)                                   // (most chunks lie inside one window)
        while (e >= d.window_offsets[w + 1]) ++w;
      wv[u] = w;
    }
  }
  const WinConst wc0 = a.wcs[w0];
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
    WinConst wc = wc0;
    if (w != w0) wc = a.wcs[w];
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
    const uint32_t strip = static_cast<uint32_t>((static_cast<uint64_t>(y) * a.r_magic) >> 32);     // y / R
    const uint32_t cell = static_cast<uint32_t>((y - static_cast<long long>(strip) * a.R) * W + x);
    const uint32_t frac = static_cast<uint32_t>(__double2int_rn(__dmul_rn(__dsub_rn(tn, fl), 1073741824.0)));   // in [0, 2^30]
    rec[u] = make_uint2(cell | (static_cast<uint32_t>(static_cast<int>(fl) + 1) << 16) | (pw < 0.f ? 0x80000000u : 0u), frac);
    where[u] = (w - w0) * S + static_cast<int>(strip);
  }
  // phase 2: rounds over groups of windows whose counters fit (one round for every realistic stream)
  const int G = max(1, kMaxCounters / S);
  const int64_t tb = table_base(a, blockIdx.x, w0);
  uint32_t running = 0u;
  for (int wg = w0; wg <= w1; wg += G) {
    const int n = min(G, w1 - wg + 1) * S;
    const int lo = (wg - w0) * S;
    for (int i = threadIdx.x; i < n; i += kSortThreads) h_s[i] = 0u;
    __syncthreads();
    uint32_t slot[kPer];
#pragma unroll
    for (int u = 0; u < kPer; ++u)
      if (where[u] >= lo && where[u] < lo + n) slot[u] = atomicAdd(&h_s[where[u] - lo], 1u);
    __syncthreads();
    const uint32_t total = cta_exclusive_scan(h_s, n, warp_tot);
    for (int i = threadIdx.x; i < n; i += kSortThreads) a.tab[tb + lo + i] = static_cast<uint16_t>(running + h_s[i]);
#pragma unroll
    for (int u = 0; u < kPer; ++u)
      if (where[u] >= lo && where[u] < lo + n) stage[running + h_s[where[u] - lo] + slot[u]] = rec[u];
    running += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) a.tab[tb + static_cast<int64_t>(w1 - w0 + 1) * S] = static_cast<uint16_t>(running);
  for (uint32_t i = threadIdx.x; i < running; i += kSortThreads) a.records[c0 + i] = stage[i];
  if (d.dropped && ndrop) atomicAdd(reinterpret_cast<unsigned long long*>(d.dropped), static_cast<unsigned long long>(ndrop));
}

// Scatter pass.  A task is (window, group of kGroup consecutive strips); an item is one strip of it: all bins of rows
// [r0, r0+rows) as fixed-point words in shared memory.
//  * The window's run offsets for the whole group are fetched once per task: lane j of EVERY warp holds chunk j's kGroup+1
//    table entries in registers (shifted down by one per item, so the current item is always entries 0 and 1): no shared
//    staging, no barrier, and the run lookups of an item are register shuffles.
//  * ONE barrier per item.  Items of at most 255 records (the common case) need one 32-bit word per cell, so the tile holds
//    two accumulators used alternately; the write-out of an item clears each cell as it reads it.  A thread that has
//    written its share of item k goes straight on to add the records of item k+1 into the other accumulator; the only
//    barrier sits between the adds and the write-out of the same item, and it also orders "clear of item k-1" before
//    "adds of item k+1" on the same accumulator.  Items that need the whole tile (two words or one 64-bit word per cell)
//    put a barrier on either side.
constexpr int kGroup = 8;

__global__ void __launch_bounds__(kItemThreads, 8) scatter_sorted_kernel(const SortedArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const v2v_scatter_desc& d = a.d;
  const int W = d.W, H = d.H, B = d.num_bins, R = a.R, S = a.S;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int plane = (R * W + 3) / 4 * 4 + 4;          // words of one bin of the tile: the strip's cells + room for the phase shift
  const int half = B * plane;                         // words of one accumulator: tile = [2][bin][plane] 32-bit words
  int* tile = reinterpret_cast<int*>(smem_raw);
  const int64_t HW = static_cast<int64_t>(H) * W;
  const bool flat = d.out_dtype != V2V_F64 && (HW & 3) == 0 && 8 * B <= kItemThreads;
  {
    int4* z = reinterpret_cast<int4*>(smem_raw);
    for (int i = threadIdx.x; i < (2 * half + 3) / 4; i += kItemThreads) z[i] = make_int4(0, 0, 0, 0);
  }
  __syncthreads();
  int pp = 0;                                         // accumulator of the next one-word item
  const int groups = (S + kGroup - 1) / kGroup;
  const int64_t tasks = static_cast<int64_t>(d.num_windows) * groups;
  for (int64_t task = blockIdx.x; task < tasks; task += gridDim.x) {
    const int win = static_cast<int>(task / groups), s0 = static_cast<int>(task - static_cast<int64_t>(win) * groups) * kGroup;
    const int ns = min(kGroup, S - s0);
    const int64_t e0 = d.window_offsets[win], e1 = d.window_offsets[win + 1];
    const int64_t cf = e0 / kChunk;
    const int nch = e1 > e0 ? static_cast<int>((e1 - 1) / kChunk - cf + 1) : 0;   // chunks that overlap the window
    const int w0f = nch ? a.win_w0f[win] : 0;
    // table row of the item group in chunk cf + j (every later chunk than the first starts inside this window)
    auto row_of = [&](int j) -> const uint16_t* {
      const int64_t c = cf + j;
      const int w0c = j == 0 ? w0f : win;
      return a.tab + table_base(a, c, w0c) + static_cast<int64_t>(win - w0c) * S + s0;
    };
    uint32_t t[kGroup + 1];                           // lane j: run offsets of chunk cf + j for strips s0 .. s0 + ns
#pragma unroll
    for (int k = 0; k <= kGroup; ++k) t[k] = 0u;
    if (lane < nch) {
      const uint16_t* row = row_of(lane);
#pragma unroll
      for (int k = 0; k <= kGroup; ++k) t[k] = row[min(k, ns)];
    }
    const int64_t my_chunk0 = (cf + lane) * kChunk;
#pragma unroll 1
    for (int k = 0; k < ns; ++k) {
      const int strip = s0 + k;
      const int r0 = strip * R, rows = min(R, H - r0);
      const int cells = rows * W;                     // per bin
      const int64_t out_base = (static_cast<int64_t>(win) * B * H + r0) * W;      // + bin * H * W
      // float32 planes are written with 16-byte stores: the strip sits in the tile at the phase of its global address, so the
      // same four cells are one aligned 128-bit word on both sides
      const int shift = flat ? static_cast<int>(out_base & 3) : 0;
      const uint32_t my_beg = t[0], my_len = t[1] - t[0];
#pragma unroll
      for (int q = 0; q < kGroup; ++q) t[q] = t[q + 1];                         // the next item's entries move to 0 and 1
      uint32_t tot = my_len;
      for (int j = 32 + lane; j < nch; j += 32) {     // (windows of more than 32 chunks)
        const uint16_t* row = row_of(j);
        tot += static_cast<uint32_t>(row[k + 1]) - row[k];
      }
      const uint32_t nrec = __reduce_add_sync(0xffffffffu, tot);
      if (nrec == 0u) {                               // no event in these rows: zeros straight to HBM
        for (int b = 0; b < B; ++b) {
          const int64_t ob = out_base + static_cast<int64_t>(b) * HW;
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
      // 2^23 < 2^31; error <= n_cell * 2^-24 per cell).  Larger items (hot rows, long windows) take the exact two-word form
      // (2^-30 units split 15 + 15 bits), and an item with more than 65535 records, which could overflow that (65536
      // same-sign unit weights on one cell), ONE 64-bit word per cell: same footprint, exact for any count.
      const int form = nrec <= 255u ? 0 : (nrec <= 65535u ? 1 : 2);
      int* acc_hi = tile + (form == 0 ? pp * half : 0);
      unsigned int* acc_lo = reinterpret_cast<unsigned int*>(tile + half);
      unsigned long long* acc64 = reinterpret_cast<unsigned long long*>(smem_raw);
      if (form != 0) __syncthreads();                 // the whole tile: every earlier write-out (and its clearing) is done
      auto add_record = [&](const uint2 rec) {
        const int cell = static_cast<int>(rec.x & 0xffffu) + shift;
        const int b0 = static_cast<int>((rec.x >> 16) & 0xffu) - 1;          // floor(t_norm)
        const bool negp = (rec.x >> 31) != 0u;
        if (form == 0) {
          const int f1 = static_cast<int>((rec.y + 64u) >> 7), f0 = 8388608 - f1;    // weights of bins b0+1 and b0 (:79), 2^-23 units
          if (b0 >= 0 && f0 != 0) atomicAdd(&acc_hi[b0 * plane + cell], negp ? -f0 : f0);
          if (b0 + 1 < B && f1 != 0) atomicAdd(&acc_hi[(b0 + 1) * plane + cell], negp ? -f1 : f1);
          return;
        }
        const long long f1 = static_cast<long long>(rec.y), f0 = 1073741824ll - f1;   // 2^-30 units
        if (form == 2) {
          if (b0 >= 0 && f0 != 0) atomicAdd(&acc64[b0 * plane + cell], static_cast<unsigned long long>(negp ? -f0 : f0));
          if (b0 + 1 < B && f1 != 0) atomicAdd(&acc64[(b0 + 1) * plane + cell], static_cast<unsigned long long>(negp ? -f1 : f1));
          return;
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
      };
      // one warp per run (the item's records of one chunk): coalesced 8-byte loads; the run comes from lane j's registers
      for (int j = warp; j < min(nch, 32); j += kItemThreads / 32) {
        const int64_t beg = __shfl_sync(0xffffffffu, my_chunk0 + my_beg, j);
        const uint32_t len = __shfl_sync(0xffffffffu, my_len, j);
        for (uint32_t r = lane; r < len; r += 32) add_record(a.records[beg + r]);
      }
      for (int j = 32 + warp; j < nch; j += kItemThreads / 32) {           // (windows of more than 32 chunks)
        const uint16_t* row = row_of(j);
        const int64_t beg = (cf + j) * kChunk + row[k];
        const uint32_t len = static_cast<uint32_t>(row[k + 1]) - row[k];
        for (uint32_t r = lane; r < len; r += 32) add_record(a.records[beg + r]);
      }
      __syncthreads();
      // Write-out: one specialised, branch-free loop per accumulator form (a per-cell "is it zero" test compiles to a
      // divergent branch per cell and was 60 % of this kernel's instructions); every cell is cleared as it is read.
      auto write_out = [&](auto form_tag) {
        constexpr int FORM = decltype(form_tag)::value;                        // 0 one word, 1 two words, 2 one 64-bit word
        auto take = [&](int idx) -> long long {                               // cell idx = bin * plane + i: its sum, cleared
          if (FORM == 0) {
            const int v = acc_hi[idx];
            acc_hi[idx] = 0;
            return v;
          }
          if (FORM == 2) {
            const long long v = static_cast<long long>(acc64[idx]);
            acc64[idx] = 0ull;
            return v;
          }
          const long long v = static_cast<long long>(acc_hi[idx]) * (1 << kLoBits) + static_cast<long long>(acc_lo[idx]);
          acc_hi[idx] = 0;
          acc_lo[idx] = 0u;
          return v;
        };
        // integer -> float32 is ONE correctly rounded conversion, the power-of-two scale is exact
        auto takef = [&](int idx) -> float {
          if (FORM == 0) return __fmul_rn(static_cast<float>(static_cast<int>(take(idx))), 1.0f / 8388608.0f);
          return __fmul_rn(__ll2float_rn(take(idx)), 1.0f / 1073741824.0f);
        };
        if (flat) {
          // float32 planes whose size is a multiple of 4 cells: every bin's slice of the strip has the same 16-byte phase,
          // so the B slices are written by ONE loop over (bin, group of 4 cells); (bin, q) advance incrementally
          float* o0 = static_cast<float*>(d.voxel) + out_base;
          const int head = min(static_cast<int>((4 - (out_base & 3)) & 3), cells);
          const int n4 = (cells - head) / 4;
          if (n4 > 0) {
            int b = 0, q = threadIdx.x;
            while (q >= n4 && b < B) q -= n4, ++b;
            while (b < B) {
              const int i = head + 4 * q, idx = b * plane + shift + i;         // idx is a multiple of 4
              float v0, v1, v2, v3;
              if (FORM == 0) {
                int4* c = reinterpret_cast<int4*>(acc_hi + idx);
                const int4 h = *c;
                *c = make_int4(0, 0, 0, 0);
                constexpr float sc = 1.0f / 8388608.0f;
                v0 = __fmul_rn(static_cast<float>(h.x), sc), v1 = __fmul_rn(static_cast<float>(h.y), sc);
                v2 = __fmul_rn(static_cast<float>(h.z), sc), v3 = __fmul_rn(static_cast<float>(h.w), sc);
              } else if (FORM == 1) {
                int4* ch = reinterpret_cast<int4*>(acc_hi + idx);
                uint4* cl = reinterpret_cast<uint4*>(acc_lo + idx);
                const int4 h = *ch;
                const uint4 l = *cl;
                *ch = make_int4(0, 0, 0, 0);
                *cl = make_uint4(0u, 0u, 0u, 0u);
                constexpr float sc = 1.0f / 1073741824.0f;
                auto f = [&](int hv, unsigned int lv) { return __fmul_rn(__ll2float_rn(static_cast<long long>(hv) * (1 << kLoBits) + static_cast<long long>(lv)), sc); };
                v0 = f(h.x, l.x), v1 = f(h.y, l.y), v2 = f(h.z, l.z), v3 = f(h.w, l.w);
              } else {
                v0 = takef(idx), v1 = takef(idx + 1), v2 = takef(idx + 2), v3 = takef(idx + 3);
              }
              st_stream_f32x4(o0 + b * HW + i, v0, v1, v2, v3);
              q += kItemThreads;
              while (q >= n4 && b < B) q -= n4, ++b;
            }
          }
          const int tail0 = head + 4 * n4;
          if (static_cast<int>(threadIdx.x) < 8 * B) {                       // up to three cells before and after the aligned part
            const int b = threadIdx.x >> 3, j = threadIdx.x & 7;
            const int i = j < 4 ? j : tail0 + j - 4;
            if (j < 4 ? j < head : i < cells) st_stream_f32(o0 + b * HW + i, takef(b * plane + shift + i));
          }
          return;
        }
        for (int b = 0; b < B; ++b) {
          const int64_t ob = out_base + static_cast<int64_t>(b) * HW;
          if (d.out_dtype == V2V_F64) {
            double* o = static_cast<double*>(d.voxel) + ob;
            for (int i = threadIdx.x; i < cells; i += kItemThreads)
              o[i] = static_cast<double>(take(b * plane + i)) * (FORM == 0 ? 1.0 / 8388608.0 : 1.0 / 1073741824.0);
          } else {
            float* o = static_cast<float*>(d.voxel) + ob;
            for (int i = threadIdx.x; i < cells; i += kItemThreads) st_stream_f32(o + i, takef(b * plane + i));
          }
        }
      };
      if (form == 2) write_out(std::integral_constant<int, 2>{});
      else if (form == 0) write_out(std::integral_constant<int, 0>{});
      else write_out(std::integral_constant<int, 1>{});
      if (form != 0) __syncthreads();                 // (the next item may add into either half)
      else pp ^= 1;
    }
  }
}

}  // namespace

static size_t align16(size_t n) { return (n + 15) / 16 * 16; }

size_t scatter_sorted_workspace_bytes(const v2v_scatter_desc& d, int* rows_per_strip, int* strips) {
  if (d.H <= 0 || d.W <= 0) return 0;
  const int64_t per_row = static_cast<int64_t>(d.num_bins) * d.W * 8;
  const int64_t pad = static_cast<int64_t>(d.num_bins) * 7 * 8;                                 // (+ up to 7 padding words per bin)
  int R = static_cast<int>((kTileBudget - pad) / per_row);
  if (R < 1) R = static_cast<int>((kTileBudgetLarge - pad) / per_row);
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
  const size_t chunks = static_cast<size_t>((d.num_events + kChunk - 1) / kChunk);
  return align16(static_cast<size_t>(d.num_events) * 8) + align16((chunks + 1) * 4) + align16(static_cast<size_t>(d.num_windows) * 4) +
         align16(((chunks + static_cast<size_t>(d.num_windows)) * S + chunks + 1) * 2) +
         static_cast<size_t>(d.num_windows) * sizeof(WinConst) + 64;
}

bool scatter_sorted_eligible(const v2v_scatter_desc& d) {
  int R, S;
  if (d.mode != V2V_SCATTER_H5_INTERP || d.polarity_mode == V2V_POL_SPLIT || d.num_bins > 254 || d.H > 65535 || d.num_events >= (1ll << 32)) return false;
  const size_t need = scatter_sorted_workspace_bytes(d, &R, &S);
  if (need == 0 || !d.workspace || static_cast<size_t>(d.workspace_bytes) < need || !aligned(d.workspace, 16)) return false;
  // a work item is one CTA's job: a few huge windows (the offline cache builder) are better served by the contiguous-range
  // kernel, which slices a window over many CTAs
  return d.num_events <= static_cast<int64_t>(d.num_windows) * S * 32768;
}

int launch_scatter_sorted(const v2v_scatter_desc& d, cudaStream_t s) {
  SortedArgs a;
  a.d = d;
  scatter_sorted_workspace_bytes(d, &a.R, &a.S);
  a.r_magic = ((1ull << 32) + a.R - 1) / a.R;                                 // exact for y < 65536 (y * (R-1) < 2^32)
  a.items = static_cast<int64_t>(d.num_windows) * a.S;
  const size_t chunks = static_cast<size_t>((d.num_events + kChunk - 1) / kChunk);
  char* p = static_cast<char*>(d.workspace);
  a.records = reinterpret_cast<uint2*>(p);
  p += align16(static_cast<size_t>(d.num_events) * 8);
  a.chunk_w0 = reinterpret_cast<int32_t*>(p);
  p += align16((chunks + 1) * 4);
  a.win_w0f = reinterpret_cast<int32_t*>(p);
  p += align16(static_cast<size_t>(d.num_windows) * 4);
  a.chunks = static_cast<int64_t>(chunks);
  a.tab = reinterpret_cast<uint16_t*>(p);
  p += align16(((chunks + static_cast<size_t>(d.num_windows)) * a.S + chunks + 1) * 2);
  a.wcs = reinterpret_cast<WinConst*>(p);
  window_constants_kernel<<<(d.num_windows + 127) / 128, 128, 0, s>>>(a);
  if (chunks > 0) sort_chunks_kernel<<<static_cast<unsigned int>(chunks), kSortThreads, 0, s>>>(a);
  const size_t smem = static_cast<size_t>(2) * d.num_bins * ((a.R * d.W + 3) / 4 * 4 + 4) * 4 + 16;
  static std::atomic<uint64_t> configured{0};
  int dev = 0, sms = 148;
  V2V_CUDA(cudaGetDevice(&dev));
  const uint64_t bit = 1ull << (dev & 63);
  if (!(configured.load(std::memory_order_acquire) & bit)) {
    V2V_CUDA(cudaFuncSetAttribute(scatter_sorted_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileBudgetLarge + 1024));
    configured.fetch_or(bit, std::memory_order_release);
  }
  V2V_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t tasks = static_cast<int64_t>(d.num_windows) * ((a.S + kGroup - 1) / kGroup);
  const int grid = static_cast<int>(tasks < 32ll * sms ? tasks : 32ll * sms);
  scatter_sorted_kernel<<<grid, kItemThreads, smem, s>>>(a);
  count_launch(3);
  V2V_CUDA(cudaGetLastError());
  return V2V_OK;
}

}  // namespace v2v

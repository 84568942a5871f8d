// Event stream -> voxel / image scatter (sm_100a).
//
// Replaces TestH5Dataset.make_voxel (reference data/testh5.py:60-90),
// events_to_voxel_torch / events_to_neg_pos_voxel_torch
// (utils/event_utils.py:466-541) and events_to_image_torch (:330-376).
//
// Voxel kernel: shared-memory privatised, write-once.  A work item is
// (window, strip of rows): the CTA zeroes a [bins, rows, W] tile of accumulators
// in shared memory, scans the window's events (coalesced reads of ys; the other
// three streams are only touched for events that fall in the strip), adds with
// shared-memory integer atomics, then streams the finished strip to HBM once.
// The output is never zero-filled or read back: HBM traffic is the event
// stream (first strip; the others hit L2) plus 4 B per voxel cell.
//
// Accumulation is exact and order independent where the reference's is not:
//   h5 discrete  : int32 counts (reference: float64 adds of +-1, exact too);
//   h5 interp    : fixed point, round(w*2^30) split in two int32 words
//                  (reference: sequential float64) -> |err| <= n*2^-31 per cell;
//   torch modes  : float32 shared atomics (reference: sequential float32).
#include "common.cuh"

namespace v2v {
namespace {

constexpr int kScatterThreads = 512;
constexpr int kSmemBudget = 200 * 1024;   // one CTA per SM
constexpr int kFixShift = 30, kLoBits = 15;

struct ScatterArgs {
  v2v_scatter_desc d;
  int rows_per_strip, num_strips;
};

__device__ __forceinline__ long long load_int(const void* p, int dtype, int64_t i, bool* ok) {
  switch (dtype) {
    case V2V_U8: return static_cast<const uint8_t*>(p)[i];
    case V2V_I8: return static_cast<const int8_t*>(p)[i];
    case V2V_U16: return static_cast<const uint16_t*>(p)[i];
    case V2V_I16: return static_cast<const int16_t*>(p)[i];
    case V2V_I32: return static_cast<const int32_t*>(p)[i];
    case V2V_I64: return static_cast<const int64_t*>(p)[i];
    case V2V_F32: {   // .long() / .to(int): truncation toward zero
      float f = static_cast<const float*>(p)[i];
      if (!(fabsf(f) < 1.0e9f)) { *ok = false; return 0; }
      return static_cast<long long>(f);
    }
    case V2V_F64: {
      double f = static_cast<const double*>(p)[i];
      if (!(fabs(f) < 1.0e9)) { *ok = false; return 0; }
      return static_cast<long long>(f);
    }
  }
  *ok = false;
  return 0;
}

__device__ __forceinline__ float load_f32(const void* p, int dtype, int64_t i) {
  switch (dtype) {
    case V2V_U8: return static_cast<float>(static_cast<const uint8_t*>(p)[i]);
    case V2V_I8: return static_cast<float>(static_cast<const int8_t*>(p)[i]);
    case V2V_F32: return static_cast<const float*>(p)[i];
    case V2V_F64: return static_cast<float>(static_cast<const double*>(p)[i]);
    case V2V_I32: return static_cast<float>(static_cast<const int32_t*>(p)[i]);
    case V2V_I64: return static_cast<float>(static_cast<const int64_t*>(p)[i]);
    case V2V_U16: return static_cast<float>(static_cast<const uint16_t*>(p)[i]);
    case V2V_I16: return static_cast<float>(static_cast<const int16_t*>(p)[i]);
  }
  return 0.f;
}

// µs since the window start, exactly as ((ts - ts[0]) * 1e6).astype(int64)
// evaluates in the dtype of the stored timestamps (data/testh5.py:68).
__device__ __forceinline__ long long tau_us(const void* ts, int dtype, int64_t i, int64_t i0) {
  if (dtype == V2V_F64) {
    const double* t = static_cast<const double*>(ts);
    return static_cast<long long>(__dmul_rn(__dsub_rn(t[i], t[i0]), 1e6));
  }
  const float* t = static_cast<const float*>(ts);
  return static_cast<long long>(__fmul_rn(__fsub_rn(t[i], t[i0]), 1e6f));
}

template <int MODE>
__global__ void __launch_bounds__(kScatterThreads, 1) scatter_kernel(const ScatterArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const v2v_scatter_desc& d = a.d;
  const int W = d.W, H = d.H, B = d.num_bins;
  const int R = a.rows_per_strip;
  constexpr bool kInterp = MODE == V2V_SCATTER_H5_INTERP;
  constexpr bool kTorch = MODE == V2V_SCATTER_TORCH_DISCRETE || MODE == V2V_SCATTER_TORCH_BILINEAR;
  constexpr bool kH5 = !kTorch;

  const int64_t items = static_cast<int64_t>(d.num_windows) * a.num_strips;
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int win = static_cast<int>(item / a.num_strips);
    const int strip = static_cast<int>(item - static_cast<int64_t>(win) * a.num_strips);
    const int r0 = strip * R;
    const int rows = min(R, H - r0);
    const int cells = B * rows * W;             // tile cells; plane stride inside the tile = rows*W
    int* acc_i = reinterpret_cast<int*>(smem_raw);
    float* acc_f = reinterpret_cast<float*>(smem_raw);
    int* acc_lo = acc_i + B * R * W;            // second word of the fixed-point pair (interp only)

    // zero the tile
    {
      const int words = kInterp ? 2 * B * R * W : cells;
      for (int i = threadIdx.x; i < words; i += kScatterThreads) acc_i[i] = 0;
    }
    __syncthreads();

    const int64_t e0 = d.window_offsets[win], e1 = d.window_offsets[win + 1];
    if (e1 > e0) {
      // window-level constants
      double h5_tpb = 0.0, h5_den = 0.0;
      float t_first = 0.f, t_span = 0.f, t_tpb = 0.f;
      if (kH5) {
        const long long tl = tau_us(d.ts, d.ts_dtype, e1 - 1, e0);
        h5_tpb = __ddiv_rn(__dadd_rn(static_cast<double>(tl), 0.001), static_cast<double>(B));       // :71
        h5_den = __dadd_rn(static_cast<double>(tl), 0.0001);                                         // :76-77 (ts[0]==0)
      } else {
        t_first = load_f32(d.ts, d.ts_dtype, e0);
        t_span = __fsub_rn(load_f32(d.ts, d.ts_dtype, e1 - 1), t_first);                             // event_utils.py:489
        t_tpb = __fdiv_rn(__fadd_rn(t_span, 0.001f), static_cast<float>(B));                         // :503
      }
      long long ndrop = 0;
      for (int64_t e = e0 + threadIdx.x; e < e1; e += kScatterThreads) {
        bool ok = true;
        const long long y = load_int(d.ys, d.ys_dtype, e, &ok);
        const long long ry = y - r0;
        const bool in_strip = ok && ry >= 0 && ry < rows;
        if (!in_strip) {
          if (strip == 0 && (!ok || y < 0 || y >= H)) ++ndrop;      // counted once per event
          continue;
        }
        const long long x = load_int(d.xs, d.xs_dtype, e, &ok);
        if (!ok || x < 0 || x >= W) { ++ndrop; continue; }
        const int cell = static_cast<int>(ry) * W + static_cast<int>(x);
        const int plane = rows * W;

        // polarity -> weight
        float pw;
        {
          const float p = load_f32(d.ps, d.ps_dtype, e);
          if (d.polarity_mode == V2V_POL_POS_ONLY) pw = p > 0.f ? 1.f : 0.f;          // event_utils.py:533
          else if (d.polarity_mode == V2V_POL_NEG_ONLY) pw = p <= 0.f ? 1.f : 0.f;    // :534
          else pw = kH5 ? (2.f * p - 1.f) : p;                                        // testh5.py:67
        }

        if (MODE == V2V_SCATTER_H5_DISCRETE) {
          const long long tau = tau_us(d.ts, d.ts_dtype, e, e0);
          const double bf = floor(__ddiv_rn(static_cast<double>(tau), h5_tpb));       // :72
          if (!(bf >= 0.0 && bf < static_cast<double>(B))) { ++ndrop; continue; }
          atomicAdd(&acc_i[static_cast<int>(bf) * plane + cell], static_cast<int>(pw));
        } else if (MODE == V2V_SCATTER_H5_INTERP) {
          const long long tau = tau_us(d.ts, d.ts_dtype, e, e0);
          const double tn = __dmul_rn(__ddiv_rn(static_cast<double>(tau), h5_den), static_cast<double>(B - 1));   // :77
          const double fl = floor(tn);
          if (!(fl >= 0.0 && fl < static_cast<double>(B))) { ++ndrop; continue; }
          const int b0 = static_cast<int>(fl);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int bi = b0 + k;
            if (bi >= B) break;
            const double wgt = fmax(0.0, __dsub_rn(1.0, fabs(__dsub_rn(tn, static_cast<double>(bi)))));   // :79
            const long long fx = __double2ll_rn(__dmul_rn(wgt * static_cast<double>(pw), 1073741824.0));
            if (fx != 0) {
              const int hi = static_cast<int>(fx >> kLoBits);
              const unsigned int lo = static_cast<unsigned int>(fx & ((1 << kLoBits) - 1));
              if (hi) atomicAdd(&acc_i[bi * plane + cell], hi);
              if (lo) atomicAdd(reinterpret_cast<unsigned int*>(&acc_lo[bi * plane + cell]), lo);
            }
          }
        } else if (MODE == V2V_SCATTER_TORCH_DISCRETE) {
          const float rel = __fsub_rn(load_f32(d.ts, d.ts_dtype, e), t_first);
          const float bf = floorf(__fdiv_rn(rel, t_tpb));                                                     // :504
          if (!(bf >= 0.f && bf < static_cast<float>(B))) { ++ndrop; continue; }
          atomicAdd(&acc_f[static_cast<int>(bf) * plane + cell], pw);
        } else {   // TORCH_BILINEAR
          const float rel = __fsub_rn(load_f32(d.ts, d.ts_dtype, e), t_first);
          const float tn = __fmul_rn(__fdiv_rn(rel, t_span), static_cast<float>(B - 1));                      // :490
          const float fl = floorf(tn);
          if (!(fl >= 0.f && fl < static_cast<float>(B))) { ++ndrop; continue; }
          const int b0 = static_cast<int>(fl);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int bi = b0 + k;
            if (bi >= B) break;
            const float wgt = fmaxf(0.f, __fsub_rn(1.0f, fabsf(__fsub_rn(tn, static_cast<float>(bi)))));     // :494
            const float v = __fmul_rn(pw, wgt);                                                               // :495
            if (v != 0.f) atomicAdd(&acc_f[bi * plane + cell], v);
          }
        }
      }
      if (d.dropped && ndrop) atomicAdd(reinterpret_cast<unsigned long long*>(d.dropped), static_cast<unsigned long long>(ndrop));
    }
    __syncthreads();

    // stream the strip out: [win, b, r0 + r, x]
    {
      const int plane = rows * W;
      const int64_t out_base = (static_cast<int64_t>(win) * B) * H * W + static_cast<int64_t>(r0) * W;
      for (int i = threadIdx.x; i < cells; i += kScatterThreads) {
        const int b = i / plane, rem = i - b * plane;
        double v;
        if (kInterp) {
          const long long tot = static_cast<long long>(acc_i[i]) * (1 << kLoBits) +
                                static_cast<long long>(reinterpret_cast<unsigned int*>(acc_lo)[i]);
          v = static_cast<double>(tot) * (1.0 / 1073741824.0);
        } else if (kTorch) {
          v = static_cast<double>(acc_f[i]);
        } else {
          v = static_cast<double>(acc_i[i]);
        }
        const int64_t o = out_base + static_cast<int64_t>(b) * H * W + rem;
        if (d.out_dtype == V2V_F64) static_cast<double*>(d.voxel)[o] = v;
        else st_stream_f32(static_cast<float*>(d.voxel) + o, static_cast<float>(v));
      }
    }
    __syncthreads();
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
  V2V_REQUIRE(d.polarity_mode >= 0 && d.polarity_mode <= 2, V2V_ERR_INVALID_ARG, "bad polarity_mode %d", d.polarity_mode);
  V2V_REQUIRE(d.out_dtype == V2V_F32 || d.out_dtype == V2V_F64, V2V_ERR_INVALID_ARG, "out_dtype must be F32 or F64");
  if (d.num_windows == 0 || d.H == 0 || d.W == 0) return V2V_OK;
  V2V_REQUIRE(d.window_offsets && d.voxel, V2V_ERR_INVALID_ARG, "window_offsets and voxel must be non-NULL");
  V2V_REQUIRE(d.num_events == 0 || (d.xs && d.ys && d.ts && d.ps), V2V_ERR_INVALID_ARG, "event arrays must be non-NULL");
  V2V_REQUIRE(coord_dtype_ok(d.xs_dtype) && coord_dtype_ok(d.ys_dtype), V2V_ERR_INVALID_ARG, "bad coordinate dtype");
  const bool h5 = d.mode == V2V_SCATTER_H5_DISCRETE || d.mode == V2V_SCATTER_H5_INTERP;
  V2V_REQUIRE(!h5 || d.ts_dtype == V2V_F64 || d.ts_dtype == V2V_F32, V2V_ERR_INVALID_ARG, "h5 modes need F64 or F32 timestamps");
  V2V_REQUIRE(h5 || d.ts_dtype == V2V_F32 || d.ts_dtype == V2V_F64, V2V_ERR_INVALID_ARG, "torch modes need float timestamps");
  V2V_REQUIRE(d.ps_dtype == V2V_U8 || d.ps_dtype == V2V_I8 || d.ps_dtype == V2V_F32, V2V_ERR_INVALID_ARG, "bad polarity dtype");

  ScatterArgs a;
  a.d = d;
  const int cell_bytes = d.mode == V2V_SCATTER_H5_INTERP ? 8 : 4;
  const int64_t row_bytes = static_cast<int64_t>(d.num_bins) * d.W * cell_bytes;
  V2V_REQUIRE(row_bytes <= kSmemBudget, V2V_ERR_UNSUPPORTED, "num_bins*W=%d*%d does not fit one shared-memory row tile", d.num_bins, d.W);
  a.rows_per_strip = static_cast<int>(kSmemBudget / row_bytes);
  if (a.rows_per_strip > d.H) a.rows_per_strip = d.H;
  a.num_strips = (d.H + a.rows_per_strip - 1) / a.rows_per_strip;
  const size_t smem = static_cast<size_t>(a.rows_per_strip) * row_bytes;
  const int64_t items = static_cast<int64_t>(d.num_windows) * a.num_strips;
  int dev = 0, sms = 148;
  V2V_CUDA(cudaGetDevice(&dev));
  V2V_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = static_cast<int>(items < 4LL * sms ? items : 4LL * sms);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define V2V_LAUNCH(M)                                                                                      \
  do {                                                                                                     \
    V2V_CUDA(cudaFuncSetAttribute(scatter_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget)); \
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

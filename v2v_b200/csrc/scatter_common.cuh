// Pieces shared by the two event-stream scatter paths (scatter.cu: contiguous bin ranges; scatter_sorted.cu: counting
// sort by strip): dtype-dispatched loads and the reference's timestamp / bin arithmetic in the reference's dtypes.
#pragma once
#include "common.cuh"

#include <atomic>

namespace v2v {

constexpr int kFixShift = 30, kLoBits = 15;     // interpolated weights: round(w * 2^30), split into two 32-bit words

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

// Bin index of event e in its window (the quantity that orders the events of a window: timestamps are
// non-decreasing inside a window, so "bin(e) >= b" is a monotone predicate and every bin owns a contiguous range).
struct WinConst {
  double h5_tpb, h5_den;
  float t_first, t_span, t_tpb;
  int64_t e0;
};

template <int MODE>
__device__ __forceinline__ WinConst window_constants(const v2v_scatter_desc& d, int64_t e0, int64_t e1) {
  WinConst c;
  c.e0 = e0;
  c.h5_tpb = c.h5_den = 0.0;
  c.t_first = c.t_span = c.t_tpb = 0.f;
  const int B = d.num_bins;
  if (MODE == V2V_SCATTER_H5_DISCRETE || MODE == V2V_SCATTER_H5_INTERP) {
    const long long tl = tau_us(d.ts, d.ts_dtype, e1 - 1, e0);
    c.h5_tpb = __ddiv_rn(__dadd_rn(static_cast<double>(tl), 0.001), static_cast<double>(B));       // testh5.py:71
    c.h5_den = __dadd_rn(static_cast<double>(tl), 0.0001);                                         // :76-77 (ts[0]==0)
  } else {
    c.t_first = load_f32(d.ts, d.ts_dtype, e0);
    c.t_span = __fsub_rn(load_f32(d.ts, d.ts_dtype, e1 - 1), c.t_first);                           // event_utils.py:489
    c.t_tpb = __fdiv_rn(__fadd_rn(c.t_span, 0.001f), static_cast<float>(B));                       // :503
  }
  return c;
}

// floor of the (possibly fractional) bin coordinate of event e; *frac_coord receives the coordinate for the
// interpolating modes.  Same expressions, same dtypes as the reference.
template <int MODE>
__device__ __forceinline__ double bin_floor(const v2v_scatter_desc& d, const WinConst& c, int64_t e, double* coord) {
  const int B = d.num_bins;
  if (MODE == V2V_SCATTER_H5_DISCRETE) {
    const long long tau = tau_us(d.ts, d.ts_dtype, e, c.e0);
    return floor(__ddiv_rn(static_cast<double>(tau), c.h5_tpb));                                   // testh5.py:72
  } else if (MODE == V2V_SCATTER_H5_INTERP) {
    const long long tau = tau_us(d.ts, d.ts_dtype, e, c.e0);
    const double tn = __dmul_rn(__ddiv_rn(static_cast<double>(tau), c.h5_den), static_cast<double>(B - 1));   // :77
    *coord = tn;
    return floor(tn);
  } else if (MODE == V2V_SCATTER_TORCH_DISCRETE) {
    const float rel = __fsub_rn(load_f32(d.ts, d.ts_dtype, e), c.t_first);
    return static_cast<double>(floorf(__fdiv_rn(rel, c.t_tpb)));                                   // event_utils.py:504
  } else {
    const float rel = __fsub_rn(load_f32(d.ts, d.ts_dtype, e), c.t_first);
    const float tn = __fmul_rn(__fdiv_rn(rel, c.t_span), static_cast<float>(B - 1));               // :490
    *coord = static_cast<double>(tn);
    return static_cast<double>(floorf(tn));
  }
}


// scatter_sorted.cu
bool scatter_sorted_eligible(const v2v_scatter_desc& d);
int launch_scatter_sorted(const v2v_scatter_desc& d, cudaStream_t s);
size_t scatter_sorted_workspace_bytes(const v2v_scatter_desc& d, int* rows_per_strip, int* strips);

}  // namespace v2v

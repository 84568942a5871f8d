// Small helpers either side of the scatter path: window segmentation of an event stream by
// time borders and the raw [N,5] event packing the NER-Net loader hands to its model.
//
// Replaces  np.searchsorted(f["events/ts"], border_timestamps)     data/testh5.py:468-474 (FPS_H5Dataset.__init__)
//           np.stack([xs, ys, ts, ps*2-1, 0], axis=1) as float64   data/testh5.py:329-339 (TestH5EventDataset.__getitem__)
#include "common.cuh"

namespace v2v {
namespace {

// out[i] = first index e in [0, n) with ts[e] >= borders[i]   (np.searchsorted side='left' on a sorted array)
__global__ void searchsorted_kernel(const double* __restrict__ ts, int64_t n, const double* __restrict__ borders, int64_t nb,
                                    int64_t* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  const double b = borders[i];
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = lo + ((hi - lo) >> 1);
    if (ts[mid] < b) lo = mid + 1; else hi = mid;
  }
  out[i] = lo;
}

__device__ __forceinline__ double load_as_f64(const void* p, int dtype, int64_t i) {
  switch (dtype) {
    case V2V_U8: return static_cast<double>(static_cast<const uint8_t*>(p)[i]);
    case V2V_I8: return static_cast<double>(static_cast<const int8_t*>(p)[i]);
    case V2V_U16: return static_cast<double>(static_cast<const uint16_t*>(p)[i]);
    case V2V_I16: return static_cast<double>(static_cast<const int16_t*>(p)[i]);
    case V2V_I32: return static_cast<double>(static_cast<const int32_t*>(p)[i]);
    case V2V_I64: return static_cast<double>(static_cast<const int64_t*>(p)[i]);
    case V2V_F32: return static_cast<double>(static_cast<const float*>(p)[i]);
    case V2V_F64: return static_cast<const double*>(p)[i];
  }
  return 0.0;
}

struct PackArgs {
  const void *xs, *ys, *ts, *ps;
  int xs_dtype, ys_dtype, ts_dtype, ps_dtype;
  int64_t n;
  double* out;
};

// out[e] = [x, y, t, 2p-1, 0] as float64; 40 bytes out per event, written as 5 coalesced-ish doubles per thread
__global__ void pack_events_kernel(const PackArgs a) {
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < a.n;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    double* o = a.out + 5 * e;
    o[0] = load_as_f64(a.xs, a.xs_dtype, e);
    o[1] = load_as_f64(a.ys, a.ys_dtype, e);
    o[2] = load_as_f64(a.ts, a.ts_dtype, e);
    o[3] = load_as_f64(a.ps, a.ps_dtype, e) * 2 - 1;          // ps * 2 - 1   (data/testh5.py:334)
    o[4] = 0.0;                                               // batch index (asserted batch size 1, :337)
  }
}

}  // namespace
}  // namespace v2v

extern "C" int v2v_searchsorted_f64(const double* sorted, int64_t n, const double* values, int64_t num_values, int64_t* out,
                                    void* stream) {
  using namespace v2v;
  V2V_REQUIRE(n >= 0 && num_values >= 0, V2V_ERR_INVALID_ARG, "negative size");
  if (num_values == 0) return V2V_OK;
  V2V_REQUIRE((sorted || n == 0) && values && out, V2V_ERR_INVALID_ARG, "NULL pointer");
  searchsorted_kernel<<<static_cast<unsigned int>((num_values + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      sorted, n, values, num_values, out);
  count_launch();
  V2V_CUDA(cudaGetLastError());
  return V2V_OK;
}

extern "C" int v2v_pack_events_n5(const void* xs, int xs_dtype, const void* ys, int ys_dtype, const void* ts, int ts_dtype,
                                  const void* ps, int ps_dtype, int64_t num_events, double* out, void* stream) {
  using namespace v2v;
  V2V_REQUIRE(num_events >= 0, V2V_ERR_INVALID_ARG, "negative size");
  if (num_events == 0) return V2V_OK;
  V2V_REQUIRE(xs && ys && ts && ps && out, V2V_ERR_INVALID_ARG, "NULL pointer");
  for (int t : {xs_dtype, ys_dtype, ts_dtype, ps_dtype}) V2V_REQUIRE(t >= V2V_U8 && t <= V2V_F64, V2V_ERR_INVALID_ARG, "bad dtype %d", t);
  PackArgs a{xs, ys, ts, ps, xs_dtype, ys_dtype, ts_dtype, ps_dtype, num_events, out};
  int64_t blocks = (num_events + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  pack_events_kernel<<<static_cast<unsigned int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  count_launch();
  V2V_CUDA(cudaGetLastError());
  return V2V_OK;
}

// ---------------------------------------------------------------------------------------------
// Voxel-space noise augmentation of the cached-voxel loader
//   replaces  add_noise_to_voxel          data/esim_dataset.py:33-46
//             add_hot_pixels_to_voxels    data/esim_dataset.py:7-30 (the broadcast add of the [H,W] noise map)
// ---------------------------------------------------------------------------------------------
namespace v2v {
namespace {

struct VoxNoiseArgs {
  float* voxel;            // [n] in/out
  int64_t n;
  const double* noise;     // explicit: [n] noise values (already scaled), or NULL
  const double* mask_u;    // explicit: [n] uniforms; element keeps its noise iff mask_u < noise_fraction (NULL: keep all)
  double noise_std, noise_fraction, lambda;
  int integer_noise, philox;
  uint32_t rk[20];
  uint64_t stream_id;
};

__device__ __forceinline__ int poisson_knuth(float lam, uint32_t w0, uint32_t w1) {
  // inversion on a 32-bit uniform, second word continues the search for large counts
  const float u = (static_cast<float>(w0 >> 8) + 0.5f) * (1.0f / 16777216.0f);
  float p = __expf(-lam), cdf = p;
  int k = 0;
  while (u >= cdf && k < 256) {
    ++k;
    p *= lam / static_cast<float>(k);
    cdf += p;
    if (p < 1e-12f) break;
  }
  (void)w1;
  return k;
}

__global__ void voxel_noise_kernel(const VoxNoiseArgs a) {
  // 4 elements per thread and Philox call
  const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t e0 = g * 4;
  if (e0 >= a.n) return;
  uint4 r = make_uint4(0, 0, 0, 0), r2 = make_uint4(0, 0, 0, 0);
  if (a.philox) {
    r = Philox::run_rk(make_uint4(static_cast<uint32_t>(g), static_cast<uint32_t>(g >> 32), static_cast<uint32_t>(a.stream_id), 0x51u), a.rk);
    r2 = Philox::run_rk(make_uint4(static_cast<uint32_t>(g), static_cast<uint32_t>(g >> 32), static_cast<uint32_t>(a.stream_id), 0x52u), a.rk);
  }
  const uint32_t w[4] = {r.x, r.y, r.z, r.w}, w2[4] = {r2.x, r2.y, r2.z, r2.w};
  float2 n01 = make_float2(0.f, 0.f), n23 = make_float2(0.f, 0.f);
  if (a.philox && !a.integer_noise) {
    n01 = box_muller(r.x, r.y);
    n23 = box_muller(r.z, r.w);
  }
  const float z[4] = {n01.x, n01.y, n23.x, n23.y};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t e = e0 + k;
    if (e >= a.n) break;
    double noise;
    bool keep = true;
    if (a.philox) {
      if (a.integer_noise) {      // y ~ Poisson(lambda), sign +-1 with equal probability   (:36-39)
        const int y = poisson_knuth(static_cast<float>(a.lambda), w[k], 0u);
        noise = static_cast<double>((w2[k] & 1u) ? y : -y);
      } else {
        noise = a.noise_std * static_cast<double>(z[k]);                                   // :41
      }
      if (a.noise_fraction < 1.0) keep = (static_cast<double>(w2[k] >> 8) * (1.0 / 16777216.0)) < a.noise_fraction;   // :43-45
    } else {
      noise = a.noise ? a.noise[e] : 0.0;
      if (a.mask_u && a.noise_fraction < 1.0) keep = a.mask_u[e] < a.noise_fraction;       // mask = rand >= fraction -> 0
    }
    if (keep) a.voxel[e] = static_cast<float>(static_cast<double>(a.voxel[e]) + noise);    // voxel + noise in float64 (:46)
  }
}

// voxel[p, i] += map[i] for every plane p (hot-pixel map broadcast over T and C, data/esim_dataset.py:27-29)
__global__ void voxel_add_map_kernel(float* voxel, const double* map, int64_t planes, int64_t hw) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= hw) return;
  const double m = map[i];
  if (m == 0.0) return;
  for (int64_t p = blockIdx.y; p < planes; p += gridDim.y) {
    float* v = voxel + p * hw + i;
    *v = static_cast<float>(static_cast<double>(*v) + m);
  }
}

}  // namespace
}  // namespace v2v

extern "C" int v2v_voxel_add_noise(float* voxel, int64_t n, const double* noise, const double* mask_u, double noise_std,
                                   double noise_fraction, int integer_noise, int philox, uint64_t seed, uint64_t stream_id,
                                   void* stream) {
  using namespace v2v;
  V2V_REQUIRE(n >= 0, V2V_ERR_INVALID_ARG, "negative size");
  if (n == 0) return V2V_OK;
  V2V_REQUIRE(voxel != nullptr, V2V_ERR_INVALID_ARG, "voxel is NULL");
  V2V_REQUIRE(philox || noise, V2V_ERR_INVALID_ARG, "explicit mode needs the noise field");
  VoxNoiseArgs a;
  a.voxel = voxel;
  a.n = n;
  a.noise = noise;
  a.mask_u = mask_u;
  a.noise_std = noise_std;
  a.noise_fraction = noise_fraction;
  a.lambda = (-1.0 + sqrt(1.0 + 4.0 * noise_std * noise_std)) / 2.0;       // data/esim_dataset.py:36
  a.integer_noise = integer_noise;
  a.philox = philox;
  a.stream_id = stream_id;
  Philox::round_keys(seed, a.rk);
  const int64_t groups = (n + 3) / 4;
  voxel_noise_kernel<<<static_cast<unsigned int>((groups + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  count_launch();
  V2V_CUDA(cudaGetLastError());
  return V2V_OK;
}

// bgr_to_gray (data/v2v_datasets.py:19-22): uint8(dot(img[..., :3], [0.5870, 0.1140, 0.2989])).  NumPy evaluates the
// length-3 dot product as fma(c2, w2, fma(c1, w1, c0*w0)) on FMA hosts (pinned by tests/golden: 261 of the 2^24 colour
// triples land exactly on an integer in rational arithmetic, and the summation order decides which side they fall on).
namespace v2v {
namespace {
__global__ void bgr_to_gray_kernel(const uint8_t* __restrict__ img, int channels, uint8_t* __restrict__ gray, int64_t n) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint8_t* p = img + i * channels;
    const double s = __fma_rn(static_cast<double>(p[2]), 0.2989,
                              __fma_rn(static_cast<double>(p[1]), 0.1140, __dmul_rn(static_cast<double>(p[0]), 0.5870)));
    gray[i] = static_cast<uint8_t>(static_cast<int>(s));          // astype(np.uint8) of a value in [0, 255): truncation
  }
}
}  // namespace
}  // namespace v2v

extern "C" int v2v_bgr_to_gray(const uint8_t* img, int32_t channels, uint8_t* gray, int64_t num_pixels, void* stream) {
  using namespace v2v;
  V2V_REQUIRE(num_pixels >= 0 && channels >= 3, V2V_ERR_INVALID_ARG, "need num_pixels >= 0 and channels >= 3");
  if (num_pixels == 0) return V2V_OK;
  V2V_REQUIRE(img && gray, V2V_ERR_INVALID_ARG, "NULL pointer");
  const int64_t blocks = (num_pixels + 255) / 256;
  bgr_to_gray_kernel<<<static_cast<unsigned int>(blocks < 148 * 32 ? blocks : 148 * 32), 256, 0, static_cast<cudaStream_t>(stream)>>>(img, channels, gray, num_pixels);
  count_launch();
  V2V_CUDA(cudaGetLastError());
  return V2V_OK;
}

extern "C" int v2v_voxel_add_map(float* voxel, int64_t planes, int64_t hw, const double* map, void* stream) {
  using namespace v2v;
  V2V_REQUIRE(planes >= 0 && hw >= 0, V2V_ERR_INVALID_ARG, "negative size");
  if (planes == 0 || hw == 0) return V2V_OK;
  V2V_REQUIRE(voxel && map, V2V_ERR_INVALID_ARG, "NULL pointer");
  dim3 grid(static_cast<unsigned int>((hw + 255) / 256), static_cast<unsigned int>(planes < 64 ? planes : 64));
  voxel_add_map_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(voxel, map, planes, hw);
  count_launch();
  V2V_CUDA(cudaGetLastError());
  return V2V_OK;
}

// ---------------------------------------------------------------------------------------------
// Consumer-side helpers of the voxel tensor
//   per-bin |count| sums of a [planes, bins, plane_elems] voxel batch (statistics vector of the all-reduce)
//   normalize_batch_voxel            model/train_utils.py:147-166
// ---------------------------------------------------------------------------------------------
namespace v2v {
namespace {

constexpr int kHistK = 255;                 // integer voxel values -255..255 are histogrammed exactly

// sums[b] += sum |voxel[g, b, :]| over all groups g: double accumulation (exact for integer-valued voxels), one atomic per CTA
__global__ void bin_abs_sums_kernel(const float* __restrict__ voxel, int64_t groups, int bins, int64_t plane, double* sums) {
  const int b = blockIdx.y;
  double acc = 0.0;
  const int64_t per_bin = groups * plane;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < per_bin; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t g = i / plane, e = i - g * plane;
    acc += fabs(static_cast<double>(voxel[(g * bins + b) * plane + e]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) t += part[w];
    if (t != 0.0) atomicAdd(sums + b, t);
  }
}

// hist[b, v + K] += #elements of clip b equal to the integer v (|v| <= K); hist[b, 2K+1] += #elements that are not
__global__ void value_hist_kernel(const float* __restrict__ voxel, int64_t per_clip, long long* hist) {
  __shared__ unsigned int h_s[2 * kHistK + 2];
  for (int i = threadIdx.x; i < 2 * kHistK + 2; i += blockDim.x) h_s[i] = 0u;
  __syncthreads();
  const float* v = voxel + static_cast<int64_t>(blockIdx.y) * per_clip;
  unsigned int zeros = 0u;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < per_clip; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float x = v[i];
    if (x == 0.f) { ++zeros; continue; }                 // the bulk: counted in a register
    const float r = rintf(x);
    const bool ok = r == x && fabsf(x) <= static_cast<float>(kHistK);
    atomicAdd(&h_s[ok ? static_cast<int>(r) + kHistK : 2 * kHistK + 1], 1u);
  }
  zeros = __reduce_add_sync(0xffffffffu, zeros);
  if ((threadIdx.x & 31) == 0 && zeros) atomicAdd(&h_s[kHistK], zeros);
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * kHistK + 2; i += blockDim.x)
    if (h_s[i]) atomicAdd(reinterpret_cast<unsigned long long*>(hist + static_cast<int64_t>(blockIdx.y) * (2 * kHistK + 2) + i), static_cast<unsigned long long>(h_s[i]));
}

// norm = where(voxel > 0, voxel / pos_max[b], voxel / neg_max[b]) in float32 (model/train_utils.py:165), in place
__global__ void normalize_kernel(float* voxel, int64_t per_clip, const float* pos_max, const float* neg_max) {
  const float pm = pos_max[blockIdx.y], nm = neg_max[blockIdx.y];
  float* v = voxel + static_cast<int64_t>(blockIdx.y) * per_clip;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < per_clip; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float x = v[i];
    if (x != 0.f) v[i] = x > 0.f ? __fdiv_rn(x, pm) : __fdiv_rn(x, nm);
  }
}

}  // namespace
}  // namespace v2v

extern "C" int v2v_voxel_bin_abs_sums(const float* voxel, int64_t groups, int32_t bins, int64_t plane_elems, double* sums, void* stream) {
  using namespace v2v;
  V2V_REQUIRE(groups >= 0 && bins >= 1 && bins <= 65535 && plane_elems >= 0, V2V_ERR_INVALID_ARG, "bad sizes");
  if (groups == 0 || plane_elems == 0) return V2V_OK;
  V2V_REQUIRE(voxel && sums, V2V_ERR_INVALID_ARG, "NULL pointer");
  int64_t blocks = (groups * plane_elems + 256 * 16 - 1) / (256 * 16);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  bin_abs_sums_kernel<<<dim3(static_cast<unsigned int>(blocks), static_cast<unsigned int>(bins)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      voxel, groups, bins, plane_elems, sums);
  count_launch();
  V2V_CUDA(cudaGetLastError());
  return V2V_OK;
}

extern "C" int v2v_voxel_value_hist(const float* voxel, int32_t clips, int64_t elems_per_clip, long long* hist, void* stream) {
  using namespace v2v;
  V2V_REQUIRE(clips >= 0 && clips <= 65535 && elems_per_clip >= 0, V2V_ERR_INVALID_ARG, "bad sizes");
  if (clips == 0 || elems_per_clip == 0) return V2V_OK;
  V2V_REQUIRE(voxel && hist, V2V_ERR_INVALID_ARG, "NULL pointer");
  int64_t blocks = (elems_per_clip + 256 * 32 - 1) / (256 * 32);
  const int64_t cap = (148 * 16 + clips - 1) / clips;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  value_hist_kernel<<<dim3(static_cast<unsigned int>(blocks), static_cast<unsigned int>(clips)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      voxel, elems_per_clip, hist);
  count_launch();
  V2V_CUDA(cudaGetLastError());
  return V2V_OK;
}

extern "C" int v2v_voxel_normalize(float* voxel, int32_t clips, int64_t elems_per_clip, const float* pos_max, const float* neg_max,
                                   void* stream) {
  using namespace v2v;
  V2V_REQUIRE(clips >= 0 && clips <= 65535 && elems_per_clip >= 0, V2V_ERR_INVALID_ARG, "bad sizes");
  if (clips == 0 || elems_per_clip == 0) return V2V_OK;
  V2V_REQUIRE(voxel && pos_max && neg_max, V2V_ERR_INVALID_ARG, "NULL pointer");
  int64_t blocks = (elems_per_clip + 256 * 16 - 1) / (256 * 16);
  const int64_t cap = (148 * 16 + clips - 1) / clips;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  normalize_kernel<<<dim3(static_cast<unsigned int>(blocks), static_cast<unsigned int>(clips)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      voxel, elems_per_clip, pos_max, neg_max);
  count_launch();
  V2V_CUDA(cudaGetLastError());
  return V2V_OK;
}

// ---- learned-representation scatter (NER-Net quantization layer, model/nernet/representation_modules.py:143-168) ----
namespace v2v {
namespace {
template <bool PUT>
__global__ void __launch_bounds__(256) put_take_bins_kernel(float* out, int64_t numel, const int64_t* idx, float* values, int64_t n, int bins,
                                                            int64_t stride, long long* bad) {
  long long nbad = 0;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; e < n; e += static_cast<int64_t>(gridDim.x) * 256) {
    const int64_t base = idx[e];                     // one index load serves every bin of the event
    for (int b = 0; b < bins; ++b) {
      int64_t i = base + stride * b;
      if (i > numel - 1) i = numel - 1;              // torch.clamp(idx, max=numel-1), :166
      if (i < 0) {
        ++nbad;
        if (!PUT) values[static_cast<int64_t>(b) * n + e] = 0.f;
        continue;
      }
      if (PUT) atomicAdd(out + i, values[static_cast<int64_t>(b) * n + e]);
      else values[static_cast<int64_t>(b) * n + e] = out[i];
    }
  }
  if (PUT && bad && nbad) atomicAdd(reinterpret_cast<unsigned long long*>(bad), static_cast<unsigned long long>(nbad));
}
}  // namespace
}  // namespace v2v

extern "C" int v2v_put_accumulate_bins(float* out, int64_t out_numel, const int64_t* idx, const float* values, int64_t n, int32_t num_bins,
                                       int64_t bin_stride, long long* bad, void* stream) {
  using namespace v2v;
  V2V_REQUIRE(out_numel >= 1 && n >= 0 && num_bins >= 1, V2V_ERR_INVALID_ARG, "bad sizes");
  if (n == 0) return V2V_OK;
  V2V_REQUIRE(out && idx && values, V2V_ERR_INVALID_ARG, "NULL pointer");
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  put_take_bins_kernel<true><<<static_cast<unsigned int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      out, out_numel, idx, const_cast<float*>(values), n, num_bins, bin_stride, bad);
  count_launch();
  V2V_CUDA(cudaGetLastError());
  return V2V_OK;
}

extern "C" int v2v_take_bins(const float* src, int64_t src_numel, const int64_t* idx, float* values_out, int64_t n, int32_t num_bins,
                             int64_t bin_stride, void* stream) {
  using namespace v2v;
  V2V_REQUIRE(src_numel >= 1 && n >= 0 && num_bins >= 1, V2V_ERR_INVALID_ARG, "bad sizes");
  if (n == 0) return V2V_OK;
  V2V_REQUIRE(src && idx && values_out, V2V_ERR_INVALID_ARG, "NULL pointer");
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  put_take_bins_kernel<false><<<static_cast<unsigned int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      const_cast<float*>(src), src_numel, idx, values_out, n, num_bins, bin_stride, nullptr);
  count_launch();
  V2V_CUDA(cudaGetLastError());
  return V2V_OK;
}

/*
 * v2v_b200.h — C ABI of the B200-native video-to-voxel hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.
 * Every compute entry point is stream-ordered and non-blocking, allocates
 * nothing, keeps no global mutable state, and returns 0 on success or a
 * negative V2V_ERR_* code (never throws across the ABI).  All data pointers are
 * DEVICE pointers unless a name ends in `_host`; the caller owns every buffer
 * and keeps it alive until the stream has passed the call.  `stream` is a
 * `cudaStream_t` passed as `void*` (NULL = legacy default stream).
 *
 * The reference (HYLZ-2019/V2V) has no FFI: its seams are Python call
 * signatures.  Each entry point cites the reference function(s) it replaces
 * (paths relative to the reference checkout); INTEGRATION.md shows the ctypes
 * binding a maintainer adds on the reference side.
 */
#ifndef V2V_B200_H_
#define V2V_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define V2V_ABI_VERSION 2

/* ---- error codes ------------------------------------------------------- */
#define V2V_OK 0
#define V2V_ERR_INVALID_ARG (-1)   /* null pointer / non-positive size / bad enum  */
#define V2V_ERR_SHAPE (-2)         /* (N-1) % (num_bins*frames_per_bin) != 0, ...  */
#define V2V_ERR_ALIGNMENT (-3)     /* pointer not aligned for its element type     */
#define V2V_ERR_CUDA (-4)          /* a CUDA runtime call failed (see last_error)  */
#define V2V_ERR_UNSUPPORTED (-5)   /* combination not implemented                  */
#define V2V_ERR_NO_DEVICE (-6)     /* no sm_100 device                             */

/* ---- queries ----------------------------------------------------------- */
int v2v_abi_version(void);
/* Thread-local description of the last error returned on this thread. */
const char* v2v_last_error(void);
/* Compute capability / SM count of `device`; V2V_ERR_NO_DEVICE without a GPU. */
int v2v_device_info(int device, int* cc_major, int* cc_minor, int* sm_count);
/* Number of kernels this library has launched in this process (all threads). */
long long v2v_launch_count(void);

/* ======================================================================= *
 * 1. ESIM-style frames -> voxel
 *    replaces  EventEmulator.video_to_voxel      data/v2v_core_esim.py:26-69
 *              reverse_gamma_correction + log    data/v2v_core_esim.py:3-4,34 (as the 256-entry LUT)
 *              bin accumulation                  data/v2v_datasets.py:365-366,399-400
 *              float32 packing of events/frame   data/v2v_datasets.py:328-356
 * ======================================================================= */

enum v2v_noise_mode {
  V2V_NOISE_NONE = 0,     /* no base noise, hot noise only if `hot_noise` given        */
  V2V_NOISE_EXPLICIT = 1, /* caller passes the reference's random fields (bit parity)  */
  V2V_NOISE_PHILOX = 2    /* in-kernel generator (throughput mode): Philox4x32-10 root, one 64-bit LCG stream per (clip, 4-pixel group) */
};

enum v2v_threshold_mode {
  V2V_THRES_PER_CLIP = 0, /* pos_thres/neg_thres are [B]      (the ESIM core)          */
  V2V_THRES_PER_PIXEL = 1 /* pos_thres/neg_thres are [B,H,W]  (per-pixel maps)         */
};

/* v2v_esim_desc.kernel_flags: explicit per-call kernel selection instead of process-wide environment variables */
#define V2V_ESIM_FLAG_GENERIC 1      /* always the generic kernel (esim.cu); tests compare the two paths bit for bit   */
#define V2V_ESIM_FLAG_SMALL_FAST 2   /* launches below 148*2048 pixels: throughput kernel                             */
#define V2V_ESIM_FLAG_SMALL_P1 4     /* launches below 148*2048 pixels: one pixel per thread                          */
#define V2V_ESIM_FLAG_STAGED 16      /* noise-free throughput kernel: frames through a cp.async.bulk ring in shared memory (A/B)  */
#define V2V_ESIM_FLAG_GEOM(g) (((g) & 0xf) << 8) /* CTA geometry index of the throughput kernel (tuning sweeps)       */

typedef struct v2v_esim_desc {
  /* shapes */
  int32_t B, N, H, W;            /* clips, frames per clip (>=2), height, width         */
  int32_t num_bins;              /* voxel channels                                      */
  int32_t frames_per_bin;        /* consecutive intervals summed per bin (>=1)          */
  int32_t noise_mode;            /* enum v2v_noise_mode                                 */
  int32_t put_noise_external;    /* data/v2v_core_esim.py:46,62                         */
  int32_t threshold_mode;        /* enum v2v_threshold_mode                             */
  int32_t frame_out_mode;        /* 0: none; 1: frames (t+1)*bins*fpb, t<T; 2: frames t*bins*fpb, t<=T */
  /* inputs */
  const uint8_t* frames;         /* [B,N,H,W] uint8, contiguous                         */
  const double* lut;             /* [256] float64: log(0.001 + gamma^-1(v)) built on the host */
  const double* pos_thres;       /* [B] or [B,H,W], > 0                                  */
  const double* neg_thres;       /* [B] or [B,H,W], > 0                                  */
  const double* base_noise_std;  /* [B]; may be NULL for V2V_NOISE_NONE                 */
  /* explicit random fields (V2V_NOISE_EXPLICIT; any may be NULL = absent) */
  const double* u0;              /* [B,H,W] uniform [0,1): potential0 = u0*(pos+neg)-neg */
  const double* hot_noise;       /* [B,H,W] masked+scaled hot-pixel noise               */
  const double* base_gauss;      /* [B,N-1,H,W] standard normal per interval            */
  /* Philox parameters (V2V_NOISE_PHILOX) */
  const double* hot_pixel_fraction; /* [B]                                              */
  const double* hot_pixel_std;      /* [B]                                              */
  uint64_t seed;                 /* key; clip b uses counter word b + clip_index_base   */
  uint64_t clip_index_base;      /* global index of clip 0 (so shards draw distinct streams) */
  /* optional carried state: overrides u0 / Philox initial potential when given */
  const double* potential_in;    /* [B,H,W] or NULL                                     */
  double* potential_out;         /* [B,H,W] or NULL: potential after the last interval  */
  /* outputs */
  float* voxel;                  /* [B,T,num_bins,Hp,Wp] float32, T=(N-1)/(bins*fpb)    */
  int64_t voxel_row_stride;      /* elements between rows   (0 -> W)                    */
  int64_t voxel_plane_stride;    /* elements between planes (0 -> H*W); pads are NOT written */
  float* frame_out;              /* [B,T or T+1,1,H,W] float32 = frame/255, or NULL     */
  long long* stats;              /* [B,2] int64 += {positive events, negative events}, or NULL */
  /* optional frame-side packing fused into the pass: the dataset's pause gather
   * `all_imgs = np.stack([raw_imgs[i] for i in img_idxes])` (data/v2v_datasets.py:285-311) and the HDR/LDR degrade
   * `np.clip((img-127.5)*scale+127.5, 0, 255).astype(np.uint8)` (:473-483), a function of the pixel value alone */
  const int32_t* frame_index;    /* [B,N] raw frame used as frame n of the clip (clamped to the clip), or NULL = identity */
  int32_t raw_frames_per_clip;   /* frames per clip in `frames` when frame_index is given ([B,M,H,W]); 0 = N              */
  int32_t kernel_flags;          /* 0 = library's choice; V2V_ESIM_FLAG_* (tests and tuning: kernel selection never changes results) */
  const uint8_t* value_map;      /* [B,256] uint8 -> uint8 applied to every pixel before the LUT and frame_out, or NULL   */
} v2v_esim_desc;

int v2v_esim_frames_to_voxel(const v2v_esim_desc* desc, void* stream);

/* Audit hook for V2V_NOISE_PHILOX: writes the random fields the kernel draws for (seed, clip_index_base,
 * B, N, H, W, base_noise_std, hot_pixel_fraction, hot_pixel_std) of `desc`: u0 [B,H,W], hot_noise [B,H,W],
 * base_noise [B,N-1,H,W] (already multiplied by base_noise_std).  Any output may be NULL.  Replaying them with
 * V2V_NOISE_EXPLICIT and base_noise_std = 1 reproduces the PHILOX run exactly, which is how the in-kernel
 * generator is checked against the CPU oracle. */
int v2v_esim_philox_fields(const v2v_esim_desc* desc, double* u0, double* hot_noise, double* base_noise, void* stream);

/* Audit hook for the generators themselves: writes (a) the Philox4x32-10 block of an arbitrary counter/key
 * (known-answer tests of the published algorithm) to philox_out[4] and (b) the first `n_words` 32-bit outputs of the
 * ESIM base-noise stream of (seed, clip_index, pixel group) — the Philox-seeded 64-bit LCG of esim_common.cuh —
 * to words_out[n_words].  Device pointers; either output may be NULL. */
/* Copies the generator's 2048-entry direction table to HOST memory: table_host[2k], table_host[2k+1] = high 32-bit
 * words of double(cos t_k), double(sin t_k), t_k = (2k+1)*pi/2048 (low words are zero).  No GPU needed. */
int v2v_noise_direction_table(uint32_t* table_host);
int v2v_rng_words(const uint32_t counter[4], const uint32_t key[2], uint32_t* philox_out, uint64_t seed, uint64_t clip_index,
                  uint64_t pixel_group, int32_t n_words, uint32_t* words_out, void* stream);

/* ======================================================================= *
 * 2. v2e-style frames -> voxel
 *    replaces  video_to_voxel / EventEmulator.generate_events
 *                                              data/v2v_core_v2e.py:401-581
 *              lin_log (as a 256-entry float32 LUT)   data/v2v_core_v2e.py:108-137
 *              low_pass_filter :139-182, subtract_leak_current :192-211,
 *              compute_event_map :42-62, generate_shot_noise :65-105
 * ======================================================================= */

#define V2V_V2E_FLAG_GENERIC 1       /* always the generic kernel (v2e.cu)                                          */
#define V2V_V2E_FLAG_FAST 2          /* throughput kernel also below 148*2048 pixels per launch                     */
#define V2V_V2E_FLAG_DIVERGENT_DIV 4 /* throughput kernel: single-crossing fast path + divergent exact division     */
#define V2V_V2E_FLAG_U8_INTENSITY 8 /* NOT a tuning flag: the caller's video was a uint8 array, so rescale_intensity_frame
                                       (data/v2v_core_v2e.py:190) computed new_frame+20 in uint8 and wrapped for values >= 236:
                                       inten01 = ((v+20) & 255)/275.  Without it (float video) inten01 = (v+20)/275.     */

typedef struct v2v_v2e_desc {
  int32_t B, N, H, W;
  int32_t num_bins, frames_per_bin; /* num_bins*frames_per_bin must divide N-1; use 1,1 for raw intervals */
  int32_t noise_mode;            /* EXPLICIT: leak_randn / shot tensors given; PHILOX: generated; NONE: leak jitter & shot off */
  int32_t state_f32;             /* 1: base/diff kept in float32 (reference when cutoff_hz<=0 and leak_rate_hz<=0) */
  double fps;                    /* frame k is at t=k/fps                                */
  double cutoff_hz, leak_rate_hz, shot_noise_rate_hz, leak_jitter_fraction;
  const uint8_t* frames;         /* [B,N,H,W]                                            */
  const float* lut;              /* [256] float32(log(v/255+0.01)) built on the host     */
  const double* pos_thres;       /* [B,H,W] per-pixel ON thresholds  (>=0.01)            */
  const double* neg_thres;       /* [B,H,W] per-pixel OFF thresholds (>=0.01)            */
  const float* noise_rate;       /* [B,H,W] float32 leak-rate multipliers; NULL if leak_rate_hz<=0 */
  const double* leak_randn;      /* [B,N-1,H,W] (EXPLICIT, leak_rate_hz>0)               */
  const int32_t* pos_shot;       /* [B,N-1,H,W] Poisson draws (EXPLICIT, shot>0)         */
  const int32_t* neg_shot;       /* [B,N-1,H,W]                                          */
  const double* shot_pos_scale;  /* [B,N-1] PHILOX shot noise: (rate/2*dt)/mean(pos_factor) per frame, from v2v_v2e_shot_scales */
  const double* shot_neg_scale;  /* [B,N-1]                                              */
  double pos_thres_nominal, neg_thres_nominal; /* data/v2v_core_v2e.py:297-298           */
  uint64_t seed, clip_index_base;
  float* voxel;                  /* [B,T,num_bins,H,W] float32                           */
  long long* stats;              /* [B,2] or NULL                                        */
  /* optional fused frame-side packing, as in v2v_esim_desc (pause gather + HDR/LDR degrade of the dataset) */
  const int32_t* frame_index;    /* [B,N] raw frame used as frame n (clamped to the clip), or NULL = identity            */
  int32_t raw_frames_per_clip;   /* frames per clip in `frames` when frame_index is given; 0 = N                         */
  int32_t kernel_flags;          /* 0 = library's choice; V2V_V2E_FLAG_* (kernel selection never changes results; U8_INTENSITY selects the reference's uint8 arithmetic) */
  const uint8_t* value_map;      /* [B,256] uint8 -> uint8 applied to every pixel before any use of its value, or NULL   */
  int32_t thres_per_interval;    /* 1: pos_thres / neg_thres are [B,N-1,H,W], the maps in force at frame i = interval i-1
                                  * (threshold_model "spatial_temporal_independent", data/v2v_core_v2e.py:417-421; noise_mode
                                  * NONE or EXPLICIT: the host draws the maps in the reference's order)                   */
  int32_t reserved1;
} v2v_v2e_desc;

int v2v_v2e_frames_to_voxel(const v2v_v2e_desc* desc, void* stream);

/* Per-frame shot-noise normalisers (the full-frame means of generate_shot_noise,
 * data/v2v_core_v2e.py:90-96): scales[b,k-1] = (rate/2*dt_k)/mean_pixels(inten_factor*nominal/thres). */
int v2v_v2e_shot_scales(const v2v_v2e_desc* desc, double* shot_pos_scale, double* shot_neg_scale, void* stream);

/* Audit hook for V2V_NOISE_PHILOX (like v2v_esim_philox_fields): leak_randn [B,N-1,H,W] float64, pos_shot /
 * neg_shot [B,N-1,H,W] int32 exactly as the kernel draws them; replay with V2V_NOISE_EXPLICIT is bit-identical. */
int v2v_v2e_philox_fields(const v2v_v2e_desc* desc, double* leak_randn, int32_t* pos_shot, int32_t* neg_shot, void* stream);

/* ======================================================================= *
 * 3. Event stream -> voxel (segmented by windows)
 *    replaces  TestH5Dataset.make_voxel          data/testh5.py:60-90
 *              events_to_voxel_torch             utils/event_utils.py:466-507
 *              events_to_neg_pos_voxel_torch     utils/event_utils.py:509-541
 *    Precondition: timestamps are non-decreasing inside every window (true for every h5 file the
 *    converters of scripts/*_to_h5.py write); each bin then owns a contiguous range of events.
 * ======================================================================= */

enum v2v_dtype {
  V2V_U8 = 0, V2V_I8 = 1, V2V_U16 = 2, V2V_I16 = 3, V2V_I32 = 4, V2V_I64 = 5, V2V_F32 = 6, V2V_F64 = 7
};

enum v2v_scatter_mode {
  V2V_SCATTER_H5_DISCRETE = 0,   /* data/testh5.py:70-73  (µs timestamps, integer bins)   */
  V2V_SCATTER_H5_INTERP = 1,     /* data/testh5.py:74-80  (µs timestamps, 2-tap in time)  */
  V2V_SCATTER_TORCH_DISCRETE = 2,/* utils/event_utils.py:501-505 (float32 arithmetic)      */
  V2V_SCATTER_TORCH_BILINEAR = 3 /* utils/event_utils.py:490-500                           */
};

enum v2v_polarity_mode {
  V2V_POL_SIGNED = 0,  /* weight = p (torch modes) or 2p-1 (h5 modes, p in {0,1})          */
  V2V_POL_POS_ONLY = 1,/* weight = 1[p>0]   (events_to_neg_pos_voxel_torch, pos half)      */
  V2V_POL_NEG_ONLY = 2,/* weight = 1[p<=0]                                                */
  V2V_POL_SPLIT = 3    /* both of the above in one launch: voxel is [Wn,2,num_bins,H,W], slot 0 = POS_ONLY, slot 1 = NEG_ONLY
                        * (events_to_neg_pos_voxel_torch; the caller concatenates them to 2B channels, data/dataset.py:331-333) */
};

typedef struct v2v_scatter_desc {
  int64_t num_events;
  int32_t num_windows;           /* Wn                                                    */
  int32_t num_bins, H, W;
  int32_t mode;                  /* enum v2v_scatter_mode                                 */
  int32_t polarity_mode;         /* enum v2v_polarity_mode                                */
  int32_t xs_dtype, ys_dtype;    /* U16 / I16 / I32 / I64 / F32                           */
  int32_t ts_dtype;              /* F64 / F32 (seconds)                                   */
  int32_t ps_dtype;              /* U8 / I8 / F32                                         */
  int32_t out_dtype;             /* F32 or F64                                            */
  const void* xs;
  const void* ys;
  const void* ts;
  const void* ps;
  const int64_t* window_offsets; /* [Wn+1] ascending event offsets; window w = [off[w], off[w+1]) */
  void* voxel;                   /* [Wn,num_bins,H,W] ([Wn,2,num_bins,H,W] for V2V_POL_SPLIT); fully written (zeros where no event) */
  long long* dropped;            /* [1] += events skipped (out-of-sensor / out-of-range bin), or NULL */
  void* workspace;               /* optional scratch, 16-byte aligned.  >= Wn*((num_bins+2)*8+64) bytes enables the bin-boundary
                                  * pre-pass; >= v2v_scatter_workspace_bytes(desc) additionally enables the one-visit path of
                                  * the interpolated mode (counting sort by strip: any event order, 8 bytes per event)          */
  int64_t workspace_bytes;
  long long* unsorted;           /* [1] += positions where a timestamp decreases inside a window, or NULL = not checked.  The
                                  * discrete and torch modes need non-decreasing timestamps inside every window (their result is
                                  * undefined otherwise: check this counter); the interpolated mode on the one-visit path does not */
  int32_t kernel_flags;          /* 0 = library's choice; V2V_SCATTER_FLAG_* (tests and tuning)                                */
  int32_t tuning_splits;         /* 0 = library's choice; CTAs sharing one (window, bin, strip) of a few, very large windows   */
  int32_t tuning_smem_kb;        /* 0 = library's choice; shared-memory budget of one accumulator tile in KB                   */
  int32_t reserved0;
} v2v_scatter_desc;

#define V2V_SCATTER_FLAG_RANGES 1        /* interpolated mode: contiguous-range kernel even when the workspace allows the one-visit path */
#define V2V_SCATTER_FLAG_GENERIC_SCAN 2  /* never take the 16-bit coordinate scan                                                       */
#define V2V_SCATTER_FLAG_NO_PACKED16 4   /* discrete mode: 32-bit instead of packed 16-bit counters                                      */

/* Accumulation is integer and order independent.  Discrete h5 mode: exact for any count (packed 16-bit counters while a bin
 * holds <= 32767 events, 32-bit counters otherwise).  Interpolated h5 mode: weights are rounded to 2^-30 once
 * (|error| <= n * 2^-31 per cell, n = events on that cell, bin and window; 2^-24 per event for work items of <= 255 events);
 * items that could put more than 65535 events on one cell switch to one 64-bit word per cell, so there is no count at which
 * a cell overflows.  Torch modes: float32 atomics (the reference is a sequential float32 sum). */

/* Bytes of `workspace` that enable every path for this descriptor (sizes and mode are read; pointers are not). */
int64_t v2v_scatter_workspace_bytes(const v2v_scatter_desc* desc);

int v2v_events_to_voxel(const v2v_scatter_desc* desc, void* stream);

/* ======================================================================= *
 * 4. Event stream -> image
 *    replaces  events_to_image_torch / interpolate_to_image
 *                                              utils/event_utils.py:330-376,176-184
 *              events_to_image (bincount)      utils/event_utils.py:155-174
 *              per-pixel event-count map       scripts/testset_evcnt_maps.py:19-25
 * ======================================================================= */
typedef struct v2v_image_desc {
  int64_t num_events;
  int32_t H, W;                  /* sensor size                                           */
  int32_t bilinear;              /* 0 nearest (truncate), 1 spatial 4-tap                 */
  int32_t padding;               /* bilinear: output is [H+1,W+1]                         */
  int32_t clip_out_of_range;     /* bilinear: mask events at/after the last row/column    */
  int32_t xs_dtype, ys_dtype, ps_dtype; /* ps may be NULL: weight 1 (count map)          */
  int32_t out_dtype;             /* F32, F64 or I64 (count map)                           */
  const void* xs;
  const void* ys;
  const void* ps;
  void* image;                   /* [Ho,Wo]; fully written                                */
  long long* dropped;
} v2v_image_desc;

int v2v_events_to_image(const v2v_image_desc* desc, void* stream);

/* ======================================================================= *
 * 5. Either side of the scatter
 *    replaces  np.searchsorted(events/ts, border_timestamps)   data/testh5.py:468-474 (FPS_H5Dataset windows)
 *              raw [N,5] float64 events [x, y, t, 2p-1, 0]     data/testh5.py:329-339 (TestH5EventDataset, NER-Net)
 * ======================================================================= */
/* out[i] = first index e with sorted[e] >= values[i] (np.searchsorted, side='left'); all device pointers. */
int v2v_searchsorted_f64(const double* sorted, int64_t n, const double* values, int64_t num_values, int64_t* out, void* stream);
/* out [N,5] float64; dtypes are enum v2v_dtype of the stored arrays. */
int v2v_pack_events_n5(const void* xs, int xs_dtype, const void* ys, int ys_dtype, const void* ts, int ts_dtype,
                       const void* ps, int ps_dtype, int64_t num_events, double* out, void* stream);

/* ======================================================================= *
 * 6. Voxel-space noise augmentation of cached voxels
 *    replaces  add_noise_to_voxel        data/esim_dataset.py:33-46
 *              add_hot_pixels_to_voxels  data/esim_dataset.py:7-30 (broadcast add of the [H,W] hot-pixel map;
 *                                        the map itself is an event image: v2v_events_to_image with weights)
 * ======================================================================= */
/* voxel [n] float32 in place.  philox=0: noise [n] float64 (already scaled) and optional mask_u [n] uniforms are the
 * reference's fields (element keeps its noise iff mask_u < noise_fraction).  philox=1: generated in the kernel:
 * Gaussian noise_std*z, or (integer_noise) Poisson((-1+sqrt(1+4 std^2))/2) * random sign. */
int v2v_voxel_add_noise(float* voxel, int64_t n, const double* noise, const double* mask_u, double noise_std,
                        double noise_fraction, int integer_noise, int philox, uint64_t seed, uint64_t stream_id, void* stream);
/* voxel [planes, hw] float32 += map [hw] float64 for every plane. */
/* bgr_to_gray of the dataset (data/v2v_datasets.py:19-22): gray = uint8(dot(img[..., :3], [0.5870, 0.1140, 0.2989])),
 * evaluated as fma(c2, w2, fma(c1, w1, c0*w0)) in float64 (what NumPy's dot computes for a length-3 row on FMA hosts;
 * pinned by tests/golden), truncated toward zero.  `channels` >= 3 values per pixel, the first three are used. */
int v2v_bgr_to_gray(const uint8_t* img, int32_t channels, uint8_t* gray, int64_t num_pixels, void* stream);

int v2v_voxel_add_map(float* voxel, int64_t planes, int64_t hw, const double* map, void* stream);

/* ======================================================================= *
 * 7. Consumer side of the voxel tensor
 *    replaces  normalize_batch_voxel                 model/train_utils.py:147-166
 *              per-bin |count| sums of the statistics vector
 * ======================================================================= */
/* sums[b] += sum over g, e of |voxel[g, b, e]| for a contiguous float32 [groups, bins, plane_elems] batch (float64
 * accumulation: exact for integer-valued voxels).  `sums` [bins] device float64, accumulated into (zero it first). */
int v2v_voxel_bin_abs_sums(const float* voxel, int64_t groups, int32_t bins, int64_t plane_elems, double* sums, void* stream);
/* hist[c, v + 255] += number of elements of clip c equal to the integer v, |v| <= 255; hist[c, 511] += number of elements
 * that are not such integers.  `hist` [clips, 512] device int64, accumulated into.  One read of the voxel batch: the
 * order statistics of normalize_batch_voxel (kthvalue at 1 % / 99 %) follow exactly from it when hist[c, 511] == 0. */
int v2v_voxel_value_hist(const float* voxel, int32_t clips, int64_t elems_per_clip, long long* hist, void* stream);
/* In place: voxel[c, e] = voxel > 0 ? voxel / pos_max[c] : voxel / neg_max[c]  (float32 division, :165). */
int v2v_voxel_normalize(float* voxel, int32_t clips, int64_t elems_per_clip, const float* pos_max, const float* neg_max, void* stream);

/* ======================================================================= *
 * 8. Learned-representation scatter (NER-Net quantization layer)
 *    replaces  vox.put_(idx, values, accumulate=True) called once per temporal bin
 *                                              model/nernet/representation_modules.py:143-168, 228-248
 *    out[min(idx[e] + b*bin_stride, out_numel-1)] += values[b*n + e]   for b in [0, num_bins), e in [0, n)
 *    (the reference clamps the index from above, :166/:246; a negative index is counted in `bad` and skipped, where
 *    put_ would raise).  All bins in ONE launch, float32 atomics like torch's CUDA put_.  `v2v_take_bins` is the
 *    gather of the backward pass: grad_values[b*n + e] = grad_out[that index].
 * ======================================================================= */
int v2v_put_accumulate_bins(float* out, int64_t out_numel, const int64_t* idx, const float* values, int64_t n, int32_t num_bins,
                            int64_t bin_stride, long long* bad, void* stream);
int v2v_take_bins(const float* src, int64_t src_numel, const int64_t* idx, float* values_out, int64_t n, int32_t num_bins,
                  int64_t bin_stride, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* V2V_B200_H_ */

/*
 * C restatement of the reference's hot path (TEST INFRASTRUCTURE ONLY — see oracle/v2v_oracle.py for the
 * contract; the product never links this file).  Scalar loops, one pixel / one event at a time, so that full-size
 * inputs (121x480x640 clips, 10 M events) can be checked in seconds.  Pinned to the reference through the same
 * golden vectors as the NumPy oracle (tests/test_oracle_golden.py::test_c_oracle_*).
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: no FMA contraction, IEEE double throughout)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* numpy's float floor_divide (third-party dependency of the reference, numpy/core/src/npymath: npy_divmod):
 * fmod-based quotient snapped to the nearest integer.  data/v2v_core_esim.py:51,54 call it through np.floor_divide. */
static double np_floor_divide(double a, double b) {
  if (b == 0.0) return a / b;
  double mod = fmod(a, b);
  double div = (a - mod) / b;
  if (mod != 0.0) {
    if ((b < 0) != (mod < 0)) {
      mod += b;
      div -= 1.0;
    }
  }
  if (div != 0.0) {
    double fl = floor(div);
    if (div - fl > 0.5) fl += 1.0;
    return fl;
  }
  return copysign(0.0, a / b);
}

/* EventEmulator.video_to_voxel, data/v2v_core_esim.py:26-69, random fields given explicitly.
 * video [N,HW] uint8; lut [256]; u0, hot [HW]; g [N-1,HW] or NULL (= zeros); out [N-1,HW]; pot_out [HW] or NULL. */
void orc_esim_video_to_voxel(const uint8_t* video, int n, int64_t hw, const double* lut, double pos, double neg,
                             double base_noise_std, const double* u0, const double* hot, const double* g, int external,
                             double* out, double* pot_out) {
  for (int64_t p = 0; p < hw; ++p) {
    double pot = u0[p] * (pos + neg) - neg;                                   /* :29 */
    const double h = hot ? hot[p] : 0.0;
    for (int i = 0; i < n - 1; ++i) {
      const double d = lut[video[(int64_t)(i + 1) * hw + p]] - lut[video[(int64_t)i * hw + p]];   /* :42 */
      pot += d;                                                               /* :43 */
      const double bn = base_noise_std * (g ? g[(int64_t)i * hw + p] : 0.0);  /* :44 */
      if (!external) {                                                        /* :46-49 */
        pot += bn;
        pot += h;
      }
      const double pe = pot >= pos ? np_floor_divide(pot, pos) : 0.0;         /* :51-52 */
      const double ne = pot <= -neg ? np_floor_divide(-pot, neg) : 0.0;       /* :54-55 */
      pot -= pe * pos;                                                        /* :57 */
      pot += ne * neg;                                                        /* :58 */
      double v = pe - ne;                                                     /* :60 */
      if (external) {                                                         /* :62-65 */
        v = v + bn;
        v = v + h;
      }
      out[(int64_t)i * hw + p] = v;
    }
    if (pot_out) pot_out[p] = pot;
  }
}

/* TestH5Dataset.make_voxel, data/testh5.py:60-90.  ts in seconds (float64, or float32 when ts_f32), xs/ys int64,
 * ps in {0,1}; vox [bins,H,W] zero-filled by the caller. */
void orc_make_voxel(const void* ts, int ts_f32, const int64_t* xs, const int64_t* ys, const uint8_t* ps, int64_t ne,
                    int bins, int h, int w, int interpolate, double* vox) {
  if (ne == 0) return;                                                        /* :63-64 */
  const double* td = (const double*)ts;
  const float* tf = (const float*)ts;
#define TAU(e) (ts_f32 ? (int64_t)((float)(tf[e] - tf[0]) * 1e6f) : (int64_t)((td[e] - td[0]) * 1e6))   /* :68 */
  const int64_t tl = TAU(ne - 1);
  if (!interpolate) {
    const double tpb = ((double)tl + 0.001) / bins;                           /* :71 */
    for (int64_t e = 0; e < ne; ++e) {
      const int b = (int)(uint8_t)floor((double)TAU(e) / tpb);                /* :72 */
      vox[((int64_t)b * h + ys[e]) * w + xs[e]] += (double)(2 * (int)ps[e] - 1);   /* :67,73 */
    }
  } else {
    const double den = (double)(tl - 0) + 0.0001;                             /* :76-77 */
    for (int bi = 0; bi < bins; ++bi) {                                       /* :78-80 */
      for (int64_t e = 0; e < ne; ++e) {
        const double tn = (double)TAU(e) / den * (bins - 1);
        const double wgt = fmax(0.0, 1.0 - fabs(tn - bi));
        vox[((int64_t)bi * h + ys[e]) * w + xs[e]] += wgt * (double)(2 * (int)ps[e] - 1);
      }
    }
  }
#undef TAU
}

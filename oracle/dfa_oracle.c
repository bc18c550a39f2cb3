/*
 * dfa_oracle.c — CPU restatement of SimPB's deformable_aggregation op.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (simpb_b200/) may link,
 * import or call this file; it is the checker used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline leg.
 *
 * What it restates (all paths relative to /root/reference):
 *   forward  : projects/mmdet3d_plugin/ops/src/deformable_aggregation_cuda.cu:13-59
 *              (bilinear_sampling) and :129-187 (deformable_aggregation_kernel)
 *   backward : same file :62-126 (bilinear_sampling_grad) and :190-262
 *   layouts  : projects/mmdet3d_plugin/ops/src/deformable_aggregation.cpp:22-28
 *
 * Arithmetic contract
 *   - Geometry (validity test, pixel coordinate, floor, corner rows, in-bounds
 *     flags, fractional weights lh/lw/hh/hw) is evaluated in IEEE binary32 exactly
 *     as the compiled reference does.  The reference source writes
 *     `loc_h * h - 0.5` (…_cuda.cu:180-181); nvcc's default -fmad=true contracts it
 *     into ONE fused multiply-add (FFMA loc, (float)h, -0.5) — `fma_mode = 1`
 *     below.  `fma_mode = 0` is the two-rounding reading of the source text.
 *   - Everything after the geometry (products with feature values, the 312-way sum)
 *     is carried in binary64, because the reference accumulates through float
 *     atomics in a run-to-run varying order; binary64 is the tie-breaker against
 *     which the reference op, the grid_sample fallback and the new kernels are each
 *     compared with the 1e-5 (fp32) / 1e-2 (bf16 features) relative tolerance.
 *   - Integer side channel for the bit-exact checks: valid[b,a,p,k] (uint8) and
 *     corner_rows[b,a,p,k,l,4] (int32 row index into num_feat in the order
 *     (h_low,w_low),(h_low,w_high),(h_high,w_low),(h_high,w_high); -1 when the
 *     corner is out of bounds or the sample is invalid).
 *
 * Layouts (row-major, innermost last):
 *   feat   [bs, num_feat, C]            float
 *   shape  [K, L, 2]  (H, W)            int32
 *   start  [K, L]                       int32
 *   loc    [bs, A, P, K, 2] (x, y)      float
 *   w      [bs, A, P, K, L, G]          float
 *   out    [bs, A, C]                   double
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int valid;        /* sample passes 0<x<1 && 0<y<1 (exclusive)                */
  int h_low, w_low; /* floor of the pixel coordinate                           */
  float lh, lw, hh, hw;
  int row[4];       /* row index inside one batch item, -1 if out of bounds    */
} tap_geom;

/* …_cuda.cu:168-171 (mask), :180-181 (pixel coordinate), :18-25 (floor/fractions),
 * :33-52 (in-bounds tests). */
static void geometry(float loc_w, float loc_h, int H, int W, int start, int fma_mode,
                     tap_geom *g) {
  g->valid = !(loc_w <= 0.0f || loc_w >= 1.0f) && !(loc_h <= 0.0f || loc_h >= 1.0f);
  float h_im, w_im;
  if (fma_mode) {
    h_im = fmaf(loc_h, (float)H, -0.5f);
    w_im = fmaf(loc_w, (float)W, -0.5f);
  } else {
    volatile float th = loc_h * (float)H, tw = loc_w * (float)W;
    /* source text: float product, then `- 0.5` in double, rounded back to float */
    h_im = (float)((double)th - 0.5);
    w_im = (float)((double)tw - 0.5);
  }
  const float fh = floorf(h_im), fw = floorf(w_im);
  g->h_low = (int)fh;
  g->w_low = (int)fw;
  const int h_high = g->h_low + 1, w_high = g->w_low + 1;
  g->lh = h_im - (float)g->h_low;
  g->lw = w_im - (float)g->w_low;
  g->hh = 1.0f - g->lh;
  g->hw = 1.0f - g->lw;
  const int ok_hl = g->h_low >= 0, ok_wl = g->w_low >= 0;
  const int ok_hh = h_high <= H - 1, ok_wh = w_high <= W - 1;
  g->row[0] = (ok_hl && ok_wl) ? start + g->h_low * W + g->w_low : -1;
  g->row[1] = (ok_hl && ok_wh) ? start + g->h_low * W + w_high : -1;
  g->row[2] = (ok_hh && ok_wl) ? start + h_high * W + g->w_low : -1;
  g->row[3] = (ok_hh && ok_wh) ? start + h_high * W + w_high : -1;
  if (!g->valid) g->row[0] = g->row[1] = g->row[2] = g->row[3] = -1;
}

/* Forward.  `valid_out` / `rows_out` may be NULL. */
int dfa_oracle_forward(const float *feat, const int32_t *shape, const int32_t *start,
                       const float *loc, const float *w, double *out, uint8_t *valid_out,
                       int32_t *rows_out, int bs, int num_feat, int C, int K, int L, int A,
                       int P, int G, int fma_mode) {
  if (G <= 0 || C % G != 0) return 1;
  const int cpg = C / G;
  memset(out, 0, sizeof(double) * (size_t)bs * A * C);
  for (int b = 0; b < bs; ++b)
    for (int a = 0; a < A; ++a) {
      double *o = out + ((size_t)b * A + a) * C;
      for (int p = 0; p < P; ++p)
        for (int k = 0; k < K; ++k) {
          const size_t s = (((size_t)b * A + a) * P + p) * K + k;
          const float x = loc[2 * s], y = loc[2 * s + 1];
          for (int l = 0; l < L; ++l) {
            const int H = shape[(k * L + l) * 2], W = shape[(k * L + l) * 2 + 1];
            tap_geom g;
            geometry(x, y, H, W, start[k * L + l], fma_mode, &g);
            if (l == 0 && valid_out) valid_out[s] = (uint8_t)g.valid;
            if (rows_out) memcpy(rows_out + (s * L + l) * 4, g.row, sizeof(g.row));
            if (!g.valid) continue;
            const double cw[4] = {(double)g.hh * g.hw, (double)g.hh * g.lw,
                                  (double)g.lh * g.hw, (double)g.lh * g.lw};
            const float *wp = w + (s * L + l) * G;
            for (int q = 0; q < 4; ++q) {
              if (g.row[q] < 0) continue;
              const float *v = feat + ((size_t)b * num_feat + g.row[q]) * C;
              for (int c = 0; c < C; ++c) o[c] += cw[q] * (double)v[c] * (double)wp[c / cpg];
            }
          }
        }
    }
  return 0;
}

/* Backward (…_cuda.cu:62-126, :237-261).  All three gradients in binary64,
 * written (not accumulated).  Any of the outputs may be NULL. */
int dfa_oracle_backward(const float *feat, const int32_t *shape, const int32_t *start,
                        const float *loc, const float *w, const float *grad_out,
                        double *grad_feat, double *grad_loc, double *grad_w, int bs,
                        int num_feat, int C, int K, int L, int A, int P, int G, int fma_mode) {
  if (G <= 0 || C % G != 0) return 1;
  const int cpg = C / G;
  if (grad_feat) memset(grad_feat, 0, sizeof(double) * (size_t)bs * num_feat * C);
  if (grad_loc) memset(grad_loc, 0, sizeof(double) * (size_t)bs * A * P * K * 2);
  if (grad_w) memset(grad_w, 0, sizeof(double) * (size_t)bs * A * P * K * L * G);
  for (int b = 0; b < bs; ++b)
    for (int a = 0; a < A; ++a) {
      const float *go = grad_out + ((size_t)b * A + a) * C;
      for (int p = 0; p < P; ++p)
        for (int k = 0; k < K; ++k) {
          const size_t s = (((size_t)b * A + a) * P + p) * K + k;
          const float x = loc[2 * s], y = loc[2 * s + 1];
          for (int l = 0; l < L; ++l) {
            const int H = shape[(k * L + l) * 2], W = shape[(k * L + l) * 2 + 1];
            tap_geom g;
            geometry(x, y, H, W, start[k * L + l], fma_mode, &g);
            if (!g.valid) continue;
            const double hh = g.hh, hw = g.hw, lh = g.lh, lw = g.lw;
            const double cw[4] = {hh * hw, hh * lw, lh * hw, lh * lw};
            /* d(val)/d(h_im) and d(val)/d(w_im) coefficients per corner (:91-118) */
            const double ch[4] = {-hw, -lw, hw, lw};
            const double cx[4] = {-hh, hh, -lh, lh};
            const float *wp = w + (s * L + l) * G;
            for (int c = 0; c < C; ++c) {
              const double gr = go[c], wt = wp[c / cpg], t = gr * wt;
              double val = 0, gh = 0, gw_ = 0;
              for (int q = 0; q < 4; ++q) {
                if (g.row[q] < 0) continue;
                const size_t fi = ((size_t)b * num_feat + g.row[q]) * C + c;
                const double v = feat[fi];
                val += cw[q] * v;
                gh += ch[q] * v;
                gw_ += cx[q] * v;
                if (grad_feat) grad_feat[fi] += cw[q] * t;
              }
              if (grad_w) grad_w[(s * L + l) * G + c / cpg] += gr * val;
              if (grad_loc) {
                grad_loc[2 * s] += (double)W * gw_ * t;
                grad_loc[2 * s + 1] += (double)H * gh * t;
              }
            }
          }
        }
    }
  return 0;
}

/* Number of distinct feature rows (b,row) referenced by at least one in-bounds corner
 * of a valid sample — the `U` of SURVEY.md §8(d)'s algorithmic-byte formula. */
long long dfa_oracle_distinct_rows(const int32_t *shape, const int32_t *start, const float *loc,
                                   int bs, int num_feat, int K, int L, int A, int P,
                                   int fma_mode) {
  uint8_t *seen = (uint8_t *)calloc((size_t)bs * num_feat, 1);
  if (!seen) return -1;
  long long n = 0;
  for (int b = 0; b < bs; ++b)
    for (size_t s = (size_t)b * A * P * K; s < (size_t)(b + 1) * A * P * K; ++s) {
      const int k = (int)(s % K);
      for (int l = 0; l < L; ++l) {
        tap_geom g;
        geometry(loc[2 * s], loc[2 * s + 1], shape[(k * L + l) * 2], shape[(k * L + l) * 2 + 1],
                 start[k * L + l], fma_mode, &g);
        for (int q = 0; q < 4; ++q)
          if (g.row[q] >= 0 && !seen[(size_t)b * num_feat + g.row[q]]) {
            seen[(size_t)b * num_feat + g.row[q]] = 1;
            ++n;
          }
      }
    }
  free(seen);
  return n;
}

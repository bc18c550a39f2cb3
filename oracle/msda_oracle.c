/*
 * msda_oracle.c — CPU restatement of multi-scale deformable attention as SimPB's 2-D decoder calls
 * it (TEST INFRASTRUCTURE ONLY — see oracle/__init__.py).
 *
 * Call site in the reference: projects/mmdet3d_plugin/models/group_attn.py:229-233,
 *   MultiScaleDeformableAttnFunction.apply(value[:, i], spatial_shapes, level_start_index,
 *                                          sampling_locations[:, q0:q1], attention_weights[:, q0:q1],
 *                                          im2col_step)
 * The function itself lives in a third-party dependency that is NOT vendored under /root/reference:
 * mmcv-full 1.7.1 (requirement.txt:2), mmcv/ops/multi_scale_deform_attn.py +
 * mmcv/ops/csrc/common/cuda/ms_deform_attn_cuda_kernel.cuh.  PARITY UNPINNED by reference vectors
 * (mmcv cannot be installed here); this file restates mmcv's published algorithm:
 *
 *   out[b,q,m*D+d] = sum_{l,p} w[b,q,m,l,p] * bilinear(value[b, start_l + ., m, d], h_im, w_im)
 *       h_im = loc_y * H_l - 0.5,  w_im = loc_x * W_l - 0.5          (loc[...,0] = x, loc[...,1] = y)
 *       taken only if h_im > -1 && w_im > -1 && h_im < H_l && w_im < W_l
 *       bilinear with zero padding: corner (h_low, w_low) needs h_low >= 0 && w_low >= 0, the
 *       high corners need h_high <= H-1 / w_high <= W-1
 *   backward (col2im): with t = grad_out * w,
 *       grad_value[corner_i]   += bilinear_weight_i * t
 *       grad_w                  = sum_d grad_out * val
 *       grad_loc.x              = W * sum_d t * (-hh v1 + hh v2 - lh v3 + lh v4)
 *       grad_loc.y              = H * sum_d t * (-hw v1 - lw v2 + hw v3 + lw v4)
 * and is cross-checked in tests/ against the same formulation written with torch's grid_sample
 * (mmcv's own CPU fallback `multi_scale_deformable_attn_pytorch`, restated in oracle/msda_ref.py).
 * The pixel coordinate is evaluated in binary32 (fma_mode 1: one fused multiply-add, as nvcc emits
 * for the analogous expression of SimPB's own op; 0: product rounded, then the subtraction), the
 * sums in binary64.
 *
 * Layouts: value [bs, S, M, D], shapes [L,2] (H,W), start [L], loc [bs,Q,M,L,P,2], w [bs,Q,M,L,P],
 *          out [bs, Q, M*D] (double).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct {
  int valid;
  float lh, lw, hh, hw;
  int idx[4]; /* position inside the level (h*W + w), -1 = out of bounds */
} msda_geom;

static void msda_geometry(float loc_w, float loc_h, int H, int W, int fma_mode, msda_geom *g) {
  float h_im, w_im;
  if (fma_mode) {
    h_im = fmaf(loc_h, (float)H, -0.5f);
    w_im = fmaf(loc_w, (float)W, -0.5f);
  } else {
    volatile float th = loc_h * (float)H, tw = loc_w * (float)W;
    h_im = (float)((double)th - 0.5);
    w_im = (float)((double)tw - 0.5);
  }
  g->valid = h_im > -1.0f && w_im > -1.0f && h_im < (float)H && w_im < (float)W;
  const int h_low = (int)floorf(h_im), w_low = (int)floorf(w_im);
  const int h_high = h_low + 1, w_high = w_low + 1;
  g->lh = h_im - (float)h_low;
  g->lw = w_im - (float)w_low;
  g->hh = 1.0f - g->lh;
  g->hw = 1.0f - g->lw;
  const int hl = h_low >= 0, wl = w_low >= 0, hh = h_high <= H - 1, wh = w_high <= W - 1;
  g->idx[0] = (g->valid && hl && wl) ? h_low * W + w_low : -1;
  g->idx[1] = (g->valid && hl && wh) ? h_low * W + w_high : -1;
  g->idx[2] = (g->valid && hh && wl) ? h_high * W + w_low : -1;
  g->idx[3] = (g->valid && hh && wh) ? h_high * W + w_high : -1;
}

int msda_oracle_forward(const float *value, const int32_t *shapes, const int32_t *start,
                        const float *loc, const float *w, double *out, int bs, int S, int M, int D,
                        int Q, int L, int P, int fma_mode) {
  for (long long bq = 0; bq < (long long)bs * Q; ++bq) {
    const int b = (int)(bq / Q);
    for (int m = 0; m < M; ++m) {
      double *o = out + (bq * M + m) * D;
      for (int d = 0; d < D; ++d) o[d] = 0.0;
      for (int l = 0; l < L; ++l) {
        const int H = shapes[2 * l], W = shapes[2 * l + 1];
        for (int p = 0; p < P; ++p) {
          const long long t = ((bq * M + m) * L + l) * P + p;
          msda_geom g;
          msda_geometry(loc[2 * t], loc[2 * t + 1], H, W, fma_mode, &g);
          if (!g.valid) continue;
          const double bw[4] = {(double)g.hh * g.hw, (double)g.hh * g.lw, (double)g.lh * g.hw,
                                (double)g.lh * g.lw};
          for (int c = 0; c < 4; ++c) {
            if (g.idx[c] < 0) continue;
            const float *v = value + (((long long)b * S + start[l] + g.idx[c]) * M + m) * D;
            for (int d = 0; d < D; ++d) o[d] += (double)w[t] * bw[c] * (double)v[d];
          }
        }
      }
    }
  }
  return 0;
}

int msda_oracle_backward(const float *value, const int32_t *shapes, const int32_t *start,
                         const float *loc, const float *w, const float *grad_out, double *grad_value,
                         double *grad_loc, double *grad_w, int bs, int S, int M, int D, int Q, int L,
                         int P, int fma_mode) {
  if (grad_value) memset(grad_value, 0, sizeof(double) * (size_t)bs * S * M * D);
  for (long long bq = 0; bq < (long long)bs * Q; ++bq) {
    const int b = (int)(bq / Q);
    for (int m = 0; m < M; ++m) {
      const float *go = grad_out + (bq * M + m) * D;
      for (int l = 0; l < L; ++l) {
        const int H = shapes[2 * l], W = shapes[2 * l + 1];
        for (int p = 0; p < P; ++p) {
          const long long t = ((bq * M + m) * L + l) * P + p;
          grad_w[t] = 0.0, grad_loc[2 * t] = 0.0, grad_loc[2 * t + 1] = 0.0;
          msda_geom g;
          msda_geometry(loc[2 * t], loc[2 * t + 1], H, W, fma_mode, &g);
          if (!g.valid) continue;
          const double bw[4] = {(double)g.hh * g.hw, (double)g.hh * g.lw, (double)g.lh * g.hw,
                                (double)g.lh * g.lw};
          const double cx[4] = {-(double)g.hh, (double)g.hh, -(double)g.lh, (double)g.lh};
          const double cy[4] = {-(double)g.hw, -(double)g.lw, (double)g.hw, (double)g.lw};
          double ga = 0.0, gx = 0.0, gy = 0.0;
          for (int c = 0; c < 4; ++c) {
            if (g.idx[c] < 0) continue;
            const long long vi = (((long long)b * S + start[l] + g.idx[c]) * M + m) * D;
            for (int d = 0; d < D; ++d) {
              const double gd = go[d], v = value[vi + d];
              ga += gd * bw[c] * v;
              gx += gd * cx[c] * v;
              gy += gd * cy[c] * v;
              if (grad_value) grad_value[vi + d] += bw[c] * gd * (double)w[t];
            }
          }
          grad_w[t] = ga;
          grad_loc[2 * t] = (double)W * (double)w[t] * gx;
          grad_loc[2 * t + 1] = (double)H * (double)w[t] * gy;
        }
      }
    }
  }
  return 0;
}

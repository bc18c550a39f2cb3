// dfa_msda.cu — multi-scale deformable attention for SimPB's 2-D query branch (the decoder's other
// gather: models/group_attn.py:229-233 calls mmcv's MultiScaleDeformableAttnFunction once per camera
// group on the SAME channel-last feature table the 3-D branch uses) and its C ABI.
//
// Semantics: mmcv-full 1.7.1, mmcv/ops/csrc/common/cuda/ms_deform_attn_cuda_kernel.cuh
// (ms_deformable_im2col / col2im):
//   out[b,q,m*D+d] = sum_{l,p} w[b,q,m,l,p] * bilinear(value[b, start_l + ., m, d]; loc[b,q,m,l,p])
// with h_im = y*H - 0.5, w_im = x*W - 0.5, a tap taken iff h_im > -1 && w_im > -1 && h_im < H &&
// w_im < W, zero padding outside the map.  Unlike SimPB's own op the location differs per head.
//
// One CTA owns one query (b, q); warp m owns head m.  The query's sampling locations and attention
// weights (M*L*P*12 bytes, contiguous) arrive by two TMA bulk copies.  Lanes of a warp are split as
// [sub-tap][corner][16-byte vector of the head's channels], so one 128-bit load per lane fetches the
// head's slice of all four corners; the weighted sum stays in registers and is written once (mmcv:
// one thread per output channel, scalar loads).  In the backward every tap's grad_attn_weight and
// grad_sampling_loc are produced by exactly one warp through shuffles and written once — no shared
// memory reductions, no atomics, no zero-filled buffers; grad_value is scattered with
// red.global.add.v4.f32.
#include "dfa_common.cuh"

namespace {

struct MsdaDims {
  int bs, S, M, D, Q, L, P;
  int K;  // value tables per batch item (camera groups); query q reads table qcam[q]
};

// first element of the value table query (b, q) samples
__device__ __forceinline__ size_t msda_table(const MsdaDims &d, const int *qcam, int b, int q) {
  const int cam = qcam ? __ldg(qcam + q) : 0;
  return (static_cast<size_t>(b) * d.K + cam) * d.S * d.M * d.D;
}

struct MsdaGeom {
  int idx[4];  // position inside the level (h*W + w), -1 when the corner is outside or the tap is not taken
  float lh, lw, hh, hw;
};

__device__ __forceinline__ void msda_geometry(float x, float y, int H, int W, MsdaGeom &g) {
  const float h_im = fmaf(y, static_cast<float>(H), -0.5f);
  const float w_im = fmaf(x, static_cast<float>(W), -0.5f);
  const bool ok = h_im > -1.f && w_im > -1.f && h_im < static_cast<float>(H) && w_im < static_cast<float>(W);
  const float fh = floorf(h_im), fw = floorf(w_im);
  const int h_low = static_cast<int>(fh), w_low = static_cast<int>(fw);
  g.lh = h_im - fh, g.lw = w_im - fw, g.hh = 1.f - g.lh, g.hw = 1.f - g.lw;
  const bool hl = h_low >= 0, wl = w_low >= 0, hh = h_low + 1 <= H - 1, wh = w_low + 1 <= W - 1;
  const int base = h_low * W + w_low;
  g.idx[0] = (ok && hl && wl) ? base : -1;
  g.idx[1] = (ok && hl && wh) ? base + 1 : -1;
  g.idx[2] = (ok && hh && wl) ? base + W : -1;
  g.idx[3] = (ok && hh && wh) ? base + W + 1 : -1;
}

struct MsdaRec {  // one per (tap, corner)
  int off;        // element offset of the corner's head slice inside the batch item, -1 = skip
  float a, b, c;  // forward: a = bilinear weight x attention weight
                  // backward: a = bilinear weight, b = d/dx coefficient x W, c = d/dy coefficient x H
};

// Stage loc / w of one query into shared memory (TMA when aligned), all warps return after the data
// is visible.
template <bool TMA>
__device__ __forceinline__ void msda_stage(const float *loc_g, const float *w_g, float *s_loc, float *s_w,
                                           uint64_t *bar, int n_tap) {
  const int tid = threadIdx.x;
  if (TMA) {
    if (tid == 0) {
      mbar_init(bar, 1);
      fence_mbar_init();
      mbar_expect_tx(bar, 12u * n_tap);
      tma_bulk_g2s(s_loc, loc_g, 8u * n_tap, bar);
      tma_bulk_g2s(s_w, w_g, 4u * n_tap, bar);
    }
    __syncthreads();
    mbar_wait(bar, 0);
  } else {
    for (int i = tid; i < 2 * n_tap; i += blockDim.x) s_loc[i] = __ldg(loc_g + i);
    for (int i = tid; i < n_tap; i += blockDim.x) s_w[i] = __ldg(w_g + i);
    __syncthreads();
  }
}

// T: value type.  LPG: lanes per corner = D*sizeof(T)/16.  U: taps in flight per lane group.
template <typename T, int LPG, int U, bool TMA>
__global__ void __launch_bounds__(1024)
    msda_fwd_kernel(const T *__restrict__ value, const int *__restrict__ shapes,
                    const int *__restrict__ start, const float *__restrict__ loc,
                    const float *__restrict__ w, float *__restrict__ out, MsdaDims d,
                    const int *__restrict__ qcam) {
  constexpr int VEC = FeatVec<T>::VEC;
  constexpr int TPW = 32 / (4 * LPG);
  extern __shared__ __align__(128) unsigned char smem[];
  const int LP = d.L * d.P, n_tap = d.M * LP;
  const int lp_pad = (LP + TPW * U - 1) / (TPW * U) * (TPW * U);
  float *s_loc = reinterpret_cast<float *>(smem);
  float *s_w = s_loc + 2 * n_tap;
  MsdaRec *s_rec = reinterpret_cast<MsdaRec *>(smem + align_up(12u * n_tap, 16));
  uint64_t *bar = reinterpret_cast<uint64_t *>(s_rec + static_cast<size_t>(d.M) * lp_pad * 4);
  const int lane = threadIdx.x & 31, m = threadIdx.x >> 5;
  const long long bq = blockIdx.x;
  const int b = static_cast<int>(bq / d.Q);
  msda_stage<TMA>(loc + bq * n_tap * 2, w + bq * n_tap, s_loc, s_w, bar, n_tap);

  MsdaRec *rec = s_rec + static_cast<size_t>(m) * lp_pad * 4;
  for (int t = lane; t < lp_pad; t += 32) {
    MsdaRec r[4] = {{-1, 0.f, 0.f, 0.f}, {-1, 0.f, 0.f, 0.f}, {-1, 0.f, 0.f, 0.f}, {-1, 0.f, 0.f, 0.f}};
    if (t < LP) {
      const int l = t / d.P, i = m * LP + t;
      const int H = __ldg(shapes + 2 * l), W = __ldg(shapes + 2 * l + 1), s0 = __ldg(start + l);
      MsdaGeom g;
      msda_geometry(s_loc[2 * i], s_loc[2 * i + 1], H, W, g);
      const float aw = s_w[i];
      const float bw[4] = {g.hh * g.hw, g.hh * g.lw, g.lh * g.hw, g.lh * g.lw};
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (g.idx[c] >= 0) r[c].off = ((s0 + g.idx[c]) * d.M + m) * d.D, r[c].a = bw[c] * aw;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) rec[4 * t + c] = r[c];
  }
  __syncwarp();

  const int j = lane % LPG, q = (lane / LPG) & 3, sub = lane / (4 * LPG);
  const T *vb = value + msda_table(d, qcam, b, static_cast<int>(bq - static_cast<long long>(b) * d.Q)) + j * VEC;
  float acc[VEC];
#pragma unroll
  for (int c = 0; c < VEC; ++c) acc[c] = 0.f;
  for (int t0 = 0; t0 < lp_pad; t0 += TPW * U) {
    typename FeatVec<T>::raw_t v[U];
    float cw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const MsdaRec r = rec[4 * (t0 + u * TPW + sub) + q];
      cw[u] = r.a;
      v[u] = r.off >= 0 ? FeatVec<T>::load_raw(vb + r.off) : FeatVec<T>::zero_raw();
    }
#pragma unroll
    for (int u = 0; u < U; ++u) FeatVec<T>::fma(acc, cw[u], v[u]);
  }
#pragma unroll
  for (int mk = LPG; mk < 32; mk <<= 1)
#pragma unroll
    for (int c = 0; c < VEC; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], mk);
  if (lane < LPG) {
    float4 *o = reinterpret_cast<float4 *>(out + (bq * d.M + m) * d.D + j * VEC);
#pragma unroll
    for (int c = 0; c < VEC / 4; ++c)
      o[c] = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
  }
}

template <typename T, int LPG, int U, bool TMA>
__global__ void __launch_bounds__(1024)
    msda_bwd_kernel(const T *__restrict__ value, const int *__restrict__ shapes,
                    const int *__restrict__ start, const float *__restrict__ loc,
                    const float *__restrict__ w, const float *__restrict__ grad_out,
                    float *__restrict__ grad_value, float *__restrict__ grad_loc,
                    float *__restrict__ grad_w, MsdaDims d, const int *__restrict__ qcam) {
  constexpr int VEC = FeatVec<T>::VEC;
  constexpr int TPW = 32 / (4 * LPG);
  extern __shared__ __align__(128) unsigned char smem[];
  const int LP = d.L * d.P, n_tap = d.M * LP;
  const int lp_pad = (LP + TPW * U - 1) / (TPW * U) * (TPW * U);
  float *s_loc = reinterpret_cast<float *>(smem);
  float *s_w = s_loc + 2 * n_tap;
  MsdaRec *s_rec = reinterpret_cast<MsdaRec *>(smem + align_up(12u * n_tap, 16));
  uint64_t *bar = reinterpret_cast<uint64_t *>(s_rec + static_cast<size_t>(d.M) * lp_pad * 4);
  const int lane = threadIdx.x & 31, m = threadIdx.x >> 5;
  const long long bq = blockIdx.x;
  const int b = static_cast<int>(bq / d.Q);
  msda_stage<TMA>(loc + bq * n_tap * 2, w + bq * n_tap, s_loc, s_w, bar, n_tap);

  MsdaRec *rec = s_rec + static_cast<size_t>(m) * lp_pad * 4;
  for (int t = lane; t < lp_pad; t += 32) {
    MsdaRec r[4] = {{-1, 0.f, 0.f, 0.f}, {-1, 0.f, 0.f, 0.f}, {-1, 0.f, 0.f, 0.f}, {-1, 0.f, 0.f, 0.f}};
    if (t < LP) {
      const int l = t / d.P, i = m * LP + t;
      const int H = __ldg(shapes + 2 * l), W = __ldg(shapes + 2 * l + 1), s0 = __ldg(start + l);
      MsdaGeom g;
      msda_geometry(s_loc[2 * i], s_loc[2 * i + 1], H, W, g);
      const float Wf = static_cast<float>(W), Hf = static_cast<float>(H);
      const float bw[4] = {g.hh * g.hw, g.hh * g.lw, g.lh * g.hw, g.lh * g.lw};
      const float cx[4] = {-g.hh * Wf, g.hh * Wf, -g.lh * Wf, g.lh * Wf};
      const float cy[4] = {-g.hw * Hf, -g.lw * Hf, g.hw * Hf, g.lw * Hf};
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (g.idx[c] >= 0)
          r[c].off = ((s0 + g.idx[c]) * d.M + m) * d.D, r[c].a = bw[c], r[c].b = cx[c], r[c].c = cy[c];
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) rec[4 * t + c] = r[c];
  }
  __syncwarp();

  const int j = lane % LPG, q = (lane / LPG) & 3, sub = lane / (4 * LPG);
  const size_t vbase = msda_table(d, qcam, b, static_cast<int>(bq - static_cast<long long>(b) * d.Q)) + j * VEC;
  const T *vb = value + vbase;
  float *gvb = grad_value ? grad_value + vbase : nullptr;
  float go[VEC];
  {
    const float4 *p = reinterpret_cast<const float4 *>(grad_out + (bq * d.M + m) * d.D + j * VEC);
#pragma unroll
    for (int c = 0; c < VEC / 4; ++c) {
      const float4 t = __ldg(p + c);
      go[4 * c] = t.x, go[4 * c + 1] = t.y, go[4 * c + 2] = t.z, go[4 * c + 3] = t.w;
    }
  }
  float *gw_q = grad_w + (bq * d.M + m) * LP;
  float *gl_q = grad_loc + (bq * d.M + m) * LP * 2;
  for (int t0 = 0; t0 < lp_pad; t0 += TPW * U) {
    float dd[U];
    MsdaRec r[U];
    float aw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int t = t0 + u * TPW + sub;
      r[u] = rec[4 * t + q];
      aw[u] = t < LP ? s_w[m * LP + t] : 0.f;
      float s = 0.f;
      if (r[u].off >= 0) {
        float v[VEC];
        FeatVec<T>::load(vb + r[u].off, v);
#pragma unroll
        for (int c = 0; c < VEC; ++c) s = fmaf(go[c], v[c], s);
        if (gvb) {
          const float coef = r[u].a * aw[u];
#pragma unroll
          for (int c = 0; c < VEC; c += 4)
            red_add_v4(gvb + r[u].off + c, coef * go[c], coef * go[c + 1], coef * go[c + 2],
                       coef * go[c + 3]);
        }
      }
      dd[u] = s;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float s = dd[u];
#pragma unroll
      for (int mk = 1; mk < LPG; mk <<= 1) s += __shfl_xor_sync(0xffffffffu, s, mk);
      float pa = r[u].a * s, px = r[u].b * s, py = r[u].c * s;
#pragma unroll
      for (int mk = LPG; mk < 4 * LPG; mk <<= 1) {
        pa += __shfl_xor_sync(0xffffffffu, pa, mk);
        px += __shfl_xor_sync(0xffffffffu, px, mk);
        py += __shfl_xor_sync(0xffffffffu, py, mk);
      }
      const int t = t0 + u * TPW + sub;
      if (j == 0 && q == 0 && t < LP) {  // every tap is written, taken or not: no memset needed
        gw_q[t] = pa;
        *reinterpret_cast<float2 *>(gl_q + 2 * t) = make_float2(px * aw[u], py * aw[u]);
      }
    }
  }
}

// Shape-generic kernels (any D, any alignment): one warp per (b, q, m), lanes stride over channels.
template <typename T>
__global__ void __launch_bounds__(256)
    msda_fwd_generic_kernel(const T *__restrict__ value, const int *__restrict__ shapes,
                            const int *__restrict__ start, const float *__restrict__ loc,
                            const float *__restrict__ w, float *__restrict__ out, MsdaDims d,
                            const int *__restrict__ qcam) {
  const long long wid = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= static_cast<long long>(d.bs) * d.Q * d.M) return;
  const int m = static_cast<int>(wid % d.M);
  const int b = static_cast<int>(wid / d.M / d.Q);
  const int LP = d.L * d.P;
  const T *vb = value + msda_table(d, qcam, b, static_cast<int>((wid / d.M) % d.Q));
  for (int c0 = 0; c0 < d.D; c0 += 32) {
    const int c = c0 + lane;
    float acc = 0.f;
    for (int t = 0; t < LP; ++t) {
      const int l = t / d.P;
      MsdaGeom g;
      msda_geometry(loc[(wid * LP + t) * 2], loc[(wid * LP + t) * 2 + 1], shapes[2 * l], shapes[2 * l + 1], g);
      const float bw[4] = {g.hh * g.hw, g.hh * g.lw, g.lh * g.hw, g.lh * g.lw};
      float val = 0.f;
      if (c < d.D)
        for (int k = 0; k < 4; ++k)
          if (g.idx[k] >= 0)
            val = fmaf(bw[k], static_cast<float>(vb[(static_cast<size_t>(start[l] + g.idx[k]) * d.M + m) * d.D + c]), val);
      acc = fmaf(w[wid * LP + t], val, acc);
    }
    if (c < d.D) out[wid * d.D + c] = acc;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
    msda_bwd_generic_kernel(const T *__restrict__ value, const int *__restrict__ shapes,
                            const int *__restrict__ start, const float *__restrict__ loc,
                            const float *__restrict__ w, const float *__restrict__ grad_out,
                            float *__restrict__ grad_value, float *__restrict__ grad_loc,
                            float *__restrict__ grad_w, MsdaDims d, const int *__restrict__ qcam) {
  const long long wid = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= static_cast<long long>(d.bs) * d.Q * d.M) return;
  const int m = static_cast<int>(wid % d.M);
  const int b = static_cast<int>(wid / d.M / d.Q);
  const int LP = d.L * d.P;
  const size_t vbase = msda_table(d, qcam, b, static_cast<int>((wid / d.M) % d.Q));
  for (int t = 0; t < LP; ++t) {
    const int l = t / d.P;
    const int H = shapes[2 * l], W = shapes[2 * l + 1];
    MsdaGeom g;
    msda_geometry(loc[(wid * LP + t) * 2], loc[(wid * LP + t) * 2 + 1], H, W, g);
    const float aw = w[wid * LP + t];
    const float bw[4] = {g.hh * g.hw, g.hh * g.lw, g.lh * g.hw, g.lh * g.lw};
    const float cx[4] = {-g.hh, g.hh, -g.lh, g.lh}, cy[4] = {-g.hw, -g.lw, g.hw, g.lw};
    float pa = 0.f, px = 0.f, py = 0.f;
    for (int c = lane; c < d.D; c += 32) {
      const float gr = grad_out[wid * d.D + c];
      for (int k = 0; k < 4; ++k) {
        if (g.idx[k] < 0) continue;
        const size_t vi = vbase + (static_cast<size_t>(start[l] + g.idx[k]) * d.M + m) * d.D + c;
        const float v = static_cast<float>(value[vi]);
        pa = fmaf(bw[k] * gr, v, pa), px = fmaf(cx[k] * gr, v, px), py = fmaf(cy[k] * gr, v, py);
        if (grad_value) atomicAdd(grad_value + vi, bw[k] * aw * gr);
      }
    }
    for (int mk = 16; mk > 0; mk >>= 1) {
      pa += __shfl_xor_sync(0xffffffffu, pa, mk);
      px += __shfl_xor_sync(0xffffffffu, px, mk);
      py += __shfl_xor_sync(0xffffffffu, py, mk);
    }
    if (lane == 0) {
      grad_w[wid * LP + t] = pa;
      grad_loc[(wid * LP + t) * 2] = px * aw * static_cast<float>(W);
      grad_loc[(wid * LP + t) * 2 + 1] = py * aw * static_cast<float>(H);
    }
  }
}

int msda_check(int bs, int S, int M, int D, int Q, int L, int P, int K, MsdaDims &d) {
  if (bs <= 0 || S <= 0 || M <= 0 || D <= 0 || Q <= 0 || L <= 0 || P <= 0 || K <= 0) return DFA_ERR_BAD_DIMS;
  if (static_cast<long long>(S) * M * D >= (1ll << 31)) return DFA_ERR_BAD_DIMS;  // 32-bit offsets per item
  if (static_cast<long long>(bs) * Q * M >= (1ll << 31)) return DFA_ERR_BAD_DIMS;
  d = MsdaDims{bs, S, M, D, Q, L, P, K};
  return 0;
}

template <typename T>
int msda_lpg(const MsdaDims &d, const void *value) {
  const int bytes = d.D * static_cast<int>(sizeof(T));
  if (bytes % 16 != 0 || !aligned(value, 16) || d.M > 32) return 0;
  const int lpg = bytes / 16;
  return (lpg == 1 || lpg == 2 || lpg == 4 || lpg == 8) ? lpg : 0;
}

inline uint32_t msda_smem(const MsdaDims &d, int tpw_u) {
  const int LP = d.L * d.P, lp_pad = (LP + tpw_u - 1) / tpw_u * tpw_u;
  return align_up(12u * d.M * LP, 16) + 16u * 4u * d.M * lp_pad + 16u;
}

inline bool msda_tma_ok(const MsdaDims &d, const float *loc, const float *w) {
  const long long n = static_cast<long long>(d.M) * d.L * d.P;
  return (4 * n) % 16 == 0 && 12 * n < (1ll << 20) && aligned(loc, 16) && aligned(w, 16);
}

constexpr int MSDA_U = 4;

template <typename T, int LPG>
int msda_launch_fwd(const void *value, const int *shapes, const int *start, const float *loc,
                    const float *w, float *out, const MsdaDims &d, const int *qcam, cudaStream_t st) {
  constexpr int TPW = 32 / (4 * LPG);
  const uint32_t smem = msda_smem(d, TPW * MSDA_U);
  const bool tma = msda_tma_ok(d, loc, w);
  auto kern = tma ? msda_fwd_kernel<T, LPG, MSDA_U, true> : msda_fwd_kernel<T, LPG, MSDA_U, false>;
  if (int rc = set_smem(kern, smem)) return rc;
  kern<<<d.bs * d.Q, 32 * d.M, smem, st>>>(static_cast<const T *>(value), shapes, start, loc, w, out, d,
                                          qcam);
  return static_cast<int>(cudaGetLastError());
}

template <typename T, int LPG>
int msda_launch_bwd(const void *value, const int *shapes, const int *start, const float *loc,
                    const float *w, const float *go, float *gv, float *gl, float *gw, const MsdaDims &d,
                    const int *qcam, cudaStream_t st) {
  constexpr int TPW = 32 / (4 * LPG);
  const uint32_t smem = msda_smem(d, TPW * MSDA_U);
  const bool tma = msda_tma_ok(d, loc, w);
  auto kern = tma ? msda_bwd_kernel<T, LPG, MSDA_U, true> : msda_bwd_kernel<T, LPG, MSDA_U, false>;
  if (int rc = set_smem(kern, smem)) return rc;
  kern<<<d.bs * d.Q, 32 * d.M, smem, st>>>(static_cast<const T *>(value), shapes, start, loc, w, go, gv,
                                          gl, gw, d, qcam);
  return static_cast<int>(cudaGetLastError());
}

template <typename T>
int msda_forward_typed(const void *value, const int *shapes, const int *start, const float *loc,
                       const float *w, float *out, const MsdaDims &d, const int *qcam, cudaStream_t st) {
  const int lpg = msda_lpg<T>(d, value);
  if (lpg && aligned(out, 16) && msda_smem(d, 32) <= 200u * 1024u) {
    switch (lpg) {
      case 8: return msda_launch_fwd<T, 8>(value, shapes, start, loc, w, out, d, qcam, st);
      case 4: return msda_launch_fwd<T, 4>(value, shapes, start, loc, w, out, d, qcam, st);
      case 2: return msda_launch_fwd<T, 2>(value, shapes, start, loc, w, out, d, qcam, st);
      default: return msda_launch_fwd<T, 1>(value, shapes, start, loc, w, out, d, qcam, st);
    }
  }
  const long long warps = static_cast<long long>(d.bs) * d.Q * d.M;
  msda_fwd_generic_kernel<T><<<static_cast<int>((warps + 7) / 8), 256, 0, st>>>(
      static_cast<const T *>(value), shapes, start, loc, w, out, d, qcam);
  return static_cast<int>(cudaGetLastError());
}


// ------------------------------------------------------------------------------------------
// forward on the UNPROJECTED table (inference): gather first, project afterwards
// ------------------------------------------------------------------------------------------
// The module computes value = value_proj(table) for all S rows of every camera (a [bs*K*S, C] x [C, C]
// GEMM: 11.8 GFLOP per layer at SimPB's size, 0.39 ms in fp32) and then samples M*L*P taps per query
// from it.  Sampling is linear, so the projection can move behind it:
//     out[q, m, :] = W_m . ( sum_taps a * bilinear(table) ) + b_m * sum_taps a * (in-map corner weight)
// Here warp m of the query's CTA gathers the WHOLE rows (C channels) of head m's taps from the raw
// table: g[b,q,m,:] (C floats) and s[b,q,m] (the sum that multiplies the bias).  The M small
// [Q, C] x [C, D] products that remain are 126 MFLOP.  The gather moves M times the bytes of the
// projected one — 0.98 GB per layer from an L2-resident table — and still costs a quarter of the GEMM.
// VPL: 16-byte vectors per lane per row (row bytes = 512 * VPL).  Two corners' loads of all taps in
// flight: 4 * VPL loads per lane per tap.
template <typename T, int VPL, bool TMA>
__global__ void __launch_bounds__(1024)
    msda_raw_fwd_kernel(const T *__restrict__ table, const int *__restrict__ shapes,
                        const int *__restrict__ start, const float *__restrict__ loc,
                        const float *__restrict__ w, float *__restrict__ out_g, float *__restrict__ out_s,
                        MsdaDims d, const int *__restrict__ qcam) {
  constexpr int VEC = FeatVec<T>::VEC;
  constexpr int C = 32 * VPL * VEC;
  extern __shared__ __align__(128) unsigned char smem[];
  const int LP = d.L * d.P, n_tap = d.M * LP;
  float *s_loc = reinterpret_cast<float *>(smem);
  float *s_w = s_loc + 2 * n_tap;
  MsdaRec *s_rec = reinterpret_cast<MsdaRec *>(smem + align_up(12u * n_tap, 16));
  uint64_t *bar = reinterpret_cast<uint64_t *>(s_rec + static_cast<size_t>(d.M) * LP * 4);
  const int lane = threadIdx.x & 31, m = threadIdx.x >> 5;
  const long long bq = blockIdx.x;
  const int b = static_cast<int>(bq / d.Q), q = static_cast<int>(bq - static_cast<long long>(b) * d.Q);
  msda_stage<TMA>(loc + bq * n_tap * 2, w + bq * n_tap, s_loc, s_w, bar, n_tap);

  MsdaRec *rec = s_rec + static_cast<size_t>(m) * LP * 4;
  for (int t = lane; t < LP; t += 32) {
    const int l = t / d.P, i = m * LP + t;
    const int H = __ldg(shapes + 2 * l), W = __ldg(shapes + 2 * l + 1), s0 = __ldg(start + l);
    MsdaGeom g;
    msda_geometry(s_loc[2 * i], s_loc[2 * i + 1], H, W, g);
    const float aw = s_w[i];
    const float bw[4] = {g.hh * g.hw, g.hh * g.lw, g.lh * g.hw, g.lh * g.lw};
    // an out-of-map corner (zero padding) replays an in-map corner of the tap with coefficient 0; a tap
    // that is not taken at all replays row 0
    const int safe = g.idx[0] >= 0 ? g.idx[0] : g.idx[1] >= 0 ? g.idx[1] : g.idx[2] >= 0 ? g.idx[2]
                   : g.idx[3] >= 0 ? g.idx[3] : -s0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      MsdaRec r;
      r.off = (s0 + (g.idx[c] >= 0 ? g.idx[c] : safe)) * C;
      r.a = g.idx[c] >= 0 ? bw[c] * aw : 0.f;
      r.b = r.c = 0.f;
      rec[4 * t + c] = r;
    }
  }
  __syncwarp();

  const int cam = qcam ? __ldg(qcam + q) : 0;
  const T *tb = table + (static_cast<size_t>(b) * d.K + cam) * d.S * C + lane * VEC;
  float acc[VPL][VEC];
#pragma unroll
  for (int v = 0; v < VPL; ++v)
#pragma unroll
    for (int c = 0; c < VEC; ++c) acc[v][c] = 0.f;
  float ssum = 0.f;
  for (int t = 0; t < LP; ++t) {
    typename FeatVec<T>::raw_t val[4][VPL];
    float cw[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const MsdaRec r = rec[4 * t + c];
      cw[c] = r.a;
#pragma unroll
      for (int v = 0; v < VPL; ++v) val[c][v] = FeatVec<T>::load_raw(tb + r.off + v * 32 * VEC);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      ssum += cw[c];
#pragma unroll
      for (int v = 0; v < VPL; ++v) FeatVec<T>::fma(acc[v], cw[c], val[c][v]);
    }
  }
  float *og = out_g + (bq * d.M + m) * C + lane * VEC;
#pragma unroll
  for (int v = 0; v < VPL; ++v)
#pragma unroll
    for (int c = 0; c < VEC / 4; ++c)
      reinterpret_cast<float4 *>(og + v * 32 * VEC)[c] =
          make_float4(acc[v][4 * c], acc[v][4 * c + 1], acc[v][4 * c + 2], acc[v][4 * c + 3]);
  if (lane == 0) out_s[bq * d.M + m] = ssum;
}

template <typename T, int VPL>
int msda_launch_raw(const void *table, const int *shapes, const int *start, const float *loc, const float *w,
                    float *out_g, float *out_s, const MsdaDims &d, const int *qcam, cudaStream_t st) {
  const uint32_t smem = align_up(12u * d.M * d.L * d.P, 16) + 16u * 4u * d.M * d.L * d.P + 16u;
  const long long grid = static_cast<long long>(d.bs) * d.Q;
  if (smem > 200u * 1024u) return DFA_ERR_UNSUPPORTED;
  auto go = [&](auto kern) -> int {
    if (int rc = set_smem(kern, smem)) return rc;
    kern<<<static_cast<unsigned int>(grid), 32 * d.M, smem, st>>>(static_cast<const T *>(table), shapes, start, loc, w,
                                                                 out_g, out_s, d, qcam);
    return static_cast<int>(cudaGetLastError());
  };
  return msda_tma_ok(d, loc, w) ? go(msda_raw_fwd_kernel<T, VPL, true>) : go(msda_raw_fwd_kernel<T, VPL, false>);
}

template <typename T>
int msda_backward_typed(const void *value, const int *shapes, const int *start, const float *loc,
                        const float *w, const float *go, float *gv, float *gl, float *gw,
                        const MsdaDims &d, const int *qcam, cudaStream_t st) {
  const int lpg = msda_lpg<T>(d, value);
  if (lpg && aligned(go, 16) && aligned(gv, 16) && aligned(gl, 8) && msda_smem(d, 32) <= 200u * 1024u) {
    switch (lpg) {
      case 8: return msda_launch_bwd<T, 8>(value, shapes, start, loc, w, go, gv, gl, gw, d, qcam, st);
      case 4: return msda_launch_bwd<T, 4>(value, shapes, start, loc, w, go, gv, gl, gw, d, qcam, st);
      case 2: return msda_launch_bwd<T, 2>(value, shapes, start, loc, w, go, gv, gl, gw, d, qcam, st);
      default: return msda_launch_bwd<T, 1>(value, shapes, start, loc, w, go, gv, gl, gw, d, qcam, st);
    }
  }
  const long long warps = static_cast<long long>(d.bs) * d.Q * d.M;
  msda_bwd_generic_kernel<T><<<static_cast<int>((warps + 7) / 8), 256, 0, st>>>(
      static_cast<const T *>(value), shapes, start, loc, w, go, gv, gl, gw, d, qcam);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace

extern "C" {

int dfa_msda_forward(const void *value, int value_dtype, const int32_t *spatial_shapes,
                     const int32_t *level_start_index, const float *sampling_loc,
                     const float *attn_weight, float *output, int bs, int num_value, int num_heads,
                     int head_dim, int num_query, int num_levels, int num_points, int num_tables,
                     const int32_t *query_table, void *stream) {
  if (!value || !spatial_shapes || !level_start_index || !sampling_loc || !attn_weight || !output)
    return DFA_ERR_NULL_POINTER;
  MsdaDims d;
  if (num_tables > 1 && !query_table) return DFA_ERR_NULL_POINTER;
  if (int rc = msda_check(bs, num_value, num_heads, head_dim, num_query, num_levels, num_points,
                          num_tables, d))
    return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (value_dtype == DFA_F32)
    return msda_forward_typed<float>(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                                     output, d, query_table, st);
  if (value_dtype == DFA_BF16)
    return msda_forward_typed<__nv_bfloat16>(value, spatial_shapes, level_start_index, sampling_loc,
                                             attn_weight, output, d, query_table, st);
  return DFA_ERR_BAD_DTYPE;
}

int dfa_msda_forward_raw(const void *table, int table_dtype, const int32_t *spatial_shapes,
                         const int32_t *level_start_index, const float *sampling_loc,
                         const float *attn_weight, float *out_gathered, float *out_weight_sum, int bs,
                         int num_value, int channels, int num_heads, int num_query, int num_levels,
                         int num_points, int num_tables, const int32_t *query_table, void *stream) {
  if (!table || !spatial_shapes || !level_start_index || !sampling_loc || !attn_weight || !out_gathered ||
      !out_weight_sum)
    return DFA_ERR_NULL_POINTER;
  if (num_tables > 1 && !query_table) return DFA_ERR_NULL_POINTER;
  MsdaDims d;
  if (int rc = msda_check(bs, num_value, 1, channels, num_query, num_levels, num_points, num_tables, d)) return rc;
  if (num_heads <= 0 || num_heads > 32) return DFA_ERR_BAD_DIMS;
  if (static_cast<long long>(bs) * num_query * num_heads >= (1ll << 31)) return DFA_ERR_BAD_DIMS;
  d.M = num_heads;  // heads = warps of a query's CTA; the table itself has no head dimension
  d.D = channels;
  if (table_dtype != DFA_F32 && table_dtype != DFA_BF16) return DFA_ERR_BAD_DTYPE;
  const long long rb = static_cast<long long>(channels) * (table_dtype == DFA_BF16 ? 2 : 4);
  if ((rb != 512 && rb != 1024) || !aligned(table, 16) || !aligned(out_gathered, 16)) return DFA_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (table_dtype == DFA_F32)
    return rb == 1024 ? msda_launch_raw<float, 2>(table, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                                                  out_gathered, out_weight_sum, d, query_table, st)
                      : msda_launch_raw<float, 1>(table, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                                                  out_gathered, out_weight_sum, d, query_table, st);
  return rb == 1024 ? msda_launch_raw<__nv_bfloat16, 2>(table, spatial_shapes, level_start_index, sampling_loc,
                                                        attn_weight, out_gathered, out_weight_sum, d, query_table, st)
                    : msda_launch_raw<__nv_bfloat16, 1>(table, spatial_shapes, level_start_index, sampling_loc,
                                                        attn_weight, out_gathered, out_weight_sum, d, query_table, st);
}

int dfa_msda_backward(const void *value, int value_dtype, const int32_t *spatial_shapes,
                      const int32_t *level_start_index, const float *sampling_loc,
                      const float *attn_weight, const float *grad_output, float *grad_value,
                      float *grad_sampling_loc, float *grad_attn_weight, int bs, int num_value,
                      int num_heads, int head_dim, int num_query, int num_levels, int num_points,
                      int num_tables, const int32_t *query_table, int zero_grad_value, void *stream) {
  if (!value || !spatial_shapes || !level_start_index || !sampling_loc || !attn_weight || !grad_output ||
      !grad_sampling_loc || !grad_attn_weight)
    return DFA_ERR_NULL_POINTER;  // grad_value may be NULL: the value gradient is skipped
  MsdaDims d;
  if (num_tables > 1 && !query_table) return DFA_ERR_NULL_POINTER;
  if (int rc = msda_check(bs, num_value, num_heads, head_dim, num_query, num_levels, num_points,
                          num_tables, d))
    return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (zero_grad_value && grad_value) {
    cudaError_t e = cudaMemsetAsync(
        grad_value, 0, sizeof(float) * static_cast<size_t>(bs) * num_tables * num_value * num_heads * head_dim, st);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  if (value_dtype == DFA_F32)
    return msda_backward_typed<float>(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                                      grad_output, grad_value, grad_sampling_loc, grad_attn_weight, d,
                                      query_table, st);
  if (value_dtype == DFA_BF16)
    return msda_backward_typed<__nv_bfloat16>(value, spatial_shapes, level_start_index, sampling_loc,
                                              attn_weight, grad_output, grad_value, grad_sampling_loc,
                                              grad_attn_weight, d, query_table, st);
  return DFA_ERR_BAD_DTYPE;
}

}  // extern "C"

// dfa_forward.cu — forward kernels of the deformable feature aggregation and their C ABI
// (dfa_forward, dfa_forward_host, dfa_debug_indices).  See DESIGN.md §4.1.
//
// One CTA owns one anchor (b, a).  Its sampling locations (P*K*2 floats) and weights (P*K*L*G
// floats) are contiguous per anchor and are staged into shared memory with TMA bulk copies
// (cp.async.bulk -> UBLKCP) completing on mbarriers; warp 0 compacts the samples that pass the op's
// exclusive (0,1) test.  Three kernel families follow, selected by shape (and DFA_FWD_VARIANT):
// row-sliced (default), row-merging, one warp per group; plus a shape-generic fallback.
// The weighted sum lives in registers and the output row is written once: no atomics and no
// zero-filled output (the reference: one float atomic per thread on an at::zeros tensor).
#include <atomic>
#include <mutex>

#include "dfa_common.cuh"

std::atomic<int> dfa_knob_generation{0};

namespace {

#ifdef DFA_PHASE_TIMING
// Tool-only build (tools/phase_timing.py): per-warp clock64() stamps at the phase boundaries of the
// merging forward kernel, written to a caller-provided buffer [anchor][warp][8].
__device__ long long *g_phase_buf = nullptr;
#define DFA_STAMP(i)                                                                      \
  do {                                                                                    \
    if (g_phase_buf && lane == 0)                                                         \
      g_phase_buf[(static_cast<size_t>(blockIdx.x) * NW + warp) * 8 + (i)] = clock64();   \
  } while (0)
#define DFA_GSTAMP(i)                                                                      \
  do {                                                                                     \
    if (g_phase_buf && lane == 0) {                                                        \
      unsigned long long gt_;                                                              \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));                              \
      g_phase_buf[(static_cast<size_t>(blockIdx.x) * NW + warp) * 8 + (i)] = static_cast<long long>(gt_); \
    }                                                                                      \
  } while (0)
#else
#define DFA_STAMP(i) do {} while (0)
#define DFA_GSTAMP(i) do {} while (0)
#endif

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
// T: feature element type.  LPG: lanes per group row = (C/G)*sizeof(T)/16.  U: taps in flight
// per lane.  Block = 32*G threads (warp g = group g).
template <typename T, int LPG, int U, bool TMA, int MAXT>
__global__ void __launch_bounds__(MAXT, (MAXT <= 256) ? ((sizeof(T) == 4 ? 1536 : 1024) / MAXT) : 1)
    dfa_fwd_kernel(const T *__restrict__ feat, const int *__restrict__ shape,
                   const int *__restrict__ start, const float *__restrict__ loc,
                   const float *__restrict__ weights, float *__restrict__ out, Dims d) {
  constexpr int VEC = FeatVec<T>::VEC;
  constexpr int TPW = 32 / (4 * LPG);  // taps one warp instruction covers
  extern __shared__ __align__(128) unsigned char smem[];
  const SmemLayout lay = smem_layout(d.P, d.K, d.L, d.G, TPW * U, false);
  float *s_w = reinterpret_cast<float *>(smem + lay.w);
  float *s_loc = reinterpret_cast<float *>(smem + lay.loc);
  TapQ *s_rec = reinterpret_cast<TapQ *>(smem + lay.rec);
  int *s_widx = reinterpret_cast<int *>(smem + lay.widx);
  int *s_list = reinterpret_cast<int *>(smem + lay.list);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + lay.bar);
  int *s_nvalid = reinterpret_cast<int *>(bars + 2);

  const int tid = threadIdx.x, lane = tid & 31, g = tid >> 5;
  const int anchor = blockIdx.x;  // b * A + a
  const int b = anchor / d.A;
  const int PK = d.P * d.K, wcount = PK * d.L * d.G;

  const int nv = stage_and_compact<TMA>(loc + static_cast<size_t>(anchor) * PK * 2,
                                        weights + static_cast<size_t>(anchor) * wcount, s_w,
                                        s_loc, s_list, bars, s_nvalid, PK, wcount);
  const int ntaps = nv * d.L;
  const int ntaps_pad = (ntaps + TPW * U - 1) / (TPW * U) * (TPW * U);

  // tap records, level-major so that the coarse levels' shared rows are touched back to back
  for (int t = tid; t < ntaps_pad; t += blockDim.x) {
    TapQ r[4] = {{-1, 0.f}, {-1, 0.f}, {-1, 0.f}, {-1, 0.f}};
    int widx = 0;
    if (t < ntaps) {
      const int l = t / nv, i = t - l * nv;
      const int s = s_list[i];
      const int k = s % d.K;
      const int kl = k * d.L + l;
      const int H = __ldg(shape + 2 * kl), W = __ldg(shape + 2 * kl + 1);
      TapGeom gm;
      tap_geometry(s_loc[2 * s], s_loc[2 * s + 1], H, W, __ldg(start + kl), gm);
      const float bw[4] = {gm.hh * gm.hw, gm.hh * gm.lw, gm.lh * gm.hw, gm.lh * gm.lw};
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (gm.row[q] >= 0) r[q].off = gm.row[q] * d.C, r[q].bw = bw[q];
      widx = (s * d.L + l) * d.G;
    }
    int4 *dst = reinterpret_cast<int4 *>(s_rec + 4 * t);
    dst[0] = make_int4(r[0].off, __float_as_int(r[0].bw), r[1].off, __float_as_int(r[1].bw));
    dst[1] = make_int4(r[2].off, __float_as_int(r[2].bw), r[3].off, __float_as_int(r[3].bw));
    s_widx[t] = widx;
  }
  __syncthreads();
  if (TMA) mbar_wait(&bars[1], 0);  // weights have landed

  const int j = lane % LPG, q = (lane / LPG) & 3, sub = lane / (4 * LPG);
  const int cpg = d.C / d.G;
  const T *fb = feat + static_cast<size_t>(b) * d.num_feat * d.C + g * cpg + j * VEC;
  float acc[VEC];
#pragma unroll
  for (int c = 0; c < VEC; ++c) acc[c] = 0.f;

  for (int t0 = 0; t0 < ntaps_pad; t0 += TPW * U) {
    float v[U][VEC], cw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int t = t0 + u * TPW + sub;
      const int2 rq = *reinterpret_cast<const int2 *>(s_rec + 4 * t + q);
      const float wgt = s_w[s_widx[t] + g];
      if (rq.x >= 0) {
        cw[u] = __int_as_float(rq.y) * wgt;
        FeatVec<T>::load(fb + rq.x, v[u]);
      } else {  // corner outside the map (zero padding) or padding tap: contributes nothing
        cw[u] = 0.f;
#pragma unroll
        for (int c = 0; c < VEC; ++c) v[u][c] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int c = 0; c < VEC; ++c) acc[c] = fmaf(cw[u], v[u][c], acc[c]);
  }
  // fold corners (and sub-taps): lanes differing in bits >= log2(LPG)
#pragma unroll
  for (int m = LPG; m < 32; m <<= 1)
#pragma unroll
    for (int c = 0; c < VEC; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], m);
  if (lane < LPG) {
    float4 *o = reinterpret_cast<float4 *>(out + static_cast<size_t>(anchor) * d.C + g * cpg + j * VEC);
#pragma unroll
    for (int c = 0; c < VEC / 4; ++c)
      o[c] = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
  }
}


// ------------------------------------------------------------------------------------------
// forward, row-sliced mapping
// ------------------------------------------------------------------------------------------
// A feature row (C channels) is `vpr` 16-byte vectors.  The CTA is split into NT/vpr slices;
// slice s owns taps s, s+slices, ... and inside a slice thread v owns vector v of the row, i.e.
// VEC consecutive channels of ONE group, for all four corners of the tap.  Per tap a thread
// issues four 128-bit loads (a warp covers 512 contiguous bytes of each corner row), reads the
// tap record with two broadcast LDS.128 and accumulates with packed FFMA2.  Slices are folded
// through shared memory at the end.  Compared with the one-warp-per-group mapping this needs
// ~3.5x fewer instructions per byte gathered and shortens an anchor's serial chain by `slices`.
struct SmemLayout2 {
  uint32_t w, loc, off, bw, widx, list, tab, red, bar, total;
};
__host__ __device__ inline SmemLayout2 smem_layout2(int P, int K, int L, int G, int C, int slices,
                                                    int tap_pad) {
  SmemLayout2 s;
  const uint32_t taps = align_up(static_cast<uint32_t>(P) * K * L, tap_pad) + tap_pad;
  uint32_t o = 0;
  s.w = o, o = align_up(o + 4u * P * K * L * G, 16);
  s.loc = o, o = align_up(o + 8u * P * K, 16);
  s.off = o, o = align_up(o + 16u * taps, 16);
  s.bw = o, o = align_up(o + 16u * taps, 16);
  s.widx = o, o = align_up(o + 2u * taps, 16);  // 16-bit: rows_vpr() checks P*K*L*G < 65536
  s.list = o, o = align_up(o + 4u * P * K, 16);
  s.tab = o, o = align_up(o + 12u * K * L, 16);
  s.red = o, o = align_up(o + 4u * slices * C, 16);
  s.bar = o, o += 32;
  s.total = o;
  return s;
}

template <typename T, int U, bool TMA, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
    dfa_fwd_rows_kernel(const T *__restrict__ feat, const int *__restrict__ shape,
                        const int *__restrict__ start, const float *__restrict__ loc,
                        const float *__restrict__ weights, float *__restrict__ out, Dims d,
                        int vpr_log2, int split_from, int split_log2) {
  constexpr int VEC = FeatVec<T>::VEC;
  extern __shared__ __align__(128) unsigned char smem[];
  // Channel split (small grids): CTAs from `split_from` on share an anchor 2^split_log2 ways, each
  // taking a contiguous block of channels — a quarter row per load and four times the slices, so the
  // anchor's chain of dependent load rounds is a quarter as long; every CTA writes its own channels,
  // nothing is combined.  The launcher splits the anchors that do not fit the first wave.
  int anchor = blockIdx.x;  // b * A + a
  int ch_base = 0, Cs = d.C;
  if (static_cast<int>(blockIdx.x) >= split_from) {
    const int r = blockIdx.x - split_from;
    anchor = split_from + (r >> split_log2);
    Cs = d.C >> split_log2;
    ch_base = (r & ((1 << split_log2) - 1)) * Cs;
    vpr_log2 -= split_log2;
  }
  const int vpr = 1 << vpr_log2;
  const int slices = NT >> vpr_log2;
  const SmemLayout2 lay = smem_layout2(d.P, d.K, d.L, d.G, Cs, slices, slices * U);
  float *s_w = reinterpret_cast<float *>(smem + lay.w);
  float *s_loc = reinterpret_cast<float *>(smem + lay.loc);
  uint4 *s_off = reinterpret_cast<uint4 *>(smem + lay.off);
  float4 *s_bw = reinterpret_cast<float4 *>(smem + lay.bw);
  uint16_t *s_widx = reinterpret_cast<uint16_t *>(smem + lay.widx);
  int *s_list = reinterpret_cast<int *>(smem + lay.list);
  int *s_tab = reinterpret_cast<int *>(smem + lay.tab);
  float *s_red = reinterpret_cast<float *>(smem + lay.red);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + lay.bar);
  int *s_nvalid = reinterpret_cast<int *>(bars + 2);

  const int tid = threadIdx.x;
#ifdef DFA_PHASE_TIMING
  constexpr int NW = NT / 32;
  const int lane = tid & 31, warp = tid >> 5;
#endif
  DFA_STAMP(0);
  DFA_GSTAMP(6);
  const int b = anchor / d.A;
  const int PK = d.P * d.K, wcount = PK * d.L * d.G;

  // operands first (tried: plain loads for the locations, and weight lines of the valid samples only
  // after the mask is known — both slower at this shape), then the level tables → shared memory
  // while the TMA copies are in flight
  stage_issue<TMA>(loc + static_cast<size_t>(anchor) * PK * 2, weights + static_cast<size_t>(anchor) * wcount,
                   s_w, s_loc, bars, PK, wcount);
  for (int i = tid; i < d.K * d.L; i += NT) {
    s_tab[3 * i] = __ldg(shape + 2 * i);
    s_tab[3 * i + 1] = __ldg(shape + 2 * i + 1);
    s_tab[3 * i + 2] = __ldg(start + i);
  }
  const int nv = stage_and_compact<TMA>(loc + static_cast<size_t>(anchor) * PK * 2,
                                        weights + static_cast<size_t>(anchor) * wcount, s_w,
                                        s_loc, s_list, bars, s_nvalid, PK, wcount, true);
  DFA_STAMP(1);
  const int ntaps = nv * d.L;
  const int step = slices * U;
  const int ntaps_pad = (ntaps + step - 1) / step * step;

  // Tap records.  Corner offsets are BYTE offsets inside the batch item.  A corner that falls
  // outside the map (zero padding) is redirected to an in-bounds corner of the same tap with a
  // zero bilinear weight — a valid sample always has one — so the main loop needs no predicates
  // (and a non-finite feature there would reach the reference's result through the in-bounds
  // corner as well).  Padding taps replay tap 0 with zero weights.
  for (int t = tid; t < ntaps_pad; t += NT) {
    const int tt = t < ntaps ? t : 0;
    const int l = tt / nv, i = tt - l * nv;  // level-major: coarse-level neighbours back to back
    const int s = s_list[i];
    const int kl = (s % d.K) * d.L + l;
    TapGeom gm;
    tap_geometry(s_loc[2 * s], s_loc[2 * s + 1], s_tab[3 * kl], s_tab[3 * kl + 1], s_tab[3 * kl + 2],
                 gm);
    const int safe = gm.row[0] >= 0 ? gm.row[0] : gm.row[1] >= 0 ? gm.row[1]
                   : gm.row[2] >= 0 ? gm.row[2] : gm.row[3];
    const float live = t < ntaps ? 1.f : 0.f;
    const uint32_t rb = static_cast<uint32_t>(d.C) * sizeof(T);
    uint4 off;
    float4 bw;
    off.x = (gm.row[0] >= 0 ? gm.row[0] : safe) * rb, bw.x = gm.row[0] >= 0 ? live * gm.hh * gm.hw : 0.f;
    off.y = (gm.row[1] >= 0 ? gm.row[1] : safe) * rb, bw.y = gm.row[1] >= 0 ? live * gm.hh * gm.lw : 0.f;
    off.z = (gm.row[2] >= 0 ? gm.row[2] : safe) * rb, bw.z = gm.row[2] >= 0 ? live * gm.lh * gm.hw : 0.f;
    off.w = (gm.row[3] >= 0 ? gm.row[3] : safe) * rb, bw.w = gm.row[3] >= 0 ? live * gm.lh * gm.lw : 0.f;
    s_off[t] = off, s_bw[t] = bw, s_widx[t] = static_cast<uint16_t>((s * d.L + l) * d.G);
  }
  DFA_STAMP(2);
  __syncthreads();
  if (TMA) mbar_wait(&bars[1], 0);  // weights have landed
  DFA_STAMP(3);

  const int slice = tid >> vpr_log2, v = tid & (vpr - 1);
  const int ch = ch_base + v * VEC;
  float acc[VEC];
#pragma unroll
  for (int c = 0; c < VEC; ++c) acc[c] = 0.f;

  if (slice < slices) {
    const unsigned char *fb = reinterpret_cast<const unsigned char *>(feat) +
                              static_cast<size_t>(b) * d.num_feat * d.C * sizeof(T);
    const uint32_t lane_off = static_cast<uint32_t>(ch) * sizeof(T);
    const float *s_wg = s_w + ch / (d.C / d.G);
    for (int t0 = slice; t0 < ntaps_pad; t0 += step) {
      typename FeatVec<T>::raw_t val[U][4];
      float cw[U][4];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int t = t0 + u * slices;
        const uint4 off = s_off[t];
        val[u][0] = FeatVec<T>::load_raw(reinterpret_cast<const T *>(fb + (off.x + lane_off)));
        val[u][1] = FeatVec<T>::load_raw(reinterpret_cast<const T *>(fb + (off.y + lane_off)));
        val[u][2] = FeatVec<T>::load_raw(reinterpret_cast<const T *>(fb + (off.z + lane_off)));
        val[u][3] = FeatVec<T>::load_raw(reinterpret_cast<const T *>(fb + (off.w + lane_off)));
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int t = t0 + u * slices;
        const float4 bw = s_bw[t];
        const float wgt = s_wg[s_widx[t]];
        cw[u][0] = bw.x * wgt, cw[u][1] = bw.y * wgt, cw[u][2] = bw.z * wgt, cw[u][3] = bw.w * wgt;
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int q = 0; q < 4; ++q) FeatVec<T>::fma(acc, cw[u][q], val[u][q]);
    }
    float4 *r = reinterpret_cast<float4 *>(s_red + slice * Cs + (ch - ch_base));
#pragma unroll
    for (int c = 0; c < VEC / 4; ++c)
      r[c] = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
  }
  DFA_STAMP(4);
  __syncthreads();
  for (int c = tid; c < Cs; c += NT) {
    float sum = 0.f;
    for (int sl = 0; sl < slices; ++sl) sum += s_red[sl * Cs + c];
    out[static_cast<size_t>(anchor) * d.C + ch_base + c] = sum;
  }
  DFA_STAMP(5);
  DFA_GSTAMP(7);
}



// ------------------------------------------------------------------------------------------
// fused module forward: key points + projection + softmax of the attention logits + gather
// ------------------------------------------------------------------------------------------
// The row-sliced kernel with the module's front end folded into its prologue (inference): instead
// of reading sampling locations and softmaxed weights that three earlier kernels wrote, the CTA
// takes the anchor (11 floats), the learnable-offset logits and the attention LOGITS and computes
//   * the anchor's P key points and their K camera projections (same device code as
//     dfa_keypoints_project, so the sampling locations are bit-identical),
//   * exp(logit - max) for the K*L*P*G logits, kept UN-normalised in shared memory in their natural
//     (k,l,p,g) order; the softmax denominator is per (anchor, group), so it is applied once to the
//     finished output channels instead of to 2,496 weights,
// and then gathers exactly as dfa_fwd_rows_kernel does.  Neither the [bs,A,P,K,2] locations nor the
// 9 MB [bs,A,P,K,L,G] weights tensor exist.  Logits may be split (weights_fc linearity, see the
// front-end file): logits_a [bs,A,L*P*G] + logits_k [bs,K,L*P*G]; logits_k = NULL means logits_a is
// the full [bs,A,K,L*P*G] tensor.
struct FusedArgs {
  const float *anchor, *fix_scale, *off_logits, *proj, *wh, *logits_a, *logits_k;
  float *loc_out;  // optional [bs,A,P,K,2] copy of the sampling locations (tests)
  int num_fix;
};

struct SmemLayoutF {
  SmemLayout2 r;
  uint32_t kp, red, inv, la, total;
};
__host__ __device__ inline SmemLayoutF smem_layout_fused(int P, int K, int L, int G, int C, int slices,
                                                         int nt) {
  SmemLayoutF s;
  s.r = smem_layout2(P, K, L, G, C, slices, slices);
  uint32_t o = s.r.total;
  s.kp = o, o = align_up(o + 12u * P, 16);
  s.red = o, o = align_up(o + 4u * nt, 16);
  s.inv = o, o = align_up(o + 4u * G, 16);
  s.la = o, o = align_up(o + 4u * L * P * G, 16);
  s.total = o;
  return s;
}

template <typename T, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
    dfa_fwd_fused_kernel(const T *__restrict__ feat, const int *__restrict__ shape,
                         const int *__restrict__ start, FusedArgs fa, float *__restrict__ out, Dims d,
                         int vpr_log2, int split_from, int split_log2) {
  constexpr int VEC = FeatVec<T>::VEC;
  extern __shared__ __align__(128) unsigned char smem[];
  // channel split of the last wave's anchors, as in dfa_fwd_rows_kernel: every split CTA runs the
  // whole prologue and gathers its own block of channels with 2^split_log2 times the slices
  int anchor = blockIdx.x;  // b * A + a
  int ch_base = 0, Cs = d.C;
  if (static_cast<int>(blockIdx.x) >= split_from) {
    const int r = blockIdx.x - split_from;
    anchor = split_from + (r >> split_log2);
    Cs = d.C >> split_log2;
    ch_base = (r & ((1 << split_log2) - 1)) * Cs;
    vpr_log2 -= split_log2;
  }
  const int vpr = 1 << vpr_log2;
  const int slices = NT >> vpr_log2;
  const SmemLayoutF layf = smem_layout_fused(d.P, d.K, d.L, d.G, Cs, slices, NT);
  const SmemLayout2 &lay = layf.r;
  float *s_w = reinterpret_cast<float *>(smem + lay.w);
  float *s_loc = reinterpret_cast<float *>(smem + lay.loc);
  uint4 *s_off = reinterpret_cast<uint4 *>(smem + lay.off);
  float4 *s_bw = reinterpret_cast<float4 *>(smem + lay.bw);
  uint16_t *s_widx = reinterpret_cast<uint16_t *>(smem + lay.widx);
  int *s_list = reinterpret_cast<int *>(smem + lay.list);
  int *s_tab = reinterpret_cast<int *>(smem + lay.tab);
  float *s_red = reinterpret_cast<float *>(smem + lay.red);
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + lay.bar);
  int *s_nvalid = reinterpret_cast<int *>(bar + 1);
  float *s_la = reinterpret_cast<float *>(smem + layf.la);
  float *s_kp = reinterpret_cast<float *>(smem + layf.kp);
  float *s_gred = reinterpret_cast<float *>(smem + layf.red);
  float *s_inv = reinterpret_cast<float *>(smem + layf.inv);

  const int tid = threadIdx.x;
  const int b = anchor / d.A;
  const int PK = d.P * d.K, LP = d.L * d.P, lpg = LP * d.G, n_el = d.K * lpg;
#ifdef DFA_PHASE_TIMING
  constexpr int NW = NT / 32;
  const int lane = tid & 31, warp = tid >> 5;
#endif
  DFA_STAMP(0);

  // The logits are the long pole of the prologue: start their TMA bulk copies first (camera part or
  // the full block straight into the weight buffer, anchor part beside it) and compute key points,
  // projections and the sample mask while they are in flight.
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
    if (fa.logits_k) {
      mbar_expect_tx(bar, 4u * (n_el + lpg));
      tma_bulk_g2s(s_w, fa.logits_k + static_cast<size_t>(b) * n_el, 4u * n_el, bar);
      tma_bulk_g2s(s_la, fa.logits_a + static_cast<size_t>(anchor) * lpg, 4u * lpg, bar);
    } else {
      mbar_expect_tx(bar, 4u * n_el);
      tma_bulk_g2s(s_w, fa.logits_a + static_cast<size_t>(anchor) * n_el, 4u * n_el, bar);
    }
  }
  for (int i = tid; i < d.K * d.L; i += NT) {
    s_tab[3 * i] = __ldg(shape + 2 * i);
    s_tab[3 * i + 1] = __ldg(shape + 2 * i + 1);
    s_tab[3 * i + 2] = __ldg(start + i);
  }
  // ---- key points, then their projections ---------------------------------------------------------
  for (int p = tid; p < d.P; p += NT)
    key_point(fa.anchor + static_cast<size_t>(anchor) * 11, fa.fix_scale, fa.num_fix,
              fa.off_logits ? fa.off_logits + static_cast<size_t>(anchor) * (d.P - fa.num_fix) * 3 : nullptr,
              p, s_kp[3 * p], s_kp[3 * p + 1], s_kp[3 * p + 2]);
  __syncthreads();
  DFA_STAMP(1);
  for (int s = tid; s < PK; s += NT) {
    const int p = s / d.K, k = s - p * d.K;
    float px, py;
    project_point(fa.proj + (static_cast<size_t>(b) * d.K + k) * 16,
                  fa.wh ? fa.wh + (b * d.K + k) * 2 : nullptr, s_kp[3 * p], s_kp[3 * p + 1], s_kp[3 * p + 2],
                  px, py);
    s_loc[2 * s] = px, s_loc[2 * s + 1] = py;
    if (fa.loc_out && ch_base == 0) {
      fa.loc_out[(static_cast<size_t>(anchor) * PK + s) * 2] = px;
      fa.loc_out[(static_cast<size_t>(anchor) * PK + s) * 2 + 1] = py;
    }
  }
  DFA_STAMP(2);
  // ---- compaction of the samples that pass the (0,1) test -------------------------------------------
  __syncthreads();  // s_loc complete (and the mbarrier initialised for every waiter)
  if (tid < 32) {
    int n = 0;
    for (int base = 0; base < PK; base += 32) {
      const int s = base + tid;
      bool v = false;
      if (s < PK) v = sample_valid(s_loc[2 * s], s_loc[2 * s + 1]);
      const unsigned m = __ballot_sync(0xffffffffu, v);
      if (v) s_list[n + __popc(m & ((1u << tid) - 1u))] = s;
      n += __popc(m);
    }
    if (tid == 0) *s_nvalid = n;
  }
  // ---- softmax numerators of the attention logits, natural (k,l,p,g) order ------------------------
  // element e = i * NT + tid: its group is tid % G for every i (G divides NT), so a thread reduces
  // its own elements first and the groups are combined once.
  mbar_wait(bar, 0);
  DFA_STAMP(3);
  {
    const bool split = fa.logits_k != nullptr;
    float mx = -INFINITY;
    int ia = tid % lpg;
    for (int e = tid; e < n_el; e += NT) {
      float v = s_w[e];
      if (split) {
        v += s_la[ia];
        ia += NT;
        while (ia >= lpg) ia -= lpg;
        s_w[e] = v;
      }
      mx = fmaxf(mx, v);
    }
    mx = group_reduce<NT>(mx, s_gred, tid, d.G, true);
    float sum = 0.f;
    int e = tid;
    for (; e + 3 * NT < n_el; e += 4 * NT) {  // four independent exponentials in flight
      const float v0 = expf(s_w[e] - mx), v1 = expf(s_w[e + NT] - mx), v2 = expf(s_w[e + 2 * NT] - mx),
                  v3 = expf(s_w[e + 3 * NT] - mx);
      s_w[e] = v0, s_w[e + NT] = v1, s_w[e + 2 * NT] = v2, s_w[e + 3 * NT] = v3;
      sum += (v0 + v1) + (v2 + v3);
    }
    for (; e < n_el; e += NT) {
      const float v = expf(s_w[e] - mx);
      s_w[e] = v;
      sum += v;
    }
    sum = group_reduce<NT>(sum, s_gred, tid, d.G, false);
    if (tid < d.G) s_inv[tid] = 1.f / sum;
  }
  DFA_STAMP(4);
  __syncthreads();
  const int nv = *s_nvalid;
  const int ntaps = nv * d.L;
  const int step = slices;
  const int ntaps_pad = (ntaps + step - 1) / step * step;
  for (int t = tid; t < ntaps_pad; t += NT) {
    const int tt = t < ntaps ? t : 0;
    const int l = tt / nv, i = tt - l * nv;
    const int s = s_list[i];
    const int p = s / d.K, k = s - p * d.K;
    const int kl = k * d.L + l;
    TapGeom gm;
    tap_geometry(s_loc[2 * s], s_loc[2 * s + 1], s_tab[3 * kl], s_tab[3 * kl + 1], s_tab[3 * kl + 2], gm);
    const int safe = gm.row[0] >= 0 ? gm.row[0] : gm.row[1] >= 0 ? gm.row[1]
                   : gm.row[2] >= 0 ? gm.row[2] : gm.row[3];
    const float live = t < ntaps ? 1.f : 0.f;
    const uint32_t rb = static_cast<uint32_t>(d.C) * sizeof(T);
    uint4 off;
    float4 bw;
    off.x = (gm.row[0] >= 0 ? gm.row[0] : safe) * rb, bw.x = gm.row[0] >= 0 ? live * gm.hh * gm.hw : 0.f;
    off.y = (gm.row[1] >= 0 ? gm.row[1] : safe) * rb, bw.y = gm.row[1] >= 0 ? live * gm.hh * gm.lw : 0.f;
    off.z = (gm.row[2] >= 0 ? gm.row[2] : safe) * rb, bw.z = gm.row[2] >= 0 ? live * gm.lh * gm.hw : 0.f;
    off.w = (gm.row[3] >= 0 ? gm.row[3] : safe) * rb, bw.w = gm.row[3] >= 0 ? live * gm.lh * gm.lw : 0.f;
    s_off[t] = off, s_bw[t] = bw, s_widx[t] = static_cast<uint16_t>(((k * d.L + l) * d.P + p) * d.G);  // (k,l,p,g) order
  }
  __syncthreads();
  DFA_STAMP(5);

  const int slice = tid >> vpr_log2, v = tid & (vpr - 1);
  const int ch = ch_base + v * VEC;
  float acc[VEC];
#pragma unroll
  for (int c = 0; c < VEC; ++c) acc[c] = 0.f;
  if (slice < slices) {
    const unsigned char *fb = reinterpret_cast<const unsigned char *>(feat) +
                              static_cast<size_t>(b) * d.num_feat * d.C * sizeof(T);
    const uint32_t lane_off = static_cast<uint32_t>(ch) * sizeof(T);
    const float *s_wg = s_w + ch / (d.C / d.G);
    for (int t = slice; t < ntaps_pad; t += step) {
      typename FeatVec<T>::raw_t val[4];
      const uint4 off = s_off[t];
      val[0] = FeatVec<T>::load_raw(reinterpret_cast<const T *>(fb + (off.x + lane_off)));
      val[1] = FeatVec<T>::load_raw(reinterpret_cast<const T *>(fb + (off.y + lane_off)));
      val[2] = FeatVec<T>::load_raw(reinterpret_cast<const T *>(fb + (off.z + lane_off)));
      val[3] = FeatVec<T>::load_raw(reinterpret_cast<const T *>(fb + (off.w + lane_off)));
      const float4 bw = s_bw[t];
      const float wgt = s_wg[s_widx[t]];
      FeatVec<T>::fma(acc, bw.x * wgt, val[0]);
      FeatVec<T>::fma(acc, bw.y * wgt, val[1]);
      FeatVec<T>::fma(acc, bw.z * wgt, val[2]);
      FeatVec<T>::fma(acc, bw.w * wgt, val[3]);
    }
    float4 *r = reinterpret_cast<float4 *>(s_red + slice * Cs + (ch - ch_base));
#pragma unroll
    for (int c = 0; c < VEC / 4; ++c)
      r[c] = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
  }
  DFA_STAMP(6);
  __syncthreads();
  const int cpg = d.C / d.G;
  for (int c = tid; c < Cs; c += NT) {
    float sum = 0.f;
    for (int sl = 0; sl < slices; ++sl) sum += s_red[sl * Cs + c];
    out[static_cast<size_t>(anchor) * d.C + ch_base + c] = sum * s_inv[(ch_base + c) / cpg];  // the softmax denominator, once
  }
  DFA_STAMP(7);
}

// Shape-generic forward (any C, G with C % G == 0, any alignment): one CTA per anchor, threads
// stride over channels, scalar loads.  Same staging/geometry code, no atomics.
template <typename T>
__global__ void __launch_bounds__(256)
    dfa_fwd_generic_kernel(const T *__restrict__ feat, const int *__restrict__ shape,
                           const int *__restrict__ start, const float *__restrict__ loc,
                           const float *__restrict__ weights, float *__restrict__ out, Dims d) {
  extern __shared__ __align__(128) unsigned char smem[];
  const SmemLayout lay = smem_layout(d.P, d.K, d.L, d.G, 1, false);
  float *s_w = reinterpret_cast<float *>(smem + lay.w);
  float *s_loc = reinterpret_cast<float *>(smem + lay.loc);
  TapQ *s_rec = reinterpret_cast<TapQ *>(smem + lay.rec);
  int *s_widx = reinterpret_cast<int *>(smem + lay.widx);
  int *s_list = reinterpret_cast<int *>(smem + lay.list);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + lay.bar);
  int *s_nvalid = reinterpret_cast<int *>(bars + 2);
  const int tid = threadIdx.x, anchor = blockIdx.x, b = anchor / d.A;
  const int PK = d.P * d.K, wcount = PK * d.L * d.G;
  const int nv = stage_and_compact<false>(loc + static_cast<size_t>(anchor) * PK * 2,
                                          weights + static_cast<size_t>(anchor) * wcount, s_w,
                                          s_loc, s_list, bars, s_nvalid, PK, wcount);
  const int ntaps = nv * d.L;
  for (int t = tid; t < ntaps; t += blockDim.x) {
    const int l = t / nv, i = t - l * nv, s = s_list[i], k = s % d.K, kl = k * d.L + l;
    const int H = __ldg(shape + 2 * kl), W = __ldg(shape + 2 * kl + 1);
    TapGeom gm;
    tap_geometry(s_loc[2 * s], s_loc[2 * s + 1], H, W, __ldg(start + kl), gm);
    const float bw[4] = {gm.hh * gm.hw, gm.hh * gm.lw, gm.lh * gm.hw, gm.lh * gm.lw};
    for (int q = 0; q < 4; ++q) {
      s_rec[4 * t + q].off = gm.row[q] >= 0 ? gm.row[q] * d.C : -1;
      s_rec[4 * t + q].bw = gm.row[q] >= 0 ? bw[q] : 0.f;
    }
    s_widx[t] = (s * d.L + l) * d.G;
  }
  __syncthreads();
  const int cpg = d.C / d.G;
  const T *fb = feat + static_cast<size_t>(b) * d.num_feat * d.C;
  for (int c = tid; c < d.C; c += blockDim.x) {
    const int grp = c / cpg;
    float acc = 0.f;
    for (int t = 0; t < ntaps; ++t) {
      const float wgt = s_w[s_widx[t] + grp];
      float val = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const TapQ r = s_rec[4 * t + q];
        if (r.off >= 0) val = fmaf(r.bw, static_cast<float>(fb[r.off + c]), val);
      }
      acc = fmaf(wgt, val, acc);
    }
    out[static_cast<size_t>(anchor) * d.C + c] = acc;
  }
}

// ------------------------------------------------------------------------------------------
// side channel for the bit-exact parity tests: the geometry above, nothing else
// ------------------------------------------------------------------------------------------
__global__ void dfa_debug_indices_kernel(const int *__restrict__ shape, const int *__restrict__ start,
                                         const float *__restrict__ loc, uint8_t *__restrict__ valid,
                                         int *__restrict__ rows, Dims d) {
  const long long n = static_cast<long long>(d.bs) * d.A * d.P * d.K;
  for (long long s = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; s < n;
       s += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(s % d.K);
    const float x = loc[2 * s], y = loc[2 * s + 1];
    const bool ok = sample_valid(x, y);
    valid[s] = ok ? 1 : 0;
    for (int l = 0; l < d.L; ++l) {
      const int kl = k * d.L + l;
      TapGeom gm;
      tap_geometry(x, y, shape[2 * kl], shape[2 * kl + 1], start[kl], gm);
      for (int q = 0; q < 4; ++q) rows[(s * d.L + l) * 4 + q] = ok ? gm.row[q] : -1;
    }
  }
}


// ------------------------------------------------------------------------------------------
// dfa_forward_host, pull mode: only what the forward will read crosses the host link
// ------------------------------------------------------------------------------------------
// With camera-rig inputs ~19 % of the samples are valid and ~27 % of the feature rows are referenced at
// all.  When the host buffers are pinned and mapped (device-accessible), the device marks the rows the
// forward will touch — same sample_valid / tap_geometry as every kernel, so the set is exact — and
// pulls exactly those rows, and the weight lines of the valid samples, straight from host memory.
// Workspace header: [0] rows pulled, [1] 16-byte weight vectors pulled, [2] bytes per row, [3] bytes of
// the small operands copied whole.
__global__ void dfa_host_mark_kernel(const int *__restrict__ shape, const int *__restrict__ start,
                                     const float *__restrict__ loc, uint32_t *__restrict__ bitmap,
                                     uint8_t *__restrict__ svalid, unsigned long long *__restrict__ hdr,
                                     unsigned long long row_bytes, unsigned long long fixed_bytes, Dims d) {
  if (blockIdx.x == 0 && threadIdx.x == 0) hdr[2] = row_bytes, hdr[3] = fixed_bytes;
  const long long n = static_cast<long long>(d.bs) * d.A * d.P * d.K;
  const long long per_item = static_cast<long long>(d.A) * d.P * d.K;
  for (long long s = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; s < n;
       s += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(s % d.K);
    const long long row0 = (s / per_item) * d.num_feat;
    const float x = loc[2 * s], y = loc[2 * s + 1];
    const bool ok = sample_valid(x, y);
    svalid[s] = ok ? 1 : 0;
    if (!ok) continue;
    for (int l = 0; l < d.L; ++l) {
      const int kl = k * d.L + l;
      TapGeom gm;
      tap_geometry(x, y, __ldg(shape + 2 * kl), __ldg(shape + 2 * kl + 1), __ldg(start + kl), gm);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (gm.row[q] >= 0) {
          const long long r = row0 + gm.row[q];
          atomicOr(bitmap + (r >> 5), 1u << (r & 31));
        }
    }
  }
}

// One warp per bitmap word (32 rows); up to four marked rows are copied at a time so that a lane keeps
// several 16-byte host reads in flight.  `src` is the device alias of the pinned host table.
__global__ void dfa_host_pull_rows_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst,
                                          const uint32_t *__restrict__ bitmap, long long nwords,
                                          long long nrows, int row_vecs, unsigned long long *__restrict__ hdr) {
  const int lane = threadIdx.x & 31;
  const long long nwarps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  unsigned long long pulled = 0;
  for (long long w = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5); w < nwords;
       w += nwarps) {
    uint32_t m = bitmap[w];
    if (w * 32 + 32 > nrows) m &= (1u << (nrows - w * 32)) - 1u;
    pulled += __popc(m);
    while (m) {
      long long r[4];
      int nr = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        r[i] = -1;
        if (m) r[i] = (w * 32 + (__ffs(m) - 1)) * row_vecs, m &= m - 1, ++nr;
      }
      for (int v = lane; v < row_vecs; v += 32) {
        uint4 val[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i < nr) val[i] = __ldcv(src + r[i] + v);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i < nr) dst[r[i] + v] = val[i];
      }
    }
  }
  if (lane == 0 && pulled) atomicAdd(hdr, pulled);
}

// Weight lines (L*G floats = line_vecs 16-byte vectors) of the valid samples.
__global__ void dfa_host_pull_weights_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst,
                                             const uint8_t *__restrict__ svalid, long long nvecs, int line_vecs,
                                             unsigned long long *__restrict__ hdr) {
  unsigned long long pulled = 0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvecs;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    if (svalid[i / line_vecs]) dst[i] = __ldcv(src + i), ++pulled;
  }
  for (int m = 16; m; m >>= 1) pulled += __shfl_xor_sync(0xffffffffu, pulled, m);
  if ((threadIdx.x & 31) == 0 && pulled) atomicAdd(hdr + 1, pulled);
}

}  // namespace

#include "dfa_forward_win.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// launchers and dispatch
// ------------------------------------------------------------------------------------------
constexpr int FWD_U = 4;

template <typename T, int LPG, bool TMA, int MAXT>
int launch_fwd_t(const void *feat, const int *shape, const int *start, const float *loc,
                 const float *w, float *out, const Dims &d, cudaStream_t st) {
  auto kern = dfa_fwd_kernel<T, LPG, FWD_U, TMA, MAXT>;
  constexpr int TPW = 32 / (4 * LPG);
  const SmemLayout lay = smem_layout(d.P, d.K, d.L, d.G, TPW * FWD_U, false);
  if (int rc = set_smem(kern, lay.total)) return rc;
  kern<<<d.bs * d.A, 32 * d.G, lay.total, st>>>(static_cast<const T *>(feat), shape, start, loc, w,
                                               out, d);
  return static_cast<int>(cudaGetLastError());
}

#define DFA_DISPATCH_LPG(CALL)                    \
  switch (lpg) {                                  \
    case 8: return CALL(8);                       \
    case 4: return CALL(4);                       \
    case 2: return CALL(2);                       \
    default: return CALL(1);                      \
  }

// Channel split of the last wave.  The anchors of the last, partial wave of resident CTAs start only
// when earlier anchors retire and then end the kernel (R50, bs=1: 900 anchors on 888 slots; the 12
// late CTAs were the last to finish): they are split four ways by channels, which makes their CTAs a
// quarter as long.  Measured: 19.2 -> 17.9 us at bs=1 / 900 anchors, 24.8 -> 23.5 us at 1220 anchors,
// 31.6 -> 30.3 us at bs=2, neutral from bs=4 on.  DFA_FWD_SPLIT: 0 = never, 1 = as described
// (default), 2 / 3 = every anchor two / four ways (experiments: par / slower).
// SM count of the current device, queried once per device and process.
inline int device_sm_count() {
  static std::atomic<int> sm_count[64];  // 0 = not queried yet; racing threads store the same value
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  int sms = sm_count[dev].load(std::memory_order_relaxed);
  if (!sms && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess)
    sm_count[dev].store(sms, std::memory_order_relaxed);
  return sms;
}

template <typename T>
void last_wave_split(const Dims &d, int vpr, int ctas_per_sm, long long &split_from, int &split_log2,
                     long long &grid) {
  constexpr int VEC = FeatVec<T>::VEC;
  const long long total = static_cast<long long>(d.bs) * d.A;
  split_from = total, grid = total, split_log2 = 0;
  const int mode = DFA_KNOB("DFA_FWD_SPLIT", 1);
  const int max_log2 = ((d.C / 4) % VEC == 0 && vpr >= 4) ? 2 : (((d.C / 2) % VEC == 0 && vpr >= 2) ? 1 : 0);
  if (mode == 1 && max_log2 == 2) {
    const int sms = device_sm_count();
    {
      const long long slots = static_cast<long long>(sms) * ctas_per_sm;
      const long long rem = slots > 0 ? total % slots : 0;  // anchors of the last, partial wave
      const int frac = DFA_KNOB("DFA_FWD_SPLIT_FRAC", 2);  // split when the last wave is at most 1/frac full
      if (total > slots && rem > 0 && frac > 0 && rem <= slots / frac)
        split_from = total - rem, split_log2 = 2, grid = split_from + 4 * rem;
    }
  } else if (mode >= 2 && max_log2 >= mode - 1) {
    split_from = 0, split_log2 = mode - 1, grid = total << split_log2;
  }
}

template <typename T, int U, bool TMA, int NT, int MINB>
int launch_fwd_rows(const void *feat, const int *shape, const int *start, const float *loc,
                    const float *w, float *out, const Dims &d, int vpr, cudaStream_t st) {
  auto kern = dfa_fwd_rows_kernel<T, U, TMA, NT, MINB>;
  const int slices = NT / vpr;
  int vpr_log2 = 0;
  while ((1 << vpr_log2) < vpr) ++vpr_log2;
  long long split_from = 0, grid = 0;
  int split_log2 = 0;
  last_wave_split<T>(d, vpr, MINB, split_from, split_log2, grid);
  const int max_slices = slices << split_log2;
  SmemLayout2 lay = smem_layout2(d.P, d.K, d.L, d.G, d.C, slices, slices * U);
  const SmemLayout2 lay_s = smem_layout2(d.P, d.K, d.L, d.G, d.C >> split_log2, max_slices, max_slices * U);
  if (lay_s.total > lay.total) lay = lay_s;
  if (int rc = set_smem(kern, lay.total)) return rc;
  kern<<<static_cast<unsigned int>(grid), NT, lay.total, st>>>(static_cast<const T *>(feat), shape, start, loc, w,
                                                               out, d, vpr_log2, static_cast<int>(split_from),
                                                               split_log2);
  return static_cast<int>(cudaGetLastError());
}

// Row-sliced fast path: the row is a power-of-two number of 16-byte vectors (<= block size) and
// every vector lies inside one channel group.
template <typename T>
int rows_vpr(const Dims &d, const void *feat, int nt) {
  constexpr int VEC = FeatVec<T>::VEC;
  if (d.C % VEC != 0 || (d.C / d.G) % VEC != 0 || !aligned(feat, 16)) return 0;
  const int vpr = d.C / VEC;
  if (vpr > nt || nt % vpr != 0 || (vpr & (vpr - 1)) != 0) return 0;
  if (static_cast<long long>(d.num_feat) * d.C * static_cast<long long>(sizeof(T)) >= (1ll << 32)) return 0;
  if (static_cast<long long>(d.P) * d.K * d.L * d.G >= 65536) return 0;  // 16-bit weight index per tap
  return vpr;
}

template <typename T, int VPL, bool TMA, int NW, int MINB>
int launch_fwd_win(const void *feat, const int *shape, const int *start, const float *loc,
                   const float *w, float *out, const Dims &d, cudaStream_t st) {
  auto kern = dfa_fwd_win_kernel<T, VPL, TMA, NW, MINB>;
  // A grid of about one wave is bound by latency: fetch the whole weights block with the locations
  // instead of the valid samples' weights next to the row loads.
  const long long grid = static_cast<long long>(d.bs) * d.A;
  const int whole = TMA ? DFA_KNOB("DFA_FWD_WHOLE_WEIGHTS", grid <= 148 * 8 ? 1 : 0) : 0;
  const int merge_pix = DFA_KNOB("DFA_FWD_MERGE_PIX", WIN_MAP);
  const WinLayout lay = win_layout(d.P, d.K, d.L, d.G, d.C, NW, whole != 0);
  if (int rc = set_smem(kern, lay.total)) return rc;
  // Channel split of the anchors that start last (see the kernel): DFA_FWD_TAIL_SPLIT = how many
  // waves of resident CTAs, in tenths, run split at the end of the grid.
  long long split_from = grid;
  if (VPL == 2) {
    const long long slots = 148ll * MINB;
    const long long n = slots * DFA_KNOB("DFA_FWD_TAIL_SPLIT", 10) / 10 / 2;  // anchors -> two CTAs each
    if (grid > 2 * slots && n > 0) split_from = grid - n;
  }
  kern<<<static_cast<unsigned int>(split_from + 2 * (grid - split_from)), NW * 32, lay.total, st>>>(
      static_cast<const T *>(feat), shape, start, loc, w, out, d, merge_pix, whole, static_cast<int>(split_from));
  return static_cast<int>(cudaGetLastError());
}

template <typename T>
int forward_typed(const void *feat, const int *shape, const int *start, const float *loc,
                  const float *w, float *out, const Dims &d, cudaStream_t st) {
  // DFA_FWD_VARIANT (tuning knob): 1..4 = row-sliced kernel with (threads, taps in flight) =
  // (256,1) (256,2) (512,1) (512,2); 30..33 = warp-autonomous window-merging kernel with (warps,
  // CTAs per SM) = (4,8) (8,4) (4,7) (2,16); 0 = one-warp-per-group kernel.  A variant whose shape
  // constraints are not met falls through to the next family.
  // Defaults, each the fastest measured on B200 (tools/sweep_fwd.py, tools/op_sweep.py, profiles/):
  //  * the row-sliced kernel with one tap in flight per thread, fp32 and bfloat16 tables alike;
  //  * with many potential taps per anchor (more key points than SimPB's 13: P*K*L >= 400) the chains
  //    of dependent load rounds are long enough that two taps in flight at 4 CTAs per SM win (26.7
  //    vs 32.0 us at 20 key points, 38.2 vs 42.2 us at 32; 900 anchors, fp32);
  //  * many waves of anchors on a table that mostly lives in L2 (fp32 rows of 1 KB, at least 7,000
  //    anchors, at most 128 MB per batch item — the training shapes): the window-merging kernel with
  //    two warps per anchor.  There the path is bound by the L2 -> SM fabric, and merging the coarse
  //    levels' taps moves 23 % fewer bytes over it: 88.0 vs 94.7 us at bs=8 x 900 anchors, 111.8 vs
  //    120.8 us at 8 x 1,220.  On small grids its longer per-anchor chain loses (bs=1: 24 vs 18 us),
  //    and R101 maps (368 MB per item, DRAM-bound) are par: both stay on the row-sliced kernel.
  const bool long_chains = static_cast<long long>(d.P) * d.K * d.L >= 400;
  int dflt = long_chains ? 2 : 1;
  if (!long_chains && sizeof(T) == 4 && static_cast<long long>(d.bs) * d.A >= 7000 &&
      static_cast<long long>(d.num_feat) * d.C * 4 <= (128ll << 20) && win_vpl<T>(d, feat) == 2)
    dflt = 33;
  const int variant = DFA_KNOB("DFA_FWD_VARIANT", dflt);
  if (variant >= 30 && variant < 40) {  // warp-autonomous window-merging kernel
    const int vpl = win_vpl<T>(d, feat);
    if (vpl) {
      const bool tma = tma_ok(d, loc, w);
#define WIN(NW, MINB)                                                                             \
  (vpl == 2 ? (tma ? launch_fwd_win<T, 2, true, NW, MINB>(feat, shape, start, loc, w, out, d, st)    \
                   : launch_fwd_win<T, 2, false, NW, MINB>(feat, shape, start, loc, w, out, d, st))  \
            : (tma ? launch_fwd_win<T, 1, true, NW, MINB>(feat, shape, start, loc, w, out, d, st)    \
                   : launch_fwd_win<T, 1, false, NW, MINB>(feat, shape, start, loc, w, out, d, st)))
      switch (variant) {
        case 31: return WIN(8, 4);
        case 32: return WIN(4, 7);
        case 33: return WIN(2, 16);
        default: return WIN(4, 8);
      }
#undef WIN
    }
  }
  const int rvariant = variant >= 5 ? 1 : variant;
  if (rvariant >= 1) {
    const int nt = rvariant >= 3 ? 512 : 256;
    const int vpr = rows_vpr<T>(d, feat, nt);
    if (vpr) {
      const bool tma = tma_ok(d, loc, w);
#define ROWS(U, NT, MINB)                                                                        \
  (tma ? launch_fwd_rows<T, U, true, NT, MINB>(feat, shape, start, loc, w, out, d, vpr, st)        \
       : launch_fwd_rows<T, U, false, NT, MINB>(feat, shape, start, loc, w, out, d, vpr, st))
      switch (rvariant) {
        case 1: return ROWS(1, 256, 6);
        case 2: return ROWS(2, 256, 4);
        case 3: return ROWS(1, 512, 3);
        default: return ROWS(2, 512, 2);
      }
#undef ROWS
    }
  }
  const int lpg = fast_lpg<T>(d, feat);
  if (lpg && aligned(out, 16)) {
    const bool tma = tma_ok(d, loc, w);
    const bool small = 32 * d.G <= 256;
#define CALL_FWD(N)                                                                            \
  (tma ? (small ? launch_fwd_t<T, N, true, 256>(feat, shape, start, loc, w, out, d, st)        \
                : launch_fwd_t<T, N, true, 1024>(feat, shape, start, loc, w, out, d, st))      \
       : (small ? launch_fwd_t<T, N, false, 256>(feat, shape, start, loc, w, out, d, st)       \
                : launch_fwd_t<T, N, false, 1024>(feat, shape, start, loc, w, out, d, st)))
    DFA_DISPATCH_LPG(CALL_FWD)
#undef CALL_FWD
  }
  auto kern = dfa_fwd_generic_kernel<T>;
  const SmemLayout lay = smem_layout(d.P, d.K, d.L, d.G, 1, false);
  if (int rc = set_smem(kern, lay.total)) return rc;
  const int threads = d.C >= 256 ? 256 : ((d.C + 31) / 32) * 32;
  kern<<<d.bs * d.A, threads, lay.total, st>>>(static_cast<const T *>(feat), shape, start, loc, w,
                                              out, d);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace

extern "C" {

int dfa_version(void) { return DFA_B200_VERSION; }

void dfa_debug_reload_knobs(void) { dfa_knob_generation.fetch_add(1, std::memory_order_relaxed); }

#ifdef DFA_PHASE_TIMING
int dfa_debug_set_phase_buffer(long long *buf) {
  return static_cast<int>(cudaMemcpyToSymbol(g_phase_buf, &buf, sizeof(buf)));
}
#endif

const char *dfa_error_string(int code) {
  switch (code) {
    case 0: return "success";
    case DFA_ERR_NULL_POINTER: return "dfa: null pointer argument";
    case DFA_ERR_BAD_DIMS: return "dfa: bad dimensions (non-positive, C % G != 0, or index overflow)";
    case DFA_ERR_BAD_DTYPE: return "dfa: unsupported feature dtype";
    case DFA_ERR_MISALIGNED: return "dfa: pointer not aligned for its element type";
    case DFA_ERR_UNSUPPORTED: return "dfa: configuration not supported (shared memory budget)";
    default: return code > 0 ? cudaGetErrorString(static_cast<cudaError_t>(code)) : "dfa: unknown error";
  }
}

int dfa_forward(const void *mc_ms_feat, int feat_dtype, const int32_t *spatial_shape,
                const int32_t *scale_start_index, const float *sampling_location,
                const float *weights, float *output, const dfa_dims *dims, void *stream) {
  if (!mc_ms_feat || !spatial_shape || !scale_start_index || !sampling_location || !weights || !output)
    return DFA_ERR_NULL_POINTER;
  Dims d;
  if (int rc = check_dims(dims, d)) return rc;
  if (!aligned(sampling_location, 4) || !aligned(weights, 4) || !aligned(output, 4) ||
      !aligned(spatial_shape, 4) || !aligned(scale_start_index, 4))
    return DFA_ERR_MISALIGNED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (feat_dtype == DFA_F32) {
    if (!aligned(mc_ms_feat, 4)) return DFA_ERR_MISALIGNED;
    return forward_typed<float>(mc_ms_feat, spatial_shape, scale_start_index, sampling_location,
                                weights, output, d, st);
  }
  if (feat_dtype == DFA_BF16) {
    if (!aligned(mc_ms_feat, 2)) return DFA_ERR_MISALIGNED;
    return forward_typed<__nv_bfloat16>(mc_ms_feat, spatial_shape, scale_start_index,
                                        sampling_location, weights, output, d, st);
  }
  return DFA_ERR_BAD_DTYPE;
}

int dfa_forward_fused(const void *mc_ms_feat, int feat_dtype, const int32_t *spatial_shape,
                      const int32_t *scale_start_index, const float *anchor, const float *fix_scale,
                      int num_fix, const float *learnable_logits, const float *projection_mat,
                      const float *image_wh, const float *logits_anchor, const float *logits_cam,
                      float *output, float *sampling_location_out, const dfa_dims *dims, void *stream) {
  if (!mc_ms_feat || !spatial_shape || !scale_start_index || !anchor || !fix_scale || !projection_mat ||
      !logits_anchor || !output)
    return DFA_ERR_NULL_POINTER;
  Dims d;
  if (int rc = check_dims(dims, d)) return rc;
  if (num_fix < 0 || num_fix > d.P) return DFA_ERR_BAD_DIMS;
  if (num_fix < d.P && !learnable_logits) return DFA_ERR_NULL_POINTER;
  constexpr int NT = 256;
  if (d.G > 32 || (d.G & (d.G - 1)) != 0) return DFA_ERR_UNSUPPORTED;  // group reductions by shuffles
  {  // the logits are staged by TMA bulk copies: 16-byte sized / aligned blocks, mbarrier-countable
    const long long lpg = static_cast<long long>(d.L) * d.P * d.G;
    if ((4 * lpg) % 16 != 0 || 4 * lpg * (d.K + 1) >= (1ll << 20) || !aligned(logits_anchor, 16) ||
        (logits_cam && !aligned(logits_cam, 16)))
      return DFA_ERR_UNSUPPORTED;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const FusedArgs fa{anchor, fix_scale, learnable_logits, projection_mat, image_wh, logits_anchor, logits_cam,
                     sampling_location_out, num_fix};
  auto launch = [&](auto tag) -> int {
    using T = decltype(tag);
    const int vpr = rows_vpr<T>(d, mc_ms_feat, NT);
    if (!vpr) return DFA_ERR_UNSUPPORTED;
    int vpr_log2 = 0;
    while ((1 << vpr_log2) < vpr) ++vpr_log2;
    auto kern = dfa_fwd_fused_kernel<T, NT, 6>;
    long long split_from = 0, grid = 0;
    int split_log2 = 0;
    last_wave_split<T>(d, vpr, 6, split_from, split_log2, grid);
    SmemLayoutF lay = smem_layout_fused(d.P, d.K, d.L, d.G, d.C, NT / vpr, NT);
    const SmemLayoutF lay_s = smem_layout_fused(d.P, d.K, d.L, d.G, d.C >> split_log2, (NT / vpr) << split_log2, NT);
    if (lay_s.total > lay.total) lay = lay_s;
    if (int rc = set_smem(kern, lay.total)) return rc;
    kern<<<static_cast<unsigned int>(grid), NT, lay.total, st>>>(static_cast<const T *>(mc_ms_feat), spatial_shape,
                                                                 scale_start_index, fa, output, d, vpr_log2,
                                                                 static_cast<int>(split_from), split_log2);
    return static_cast<int>(cudaGetLastError());
  };
  if (feat_dtype == DFA_F32) return launch(float{});
  if (feat_dtype == DFA_BF16) return launch(__nv_bfloat16{});
  return DFA_ERR_BAD_DTYPE;
}

int dfa_debug_indices(const int32_t *spatial_shape, const int32_t *scale_start_index,
                      const float *sampling_location, uint8_t *valid, int32_t *corner_rows,
                      const dfa_dims *dims, void *stream) {
  if (!spatial_shape || !scale_start_index || !sampling_location || !valid || !corner_rows)
    return DFA_ERR_NULL_POINTER;
  Dims d;
  if (int rc = check_dims(dims, d)) return rc;
  const long long n = static_cast<long long>(d.bs) * d.A * d.P * d.K;
  const int blocks = static_cast<int>((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  dfa_debug_indices_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      spatial_shape, scale_start_index, sampling_location, valid, corner_rows, d);
  return static_cast<int>(cudaGetLastError());
}

// Device pointer aliasing a pinned, mapped host buffer (NULL when the buffer is pageable or not mapped).
static const void *host_alias(const void *h) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, h) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return (at.type == cudaMemoryTypeHost && at.devicePointer) ? at.devicePointer : nullptr;
}

int64_t dfa_forward_host_workspace_bytes(int feat_dtype, const dfa_dims *dims) {
  Dims d;
  if (check_dims(dims, d)) return -1;
  const int64_t esz = feat_dtype == DFA_BF16 ? 2 : 4;
  auto up = [](int64_t x) { return (x + 255) / 256 * 256; };
  const int64_t rows = static_cast<int64_t>(d.bs) * d.num_feat;
  return up(esz * d.bs * d.num_feat * d.C) + up(8ll * d.K * d.L) + up(4ll * d.K * d.L) +
         up(8ll * d.bs * d.A * d.P * d.K) + up(4ll * d.bs * d.A * d.P * d.K * d.L * d.G) +
         up(4ll * d.bs * d.A * d.C) + 256 + up((rows + 31) / 32 * 4) + up(1ll * d.bs * d.A * d.P * d.K);
}

int dfa_forward_host(const void *h_feat, int feat_dtype, const int32_t *h_shape,
                     const int32_t *h_start, const float *h_loc, const float *h_w, float *h_out,
                     const dfa_dims *dims, void *workspace, int64_t workspace_bytes, void *stream) {
  if (!h_feat || !h_shape || !h_start || !h_loc || !h_w || !h_out || !workspace)
    return DFA_ERR_NULL_POINTER;
  Dims d;
  if (int rc = check_dims(dims, d)) return rc;
  if (feat_dtype != DFA_F32 && feat_dtype != DFA_BF16) return DFA_ERR_BAD_DTYPE;
  if (workspace_bytes < dfa_forward_host_workspace_bytes(feat_dtype, dims)) return DFA_ERR_BAD_DIMS;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t esz = feat_dtype == DFA_BF16 ? 2 : 4;
  auto up = [](int64_t x) { return (x + 255) / 256 * 256; };
  char *p = static_cast<char *>(workspace);
  const int64_t nb_feat = esz * d.bs * d.num_feat * d.C, nb_shape = 8ll * d.K * d.L,
                nb_start = 4ll * d.K * d.L, nb_loc = 8ll * d.bs * d.A * d.P * d.K,
                nb_w = 4ll * d.bs * d.A * d.P * d.K * d.L * d.G, nb_out = 4ll * d.bs * d.A * d.C;
  const int64_t rows = static_cast<int64_t>(d.bs) * d.num_feat, nwords = (rows + 31) / 32;
  const int64_t nsamples = static_cast<int64_t>(d.bs) * d.A * d.P * d.K;
  void *d_feat = p; p += up(nb_feat);
  int32_t *d_shape = reinterpret_cast<int32_t *>(p); p += up(nb_shape);
  int32_t *d_start = reinterpret_cast<int32_t *>(p); p += up(nb_start);
  float *d_loc = reinterpret_cast<float *>(p); p += up(nb_loc);
  float *d_w = reinterpret_cast<float *>(p); p += up(nb_w);
  float *d_out = reinterpret_cast<float *>(p); p += up(nb_out);
  unsigned long long *d_hdr = reinterpret_cast<unsigned long long *>(p); p += 256;
  uint32_t *d_bitmap = reinterpret_cast<uint32_t *>(p); p += up(nwords * 4);
  uint8_t *d_svalid = reinterpret_cast<uint8_t *>(p);
  cudaError_t e;
  // Pull mode (default when it applies; DFA_HOST_PULL=0 forces whole copies): the feature table is
  // pinned + mapped host memory and a row is a whole number of 16-byte vectors.
  const int64_t row_bytes = esz * d.C, line_bytes = 4ll * d.L * d.G;
  const void *a_feat = (DFA_KNOB("DFA_HOST_PULL", 1) && row_bytes % 16 == 0 && aligned(h_feat, 16))
                           ? host_alias(h_feat) : nullptr;
  const void *a_w = (a_feat && line_bytes % 16 == 0 && aligned(h_w, 16)) ? host_alias(h_w) : nullptr;
  // small operands first so the kernel's staging data is resident before the big transfer ends
  if ((e = cudaMemcpyAsync(d_shape, h_shape, nb_shape, cudaMemcpyHostToDevice, st))) return e;
  if ((e = cudaMemcpyAsync(d_start, h_start, nb_start, cudaMemcpyHostToDevice, st))) return e;
  if ((e = cudaMemcpyAsync(d_loc, h_loc, nb_loc, cudaMemcpyHostToDevice, st))) return e;
  if (a_feat) {
    const int sms = device_sm_count() > 0 ? device_sm_count() : 132;
    if ((e = cudaMemsetAsync(d_hdr, 0, 256 + up(nwords * 4), st))) return e;  // header + bitmap (adjacent)
    const unsigned long long fixed = nb_shape + nb_start + nb_loc + (a_w ? 0 : nb_w);
    const int mark_blocks = static_cast<int>((nsamples + 255) / 256 < sms * 8 ? (nsamples + 255) / 256 : sms * 8);
    dfa_host_mark_kernel<<<mark_blocks, 256, 0, st>>>(d_shape, d_start, d_loc, d_bitmap, d_svalid, d_hdr,
                                                      static_cast<unsigned long long>(row_bytes), fixed, d);
    if ((e = cudaGetLastError())) return e;
    if (a_w) {
      const int64_t nvecs = nb_w / 16;
      const int wb = static_cast<int>((nvecs + 255) / 256 < sms * 16 ? (nvecs + 255) / 256 : sms * 16);
      dfa_host_pull_weights_kernel<<<wb, 256, 0, st>>>(static_cast<const uint4 *>(a_w), reinterpret_cast<uint4 *>(d_w),
                                                       d_svalid, nvecs, static_cast<int>(line_bytes / 16), d_hdr);
      if ((e = cudaGetLastError())) return e;
    } else if ((e = cudaMemcpyAsync(d_w, h_w, nb_w, cudaMemcpyHostToDevice, st))) {
      return e;
    }
    const int rb = static_cast<int>((nwords + 7) / 8 < sms * 8 ? (nwords + 7) / 8 : sms * 8);
    dfa_host_pull_rows_kernel<<<rb, 256, 0, st>>>(static_cast<const uint4 *>(a_feat), static_cast<uint4 *>(d_feat),
                                                  d_bitmap, nwords, rows, static_cast<int>(row_bytes / 16), d_hdr);
    if ((e = cudaGetLastError())) return e;
  } else {
    const unsigned long long hdr[4] = {static_cast<unsigned long long>(rows), static_cast<unsigned long long>(nb_w / 16),
                                       static_cast<unsigned long long>(row_bytes),
                                       static_cast<unsigned long long>(nb_shape + nb_start + nb_loc)};
    if ((e = cudaMemcpyAsync(d_hdr, hdr, sizeof(hdr), cudaMemcpyHostToDevice, st))) return e;
    if ((e = cudaMemcpyAsync(d_w, h_w, nb_w, cudaMemcpyHostToDevice, st))) return e;
    if ((e = cudaMemcpyAsync(d_feat, h_feat, nb_feat, cudaMemcpyHostToDevice, st))) return e;
  }
  if (int rc = dfa_forward(d_feat, feat_dtype, d_shape, d_start, d_loc, d_w, d_out, dims, stream))
    return rc;
  if ((e = cudaMemcpyAsync(h_out, d_out, nb_out, cudaMemcpyDeviceToHost, st))) return e;
  return static_cast<int>(cudaStreamSynchronize(st));
}

int dfa_forward_host_stats(const void *workspace, int feat_dtype, const dfa_dims *dims, void *stream,
                           int64_t *h2d_bytes, int64_t *rows_moved, int64_t *weight_bytes_moved) {
  if (!workspace || !h2d_bytes) return DFA_ERR_NULL_POINTER;
  Dims d;
  if (int rc = check_dims(dims, d)) return rc;
  const int64_t esz = feat_dtype == DFA_BF16 ? 2 : 4;
  auto up = [](int64_t x) { return (x + 255) / 256 * 256; };
  const char *p = static_cast<const char *>(workspace);
  p += up(esz * d.bs * d.num_feat * d.C) + up(8ll * d.K * d.L) + up(4ll * d.K * d.L) +
       up(8ll * d.bs * d.A * d.P * d.K) + up(4ll * d.bs * d.A * d.P * d.K * d.L * d.G) + up(4ll * d.bs * d.A * d.C);
  unsigned long long hdr[4];
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  if ((e = cudaMemcpyAsync(hdr, p, sizeof(hdr), cudaMemcpyDeviceToHost, st))) return e;
  if ((e = cudaStreamSynchronize(st))) return e;
  *h2d_bytes = static_cast<int64_t>(hdr[0] * hdr[2] + hdr[1] * 16 + hdr[3]);
  if (rows_moved) *rows_moved = static_cast<int64_t>(hdr[0]);
  if (weight_bytes_moved) *weight_bytes_moved = static_cast<int64_t>(hdr[1] * 16);
  return 0;
}

}  // extern "C"

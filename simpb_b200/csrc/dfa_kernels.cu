// dfa_kernels.cu — deformable feature aggregation for SimPB, hand-written for sm_100a (B200).
//
// One CTA owns one anchor (b, a).  Its sampling locations (P*K*2 floats) and weights
// (P*K*L*G floats) are contiguous per anchor and are staged into shared memory with two
// TMA bulk copies (cp.async.bulk → UBLKCP) completing on mbarriers.  Warp 0 compacts the
// samples that pass the op's exclusive (0,1) test; the CTA then builds one 32-byte "tap"
// record per (valid sample, level): the four corner element offsets and bilinear weights.
// In the main loop warp g owns channel group g: its lanes are split as
//     lane = [sub-tap s][corner q][16-byte vector j]
// so one 128-bit load per lane fetches the group's contiguous channels of all four corners
// (fp32: 8 lanes x 16 B = the group's 128 B per corner).  The weighted sum lives in
// registers; corners are folded with two warp shuffles at the end and the anchor's output row
// is written once with plain vector stores — no atomics and no zero-filled output.
//
// The backward keeps the same ownership: grad_weights[b,a,p,k,l,g] is produced by exactly one
// warp (shuffle reduction, plain store) and grad_sampling_location[b,a,p,k,:] by exactly one
// CTA (per-warp shared-memory rows, summed in a fixed order).  Only grad_mc_ms_feat, where
// different anchors meet on the same pixel, is scattered — with 128-bit vector reductions
// (red.global.add.v4.f32).
//
// Semantics follow /root/reference/projects/mmdet3d_plugin/ops/src/deformable_aggregation_cuda.cu
// (forward :129-187 + :13-59, backward :190-262 + :62-126); the pixel coordinate uses the single
// fused multiply-add the compiled reference uses (SURVEY.md §7 "bit-exact indices").
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "dfa_b200.h"

namespace {

// ------------------------------------------------------------------------------------------
// small PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA 1-D bulk copy global → shared, completion counted in bytes on an mbarrier.
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes,
                                             uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// TMA bulk prefetch of `bytes` (multiple of 16) into L2: no registers, no shared memory.
__device__ __forceinline__ void tma_prefetch_l2(const void *src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

// ------------------------------------------------------------------------------------------
// geometry shared by every kernel (and by the debug side channel the parity tests read)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool sample_valid(float x, float y) {
  // …_cuda.cu:168-171: `if (loc <= 0 || loc >= 1) return;` — exclusive on both ends
  return !(x <= 0.f || x >= 1.f) && !(y <= 0.f || y >= 1.f);
}

struct TapGeom {
  int row[4];  // row inside the batch item's feature table, -1 when the corner is outside
  float lh, lw, hh, hw;
};

__device__ __forceinline__ void tap_geometry(float x, float y, int H, int W, int start, TapGeom &g) {
  // …_cuda.cu:180-181 as compiled: one FFMA(loc, size, -0.5), then floor (:18-25)
  const float h_im = fmaf(y, static_cast<float>(H), -0.5f);
  const float w_im = fmaf(x, static_cast<float>(W), -0.5f);
  const float fh = floorf(h_im), fw = floorf(w_im);
  const int h_low = static_cast<int>(fh), w_low = static_cast<int>(fw);
  g.lh = h_im - fh;
  g.lw = w_im - fw;
  g.hh = 1.f - g.lh;
  g.hw = 1.f - g.lw;
  const bool hl = h_low >= 0, wl = w_low >= 0;            // :33, :38, :43, :48
  const bool hh = h_low + 1 <= H - 1, wh = w_low + 1 <= W - 1;
  const int base = start + h_low * W + w_low;
  g.row[0] = (hl && wl) ? base : -1;
  g.row[1] = (hl && wh) ? base + 1 : -1;
  g.row[2] = (hh && wl) ? base + W : -1;
  g.row[3] = (hh && wh) ? base + W + 1 : -1;
}

// ------------------------------------------------------------------------------------------
// feature vector access: one 16-byte load = VEC channels
// ------------------------------------------------------------------------------------------
template <typename T>
struct FeatVec;
template <>
struct FeatVec<float> {
  static constexpr int VEC = 4;
  typedef float4 raw_t;
  __device__ static __forceinline__ raw_t load_raw(const float *p) {
    return __ldg(reinterpret_cast<const float4 *>(p));
  }
  __device__ static __forceinline__ raw_t zero_raw() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ static __forceinline__ void unpack(const raw_t &t, float (&v)[4]) {
    v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
  }
  __device__ static __forceinline__ void load(const float *p, float (&v)[4]) { unpack(load_raw(p), v); }
  // acc += c * raw, packed FFMA2 (sm_100 fma.rn.f32x2)
  __device__ static __forceinline__ void fma(float (&acc)[4], float c, const raw_t &t) {
    const float2 cc = make_float2(c, c);
    float2 a0 = __ffma2_rn(cc, make_float2(t.x, t.y), make_float2(acc[0], acc[1]));
    float2 a1 = __ffma2_rn(cc, make_float2(t.z, t.w), make_float2(acc[2], acc[3]));
    acc[0] = a0.x, acc[1] = a0.y, acc[2] = a1.x, acc[3] = a1.y;
  }
};
template <>
struct FeatVec<__nv_bfloat16> {
  static constexpr int VEC = 8;
  typedef uint4 raw_t;
  __device__ static __forceinline__ raw_t load_raw(const __nv_bfloat16 *p) {
    return __ldg(reinterpret_cast<const uint4 *>(p));
  }
  __device__ static __forceinline__ raw_t zero_raw() { return make_uint4(0u, 0u, 0u, 0u); }
  __device__ static __forceinline__ void unpack(const raw_t &t, float (&v)[8]) {
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // bf16 → fp32 is a 16-bit shift
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ static __forceinline__ void load(const __nv_bfloat16 *p, float (&v)[8]) {
    unpack(load_raw(p), v);
  }
  __device__ static __forceinline__ void fma(float (&acc)[8], float c, const raw_t &t) {
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
    const float2 cc = make_float2(c, c);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 x = make_float2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xffff0000u));
      const float2 a = __ffma2_rn(cc, x, make_float2(acc[2 * i], acc[2 * i + 1]));
      acc[2 * i] = a.x, acc[2 * i + 1] = a.y;
    }
  }
};

struct Dims {
  int bs, K, num_feat, C, L, A, P, G;
};

// shared-memory carve-up, identical on host and device
struct SmemLayout {
  uint32_t w, loc, rec, widx, list, gl, bar, total;
};
__host__ __device__ inline uint32_t align_up(uint32_t x, uint32_t a) { return (x + a - 1) / a * a; }
__host__ __device__ inline SmemLayout smem_layout(int P, int K, int L, int G, int tap_pad,
                                                  bool backward) {
  SmemLayout s;
  const uint32_t taps = align_up(static_cast<uint32_t>(P) * K * L, tap_pad) + tap_pad;
  uint32_t o = 0;
  s.w = o, o = align_up(o + 4u * P * K * L * G, 16);
  s.loc = o, o = align_up(o + 8u * P * K, 16);
  s.rec = o, o = align_up(o + 32u * taps, 16);
  s.widx = o, o = align_up(o + 4u * taps, 16);
  s.list = o, o = align_up(o + 4u * P * K, 16);
  s.gl = o, o = align_up(o + (backward ? 8u * P * K * G : 0u), 16);
  s.bar = o, o += 32;
  s.total = o;
  return s;
}

struct TapQ {  // forward record, one per corner
  int off;     // element offset of the corner's row inside the batch item, -1 = skip
  float bw;    // bilinear weight
};
struct TapB {  // backward record, one per tap
  int off[4];
  float lh, lw, Wf, Hf;
};

// Stage the anchor's locations and weights, compact valid samples.  Returns n_valid.
template <bool TMA>
__device__ __forceinline__ int stage_and_compact(const float *__restrict__ loc_g,
                                                 const float *__restrict__ w_g, float *s_w,
                                                 float *s_loc, int *s_list, uint64_t *bars,
                                                 int *s_nvalid, int PK, int wcount) {
  const int tid = threadIdx.x;
  if (TMA) {
    if (tid == 0) {
      mbar_init(&bars[0], 1);
      mbar_init(&bars[1], 1);
      fence_mbar_init();
      mbar_expect_tx(&bars[0], 8u * PK);
      tma_bulk_g2s(s_loc, loc_g, 8u * PK, &bars[0]);
      mbar_expect_tx(&bars[1], 4u * wcount);
      tma_bulk_g2s(s_w, w_g, 4u * wcount, &bars[1]);
    }
    __syncthreads();  // barrier init visible to every waiter
    if (tid < 32) mbar_wait(&bars[0], 0);
  } else {
    for (int i = tid; i < 2 * PK; i += blockDim.x) s_loc[i] = __ldg(loc_g + i);
    for (int i = tid; i < wcount; i += blockDim.x) s_w[i] = __ldg(w_g + i);
    __syncthreads();
  }
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
  __syncthreads();
  return *s_nvalid;
}

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
  s.widx = o, o = align_up(o + 4u * taps, 16);
  s.list = o, o = align_up(o + 4u * P * K, 16);
  s.tab = o, o = align_up(o + 12u * K * L, 16);
  s.red = o, o = align_up(o + 4u * slices * C, 16);
  s.bar = o, o += 32;
  s.total = o;
  return s;
}

template <typename T, int U, bool TMA, int NT, int MINB, bool PF>
__global__ void __launch_bounds__(NT, MINB)
    dfa_fwd_rows_kernel(const T *__restrict__ feat, const int *__restrict__ shape,
                        const int *__restrict__ start, const float *__restrict__ loc,
                        const float *__restrict__ weights, float *__restrict__ out, Dims d,
                        int vpr_log2) {
  constexpr int VEC = FeatVec<T>::VEC;
  extern __shared__ __align__(128) unsigned char smem[];
  const int vpr = 1 << vpr_log2;
  const int slices = NT >> vpr_log2;
  const SmemLayout2 lay = smem_layout2(d.P, d.K, d.L, d.G, d.C, slices, slices * U);
  float *s_w = reinterpret_cast<float *>(smem + lay.w);
  float *s_loc = reinterpret_cast<float *>(smem + lay.loc);
  uint4 *s_off = reinterpret_cast<uint4 *>(smem + lay.off);
  float4 *s_bw = reinterpret_cast<float4 *>(smem + lay.bw);
  int *s_widx = reinterpret_cast<int *>(smem + lay.widx);
  int *s_list = reinterpret_cast<int *>(smem + lay.list);
  int *s_tab = reinterpret_cast<int *>(smem + lay.tab);
  float *s_red = reinterpret_cast<float *>(smem + lay.red);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + lay.bar);
  int *s_nvalid = reinterpret_cast<int *>(bars + 2);

  const int tid = threadIdx.x;
  const int anchor = blockIdx.x;  // b * A + a
  const int b = anchor / d.A;
  const int PK = d.P * d.K, wcount = PK * d.L * d.G;

  // level tables → shared memory while the TMA copies are in flight
  for (int i = tid; i < d.K * d.L; i += NT) {
    s_tab[3 * i] = __ldg(shape + 2 * i);
    s_tab[3 * i + 1] = __ldg(shape + 2 * i + 1);
    s_tab[3 * i + 2] = __ldg(start + i);
  }
  const int nv = stage_and_compact<TMA>(loc + static_cast<size_t>(anchor) * PK * 2,
                                        weights + static_cast<size_t>(anchor) * wcount, s_w,
                                        s_loc, s_list, bars, s_nvalid, PK, wcount);
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
    s_off[t] = off, s_bw[t] = bw, s_widx[t] = (s * d.L + l) * d.G;
    if (PF && t < ntaps) {
      // start the rows' DRAM → L2 transfers now, long before the first register load needs them
      const unsigned char *fr = reinterpret_cast<const unsigned char *>(feat) +
                                static_cast<size_t>(b) * d.num_feat * rb;
      if (gm.row[0] >= 0) tma_prefetch_l2(fr + off.x, rb);
      if (gm.row[1] >= 0) tma_prefetch_l2(fr + off.y, rb);
      if (gm.row[2] >= 0) tma_prefetch_l2(fr + off.z, rb);
      if (gm.row[3] >= 0) tma_prefetch_l2(fr + off.w, rb);
    }
  }
  __syncthreads();
  if (TMA) mbar_wait(&bars[1], 0);  // weights have landed

  const int slice = tid >> vpr_log2, v = tid & (vpr - 1);
  const int ch = v * VEC;
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
    float4 *r = reinterpret_cast<float4 *>(s_red + slice * d.C + ch);
#pragma unroll
    for (int c = 0; c < VEC / 4; ++c)
      r[c] = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
  }
  __syncthreads();
  for (int c = tid; c < d.C; c += NT) {
    float sum = 0.f;
    for (int sl = 0; sl < slices; ++sl) sum += s_red[sl * d.C + c];
    out[static_cast<size_t>(anchor) * d.C + c] = sum;
  }
}


// ------------------------------------------------------------------------------------------
// forward with row merging and a balanced gather (the default for SimPB's shapes)
// ------------------------------------------------------------------------------------------
// What bounds the gather on B200 is not HBM but the SM's load path — every 128 bytes a warp pulls
// into registers is one L1 wavefront, hit or miss (ncu: l1tex data pipe 71 % busy on the
// row-sliced kernel while DRAM sits at 34 %) — and, at one batch item, the length of each warp's
// dependent chain.  So this kernel pulls fewer rows and keeps every warp's loads independent.
// The key points of one anchor project close together: at the coarse levels their bilinear
// corners name the same feature rows again and again (camera-rig inputs: 230 corner references
// per anchor, 106 distinct rows).  They are merged without sorting or atomics:
//
//   prologue (warp 0)  the anchor's sampling locations arrive by one TMA bulk copy; the warp
//                      compacts the samples that pass the op's (0,1) test and issues the weights
//                      as TMA bulk copies too — one 128-byte line per VALID sample when the anchor
//                      is sparse (a sample's L*G weights are contiguous), the whole block when it
//                      is dense or when the grid is so small that latency, not bandwidth, rules.
//   merge              warp l owns level l.  A chunk = up to 8 valid samples = 32 corner references,
//                      one per lane.  __match_any_sync groups lanes that name the same row; the
//                      lowest lane of each group sums the group's (bilinear weight x group weight)
//                      coefficients with shuffles in ascending lane order.  Across the chunks of
//                      the level a small direct-mapped table (row -> slot) in shared memory lets a
//                      leader find a row an earlier chunk already listed and add to its
//                      coefficients instead.  Result: a list of (row offset, coef[G]) per warp.
//   gather             after one barrier every warp takes the same share of every list (slot p of
//                      list c goes to warp (p + c) mod NW), so the warps finish together.  A lane
//                      owns VPL 16-byte vectors of a row: one warp instruction covers 512 contiguous
//                      bytes, U rows are in flight per lane, and the weighted sum stays in
//                      registers (packed FFMA2).
//   epilogue           the warps' partial rows are folded through shared memory and the anchor's
//                      output row is written once.
//
// Merge and summation order are fixed, so results are bitwise reproducible.
#ifdef DFA_PHASE_TIMING
// Tool-only build (tools/phase_timing.py): per-warp clock64() stamps at the phase boundaries of the
// merging forward kernel, written to a caller-provided buffer [anchor][warp][8].
__device__ long long *g_phase_buf = nullptr;
#define DFA_STAMP(i)                                                                      \
  do {                                                                                    \
    if (g_phase_buf && lane == 0)                                                         \
      g_phase_buf[(static_cast<size_t>(blockIdx.x) * NW + warp) * 8 + (i)] = clock64();   \
  } while (0)
#else
#define DFA_STAMP(i) do {} while (0)
#endif

constexpr int MERGE_CAP = 64;     // slots per warp list and round (two chunks)
constexpr int MERGE_TABLE = 128;  // entries of a warp's row -> slot table
constexpr int MERGE_WPAD = 8;     // floats between the weight lines of a sparse anchor (bank spread)

struct MergeLayout {
  uint32_t w, loc, list, tab, rowoff, coef, table, cnt, mine_off, mine_slot, mine_stride, bar, total;
};
__host__ __device__ inline MergeLayout merge_layout(int P, int K, int L, int G, int NW, int U) {
  MergeLayout s;
  const uint32_t slots = static_cast<uint32_t>(NW) * MERGE_CAP;
  uint32_t o = 0;
  s.w = o, o = align_up(o + 4u * P * K * L * G, 128);
  s.loc = o, o = align_up(o + 8u * P * K, 16);
  s.list = o, o = align_up(o + 4u * P * K, 16);
  s.tab = o, o = align_up(o + 16u * K * L, 128);
  s.rowoff = o, o = align_up(o + 4u * slots, 128);
  s.coef = o, o = align_up(o + 4u * G * slots, 128);  // reused for the NW partial rows
  s.table = o, o = align_up(o + 4u * MERGE_TABLE * NW, 128);
  s.cnt = o, o = align_up(o + 4u * NW, 16);
  s.mine_stride = MERGE_CAP + U;  // entries per warp (a warp's share of NW lists, padded to U)
  s.mine_off = o, o = align_up(o + 4u * s.mine_stride * NW, 16);
  s.mine_slot = o, o = align_up(o + 2u * s.mine_stride * NW, 16);
  s.bar = o, o += 32;
  s.total = o;
  return s;
}

// T: feature type.  VPL: 16-byte vectors per lane per row (row bytes = 512 * VPL).  G: groups.
// NW: warps per CTA.  U: rows in flight per lane.
template <typename T, int VPL, int G, int NW, int U, bool TMA, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB)
    dfa_fwd_merge_kernel(const T *__restrict__ feat, const int *__restrict__ shape,
                         const int *__restrict__ start, const float *__restrict__ loc,
                         const float *__restrict__ weights, float *__restrict__ out, Dims d,
                         MergeLayout lay, int whole_weights, int prefetch) {
  constexpr int VEC = FeatVec<T>::VEC;
  constexpr int NT = NW * 32;
  constexpr int GPV = G / VPL;  // groups covered by one 512-byte segment of the row
  constexpr int C = 32 * VPL * VEC;
  static_assert(G % 4 == 0 && G % VPL == 0 && (32 * VPL) % G == 0,
                "a 16-byte vector must lie inside one group");
  static_assert((NW & (NW - 1)) == 0 && NW >= 2 && NW <= 32, "warps per CTA: a power of two");
  static_assert(MERGE_CAP * G >= C, "partial rows must fit the coefficient lists");
  static_assert(U == 2 || U == 4, "row offsets / slots of a batch are read with one vector load");
  extern __shared__ __align__(128) unsigned char smem[];
  float *s_w = reinterpret_cast<float *>(smem + lay.w);
  float *s_loc = reinterpret_cast<float *>(smem + lay.loc);
  uint32_t *s_list = reinterpret_cast<uint32_t *>(smem + lay.list);
  int4 *s_tab = reinterpret_cast<int4 *>(smem + lay.tab);
  uint32_t *s_rowoff = reinterpret_cast<uint32_t *>(smem + lay.rowoff);
  float *s_coef = reinterpret_cast<float *>(smem + lay.coef);
  int *s_cnt = reinterpret_cast<int *>(smem + lay.cnt);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + lay.bar);
  int *s_nvalid = reinterpret_cast<int *>(bars + 2);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int anchor = blockIdx.x;  // b * A + a
  const int b = anchor / d.A;
  const int PK = d.P * d.K, LG = d.L * G, wcount = PK * LG;
  const float *loc_g = loc + static_cast<size_t>(anchor) * PK * 2;
  const float *w_g = weights + static_cast<size_t>(anchor) * wcount;
  const unsigned lt_mask = (1u << lane) - 1u;

  // ---- prologue ------------------------------------------------------------------------------
  DFA_STAMP(0);
  if (TMA) {
    if (warp == 0) {
      if (lane == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
        mbar_expect_tx(&bars[0], 8u * PK);
        tma_bulk_g2s(s_loc, loc_g, 8u * PK, &bars[0]);
        if (whole_weights) {
          mbar_expect_tx(&bars[1], 4u * wcount);
          tma_bulk_g2s(s_w, w_g, 4u * wcount, &bars[1]);
        }
      }
      __syncwarp();
      if (NW == 1)
        for (int i = lane; i < d.K * d.L; i += 32)
          s_tab[i] = make_int4(__ldg(shape + 2 * i), __ldg(shape + 2 * i + 1), __ldg(start + i), 0);
      mbar_wait(&bars[0], 0);
    } else {
      for (int i = tid - 32; i < d.K * d.L; i += NT - 32)
        s_tab[i] = make_int4(__ldg(shape + 2 * i), __ldg(shape + 2 * i + 1), __ldg(start + i), 0);
    }
  } else {
    for (int i = tid; i < 2 * PK; i += NT) s_loc[i] = __ldg(loc_g + i);
    for (int i = tid; i < d.K * d.L; i += NT)
      s_tab[i] = make_int4(__ldg(shape + 2 * i), __ldg(shape + 2 * i + 1), __ldg(start + i), 0);
    __syncthreads();
  }
  DFA_STAMP(1);
  bool sparse_w = false;  // weight lines packed by valid-sample index (stride LG + MERGE_WPAD)
  if (warp == 0) {
    // compaction: entry = sample | camera << 16, in sample order
    const float rK = 1.0f / static_cast<float>(d.K);
    int n = 0;
    for (int base = 0; base < PK; base += 32) {
      const int s = base + lane;
      bool v = false;
      if (s < PK) {
        const float2 xy = *reinterpret_cast<const float2 *>(s_loc + 2 * s);
        v = sample_valid(xy.x, xy.y);
      }
      const unsigned m = __ballot_sync(0xffffffffu, v);
      if (v) {
        const int p = static_cast<int>((static_cast<float>(s) + 0.5f) * rK);  // exact: s < 65536
        s_list[n + __popc(m & lt_mask)] = static_cast<uint32_t>(s) | (static_cast<uint32_t>(s - p * d.K) << 16);
      }
      n += __popc(m);
    }
    sparse_w = TMA && !whole_weights && 2 * n <= PK;
    if (lane == 0) *s_nvalid = sparse_w ? -n - 1 : n;
    if (TMA && !whole_weights && n > 0) {
      if (!sparse_w) {  // dense anchor: one copy of the whole block
        if (lane == 0) {
          mbar_expect_tx(&bars[1], 4u * wcount);
          tma_bulk_g2s(s_w, w_g, 4u * wcount, &bars[1]);
        }
      } else {  // sparse anchor: one line per valid sample, packed by valid index
        if (lane == 0) mbar_expect_tx(&bars[1], 4u * LG * n);
        __syncwarp();
        for (int i = lane; i < n; i += 32) {
          const int s = s_list[i] & 0xffff;
          tma_bulk_g2s(s_w + i * (LG + MERGE_WPAD), w_g + s * LG, 4u * LG, &bars[1]);
        }
      }
    }
  }
  DFA_STAMP(2);
  __syncthreads();
  DFA_STAMP(3);
  int nv = *s_nvalid;
  sparse_w = nv < 0;
  nv = sparse_w ? -nv - 1 : nv;
  if (!TMA) {  // plain staging of the valid samples' weight lines
    for (int i = tid; i < nv * LG; i += NT) {
      const int s = s_list[i / LG] & 0xffff, r = i - (i / LG) * LG;
      s_w[s * LG + r] = __ldg(w_g + s * LG + r);
    }
    __syncthreads();
  }
  const int wstride = sparse_w ? LG + MERGE_WPAD : LG;

  const uint32_t rb = 512u * VPL;  // bytes per feature row
  const unsigned char *fb = reinterpret_cast<const unsigned char *>(feat) +
                            static_cast<size_t>(b) * d.num_feat * rb + lane * 16;
  const int gq = (lane * G) / (32 * VPL);  // this lane's group inside each 512-byte segment
  float acc[VPL][VEC];
#pragma unroll
  for (int v = 0; v < VPL; ++v)
#pragma unroll
    for (int c = 0; c < VEC; ++c) acc[v][c] = 0.f;

  // A warp's work items: (level, chunk of 8 valid samples) for its levels l = warp, warp + NW, ...
  const int cpl = (nv + 7) >> 3;  // chunks per level
  const int my_levels = warp < d.L ? (d.L - warp + NW - 1) / NW : 0;
  const int max_items = ((d.L + NW - 1) / NW) * cpl;  // warp 0 has the most
  uint32_t *s_table = reinterpret_cast<uint32_t *>(smem + lay.table) + warp * MERGE_TABLE;
  uint32_t *my_rowoff = s_rowoff + warp * MERGE_CAP;
  float *my_coef = s_coef + warp * MERGE_CAP * G;
  uint32_t *s_mine_off = reinterpret_cast<uint32_t *>(smem + lay.mine_off) + warp * lay.mine_stride;
  uint16_t *s_mine_slot = reinterpret_cast<uint16_t *>(smem + lay.mine_slot) + warp * lay.mine_stride;
  bool wready = !TMA;
  int li = 0, cj = 0;  // this warp's next item: its li-th level, chunk cj

  for (int it0 = 0; it0 < max_items; it0 += 2) {
    // ---- merge: this warp's next two chunks -------------------------------------------------------
    int cnt = 0;
    if (li < my_levels) {
#pragma unroll
      for (int x = 0; x < MERGE_TABLE / 128; ++x)
        reinterpret_cast<uint4 *>(s_table)[x * 32 + lane] =
            make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
      __syncwarp();
    }
    for (int x = 0; x < 2 && li < my_levels; ++x) {
      const int l = warp + li * NW;
      const int i = (cj << 3) + (lane >> 2), q = lane & 3;
      if (++cj == cpl) cj = 0, ++li;
      const bool live = i < nv;
      const int ii = live ? i : 0;
      const uint32_t ent = s_list[ii];
      const int s = ent & 0xffff, k = ent >> 16;
      const int4 tab = s_tab[k * d.L + l];
      const float2 xy = *reinterpret_cast<const float2 *>(s_loc + 2 * s);
      TapGeom gm;
      tap_geometry(xy.x, xy.y, tab.x, tab.y, tab.z, gm);
      const int row = q == 0 ? gm.row[0] : q == 1 ? gm.row[1] : q == 2 ? gm.row[2] : gm.row[3];
      const float bw = ((q & 2) ? gm.lh : gm.hh) * ((q & 1) ? gm.lw : gm.hw);
      const bool use = live && row >= 0;
      const unsigned key = use ? static_cast<unsigned>(row) : (0x80000000u | lane);
      const unsigned grp = __match_any_sync(0xffffffffu, key);
      const bool leader = use && (static_cast<int>(__ffs(grp)) - 1 == lane);
      if (!wready) {
        mbar_wait(&bars[1], 0);  // weights have landed
        wready = true;
      }
      // coefficients in the permuted order the gather reads them: [gq][v]  (g = v * GPV + gq)
      float cf[G];
      {
        const float4 *wp = reinterpret_cast<const float4 *>(s_w + (sparse_w ? ii : s) * wstride + l * G);
        float wv[G];
#pragma unroll
        for (int x = 0; x < G / 4; ++x) {
          const float4 t = wp[x];
          wv[4 * x] = t.x, wv[4 * x + 1] = t.y, wv[4 * x + 2] = t.z, wv[4 * x + 3] = t.w;
        }
#pragma unroll
        for (int g = 0; g < G; ++g) cf[(g % GPV) * VPL + g / GPV] = use ? bw * wv[g] : 0.f;
      }
      // leaders add the coefficients of the other lanes of their group, lowest lane first
      unsigned rem = leader ? (grp & ~(1u << lane)) : 0u;
      const int n_it = __reduce_max_sync(0xffffffffu, __popc(rem));
      for (int it = 0; it < n_it; ++it) {
        const int src = rem ? (static_cast<int>(__ffs(rem)) - 1) : lane;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const float v = __shfl_sync(0xffffffffu, cf[g], src);
          if (rem) cf[g] += v;
        }
        rem &= rem - 1;
      }
      // has an earlier chunk of this round listed the row already?
      const uint32_t h = (static_cast<uint32_t>(row) * 0x9E3779B1u) >> 25;  // MERGE_TABLE = 128
      const uint32_t e = leader ? s_table[h] : 0xffffffffu;
      const bool hit = leader && (e >> 6) == static_cast<uint32_t>(row);
      const bool fresh = leader && !hit;
      const unsigned fresh_m = __ballot_sync(0xffffffffu, fresh);
      const int slot = hit ? static_cast<int>(e & 63u) : cnt + __popc(fresh_m & lt_mask);
      float4 *cp = reinterpret_cast<float4 *>(my_coef + slot * G);
      if (hit) {
#pragma unroll
        for (int x = 0; x < G / 4; ++x) {
          float4 t = cp[x];
          t.x += cf[4 * x], t.y += cf[4 * x + 1], t.z += cf[4 * x + 2], t.w += cf[4 * x + 3];
          cp[x] = t;
        }
      } else if (fresh) {
        s_table[h] = (static_cast<uint32_t>(row) << 6) | static_cast<uint32_t>(slot);
        my_rowoff[slot] = static_cast<uint32_t>(row) * rb;
        if (prefetch == 1)  // start the row's DRAM -> L2 transfer now; the gather then runs at L2 latency
          tma_prefetch_l2(fb - lane * 16 + static_cast<uint32_t>(row) * rb, rb);
        else if (prefetch == 2)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(fb - lane * 16 + static_cast<uint32_t>(row) * rb));
#pragma unroll
        for (int x = 0; x < G / 4; ++x)
          cp[x] = make_float4(cf[4 * x], cf[4 * x + 1], cf[4 * x + 2], cf[4 * x + 3]);
      }
      cnt += __popc(fresh_m);
      __syncwarp();
    }
    if (lane == 0) s_cnt[warp] = cnt;
    DFA_STAMP(4);
    __syncthreads();
    DFA_STAMP(5);
    // ---- this warp's share: slot p of list c goes to warp (p + c) mod NW --------------------------
    int n_mine = 0;
#pragma unroll
    for (int c = 0; c < NW; ++c) {
      const int cn = s_cnt[c];
      const int first = (warp - c) & (NW - 1);
      const int mine = cn > first ? (cn - first + NW - 1) / NW : 0;
      if (lane < mine) {
        const int slot = c * MERGE_CAP + first + lane * NW;
        s_mine_off[n_mine + lane] = s_rowoff[slot];
        s_mine_slot[n_mine + lane] = static_cast<uint16_t>(slot);
      }
      n_mine += mine;
    }
    if (lane < U) {  // padding up to a multiple of U: no load, zero coefficients
      s_mine_off[n_mine + lane] = 0xffffffffu;
      s_mine_slot[n_mine + lane] = 0;
    }
    __syncwarp();
    // ---- gather: every distinct row once ----------------------------------------------------------
    for (int k0 = 0; k0 < n_mine; k0 += U) {
      typename FeatVec<T>::raw_t val[U][VPL];
      uint32_t off[U];
      int slot[U];
      if constexpr (U == 4) {
        const uint4 o4 = *reinterpret_cast<const uint4 *>(s_mine_off + k0);
        const uint2 s2 = *reinterpret_cast<const uint2 *>(s_mine_slot + k0);
        off[0] = o4.x, off[1] = o4.y, off[2] = o4.z, off[3] = o4.w;
        slot[0] = s2.x & 0xffff, slot[1] = s2.x >> 16, slot[2] = s2.y & 0xffff, slot[3] = s2.y >> 16;
      } else {
        const uint2 o2 = *reinterpret_cast<const uint2 *>(s_mine_off + k0);
        const uint32_t s1 = *reinterpret_cast<const uint32_t *>(s_mine_slot + k0);
        off[0] = o2.x, off[1] = o2.y;
        slot[0] = s1 & 0xffff, slot[1] = s1 >> 16;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool ok = off[u] != 0xffffffffu;
#pragma unroll
        for (int v = 0; v < VPL; ++v)
          val[u][v] = ok ? FeatVec<T>::load_raw(reinterpret_cast<const T *>(fb + (off[u] + 512u * v)))
                         : FeatVec<T>::zero_raw();
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool ok = off[u] != 0xffffffffu;
        const float *cp = s_coef + slot[u] * G + gq * VPL;
        float cv[VPL];
        if (VPL == 2) {
          const float2 t = *reinterpret_cast<const float2 *>(cp);
          cv[0] = ok ? t.x : 0.f, cv[VPL - 1] = ok ? t.y : 0.f;
        } else {
          cv[0] = ok ? *cp : 0.f;
        }
#pragma unroll
        for (int v = 0; v < VPL; ++v) FeatVec<T>::fma(acc[v], cv[v], val[u][v]);
      }
    }
    DFA_STAMP(6);
    __syncthreads();  // the lists are free again (next round, or the partial rows)
  }
  if (TMA && whole_weights && !wready) mbar_wait(&bars[1], 0);  // never exit with a copy in flight

  // ---- epilogue: fold the warps' partial rows ---------------------------------------------------
  {
    float *part = s_coef + warp * C;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      float4 *o = reinterpret_cast<float4 *>(part + (v * 32 + lane) * VEC);
#pragma unroll
      for (int c = 0; c < VEC / 4; ++c)
        o[c] = make_float4(acc[v][4 * c], acc[v][4 * c + 1], acc[v][4 * c + 2], acc[v][4 * c + 3]);
    }
  }
  __syncthreads();
  for (int c = tid; c < C; c += NT) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) sum += s_coef[w * C + c];
    out[static_cast<size_t>(anchor) * C + c] = sum;
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
// backward
// ------------------------------------------------------------------------------------------
template <typename T, int LPG, int U, bool TMA, int MAXT>
__global__ void __launch_bounds__(MAXT, (MAXT <= 256) ? (1024 / MAXT) : 1)
    dfa_bwd_kernel(const T *__restrict__ feat, const int *__restrict__ shape,
                   const int *__restrict__ start, const float *__restrict__ loc,
                   const float *__restrict__ weights, const float *__restrict__ grad_out,
                   float *__restrict__ grad_feat, float *__restrict__ grad_loc,
                   float *__restrict__ grad_w, Dims d, int overwrite) {
  constexpr int VEC = FeatVec<T>::VEC;
  constexpr int TPW = 32 / (4 * LPG);
  extern __shared__ __align__(128) unsigned char smem[];
  const SmemLayout lay = smem_layout(d.P, d.K, d.L, d.G, TPW * U, true);
  float *s_w = reinterpret_cast<float *>(smem + lay.w);
  float *s_loc = reinterpret_cast<float *>(smem + lay.loc);
  TapB *s_rec = reinterpret_cast<TapB *>(smem + lay.rec);
  int *s_widx = reinterpret_cast<int *>(smem + lay.widx);
  int *s_list = reinterpret_cast<int *>(smem + lay.list);
  float2 *s_gl = reinterpret_cast<float2 *>(smem + lay.gl);  // [G][P*K] per-warp rows
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + lay.bar);
  int *s_nvalid = reinterpret_cast<int *>(bars + 2);

  const int tid = threadIdx.x, lane = tid & 31, g = tid >> 5;
  const int anchor = blockIdx.x, b = anchor / d.A;
  const int PK = d.P * d.K, wcount = PK * d.L * d.G;
  float *gw_a = grad_w + static_cast<size_t>(anchor) * wcount;
  float *gl_a = grad_loc + static_cast<size_t>(anchor) * PK * 2;

  const int nv = stage_and_compact<TMA>(loc + static_cast<size_t>(anchor) * PK * 2,
                                        weights + static_cast<size_t>(anchor) * wcount, s_w,
                                        s_loc, s_list, bars, s_nvalid, PK, wcount);
  const int ntaps = nv * d.L;
  const int ntaps_pad = (ntaps + TPW * U - 1) / (TPW * U) * (TPW * U);

  if (overwrite) {  // masked samples get explicit zeros: the caller needs no memset
    for (int i = tid; i < wcount; i += blockDim.x) gw_a[i] = 0.f;
    for (int i = tid; i < 2 * PK; i += blockDim.x) gl_a[i] = 0.f;
  }
  for (int i = tid; i < PK * d.G; i += blockDim.x) s_gl[i] = make_float2(0.f, 0.f);
  for (int t = tid; t < ntaps_pad; t += blockDim.x) {
    TapB r;
    r.off[0] = r.off[1] = r.off[2] = r.off[3] = -1;
    r.lh = r.lw = r.Wf = r.Hf = 0.f;
    int widx = 0;
    if (t < ntaps) {
      const int l = t / nv, i = t - l * nv, s = s_list[i], k = s % d.K, kl = k * d.L + l;
      const int H = __ldg(shape + 2 * kl), W = __ldg(shape + 2 * kl + 1);
      TapGeom gm;
      tap_geometry(s_loc[2 * s], s_loc[2 * s + 1], H, W, __ldg(start + kl), gm);
#pragma unroll
      for (int q = 0; q < 4; ++q) r.off[q] = gm.row[q] >= 0 ? gm.row[q] * d.C : -1;
      r.lh = gm.lh, r.lw = gm.lw, r.Wf = static_cast<float>(W), r.Hf = static_cast<float>(H);
      widx = (s * d.L + l) * d.G;
    }
    s_rec[t] = r;
    s_widx[t] = widx;
  }
  __syncthreads();
  if (TMA) mbar_wait(&bars[1], 0);

  const int j = lane % LPG, q = (lane / LPG) & 3, sub = lane / (4 * LPG);
  const int cpg = d.C / d.G;
  const int choff = g * cpg + j * VEC;
  const size_t fbase = static_cast<size_t>(b) * d.num_feat * d.C + choff;
  const T *fb = feat + fbase;
  float *gfb = grad_feat + fbase;
  float go[VEC];
  {
    const float4 *p = reinterpret_cast<const float4 *>(grad_out + static_cast<size_t>(anchor) * d.C + choff);
#pragma unroll
    for (int c = 0; c < VEC / 4; ++c) {
      const float4 t = __ldg(p + c);
      go[4 * c] = t.x, go[4 * c + 1] = t.y, go[4 * c + 2] = t.z, go[4 * c + 3] = t.w;
    }
  }
  const bool qh = (q & 2) != 0, qw = (q & 1) != 0;  // corner uses h_high / w_high

  for (int t0 = 0; t0 < ntaps_pad; t0 += TPW * U) {
    float dsum[U], bwq[U], cxq[U], cyq[U], wgt[U];
    int widx[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int t = t0 + u * TPW + sub;
      const int off = s_rec[t].off[q];
      const float4 fr = *reinterpret_cast<const float4 *>(&s_rec[t].lh);  // lh, lw, W, H
      widx[u] = s_widx[t] + g;
      wgt[u] = s_w[widx[u]];
      const float ah = qh ? fr.x : 1.f - fr.x;  // lh or hh
      const float aw = qw ? fr.y : 1.f - fr.y;  // lw or hw
      bwq[u] = ah * aw;
      cxq[u] = (qw ? ah : -ah) * fr.z;  // d val / d x, already times W
      cyq[u] = (qh ? aw : -aw) * fr.w;  // d val / d y, already times H
      float v[VEC];
      float dd = 0.f;
      if (off >= 0) {
        FeatVec<T>::load(fb + off, v);
        const float coef = bwq[u] * wgt[u];
#pragma unroll
        for (int c = 0; c < VEC; ++c) dd = fmaf(go[c], v[c], dd);
        if (grad_feat) {
#pragma unroll
          for (int c = 0; c < VEC; c += 4)
            red_add_v4(gfb + off + c, coef * go[c], coef * go[c + 1], coef * go[c + 2],
                       coef * go[c + 3]);
        }
      }
      dsum[u] = dd;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float dd = dsum[u];
#pragma unroll
      for (int m = 1; m < LPG; m <<= 1) dd += __shfl_xor_sync(0xffffffffu, dd, m);
      float pa = bwq[u] * dd, px = cxq[u] * dd, py = cyq[u] * dd;
#pragma unroll
      for (int m = LPG; m < 4 * LPG; m <<= 1) {
        pa += __shfl_xor_sync(0xffffffffu, pa, m);
        px += __shfl_xor_sync(0xffffffffu, px, m);
        py += __shfl_xor_sync(0xffffffffu, py, m);
      }
      const int t = t0 + u * TPW + sub;
      if (j == 0 && q == 0 && t < ntaps) {
        const int wi = widx[u];
        if (overwrite) gw_a[wi] = pa; else gw_a[wi] += pa;
        const int i = t % nv;  // level-major tap order: sample index inside the valid list
        float2 *cell = &s_gl[g * PK + i];
        if (TPW == 1) {  // one lane per warp owns the row: plain, ordered accumulation
          float2 c = *cell;
          c.x = fmaf(px, wgt[u], c.x), c.y = fmaf(py, wgt[u], c.y);
          *cell = c;
        } else {
          atomicAdd(&cell->x, px * wgt[u]);
          atomicAdd(&cell->y, py * wgt[u]);
        }
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < nv; i += blockDim.x) {
    float gx = 0.f, gy = 0.f;
    for (int w = 0; w < d.G; ++w) gx += s_gl[w * PK + i].x, gy += s_gl[w * PK + i].y;
    const int s = s_list[i];
    if (overwrite) {
      gl_a[2 * s] = gx, gl_a[2 * s + 1] = gy;
    } else {
      gl_a[2 * s] += gx, gl_a[2 * s + 1] += gy;
    }
  }
}

// ------------------------------------------------------------------------------------------
// backward with row merging (the default for SimPB's shapes)
// ------------------------------------------------------------------------------------------
// The three gradients need each feature row only through two quantities:
//     grad_feat[row, c]  +=  coef[row, g(c)] * grad_out[c]          coef = sum of (bilinear x weight)
//     D[row, g]           =  sum_{c in g} grad_out[c] * feat[row, c]
// and grad_weights / grad_sampling_location of a tap are small combinations of the D of its four
// corner rows.  So the backward is the forward's pipeline — same prologue, same merge of duplicate
// rows, same balanced gather of every DISTINCT row once — with a dot product in place of the axpy,
// one vector reduction (red.global.add.v4.f32) per 16 bytes of a distinct row, and a short epilogue
// per tap.  Against the per-group kernel above this halves the dependent rounds of loads per warp,
// and merging removes a third or more of the atomics.  grad_weights and grad_sampling_location are
// produced without atomics, in a fixed order (bitwise reproducible).
struct MergeBwdLayout {
  MergeLayout m;
  uint32_t dot, refslot, item, gl, total;
};
__host__ __device__ inline MergeBwdLayout merge_bwd_layout(int P, int K, int L, int G, int NW, int U) {
  MergeBwdLayout s;
  s.m = merge_layout(P, K, L, G, NW, U);
  uint32_t o = s.m.total;
  s.dot = o, o = align_up(o + 4u * G * NW * MERGE_CAP, 128);
  s.refslot = o, o = align_up(o + 8u * NW * 16u, 16);  // 4 x uint16 per tap, 16 taps per warp and round
  s.item = o, o = align_up(o + 4u * NW * 2u, 16);
  s.gl = o, o = align_up(o + 8u * P * K * L, 16);
  s.total = o;
  return s;
}

template <typename T, int VPL, int G, int NW, int U, bool TMA, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB)
    dfa_bwd_merge_kernel(const T *__restrict__ feat, const int *__restrict__ shape,
                         const int *__restrict__ start, const float *__restrict__ loc,
                         const float *__restrict__ weights, const float *__restrict__ grad_out,
                         float *__restrict__ grad_feat, float *__restrict__ grad_loc,
                         float *__restrict__ grad_w, Dims d, MergeBwdLayout blay, int whole_weights,
                         int overwrite) {
  constexpr int VEC = FeatVec<T>::VEC;
  constexpr int NT = NW * 32;
  constexpr int GPV = G / VPL;
  constexpr int C = 32 * VPL * VEC;
  constexpr int LPQ = 32 * VPL / G;  // lanes that share one group inside a 512-byte segment
  constexpr int RB_SHIFT = VPL == 2 ? 10 : 9;
  static_assert(G % 4 == 0 && G % VPL == 0 && (32 * VPL) % G == 0 && G <= 32, "group geometry");
  static_assert((NW & (NW - 1)) == 0 && NW >= 2 && NW <= 32, "warps per CTA: a power of two");
  static_assert(U == 2 || U == 4, "row offsets / slots of a batch are read with one vector load");
  static_assert(NT % (8 * G) == 0 || (8 * G) % NT == 0, "epilogue item mapping");
  const MergeLayout &lay = blay.m;
  extern __shared__ __align__(128) unsigned char smem[];
  float *s_w = reinterpret_cast<float *>(smem + lay.w);
  float *s_loc = reinterpret_cast<float *>(smem + lay.loc);
  uint32_t *s_list = reinterpret_cast<uint32_t *>(smem + lay.list);
  int4 *s_tab = reinterpret_cast<int4 *>(smem + lay.tab);
  uint32_t *s_rowoff = reinterpret_cast<uint32_t *>(smem + lay.rowoff);
  float *s_coef = reinterpret_cast<float *>(smem + lay.coef);
  float *s_dot = reinterpret_cast<float *>(smem + blay.dot);
  uint2 *s_refslot = reinterpret_cast<uint2 *>(smem + blay.refslot);
  int *s_item = reinterpret_cast<int *>(smem + blay.item);
  float2 *s_gl = reinterpret_cast<float2 *>(smem + blay.gl);
  int *s_cnt = reinterpret_cast<int *>(smem + lay.cnt);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + lay.bar);
  int *s_nvalid = reinterpret_cast<int *>(bars + 2);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int anchor = blockIdx.x, b = anchor / d.A;
  const int PK = d.P * d.K, LG = d.L * G, wcount = PK * LG;
  const float *loc_g = loc + static_cast<size_t>(anchor) * PK * 2;
  const float *w_g = weights + static_cast<size_t>(anchor) * wcount;
  float *gw_a = grad_w + static_cast<size_t>(anchor) * wcount;
  float *gl_a = grad_loc + static_cast<size_t>(anchor) * PK * 2;
  const unsigned lt_mask = (1u << lane) - 1u;

  // ---- prologue (as in the forward) ------------------------------------------------------------
  if (TMA) {
    if (warp == 0) {
      if (lane == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
        mbar_expect_tx(&bars[0], 8u * PK);
        tma_bulk_g2s(s_loc, loc_g, 8u * PK, &bars[0]);
        if (whole_weights) {
          mbar_expect_tx(&bars[1], 4u * wcount);
          tma_bulk_g2s(s_w, w_g, 4u * wcount, &bars[1]);
        }
      }
      __syncwarp();
      mbar_wait(&bars[0], 0);
    } else {
      for (int i = tid - 32; i < d.K * d.L; i += NT - 32)
        s_tab[i] = make_int4(__ldg(shape + 2 * i), __ldg(shape + 2 * i + 1), __ldg(start + i), 0);
    }
  } else {
    for (int i = tid; i < 2 * PK; i += NT) s_loc[i] = __ldg(loc_g + i);
    for (int i = tid; i < d.K * d.L; i += NT)
      s_tab[i] = make_int4(__ldg(shape + 2 * i), __ldg(shape + 2 * i + 1), __ldg(start + i), 0);
    __syncthreads();
  }
  // this lane's slice of grad_out stays in registers for the whole kernel
  float go[VPL][VEC];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const float4 *p = reinterpret_cast<const float4 *>(grad_out + static_cast<size_t>(anchor) * C +
                                                       (v * 32 + lane) * VEC);
#pragma unroll
    for (int c = 0; c < VEC / 4; ++c) {
      const float4 t = __ldg(p + c);
      go[v][4 * c] = t.x, go[v][4 * c + 1] = t.y, go[v][4 * c + 2] = t.z, go[v][4 * c + 3] = t.w;
    }
  }
  if (overwrite) {  // masked samples get explicit zeros: the caller needs no memset for these two
    for (int i = tid; i < wcount; i += NT) gw_a[i] = 0.f;
    for (int i = tid; i < 2 * PK; i += NT) gl_a[i] = 0.f;
  }
  bool sparse_w = false;
  if (warp == 0) {
    const float rK = 1.0f / static_cast<float>(d.K);
    int n = 0;
    for (int base = 0; base < PK; base += 32) {
      const int s = base + lane;
      bool v = false;
      if (s < PK) {
        const float2 xy = *reinterpret_cast<const float2 *>(s_loc + 2 * s);
        v = sample_valid(xy.x, xy.y);
      }
      const unsigned m = __ballot_sync(0xffffffffu, v);
      if (v) {
        const int p = static_cast<int>((static_cast<float>(s) + 0.5f) * rK);
        s_list[n + __popc(m & lt_mask)] = static_cast<uint32_t>(s) | (static_cast<uint32_t>(s - p * d.K) << 16);
      }
      n += __popc(m);
    }
    sparse_w = TMA && !whole_weights && 2 * n <= PK;
    if (lane == 0) *s_nvalid = sparse_w ? -n - 1 : n;
    if (TMA && !whole_weights && n > 0) {
      if (!sparse_w) {
        if (lane == 0) {
          mbar_expect_tx(&bars[1], 4u * wcount);
          tma_bulk_g2s(s_w, w_g, 4u * wcount, &bars[1]);
        }
      } else {
        if (lane == 0) mbar_expect_tx(&bars[1], 4u * LG * n);
        __syncwarp();
        for (int i = lane; i < n; i += 32) {
          const int s = s_list[i] & 0xffff;
          tma_bulk_g2s(s_w + i * (LG + MERGE_WPAD), w_g + s * LG, 4u * LG, &bars[1]);
        }
      }
    }
  }
  __syncthreads();
  int nv = *s_nvalid;
  sparse_w = nv < 0;
  nv = sparse_w ? -nv - 1 : nv;
  if (!TMA) {
    for (int i = tid; i < nv * LG; i += NT) {
      const int s = s_list[i / LG] & 0xffff, r = i - (i / LG) * LG;
      s_w[s * LG + r] = __ldg(w_g + s * LG + r);
    }
    __syncthreads();
  }
  const int wstride = sparse_w ? LG + MERGE_WPAD : LG;

  const uint32_t rb = 512u * VPL;
  const unsigned char *fb = reinterpret_cast<const unsigned char *>(feat) +
                            static_cast<size_t>(b) * d.num_feat * rb + lane * 16;
  float *gfb = grad_feat + static_cast<size_t>(b) * d.num_feat * C + lane * VEC;
  const int gq = (lane * G) / (32 * VPL);

  const int cpl = (nv + 7) >> 3;
  const int my_levels = warp < d.L ? (d.L - warp + NW - 1) / NW : 0;
  const int max_items = ((d.L + NW - 1) / NW) * cpl;
  uint32_t *s_table = reinterpret_cast<uint32_t *>(smem + lay.table) + warp * MERGE_TABLE;
  uint32_t *my_rowoff = s_rowoff + warp * MERGE_CAP;
  float *my_coef = s_coef + warp * MERGE_CAP * G;
  uint32_t *s_mine_off = reinterpret_cast<uint32_t *>(smem + lay.mine_off) + warp * lay.mine_stride;
  uint16_t *s_mine_slot = reinterpret_cast<uint16_t *>(smem + lay.mine_slot) + warp * lay.mine_stride;
  bool wready = !TMA;
  int li = 0, cj = 0;

  for (int it0 = 0; it0 < max_items; it0 += 2) {
    // ---- merge: this warp's next two chunks; remember every reference's slot -------------------------
    int cnt = 0;
    if (li < my_levels) {
#pragma unroll
      for (int x = 0; x < MERGE_TABLE / 128; ++x)
        reinterpret_cast<uint4 *>(s_table)[x * 32 + lane] =
            make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
      __syncwarp();
    }
    for (int x = 0; x < 2; ++x) {
      if (li >= my_levels) {
        if (lane == 0) s_item[warp * 2 + x] = -1;
        continue;
      }
      const int l = warp + li * NW;
      if (lane == 0) s_item[warp * 2 + x] = (l << 16) | cj;
      const int i = (cj << 3) + (lane >> 2), q = lane & 3;
      if (++cj == cpl) cj = 0, ++li;
      const bool live = i < nv;
      const int ii = live ? i : 0;
      const uint32_t ent = s_list[ii];
      const int s = ent & 0xffff, k = ent >> 16;
      const int4 tab = s_tab[k * d.L + l];
      const float2 xy = *reinterpret_cast<const float2 *>(s_loc + 2 * s);
      TapGeom gm;
      tap_geometry(xy.x, xy.y, tab.x, tab.y, tab.z, gm);
      const int row = q == 0 ? gm.row[0] : q == 1 ? gm.row[1] : q == 2 ? gm.row[2] : gm.row[3];
      const float bw = ((q & 2) ? gm.lh : gm.hh) * ((q & 1) ? gm.lw : gm.hw);
      const bool use = live && row >= 0;
      const unsigned key = use ? static_cast<unsigned>(row) : (0x80000000u | lane);
      const unsigned grp = __match_any_sync(0xffffffffu, key);
      const int lead_lane = static_cast<int>(__ffs(grp)) - 1;
      const bool leader = use && lead_lane == lane;
      if (!wready) {
        mbar_wait(&bars[1], 0);
        wready = true;
      }
      float cf[G];
      {
        const float4 *wp = reinterpret_cast<const float4 *>(s_w + (sparse_w ? ii : s) * wstride + l * G);
        float wv[G];
#pragma unroll
        for (int y = 0; y < G / 4; ++y) {
          const float4 t = wp[y];
          wv[4 * y] = t.x, wv[4 * y + 1] = t.y, wv[4 * y + 2] = t.z, wv[4 * y + 3] = t.w;
        }
#pragma unroll
        for (int g = 0; g < G; ++g) cf[(g % GPV) * VPL + g / GPV] = use ? bw * wv[g] : 0.f;
      }
      unsigned rem = leader ? (grp & ~(1u << lane)) : 0u;
      const int n_it = __reduce_max_sync(0xffffffffu, __popc(rem));
      for (int it = 0; it < n_it; ++it) {
        const int src = rem ? (static_cast<int>(__ffs(rem)) - 1) : lane;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const float v = __shfl_sync(0xffffffffu, cf[g], src);
          if (rem) cf[g] += v;
        }
        rem &= rem - 1;
      }
      const uint32_t h = (static_cast<uint32_t>(row) * 0x9E3779B1u) >> 25;
      const uint32_t e = leader ? s_table[h] : 0xffffffffu;
      const bool hit = leader && (e >> 6) == static_cast<uint32_t>(row);
      const bool fresh = leader && !hit;
      const unsigned fresh_m = __ballot_sync(0xffffffffu, fresh);
      const int slot = hit ? static_cast<int>(e & 63u) : cnt + __popc(fresh_m & lt_mask);
      float4 *cp = reinterpret_cast<float4 *>(my_coef + slot * G);
      if (hit) {
#pragma unroll
        for (int y = 0; y < G / 4; ++y) {
          float4 t = cp[y];
          t.x += cf[4 * y], t.y += cf[4 * y + 1], t.z += cf[4 * y + 2], t.w += cf[4 * y + 3];
          cp[y] = t;
        }
      } else if (fresh) {
        s_table[h] = (static_cast<uint32_t>(row) << 6) | static_cast<uint32_t>(slot);
        my_rowoff[slot] = static_cast<uint32_t>(row) * rb;
#pragma unroll
        for (int y = 0; y < G / 4; ++y)
          cp[y] = make_float4(cf[4 * y], cf[4 * y + 1], cf[4 * y + 2], cf[4 * y + 3]);
      }
      cnt += __popc(fresh_m);
      // every reference learns the slot of its row from its group's leader
      const int gslot = __shfl_sync(0xffffffffu, warp * MERGE_CAP + slot, use ? lead_lane : lane);
      const uint32_t mys = use ? static_cast<uint32_t>(gslot) : 0xffffu;
      const uint32_t up = __shfl_down_sync(0xffffffffu, mys, 1);
      const uint32_t pair = mys | (up << 16);                        // corners (q, q+1)
      const uint32_t pair2 = __shfl_down_sync(0xffffffffu, pair, 2);  // corners (2, 3) at q == 0
      if (q == 0) s_refslot[(warp * 2 + x) * 8 + (lane >> 2)] = make_uint2(pair, pair2);
      __syncwarp();
    }
    if (lane == 0) s_cnt[warp] = cnt;
    __syncthreads();
    // ---- this warp's share of the distinct rows -------------------------------------------------------
    int n_mine = 0;
#pragma unroll
    for (int c = 0; c < NW; ++c) {
      const int cn = s_cnt[c];
      const int first = (warp - c) & (NW - 1);
      const int mine = cn > first ? (cn - first + NW - 1) / NW : 0;
      if (lane < mine) {
        const int slot = c * MERGE_CAP + first + lane * NW;
        s_mine_off[n_mine + lane] = s_rowoff[slot];
        s_mine_slot[n_mine + lane] = static_cast<uint16_t>(slot);
      }
      n_mine += mine;
    }
    if (lane < U) {
      s_mine_off[n_mine + lane] = 0xffffffffu;
      s_mine_slot[n_mine + lane] = 0;
    }
    __syncwarp();
    // ---- gather: scatter coef x grad_out, and D[row][g] = <grad_out, row> over the group ------------
    for (int k0 = 0; k0 < n_mine; k0 += U) {
      typename FeatVec<T>::raw_t val[U][VPL];
      uint32_t off[U];
      int slot[U];
      if constexpr (U == 4) {
        const uint4 o4 = *reinterpret_cast<const uint4 *>(s_mine_off + k0);
        const uint2 s2 = *reinterpret_cast<const uint2 *>(s_mine_slot + k0);
        off[0] = o4.x, off[1] = o4.y, off[2] = o4.z, off[3] = o4.w;
        slot[0] = s2.x & 0xffff, slot[1] = s2.x >> 16, slot[2] = s2.y & 0xffff, slot[3] = s2.y >> 16;
      } else {
        const uint2 o2 = *reinterpret_cast<const uint2 *>(s_mine_off + k0);
        const uint32_t s1 = *reinterpret_cast<const uint32_t *>(s_mine_slot + k0);
        off[0] = o2.x, off[1] = o2.y;
        slot[0] = s1 & 0xffff, slot[1] = s1 >> 16;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool ok = off[u] != 0xffffffffu;
#pragma unroll
        for (int v = 0; v < VPL; ++v)
          val[u][v] = ok ? FeatVec<T>::load_raw(reinterpret_cast<const T *>(fb + (off[u] + 512u * v)))
                         : FeatVec<T>::zero_raw();
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool ok = off[u] != 0xffffffffu;
        if (ok && grad_feat) {  // the scatter does not wait for the loads
          const float *cp = s_coef + slot[u] * G + gq * VPL;
          float *gr = gfb + static_cast<size_t>(off[u] >> RB_SHIFT) * C;
#pragma unroll
          for (int v = 0; v < VPL; ++v) {
            const float cv = cp[v];
#pragma unroll
            for (int c = 0; c < VEC; c += 4)
              red_add_v4(gr + v * 32 * VEC + c, cv * go[v][c], cv * go[v][c + 1], cv * go[v][c + 2],
                         cv * go[v][c + 3]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float dd[VPL];
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          float x[VEC];
          FeatVec<T>::unpack(val[u][v], x);
          float t = 0.f;
#pragma unroll
          for (int c = 0; c < VEC; ++c) t = fmaf(go[v][c], x[c], t);
          dd[v] = t;
        }
#pragma unroll
        for (int m = 1; m < LPQ; m <<= 1)
#pragma unroll
          for (int v = 0; v < VPL; ++v) dd[v] += __shfl_xor_sync(0xffffffffu, dd[v], m);
        if ((lane & (LPQ - 1)) == 0 && off[u] != 0xffffffffu) {
          float *dp = s_dot + slot[u] * G + gq * VPL;
#pragma unroll
          for (int v = 0; v < VPL; ++v) dp[v] = dd[v];
        }
      }
    }
    __syncthreads();
    // ---- per tap: grad_weights, and the tap's share of grad_sampling_location -------------------------
    for (int item = tid; item < NW * 2 * 8 * G; item += NT) {
      const int g = item % G, j = (item / G) & 7, wx = item / (8 * G);
      const int it = s_item[wx];
      const int l = it >> 16, i = ((it & 0xffff) << 3) + j;
      const bool live = it >= 0 && i < nv;
      float px = 0.f, py = 0.f;
      if (live) {
        const uint32_t ent = s_list[i];
        const int s = ent & 0xffff, k = ent >> 16;
        const int4 tab = s_tab[k * d.L + l];
        const float2 xy = *reinterpret_cast<const float2 *>(s_loc + 2 * s);
        TapGeom gm;
        tap_geometry(xy.x, xy.y, tab.x, tab.y, tab.z, gm);
        const uint2 rs = s_refslot[wx * 8 + j];
        const uint32_t sl[4] = {rs.x & 0xffffu, rs.x >> 16, rs.y & 0xffffu, rs.y >> 16};
        const int pg = (g % GPV) * VPL + g / GPV;
        float D[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) D[c] = sl[c] != 0xffffu ? s_dot[sl[c] * G + pg] : 0.f;
        const float pa = gm.hh * gm.hw * D[0] + gm.hh * gm.lw * D[1] + gm.lh * gm.hw * D[2] +
                         gm.lh * gm.lw * D[3];
        const float wgt = s_w[(sparse_w ? i : s) * wstride + l * G + g];
        px = wgt * static_cast<float>(tab.y) * (gm.hh * (D[1] - D[0]) + gm.lh * (D[3] - D[2]));
        py = wgt * static_cast<float>(tab.x) * (gm.hw * (D[2] - D[0]) + gm.lw * (D[3] - D[1]));
        float *o = gw_a + (s * d.L + l) * G + g;
        if (overwrite) *o = pa; else *o += pa;
      }
#pragma unroll
      for (int m = 1; m < G; m <<= 1) {
        px += __shfl_xor_sync(0xffffffffu, px, m);
        py += __shfl_xor_sync(0xffffffffu, py, m);
      }
      if (live && g == 0) s_gl[i * d.L + l] = make_float2(px, py);
    }
    __syncthreads();  // lists, D and slots are free for the next round
  }
  if (TMA && whole_weights && !wready) mbar_wait(&bars[1], 0);
  // ---- grad_sampling_location: levels summed in a fixed order -----------------------------------------
  for (int i = tid; i < nv; i += NT) {
    float gx = 0.f, gy = 0.f;
    for (int l = 0; l < d.L; ++l) gx += s_gl[i * d.L + l].x, gy += s_gl[i * d.L + l].y;
    const int s = s_list[i] & 0xffff;
    if (overwrite) {
      gl_a[2 * s] = gx, gl_a[2 * s + 1] = gy;
    } else {
      gl_a[2 * s] += gx, gl_a[2 * s + 1] += gy;
    }
  }
}

// Shape-generic backward: thread per channel, one tap at a time, block-level reductions.
template <typename T>
__global__ void __launch_bounds__(256)
    dfa_bwd_generic_kernel(const T *__restrict__ feat, const int *__restrict__ shape,
                           const int *__restrict__ start, const float *__restrict__ loc,
                           const float *__restrict__ weights, const float *__restrict__ grad_out,
                           float *__restrict__ grad_feat, float *__restrict__ grad_loc,
                           float *__restrict__ grad_w, Dims d, int overwrite) {
  const int anchor = blockIdx.x, b = anchor / d.A, tid = threadIdx.x;
  const int PK = d.P * d.K, cpg = d.C / d.G;
  const float *loc_a = loc + static_cast<size_t>(anchor) * PK * 2;
  const float *w_a = weights + static_cast<size_t>(anchor) * PK * d.L * d.G;
  float *gw_a = grad_w + static_cast<size_t>(anchor) * PK * d.L * d.G;
  float *gl_a = grad_loc + static_cast<size_t>(anchor) * PK * 2;
  const T *fb = feat + static_cast<size_t>(b) * d.num_feat * d.C;
  float *gfb = grad_feat + static_cast<size_t>(b) * d.num_feat * d.C;
  const float *go = grad_out + static_cast<size_t>(anchor) * d.C;
  __shared__ float s_red[3][256];
  for (int s = 0; s < PK; ++s) {
    const float x = __ldg(loc_a + 2 * s), y = __ldg(loc_a + 2 * s + 1);
    const bool ok = sample_valid(x, y);  // uniform across the block
    float glx = 0.f, gly = 0.f;
    for (int l = 0; l < d.L; ++l) {
      const int k = s % d.K, kl = k * d.L + l;
      float *gw_t = gw_a + (s * d.L + l) * d.G;
      if (!ok) {
        if (overwrite)
          for (int w = tid; w < d.G; w += blockDim.x) gw_t[w] = 0.f;
        continue;
      }
      const int H = __ldg(shape + 2 * kl), W = __ldg(shape + 2 * kl + 1);
      TapGeom gm;
      tap_geometry(x, y, H, W, __ldg(start + kl), gm);
      const float bw[4] = {gm.hh * gm.hw, gm.hh * gm.lw, gm.lh * gm.hw, gm.lh * gm.lw};
      const float cx[4] = {-gm.hh, gm.hh, -gm.lh, gm.lh}, cy[4] = {-gm.hw, -gm.lw, gm.hw, gm.lw};
      for (int w = 0; w < d.G; ++w) {  // one group at a time keeps the reduction simple
        float pa = 0.f, px = 0.f, py = 0.f;
        const float wgt = __ldg(w_a + (s * d.L + l) * d.G + w);
        for (int c = w * cpg + tid; c < (w + 1) * cpg; c += blockDim.x) {
          const float gr = __ldg(go + c);
          for (int q = 0; q < 4; ++q) {
            if (gm.row[q] < 0) continue;
            const size_t fi = static_cast<size_t>(gm.row[q]) * d.C + c;
            const float v = static_cast<float>(fb[fi]);
            pa = fmaf(bw[q] * gr, v, pa);
            px = fmaf(cx[q] * gr, v, px);
            py = fmaf(cy[q] * gr, v, py);
            if (grad_feat) atomicAdd(gfb + fi, bw[q] * wgt * gr);
          }
        }
        s_red[0][tid] = pa, s_red[1][tid] = px * wgt * W, s_red[2][tid] = py * wgt * H;
        __syncthreads();
        for (int m = blockDim.x / 2; m > 0; m >>= 1) {
          if (tid < m)
            for (int r = 0; r < 3; ++r) s_red[r][tid] += s_red[r][tid + m];
          __syncthreads();
        }
        if (tid == 0) {
          if (overwrite) gw_t[w] = s_red[0][0]; else gw_t[w] += s_red[0][0];
        }
        glx += s_red[1][0], gly += s_red[2][0];
        __syncthreads();
      }
    }
    if (tid == 0) {
      if (overwrite) {
        gl_a[2 * s] = glx, gl_a[2 * s + 1] = gly;
      } else if (ok) {
        gl_a[2 * s] += glx, gl_a[2 * s + 1] += gly;
      }
    }
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
// feature-map flattening: NCHW levels → [bs, K*sum(HW), C] channel-last, one pass
// ------------------------------------------------------------------------------------------
// A 32(pixels) x 32(channels) tile goes through shared memory so both the NCHW read (pixels
// contiguous) and the channel-last write (channels contiguous) are coalesced.
template <typename TO>
__global__ void __launch_bounds__(256)
    dfa_flatten_level_kernel(const float *__restrict__ src, TO *__restrict__ dst, int HW, int C,
                             int K, long long dst_rows_per_batch, int rows_per_cam, int level_row0) {
  __shared__ float tile[32][33];
  const int bk = blockIdx.z;  // b * K + k
  const int b = bk / K, k = bk - b * K;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const float *s = src + static_cast<size_t>(bk) * C * HW;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i, p = p0 + tx;
    tile[ty + 8 * i][tx] = (c < C && p < HW) ? __ldg(s + static_cast<size_t>(c) * HW + p) : 0.f;
  }
  __syncthreads();
  TO *o = dst + (static_cast<size_t>(b) * dst_rows_per_batch +
                 static_cast<size_t>(k) * rows_per_cam + level_row0) * C;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p = p0 + ty + 8 * i, c = c0 + tx;
    if (p < HW && c < C) o[static_cast<size_t>(p) * C + c] = static_cast<TO>(tile[tx][ty + 8 * i]);
  }
}

// ------------------------------------------------------------------------------------------
// key points + camera projection
// ------------------------------------------------------------------------------------------
// One thread per (b, a, p): builds the 3-D key point, then projects it into the K cameras.
// Follows models/detection3d/blocks.py:181-207 and models/blocks.py:198-213; the 4-term dot
// products are evaluated left to right with fused multiply-adds.
__global__ void __launch_bounds__(256)
    dfa_keypoints_project_kernel(const float *__restrict__ anchor, const float *__restrict__ fix_scale,
                                 int num_fix, const float *__restrict__ logits,
                                 const float *__restrict__ proj, const float *__restrict__ wh,
                                 float *__restrict__ kp_out, float *__restrict__ loc_out, int bs,
                                 int A, int P, int K) {
  const long long n = static_cast<long long>(bs) * A * P;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const int p = static_cast<int>(i % P);
  const long long ba = i / P;
  const int b = static_cast<int>(ba / A);
  const float *an = anchor + ba * 11;
  const float sx = expf(an[3]), sy = expf(an[4]), sz = expf(an[5]);  // W, L, H
  float ox, oy, oz;
  if (p < num_fix) {
    ox = fix_scale[3 * p], oy = fix_scale[3 * p + 1], oz = fix_scale[3 * p + 2];
  } else {
    const float *lg = logits + ba * (P - num_fix) * 3 + (p - num_fix) * 3;
    ox = 1.f / (1.f + expf(-lg[0])) - 0.5f;
    oy = 1.f / (1.f + expf(-lg[1])) - 0.5f;
    oz = 1.f / (1.f + expf(-lg[2])) - 0.5f;
  }
  ox *= sx, oy *= sy, oz *= sz;
  const float sn = an[6], cs = an[7];
  const float x = fmaf(cs, ox, -sn * oy) + an[0];
  const float y = fmaf(sn, ox, cs * oy) + an[1];
  const float z = oz + an[2];
  if (kp_out) kp_out[3 * i] = x, kp_out[3 * i + 1] = y, kp_out[3 * i + 2] = z;
  for (int k = 0; k < K; ++k) {
    const float *m = proj + (static_cast<size_t>(b) * K + k) * 16;
    const float u = fmaf(m[2], z, fmaf(m[1], y, m[0] * x)) + m[3];
    const float v = fmaf(m[6], z, fmaf(m[5], y, m[4] * x)) + m[7];
    const float dpt = fmaf(m[10], z, fmaf(m[9], y, m[8] * x)) + m[11];
    const float den = fmaxf(dpt, 1e-5f);
    float px = u / den, py = v / den;
    if (wh) px /= wh[(b * K + k) * 2], py /= wh[(b * K + k) * 2 + 1];
    float *o = loc_out + (static_cast<size_t>(i) * K + k) * 2;
    o[0] = px, o[1] = py;
  }
}

// Backward of the kernel above: one thread per (b, a) walks the anchor's P key points and K
// cameras in a fixed order, so the gradients need no atomics and are bitwise reproducible.
// grad_anchor [bs,A,11] (velocity entries get 0), grad_logits [bs,A,(P-F)*3] (may be NULL).
__global__ void __launch_bounds__(128)
    dfa_keypoints_project_bwd_kernel(const float *__restrict__ anchor, const float *__restrict__ fix_scale,
                                     int num_fix, const float *__restrict__ logits,
                                     const float *__restrict__ proj, const float *__restrict__ wh,
                                     const float *__restrict__ grad_loc, float *__restrict__ grad_anchor,
                                     float *__restrict__ grad_logits, int bs, int A, int P, int K) {
  const long long ba = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (ba >= static_cast<long long>(bs) * A) return;
  const int b = static_cast<int>(ba / A);
  const float *an = anchor + ba * 11;
  const float size[3] = {expf(an[3]), expf(an[4]), expf(an[5])};
  const float sn = an[6], cs = an[7];
  float g_ctr[3] = {0.f, 0.f, 0.f}, g_size[3] = {0.f, 0.f, 0.f}, g_sn = 0.f, g_cs = 0.f;
  for (int p = 0; p < P; ++p) {
    float off[3], dsig[3] = {0.f, 0.f, 0.f};
    if (p < num_fix) {
      off[0] = fix_scale[3 * p], off[1] = fix_scale[3 * p + 1], off[2] = fix_scale[3 * p + 2];
    } else {
      const float *lg = logits + ba * (P - num_fix) * 3 + (p - num_fix) * 3;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float sg = 1.f / (1.f + expf(-lg[i]));
        off[i] = sg - 0.5f, dsig[i] = sg * (1.f - sg);
      }
    }
    const float ox = off[0] * size[0], oy = off[1] * size[1], oz = off[2] * size[2];
    const float x = fmaf(cs, ox, -sn * oy) + an[0];
    const float y = fmaf(sn, ox, cs * oy) + an[1];
    const float z = oz + an[2];
    float gx = 0.f, gy = 0.f, gz = 0.f;
    for (int k = 0; k < K; ++k) {
      const float *m = proj + (static_cast<size_t>(b) * K + k) * 16;
      const float *gl = grad_loc + ((ba * P + p) * K + k) * 2;
      const float u = fmaf(m[2], z, fmaf(m[1], y, m[0] * x)) + m[3];
      const float v = fmaf(m[6], z, fmaf(m[5], y, m[4] * x)) + m[7];
      const float dpt = fmaf(m[10], z, fmaf(m[9], y, m[8] * x)) + m[11];
      const float den = fmaxf(dpt, 1e-5f);
      float gpx = gl[0], gpy = gl[1];
      if (wh) gpx /= wh[(b * K + k) * 2], gpy /= wh[(b * K + k) * 2 + 1];
      const float gu = gpx / den, gv = gpy / den;
      // d/d den of (u/den, v/den); the clamp passes the gradient where dpt >= 1e-5 (torch.clamp)
      const float gd = dpt >= 1e-5f ? -(gu * u + gv * v) / den : 0.f;
      gx += m[0] * gu + m[4] * gv + m[8] * gd;
      gy += m[1] * gu + m[5] * gv + m[9] * gd;
      gz += m[2] * gu + m[6] * gv + m[10] * gd;
    }
    g_ctr[0] += gx, g_ctr[1] += gy, g_ctr[2] += gz;
    const float go[3] = {cs * gx + sn * gy, -sn * gx + cs * gy, gz};  // wrt the rotated-back offset
    g_cs += ox * gx + oy * gy;
    g_sn += -oy * gx + ox * gy;
#pragma unroll
    for (int i = 0; i < 3; ++i) g_size[i] += off[i] * go[i];
    if (p >= num_fix && grad_logits) {
      float *o = grad_logits + ba * (P - num_fix) * 3 + (p - num_fix) * 3;
#pragma unroll
      for (int i = 0; i < 3; ++i) o[i] = go[i] * size[i] * dsig[i];
    }
  }
  float *ga = grad_anchor + ba * 11;
  ga[0] = g_ctr[0], ga[1] = g_ctr[1], ga[2] = g_ctr[2];
  ga[3] = g_size[0] * size[0], ga[4] = g_size[1] * size[1], ga[5] = g_size[2] * size[2];  // d exp
  ga[6] = g_sn, ga[7] = g_cs, ga[8] = 0.f, ga[9] = 0.f, ga[10] = 0.f;
}

// ------------------------------------------------------------------------------------------
// attention weights: softmax over (K, L, P) + attn-drop mask + permute, one pass
// ------------------------------------------------------------------------------------------
// logits [bs, A, K, L, P, G] (= weights_fc output, models/blocks.py:175-186) -> weights
// [bs, A, P, K, L, G] (the op's layout, models/blocks.py:133-144), softmax taken over the N =
// K*L*P entries of each (b, a, g).  keep [bs, A, K, P] (uint8, may be NULL) is the attn-drop keep
// mask of models/blocks.py:188-195 and `scale` its 1/(1-p).  One CTA per anchor; the anchor's
// logits are staged in shared memory once.  Thread t owns group t % G (G divides the block).
// Reduction over the threads that own the same group (tid % G, G a power of two <= 32 here): xor
// shuffles inside the warp, then one shared-memory row per warp.
template <int NT>
__device__ __forceinline__ float group_reduce(float v, float *s_red, int tid, int G, bool is_max) {
  if (G <= 32 && (G & (G - 1)) == 0) {
    for (int m = G; m < 32; m <<= 1) {
      const float o = __shfl_xor_sync(0xffffffffu, v, m);
      v = is_max ? fmaxf(v, o) : v + o;
    }
    const int lane = tid & 31, warp = tid >> 5;
    if (lane < G) s_red[warp * G + lane] = v;
    __syncthreads();
    float r = s_red[lane % G];
    for (int w = 1; w < NT / 32; ++w) r = is_max ? fmaxf(r, s_red[w * G + lane % G]) : r + s_red[w * G + lane % G];
    __syncthreads();
    return r;
  }
  s_red[tid] = v;
  __syncthreads();
  float r = s_red[tid % G];
  for (int j = (tid % G) + G; j < NT; j += G) r = is_max ? fmaxf(r, s_red[j]) : r + s_red[j];
  __syncthreads();
  return r;
}

// `logits_cam` (may be NULL) is the camera part of split logits: weights_fc is linear, so
// weights_fc(feature[b,a] + camera_embed[b,k]) = weights_fc(feature[b,a]) + W * camera_embed[b,k];
// the module then runs the GEMM on [bs*A] and [bs*K] rows instead of [bs*A*K] and this kernel adds
// the two parts on the fly: logits[b,a,k,r,g] = logits[b,a,r,g] + logits_cam[b,k,r,g], r = (l,p).
//
// Thread t owns group t % G and the rows r = t / G, t / G + NT / G, ... of every camera, so the
// loops need no integer division; the (k,l,p) -> (p,k,l) permutation and the keep mask come from a
// small table built once per CTA.
struct SoftmaxTables {
  uint16_t *perm;  // row (k,l,p) -> output row (p*K + k)*L + l
  uint8_t *keep;   // row (k,l,p) -> keep flag
};

template <int NT>
__device__ __forceinline__ void softmax_tables(SoftmaxTables t, const uint8_t *kp, int tid, int K, int L,
                                               int P) {
  const int LP = L * P, N = K * LP;
  for (int n = tid; n < N; n += NT) {
    const int k = n / LP, r = n - k * LP, l = r / P, p = r - l * P;
    t.perm[n] = static_cast<uint16_t>((p * K + k) * L + l);
    t.keep[n] = kp ? kp[k * P + p] : 1;
  }
}

// s_x[e] <- softmax numerators; returns 1 / sum for this thread's group
template <int NT>
__device__ __forceinline__ float softmax_stage(float *s_x, float *s_red, const float *la, const float *lk,
                                               int tid, int K, int LP, int G) {
  const int g = tid % G, r0 = tid / G, rs = NT / G;
  float mx = -INFINITY;
  for (int k = 0; k < K; ++k)
    for (int r = r0; r < LP; r += rs) {
      const int e = (k * LP + r) * G + g;
      const float v = lk ? __ldg(la + r * G + g) + __ldg(lk + e) : __ldg(la + e);
      s_x[e] = v;
      mx = fmaxf(mx, v);
    }
  mx = group_reduce<NT>(mx, s_red, tid, G, true);
  float sum = 0.f;
  for (int k = 0; k < K; ++k)
    for (int r = r0; r < LP; r += rs) {
      const int e = (k * LP + r) * G + g;
      const float v = expf(s_x[e] - mx);
      s_x[e] = v;
      sum += v;
    }
  sum = group_reduce<NT>(sum, s_red, tid, G, false);
  return 1.f / sum;
}

template <int NT>
__global__ void __launch_bounds__(NT)
    dfa_softmax_weights_kernel(const float *__restrict__ logits, const float *__restrict__ logits_cam,
                               const uint8_t *__restrict__ keep, float scale, float *__restrict__ w,
                               int A, int K, int L, int P, int G) {
  extern __shared__ __align__(16) float s_x[];  // N*G logits, then the tables
  __shared__ float s_red[NT];
  const int tid = threadIdx.x, LP = L * P, N = K * LP, n_el = N * G;
  SoftmaxTables tb{reinterpret_cast<uint16_t *>(s_x + n_el), nullptr};
  tb.keep = reinterpret_cast<uint8_t *>(tb.perm + N);
  const size_t base = static_cast<size_t>(blockIdx.x) * n_el;
  const float *la = logits_cam ? logits + static_cast<size_t>(blockIdx.x) * LP * G : logits + base;
  const float *lk = logits_cam ? logits_cam + static_cast<size_t>(blockIdx.x / A) * n_el : nullptr;
  softmax_tables<NT>(tb, keep ? keep + static_cast<size_t>(blockIdx.x) * K * P : nullptr, tid, K, L, P);
  const float inv = softmax_stage<NT>(s_x, s_red, la, lk, tid, K, LP, G);  // syncs inside: tables visible
  const int g = tid % G, r0 = tid / G, rs = NT / G;
  const float on = keep ? scale : 1.f;
  for (int k = 0; k < K; ++k)
    for (int r = r0; r < LP; r += rs) {
      const int n = k * LP + r;
      const float v = tb.keep[n] ? s_x[n * G + g] * inv * on : 0.f;
      w[base + static_cast<size_t>(tb.perm[n]) * G + g] = v;
    }
}

// grad_logits = y * (dy - sum_n dy_n y_n) with y = softmax(logits) recomputed and
// dy = keep * scale * grad_w (read through the permutation).  With split logits the anchor part of
// the gradient (sum over cameras) is also written: grad_anchor [bs,A,L*P*G]; the camera part is the
// sum of grad_logits over anchors, left to the caller.
template <int NT>
__global__ void __launch_bounds__(NT)
    dfa_softmax_weights_bwd_kernel(const float *__restrict__ logits, const float *__restrict__ logits_cam,
                                   const uint8_t *__restrict__ keep, float scale,
                                   const float *__restrict__ grad_w, float *__restrict__ grad_logits,
                                   float *__restrict__ grad_anchor, int A, int K, int L, int P, int G) {
  extern __shared__ __align__(16) float s_x[];  // N*G softmax values, N*G dy, then the tables
  __shared__ float s_red[NT];
  const int tid = threadIdx.x, LP = L * P, N = K * LP, n_el = N * G;
  float *s_dy = s_x + n_el;
  SoftmaxTables tb{reinterpret_cast<uint16_t *>(s_dy + n_el), nullptr};
  tb.keep = reinterpret_cast<uint8_t *>(tb.perm + N);
  const size_t base = static_cast<size_t>(blockIdx.x) * n_el;
  const float *la = logits_cam ? logits + static_cast<size_t>(blockIdx.x) * LP * G : logits + base;
  const float *lk = logits_cam ? logits_cam + static_cast<size_t>(blockIdx.x / A) * n_el : nullptr;
  softmax_tables<NT>(tb, keep ? keep + static_cast<size_t>(blockIdx.x) * K * P : nullptr, tid, K, L, P);
  const float inv = softmax_stage<NT>(s_x, s_red, la, lk, tid, K, LP, G);
  const int g = tid % G, r0 = tid / G, rs = NT / G;
  const float on = keep ? scale : 1.f;
  float dot = 0.f;
  for (int k = 0; k < K; ++k)
    for (int r = r0; r < LP; r += rs) {
      const int n = k * LP + r, e = n * G + g;
      const float dy = tb.keep[n] ? __ldg(grad_w + base + static_cast<size_t>(tb.perm[n]) * G + g) * on : 0.f;
      const float y = s_x[e] * inv;
      s_x[e] = y, s_dy[e] = dy;
      dot = fmaf(dy, y, dot);
    }
  dot = group_reduce<NT>(dot, s_red, tid, G, false);
  for (int r = r0; r < LP; r += rs) {
    float t = 0.f;
    for (int k = 0; k < K; ++k) {
      const int e = (k * LP + r) * G + g;
      const float gx = s_x[e] * (s_dy[e] - dot);
      grad_logits[base + e] = gx;
      t += gx;
    }
    if (grad_anchor) grad_anchor[static_cast<size_t>(blockIdx.x) * LP * G + r * G + g] = t;
  }
}

// Vectorised forms for G % 4 == 0 (SimPB: G = 8): a thread owns FOUR consecutive groups of a row
// (one 16-byte load / store), threads t and t + Q share their groups (Q = G / 4 quads per row).
// Same staging and tables as above; a quarter of the instructions.
template <int NT>
__device__ __forceinline__ float4 quad_reduce(float4 v, float4 *s_red4, int tid, int Q, bool is_max) {
  // Q is a power of two <= 8 here: lanes with equal tid % Q combine by xor shuffles, warps via smem
  auto comb = [is_max](float a, float b) { return is_max ? fmaxf(a, b) : a + b; };
  for (int m = Q; m < 32; m <<= 1) {
    v.x = comb(v.x, __shfl_xor_sync(0xffffffffu, v.x, m));
    v.y = comb(v.y, __shfl_xor_sync(0xffffffffu, v.y, m));
    v.z = comb(v.z, __shfl_xor_sync(0xffffffffu, v.z, m));
    v.w = comb(v.w, __shfl_xor_sync(0xffffffffu, v.w, m));
  }
  const int lane = tid & 31, warp = tid >> 5;
  if (lane < Q) s_red4[warp * Q + lane] = v;
  __syncthreads();
  float4 r = s_red4[lane % Q];
  for (int w = 1; w < NT / 32; ++w) {
    const float4 o = s_red4[w * Q + lane % Q];
    r.x = comb(r.x, o.x), r.y = comb(r.y, o.y), r.z = comb(r.z, o.z), r.w = comb(r.w, o.w);
  }
  __syncthreads();
  return r;
}

// numerators into s_x4, returns 1/sum for the thread's four groups
template <int NT>
__device__ __forceinline__ float4 softmax_stage4(float4 *s_x4, float4 *s_red4, const float4 *la,
                                                 const float4 *lk, int tid, int K, int LP, int Q) {
  const int q = tid % Q, r0 = tid / Q, rs = NT / Q;
  float4 mx = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  for (int k = 0; k < K; ++k)
    for (int r = r0; r < LP; r += rs) {
      const int e = (k * LP + r) * Q + q;
      float4 v;
      if (lk) {
        const float4 a = __ldg(la + r * Q + q), c = __ldg(lk + e);
        v = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
      } else {
        v = __ldg(la + e);
      }
      s_x4[e] = v;
      mx = make_float4(fmaxf(mx.x, v.x), fmaxf(mx.y, v.y), fmaxf(mx.z, v.z), fmaxf(mx.w, v.w));
    }
  mx = quad_reduce<NT>(mx, s_red4, tid, Q, true);
  float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k = 0; k < K; ++k)
    for (int r = r0; r < LP; r += rs) {
      const int e = (k * LP + r) * Q + q;
      float4 v = s_x4[e];
      v = make_float4(expf(v.x - mx.x), expf(v.y - mx.y), expf(v.z - mx.z), expf(v.w - mx.w));
      s_x4[e] = v;
      sum.x += v.x, sum.y += v.y, sum.z += v.z, sum.w += v.w;
    }
  sum = quad_reduce<NT>(sum, s_red4, tid, Q, false);
  return make_float4(1.f / sum.x, 1.f / sum.y, 1.f / sum.z, 1.f / sum.w);
}

template <int NT>
__global__ void __launch_bounds__(NT)
    dfa_softmax_weights4_kernel(const float *__restrict__ logits, const float *__restrict__ logits_cam,
                                const uint8_t *__restrict__ keep, float scale, float *__restrict__ w,
                                int A, int K, int L, int P, int G) {
  extern __shared__ __align__(16) float s_x[];
  __shared__ float4 s_red4[NT / 32 * 8];
  const int tid = threadIdx.x, LP = L * P, N = K * LP, n_el = N * G, Q = G / 4;
  float4 *s_x4 = reinterpret_cast<float4 *>(s_x);
  SoftmaxTables tb{reinterpret_cast<uint16_t *>(s_x + n_el), nullptr};
  tb.keep = reinterpret_cast<uint8_t *>(tb.perm + N);
  const size_t base = static_cast<size_t>(blockIdx.x) * n_el;
  const float *la = logits_cam ? logits + static_cast<size_t>(blockIdx.x) * LP * G : logits + base;
  const float *lk = logits_cam ? logits_cam + static_cast<size_t>(blockIdx.x / A) * n_el : nullptr;
  softmax_tables<NT>(tb, keep ? keep + static_cast<size_t>(blockIdx.x) * K * P : nullptr, tid, K, L, P);
  const float4 inv = softmax_stage4<NT>(s_x4, s_red4, reinterpret_cast<const float4 *>(la),
                                        reinterpret_cast<const float4 *>(lk), tid, K, LP, Q);
  const int q = tid % Q, r0 = tid / Q, rs = NT / Q;
  const float on = keep ? scale : 1.f;
  float4 *w4 = reinterpret_cast<float4 *>(w + base);
  for (int k = 0; k < K; ++k)
    for (int r = r0; r < LP; r += rs) {
      const int n = k * LP + r;
      const float4 v = s_x4[n * Q + q];
      const float m = tb.keep[n] ? on : 0.f;
      w4[static_cast<size_t>(tb.perm[n]) * Q + q] =
          make_float4(v.x * inv.x * m, v.y * inv.y * m, v.z * inv.z * m, v.w * inv.w * m);
    }
}

template <int NT>
__global__ void __launch_bounds__(NT)
    dfa_softmax_weights4_bwd_kernel(const float *__restrict__ logits, const float *__restrict__ logits_cam,
                                    const uint8_t *__restrict__ keep, float scale,
                                    const float *__restrict__ grad_w, float *__restrict__ grad_logits,
                                    float *__restrict__ grad_anchor, int A, int K, int L, int P, int G) {
  extern __shared__ __align__(16) float s_x[];
  __shared__ float4 s_red4[NT / 32 * 8];
  const int tid = threadIdx.x, LP = L * P, N = K * LP, n_el = N * G, Q = G / 4;
  float4 *s_x4 = reinterpret_cast<float4 *>(s_x), *s_dy4 = s_x4 + N * Q;
  SoftmaxTables tb{reinterpret_cast<uint16_t *>(s_x + 2 * n_el), nullptr};
  tb.keep = reinterpret_cast<uint8_t *>(tb.perm + N);
  const size_t base = static_cast<size_t>(blockIdx.x) * n_el;
  const float *la = logits_cam ? logits + static_cast<size_t>(blockIdx.x) * LP * G : logits + base;
  const float *lk = logits_cam ? logits_cam + static_cast<size_t>(blockIdx.x / A) * n_el : nullptr;
  softmax_tables<NT>(tb, keep ? keep + static_cast<size_t>(blockIdx.x) * K * P : nullptr, tid, K, L, P);
  const float4 inv = softmax_stage4<NT>(s_x4, s_red4, reinterpret_cast<const float4 *>(la),
                                        reinterpret_cast<const float4 *>(lk), tid, K, LP, Q);
  const int q = tid % Q, r0 = tid / Q, rs = NT / Q;
  const float on = keep ? scale : 1.f;
  const float4 *gw4 = reinterpret_cast<const float4 *>(grad_w + base);
  float4 dot = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k = 0; k < K; ++k)
    for (int r = r0; r < LP; r += rs) {
      const int n = k * LP + r, e = n * Q + q;
      float4 dy = __ldg(gw4 + static_cast<size_t>(tb.perm[n]) * Q + q);
      const float m = tb.keep[n] ? on : 0.f;
      dy = make_float4(dy.x * m, dy.y * m, dy.z * m, dy.w * m);
      float4 y = s_x4[e];
      y = make_float4(y.x * inv.x, y.y * inv.y, y.z * inv.z, y.w * inv.w);
      s_x4[e] = y, s_dy4[e] = dy;
      dot.x = fmaf(dy.x, y.x, dot.x), dot.y = fmaf(dy.y, y.y, dot.y);
      dot.z = fmaf(dy.z, y.z, dot.z), dot.w = fmaf(dy.w, y.w, dot.w);
    }
  dot = quad_reduce<NT>(dot, s_red4, tid, Q, false);
  float4 *gl4 = reinterpret_cast<float4 *>(grad_logits + base);
  for (int r = r0; r < LP; r += rs) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < K; ++k) {
      const int e = (k * LP + r) * Q + q;
      const float4 y = s_x4[e], dy = s_dy4[e];
      const float4 gx = make_float4(y.x * (dy.x - dot.x), y.y * (dy.y - dot.y), y.z * (dy.z - dot.z),
                                    y.w * (dy.w - dot.w));
      gl4[e] = gx;
      t.x += gx.x, t.y += gx.y, t.z += gx.z, t.w += gx.w;
    }
    if (grad_anchor)
      reinterpret_cast<float4 *>(grad_anchor + static_cast<size_t>(blockIdx.x) * LP * G)[r * Q + q] = t;
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
inline bool aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

int check_dims(const dfa_dims *dd, Dims &d) {
  if (!dd) return DFA_ERR_NULL_POINTER;
  d = Dims{dd->batch_size, dd->num_cams, dd->num_feat, dd->num_embeds,
           dd->num_scale,  dd->num_anchors, dd->num_pts, dd->num_groups};
  if (d.bs <= 0 || d.K <= 0 || d.num_feat <= 0 || d.C <= 0 || d.L <= 0 || d.A <= 0 || d.P <= 0 ||
      d.G <= 0)
    return DFA_ERR_BAD_DIMS;
  if (d.C % d.G != 0) return DFA_ERR_BAD_DIMS;
  // 32-bit element offsets inside one batch item; 31-bit anchor index
  if (static_cast<long long>(d.num_feat) * d.C >= (1ll << 31)) return DFA_ERR_BAD_DIMS;
  if (static_cast<long long>(d.bs) * d.A >= (1ll << 31)) return DFA_ERR_BAD_DIMS;
  if (static_cast<long long>(d.P) * d.K * d.L * d.G >= (1ll << 24)) return DFA_ERR_BAD_DIMS;
  return 0;
}

template <typename K>
int set_smem(K kernel, uint32_t bytes) {
  if (bytes > 227u * 1024u) return DFA_ERR_UNSUPPORTED;
  if (bytes > 48u * 1024u) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(bytes));
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  return 0;
}

constexpr int FWD_U = 4;
constexpr int BWD_U = 2;

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

template <typename T, int LPG, bool TMA, int MAXT>
int launch_bwd_t(const void *feat, const int *shape, const int *start, const float *loc,
                 const float *w, const float *go, float *gf, float *gl, float *gw, const Dims &d,
                 int overwrite, cudaStream_t st) {
  auto kern = dfa_bwd_kernel<T, LPG, BWD_U, TMA, MAXT>;
  constexpr int TPW = 32 / (4 * LPG);
  const SmemLayout lay = smem_layout(d.P, d.K, d.L, d.G, TPW * BWD_U, true);
  if (int rc = set_smem(kern, lay.total)) return rc;
  kern<<<d.bs * d.A, 32 * d.G, lay.total, st>>>(static_cast<const T *>(feat), shape, start, loc, w,
                                               go, gf, gl, gw, d, overwrite);
  return static_cast<int>(cudaGetLastError());
}

// Fast path applies when a group's channels are a power-of-two number (1..8) of 16-byte vectors
// and the block (one warp per group) fits; everything else takes the generic kernels.
template <typename T>
int fast_lpg(const Dims &d, const void *feat) {
  const int bytes = (d.C / d.G) * static_cast<int>(sizeof(T));
  if (bytes % 16 != 0 || !aligned(feat, 16) || d.G > 32) return 0;
  if ((d.C * static_cast<int>(sizeof(T))) % 16 != 0) return 0;
  const int lpg = bytes / 16;
  return (lpg == 1 || lpg == 2 || lpg == 4 || lpg == 8) ? lpg : 0;
}

inline bool tma_ok(const Dims &d, const float *loc, const float *w) {
  const long long wbytes = 4ll * d.P * d.K * d.L * d.G, lbytes = 8ll * d.P * d.K;
  return wbytes % 16 == 0 && lbytes % 16 == 0 && aligned(loc, 16) && aligned(w, 16);
}

#define DFA_DISPATCH_LPG(CALL)                    \
  switch (lpg) {                                  \
    case 8: return CALL(8);                       \
    case 4: return CALL(4);                       \
    case 2: return CALL(2);                       \
    default: return CALL(1);                      \
  }

template <typename T, int U, bool TMA, int NT, int MINB, bool PF>
int launch_fwd_rows(const void *feat, const int *shape, const int *start, const float *loc,
                    const float *w, float *out, const Dims &d, int vpr, cudaStream_t st) {
  auto kern = dfa_fwd_rows_kernel<T, U, TMA, NT, MINB, PF>;
  const int slices = NT / vpr;
  int vpr_log2 = 0;
  while ((1 << vpr_log2) < vpr) ++vpr_log2;
  const SmemLayout2 lay = smem_layout2(d.P, d.K, d.L, d.G, d.C, slices, slices * U);
  if (int rc = set_smem(kern, lay.total)) return rc;
  kern<<<d.bs * d.A, NT, lay.total, st>>>(static_cast<const T *>(feat), shape, start, loc, w, out, d,
                                          vpr_log2);
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
  return vpr;
}

inline int env_int(const char *name, int dflt) {
  const char *e = getenv(name);
  return e ? atoi(e) : dflt;
}

template <typename T, int VPL, int NW, int U, bool TMA, int MINB>
int launch_fwd_merge(const void *feat, const int *shape, const int *start, const float *loc,
                     const float *w, float *out, const Dims &d, cudaStream_t st) {
  auto kern = dfa_fwd_merge_kernel<T, VPL, 8, NW, U, TMA, MINB>;
  const MergeLayout lay = merge_layout(d.P, d.K, d.L, d.G, NW, U);
  if (int rc = set_smem(kern, lay.total)) return rc;
  // A grid that fits the machine in about one wave is bound by latency, not bandwidth: fetch the
  // whole weights block at once instead of waiting for the sample mask first.
  const long long grid = static_cast<long long>(d.bs) * d.A;
  const int whole = env_int("DFA_FWD_WHOLE_WEIGHTS", grid <= 148 * 8 ? 1 : 0);
  const int prefetch = env_int("DFA_FWD_PREFETCH", 0);
  kern<<<d.bs * d.A, NW * 32, lay.total, st>>>(static_cast<const T *>(feat), shape, start, loc, w,
                                              out, d, lay, whole, prefetch);
  return static_cast<int>(cudaGetLastError());
}

// The merging kernel applies when a feature row is 512 or 1024 bytes (a lane owns one or two
// 16-byte vectors of it) and there are 8 channel groups — SimPB's C=256 / G=8 in fp32 and bf16.
// Returns vectors per lane, 0 when the shape does not fit.
template <typename T>
int merge_vpl(const Dims &d, const void *feat) {
  const long long rb = static_cast<long long>(d.C) * static_cast<long long>(sizeof(T));
  if (d.G != 8 || !aligned(feat, 16)) return 0;
  if (rb != 512 && rb != 1024) return 0;
  if (d.K > 64 || static_cast<long long>(d.P) * d.K >= 65536) return 0;  // packed sample list
  if (static_cast<long long>(d.num_feat) * rb >= (1ll << 32)) return 0;
  const MergeLayout lay = merge_layout(d.P, d.K, d.L, d.G, 8, 8);
  if (lay.total > 200u * 1024u) return 0;
  return static_cast<int>(rb / 512);
}

// TMA staging needs 16-byte sized and aligned blocks, and byte counts an mbarrier can track.
inline bool warp_tma_ok(const Dims &d, const float *loc, const float *w) {
  const long long line = 4ll * d.L * d.G, wbytes = line * d.P * d.K, lbytes = 8ll * d.P * d.K;
  return line % 16 == 0 && lbytes % 16 == 0 && wbytes < (1ll << 20) && aligned(loc, 16) &&
         aligned(w, 16);
}

template <typename T>
int forward_typed(const void *feat, const int *shape, const int *start, const float *loc,
                  const float *w, float *out, const Dims &d, cudaStream_t st) {
  // DFA_FWD_VARIANT (tuning knob): 1..4 = row-sliced kernel with (threads, taps in flight) =
  // (256,1) (256,2) (512,1) (512,2) — 1 is the default, the fastest measured on B200 at SimPB's
  // shapes; 10.. = row-merging kernel (fewer DRAM bytes and L1 wavefronts, longer dependent chain
  // per warp: within 5-25 % of the default); 0 = one-warp-per-group kernel.
  // A variant whose shape constraints are not met falls through to the next family.
  const int variant = env_int("DFA_FWD_VARIANT", 1);
  if (variant >= 10) {  // merging kernel: (warps, rows in flight, CTAs per SM) per variant
    const int vpl = merge_vpl<T>(d, feat);
    if (vpl) {
      const bool tma = warp_tma_ok(d, loc, w);
#define WARPK(VPL, NW, U, MINB)                                                                  \
  (tma ? launch_fwd_merge<T, VPL, NW, U, true, MINB>(feat, shape, start, loc, w, out, d, st)       \
       : launch_fwd_merge<T, VPL, NW, U, false, MINB>(feat, shape, start, loc, w, out, d, st))
#define WARPV(NW, U, MINB) (vpl == 2 ? WARPK(2, NW, U, MINB) : WARPK(1, NW, U, MINB))
      switch (variant) {
        case 11: return WARPV(4, 4, 8);
        case 12: return WARPV(4, 4, 9);
        case 13: return WARPV(4, 2, 12);
        case 14: return WARPV(8, 4, 5);
        case 15: return WARPV(8, 4, 4);
        case 16: return WARPV(8, 2, 6);
        case 17: return WARPV(2, 4, 18);
        default: return WARPV(4, 4, 10);
      }
#undef WARPV
#undef WARPK
    }
  }
  const int rvariant = variant >= 10 ? 1 : variant;
  if (rvariant >= 1) {
    const int nt = rvariant == 5 ? 192 : rvariant == 6 ? 128 : rvariant >= 3 ? 512 : 256;
    const int vpr = rows_vpr<T>(d, feat, nt);
    if (vpr) {
      const bool tma = tma_ok(d, loc, w);
      const bool pf = env_int("DFA_FWD_PREFETCH", 0) != 0;
#define ROWS(U, NT, MINB)                                                                              \
  (tma ? (pf ? launch_fwd_rows<T, U, true, NT, MINB, true>(feat, shape, start, loc, w, out, d, vpr, st)  \
             : launch_fwd_rows<T, U, true, NT, MINB, false>(feat, shape, start, loc, w, out, d, vpr, st)) \
       : launch_fwd_rows<T, U, false, NT, MINB, false>(feat, shape, start, loc, w, out, d, vpr, st))
      switch (rvariant) {
        case 1: return ROWS(1, 256, 6);
        case 2: return ROWS(2, 256, 4);
        case 3: return ROWS(1, 512, 3);
        case 5: return ROWS(1, 192, 8);
        case 6: return ROWS(1, 128, 12);
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

template <typename T, int VPL, int NW, int U, bool TMA, int MINB>
int launch_bwd_merge(const void *feat, const int *shape, const int *start, const float *loc,
                     const float *w, const float *go, float *gf, float *gl, float *gw, const Dims &d,
                     int overwrite, cudaStream_t st) {
  auto kern = dfa_bwd_merge_kernel<T, VPL, 8, NW, U, TMA, MINB>;
  const MergeBwdLayout lay = merge_bwd_layout(d.P, d.K, d.L, d.G, NW, U);
  if (int rc = set_smem(kern, lay.total)) return rc;
  const long long grid = static_cast<long long>(d.bs) * d.A;
  const int whole = env_int("DFA_BWD_WHOLE_WEIGHTS", grid <= 148 * 8 ? 1 : 0);
  kern<<<d.bs * d.A, NW * 32, lay.total, st>>>(static_cast<const T *>(feat), shape, start, loc, w, go,
                                              gf, gl, gw, d, lay, whole, overwrite);
  return static_cast<int>(cudaGetLastError());
}

template <typename T>
int backward_typed(const void *feat, const int *shape, const int *start, const float *loc,
                   const float *w, const float *go, float *gf, float *gl, float *gw, const Dims &d,
                   int overwrite, cudaStream_t st) {
  // DFA_BWD_VARIANT (tuning knob): 10.. = row-merging kernel (default where the shape fits),
  // 0 = one-warp-per-group kernel.
  const int variant = env_int("DFA_BWD_VARIANT", 10);
  if (variant >= 10 && aligned(go, 16) && aligned(gf, 16)) {
    const int vpl = merge_vpl<T>(d, feat);
    const MergeBwdLayout bl = merge_bwd_layout(d.P, d.K, d.L, d.G, 8, 4);
    if (vpl && bl.total <= 200u * 1024u) {
      const bool tma = warp_tma_ok(d, loc, w);
#define BWDK(VPL, NW, U, MINB)                                                                        \
  (tma ? launch_bwd_merge<T, VPL, NW, U, true, MINB>(feat, shape, start, loc, w, go, gf, gl, gw, d,    \
                                                     overwrite, st)                                   \
       : launch_bwd_merge<T, VPL, NW, U, false, MINB>(feat, shape, start, loc, w, go, gf, gl, gw, d,   \
                                                      overwrite, st))
#define BWDV(NW, U, MINB) (vpl == 2 ? BWDK(2, NW, U, MINB) : BWDK(1, NW, U, MINB))
      switch (variant) {
        case 11: return BWDV(4, 2, 10);
        case 12: return BWDV(8, 4, 4);
        case 13: return BWDV(8, 2, 5);
        default: return BWDV(4, 4, 8);
      }
#undef BWDV
#undef BWDK
    }
  }
  const int lpg = fast_lpg<T>(d, feat);
  if (lpg && aligned(go, 16) && aligned(gf, 16) && (d.C * 4) % 16 == 0) {
    const bool tma = tma_ok(d, loc, w);
    const bool small = 32 * d.G <= 256;
#define CALL_BWD(N)                                                                                  \
  (tma ? (small ? launch_bwd_t<T, N, true, 256>(feat, shape, start, loc, w, go, gf, gl, gw, d,      \
                                               overwrite, st)                                      \
                : launch_bwd_t<T, N, true, 1024>(feat, shape, start, loc, w, go, gf, gl, gw, d,     \
                                                overwrite, st))                                    \
       : (small ? launch_bwd_t<T, N, false, 256>(feat, shape, start, loc, w, go, gf, gl, gw, d,     \
                                                overwrite, st)                                     \
                : launch_bwd_t<T, N, false, 1024>(feat, shape, start, loc, w, go, gf, gl, gw, d,    \
                                                 overwrite, st)))
    DFA_DISPATCH_LPG(CALL_BWD)
#undef CALL_BWD
  }
  dfa_bwd_generic_kernel<T><<<d.bs * d.A, 256, 0, st>>>(static_cast<const T *>(feat), shape, start,
                                                       loc, w, go, gf, gl, gw, d, overwrite);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

int dfa_version(void) { return DFA_B200_VERSION; }

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

int dfa_backward(const void *mc_ms_feat, int feat_dtype, const int32_t *spatial_shape,
                 const int32_t *scale_start_index, const float *sampling_location,
                 const float *weights, const float *grad_output, float *grad_mc_ms_feat,
                 float *grad_sampling_location, float *grad_weights, const dfa_dims *dims,
                 int flags, void *stream) {
  if (!mc_ms_feat || !spatial_shape || !scale_start_index || !sampling_location || !weights ||
      !grad_output || !grad_sampling_location || !grad_weights)
    return DFA_ERR_NULL_POINTER;  // grad_mc_ms_feat may be NULL: the feature gradient is skipped
  Dims d;
  if (int rc = check_dims(dims, d)) return rc;
  if (!aligned(sampling_location, 4) || !aligned(weights, 4) || !aligned(grad_output, 4) ||
      !aligned(grad_mc_ms_feat, 4) || !aligned(grad_sampling_location, 4) || !aligned(grad_weights, 4))
    return DFA_ERR_MISALIGNED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if ((flags & DFA_BWD_ZERO_GRAD_FEAT) && grad_mc_ms_feat) {
    cudaError_t e = cudaMemsetAsync(grad_mc_ms_feat, 0,
                                    sizeof(float) * static_cast<size_t>(d.bs) * d.num_feat * d.C, st);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  const int overwrite = (flags & DFA_BWD_OVERWRITE_SMALL) ? 1 : 0;
  if (feat_dtype == DFA_F32)
    return backward_typed<float>(mc_ms_feat, spatial_shape, scale_start_index, sampling_location,
                                 weights, grad_output, grad_mc_ms_feat, grad_sampling_location,
                                 grad_weights, d, overwrite, st);
  if (feat_dtype == DFA_BF16)
    return backward_typed<__nv_bfloat16>(mc_ms_feat, spatial_shape, scale_start_index,
                                         sampling_location, weights, grad_output, grad_mc_ms_feat,
                                         grad_sampling_location, grad_weights, d, overwrite, st);
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

int dfa_flatten_maps(const float *const *level_ptrs, const int32_t *level_hw, int num_levels,
                     int bs, int num_cams, int channels, void *col_feats, int out_dtype,
                     void *stream) {
  if (!level_ptrs || !level_hw || !col_feats) return DFA_ERR_NULL_POINTER;
  if (num_levels <= 0 || bs <= 0 || num_cams <= 0 || channels <= 0) return DFA_ERR_BAD_DIMS;
  if (out_dtype != DFA_F32 && out_dtype != DFA_BF16) return DFA_ERR_BAD_DTYPE;
  long long rows_per_cam = 0;
  for (int l = 0; l < num_levels; ++l) {
    if (!level_ptrs[l]) return DFA_ERR_NULL_POINTER;
    if (level_hw[2 * l] <= 0 || level_hw[2 * l + 1] <= 0) return DFA_ERR_BAD_DIMS;
    rows_per_cam += static_cast<long long>(level_hw[2 * l]) * level_hw[2 * l + 1];
  }
  if (rows_per_cam * num_cams >= (1ll << 31) || static_cast<long long>(bs) * num_cams > 65535)
    return DFA_ERR_BAD_DIMS;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int row0 = 0;
  for (int l = 0; l < num_levels; ++l) {
    const int HW = level_hw[2 * l] * level_hw[2 * l + 1];
    dim3 grid((HW + 31) / 32, (channels + 31) / 32, bs * num_cams);
    if (grid.y > 65535) return DFA_ERR_BAD_DIMS;
    if (out_dtype == DFA_F32)
      dfa_flatten_level_kernel<float><<<grid, 256, 0, st>>>(
          level_ptrs[l], static_cast<float *>(col_feats), HW, channels, num_cams,
          rows_per_cam * num_cams, static_cast<int>(rows_per_cam), row0);
    else
      dfa_flatten_level_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(
          level_ptrs[l], static_cast<__nv_bfloat16 *>(col_feats), HW, channels, num_cams,
          rows_per_cam * num_cams, static_cast<int>(rows_per_cam), row0);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return static_cast<int>(e);
    row0 += HW;
  }
  return 0;
}

int dfa_keypoints_project(const float *anchor, const float *fix_scale, int num_fix,
                          const float *learnable_logits, const float *projection_mat,
                          const float *image_wh, float *key_points, float *sampling_location,
                          int bs, int num_anchors, int num_pts, int num_cams, void *stream) {
  if (!anchor || !fix_scale || !projection_mat || !sampling_location) return DFA_ERR_NULL_POINTER;
  if (bs <= 0 || num_anchors <= 0 || num_pts <= 0 || num_cams <= 0 || num_fix < 0 || num_fix > num_pts)
    return DFA_ERR_BAD_DIMS;
  if (num_fix < num_pts && !learnable_logits) return DFA_ERR_NULL_POINTER;
  const long long n = static_cast<long long>(bs) * num_anchors * num_pts;
  if (n >= (1ll << 31)) return DFA_ERR_BAD_DIMS;
  dfa_keypoints_project_kernel<<<static_cast<int>((n + 255) / 256), 256, 0,
                                 static_cast<cudaStream_t>(stream)>>>(
      anchor, fix_scale, num_fix, learnable_logits, projection_mat, image_wh, key_points,
      sampling_location, bs, num_anchors, num_pts, num_cams);
  return static_cast<int>(cudaGetLastError());
}

int dfa_keypoints_project_backward(const float *anchor, const float *fix_scale, int num_fix,
                                   const float *learnable_logits, const float *projection_mat,
                                   const float *image_wh, const float *grad_sampling_location,
                                   float *grad_anchor, float *grad_learnable_logits, int bs,
                                   int num_anchors, int num_pts, int num_cams, void *stream) {
  if (!anchor || !fix_scale || !projection_mat || !grad_sampling_location || !grad_anchor)
    return DFA_ERR_NULL_POINTER;
  if (bs <= 0 || num_anchors <= 0 || num_pts <= 0 || num_cams <= 0 || num_fix < 0 || num_fix > num_pts)
    return DFA_ERR_BAD_DIMS;
  if (num_fix < num_pts && !learnable_logits) return DFA_ERR_NULL_POINTER;
  const long long n = static_cast<long long>(bs) * num_anchors;
  if (n * num_pts >= (1ll << 31)) return DFA_ERR_BAD_DIMS;
  dfa_keypoints_project_bwd_kernel<<<static_cast<int>((n + 127) / 128), 128, 0,
                                     static_cast<cudaStream_t>(stream)>>>(
      anchor, fix_scale, num_fix, learnable_logits, projection_mat, image_wh, grad_sampling_location,
      grad_anchor, grad_learnable_logits, bs, num_anchors, num_pts, num_cams);
  return static_cast<int>(cudaGetLastError());
}

namespace {
constexpr int SOFTMAX_NT = 256;
inline bool softmax_vec_ok(int G, const void *a, const void *b, const void *c, const void *e) {
  const int Q = G / 4;
  return G % 4 == 0 && (Q & (Q - 1)) == 0 && Q <= 8 && aligned(a, 16) && aligned(b, 16) &&
         aligned(c, 16) && aligned(e, 16);
}
int softmax_check(long long n_anchors, int K, int L, int P, int G, long long smem_floats, uint32_t *smem) {
  if (n_anchors <= 0 || K <= 0 || L <= 0 || P <= 0 || G <= 0) return DFA_ERR_BAD_DIMS;
  if (n_anchors >= (1ll << 31) || SOFTMAX_NT % G != 0) return DFA_ERR_UNSUPPORTED;
  if (static_cast<long long>(K) * L * P >= 65536) return DFA_ERR_UNSUPPORTED;  // 16-bit row table
  const long long bytes = 4ll * K * L * P * G * smem_floats + 3ll * K * L * P + 16;
  if (bytes > 200ll * 1024) return DFA_ERR_UNSUPPORTED;
  *smem = static_cast<uint32_t>(bytes);
  return 0;
}
}  // namespace

int dfa_softmax_weights(const float *logits, const uint8_t *keep_mask, float scale, float *weights,
                        int bs, int num_anchors, int num_cams, int num_scale, int num_pts,
                        int num_groups, void *stream) {
  if (!logits || !weights) return DFA_ERR_NULL_POINTER;
  uint32_t smem = 0;
  const long long n = static_cast<long long>(bs) * num_anchors;
  if (int rc = softmax_check(n, num_cams, num_scale, num_pts, num_groups, 1, &smem)) return rc;
  auto kern = softmax_vec_ok(num_groups, logits, weights, nullptr, nullptr)
                  ? dfa_softmax_weights4_kernel<SOFTMAX_NT>
                  : dfa_softmax_weights_kernel<SOFTMAX_NT>;
  if (int rc = set_smem(kern, smem)) return rc;
  kern<<<static_cast<int>(n), SOFTMAX_NT, smem, static_cast<cudaStream_t>(stream)>>>(
      logits, nullptr, keep_mask, scale, weights, num_anchors, num_cams, num_scale, num_pts, num_groups);
  return static_cast<int>(cudaGetLastError());
}

int dfa_softmax_weights_split(const float *logits_anchor, const float *logits_cam,
                              const uint8_t *keep_mask, float scale, float *weights, int bs,
                              int num_anchors, int num_cams, int num_scale, int num_pts,
                              int num_groups, void *stream) {
  if (!logits_anchor || !logits_cam || !weights) return DFA_ERR_NULL_POINTER;
  uint32_t smem = 0;
  const long long n = static_cast<long long>(bs) * num_anchors;
  if (int rc = softmax_check(n, num_cams, num_scale, num_pts, num_groups, 1, &smem)) return rc;
  auto kern = softmax_vec_ok(num_groups, logits_anchor, logits_cam, weights, nullptr)
                  ? dfa_softmax_weights4_kernel<SOFTMAX_NT>
                  : dfa_softmax_weights_kernel<SOFTMAX_NT>;
  if (int rc = set_smem(kern, smem)) return rc;
  kern<<<static_cast<int>(n), SOFTMAX_NT, smem, static_cast<cudaStream_t>(stream)>>>(
      logits_anchor, logits_cam, keep_mask, scale, weights, num_anchors, num_cams, num_scale, num_pts,
      num_groups);
  return static_cast<int>(cudaGetLastError());
}

int dfa_softmax_weights_backward(const float *logits, const uint8_t *keep_mask, float scale,
                                 const float *grad_weights, float *grad_logits, int bs,
                                 int num_anchors, int num_cams, int num_scale, int num_pts,
                                 int num_groups, void *stream) {
  if (!logits || !grad_weights || !grad_logits) return DFA_ERR_NULL_POINTER;
  uint32_t smem = 0;
  const long long n = static_cast<long long>(bs) * num_anchors;
  if (int rc = softmax_check(n, num_cams, num_scale, num_pts, num_groups, 2, &smem)) return rc;
  auto kern = softmax_vec_ok(num_groups, logits, grad_weights, grad_logits, nullptr)
                  ? dfa_softmax_weights4_bwd_kernel<SOFTMAX_NT>
                  : dfa_softmax_weights_bwd_kernel<SOFTMAX_NT>;
  if (int rc = set_smem(kern, smem)) return rc;
  kern<<<static_cast<int>(n), SOFTMAX_NT, smem, static_cast<cudaStream_t>(stream)>>>(
      logits, nullptr, keep_mask, scale, grad_weights, grad_logits, nullptr, num_anchors, num_cams,
      num_scale, num_pts, num_groups);
  return static_cast<int>(cudaGetLastError());
}

int dfa_softmax_weights_split_backward(const float *logits_anchor, const float *logits_cam,
                                       const uint8_t *keep_mask, float scale,
                                       const float *grad_weights, float *grad_logits_full,
                                       float *grad_logits_anchor, int bs, int num_anchors,
                                       int num_cams, int num_scale, int num_pts, int num_groups,
                                       void *stream) {
  if (!logits_anchor || !logits_cam || !grad_weights || !grad_logits_full || !grad_logits_anchor)
    return DFA_ERR_NULL_POINTER;
  uint32_t smem = 0;
  const long long n = static_cast<long long>(bs) * num_anchors;
  if (int rc = softmax_check(n, num_cams, num_scale, num_pts, num_groups, 2, &smem)) return rc;
  auto kern = softmax_vec_ok(num_groups, logits_anchor, logits_cam, grad_weights, grad_logits_full) &&
                      aligned(grad_logits_anchor, 16)
                  ? dfa_softmax_weights4_bwd_kernel<SOFTMAX_NT>
                  : dfa_softmax_weights_bwd_kernel<SOFTMAX_NT>;
  if (int rc = set_smem(kern, smem)) return rc;
  kern<<<static_cast<int>(n), SOFTMAX_NT, smem, static_cast<cudaStream_t>(stream)>>>(
      logits_anchor, logits_cam, keep_mask, scale, grad_weights, grad_logits_full, grad_logits_anchor,
      num_anchors, num_cams, num_scale, num_pts, num_groups);
  return static_cast<int>(cudaGetLastError());
}

int64_t dfa_forward_host_workspace_bytes(int feat_dtype, const dfa_dims *dims) {
  Dims d;
  if (check_dims(dims, d)) return -1;
  const int64_t esz = feat_dtype == DFA_BF16 ? 2 : 4;
  auto up = [](int64_t x) { return (x + 255) / 256 * 256; };
  return up(esz * d.bs * d.num_feat * d.C) + up(8ll * d.K * d.L) + up(4ll * d.K * d.L) +
         up(8ll * d.bs * d.A * d.P * d.K) + up(4ll * d.bs * d.A * d.P * d.K * d.L * d.G) +
         up(4ll * d.bs * d.A * d.C);
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
  void *d_feat = p; p += up(nb_feat);
  int32_t *d_shape = reinterpret_cast<int32_t *>(p); p += up(nb_shape);
  int32_t *d_start = reinterpret_cast<int32_t *>(p); p += up(nb_start);
  float *d_loc = reinterpret_cast<float *>(p); p += up(nb_loc);
  float *d_w = reinterpret_cast<float *>(p); p += up(nb_w);
  float *d_out = reinterpret_cast<float *>(p);
  cudaError_t e;
  // small operands first so the kernel's staging data is resident before the big copy ends
  if ((e = cudaMemcpyAsync(d_shape, h_shape, nb_shape, cudaMemcpyHostToDevice, st))) return e;
  if ((e = cudaMemcpyAsync(d_start, h_start, nb_start, cudaMemcpyHostToDevice, st))) return e;
  if ((e = cudaMemcpyAsync(d_loc, h_loc, nb_loc, cudaMemcpyHostToDevice, st))) return e;
  if ((e = cudaMemcpyAsync(d_w, h_w, nb_w, cudaMemcpyHostToDevice, st))) return e;
  if ((e = cudaMemcpyAsync(d_feat, h_feat, nb_feat, cudaMemcpyHostToDevice, st))) return e;
  if (int rc = dfa_forward(d_feat, feat_dtype, d_shape, d_start, d_loc, d_w, d_out, dims, stream))
    return rc;
  if ((e = cudaMemcpyAsync(h_out, d_out, nb_out, cudaMemcpyDeviceToHost, st))) return e;
  return static_cast<int>(cudaStreamSynchronize(st));
}

}  // extern "C"

// dfa_common.cuh — shared device / host helpers of libdfa_b200 (sm_100a only).
//
// PTX wrappers (mbarrier, TMA bulk copy, vector reductions), the tap geometry every kernel and the
// parity side channel share, 16-byte feature-vector access, shared-memory layouts and the host-side
// checks.  Everything lives in an anonymous namespace: each translation unit gets its own copy.
//
// Semantics follow /root/reference/projects/mmdet3d_plugin/ops/src/deformable_aggregation_cuda.cu
// (forward :129-187 + :13-59, backward :190-262 + :62-126); the pixel coordinate uses the single
// fused multiply-add the compiled reference uses (SURVEY.md §7 "bit-exact indices").
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <atomic>

#include "dfa_b200.h"

// Generation of the tuning-knob cache (defined in dfa_forward.cu): dfa_debug_reload_knobs() bumps it
// and every cached knob re-reads its environment variable on its next use.
extern std::atomic<int> dfa_knob_generation;

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
// First thing a CTA does: thread 0 starts the two TMA bulk copies of the anchor's operands, so that
// nothing else (level-table loads, index arithmetic) sits in front of that round trip.
template <bool TMA>
__device__ __forceinline__ void stage_issue(const float *__restrict__ loc_g, const float *__restrict__ w_g,
                                            float *s_w, float *s_loc, uint64_t *bars, int PK, int wcount) {
  if (TMA && threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
    mbar_expect_tx(&bars[0], 8u * PK);
    tma_bulk_g2s(s_loc, loc_g, 8u * PK, &bars[0]);
    mbar_expect_tx(&bars[1], 4u * wcount);
    tma_bulk_g2s(s_w, w_g, 4u * wcount, &bars[1]);
  }
}

// Wait for the locations (plain loads of both operands when TMA is off), compact valid samples.
// `issued`: stage_issue() already ran.  Returns n_valid.
template <bool TMA>
__device__ __forceinline__ int stage_and_compact(const float *__restrict__ loc_g,
                                                 const float *__restrict__ w_g, float *s_w,
                                                 float *s_loc, int *s_list, uint64_t *bars,
                                                 int *s_nvalid, int PK, int wcount,
                                                 bool issued = false) {
  const int tid = threadIdx.x;
  if (TMA) {
    if (!issued) stage_issue<TMA>(loc_g, w_g, s_w, s_loc, bars, PK, wcount);
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
// key points and camera projection (shared by the stand-alone front-end kernel and the fused forward)
// ------------------------------------------------------------------------------------------
// models/detection3d/blocks.py:181-207: size = exp(anchor[W,L,H]); offset = fix_scale[p] or
// sigmoid(logit) - 0.5; rotate by yaw, add the centre.  `lg` = this anchor's (P - num_fix) * 3
// learnable-offset logits (NULL when there are none).
__device__ __forceinline__ void key_point(const float *an, const float *fix_scale, int num_fix,
                                          const float *lg, int p, float &x, float &y, float &z) {
  const float sx = expf(an[3]), sy = expf(an[4]), sz = expf(an[5]);  // W, L, H
  float ox, oy, oz;
  if (p < num_fix) {
    ox = fix_scale[3 * p], oy = fix_scale[3 * p + 1], oz = fix_scale[3 * p + 2];
  } else {
    lg += (p - num_fix) * 3;
    ox = 1.f / (1.f + expf(-lg[0])) - 0.5f;
    oy = 1.f / (1.f + expf(-lg[1])) - 0.5f;
    oz = 1.f / (1.f + expf(-lg[2])) - 0.5f;
  }
  ox *= sx, oy *= sy, oz *= sz;
  const float sn = an[6], cs = an[7];
  x = fmaf(cs, ox, -sn * oy) + an[0];
  y = fmaf(sn, ox, cs * oy) + an[1];
  z = oz + an[2];
}

// models/blocks.py:198-213: m = 4x4 projection matrix (row-major), divide by clamp(depth, 1e-5) and
// by the image size (wh = (w, h), may be NULL).  The 4-term dot products are evaluated left to
// right with fused multiply-adds.
__device__ __forceinline__ void project_point(const float *m, const float *wh, float x, float y, float z,
                                              float &px, float &py) {
  const float u = fmaf(m[2], z, fmaf(m[1], y, m[0] * x)) + m[3];
  const float v = fmaf(m[6], z, fmaf(m[5], y, m[4] * x)) + m[7];
  const float dpt = fmaf(m[10], z, fmaf(m[9], y, m[8] * x)) + m[11];
  const float den = fmaxf(dpt, 1e-5f);
  px = u / den, py = v / den;
  if (wh) px /= wh[0], py /= wh[1];
}

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

// ------------------------------------------------------------------------------------------
// row-merging kernels: shared constants and shared-memory layout
// ------------------------------------------------------------------------------------------
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
  // A warp's share of the NW lists (MERGE_CAP + U entries: offsets, then 16-bit slots) lives in the warp's
  // own row -> slot table: the table is dead once the merge is done and is re-initialised by the next
  // round's merge.  6 bytes x (MERGE_CAP + U) <= 4 bytes x MERGE_TABLE.
  s.mine_stride = MERGE_CAP + U;
  s.mine_off = s.table, s.mine_slot = s.table + 4u * s.mine_stride;
  s.bar = o, o += 32;
  s.total = o;
  return s;
}

// T: feature type.  VPL: 16-byte vectors per lane per row (row bytes = 512 * VPL).  G: groups.

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

// Tuning knobs (tests and tools only): an environment variable is read ONCE per call site and
// process — not on every launch — and again only after dfa_debug_reload_knobs().  Racing host
// threads compute the same value; nothing is read from the environment on the launch path afterwards.
struct KnobCache {
  std::atomic<int> gen{-1};
  std::atomic<int> val{0};
};
constexpr int KNOB_UNSET = -0x7fffffff - 1;
inline int knob_get(KnobCache &c, const char *name, int dflt) {
  const int g = dfa_knob_generation.load(std::memory_order_relaxed);
  if (c.gen.load(std::memory_order_acquire) != g) {
    const char *e = getenv(name);
    c.val.store(e ? atoi(e) : KNOB_UNSET, std::memory_order_relaxed);
    c.gen.store(g, std::memory_order_release);
  }
  const int v = c.val.load(std::memory_order_relaxed);
  return v == KNOB_UNSET ? dflt : v;
}
#define DFA_KNOB(name, dflt)        \
  ([&]() -> int {                   \
    static KnobCache cache_;        \
    return knob_get(cache_, name, (dflt)); \
  }())

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

}  // namespace

// dfa_forward_win.cuh — warp-autonomous, window-merging forward kernel (included by
// dfa_forward.cu).  DESIGN.md §4.1.
//
// Same contract as the other forward kernels (reference: ops/src/deformable_aggregation_cuda.cu
// :129-187 + :13-59).  What bounds the gather on B200 is the SM's load path, not HBM: with one load
// per bilinear corner an anchor pulls ~230 feature rows of 1 KB through L1 of which only ~107 are
// distinct — its key points project into a few pixels of the coarse levels — and what bounds a
// small grid is the length of a CTA's chain of barriers and dependent load rounds.  So here
//
//   * a small CTA (4 warps) owns an anchor and its warps work on their own after ONE barrier: the
//     anchor's sampling locations arrive by a TMA bulk copy, warp k finds the valid key points of
//     camera k with one ballot (lane = key point), and from the camera masks alone every warp
//     derives the same deal of work units without talking to the others;
//   * fine levels are dealt tap by tap, round-robin over the warps.  A warp covers a whole feature
//     row with 16-byte vectors (VPL per lane: a warp instruction reads 512 contiguous bytes), so a
//     tap is 4 * VPL loads in flight per lane, the tap's geometry is computed once per warp and the
//     weighted sum stays in registers (packed FFMA2);
//   * a coarse level of a camera (map of at most `merge_pix` pixels) is one unit: its owner warp finds
//     the bounding window of the valid points' corners (four redux.sync); when the window has at
//     most 32 pixels the taps are MERGED before anything is loaded — lane = window pixel walks the
//     taps and adds bilinear(pixel, tap) * weight[tap][g] for the 8 groups into registers (no
//     shared-memory read-modify-write, no atomics, fixed key-point order) — and every touched pixel's
//     row is then loaded exactly once; larger windows (very close objects) fall back to taps;
//   * the warps' partial rows are folded through shared memory and the output row is written once
//     (no atomics, no zero-filled output).
//
// Summation order is fixed, so results are reproducible bit for bit.  Weights: the anchor's whole
// block by one TMA bulk copy issued together with the locations when the grid is small (latency
// rules), otherwise plain loads of the valid samples' weights next to the row loads (~80 % of the
// weight bytes never leave HBM with camera-rig inputs).
#pragma once

namespace {

constexpr int WIN_PIX = 32;     // largest merged window (one pixel per lane)
constexpr int WIN_MAP = 1024;   // levels whose maps have at most this many pixels are merge candidates

struct WinLayout {
  uint32_t loc, tab, vmask, plist, w, scr, bar, total;
  uint32_t scr_stride;  // per warp, coarse units: rec[32] float4, wts[32][8], coef[32][8] float, off[32] u32;
                        // fine levels (same space): off[32] uint4, bw[32] float4, widx[32] int
};
__host__ __device__ inline WinLayout win_layout(int P, int K, int L, int G, int C, int nw, bool whole) {
  WinLayout s;
  uint32_t o = 0;
  s.w = o, o = align_up(o + (whole ? 4u * P * K * L * G : 0u), 128);
  s.loc = o, o = align_up(o + 8u * P * K, 16);
  s.tab = o, o = align_up(o + 16u * K * L, 16);
  s.vmask = o, o = align_up(o + 4u * K, 16);
  s.plist = o, o = align_up(o + 32u * K, 16);
  s.scr_stride = 32u * 16u + 32u * 32u + 32u * 32u + 32u * 4u;  // fine records: 34 * 36 bytes fit as well
  const uint32_t scr = s.scr_stride * nw, fold = 4u * C * nw;  // the scratch doubles as fold buffer
  s.scr = o, o = align_up(o + (scr > fold ? scr : fold), 16);
  s.bar = o, o += 16;
  s.total = o;
  return s;
}

// …_cuda.cu:180-181 + :18-25 as compiled (see tap_geometry): the low corner and the two fractions
__device__ __forceinline__ void tap_low(float x, float y, int H, int W, int &h_low, int &w_low, float &lh,
                                        float &lw) {
  const float h_im = fmaf(y, static_cast<float>(H), -0.5f);
  const float w_im = fmaf(x, static_cast<float>(W), -0.5f);
  const float fh = floorf(h_im), fw = floorf(w_im);
  h_low = static_cast<int>(fh), w_low = static_cast<int>(fw);
  lh = h_im - fh, lw = w_im - fw;
}

// Tap record of (x, y) on a level: four corner byte offsets and bilinear weights.  Out-of-map
// corners (zero padding) are redirected to an in-map corner with weight 0 — a valid sample always
// has one — so the loads need no predicates.
__device__ __forceinline__ void win_record(const int4 tab, float x, float y, uint32_t rb, uint4 &off, float4 &bw) {
  int h_low, w_low;
  float lh, lw;
  tap_low(x, y, tab.x, tab.y, h_low, w_low, lh, lw);
  const bool hl = h_low >= 0, wl = w_low >= 0, hh = h_low + 1 <= tab.x - 1, wh = w_low + 1 <= tab.y - 1;
  const int base = tab.z + h_low * tab.y + w_low;
  const int r0 = (hl && wl) ? base : -1, r1 = (hl && wh) ? base + 1 : -1;
  const int r2 = (hh && wl) ? base + tab.y : -1, r3 = (hh && wh) ? base + tab.y + 1 : -1;
  const int safe = r0 >= 0 ? r0 : r1 >= 0 ? r1 : r2 >= 0 ? r2 : r3;
  const float ph = 1.f - lh, pw = 1.f - lw;
  off = make_uint4((r0 >= 0 ? r0 : safe) * rb, (r1 >= 0 ? r1 : safe) * rb, (r2 >= 0 ? r2 : safe) * rb,
                   (r3 >= 0 ? r3 : safe) * rb);
  bw = make_float4(r0 >= 0 ? ph * pw : 0.f, r1 >= 0 ? ph * lw : 0.f, r2 >= 0 ? lh * pw : 0.f,
                   r3 >= 0 ? lh * lw : 0.f);
}

__device__ __forceinline__ void prefetch_l2(const void *p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// TPI taps, whole warp: four corner rows each, VE vectors per lane per row, all 4 * VE * TPI loads
// issued back to back.  `fb` already includes the lane's byte offset inside a row; wv[u][v] = weight
// of tap u for the group of the lane's vector v.
template <typename T, int VE, int TPI>
__device__ __forceinline__ void win_taps(const unsigned char *__restrict__ fb, const uint4 (&off)[TPI],
                                         const float4 (&bw)[TPI], const float (&wv)[TPI][VE],
                                         float (&acc)[VE][FeatVec<T>::VEC]) {
  typename FeatVec<T>::raw_t val[TPI][4][VE];
#pragma unroll
  for (int u = 0; u < TPI; ++u) {
    const uint32_t o[4] = {off[u].x, off[u].y, off[u].z, off[u].w};
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int v = 0; v < VE; ++v)
        val[u][q][v] = FeatVec<T>::load_raw(reinterpret_cast<const T *>(fb + o[q] + 512u * v));
  }
#pragma unroll
  for (int u = 0; u < TPI; ++u) {
    const float b[4] = {bw[u].x, bw[u].y, bw[u].z, bw[u].w};
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int v = 0; v < VE; ++v) FeatVec<T>::fma(acc[v], b[q] * wv[u][v], val[u][q][v]);
  }
}

struct WinCtx {
  const float *s_w, *s_loc, *w_g;
  const int4 *s_tab;
  const unsigned char *s_plist;
  unsigned char *scr;  // this warp's scratch line
  unsigned myvm, cams;
  int K, L, merge_pix, whole;
  uint32_t rb;  // bytes per feature row
};

// One warp's share of an anchor: its fine-level taps, then the coarse-level units it owns.  The lane
// owns VE 16-byte vectors of a row (vector v at byte 512 * v from `fb`, channel group gv[v]); TPI
// taps are in flight together (4 * VE * TPI loads per lane).
template <typename T, int VE, int TPI, int NW>
__device__ __forceinline__ void win_work(const WinCtx &c, const unsigned char *__restrict__ fb,
                                         const int (&gv)[VE], float (&acc)[VE][FeatVec<T>::VEC], int lane,
                                         int warp) {
  constexpr int G = 8;
  constexpr int UM = 8 / VE;  // merged rows in flight per lane (8 loads, as for the taps)
  constexpr unsigned FULL = 0xffffffffu;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int LG = c.L * G;
  float4 *scr_rec = reinterpret_cast<float4 *>(c.scr);
  uint4 *rec_off = reinterpret_cast<uint4 *>(c.scr);  // fine-level tap records share the space
  float4 *rec_bw = reinterpret_cast<float4 *>(c.scr + 544);  // 34 records: a batch and its padding
  int *rec_widx = reinterpret_cast<int *>(c.scr + 1088);
  float *scr_wts = reinterpret_cast<float *>(c.scr + 512);
  float *scr_coef = reinterpret_cast<float *>(c.scr + 512 + 1024);
  unsigned *scr_off = reinterpret_cast<unsigned *>(c.scr + 512 + 2048);

  // The deal.  Work units of the anchor: its fine-level taps in (camera, level, valid point) order,
  // and one unit per (camera, coarse level).  Coarse units go to the warps from the last one down and
  // count as a few taps each; the fine taps are then dealt in contiguous blocks so that every warp
  // ends up with about the same load.  Every warp computes the same deal from the camera masks.
  const int ntk = __popc(c.myvm);
  int nf = 0;  // fine levels of camera `lane` (they must come first: FPN order)
  bool mono = true;
  if (ntk > 0) {
    bool seen_coarse = false;
    for (int l = 0; l < c.L; ++l) {
      const int4 tab = c.s_tab[lane * c.L + l];
      if (tab.x * tab.y > c.merge_pix) ++nf, mono = mono && !seen_coarse;
      else seen_coarse = true;
    }
  }
  const int merge_pix = __all_sync(FULL, mono) ? c.merge_pix : 0;  // odd level order: no merging
  if (merge_pix == 0) nf = ntk > 0 ? c.L : 0;
  const int fk = ntk * nf;
  int cum = fk;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(FULL, cum, o);
    if (lane >= o) cum += t;
  }
  const int F0 = cum - fk;  // first fine tap of camera `lane`
  const int n_fine = __shfl_sync(FULL, cum, 31);
  int my_start = 0, my_quota = 0;
  {
    int load[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) load[w] = 0;
    int u = 0, total = n_fine;
    for (unsigned cm = c.cams; cm; cm &= cm - 1) {
      const int k = __ffs(cm) - 1;
      const int nk = __shfl_sync(FULL, ntk, k), nfk = __shfl_sync(FULL, nf, k);
      for (int ci = 0; ci < c.L - nfk; ++ci, ++u) {
        const int wgt = min(ci == 0 ? 5 : 3, 1 + nk / 3);
        total += wgt;
#pragma unroll
        for (int w = 0; w < NW; ++w)
          if (((NW - 1 - u) & (NW - 1)) == w) load[w] += wgt;
      }
    }
    const int target = (total + NW - 1) / NW;
    int rest = n_fine, start = 0, quota[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      quota[w] = min(max(target - load[w], 0), rest);
      rest -= quota[w];
    }
    const int extra = (rest + NW - 1) / NW;  // (units heavier than the target: spread what is left)
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const int q = quota[w] + min(extra, rest);
      rest -= min(extra, rest);
      if (w == warp) my_start = start, my_quota = q;
      start += q;
    }
  }
  for (int base = 0; base < my_quota; base += 32) {
    const int n_mine = min(32, my_quota - base);
    const int t = my_start + base + lane;
    int kk = -1, r = 0, nn = 1;
    for (unsigned cm = c.cams; cm; cm &= cm - 1) {
      const int k = __ffs(cm) - 1;
      const int f0k = __shfl_sync(FULL, F0, k), fkk = __shfl_sync(FULL, fk, k), nk = __shfl_sync(FULL, ntk, k);
      if (t >= f0k && t < f0k + fkk) kk = k, r = t - f0k, nn = nk;
    }
    __syncwarp();  // the previous batch's readers are done
    if (lane < n_mine) {
      const int l = static_cast<int>((static_cast<float>(r) + 0.5f) * __frcp_rn(static_cast<float>(nn)));
      const int4 tab = c.s_tab[kk * c.L + l];
      const int s = c.s_plist[kk * 32 + (r - l * nn)] * c.K + kk;
      const float2 xy = *reinterpret_cast<const float2 *>(c.s_loc + 2 * s);
      uint4 off;
      float4 bw;
      win_record(tab, xy.x, xy.y, c.rb, off, bw);
      rec_off[lane] = off, rec_bw[lane] = bw, rec_widx[lane] = s * LG + l * G;
    }
    if (lane < TPI) {  // padding up to a multiple of TPI: row 0 of the table with zero weights
      rec_off[n_mine + lane] = make_uint4(0u, 0u, 0u, 0u), rec_bw[n_mine + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
      rec_widx[n_mine + lane] = 0;
    }
    __syncwarp();
    for (int i = 0; i < n_mine; i += TPI) {
      uint4 o4[TPI];
      float4 b4[TPI];
      float wv[TPI][VE];
#pragma unroll
      for (int u = 0; u < TPI; ++u) {
        o4[u] = rec_off[i + u], b4[u] = rec_bw[i + u];
        const int wi = rec_widx[i + u];
#pragma unroll
        for (int v = 0; v < VE; ++v) wv[u][v] = c.whole ? c.s_w[wi + gv[v]] : __ldg(c.w_g + wi + gv[v]);
      }
      win_taps<T, VE, TPI>(fb, o4, b4, wv, acc);
    }
  }
  // Coarse levels: one unit per (camera, level), dealt from the last warp down.
  int idx = 0;
  for (unsigned cm = c.cams; cm; cm &= cm - 1) {
    const int k = __ffs(cm) - 1;
    const unsigned vm = __shfl_sync(FULL, c.myvm, k);
    for (int l = __shfl_sync(FULL, nf, k); l < c.L; ++l) {
      if (((NW - 1 - idx++) & (NW - 1)) != warp) continue;
      const int4 tab = c.s_tab[k * c.L + l];
      const int nt = __popc(vm);
      const int j = __popc(vm & lt_mask);
      const bool v = (vm >> lane) & 1u;
      const int s = lane * c.K + k;
      // lane = key point: its tap's low corner, then the window of all in-map corners
      int h_low = 0, w_low = 0;
      float lh = 0.f, lw = 0.f;
      int xl = 0x7fffffff, yl = 0x7fffffff, xh = -0x7fffffff, yh = -0x7fffffff;
      if (v) {
        const float2 xy = *reinterpret_cast<const float2 *>(c.s_loc + 2 * s);
        tap_low(xy.x, xy.y, tab.x, tab.y, h_low, w_low, lh, lw);
        xl = max(w_low, 0), xh = min(w_low + 1, tab.y - 1);
        yl = max(h_low, 0), yh = min(h_low + 1, tab.x - 1);
      }
      xl = __reduce_min_sync(FULL, xl), yl = __reduce_min_sync(FULL, yl);
      xh = __reduce_max_sync(FULL, xh), yh = __reduce_max_sync(FULL, yh);
      const int ww = xh - xl + 1, area = ww * (yh - yl + 1);
      if (area > WIN_PIX) {  // a very close object: tap by tap
        for (unsigned pm = vm; pm; pm &= pm - 1) {
          const int s2 = (__ffs(pm) - 1) * c.K + k;
          const float2 xy = *reinterpret_cast<const float2 *>(c.s_loc + 2 * s2);
          uint4 o4[1];
          float4 b4[1];
          float wv[1][VE];
#pragma unroll
          for (int vv = 0; vv < VE; ++vv)
            wv[0][vv] = c.whole ? c.s_w[s2 * LG + l * G + gv[vv]] : __ldg(c.w_g + s2 * LG + l * G + gv[vv]);
          win_record(tab, xy.x, xy.y, c.rb, o4[0], b4[0]);
          win_taps<T, VE, 1>(fb, o4, b4, wv, acc);
        }
        continue;
      }
      // merge: tap records and the taps' weights -> scratch
      __syncwarp();
      if (v) {
        scr_rec[j] = make_float4(__int_as_float(w_low), __int_as_float(h_low), lw, lh);
        float4 w0, w1;
        if (c.whole) {
          const float4 *wp = reinterpret_cast<const float4 *>(c.s_w + s * LG + l * G);
          w0 = wp[0], w1 = wp[1];
        } else {
          const float4 *wp = reinterpret_cast<const float4 *>(c.w_g + s * LG + l * G);
          w0 = __ldg(wp), w1 = __ldg(wp + 1);
        }
        reinterpret_cast<float4 *>(scr_wts + j * G)[0] = w0;
        reinterpret_cast<float4 *>(scr_wts + j * G)[1] = w1;
      }
      __syncwarp();
      // lane = window pixel
      const int py = static_cast<int>((static_cast<float>(lane) + 0.5f) * __frcp_rn(static_cast<float>(ww)));
      const int px = xl + lane - py * ww, pyy = yl + py;
      float cf[G];
#pragma unroll
      for (int g = 0; g < G; ++g) cf[g] = 0.f;
      bool touched = false;
#pragma unroll 2
      for (int t = 0; t < nt; ++t) {
        const float4 r = scr_rec[t];
        const int dx = px - __float_as_int(r.x), dy = pyy - __float_as_int(r.y);
        const float wx = dx == 0 ? 1.f - r.z : r.z, wy = dy == 0 ? 1.f - r.w : r.w;
        const bool hit = static_cast<unsigned>(dx) < 2u && static_cast<unsigned>(dy) < 2u;
        const float bw = wy * wx;
        touched |= hit;
        const float4 w0 = reinterpret_cast<const float4 *>(scr_wts + t * G)[0],
                     w1 = reinterpret_cast<const float4 *>(scr_wts + t * G)[1];
        if (hit) {
          cf[0] = fmaf(bw, w0.x, cf[0]), cf[1] = fmaf(bw, w0.y, cf[1]);
          cf[2] = fmaf(bw, w0.z, cf[2]), cf[3] = fmaf(bw, w0.w, cf[3]);
          cf[4] = fmaf(bw, w1.x, cf[4]), cf[5] = fmaf(bw, w1.y, cf[5]);
          cf[6] = fmaf(bw, w1.z, cf[6]), cf[7] = fmaf(bw, w1.w, cf[7]);
        }
      }
      touched = touched && lane < area;
      const unsigned tm = __ballot_sync(FULL, touched);
      const int n_rows = __popc(tm);
      if (touched) {
        const int slot = __popc(tm & lt_mask);
        scr_off[slot] = static_cast<unsigned>(tab.z + pyy * tab.y + px) * c.rb;
        float4 *cp = reinterpret_cast<float4 *>(scr_coef + slot * G);
        cp[0] = make_float4(cf[0], cf[1], cf[2], cf[3]);
        cp[1] = make_float4(cf[4], cf[5], cf[6], cf[7]);
      }
      __syncwarp();
      // every distinct row of the window once
      for (int r0 = 0; r0 < n_rows; r0 += UM) {
        typename FeatVec<T>::raw_t val[UM][VE];
        float cv[UM][VE];
#pragma unroll
        for (int u = 0; u < UM; ++u) {
          const int r = min(r0 + u, n_rows - 1);  // the tail repeats the last row with weight 0
          const unsigned off = scr_off[r];
#pragma unroll
          for (int vv = 0; vv < VE; ++vv) {
            val[u][vv] = FeatVec<T>::load_raw(reinterpret_cast<const T *>(fb + off + 512u * vv));
            cv[u][vv] = r0 + u < n_rows ? scr_coef[r * G + gv[vv]] : 0.f;
          }
        }
#pragma unroll
        for (int u = 0; u < UM; ++u)
#pragma unroll
          for (int vv = 0; vv < VE; ++vv) FeatVec<T>::fma(acc[vv], cv[u][vv], val[u][vv]);
      }
    }
  }
}

// T: feature type.  VPL: 16-byte vectors of a row per lane (row bytes = 512 * VPL).  NW: warps per
// CTA.  Needs G == 8, P <= 32 (lane = key point), K <= 32.
//
// Channel split (VPL == 2): CTAs from `split_from` on share an anchor two ways, each gathering one
// half of the channels (half rows, two taps in flight), so their chain of dependent load rounds is
// half as long.  The launcher splits the anchors that start last: the kernel ends with the longest-
// lived of the CTAs that are running when the queue runs dry, and halves that tail.
template <typename T, int VPL, bool TMA, int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB)
    dfa_fwd_win_kernel(const T *__restrict__ feat, const int *__restrict__ shape,
                       const int *__restrict__ start, const float *__restrict__ loc,
                       const float *__restrict__ weights, float *__restrict__ out, Dims d, int merge_pix, int whole,
                       int split_from) {
  constexpr int VEC = FeatVec<T>::VEC;
  constexpr int NT = NW * 32;
  constexpr int G = 8;
  constexpr int C = 32 * VPL * VEC;
  constexpr unsigned FULL = 0xffffffffu;
  static_assert((NW & (NW - 1)) == 0, "warps per CTA: a power of two");
  extern __shared__ __align__(128) unsigned char smem[];
  const WinLayout lay = win_layout(d.P, d.K, d.L, G, C, NW, whole != 0);
  float *s_w = reinterpret_cast<float *>(smem + lay.w);
  float *s_loc = reinterpret_cast<float *>(smem + lay.loc);
  int4 *s_tab = reinterpret_cast<int4 *>(smem + lay.tab);
  unsigned *s_vmask = reinterpret_cast<unsigned *>(smem + lay.vmask);
  unsigned char *s_plist = smem + lay.plist;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + lay.bar);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool heavy = VPL == 2 && static_cast<int>(blockIdx.x) >= split_from;  // one of two CTAs of an anchor
  const bool helper = heavy && ((blockIdx.x - split_from) & 1);               // ... the upper half of the channels
  const int anchor = heavy ? split_from + ((blockIdx.x - split_from) >> 1) : blockIdx.x;  // b * A + a
  const int b = anchor / d.A;
  const int PK = d.P * d.K, LG = d.L * G, wcount = PK * LG, KL = d.K * d.L;
  const float *loc_g = loc + static_cast<size_t>(anchor) * PK * 2;
  const float *w_g = weights + static_cast<size_t>(anchor) * wcount;
  const unsigned lt_mask = (1u << lane) - 1u;
  constexpr uint32_t rb = 512u * VPL;  // bytes per feature row
  DFA_STAMP(0);
  DFA_GSTAMP(6);

  // ---- operands on their way, level tables, camera masks ----------------------------------------
  if (TMA) {
    if (tid == 0) {
      mbar_init(&bars[0], 1);
      mbar_init(&bars[1], 1);
      fence_mbar_init();
      mbar_expect_tx(&bars[0], 8u * PK);
      tma_bulk_g2s(s_loc, loc_g, 8u * PK, &bars[0]);
      if (whole) {
        mbar_expect_tx(&bars[1], 4u * wcount);
        tma_bulk_g2s(s_w, w_g, 4u * wcount, &bars[1]);
      }
    }
  } else {
    for (int i = tid; i < 2 * PK; i += NT) s_loc[i] = __ldg(loc_g + i);
  }
  for (int i = tid; i < KL; i += NT)
    s_tab[i] = make_int4(__ldg(shape + 2 * i), __ldg(shape + 2 * i + 1), __ldg(start + i), 0);
  __syncthreads();
  for (int k = warp; k < d.K; k += NW) {
    if (TMA) mbar_wait(&bars[0], 0);
    bool v = false;
    if (lane < d.P) {
      const float2 xy = *reinterpret_cast<const float2 *>(s_loc + 2 * (lane * d.K + k));
      v = sample_valid(xy.x, xy.y);
    }
    const unsigned vm = __ballot_sync(FULL, v);
    if (lane == 0) s_vmask[k] = vm;
    if (v) s_plist[k * 32 + __popc(vm & lt_mask)] = static_cast<unsigned char>(lane);  // j-th valid point
  }
  DFA_STAMP(1);
  __syncthreads();

  // ---- every warp on its own from here ---------------------------------------------------------------
  WinCtx c;
  c.myvm = lane < d.K ? s_vmask[lane] : 0u;
  c.cams = __ballot_sync(FULL, c.myvm != 0u);
  c.s_w = s_w, c.s_loc = s_loc, c.w_g = w_g, c.s_tab = s_tab, c.s_plist = s_plist;
  c.scr = smem + lay.scr + warp * lay.scr_stride;
  c.K = d.K, c.L = d.L, c.merge_pix = merge_pix, c.whole = whole, c.rb = rb;
  const unsigned char *fb = reinterpret_cast<const unsigned char *>(feat) +
                            static_cast<size_t>(b) * d.num_feat * rb + lane * 16;
  float acc[VPL][VEC];
#pragma unroll
  for (int v = 0; v < VPL; ++v)
#pragma unroll
    for (int cc = 0; cc < VEC; ++cc) acc[v][cc] = 0.f;
  if (c.cams) {
    if (TMA && whole) mbar_wait(&bars[1], 0);  // weights have landed
    DFA_STAMP(2);
    if (VPL == 2 && heavy) {  // this CTA's half of the channels: acc[0] only
      const int gv[1] = {(helper ? 4 : 0) + lane / 8};
      float (&acc1)[1][VEC] = reinterpret_cast<float (&)[1][VEC]>(acc[0]);
      win_work<T, 1, 2, NW>(c, fb + (helper ? 512 : 0), gv, acc1, lane, warp);
    } else {
      int gv[VPL];  // channel group of the lane's vector v
#pragma unroll
      for (int v = 0; v < VPL; ++v) gv[v] = (v * 32 + lane) / (4 * VPL);
      win_work<T, VPL, 1, NW>(c, fb, gv, acc, lane, warp);
    }
  } else {
    if (TMA && whole) mbar_wait(&bars[1], 0);  // never exit with the copy in flight
  }
  DFA_STAMP(4);
  __syncthreads();  // the scratch lines are dead: their space is the fold buffer now

  // ---- fold the warps' partial rows, write the output row (or this CTA's half of it) ----------------
  float *s_red = reinterpret_cast<float *>(smem + lay.scr);
  const int nv = heavy ? 1 : VPL;
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    if (v < nv) {
      float4 *o = reinterpret_cast<float4 *>(s_red + warp * C + (v * 32 + lane) * VEC);
#pragma unroll
      for (int cc = 0; cc < VEC / 4; ++cc)
        o[cc] = make_float4(acc[v][4 * cc], acc[v][4 * cc + 1], acc[v][4 * cc + 2], acc[v][4 * cc + 3]);
    }
  }
  __syncthreads();
  const int Cs = heavy ? C / 2 : C, ch0 = (heavy && helper) ? C / 2 : 0;
  for (int cc = tid; cc < Cs; cc += NT) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) sum += s_red[w * C + cc];
    out[static_cast<size_t>(anchor) * C + ch0 + cc] = sum;
  }
  DFA_STAMP(5);
  DFA_GSTAMP(7);
}

// Shapes the kernel takes: 8 groups, at most 32 key points and cameras, rows of 512 or 1024 bytes
// (a lane owns one or two 16-byte vectors).  Returns vectors per lane, 0 when the shape does not fit.
template <typename T>
int win_vpl(const Dims &d, const void *feat) {
  const long long rbytes = static_cast<long long>(d.C) * static_cast<long long>(sizeof(T));
  if (d.G != 8 || d.P > 32 || d.K > 32 || !aligned(feat, 16)) return 0;
  if (rbytes != 512 && rbytes != 1024) return 0;
  if (static_cast<long long>(d.num_feat) * rbytes >= (1ll << 32)) return 0;
  return static_cast<int>(rbytes / 512);
}

}  // namespace

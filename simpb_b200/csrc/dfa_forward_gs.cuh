// dfa_forward_gs.cuh — group-sliced, anchor-pooled forward kernel (included by dfa_forward.cu).
// DESIGN.md §4.1.
//
// Same contract as the other forward kernels (reference: ops/src/deformable_aggregation_cuda.cu
// :129-187 + :13-59).  What ends a single-wave launch of the one-CTA-per-anchor kernels is the
// heaviest anchor: the number of valid samples varies 4x between anchors, a CTA's life is a chain of
// dependent load rounds proportional to it, and the launch is as long as its longest chain (R50 bs=1:
// median CTA done after 12 us, the 26-sample anchors after 17 us).  Here the work is cut the other way:
//
//   * a CTA owns ONE channel group (C/G channels: a 128-byte slice of every fp32 feature row) of a
//     BLOCK of M consecutive anchors; the G CTAs of a block are launched side by side.  The launcher
//     sizes M so that the grid is a whole number of waves of resident CTAs (R50 bs=1: 111 blocks of
//     8-9 anchors x 8 groups = 888 CTAs on 148 SMs x 6);
//   * the block's sampling locations arrive by one TMA bulk copy; warp w compacts the valid samples of
//     anchors w, w+8, ... with ballots; a prefix sum over the anchors' tap counts gives every tap of the
//     block a position in one flat list;
//   * one thread per tap builds the tap's record for THIS group: four corner offsets and the four
//     coefficients bilinear x weight[g] (the weight is one 4-byte load, issued before the geometry is
//     computed; weights of invalid samples are never read).  No weight block is staged;
//   * the flat list is dealt to the warps in equal contiguous ranges, so the chain of every warp of
//     every CTA of the launch is the BLOCK's mean, not an anchor's.  A warp instruction gathers one tap:
//     lane = (corner, 16-byte vector of the slice), U taps in flight, packed FFMA2 into registers;
//   * where a range crosses from one anchor to the next the warp folds its registers (two shuffles per
//     channel).  An anchor that lies inside one warp's range is stored straight to the output; an
//     anchor cut by a range boundary leaves partial rows in shared memory which are summed in warp
//     order after one barrier.  No atomics, no zero-filled output, fixed summation order: the result
//     is reproducible bit for bit (it does depend on M, i.e. on the grid the launcher picked).
//
// More taps than the record buffer holds (dense inputs) are processed in several passes over whole
// anchors.  Row offsets are kept in 16-byte units from the start of the table, so a block may span
// batch items.
#pragma once

namespace {

constexpr int GS_NT = 256;      // threads per CTA
constexpr int GS_CAP = 672;     // tap records per pass (32 bytes each): 6 CTAs per SM inside 196 KB of shared memory
constexpr int GS_MMAX = 32;     // anchors per block the kernel supports (prefix sum by one warp)

struct GsLayout {
  uint32_t loc, list, nv, base, tab, rec, part, pj, seg, bar, total, list_stride;
};
__host__ __device__ inline GsLayout gs_layout(int P, int K, int L, int mmax, int cpg) {
  GsLayout s;
  const uint32_t PK = static_cast<uint32_t>(P) * K;
  uint32_t o = 0;
  s.rec = o, o = align_up(o + 32u * GS_CAP, 128);
  s.loc = o, o = align_up(o + 8u * PK * mmax, 16);
  s.list_stride = align_up(PK, 2);
  s.list = o, o = align_up(o + 2u * s.list_stride * mmax, 16);
  s.nv = o, o = align_up(o + 4u * (mmax + 1), 16);
  s.base = o, o = align_up(o + 4u * (mmax + 1), 16);
  s.tab = o, o = align_up(o + 12u * K * L, 16);
  s.part = o, o = align_up(o + 4u * (GS_NT / 32) * 2u * cpg, 16);
  s.pj = o, o = align_up(o + 4u * (GS_NT / 32) * 2u, 16);
  s.seg = o, o = align_up(o + 16u * (GS_NT / 32) * (mmax < GS_MMAX ? mmax : GS_MMAX), 16);
  s.bar = o, o += 16;
  s.total = o;
  return s;
}

// T: feature type.  LPS: 16-byte vectors per group slice ((C/G) * sizeof(T) / 16; 8 for SimPB's fp32
// table, 4 for bf16).  U: warp load instructions in flight (one covers 32 / (4 * LPS) taps).
template <typename T, int LPS, int U, bool TMA, int MINB>
__global__ void __launch_bounds__(GS_NT, MINB)
    dfa_fwd_gs_kernel(const T *__restrict__ feat, const int *__restrict__ shape, const int *__restrict__ start,
                      const float *__restrict__ loc, const float *__restrict__ weights, float *__restrict__ out,
                      Dims d, int nblocks, int mmax, int interleave) {
  constexpr int VEC = FeatVec<T>::VEC;
  constexpr int TPI = 32 / (4 * LPS);  // taps per warp instruction
  constexpr int NW = GS_NT / 32;
  extern __shared__ __align__(128) unsigned char smem[];
  const int cpg = d.C / d.G;
  const GsLayout lay = gs_layout(d.P, d.K, d.L, mmax, cpg);
  uint2 *s_rec = reinterpret_cast<uint2 *>(smem + lay.rec);  // [tap][corner] = (offset / 16, coefficient)
  float *s_loc = reinterpret_cast<float *>(smem + lay.loc);
  uint16_t *s_list = reinterpret_cast<uint16_t *>(smem + lay.list);
  int *s_nv = reinterpret_cast<int *>(smem + lay.nv);
  int *s_base = reinterpret_cast<int *>(smem + lay.base);  // [j] = taps of the block's anchors before j
  int *s_tab = reinterpret_cast<int *>(smem + lay.tab);
  float *s_part = reinterpret_cast<float *>(smem + lay.part);
  int *s_pj = reinterpret_cast<int *>(smem + lay.pj);
  int4 *s_seg = reinterpret_cast<int4 *>(smem + lay.seg);  // [warp][segment] = (first tap, end, destination, -)
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + lay.bar);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = blockIdx.x % d.G, blk = blockIdx.x / d.G;
  const long long total = static_cast<long long>(d.bs) * d.A;
  // anchors of this block: a0 + i * astep, i < M
  long long a0;
  int M, astep;
  if (interleave) {
    a0 = blk, astep = nblocks;
    M = static_cast<int>((total - blk + nblocks - 1) / nblocks);
  } else {
    a0 = total * blk / nblocks, astep = 1;
    M = static_cast<int>(total * (blk + 1) / nblocks - a0);
  }
  const int PK = d.P * d.K;
  DFA_STAMP(0);
  DFA_GSTAMP(6);

  if (TMA) {
    if (tid == 0) {
      mbar_init(bar, 1);
      fence_mbar_init();
      mbar_expect_tx(bar, 8u * PK * M);
      if (astep == 1) {
        tma_bulk_g2s(s_loc, loc + a0 * PK * 2, 8u * PK * M, bar);
      } else {
        for (int j = 0; j < M; ++j)
          tma_bulk_g2s(s_loc + j * PK * 2, loc + (a0 + static_cast<long long>(j) * astep) * PK * 2, 8u * PK, bar);
      }
    }
  } else {
    for (int i = tid; i < 2 * PK * M; i += GS_NT) {
      const int j = i / (2 * PK);
      s_loc[i] = __ldg(loc + (a0 + static_cast<long long>(j) * astep) * PK * 2 + (i - j * 2 * PK));
    }
  }
  for (int i = tid; i < d.K * d.L; i += GS_NT) {
    s_tab[3 * i] = __ldg(shape + 2 * i);
    s_tab[3 * i + 1] = __ldg(shape + 2 * i + 1);
    s_tab[3 * i + 2] = __ldg(start + i);
  }
  __syncthreads();  // barrier initialised (TMA) / locations stored (plain loads)
  if (TMA) mbar_wait(bar, 0);

  // valid samples per anchor (…_cuda.cu:168-171), warp w: anchors w, w + NW, ...
  for (int j = warp; j < M; j += NW) {
    const float *lj = s_loc + j * PK * 2;
    uint16_t *list = s_list + j * lay.list_stride;
    int n = 0;
    for (int b0 = 0; b0 < PK; b0 += 32) {
      const int s = b0 + lane;
      const bool v = s < PK && sample_valid(lj[2 * s], lj[2 * s + 1]);
      const unsigned m = __ballot_sync(0xffffffffu, v);
      if (v) list[n + __popc(m & ((1u << lane) - 1u))] = static_cast<uint16_t>(s);
      n += __popc(m);
    }
    if (lane == 0) s_nv[j] = n;
  }
  __syncthreads();
  {  // exclusive prefix sum of the anchors' tap counts (M <= 32); every warp computes and stores the same values
    const int nt = lane < M ? s_nv[lane] * d.L : 0;
    int inc = nt;
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, inc, m);
      if (lane >= m) inc += o;
    }
    if (lane < M) s_base[lane] = inc - nt;
    if (lane == M - 1) s_base[M] = inc;
    __syncwarp();
  }
  DFA_STAMP(1);

  const uint32_t rb16 = static_cast<uint32_t>(d.C) * sizeof(T) / 16u;  // row size in 16-byte units
  const int q = (lane / LPS) & 3, v = lane % LPS, sub = lane / (4 * LPS);
  const unsigned char *fb = reinterpret_cast<const unsigned char *>(feat) +
                            (static_cast<size_t>(grp) * cpg * sizeof(T) + 16u * v);

  for (int j0 = 0; j0 < M;) {
    // pass = the longest run of whole anchors whose taps fit the record buffer
    int j1 = j0 + 1;
    const int tb = s_base[j0];
    while (j1 < M && s_base[j1 + 1] - tb <= GS_CAP) ++j1;
    const int T_pass = s_base[j1] - tb;

    // ---- records: one thread per valid (anchor, sample); its L taps sit at l * nv + i (level-major:
    // coarse-level neighbours back to back).  Weight loads first, geometry while they are in flight.
    const int NS = T_pass / d.L;
    for (int i0 = tid; i0 < NS; i0 += GS_NT) {
      int j = j0;
      while ((s_base[j + 1] - tb) <= i0 * d.L) ++j;
      const int nv = s_nv[j];
      const int i = i0 - (s_base[j] - tb) / d.L;
      const int s = s_list[j * lay.list_stride + i];
      const long long an = a0 + static_cast<long long>(j) * astep;
      const float *wp = weights + (an * PK + s) * d.L * d.G + grp;
      const uint32_t item16 = static_cast<uint32_t>(an / d.A) * static_cast<uint32_t>(d.num_feat) * rb16;
      const float x = s_loc[(j * PK + s) * 2], y = s_loc[(j * PK + s) * 2 + 1];
      const int kl0 = (s % d.K) * d.L;
      uint4 *rec = reinterpret_cast<uint4 *>(s_rec + 4 * ((s_base[j] - tb) + i));
      for (int l0 = 0; l0 < d.L; l0 += 4) {
        float wv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) wv[u] = l0 + u < d.L ? __ldg(wp + (l0 + u) * d.G) : 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int l = l0 + u;
          if (l < d.L) {
            TapGeom gm;
            tap_geometry(x, y, s_tab[3 * (kl0 + l)], s_tab[3 * (kl0 + l) + 1], s_tab[3 * (kl0 + l) + 2], gm);
            // a corner outside the map (zero padding) is redirected to an in-map corner of the same tap
            // with coefficient 0 — a valid sample always has one — so the gather needs no predicates
            const int safe = gm.row[0] >= 0 ? gm.row[0] : gm.row[1] >= 0 ? gm.row[1]
                           : gm.row[2] >= 0 ? gm.row[2] : gm.row[3];
            const float bw[4] = {gm.hh * gm.hw, gm.hh * gm.lw, gm.lh * gm.hw, gm.lh * gm.lw};
            uint32_t o[4];
            float c[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              o[k] = item16 + static_cast<uint32_t>(gm.row[k] >= 0 ? gm.row[k] : safe) * rb16;
              c[k] = gm.row[k] >= 0 ? bw[k] * wv[u] : 0.f;
            }
            uint4 *dst = rec + 2 * l * nv;
            dst[0] = make_uint4(o[0], __float_as_uint(c[0]), o[1], __float_as_uint(c[1]));
            dst[1] = make_uint4(o[2], __float_as_uint(c[2]), o[3], __float_as_uint(c[3]));
          }
        }
      }
    }
    DFA_STAMP(2);
    if (lane < 2) s_pj[warp * 2 + lane] = -1;  // this warp's two partial-row slots: unused
    __syncwarp();
    // ---- the warp's share of the pass: taps [lo, hi), cut into one segment per anchor it meets.
    // lane i looks at anchor j0 + i (a pass has at most 32 anchors).
    int nseg;
    {
      int per = (T_pass + NW - 1) / NW;
      per = (per + U * TPI - 1) / (U * TPI) * (U * TPI);
      const int lo = min(warp * per, T_pass), hi = min(lo + per, T_pass);
      const int j = j0 + lane;
      int sb = 0, se = 0, ab = 0, ae = 0;
      if (j < j1) {
        ab = s_base[j] - tb, ae = s_base[j + 1] - tb;
        sb = max(ab, lo), se = min(ae, hi);
      }
      const bool live = se > sb, cut = live && (sb != ab || se != ae);
      const unsigned mlive = __ballot_sync(0xffffffffu, live), mcut = __ballot_sync(0xffffffffu, cut);
      const unsigned below = (1u << lane) - 1u;
      if (live) {
        // destination: the output row itself (float index), or partial-row slot 0 / 1 of this warp
        // (a range cuts at most two anchors: its first and its last)
        int dst;
        if (cut) {
          const int slot = __popc(mcut & below);
          dst = -1 - slot;
          s_pj[warp * 2 + slot] = j;
        } else {
          dst = static_cast<int>((a0 + static_cast<long long>(j) * astep) * d.C);
        }
        s_seg[warp * mmax + __popc(mlive & below)] = make_int4(sb, se, dst, 0);
      }
      nseg = __popc(mlive);
    }
    __syncthreads();
    DFA_STAMP(3);

    // ---- gather
    {
      const unsigned char *recl = reinterpret_cast<const unsigned char *>(s_rec) + 8 * q + 32 * sub;
#pragma unroll 1
      for (int sg = 0; sg < nseg; ++sg) {
        const int4 sd = s_seg[warp * mmax + sg];
        const int se = sd.y;
        float acc[VEC];
#pragma unroll
        for (int c = 0; c < VEC; ++c) acc[c] = 0.f;
        // full batches: U warp instructions = U * TPI taps, all loads issued before the first use
        int t = sd.x;
#pragma unroll 1
        for (; t + U * TPI <= se; t += U * TPI) {
          typename FeatVec<T>::raw_t val[U];
          float cf[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const uint2 rc = *reinterpret_cast<const uint2 *>(recl + 32 * (t + u * TPI));
            val[u] = FeatVec<T>::load_raw(reinterpret_cast<const T *>(fb + static_cast<size_t>(rc.x) * 16u));
            cf[u] = __uint_as_float(rc.y);
          }
#pragma unroll
          for (int u = 0; u < U; ++u) FeatVec<T>::fma(acc, cf[u], val[u]);
        }
        if (t < se) {  // last, partial batch: slots past the end replay the segment's last tap with coefficient 0
          typename FeatVec<T>::raw_t val[U];
          float cf[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int tt = t + u * TPI + sub;
            const uint2 rc = s_rec[4 * min(tt, se - 1) + q];
            val[u] = FeatVec<T>::load_raw(reinterpret_cast<const T *>(fb + static_cast<size_t>(rc.x) * 16u));
            cf[u] = tt < se ? __uint_as_float(rc.y) : 0.f;
          }
#pragma unroll
          for (int u = 0; u < U; ++u) FeatVec<T>::fma(acc, cf[u], val[u]);
        }
        // fold corners (and sub-taps): lanes that hold the same vector of the slice
#pragma unroll
        for (int m = LPS; m < 32; m <<= 1)
#pragma unroll
          for (int c = 0; c < VEC; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], m);
        if (lane < LPS) {
          float4 *dst = sd.z >= 0 ? reinterpret_cast<float4 *>(out + sd.z + grp * cpg + lane * VEC)
                                  : reinterpret_cast<float4 *>(s_part + (warp * 2 + (-1 - sd.z)) * cpg + lane * VEC);
#pragma unroll
          for (int c = 0; c < VEC / 4; ++c)
            dst[c] = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
        }
      }
    }
    DFA_STAMP(4);
    __syncthreads();

    // ---- anchors cut by a range boundary: sum the partial rows in warp order; anchors without a
    // valid sample: zeros
    for (int i = tid; i < (j1 - j0) * cpg; i += GS_NT) {
      const int j = j0 + i / cpg, c = i % cpg;
      float sum = 0.f;
      int hits = 0;
#pragma unroll
      for (int p = 0; p < 2 * NW; ++p)
        if (s_pj[p] == j) sum += s_part[p * cpg + c], ++hits;
      if (hits || s_nv[j] == 0)
        out[(a0 + static_cast<long long>(j) * astep) * d.C + grp * cpg + c] = sum;
    }
    j0 = j1;
    if (j0 < M) __syncthreads();  // records and partial rows are rewritten by the next pass
  }
  DFA_STAMP(5);
  DFA_GSTAMP(7);
}

// Shape test: a group slice is 1, 2, 4 or 8 16-byte vectors, rows and the table are 16-byte aligned,
// the table fits 32-bit offsets in 16-byte units, every anchor's taps fit one pass.  Returns LPS or 0.
template <typename T>
int gs_lps(const Dims &d, const void *feat, const float *out) {
  const long long sb = static_cast<long long>(d.C / d.G) * static_cast<long long>(sizeof(T));
  if (sb % 16 != 0 || !aligned(feat, 16) || !aligned(out, 16)) return 0;
  const long long lps = sb / 16;
  if (lps != 1 && lps != 2 && lps != 4 && lps != 8) return 0;
  if (static_cast<long long>(d.P) * d.K * d.L > GS_CAP || static_cast<long long>(d.P) * d.K >= 65536 || d.L > 15)
    return 0;
  if (static_cast<long long>(d.bs) * d.num_feat * d.C * static_cast<long long>(sizeof(T)) >= (1ll << 36)) return 0;
  if (((d.C / d.G) * 4) % 16 != 0) return 0;  // float4 stores of an output slice
  if (static_cast<long long>(d.bs) * d.A * d.C >= (1ll << 31)) return 0;  // 32-bit output index in the segment records
  return static_cast<int>(lps);
}

}  // namespace

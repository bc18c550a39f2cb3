// dfa_backward.cu — backward kernels (gradients wrt features, sampling locations and weights) and
// dfa_backward.  See DESIGN.md §4.2.
//
// grad_weights[b,a,p,k,l,g] is produced by exactly one thread / warp and
// grad_sampling_location[b,a,p,k,:] by exactly one CTA, in a fixed order (the reference: 32-way and
// 1024-way same-address float atomics).  Only grad_mc_ms_feat, where different anchors meet on one
// pixel, is scattered — with 128-bit vector reductions (red.global.add.v4.f32).
#include "dfa_common.cuh"

namespace {

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
  s.dot = s.m.coef;  // D[slot][g] overwrites coef[slot][g]: the warp that gathers a row is the only reader of
                     // its coefficients, and reads them (scatter) before it stores the dot products
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
  static_assert(6 * (MERGE_CAP + U) <= 4 * MERGE_TABLE, "a warp's share lists fit its row -> slot table");
  uint32_t *s_mine_off = reinterpret_cast<uint32_t *>(smem + lay.mine_off) + warp * MERGE_TABLE;
  uint16_t *s_mine_slot = reinterpret_cast<uint16_t *>(smem + lay.mine_slot + 4u * warp * MERGE_TABLE);
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
// launchers and dispatch
// ------------------------------------------------------------------------------------------
constexpr int BWD_U = 2;

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
#define DFA_DISPATCH_LPG(CALL)                    \
  switch (lpg) {                                  \
    case 8: return CALL(8);                       \
    case 4: return CALL(4);                       \
    case 2: return CALL(2);                       \
    default: return CALL(1);                      \
  }

template <typename T, int VPL, int NW, int U, bool TMA, int MINB>
int launch_bwd_merge(const void *feat, const int *shape, const int *start, const float *loc,
                     const float *w, const float *go, float *gf, float *gl, float *gw, const Dims &d,
                     int overwrite, cudaStream_t st) {
  auto kern = dfa_bwd_merge_kernel<T, VPL, 8, NW, U, TMA, MINB>;
  const MergeBwdLayout lay = merge_bwd_layout(d.P, d.K, d.L, d.G, NW, U);
  if (int rc = set_smem(kern, lay.total)) return rc;
  const long long grid = static_cast<long long>(d.bs) * d.A;
  const int whole = DFA_KNOB("DFA_BWD_WHOLE_WEIGHTS", grid <= 148 * 8 ? 1 : 0);
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
  const int variant = DFA_KNOB("DFA_BWD_VARIANT", 10);
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

extern "C" {

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

}  // extern "C"

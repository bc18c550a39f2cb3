// dfa_forward_pool.cuh — SM-pooled forward kernel (included by dfa_forward.cu).  DESIGN.md §4.1.
//
// Same contract as the other forward kernels (reference: ops/src/deformable_aggregation_cuda.cu
// :129-187 + :13-59), different decomposition.  The one-CTA-per-anchor kernels leave the machine
// unbalanced on small grids: an anchor's work is proportional to the number of its samples that
// fall inside a camera (median 14, maximum 26 of 78 with camera-rig inputs), so the heavy anchors
// finish alone, each with a sixth of an SM's loads in flight (ncu at bs=1: SMs active 63 % of the
// kernel's duration).  Here a persistent CTA owns the whole SM (or half of it) and POOLS the taps of
// a batch of anchors over all of its warps:
//
//   * batches of up to NB consecutive anchors; the first batch of a CTA is static (its TMA copy
//     starts with the kernel), the rest are handed out by a ticket counter in global memory (one
//     fetch-and-add per batch; big batches first, single anchors at the end), so SMs that drew light
//     anchors take more of them;
//   * the sampling locations of batch n+1 (TMA bulk copy on an mbarrier, two stages) are in flight
//     while batch n is gathered; the weights are NOT staged: a tap's weight is one 4-byte load next
//     to its 16-byte row loads, so only the weight lines of valid samples leave HBM and the CTA's
//     shared memory stays small enough to leave ~150 KB of the SM to L1, where the duplicate rows of
//     an anchor hit;
//   * an anchor's taps are cut into UNITS of Q_a consecutive taps (level-major order; Q_a = the
//     smallest Q0 * 2^k that gives the anchor at most POOL_CAPA units — a function of the anchor's
//     own sample count only); warps pull units from a shared-memory counter, a warp covers a whole
//     feature row with 16-byte vectors (VPL per lane), every unit's partial row goes to shared
//     memory, and an anchor's output is the sum of its unit rows in unit order.
//
// The summation order of an output element therefore depends only on the anchor's own samples —
// not on which anchors share the batch, which warp took which unit, or the order in which CTAs drew
// batches — so results are bitwise reproducible run to run and independent of the batch
// composition, although both schedules are dynamic.  No atomics on data, output written once.
#pragma once

namespace {

constexpr int POOL_NB_MAX = 8;          // anchors per batch (= compaction warps)
constexpr int POOL_CAPA = 16;           // units per anchor
constexpr int POOL_SCHED_SLOTS = 1024;  // self-resetting ticket counters, one slot per launch in flight
__device__ unsigned int g_pool_sched[POOL_SCHED_SLOTS][2];

#ifdef DFA_PHASE_TIMING
#define POOL_STAMP(cond, i)                                                                  \
  do {                                                                                       \
    if (g_phase_buf && (cond) && it < 16)                                                    \
      g_phase_buf[(static_cast<size_t>(blockIdx.x) * 16 + it) * 16 + (i)] = clock64();       \
  } while (0)
#else
#define POOL_STAMP(cond, i) do {} while (0)
#endif

struct PoolLayout {
  uint32_t loc, lstride, tab, list, slot, urec, uinfo, off, bw, widx, part, bar, misc, total;
};
__host__ __device__ inline PoolLayout pool_layout(int P, int K, int L, int C, int NB, int capt) {
  PoolLayout s;
  const uint32_t lbytes = 8u * P * K;
  const uint32_t capu = static_cast<uint32_t>(NB) * POOL_CAPA;
  uint32_t o = 0;
  s.lstride = align_up(NB * lbytes, 16), s.loc = o, o += 2 * s.lstride;  // two stages
  s.tab = o, o = align_up(o + 12u * K * L, 16);
  s.list = o, o = align_up(o + 2u * 4u * NB * P * K, 16);                // two stages
  s.slot = o, o = align_up(o + 2u * 3u * 4u * POOL_NB_MAX, 16);          // [stage][nv | qlog | nu][slot]
  s.urec = o, o = align_up(o + 4u * capu, 16);
  s.uinfo = o, o = align_up(o + 4u * capu, 16);
  s.off = o, o = align_up(o + 16u * capt, 16);
  s.bw = o, o = align_up(o + 16u * capt, 16);
  s.widx = o, o = align_up(o + 4u * capt, 16);
  s.part = o, o = align_up(o + 4u * C * capu, 16);
  s.bar = o, o += 16;
  s.misc = o, o += 32;  // [0..3] next batch {start, n} per stage, [4] unit counter
  s.total = o;
  return s;
}

struct PoolArgs {
  int NB;     // anchors per batch
  int capt;   // tap records per pass (shared-memory capacity)
  int q0log;  // log2 of the smallest unit size
  int nb0;    // size of every CTA's static first batch (gridDim.x * nb0 <= bs * A);
              // 0 = the whole problem is dealt out evenly: CTA i takes anchors [i*T/grid, (i+1)*T/grid)
  int dynamic;            // 0 = static round-robin batches, else a ticket counter hands out t_big batches
  int t_big, big, small;  // of `big` anchors, then batches of `small` anchors
  unsigned int *sched;    // {tickets drawn, CTAs done} — zero on entry, reset by the last CTA
};

__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// T: feature type.  VPL: 16-byte vectors of a row per lane (row bytes = 512 * VPL).  U: taps in
// flight per lane.  NT: threads per CTA (1024: one CTA per SM, 512: two).
template <typename T, int VPL, int U, int NT>
__global__ void __launch_bounds__(NT, NT == 512 ? 2 : 1)
    dfa_fwd_pool_kernel(const T *__restrict__ feat, const int *__restrict__ shape,
                        const int *__restrict__ start, const float *__restrict__ loc,
                        const float *__restrict__ weights, float *__restrict__ out, Dims d,
                        PoolArgs pa) {
  constexpr int VEC = FeatVec<T>::VEC;
  extern __shared__ __align__(128) unsigned char smem[];
  const int NB = pa.NB;
  const PoolLayout lay = pool_layout(d.P, d.K, d.L, d.C, NB, pa.capt);
  int *s_tab = reinterpret_cast<int *>(smem + lay.tab);
  int *s_urec = reinterpret_cast<int *>(smem + lay.urec);
  int *s_uinfo = reinterpret_cast<int *>(smem + lay.uinfo);
  uint4 *s_off = reinterpret_cast<uint4 *>(smem + lay.off);
  float4 *s_bw = reinterpret_cast<float4 *>(smem + lay.bw);
  int *s_widx = reinterpret_cast<int *>(smem + lay.widx);
  float *s_part = reinterpret_cast<float *>(smem + lay.part);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + lay.bar);  // one per location stage
  int *s_misc = reinterpret_cast<int *>(smem + lay.misc);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int PK = d.P * d.K, wcount = PK * d.L * d.G;
  const uint32_t lbytes = 8u * PK;
  const int total = d.bs * d.A;
  const int dyn0 = pa.nb0 == 0 ? total : gridDim.x * pa.nb0;  // anchors [0, dyn0): static first batches

#ifdef DFA_PHASE_TIMING
  if (g_phase_buf && tid == 0) {  // wall clock of the CTA's first instruction
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    g_phase_buf[(static_cast<size_t>(blockIdx.x) * 16 + 15) * 16 + 0] = static_cast<long long>(gt);
  }
#endif
  int cur_start = blockIdx.x * pa.nb0, cur_n = pa.nb0;
  if (pa.nb0 == 0) {
    cur_start = static_cast<int>(static_cast<long long>(blockIdx.x) * total / gridDim.x);
    cur_n = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * total / gridDim.x) - cur_start;
  }
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
    mbar_expect_tx(&bars[0], cur_n * lbytes);
    tma_bulk_g2s(smem + lay.loc, loc + static_cast<size_t>(cur_start) * PK * 2, cur_n * lbytes, &bars[0]);
  }
  for (int i = tid; i < d.K * d.L; i += NT) {
    s_tab[3 * i] = __ldg(shape + 2 * i);
    s_tab[3 * i + 1] = __ldg(shape + 2 * i + 1);
    s_tab[3 * i + 2] = __ldg(start + i);
  }
  __syncthreads();

  const uint32_t rb = 512u * VPL;  // bytes per feature row
  const uint32_t lane_off = static_cast<uint32_t>(lane) * 16u;
  const int cpg = d.C / d.G;
  int wg[VPL];  // this lane's channel group per vector
#pragma unroll
  for (int v = 0; v < VPL; ++v) wg[v] = ((lane + 32 * v) * VEC) / cpg;

  for (int it = 0; cur_n > 0; ++it) {
    const int st = it & 1;
    const uint32_t par = (it >> 1) & 1;
    POOL_STAMP(tid == 0, 0);
    // the ticket for the next batch is drawn now and looked at only after the tap records are built
    unsigned int ticket = 0;
    if (pa.dynamic && pa.nb0 != 0 && tid == NT - 32) ticket = atomicAdd(&pa.sched[0], 1u);

    const float *s_loc = reinterpret_cast<const float *>(smem + lay.loc + st * lay.lstride);
    int *s_list = reinterpret_cast<int *>(smem + lay.list) + st * NB * PK;
    int *s_nv = reinterpret_cast<int *>(smem + lay.slot) + st * 3 * POOL_NB_MAX;
    int *s_qlog = s_nv + POOL_NB_MAX, *s_nu = s_nv + 2 * POOL_NB_MAX;
    // compaction: warp s takes the batch's anchor s
    if (warp < cur_n) {
      mbar_wait(&bars[st], par);
      POOL_STAMP(tid == 0, 11);
      const float *lc = s_loc + warp * 2 * PK;
      int *lst = s_list + warp * PK;
      int n = 0;
      for (int base = 0; base < PK; base += 32) {
        const int s = base + lane;
        bool ok = false;
        if (s < PK) ok = sample_valid(lc[2 * s], lc[2 * s + 1]);
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (ok) lst[n + __popc(m & ((1u << lane) - 1u))] = s;
        n += __popc(m);
      }
      if (lane == 0) {  // unit size of this anchor: depends on its own sample count only
        const int ntaps = n * d.L;
        int ql = pa.q0log;
        while (((ntaps + (1 << ql) - 1) >> ql) > POOL_CAPA) ++ql;
        s_nv[warp] = n, s_qlog[warp] = ql, s_nu[warp] = (ntaps + (1 << ql) - 1) >> ql;
      }
    }
    __syncthreads();
    POOL_STAMP(tid == 0, 1);

    // passes: as many whole anchors as the record capacity holds (one pass unless most samples of
    // several anchors are valid)
    int s0 = 0;
    while (s0 < cur_n) {
      int s1 = s0, units = 0, nrec = 0;
      while (s1 < cur_n) {
        const int padded = s_nu[s1] << s_qlog[s1];
        if (s1 > s0 && nrec + padded > pa.capt) break;
        nrec += padded, units += s_nu[s1], ++s1;
      }
      if (tid == 0) s_misc[4] = 0;
      // tap records.  Padding taps of an anchor's last unit replay its tap 0 with zero weights;
      // out-of-map corners are redirected to an in-map corner of the same tap with a zero bilinear
      // weight, so the gather needs no predicates.
      for (int r = tid; r < nrec; r += NT) {
        int s = s0, rbase = 0, ubase = 0;
        for (;; ++s) {
          const int padded = s_nu[s] << s_qlog[s];
          if (r < rbase + padded) break;
          rbase += padded, ubase += s_nu[s];
        }
        const int nv = s_nv[s], ql = s_qlog[s];
        const int t = r - rbase;
        const bool live_tap = t < nv * d.L;
        const int tt = live_tap ? t : 0;
        const int l = tt / nv, i = tt - l * nv;  // level-major
        const int smp = s_list[s * PK + i];
        const int kl = (smp % d.K) * d.L + l;
        const float *lc = s_loc + s * 2 * PK;
        TapGeom gm;
        tap_geometry(lc[2 * smp], lc[2 * smp + 1], s_tab[3 * kl], s_tab[3 * kl + 1], s_tab[3 * kl + 2], gm);
        const int safe = gm.row[0] >= 0 ? gm.row[0] : gm.row[1] >= 0 ? gm.row[1]
                       : gm.row[2] >= 0 ? gm.row[2] : gm.row[3];
        const float live = live_tap ? 1.f : 0.f;
        uint4 off;
        float4 bw;
        off.x = (gm.row[0] >= 0 ? gm.row[0] : safe) * rb, bw.x = gm.row[0] >= 0 ? live * gm.hh * gm.hw : 0.f;
        off.y = (gm.row[1] >= 0 ? gm.row[1] : safe) * rb, bw.y = gm.row[1] >= 0 ? live * gm.hh * gm.lw : 0.f;
        off.z = (gm.row[2] >= 0 ? gm.row[2] : safe) * rb, bw.z = gm.row[2] >= 0 ? live * gm.lh * gm.hw : 0.f;
        off.w = (gm.row[3] >= 0 ? gm.row[3] : safe) * rb, bw.w = gm.row[3] >= 0 ? live * gm.lh * gm.lw : 0.f;
        s_off[r] = off, s_bw[r] = bw, s_widx[r] = (smp * d.L + l) * d.G;
        if ((t & ((1 << ql) - 1)) == 0) {
          const int u = ubase + (t >> ql);
          s_urec[u] = r, s_uinfo[u] = s | (ql << 8);
        }
      }
      __syncthreads();
      POOL_STAMP(tid == 0, 2);
      // next batch: resolve the ticket and start the copy of its locations (stage st^1 was last read
      // before the barrier that ended the previous iteration's record building)
      if (s0 == 0 && tid == NT - 32) {
        int ns = 0, nn = 0;
        if (pa.nb0 == 0) {
          ns = total, nn = 0;
        } else if (!pa.dynamic) {
          ns = dyn0 + (it * gridDim.x + blockIdx.x) * NB;
          nn = total - ns < NB ? total - ns : NB;
        } else if (static_cast<int>(ticket) < pa.t_big) {
          ns = dyn0 + static_cast<int>(ticket) * pa.big, nn = pa.big;
        } else {
          const long long first = static_cast<long long>(dyn0) + static_cast<long long>(pa.t_big) * pa.big +
                                  static_cast<long long>(ticket - pa.t_big) * pa.small;
          ns = first < total ? static_cast<int>(first) : total;
          nn = total - ns < pa.small ? total - ns : pa.small;
        }
        if (nn < 0) nn = 0;
        s_misc[2 * (st ^ 1)] = ns, s_misc[2 * (st ^ 1) + 1] = nn;
        if (nn > 0) {
          fence_proxy_async();
          mbar_expect_tx(&bars[st ^ 1], nn * lbytes);
          tma_bulk_g2s(smem + lay.loc + (st ^ 1) * lay.lstride, loc + static_cast<size_t>(ns) * PK * 2,
                       nn * lbytes, &bars[st ^ 1]);
        }
        POOL_STAMP(true, 10);
      }

      // gather: warps pull units
      for (;;) {
        int u = 0;
        if (lane == 0) u = atomicAdd(&s_misc[4], 1);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= units) break;
        const int info = s_uinfo[u];
        const int slot = info & 0xff, nq = 1 << (info >> 8);
        const int r0 = s_urec[u];
        const int anchor = cur_start + slot;
        const unsigned char *fb = reinterpret_cast<const unsigned char *>(feat) +
                                  static_cast<size_t>(anchor / d.A) * d.num_feat * d.C * sizeof(T) + lane_off;
        const float *wp = weights + static_cast<size_t>(anchor) * wcount;
        float acc[VPL][VEC];
#pragma unroll
        for (int v = 0; v < VPL; ++v)
#pragma unroll
          for (int c = 0; c < VEC; ++c) acc[v][c] = 0.f;
        for (int r = r0; r < r0 + nq; r += U) {
          typename FeatVec<T>::raw_t val[U][4][VPL];
          float wgt[U][VPL];
#pragma unroll
          for (int k = 0; k < U; ++k) {
            const uint4 off = s_off[r + k];
            const float *wq = wp + s_widx[r + k];
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
              val[k][0][v] = FeatVec<T>::load_raw(reinterpret_cast<const T *>(fb + (off.x + 512u * v)));
              val[k][1][v] = FeatVec<T>::load_raw(reinterpret_cast<const T *>(fb + (off.y + 512u * v)));
              val[k][2][v] = FeatVec<T>::load_raw(reinterpret_cast<const T *>(fb + (off.z + 512u * v)));
              val[k][3][v] = FeatVec<T>::load_raw(reinterpret_cast<const T *>(fb + (off.w + 512u * v)));
              wgt[k][v] = __ldg(wq + wg[v]);
            }
          }
#pragma unroll
          for (int k = 0; k < U; ++k) {
            const float4 bw = s_bw[r + k];
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
              FeatVec<T>::fma(acc[v], bw.x * wgt[k][v], val[k][0][v]);
              FeatVec<T>::fma(acc[v], bw.y * wgt[k][v], val[k][1][v]);
              FeatVec<T>::fma(acc[v], bw.z * wgt[k][v], val[k][2][v]);
              FeatVec<T>::fma(acc[v], bw.w * wgt[k][v], val[k][3][v]);
            }
          }
        }
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          float4 *pr = reinterpret_cast<float4 *>(s_part + u * d.C + (lane + 32 * v) * VEC);
#pragma unroll
          for (int c = 0; c < VEC / 4; ++c)
            pr[c] = make_float4(acc[v][4 * c], acc[v][4 * c + 1], acc[v][4 * c + 2], acc[v][4 * c + 3]);
        }
      }
      POOL_STAMP(tid == 0, 4);
      __syncthreads();
      POOL_STAMP(tid == 0, 5);
      // an anchor's output row = its unit rows summed in unit order
      for (int idx = tid; idx < (s1 - s0) * d.C; idx += NT) {
        const int sl = idx / d.C, c = idx - sl * d.C;
        int ub = 0;
        for (int s = s0; s < s0 + sl; ++s) ub += s_nu[s];
        const int nu = s_nu[s0 + sl];
        const float *pp = s_part + ub * d.C + c;
        float sum = 0.f;
        for (int u = 0; u < nu; ++u) sum += pp[u * d.C];
        out[static_cast<size_t>(cur_start + s0 + sl) * d.C + c] = sum;
      }
      s0 = s1;
      if (s0 < cur_n) __syncthreads();  // the next pass rewrites the records and the unit rows
      POOL_STAMP(tid == 0, 6);
    }
    cur_start = s_misc[2 * (st ^ 1)], cur_n = s_misc[2 * (st ^ 1) + 1];
  }
#ifdef DFA_PHASE_TIMING
  if (g_phase_buf && tid == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    g_phase_buf[(static_cast<size_t>(blockIdx.x) * 16 + 15) * 16 + 1] = static_cast<long long>(gt);
  }
#endif
  if (pa.dynamic && pa.nb0 != 0 && tid == NT - 32) {  // (the drawing thread) the last CTA out re-arms the ticket slot
    __threadfence();
    const unsigned int prev = atomicAdd(&pa.sched[1], 1u);
    if (prev == gridDim.x - 1) {
      pa.sched[0] = 0u, pa.sched[1] = 0u;
      __threadfence();
    }
  }
}

// Launcher.  Returns -1 when the shape does not fit (the caller falls through to the
// one-CTA-per-anchor kernels).
template <typename T, int VPL, int U, int NT>
int launch_fwd_pool(const void *feat, const int *shape, const int *start, const float *loc,
                    const float *w, float *out, const Dims &d, cudaStream_t st) {
  static std::atomic<unsigned int> next_slot{0};
  static std::mutex mu;
  static int sm_count[64] = {0};
  static unsigned int *sched_base[64] = {nullptr};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (!sched_base[dev]) {
      void *p = nullptr;
      if (cudaGetSymbolAddress(&p, g_pool_sched) != cudaSuccess) return -1;
      if (cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
      sched_base[dev] = static_cast<unsigned int *>(p);
    }
  }
  const long long total = static_cast<long long>(d.bs) * d.A;
  const int PKL = d.P * d.K * d.L;
  const uint32_t budget = (NT == 512 ? 100u : 200u) * 1024u;
  int NB = env_int("DFA_FWD_POOL_NB", NT == 512 ? 2 : 4);
  NB = NB < 1 ? 1 : (NB > POOL_NB_MAX ? POOL_NB_MAX : NB);
  int q0log = env_int("DFA_FWD_POOL_Q0LOG", 2);
  q0log = q0log < 0 ? 0 : (q0log > 6 ? 6 : q0log);
  while ((1 << q0log) < U) ++q0log;
  // one anchor's padded taps must fit a pass: units are at most twice as large as needed
  const int need = PKL + 2 * ((PKL + POOL_CAPA - 1) / POOL_CAPA) + (1 << q0log);
  int capt = env_int("DFA_FWD_POOL_CAPT", 640);
  if (capt < need) capt = need;
  while (NB > 1 && pool_layout(d.P, d.K, d.L, d.C, NB, capt).total > budget) NB /= 2;
  const PoolLayout lay = pool_layout(d.P, d.K, d.L, d.C, NB, capt);
  if (lay.total > budget) return -1;
  auto kern = dfa_fwd_pool_kernel<T, VPL, U, NT>;
  if (set_smem(kern, lay.total)) return -1;
  const long long max_ctas = static_cast<long long>(sm_count[dev]) * (NT == 512 ? 2 : 1);
  const int grid = static_cast<int>(total < max_ctas ? total : max_ctas);
  const int pct = env_int("DFA_FWD_POOL_STATIC_PCT", 67);
  long long nb0 = total * pct / 100 / grid;
  nb0 = nb0 < 1 ? 1 : (nb0 > NB ? NB : nb0);
  if (nb0 > total / grid) nb0 = total / grid;  // every CTA's static batch must exist (grid <= total)
  // a problem that fits one batch per CTA is dealt out evenly, no second iteration
  if (env_int("DFA_FWD_POOL_EVEN", 1) && (total + grid - 1) / grid <= NB) nb0 = 0;
  PoolArgs pa;
  pa.NB = NB, pa.capt = capt, pa.q0log = q0log, pa.nb0 = static_cast<int>(nb0);
  pa.dynamic = env_int("DFA_FWD_POOL_DYNAMIC", 1);
  {  // the last `tail` anchors per CTA go out one at a time, everything before in batches of `big`
    const long long dyn_total = total - static_cast<long long>(grid) * nb0;
    long long big = dyn_total / (4ll * grid);
    big = big < 1 ? 1 : (big > NB ? NB : big);
    const int tail = env_int("DFA_FWD_POOL_TAIL", 4);
    long long t_big = (dyn_total - static_cast<long long>(tail) * grid) / big;
    pa.big = static_cast<int>(big), pa.small = 1, pa.t_big = static_cast<int>(t_big < 0 ? 0 : t_big);
  }
  pa.sched = sched_base[dev] + 2 * (next_slot.fetch_add(1) % POOL_SCHED_SLOTS);
  kern<<<grid, NT, lay.total, st>>>(static_cast<const T *>(feat), shape, start, loc, w, out, d, pa);
  return static_cast<int>(cudaGetLastError());
}

// The pooled kernel applies when a feature row is 512 or 1024 bytes and every 16-byte vector lies
// inside one channel group.  Returns vectors per lane, 0 when the shape does not fit.
template <typename T>
int pool_vpl(const Dims &d, const void *feat, const float *loc, const float *out) {
  constexpr int VEC = FeatVec<T>::VEC;
  const long long rb = static_cast<long long>(d.C) * static_cast<long long>(sizeof(T));
  if (rb != 512 && rb != 1024) return 0;
  if ((d.C / d.G) % VEC != 0 || !aligned(feat, 16) || !aligned(loc, 16)) return 0;
  if ((8ll * d.P * d.K) % 16 != 0 || 8ll * d.P * d.K * POOL_NB_MAX >= (1ll << 20)) return 0;
  if (static_cast<long long>(d.num_feat) * rb >= (1ll << 32)) return 0;
  return static_cast<int>(rb / 512);
}

}  // namespace

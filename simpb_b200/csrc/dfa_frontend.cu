// dfa_frontend.cu — the kernels around the op: feature-map flattening, key points + camera
// projection, attention weights (softmax + attn-drop + permute), each with its backward, and their
// C ABI.  See DESIGN.md §4.3.
#include "dfa_common.cuh"

namespace {

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

// Vector form for HW % 4 == 0 and C % 4 == 0 (every SimPB level): a 64(pixels) x 64(channels) tile,
// 128-bit loads along the pixels of a channel, 128-bit (fp32) / 64-bit (bf16) stores along the
// channels of a pixel — a quarter of the memory instructions of the scalar kernel.
template <typename TO>
__global__ void __launch_bounds__(256)
    dfa_flatten_level_vec_kernel(const float *__restrict__ src, TO *__restrict__ dst, int HW, int C,
                                 int K, long long dst_rows_per_batch, int rows_per_cam, int level_row0) {
  __shared__ float tile[64][65];  // [pixel][channel]
  const int bk = blockIdx.z;
  const int b = bk / K, k = bk - b * K;
  const int p0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  const int tid = threadIdx.x;
  const float *s = src + static_cast<size_t>(bk) * C * HW;
  {
    const int p4 = tid & 15, cr = tid >> 4;  // 16 float4 along the pixels x 16 channels per pass
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = cr + 16 * i, p = p0 + 4 * p4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c0 + c < C && p < HW) v = __ldg(reinterpret_cast<const float4 *>(s + static_cast<size_t>(c0 + c) * HW + p));
      tile[4 * p4][c] = v.x, tile[4 * p4 + 1][c] = v.y, tile[4 * p4 + 2][c] = v.z, tile[4 * p4 + 3][c] = v.w;
    }
  }
  __syncthreads();
  TO *o = dst + (static_cast<size_t>(b) * dst_rows_per_batch +
                 static_cast<size_t>(k) * rows_per_cam + level_row0) * C;
  {
    const int c4 = tid & 15, pr = tid >> 4;  // 16 vectors along the channels x 16 pixels per pass
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int pl = pr + 16 * i, p = p0 + pl, c = c0 + 4 * c4;
      if (p < HW && c < C) {
        const float a0 = tile[pl][4 * c4], a1 = tile[pl][4 * c4 + 1], a2 = tile[pl][4 * c4 + 2],
                    a3 = tile[pl][4 * c4 + 3];
        TO *q = o + static_cast<size_t>(p) * C + c;
        if (sizeof(TO) == 4) {
          *reinterpret_cast<float4 *>(q) = make_float4(a0, a1, a2, a3);
        } else {
          const __nv_bfloat162 lo = __floats2bfloat162_rn(a0, a1), hi = __floats2bfloat162_rn(a2, a3);
          uint2 u;
          u.x = *reinterpret_cast<const uint32_t *>(&lo), u.y = *reinterpret_cast<const uint32_t *>(&hi);
          *reinterpret_cast<uint2 *>(q) = u;
        }
      }
    }
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
  float x, y, z;
  key_point(anchor + ba * 11, fix_scale, num_fix,
            logits ? logits + ba * (P - num_fix) * 3 : nullptr, p, x, y, z);
  if (kp_out) kp_out[3 * i] = x, kp_out[3 * i + 1] = y, kp_out[3 * i + 2] = z;
  for (int k = 0; k < K; ++k) {
    float px, py;
    project_point(proj + (static_cast<size_t>(b) * K + k) * 16, wh ? wh + (b * K + k) * 2 : nullptr, x, y, z,
                  px, py);
    float *o = loc_out + (static_cast<size_t>(i) * K + k) * 2;
    o[0] = px, o[1] = py;
  }
}

// Backward of the kernel above.  Sixteen lanes share an anchor: lane j takes key points j, j+16, ...
// (walking the K cameras of each), the partial sums are combined by a fixed xor-shuffle tree, so the
// gradients need no atomics and are bitwise reproducible.
// grad_anchor [bs,A,11] (velocity entries get 0), grad_logits [bs,A,(P-F)*3] (may be NULL).
__global__ void __launch_bounds__(128)
    dfa_keypoints_project_bwd_kernel(const float *__restrict__ anchor, const float *__restrict__ fix_scale,
                                     int num_fix, const float *__restrict__ logits,
                                     const float *__restrict__ proj, const float *__restrict__ wh,
                                     const float *__restrict__ grad_loc, float *__restrict__ grad_anchor,
                                     float *__restrict__ grad_logits, int bs, int A, int P, int K) {
  const long long gid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long ba_raw = gid >> 4;
  const int j = static_cast<int>(gid & 15);
  const bool active = ba_raw < static_cast<long long>(bs) * A;
  const long long ba = active ? ba_raw : 0;  // idle half-warps compute on anchor 0 and write nothing
  const int b = static_cast<int>(ba / A);
  const float *an = anchor + ba * 11;
  const float size[3] = {expf(an[3]), expf(an[4]), expf(an[5])};
  const float sn = an[6], cs = an[7];
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // centre xyz, size xyz, sin, cos
  for (int p = j; p < P; p += 16) {
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
      const float2 gl = __ldg(reinterpret_cast<const float2 *>(grad_loc) + (ba * P + p) * K + k);
      const float u = fmaf(m[2], z, fmaf(m[1], y, m[0] * x)) + m[3];
      const float v = fmaf(m[6], z, fmaf(m[5], y, m[4] * x)) + m[7];
      const float dpt = fmaf(m[10], z, fmaf(m[9], y, m[8] * x)) + m[11];
      const float den = fmaxf(dpt, 1e-5f);
      float gpx = gl.x, gpy = gl.y;
      if (wh) gpx /= wh[(b * K + k) * 2], gpy /= wh[(b * K + k) * 2 + 1];
      const float gu = gpx / den, gv = gpy / den;
      // d/d den of (u/den, v/den); the clamp passes the gradient where dpt >= 1e-5 (torch.clamp)
      const float gd = dpt >= 1e-5f ? -(gu * u + gv * v) / den : 0.f;
      gx += m[0] * gu + m[4] * gv + m[8] * gd;
      gy += m[1] * gu + m[5] * gv + m[9] * gd;
      gz += m[2] * gu + m[6] * gv + m[10] * gd;
    }
    acc[0] += gx, acc[1] += gy, acc[2] += gz;
    const float go[3] = {cs * gx + sn * gy, -sn * gx + cs * gy, gz};  // wrt the rotated-back offset
    acc[7] += ox * gx + oy * gy;
    acc[6] += -oy * gx + ox * gy;
#pragma unroll
    for (int i = 0; i < 3; ++i) acc[3 + i] += off[i] * go[i];
    if (active && p >= num_fix && grad_logits) {
      float *o = grad_logits + ba * (P - num_fix) * 3 + (p - num_fix) * 3;
#pragma unroll
      for (int i = 0; i < 3; ++i) o[i] = go[i] * size[i] * dsig[i];
    }
  }
#pragma unroll
  for (int mk = 8; mk > 0; mk >>= 1)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], mk);
  if (active && j == 0) {
    float *ga = grad_anchor + ba * 11;
    ga[0] = acc[0], ga[1] = acc[1], ga[2] = acc[2];
    ga[3] = acc[3] * size[0], ga[4] = acc[4] * size[1], ga[5] = acc[5] * size[2];  // d exp
    ga[6] = acc[6], ga[7] = acc[7], ga[8] = 0.f, ga[9] = 0.f, ga[10] = 0.f;
  }
}

// ------------------------------------------------------------------------------------------
// attention weights: softmax over (K, L, P) + attn-drop mask + permute, one pass
// ------------------------------------------------------------------------------------------
// logits [bs, A, K, L, P, G] (= weights_fc output, models/blocks.py:175-186) -> weights
// [bs, A, P, K, L, G] (the op's layout, models/blocks.py:133-144), softmax taken over the N =
// K*L*P entries of each (b, a, g).  keep [bs, A, K, P] (uint8, may be NULL) is the attn-drop keep
// mask of models/blocks.py:188-195 and `scale` its 1/(1-p).  One CTA per anchor; the anchor's
// logits are staged in shared memory once.  Thread t owns group t % G (G divides the block).
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

// The logits of one anchor are staged by TMA bulk copies (camera part or the full block straight into
// the working buffer, anchor part beside it): one round trip instead of one per loop iteration.
struct SoftmaxStage4 {
  float4 *x4;     // N*Q working buffer
  float4 *la4;    // LP*Q anchor part (split logits only)
  uint64_t *bar;
};
__device__ __forceinline__ SoftmaxStage4 softmax_stage4_layout(float *after_tables_base, int N, int n_el,
                                                               int lpg, float4 *x4) {
  // tables (3*N bytes) sit at `after_tables_base`; the anchor part and the barrier follow, 16-byte aligned
  unsigned char *p = reinterpret_cast<unsigned char *>(after_tables_base) + align_up(3u * N, 16);
  SoftmaxStage4 st;
  st.x4 = x4;
  st.la4 = reinterpret_cast<float4 *>(p);
  st.bar = reinterpret_cast<uint64_t *>(p + 4u * lpg);
  (void)n_el;
  return st;
}
__device__ __forceinline__ void softmax_issue4(const SoftmaxStage4 &st, const float *la_g, const float *lk_g,
                                               int tid, int n_el, int lpg) {
  if (tid == 0) {
    mbar_init(st.bar, 1);
    fence_mbar_init();
    if (lk_g) {
      mbar_expect_tx(st.bar, 4u * (n_el + lpg));
      tma_bulk_g2s(st.x4, lk_g, 4u * n_el, st.bar);
      tma_bulk_g2s(st.la4, la_g, 4u * lpg, st.bar);
    } else {
      mbar_expect_tx(st.bar, 4u * n_el);
      tma_bulk_g2s(st.x4, la_g, 4u * n_el, st.bar);
    }
  }
}

// numerators into st.x4, returns 1/sum for the thread's four groups.  Element e4 = i*NT + tid has
// quad tid % Q for every i (Q divides NT).
template <int NT>
__device__ __forceinline__ float4 softmax_stage4(const SoftmaxStage4 &st, float4 *s_red4, bool split, int tid,
                                                 int n4, int lp4, int Q) {
  __syncthreads();  // barrier initialised (and the tables written) for every thread
  mbar_wait(st.bar, 0);
  float4 mx = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  int ia = tid % lp4;
  for (int e = tid; e < n4; e += NT) {
    float4 v = st.x4[e];
    if (split) {
      const float4 a = st.la4[ia];
      ia += NT;
      while (ia >= lp4) ia -= lp4;
      v = make_float4(a.x + v.x, a.y + v.y, a.z + v.z, a.w + v.w);
      st.x4[e] = v;
    }
    mx = make_float4(fmaxf(mx.x, v.x), fmaxf(mx.y, v.y), fmaxf(mx.z, v.z), fmaxf(mx.w, v.w));
  }
  mx = quad_reduce<NT>(mx, s_red4, tid, Q, true);
  float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int e = tid; e < n4; e += NT) {
    float4 v = st.x4[e];
    v = make_float4(expf(v.x - mx.x), expf(v.y - mx.y), expf(v.z - mx.z), expf(v.w - mx.w));
    st.x4[e] = v;
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
  const SoftmaxStage4 st = softmax_stage4_layout(s_x + n_el, N, n_el, LP * G, s_x4);
  softmax_issue4(st, la, lk, tid, n_el, LP * G);
  softmax_tables<NT>(tb, keep ? keep + static_cast<size_t>(blockIdx.x) * K * P : nullptr, tid, K, L, P);
  const float4 inv = softmax_stage4<NT>(st, s_red4, lk != nullptr, tid, N * Q, LP * Q, Q);
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
  const SoftmaxStage4 st = softmax_stage4_layout(s_x + 2 * n_el, N, n_el, LP * G, s_x4);
  softmax_issue4(st, la, lk, tid, n_el, LP * G);
  softmax_tables<NT>(tb, keep ? keep + static_cast<size_t>(blockIdx.x) * K * P : nullptr, tid, K, L, P);
  const float4 inv = softmax_stage4<NT>(st, s_red4, lk != nullptr, tid, N * Q, LP * Q, Q);
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

}  // namespace

extern "C" {

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
    const bool vec = HW % 4 == 0 && channels % 4 == 0 && aligned(level_ptrs[l], 16) && aligned(col_feats, 16);
    if (vec) {
      dim3 g64((HW + 63) / 64, (channels + 63) / 64, bs * num_cams);
      if (g64.y > 65535) return DFA_ERR_BAD_DIMS;
      if (out_dtype == DFA_F32)
        dfa_flatten_level_vec_kernel<float><<<g64, 256, 0, st>>>(
            level_ptrs[l], static_cast<float *>(col_feats), HW, channels, num_cams,
            rows_per_cam * num_cams, static_cast<int>(rows_per_cam), row0);
      else
        dfa_flatten_level_vec_kernel<__nv_bfloat16><<<g64, 256, 0, st>>>(
            level_ptrs[l], static_cast<__nv_bfloat16 *>(col_feats), HW, channels, num_cams,
            rows_per_cam * num_cams, static_cast<int>(rows_per_cam), row0);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return static_cast<int>(e);
      row0 += HW;
      continue;
    }
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
  if (n * num_pts >= (1ll << 31) || n * 16 >= (1ll << 31)) return DFA_ERR_BAD_DIMS;
  if (!aligned(grad_sampling_location, 8)) return DFA_ERR_MISALIGNED;  // read as (x, y) pairs
  dfa_keypoints_project_bwd_kernel<<<static_cast<int>((n * 16 + 127) / 128), 128, 0,
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
  // working buffers + row tables + (vector kernels) the TMA-staged anchor part and its mbarrier
  const long long bytes = 4ll * K * L * P * G * smem_floats + 3ll * K * L * P + 16 + 4ll * L * P * G + 32;
  if (4ll * (K + 1) * L * P * G >= (1ll << 20)) return DFA_ERR_UNSUPPORTED;
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

}  // extern "C"

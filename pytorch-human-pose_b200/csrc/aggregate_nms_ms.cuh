// Multi-scale / general-ratio fused aggregation + NMS kernel (BASELINE config 3: test scales 0.5/1.0/1.5).
// Included by aggregate_nms.cu after aggregate_nms_x2.cuh (needs AggArgs, x2::stage helpers).
//
// Same column-walking scheme as the x2 kernel (each lane owns 4 adjacent output columns and walks down
// the rows with everything in registers), generalised:
//   * every scale s has its own half-resolution stage-mean tile S_s in shared memory (lo --x2--> hi is
//     always an exact x2 in HigherHRNet; hi --> output uses the general torch taps, e.g. x4, x2, x4/3);
//   * a lane keeps, per scale, the column taps of its 4 columns in registers and a two-row cache of
//     horizontally interpolated S_s rows; the vertical taps of every (scale, output row) come from a small
//     shared table; the scales are summed sequentially and divided by n like torch.stack(...).mean(0);
//   * tags come from one scale and must be an exact x4 (as in the reference); NMS, word side arrays and tag
//     bounds are the x2 kernel's.
// Bit-identical to the generic kernel (tests/test_gpu_stages.py).
#pragma once

namespace ms {

using x2::clampi;
using x2::max3;
using x2::tap_x2;

constexpr int RB = 16;            // output rows per CTA (multiple of 4)
constexpr int NROWS = RB + 4;     // walked rows: band + 2 halo rows on each side
constexpr int TR = RB / 4 + 2;    // quarter-res tag rows staged

struct Geom {                     // shared-memory layout, computed on the host
  int off_s[HPD_MAX_SCALES];      // float offset of S_s
  int hr[HPD_MAX_SCALES], hc[HPD_MAX_SCALES];   // rows / (even) row stride of S_s
  int off_lo, lr, lc;             // quarter-res staging tile (shared by the scales)
  int off_edge, off_tab, off_rt2, off_ninf, total;
  int hr_max;
};

// Stage an in-image window of nrows x (4 * nvec) columns starting at the 4-aligned column xa into shared
// memory (row stride a multiple of 4): one warp per row, 16-byte loads (the flipped operand as the mirrored
// 16 bytes, reversed in registers), all loads of a row in flight before the first store.
template <int NW>
__device__ __forceinline__ void stage_window(float* __restrict__ dst, int dst_stride, const float* __restrict__ p,
                                             const float* __restrict__ pf, int ya, int xa, int nrows, int nvec,
                                             int w, int warp, int lane) {
  constexpr int MAXIT = 4;
  for (int r = warp; r < nrows; r += NW) {
    const int rowo = (ya + r) * w;
    for (int qb = 0; qb < nvec; qb += 32 * MAXIT) {
      float4 v[MAXIT], f[MAXIT];
#pragma unroll
      for (int u = 0; u < MAXIT; ++u) {
        const int q = qb + lane + 32 * u;
        if (q < nvec) {
          const int x = xa + 4 * q;
          v[u] = __ldg(reinterpret_cast<const float4*>(p + rowo + x));
          if (pf) f[u] = __ldg(reinterpret_cast<const float4*>(pf + rowo + (w - 4 - x)));
        }
      }
#pragma unroll
      for (int u = 0; u < MAXIT; ++u) {
        const int q = qb + lane + 32 * u;
        if (q < nvec) {
          float4 o = v[u];
          if (pf) {
            o.x = __fmul_rn(__fadd_rn(v[u].x, f[u].w), 0.5f);
            o.y = __fmul_rn(__fadd_rn(v[u].y, f[u].z), 0.5f);
            o.z = __fmul_rn(__fadd_rn(v[u].z, f[u].y), 0.5f);
            o.w = __fmul_rn(__fadd_rn(v[u].w, f[u].x), 0.5f);
          }
          *reinterpret_cast<float4*>(dst + r * dst_stride + 4 * q) = o;
        }
      }
    }
  }
}

template <int E, int NW, int NS>
__global__ void __launch_bounds__(32 * NW, 4) agg_nms_ms_kernel(const AggArgs a, const Geom g) {
  extern __shared__ __align__(16) float smem[];
  constexpr int tc = 32 * NW + 8;
  constexpr int nthr = 32 * NW;
  float* sT = smem;                                   // tags alias the S tiles (consumed first)
  float* sLo = smem + g.off_lo;
  float* sEdge = smem + g.off_edge;
  // out-row -> S_s row taps: [NS][NROWS] x (i0, i1, w0, w1)
  int* rt_i0 = (int*)(smem + g.off_tab);
  int* rt_i1 = rt_i0 + NS * NROWS;
  float* rt_w0 = (float*)(rt_i1 + NS * NROWS);
  float* rt_w1 = rt_w0 + NS * NROWS;
  // phase 2 taps: S_s row -> lo tile rows, [hr_max] x (i0, i1, w0, w1)
  int* p2_i0 = (int*)(smem + g.off_rt2);
  int* p2_i1 = p2_i0 + g.hr_max;
  float* p2_w0 = (float*)(p2_i1 + g.hr_max);
  float* p2_w1 = p2_w0 + g.hr_max;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bk = blockIdx.z, b = bk / a.K, k = bk % a.K, kf = a.flip[k];
  const int H = a.H, W = a.W;
  const int x0 = blockIdx.x * 128 * NW, y0 = blockIdx.y * RB;
  const int xw = x0 + 128 * warp;
  const int X0 = xw + 4 * lane;
  const bool active = X0 < W;
  const int ys = y0 - 2;
  const float NINF = -INFINITY;

  // ---------------- phase 0: tags (x4), identical to the x2 kernel ------------------------------------
  {
    const int txo = x0 / 4 - 4, tyo = y0 / 4 - 1;
    const float* tg0 = a.tag + (size_t)b * a.tag_sb + (size_t)k * a.tag_sc;
    const bool unflip = !a.tags_preflipped;
    const float* tg1 = (E > 1) ? a.tag_f + (size_t)b * a.tagf_sb + (size_t)(unflip ? kf : k) * a.tagf_sc : nullptr;
    x2::stage_tile<0, TR, tc, NW, false>(sT, tg0, nullptr, false, tyo, txo, a.th, a.tw, warp, lane);
    if (E > 1) x2::stage_tile<1, TR, tc, NW, false>(sT + TR * tc, nullptr, tg1, unflip, tyo, txo, a.th, a.tw, warp, lane);
    __syncthreads();
    if (active) {
      const int q = X0 >> 2;
      const int tb = q - 1 - txo;
      const bool x_first = (q == 0);
      const float wa0 = x_first ? 1.f : 0.375f, wb0 = x_first ? 0.f : 0.625f;
      const float wa1 = x_first ? 1.f : 0.125f, wb1 = x_first ? 0.f : 0.875f;
      float hA[E][4], hB[E][4];
      auto hpass = [&](int r, float (&h)[E][4], float& mn, float& mx) {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const float* t = sT + e * TR * tc + r * tc + tb;
          const float t0 = t[0], t1 = t[1], t2 = t[2];
          if (e == 0) { mn = fminf(fminf(t0, t1), t2); mx = fmaxf(fmaxf(t0, t1), t2); }
          const float a0 = x_first ? t1 : t0, b0 = x_first ? t2 : t1;
          h[e][0] = fmaf(wa0, a0, __fmul_rn(wb0, b0));
          h[e][1] = fmaf(wa1, a0, __fmul_rn(wb1, b0));
          h[e][2] = fmaf(0.875f, t1, __fmul_rn(0.125f, t2));
          h[e][3] = fmaf(0.625f, t1, __fmul_rn(0.375f, t2));
        }
      };
      float* tg_plane = a.agg_tags + ((size_t)b * a.K + k) * H * W * E;
      auto emit = [&](int y, float wy0, float wy1, const float (&A)[E][4], const float (&Bv)[E][4]) {
        if (y < y0 || y >= y0 + RB || y >= H) return;
        float o[E][4];
#pragma unroll
        for (int e = 0; e < E; ++e)
#pragma unroll
          for (int c = 0; c < 4; ++c) o[e][c] = fmaf(wy0, A[e][c], __fmul_rn(wy1, Bv[e][c]));
        float* dst = tg_plane + ((size_t)y * W + X0) * E;
        if (E == 1) {
          *reinterpret_cast<float4*>(dst) = make_float4(o[0][0], o[0][1], o[0][2], o[0][3]);
        } else {
          *reinterpret_cast<float4*>(dst) = make_float4(o[0][0], o[E - 1][0], o[0][1], o[E - 1][1]);
          *reinterpret_cast<float4*>(dst + 4) = make_float4(o[0][2], o[E - 1][2], o[0][3], o[E - 1][3]);
        }
      };
      const unsigned amask = __activemask();
      float mnA, mxA, mnB, mxB, mnC, mxC;
      hpass(0, hA, mnA, mxA);
      hpass(1, hB, mnB, mxB);
      if (y0 == 0) {
        float hC[E][4];
        hpass(2, hC, mnC, mxC);
        emit(0, 1.f, 0.f, hB, hC);
        emit(1, 1.f, 0.f, hB, hC);
      } else {
        emit(y0, 0.375f, 0.625f, hA, hB);
        emit(y0 + 1, 0.125f, 0.875f, hA, hB);
      }
      const int HB = (H + 3) >> 2;
      const size_t bbase = (((size_t)b * a.K + k) * HB + (y0 >> 2)) * a.wpr + (X0 >> 5);
#pragma unroll 1
      for (int gq = 0; gq < RB / 4; ++gq) {
#pragma unroll
        for (int e = 0; e < E; ++e)
#pragma unroll
          for (int c = 0; c < 4; ++c) hA[e][c] = hB[e][c];
        hpass(gq + 2, hB, mnC, mxC);
        if (y0 + 4 * gq < H) {
          float mn = fminf(fminf(mnA, mnB), mnC), mx = fmaxf(fmaxf(mxA, mxB), mxC);
#pragma unroll
          for (int o = 1; o < 8; o <<= 1) {
            mn = fminf(mn, __shfl_xor_sync(amask, mn, o));
            mx = fmaxf(mx, __shfl_xor_sync(amask, mx, o));
          }
          if ((lane & 7) == 0) {
            a.tmin[bbase + (size_t)gq * a.wpr] = mn;
            a.tmax[bbase + (size_t)gq * a.wpr] = mx;
          }
        }
        mnA = mnB; mxA = mxB; mnB = mnC; mxB = mxC;
        const int y = y0 + 4 * gq + 2;
        emit(y, 0.875f, 0.125f, hA, hB);
        emit(y + 1, 0.625f, 0.375f, hA, hB);
        if (gq + 1 < RB / 4) {
          emit(y + 2, 0.375f, 0.625f, hA, hB);
          emit(y + 3, 0.125f, 0.875f, hA, hB);
        }
      }
    }
    __syncthreads();
  }

  // ---------------- phases 1+2 per scale: S_s = (up2(avg lo) + avg hi) * 0.5 in shared memory ------------
  const int oya = max(y0 - 2, 0), oyb = min(y0 + RB + 2, H) - 1;
  const int oxa = max(x0 - 2, 0), oxb = min(x0 + 128 * NW + 2, W) - 1;
  // (a real loop over the scales: the body is large and runs once per CTA -- unrolling it only thrashes
  // the instruction cache)
  __shared__ int s_hxa[HPD_MAX_SCALES];
#pragma unroll 1
  for (int s = 0; s < NS; ++s) {
    const ScaleDev& S = a.sc[s];
    float* sS = smem + g.off_s[s];
    const int hc = g.hc[s];
    const int hya = axis_tap(S.s_hi_y, oya, S.hh, H).i0, hyb = axis_tap(S.s_hi_y, oyb, S.hh, H).i1;
    // column origins are rounded down to a multiple of 4 and widths up, so that rows move as 16-byte vectors
    const int hxa = axis_tap(S.s_hi_x, oxa, S.hw, W).i0 & ~3, hxb = axis_tap(S.s_hi_x, oxb, S.hw, W).i1 | 3;
    if (tid == 0) s_hxa[s] = hxa;
    const int nhy = hyb - hya + 1, nhx = min(hxb, S.hw - 1) - hxa + 1;
    int t0, t1; float tw0, tw1;
    tap_x2(hya, S.lh, t0, t1, tw0, tw1);
    const int lya = t0;
    tap_x2(hyb, S.lh, t0, t1, tw0, tw1);
    const int nly = t1 - lya + 1;
    tap_x2(hxa, S.lw, t0, t1, tw0, tw1);
    const int lxa = t0 & ~3;
    tap_x2(min(hxb, S.hw - 1), S.lw, t0, t1, tw0, tw1);
    const int nlx = min(t1 | 3, S.lw - 1) - lxa + 1;
    stage_window<NW>(sLo, g.lc, S.lo + (size_t)b * S.lo_sb + (size_t)k * S.lo_sc,
                     S.lo_f ? S.lo_f + (size_t)b * S.lof_sb + (size_t)kf * S.lof_sc : nullptr, lya, lxa, nly, nlx >> 2,
                     S.lw, warp, lane);
    stage_window<NW>(sS, hc, S.hi + (size_t)b * S.hi_sb + (size_t)k * S.hi_sc,
                     S.hi_f ? S.hi_f + (size_t)b * S.hif_sb + (size_t)kf * S.hif_sc : nullptr, hya, hxa, nhy, nhx >> 2,
                     S.hw, warp, lane);
    for (int r = tid; r < nhy; r += nthr) {       // S_s row -> lo tile rows
      int i0, i1; float w0, w1;
      tap_x2(hya + r, S.lh, i0, i1, w0, w1);
      p2_i0[r] = i0 - lya; p2_i1[r] = i1 - lya; p2_w0[r] = w0; p2_w1[r] = w1;
    }
    for (int t = tid; t < NROWS; t += nthr) {     // output row -> S_s rows
      const int y = ys + t;
      if (y >= 0 && y < H) {
        const Tap tp = axis_tap(S.s_hi_y, y, S.hh, H);
        rt_i0[s * NROWS + t] = tp.i0 - hya; rt_i1[s * NROWS + t] = tp.i1 - hya;
        rt_w0[s * NROWS + t] = tp.w0; rt_w1[s * NROWS + t] = tp.w1;
      } else {
        rt_i0[s * NROWS + t] = -1;
      }
    }
    __syncthreads();
    for (int c = tid; c < nhx; c += nthr) {
      int c0, c1; float wx0, wx1;
      tap_x2(hxa + c, S.lw, c0, c1, wx0, wx1);
      c0 -= lxa; c1 -= lxa;
      for (int r = 0; r < nhy; ++r) {
        const float* r0 = sLo + p2_i0[r] * g.lc;
        const float* r1 = sLo + p2_i1[r] * g.lc;
        const float up = lerp2(wx0, wx1, p2_w0[r], p2_w1[r], r0[c0], r0[c1], r1[c0], r1[c1]);
        sS[r * hc + c] = __fmul_rn(__fadd_rn(up, sS[r * hc + c]), 0.5f);
      }
    }
    __syncthreads();    // sLo / p2 tables are re-used by the next scale; S_s is complete
  }

  int hxa_s[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) hxa_s[s] = s_hxa[s];
  // per-lane column taps of every scale, kept in registers: tile-relative index of the first column's
  // left tap, then per column a 2-bit offset to its own left tap (hw <= W, so 4 adjacent output columns
  // span at most 4 source columns) and 1 bit "right tap = left tap + 1"
  int cbase[NS];
  unsigned cpack[NS];
  float cw1[NS][4];   // w0 is 1 - w1 exactly as axis_tap computes it (recomputed on use)
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    cpack[s] = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const Tap tp = axis_tap(a.sc[s].s_hi_x, min(X0 + c, W - 1), a.sc[s].hw, W);
      if (c == 0) cbase[s] = tp.i0 - hxa_s[s];
      cpack[s] |= (unsigned)(tp.i0 - hxa_s[s] - cbase[s]) << (2 * c);
      cpack[s] |= (unsigned)(tp.i1 - tp.i0) << (8 + c);
      cw1[s][c] = tp.w1;
    }
  }

  // ---------------- phase 3a: values of the 4 halo columns of this warp's strip ------------------------
  float* edge = sEdge + warp * NROWS * 4;
#pragma unroll
  for (int u = 0; u < (NROWS * 4 + 31) / 32; ++u) {
    const int i = lane + 32 * u;
    if (i < NROWS * 4) {
      const int t = i >> 2, e = i & 3;
      const int xe = (e < 2) ? xw - 2 + e : xw + 126 + e;
      const int y = ys + t;
      float v = NINF;
      if (xe >= 0 && xe < W && y >= 0 && y < H) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
          const Tap tx = axis_tap(a.sc[s].s_hi_x, xe, a.sc[s].hw, W);
          const float* sS = smem + g.off_s[s];
          const float* r0 = sS + rt_i0[s * NROWS + t] * g.hc[s];
          const float* r1 = sS + rt_i1[s * NROWS + t] * g.hc[s];
          const int c0 = tx.i0 - hxa_s[s], c1 = tx.i1 - hxa_s[s];
          const float vs = lerp2(tx.w0, tx.w1, rt_w0[s * NROWS + t], rt_w1[s * NROWS + t], r0[c0], r0[c1], r1[c0], r1[c1]);
          v = (s == 0) ? vs : __fadd_rn(v, vs);
        }
        if (NS > 1) v = __fdiv_rn(v, (float)NS);
      }
      edge[i] = v;
    }
  }
  __syncwarp();

  // ---------------- phase 3b: walk the rows ---------------------------------------------------------------
  {
    float hX[NS][4], hY[NS][4];   // two cached horizontally interpolated S_s rows per scale
    int xi[NS], yi[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) { xi[s] = yi[s] = -1000; }
    auto hpass = [&](int s, int r, float (&h)[4]) {
      const float* sp = smem + g.off_s[s] + r * g.hc[s];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int i0 = cbase[s] + (int)((cpack[s] >> (2 * c)) & 3u);
        const float v0 = sp[i0], v1 = sp[i0 + (int)((cpack[s] >> (8 + c)) & 1u)];
        h[c] = fmaf(1.f - cw1[s][c], v0, __fmul_rn(cw1[s][c], v1));
      }
    };
    float r0[4], r1[4], r2[4], r3[4], vq0[4], vq1[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) { r0[c] = r1[c] = r2[c] = r3[c] = NINF; vq0[c] = vq1[c] = NINF; }
    float* const hm_plane = a.agg_hm + ((size_t)b * a.K + k) * H * W + X0;
    const size_t wbase = ((size_t)b * a.K + k) * H * a.wpr + (X0 >> 5);
    const bool word_writer = active && (lane & 7) == 0;
    const int shl = 4 * (lane & 7);

    auto row_value = [&](int t, float (&v)[4]) {
      const int y = ys + t;
      if (y < 0 || y >= H) {
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] = NINF;
        return;
      }
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const int i0 = rt_i0[s * NROWS + t], i1 = rt_i1[s * NROWS + t];
        const float w0 = rt_w0[s * NROWS + t], w1 = rt_w1[s * NROWS + t];
        float A[4], Bv[4];
        if (i0 == xi[s]) {
#pragma unroll
          for (int c = 0; c < 4; ++c) A[c] = hX[s][c];
        } else if (i0 == yi[s]) {
#pragma unroll
          for (int c = 0; c < 4; ++c) A[c] = hY[s][c];
        } else {
          hpass(s, i0, A);
        }
        if (i1 == i0) {
#pragma unroll
          for (int c = 0; c < 4; ++c) Bv[c] = A[c];
        } else if (i1 == yi[s]) {
#pragma unroll
          for (int c = 0; c < 4; ++c) Bv[c] = hY[s][c];
        } else if (i1 == xi[s]) {
#pragma unroll
          for (int c = 0; c < 4; ++c) Bv[c] = hX[s][c];
        } else {
          hpass(s, i1, Bv);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          hX[s][c] = A[c];
          hY[s][c] = Bv[c];
          const float vs = fmaf(w0, A[c], __fmul_rn(w1, Bv[c]));
          v[c] = (s == 0) ? vs : __fadd_rn(v[c], vs);
        }
        xi[s] = i0;
        yi[s] = i1;
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (NS > 1) v[c] = __fdiv_rn(v[c], (float)NS);
        if (!active) v[c] = NINF;
      }
    };

    auto process_row = [&](int t, const float (&v)[4]) {
      const int y = ys + t;
      if (t >= 2 && t < RB + 2 && y < H && active)
        *reinterpret_cast<float4*>(hm_plane + (size_t)y * W) = make_float4(v[0], v[1], v[2], v[3]);
      float l2 = __shfl_up_sync(kFull, v[2], 1), l3 = __shfl_up_sync(kFull, v[3], 1);
      float q0 = __shfl_down_sync(kFull, v[0], 1), q1 = __shfl_down_sync(kFull, v[1], 1);
      if (lane == 0) { const float2 e = *reinterpret_cast<const float2*>(edge + t * 4); l2 = e.x; l3 = e.y; }
      if (lane == 31) { const float2 e = *reinterpret_cast<const float2*>(edge + t * 4 + 2); q0 = e.x; q1 = e.y; }
      const float pb = fmaxf(v[0], v[1]), pc = fmaxf(v[2], v[3]);
      float rm[4];
      rm[0] = max3(fmaxf(l2, l3), pb, v[2]);
      rm[1] = max3(l3, pb, pc);
      rm[2] = max3(pb, pc, q0);
      rm[3] = max3(v[1], pc, fmaxf(q0, q1));
      const int yc = y - 2;
      if (t >= 4 && yc < H) {
        const float m0 = max3(max3(r0[0], r1[0], r2[0]), r3[0], rm[0]);
        const float m1 = max3(max3(r0[1], r1[1], r2[1]), r3[1], rm[1]);
        const float m2 = max3(max3(r0[2], r1[2], r2[2]), r3[2], rm[2]);
        const float m3 = max3(max3(r0[3], r1[3], r2[3]), r3[3], rm[3]);
        const bool k0 = (m0 == vq0[0]), k1 = (m1 == vq0[1]), k2 = (m2 == vq0[2]), k3 = (m3 == vq0[3]);
        float wm4 = fmaxf(fmaxf(k0 ? vq0[0] : 0.f, k1 ? vq0[1] : 0.f), fmaxf(k2 ? vq0[2] : 0.f, k3 ? vq0[3] : 0.f));
        float hm4 = fmaxf(fmaxf(vq0[0], vq0[1]), fmaxf(vq0[2], vq0[3]));
        unsigned bits = ((k0 ? 1u : 0u) | (k1 ? 2u : 0u) | (k2 ? 4u : 0u) | (k3 ? 8u : 0u)) << shl;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
          bits |= __shfl_xor_sync(kFull, bits, o);
          hm4 = fmaxf(hm4, __shfl_xor_sync(kFull, hm4, o));
          wm4 = fmaxf(wm4, __shfl_xor_sync(kFull, wm4, o));
        }
        if (word_writer) {
          const size_t w = wbase + (size_t)yc * a.wpr;
          a.mask[w] = bits;
          a.wmax[w] = wm4;
          a.hmax[w] = hm4;
        }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        r0[c] = r1[c]; r1[c] = r2[c]; r2[c] = r3[c]; r3[c] = rm[c];
        vq0[c] = vq1[c]; vq1[c] = v[c];
      }
    };

    // (not unrolled: with NS scales the body is large, and one copy keeps it in the instruction cache)
#pragma unroll 1
    for (int t = 0; t < NROWS; ++t) {
      float v[4];
      row_value(t, v);
      process_row(t, v);
    }
  }
}

}  // namespace ms

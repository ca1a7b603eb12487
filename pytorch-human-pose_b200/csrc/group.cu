// Stage (d): associative-embedding grouping -- MPPEHeatmapParser.match_by_tag
// (/root/reference/src/keypoints/grouping.py:85-145), py_max_match (:55-59, munkres 1.1.4
// Munkres.compute) and the empty-scene fallback of parse (:262-269).
//
// One warp per image (the 17 joint steps are strictly sequential and every step works on at
// most 32 detections x 32 persons, so a warp is the natural unit; a batch fills the GPU with
// one warp per image and the kernel is latency-bound by design -- bench.py reports its
// occupancy and microseconds per image instead of a bandwidth fraction).
//
// Exactness notes:
//   * det_thr and tag_thr are compared in float64; costs are float64:
//       cost = rint(sqrt(sum_e (tag_e - mean_e)^2)) * 100 - score     (no contraction anywhere)
//   * the mean tag of a person is numpy's float32 np.mean over its tag list (pairwise-8 for
//     E = 1, sequential for E = 2);
//   * persons are dict entries keyed by the float32 value of tag[0]: an equal key overwrites that
//     person's joint and resets its tag list; only the first M persons are match candidates or
//     outputs, later ones are counted but not stored;
//   * the Hungarian solver replays munkres 1.1.4 step by step: lane r owns matrix row r and a
//     32-bit mask of its zero entries, so "find the last uncovered zero in cyclic order" is bit
//     arithmetic; the (C + m) - m update order of step 6 is kept.
#include "common.cuh"

namespace hpd {

namespace {

constexpr int NP = HPD_MAX_PEOPLE;      // 32
constexpr int CS = NP + 1;              // padded row stride of the float64 matrices

struct GroupSmem {
  double C[NP][CS];                     // cost matrix, edited in place by the solver
  double D[NP][CS];                     // saved distances (diff_saved)
  float taglist[NP][HPD_MAX_KPTS][HPD_MAX_EMB];
};

__device__ __forceinline__ double warp_min_double(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double w = __shfl_xor_sync(kFull, v, o);
    v = (w < v) ? w : v;
  }
  return v;
}

// munkres 1.1.4 Munkres.compute on the n x n matrix in shared memory; returns the starred
// column of this lane's row (valid for lane < n).
__device__ int munkres_warp(double (*C)[CS], int n, int lane) {
  const unsigned nmask = (n >= 32) ? kFull : ((1u << n) - 1u);
  unsigned zmask = 0;
  // step 1: subtract the row minimum
  if (lane < n) {
    double m = C[lane][0];
    for (int j = 1; j < n; ++j) {
      const double v = C[lane][j];
      if (v < m) m = v;
    }
    for (int j = 0; j < n; ++j) {
      const double v = __dsub_rn(C[lane][j], m);
      C[lane][j] = v;
      if (v == 0.0) zmask |= 1u << j;
    }
  }
  // step 2: rows ascending, star the first zero in an uncovered column
  int star = -1;
  unsigned colcov = 0, rowcov = 0;
  for (int i = 0; i < n; ++i) {
    const unsigned c = __shfl_sync(kFull, zmask, i) & ~colcov;
    if (c) {
      const int j = __ffs(c) - 1;
      if (lane == i) star = j;
      colcov |= 1u << j;
    }
  }
  colcov = 0;
  int prime = -1;
  while (true) {
    // step 3: cover starred columns
    colcov |= __reduce_or_sync(kFull, (lane < n && star >= 0) ? (1u << star) : 0u);
    if (__popc(colcov) >= n) break;
    int z0r = 0, z0c = 0;
    while (true) {
      // step 4
      int row = 0, col = 0;
      bool augment = false;
      while (true) {
        const unsigned cand = (lane < n && !((rowcov >> lane) & 1u)) ? (zmask & ~colcov & nmask) : 0u;
        const unsigned rows_with = __ballot_sync(kFull, cand != 0u);
        if (!rows_with) break;
        const unsigned hi = rows_with >> row;
        const int r = hi ? (row + __ffs(hi) - 1) : (__ffs(rows_with) - 1);
        const unsigned c = __shfl_sync(kFull, cand, r);
        const unsigned low = c & ((1u << col) - 1u);
        const int cc = low ? (31 - __clz(low)) : (31 - __clz(c));
        if (lane == r) prime = cc;
        const int sc = __shfl_sync(kFull, star, r);
        row = r;
        if (sc >= 0) {
          col = sc;
          rowcov |= 1u << r;
          colcov &= ~(1u << sc);
        } else {
          z0r = r;
          z0c = cc;
          augment = true;
          break;
        }
      }
      if (augment) break;
      // step 6
      double m = 9.223372036854775807e18;
      if (lane < n && !((rowcov >> lane) & 1u)) {
        for (int j = 0; j < n; ++j)
          if (!((colcov >> j) & 1u)) {
            const double v = C[lane][j];
            if (m > v) m = v;
          }
      }
      m = warp_min_double(m);
      if (lane < n) {
        const bool rc = (rowcov >> lane) & 1u;
        zmask = 0;
        for (int j = 0; j < n; ++j) {
          double v = C[lane][j];
          if (rc) v = __dadd_rn(v, m);
          if (!((colcov >> j) & 1u)) v = __dsub_rn(v, m);
          C[lane][j] = v;
          if (v == 0.0) zmask |= 1u << j;
        }
      }
    }
    // step 5: augment along the alternating star/prime path from Z0
    {
      int r = z0r, c = z0c;
      while (true) {
        const unsigned b = __ballot_sync(kFull, lane < n && star == c);
        if (lane == r) star = c;
        if (!b) break;
        const int r2 = __ffs(b) - 1;
        c = __shfl_sync(kFull, prime, r2);
        r = r2;
      }
      rowcov = colcov = 0;
      prime = -1;
    }
  }
  return star;
}

__global__ void __launch_bounds__(32) group_kernel(const float* __restrict__ scores_k,
                                                   const int32_t* __restrict__ coords_k,
                                                   const float* __restrict__ tags_k, int K, int M, int E,
                                                   double det_thr, double tag_thr, const HpdParams prm,
                                                   float* __restrict__ poses, int32_t* __restrict__ n_person,
                                                   int32_t* __restrict__ flags) {
  __shared__ GroupSmem sm;
  const int lane = threadIdx.x;
  const int b = blockIdx.x;
  const int D = 3 + E;
  const float* sc_b = scores_k + (size_t)b * K * M;
  const int32_t* co_b = coords_k + (size_t)b * K * M * 2;
  const float* tg_b = tags_k + (size_t)b * K * M * E;
  float* out = poses + (size_t)b * M * K * D;
  for (int i = lane; i < M * K * D; i += 32) out[i] = 0.f;
  __syncwarp();

  int P = 0;          // persons stored (<= M), uniform
  int Ptotal = 0;     // persons created, uniform
  float key = 0.f;    // lane p: dict key of person p
  int ntag = 0;       // lane p: length of person p's tag list

  for (int it = 0; it < K; ++it) {
    const int k = prm.joints_order[it];
    // lane r looks at rank r of joint k
    const float my_score_r = (lane < M) ? sc_b[k * M + lane] : 0.f;
    const unsigned rowmask = __ballot_sync(kFull, lane < M && (double)my_score_r > det_thr);
    const int nr = __popc(rowmask);
    if (nr == 0) continue;
    // lane a < nr owns the a-th detection above threshold
    int my_r = -1;
    float a_score = 0.f, a_x = 0.f, a_y = 0.f, a_t0 = 0.f, a_t1 = 0.f;
    if (lane < nr) {
      my_r = __fns(rowmask, 0, lane + 1);
      a_score = sc_b[k * M + my_r];
      a_x = (float)co_b[(k * M + my_r) * 2 + 0];
      a_y = (float)co_b[(k * M + my_r) * 2 + 1];
      a_t0 = tg_b[(size_t)(k * M + my_r) * E];
      a_t1 = (E > 1) ? tg_b[(size_t)(k * M + my_r) * E + 1] : 0.f;
    }

    // writes detection a into person p (lane a does the stores)
    auto put_joint = [&](int a, int p) {
      if (lane == a) {
        float* d = out + ((size_t)p * K + k) * D;
        d[0] = a_x; d[1] = a_y; d[2] = a_score; d[3] = a_t0;
        if (E > 1) d[4] = a_t1;
      }
    };
    // dict.setdefault(key)[idx] = joint ; tag_dict[key] = [tag]   (grouping.py:109-111,141-143)
    auto new_or_collide = [&](int a) {
      const float t0 = __shfl_sync(kFull, a_t0, a);
      const unsigned hit = __ballot_sync(kFull, lane < P && key == t0);
      int p;
      if (hit) {
        p = __ffs(hit) - 1;
      } else {
        ++Ptotal;
        if (P < M) {
          p = P++;
          if (lane == p) key = t0;
        } else {
          p = -1;   // person beyond the first M: never read again
        }
      }
      if (p >= 0) {
        put_joint(a, p);
        if (lane == a) {
          sm.taglist[p][0][0] = a_t0;
          sm.taglist[p][0][1] = a_t1;
        }
        if (lane == p) ntag = 1;
      }
      __syncwarp();
    };

    if (it == 0 || Ptotal == 0) {
      for (int a = 0; a < nr; ++a) new_or_collide(a);
      continue;
    }
    const int G = P;
    const int n = max(G, nr);
    // mean tag per existing person (grouping.py:114)
    float mean0 = 0.f, mean1 = 0.f;
    if (lane < G) {
      float mv[HPD_MAX_EMB];
      np_mean_vectors(&sm.taglist[lane][0][0], ntag, E, HPD_MAX_EMB, mv);
      mean0 = mv[0];
      mean1 = (E > 1) ? mv[1] : 0.f;
    }
    // cost matrix (grouping.py:116-128), one (detection, person) pair per lane and step
    {
      const int npairs = nr * G;
      const unsigned inv = (65536u + G - 1) / G;         // i / G == (i * inv) >> 16 for i < 2048
      for (int i = lane; i < ((npairs + 31) & ~31); i += 32) {
        const int a = min((int)(((unsigned)i * inv) >> 16), nr - 1);
        const int p = (i < npairs) ? i - a * G : 0;
        const float t0 = __shfl_sync(kFull, a_t0, a), t1 = __shfl_sync(kFull, a_t1, a);
        const float sc_a = __shfl_sync(kFull, a_score, a);
        const float m0 = __shfl_sync(kFull, mean0, p), m1 = __shfl_sync(kFull, mean1, p);
        if (i < npairs) {
          const double d0 = __dsub_rn((double)t0, (double)m0);
          double s = __dmul_rn(d0, d0);
          if (E > 1) {
            const double d1 = __dsub_rn((double)t1, (double)m1);
            s = __dadd_rn(s, __dmul_rn(d1, d1));
          }
          const double dn = __dsqrt_rn(s);
          sm.D[a][p] = dn;
          sm.C[a][p] = __dsub_rn(__dmul_rn(rint(dn), 100.0), (double)sc_a);
        }
      }
      // more detections than persons: 1e10 columns (grouping.py:126-128); fewer: munkres pads zero rows
      if (lane < nr) for (int p = G; p < n; ++p) sm.C[lane][p] = 1e10;
      if (lane < n) for (int a = nr; a < n; ++a) sm.C[a][lane] = 0.0;
    }
    __syncwarp();
    const int star = munkres_warp(sm.C, n, lane);
    __syncwarp();
    // grouping.py:131-143
    for (int a = 0; a < nr; ++a) {
      const int c = __shfl_sync(kFull, star, a);
      int ok = 0;
      if (lane == a) ok = (c < G && sm.D[a][c] < tag_thr) ? 1 : 0;
      ok = __shfl_sync(kFull, ok, a);
      if (ok) {
        put_joint(a, c);
        const int nt = __shfl_sync(kFull, ntag, c);
        if (lane == a) {
          sm.taglist[c][nt][0] = a_t0;
          sm.taglist[c][nt][1] = a_t1;
        }
        if (lane == c) ntag = nt + 1;
        __syncwarp();
      } else {
        new_or_collide(a);
      }
    }
  }

  int fl = 0;
  if (Ptotal == 0) {
    // grouping.py:262-269: one pseudo-person from the best candidate of every joint, score := 0.01
    fl = 1;
    P = 1;
    for (int k = lane; k < K; k += 32) {
      float* d = out + (size_t)k * D;
      d[0] = (float)co_b[(k * M) * 2 + 0];
      d[1] = (float)co_b[(k * M) * 2 + 1];
      d[2] = 0.01f;
      for (int e = 0; e < E; ++e) {
        const float t = tg_b[(size_t)(k * M) * E + e];
        d[3 + e] = (t != t) ? 0.f : t;
      }
    }
  }
  if (lane == 0) {
    n_person[b] = P;
    flags[b] = fl;
  }
}

}  // namespace

int launch_group(const HpdParams* p, const HpdBuffers* buf, cudaStream_t st) {
  if (!buf->scores_k || !buf->coords_k || !buf->tags_k || !buf->poses || !buf->n_person || !buf->flags) {
    set_error("hpd_group: scores_k, coords_k, tags_k, poses, n_person, flags are required");
    return HPD_EINVAL;
  }
  group_kernel<<<p->batch, 32, 0, st>>>(buf->scores_k, buf->coords_k, buf->tags_k, p->num_kpts, p->max_people, p->emb,
                                        p->det_thr, p->tag_thr, *p, buf->poses, buf->n_person, buf->flags);
  count_launch();
  return check_launch("group_kernel");
}

}  // namespace hpd

// Stage (c): per-joint top-k (k = max_num_people) of the NMS'd heatmap with the tie order of
// the reference's CPU run (grouping.py:147-170).
//
// torch's CPU topk for k*64 <= n is std::partial_sort over (value, index) pairs with a
// value-only comparator (aten/src/ATen/native/TopKImpl.h), i.e. libstdc++'s __heap_select +
// __sort_heap: a k-entry heap whose top is the smallest kept value, fed by the stream in index
// order; an element enters only if it is STRICTLY greater than the current top.  Among equal
// values the final order is therefore a function of the heap's history, not of the indices
// (SURVEY.md App. A.4).  This kernel replays that history exactly:
//   * one warp per (image, joint) row; the heap lives in shared memory and is edited by lane 0
//     with the libstdc++ __adjust_heap / __push_heap control flow;
//   * the heap top never decreases, so the stream is pre-filtered 32 words (1024 pixels) at a
//     time with the per-word maximum of the NMS'd values written by the aggregation kernel;
//     only words that can still contain an entering element are expanded (one coalesced 128-byte
//     line of the heatmap + the survivor mask word), and entering elements are replayed in index
//     order with the live top.
// The NMS'd value of a suppressed pixel is x*0 (sign of x): zeros do enter while the top is
// negative and their indices are part of the bit-exact contract.
#include "common.cuh"

namespace hpd {

namespace {

constexpr int kTopkWarps = 4;
constexpr int kLogCap = 768;   // entering elements logged per row (typical: 100-300)

struct HeapRef {
  float* v;
  int* i;
};

// comp(a, b) of the reference's lambda: a.value > b.value (finite inputs)
__device__ __forceinline__ bool comp_gt(float a, float b) { return a > b; }

// libstdc++ std::__push_heap
__device__ __forceinline__ void push_heap(HeapRef h, int hole, int top, float val, int idx) {
  int parent = (hole - 1) / 2;
  while (hole > top && comp_gt(h.v[parent], val)) {
    h.v[hole] = h.v[parent];
    h.i[hole] = h.i[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  h.v[hole] = val;
  h.i[hole] = idx;
}

// libstdc++ std::__adjust_heap
__device__ __forceinline__ void adjust_heap(HeapRef h, int hole, int len, float val, int idx) {
  const int top = hole;
  int child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (comp_gt(h.v[child], h.v[child - 1])) --child;
    h.v[hole] = h.v[child];
    h.i[hole] = h.i[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    h.v[hole] = h.v[child - 1];
    h.i[hole] = h.i[child - 1];
    hole = child - 1;
  }
  push_heap(h, hole, top, val, idx);
}

__device__ __forceinline__ void make_heap(HeapRef h, int len) {
  if (len < 2) return;
  int parent = (len - 2) / 2;
  while (true) {
    const float v = h.v[parent];
    const int i = h.i[parent];
    adjust_heap(h, parent, len, v, i);
    if (parent == 0) return;
    --parent;
  }
}

__device__ __forceinline__ void sort_heap(HeapRef h, int len) {
  int last = len;
  while (last > 1) {
    --last;
    const float v = h.v[last];
    const int i = h.i[last];
    h.v[last] = h.v[0];
    h.i[last] = h.i[0];
    adjust_heap(h, 0, last, v, i);
  }
}

// ---- exact sink: the libstdc++ heap in shared memory, edited by lane 0 -----------------------------
struct HeapSink {
  HeapRef h;
  int M, lane;
  float top;
  __device__ __forceinline__ void init(float nv) {   // lane < M holds element `lane` of the row
    if (lane < M) { h.v[lane] = nv; h.i[lane] = lane; }
    __syncwarp();
    if (lane == 0) make_heap(h, M);
    __syncwarp();
    top = h.v[0];
  }
  __device__ __forceinline__ void insert(float cv, int idx) {   // std::__heap_select: __pop_heap(first, middle, i)
    if (lane == 0) adjust_heap(h, 0, M, cv, idx);
    __syncwarp();
    top = h.v[0];
  }
};

// ---- fast sink: the M kept elements sorted by value across the lanes of the warp ---------------------
// The heap's VALUE dynamics do not depend on its arrangement (an element enters iff it is strictly
// greater than the smallest kept value, and a smallest-valued element leaves), so a sorted register
// array evolves through the same multisets.  Which of several EQUAL elements leaves, and the final
// order of equal elements, is where the heap arrangement matters; `ambiguous()` detects every such
// case (two equal values among the final M, or an evicted value equal to the final minimum -- the
// +-0 tail of a channel with fewer than M positive peaks always lands here) and the row is then
// replayed with the exact sink.  Otherwise the result (distinct values, descending) is what
// sort_heap produces from any arrangement.
struct SortedSink {
  int M, lane;
  float sv;       // lane i < M: i-th largest kept value (lanes >= M: -inf)
  int si;
  float top;      // smallest kept value
  float evicted;  // value of the last evicted element (evictions are non-decreasing)
  bool any_evicted;
  // log of the entering elements, in order: the exact sink goes through the same VALUE history, so a
  // replay only has to feed it this list instead of scanning the row again
  float* log_v;
  int* log_i;
  int n_log, log_cap;
  float rej_eq;   // merge only: largest value that arrived equal to the then-smallest kept value
  __device__ __forceinline__ void init_empty() {   // M slots of -inf: the first M arrivals fill them
    n_log = 0;
    sv = -INFINITY;
    si = -1;
    top = -INFINITY;
    evicted = -INFINITY;
    rej_eq = -INFINITY;
    any_evicted = false;
  }
  __device__ __forceinline__ void offer(float cv, int idx) {   // merge step: any value may be offered
    if (cv > top) insert(cv, idx);
    else if (cv == top) rej_eq = cv;
  }
  __device__ __forceinline__ void init(float nv) {
    n_log = 0;
    rej_eq = -INFINITY;
    sv = lane < M ? nv : -INFINITY;
    si = lane;
    // bitonic sort, descending by value
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
      for (int j = k >> 1; j > 0; j >>= 1) {
        const float ov = __shfl_xor_sync(kFull, sv, j);
        const int oi = __shfl_xor_sync(kFull, si, j);
        const bool up = ((lane & k) == 0);          // this k-block sorts descending
        const bool lower = ((lane & j) == 0);       // this lane keeps the larger (if up) element
        const bool take = (lower == up) ? (ov > sv) : (ov < sv);
        if (take) { sv = ov; si = oi; }
      }
    }
    top = __shfl_sync(kFull, sv, M - 1);
    evicted = -INFINITY;
    any_evicted = false;
  }
  __device__ __forceinline__ void insert(float cv, int idx) {   // requires cv > top
    const int pos = __popc(__ballot_sync(kFull, lane < M && sv >= cv));
    const float uv = __shfl_up_sync(kFull, sv, 1);
    const int ui = __shfl_up_sync(kFull, si, 1);
    evicted = top;
    any_evicted = true;
    if (lane == 0 && n_log < log_cap) { log_v[n_log] = cv; log_i[n_log] = idx; }
    ++n_log;
    if (lane < M) {
      if (lane > pos) { sv = uv; si = ui; }
      else if (lane == pos) { sv = cv; si = idx; }
    }
    top = __shfl_sync(kFull, sv, M - 1);
  }
  __device__ __forceinline__ bool ambiguous() const {
    const float nxt = __shfl_down_sync(kFull, sv, 1);
    const bool tie = __ballot_sync(kFull, lane < M - 1 && sv == nxt) != 0u;
    return tie || (any_evicted && evicted == top) || rej_eq == top;
  }
};

// Stream one row through a sink in index order.  The per-word maxima are scanned 128 words at a time
// (the next 128 are prefetched while this group is processed); words that can still hold an
// entering element are expanded up to 8 at a time so that their heatmap lines and mask words are
// fetched together instead of one DRAM round trip each.
template <typename Sink>
__device__ __forceinline__ void scan_range(Sink& sink, const float* __restrict__ hm, const uint32_t* __restrict__ mk,
                                           const float* __restrict__ wm, int W, int wpr, int w_begin, int w_end,
                                           int first_idx, int lane) {
  // words [w_begin, w_end) of the row; pixels with flat index < first_idx are not offered (they are the
  // initial heap of the sequential algorithm)
  const int nwords = w_end;
  constexpr int kGroup = 4, kSlots = 8, kAhead = 4;
  // word maxima are loaded kAhead groups (of 128 words) ahead of their use
  float ring[kAhead][kGroup];
#pragma unroll
  for (int a = 0; a < kAhead; ++a)
#pragma unroll
    for (int u = 0; u < kGroup; ++u) {
      const int wd = w_begin + 32 * (a * kGroup + u) + lane;
      ring[a][u] = wd < nwords ? wm[wd] : -INFINITY;
    }
  for (int base = w_begin; base < nwords; base += 32 * kGroup) {
    float cur[kGroup];
#pragma unroll
    for (int u = 0; u < kGroup; ++u) {
      cur[u] = ring[0][u];
#pragma unroll
      for (int a = 0; a + 1 < kAhead; ++a) ring[a][u] = ring[a + 1][u];
      const int wd = base + 32 * (kAhead * kGroup + u) + lane;
      ring[kAhead - 1][u] = wd < nwords ? wm[wd] : -INFINITY;
    }
#pragma unroll
    for (int u = 0; u < kGroup; ++u) {
      const int cbase = base + 32 * u;
      uint32_t pass = __ballot_sync(kFull, cur[u] > sink.top);
      while (pass) {
        float hv[kSlots];
        uint32_t mw[kSlots];
        int wl[kSlots];
#pragma unroll
        for (int s = 0; s < kSlots; ++s) {
          wl[s] = -1;
          hv[s] = 0.f;
          mw[s] = 0u;
          if (pass) {
            const int l = __ffs(pass) - 1;
            pass &= pass - 1;
            wl[s] = l;
            const int w2 = cbase + l;
            const int y = w2 / wpr, x = (w2 - y * wpr) * 32 + lane;
            const int idx = y * W + x;
            if (x < W && idx >= first_idx) hv[s] = hm[idx];
            mw[s] = mk[w2];
          }
        }
#pragma unroll
        for (int s = 0; s < kSlots; ++s) {
          if (wl[s] < 0) break;
          if (!(__shfl_sync(kFull, cur[u], wl[s]) > sink.top)) continue;   // the top may have risen meanwhile
          const int w2 = cbase + wl[s];
          const int y = w2 / wpr, x0 = (w2 - y * wpr) * 32;
          const bool valid = (x0 + lane < W) && (y * W + x0 + lane >= first_idx);
          const bool keep = (mw[s] >> lane) & 1u;
          const float nv = keep ? hv[s] : __fmul_rn(hv[s], 0.0f);
          uint32_t cand = __ballot_sync(kFull, valid && nv > sink.top);
          while (cand) {
            const int j = __ffs(cand) - 1;
            cand &= cand - 1;
            const float cv = __shfl_sync(kFull, nv, j);
            if (cv > sink.top) sink.insert(cv, y * W + x0 + j);
          }
        }
      }
    }
  }
}

// NMS'd value of flat element `lane` (the sequential algorithm's initial heap is elements 0..M-1)
__device__ __forceinline__ float first_element(const float* __restrict__ hm, const uint32_t* __restrict__ mk, int W,
                                               int wpr, int M, int lane) {
  float nv = 0.f;
  if (lane < M) {
    const int y = lane / W, x = lane % W;
    const float v = hm[lane];
    const bool keep = (mk[(size_t)y * wpr + (x >> 5)] >> (x & 31)) & 1u;
    nv = keep ? v : __fmul_rn(v, 0.0f);
  }
  return nv;
}

template <typename Sink>
__device__ __forceinline__ void scan_row(Sink& sink, const float* __restrict__ hm, const uint32_t* __restrict__ mk,
                                         const float* __restrict__ wm, int H, int W, int wpr, int M, int lane) {
  sink.init(first_element(hm, mk, W, wpr, M, lane));
  scan_range(sink, hm, mk, wm, W, wpr, 0, H * wpr, M, lane);
}

__global__ void __launch_bounds__(kTopkWarps * 32) topk_kernel(const float* __restrict__ agg_hm,
                                                              const float* __restrict__ agg_tags,
                                                              const uint32_t* __restrict__ mask,
                                                              const float* __restrict__ wmax, int rows, int H, int W,
                                                              int wpr, int E, int M, int force_exact,
                                                              float* __restrict__ scores_k,
                                                              int32_t* __restrict__ idx_k,
                                                              int32_t* __restrict__ coords_k,
                                                              float* __restrict__ tags_k) {
  __shared__ float s_v[kTopkWarps][32];
  __shared__ int s_i[kTopkWarps][32];
  __shared__ float s_logv[kTopkWarps][kLogCap];
  __shared__ int s_logi[kTopkWarps][kLogCap];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * kTopkWarps + warp;
  if (row >= rows) return;
  const float* hm = agg_hm + (size_t)row * H * W;
  const uint32_t* mk = mask + (size_t)row * H * wpr;
  const float* wm = wmax + (size_t)row * H * wpr;

  float out_v = 0.f;
  int out_i = 0;
  bool done = false;
  int n_logged = -1;   // >= 0: the fast pass logged every entering element of this row
  if (!force_exact) {
    SortedSink fast;
    fast.M = M; fast.lane = lane;
    fast.log_v = s_logv[warp]; fast.log_i = s_logi[warp]; fast.log_cap = kLogCap;
    scan_row(fast, hm, mk, wm, H, W, wpr, M, lane);
    if (!fast.ambiguous()) { out_v = fast.sv; out_i = fast.si; done = true; }
    else n_logged = fast.n_log;
  }
  if (!done) {   // ties in play: replay the row with the exact libstdc++ heap
    HeapSink exact;
    exact.h = HeapRef{s_v[warp], s_i[warp]};
    exact.M = M; exact.lane = lane;
    if (n_logged >= 0 && n_logged <= kLogCap) {
      // same entering elements, same order: build the heap from the first M elements and feed the log
      exact.init(first_element(hm, mk, W, wpr, M, lane));
      __syncwarp();
      if (lane == 0)
        for (int i = 0; i < n_logged; ++i) adjust_heap(exact.h, 0, M, s_logv[warp][i], s_logi[warp][i]);
    } else {
      scan_row(exact, hm, mk, wm, H, W, wpr, M, lane);
    }
    __syncwarp();
    if (lane == 0) sort_heap(exact.h, M);
    __syncwarp();
    if (lane < M) { out_v = exact.h.v[lane]; out_i = exact.h.i[lane]; }
  }
  if (lane < M) {
    const size_t o = (size_t)row * M + lane;
    scores_k[o] = out_v;
    idx_k[o] = out_i;
    coords_k[o * 2 + 0] = out_i % W;
    coords_k[o * 2 + 1] = out_i / W;
    for (int e = 0; e < E; ++e) tags_k[o * E + e] = agg_tags[((size_t)row * H * W + out_i) * E + e];
  }
}

// Small batches: kSplitWarps warps per (image, joint) row.  Every warp streams one contiguous segment of
// the row through its own sorted sink, started empty (a segment's smallest kept value never exceeds the
// sequential algorithm's at the same position, so the segment keeps and logs a superset of what enters
// there); warp 0 merges the segment lists.  The merged top M is final iff it is free of ties (no equal
// values among the top M, no other listed element equal to the M-th value, no segment evicted an element
// equal to it); otherwise warp 0 replays the concatenated segment logs, which are in index order, through
// the exact libstdc++ heap.
constexpr int kSplitWarps = 8;
constexpr int kSegLogCap = 320;

__global__ void __launch_bounds__(kSplitWarps * 32) topk_split_kernel(const float* __restrict__ agg_hm,
                                                                     const float* __restrict__ agg_tags,
                                                                     const uint32_t* __restrict__ mask,
                                                                     const float* __restrict__ wmax, int H, int W,
                                                                     int wpr, int E, int M, int force_exact,
                                                                     float* __restrict__ scores_k,
                                                                     int32_t* __restrict__ idx_k,
                                                                     int32_t* __restrict__ coords_k,
                                                                     float* __restrict__ tags_k) {
  __shared__ float s_v[kSplitWarps][32];
  __shared__ int s_i[kSplitWarps][32];
  __shared__ float s_logv[kSplitWarps][kSegLogCap];
  __shared__ int s_logi[kSplitWarps][kSegLogCap];
  __shared__ float s_evicted[kSplitWarps];
  __shared__ int s_nlog[kSplitWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x;
  const float* hm = agg_hm + (size_t)row * H * W;
  const uint32_t* mk = mask + (size_t)row * H * wpr;
  const float* wm = wmax + (size_t)row * H * wpr;
  const int nwords = H * wpr;

  if (!force_exact) {
    const int groups = (nwords + 127) / 128;   // segment bounds in whole groups of 128 words
    const int gb = (groups * warp) / kSplitWarps, ge = (groups * (warp + 1)) / kSplitWarps;
    SortedSink seg;
    seg.M = M; seg.lane = lane;
    seg.log_v = s_logv[warp]; seg.log_i = s_logi[warp]; seg.log_cap = kSegLogCap;
    seg.init_empty();
    scan_range(seg, hm, mk, wm, W, wpr, gb * 128, min(ge * 128, nwords), 0, lane);
    s_v[warp][lane] = seg.sv;
    s_i[warp][lane] = seg.si;
    if (lane == 0) {
      s_evicted[warp] = seg.any_evicted ? seg.evicted : -INFINITY;
      s_nlog[warp] = seg.n_log;
    }
  }
  __syncthreads();
  if (warp != 0) return;

  float out_v = 0.f;
  int out_i = 0;
  bool done = false, logs_ok = !force_exact;
  if (!force_exact) {
    SortedSink all;
    all.M = M; all.lane = lane;
    all.log_v = nullptr; all.log_i = nullptr; all.log_cap = 0;
    all.init_empty();
    for (int w = 0; w < kSplitWarps; ++w) {
      const float cvl = s_v[w][lane];
      const int cil = s_i[w][lane];
      for (int j = 0; j < M; ++j) {           // descending: stop at the first value that cannot matter
        const float cv = __shfl_sync(kFull, cvl, j);
        if (cv < all.top || cv == -INFINITY) break;
        all.offer(cv, __shfl_sync(kFull, cil, j));
      }
      logs_ok = logs_ok && s_nlog[w] <= kSegLogCap;
    }
    bool amb = all.ambiguous();
    for (int w = 0; w < kSplitWarps; ++w) amb = amb || (s_evicted[w] == all.top);
    if (!amb) { out_v = all.sv; out_i = all.si; done = true; }
  }
  if (!done) {
    HeapSink exact;
    exact.h = HeapRef{s_v[0], s_i[0]};
    exact.M = M; exact.lane = lane;
    if (logs_ok) {
      exact.init(first_element(hm, mk, W, wpr, M, lane));
      __syncwarp();
      // the logs hold a superset of the entering elements in index order: filter 32 entries at a time
      // against the live top, replay the survivors one by one
      float top = exact.top;
      for (int w = 0; w < kSplitWarps; ++w) {
        const int n = s_nlog[w];
        for (int i0 = 0; i0 < n; i0 += 32) {
          const int i = i0 + lane;
          const float cv = i < n ? s_logv[w][i] : -INFINITY;
          const int ci = i < n ? s_logi[w][i] : 0;
          uint32_t pass = __ballot_sync(kFull, i < n && ci >= M && cv > top);
          while (pass) {
            const int j = __ffs(pass) - 1;
            pass &= pass - 1;
            const float cvj = __shfl_sync(kFull, cv, j);
            const int cij = __shfl_sync(kFull, ci, j);
            if (cvj > top) {
              if (lane == 0) adjust_heap(exact.h, 0, M, cvj, cij);
              __syncwarp();
              top = exact.h.v[0];
            }
          }
        }
      }
    } else {
      scan_row(exact, hm, mk, wm, H, W, wpr, M, lane);
    }
    __syncwarp();
    if (lane == 0) sort_heap(exact.h, M);
    __syncwarp();
    if (lane < M) { out_v = exact.h.v[lane]; out_i = exact.h.i[lane]; }
  }
  if (lane < M) {
    const size_t o = (size_t)row * M + lane;
    scores_k[o] = out_v;
    idx_k[o] = out_i;
    coords_k[o * 2 + 0] = out_i % W;
    coords_k[o * 2 + 1] = out_i / W;
    for (int e = 0; e < E; ++e) tags_k[o * E + e] = agg_tags[((size_t)row * H * W + out_i) * E + e];
  }
}

}  // namespace

int launch_topk(const HpdParams* p, const HpdBuffers* buf, cudaStream_t st) {
  if (!buf->agg_hm || !buf->agg_tags || !buf->nms_mask || !buf->nms_wmax || !buf->scores_k || !buf->idx_k ||
      !buf->coords_k || !buf->tags_k) {
    set_error("hpd_topk: agg_hm, agg_tags, nms_mask, nms_wmax, scores_k, idx_k, coords_k, tags_k are required");
    return HPD_EINVAL;
  }
  const int rows = p->batch * p->num_kpts;
  const int wpr = (p->out_w + 31) / 32;
  // small batches: kSplitWarps warps per row cut the per-row latency; large ones fill the GPU with one warp per row
  const bool split = rows <= 512 && (long long)p->out_h * wpr >= 128 * kSplitWarps && !(p->force_generic & 2);
  if (split)
    topk_split_kernel<<<rows, kSplitWarps * 32, 0, st>>>(buf->agg_hm, buf->agg_tags, buf->nms_mask, buf->nms_wmax, p->out_h,
                                                         p->out_w, wpr, p->emb, p->max_people, p->force_generic & 1,
                                                         buf->scores_k, buf->idx_k, buf->coords_k, buf->tags_k);
  else
    topk_kernel<<<(rows + kTopkWarps - 1) / kTopkWarps, kTopkWarps * 32, 0, st>>>(
        buf->agg_hm, buf->agg_tags, buf->nms_mask, buf->nms_wmax, rows, p->out_h, p->out_w, wpr, p->emb, p->max_people,
        p->force_generic & 1, buf->scores_k, buf->idx_k, buf->coords_k, buf->tags_k);
  count_launch();
  return check_launch("topk_kernel");
}

}  // namespace hpd

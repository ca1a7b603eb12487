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

__global__ void __launch_bounds__(kTopkWarps * 32) topk_kernel(const float* __restrict__ agg_hm,
                                                              const float* __restrict__ agg_tags,
                                                              const uint32_t* __restrict__ mask,
                                                              const float* __restrict__ wmax, int rows, int H, int W,
                                                              int wpr, int E, int M, float* __restrict__ scores_k,
                                                              int32_t* __restrict__ idx_k,
                                                              int32_t* __restrict__ coords_k,
                                                              float* __restrict__ tags_k) {
  __shared__ float s_v[kTopkWarps][32];
  __shared__ int s_i[kTopkWarps][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * kTopkWarps + warp;
  if (row >= rows) return;
  const float* hm = agg_hm + (size_t)row * H * W;
  const uint32_t* mk = mask + (size_t)row * H * wpr;
  const float* wm = wmax + (size_t)row * H * wpr;
  HeapRef h{s_v[warp], s_i[warp]};

  // heap <- first M elements of the row (their NMS'd values)
  if (lane < M) {
    const int y = lane / W, x = lane % W;
    const float v = hm[lane];
    const bool keep = (mk[(size_t)y * wpr + (x >> 5)] >> (x & 31)) & 1u;
    h.v[lane] = keep ? v : __fmul_rn(v, 0.0f);
    h.i[lane] = lane;
  }
  __syncwarp();
  if (lane == 0) make_heap(h, M);
  __syncwarp();
  float top = h.v[0];

  const int nwords = H * wpr;
  for (int base = 0; base < nwords; base += 32) {
    const int wd = base + lane;
    const float wv = wd < nwords ? wm[wd] : -INFINITY;
    uint32_t pass = __ballot_sync(kFull, wv > top);
    while (pass) {
      const int l = __ffs(pass) - 1;
      pass &= pass - 1;
      if (!(__shfl_sync(kFull, wv, l) > top)) continue;   // the top may have risen meanwhile
      const int w2 = base + l;
      const int y = w2 / wpr, x = (w2 % wpr) * 32 + lane;
      const int idx = y * W + x;
      const bool valid = (x < W) && (idx >= M);
      const float v = valid ? hm[idx] : 0.f;
      const bool keep = (mk[w2] >> lane) & 1u;
      const float nv = keep ? v : __fmul_rn(v, 0.0f);
      uint32_t cand = __ballot_sync(kFull, valid && nv > top);
      while (cand) {
        const int j = __ffs(cand) - 1;
        cand &= cand - 1;
        const float cv = __shfl_sync(kFull, nv, j);
        if (cv > top) {   // std::__heap_select: comp(i, first) -> __pop_heap(first, middle, i)
          if (lane == 0) adjust_heap(h, 0, M, cv, y * W + (w2 % wpr) * 32 + j);
          __syncwarp();
          top = h.v[0];
        }
      }
    }
  }
  __syncwarp();
  if (lane == 0) sort_heap(h, M);
  __syncwarp();
  if (lane < M) {
    const int idx = h.i[lane];
    const size_t o = (size_t)row * M + lane;
    scores_k[o] = h.v[lane];
    idx_k[o] = idx;
    coords_k[o * 2 + 0] = idx % W;
    coords_k[o * 2 + 1] = idx / W;
    for (int e = 0; e < E; ++e) tags_k[o * E + e] = agg_tags[((size_t)row * H * W + idx) * E + e];
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
  topk_kernel<<<(rows + kTopkWarps - 1) / kTopkWarps, kTopkWarps * 32, 0, st>>>(
      buf->agg_hm, buf->agg_tags, buf->nms_mask, buf->nms_wmax, rows, p->out_h, p->out_w, wpr, p->emb, p->max_people,
      buf->scores_k, buf->idx_k, buf->coords_k, buf->tags_k);
  count_launch();
  return check_launch("topk_kernel");
}

}  // namespace hpd

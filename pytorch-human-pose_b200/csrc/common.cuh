// Shared device/host helpers for libhpdecode (sm_100a).  Compiled with -fmad=false: the only
// fused multiply-adds in the library are the explicit fmaf() calls that reproduce torch's CPU
// bilinear kernel.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/hpdecode.h"

namespace hpd {

constexpr unsigned kFull = 0xffffffffu;

// ---- launch bookkeeping / error reporting (api.cu) --------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_launch(const char* what);
int ensure_dynamic_smem(const void* kernel, size_t bytes, const char* name);

// ---- bilinear taps, torch CPU semantics (results.py:51,59,65; aten UpSampleKernel.cpp) -----
struct Tap {
  int i0, i1;
  float w0, w1;
};

__host__ __device__ __forceinline__ Tap axis_tap(float scale, int o, int in_size, int out_size) {
  Tap t;
  if (in_size == out_size) {
    t.i0 = t.i1 = o;
    t.w0 = 1.f;
    t.w1 = 0.f;
    return t;
  }
  float src = fmaf(scale, (float)o + 0.5f, -0.5f);
  if (src < 0.f) src = 0.f;
  int i0 = (int)src;
  if (i0 > in_size - 1) i0 = in_size - 1;
  float l1 = src - (float)i0;
  l1 = fminf(fmaxf(l1, 0.f), 1.f);
  t.i0 = i0;
  t.i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  t.w1 = l1;
  t.w0 = 1.f - l1;
  return t;
}

#ifdef __CUDACC__
// out = hy0*(wx0*v00 + wx1*v01) + hy1*(wx0*v10 + wx1*v11) with torch's FMA nesting
__device__ __forceinline__ float lerp2(float wx0, float wx1, float hy0, float hy1, float v00, float v01, float v10,
                                       float v11) {
  const float top = fmaf(wx0, v00, __fmul_rn(wx1, v01));
  const float bot = fmaf(wx0, v10, __fmul_rn(wx1, v11));
  return fmaf(hy0, top, __fmul_rn(hy1, bot));
}

// order-preserving float -> signed int map (for redux.sync max)
__device__ __forceinline__ int float_order_int(float f) {
  const int i = __float_as_int(f);
  return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float order_int_float(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

__device__ __forceinline__ float warp_max_float(float v) {
  return order_int_float(__reduce_max_sync(kFull, float_order_int(v)));
}

// numpy float32 pairwise-8 summation of n <= 128 strided values (np.mean / np.sum inner loop)
__device__ __forceinline__ float np_sum_pairwise8(const float* a, int n, int stride) {
  if (n < 8) {
    float r = -0.0f;
    for (int i = 0; i < n; ++i) r = __fadd_rn(r, a[i * stride]);
    return r;
  }
  float r0 = a[0], r1 = a[stride], r2 = a[2 * stride], r3 = a[3 * stride], r4 = a[4 * stride], r5 = a[5 * stride],
        r6 = a[6 * stride], r7 = a[7 * stride];
  int i = 8;
  for (; i < n - (n % 8); i += 8) {
    r0 = __fadd_rn(r0, a[(i + 0) * stride]);
    r1 = __fadd_rn(r1, a[(i + 1) * stride]);
    r2 = __fadd_rn(r2, a[(i + 2) * stride]);
    r3 = __fadd_rn(r3, a[(i + 3) * stride]);
    r4 = __fadd_rn(r4, a[(i + 4) * stride]);
    r5 = __fadd_rn(r5, a[(i + 5) * stride]);
    r6 = __fadd_rn(r6, a[(i + 6) * stride]);
    r7 = __fadd_rn(r7, a[(i + 7) * stride]);
  }
  float res = __fadd_rn(__fadd_rn(__fadd_rn(r0, r1), __fadd_rn(r2, r3)),
                        __fadd_rn(__fadd_rn(r4, r5), __fadd_rn(r6, r7)));
  for (; i < n; ++i) res = __fadd_rn(res, a[i * stride]);
  return res;
}

// np.mean(list of n float32 vectors of E components, axis=0) (grouping.py:114,213):
// E == 1 reduces a contiguous axis (pairwise-8), E == 2 accumulates row by row (sequential).
__device__ __forceinline__ void np_mean_vectors(const float* v, int n, int E, int row_stride, float* out) {
  if (E == 1) {
    out[0] = __fdiv_rn(__fadd_rn(0.0f, np_sum_pairwise8(v, n, row_stride)), (float)n);
  } else {
    for (int e = 0; e < E; ++e) {
      float s = 0.0f;
      for (int i = 0; i < n; ++i) s = __fadd_rn(s, v[i * row_stride + e]);
      out[e] = __fdiv_rn(s, (float)n);
    }
  }
}
#endif  // __CUDACC__

// ---- per-image result record (include/hpdecode.h: HpdRecordLayout) ------------------------------
__host__ __device__ inline HpdRecordLayout record_layout(int K, int M, int E) {
  HpdRecordLayout L;
  L.coco_stride = 3 * K + 1;
  L.reserved_ = 0;
  long long off = 0;
  L.off_coco = off;          off += 8LL * M * L.coco_stride;
  L.off_poses = off;         off += 4LL * M * K * (3 + E);
  L.off_person_scores = off; off += 4LL * M;
  L.off_n_person = off;      off += 4;
  L.off_flags = off;         off += 4;
  L.row_bytes = (off + 7) / 8 * 8;
  return L;
}

// ---- per-stage launchers (defined in the stage .cu files) -------------------------------------
int launch_aggregate_nms(const HpdParams* p, const HpdScaleInputs* scales, const HpdBuffers* buf, cudaStream_t st);
int launch_nms(const HpdParams* p, const HpdBuffers* buf, float* nms_out, cudaStream_t st);
int launch_topk(const HpdParams* p, const HpdBuffers* buf, cudaStream_t st);
int launch_group(const HpdParams* p, const HpdBuffers* buf, cudaStream_t st);
int launch_adjust_refine(const HpdParams* p, const HpdBuffers* buf, void* ws, size_t ws_bytes, cudaStream_t st);
size_t refine_workspace_bytes(const HpdParams* p);
int launch_resize(const HpdMap* in, int batch, int channels, float* out, int oh, int ow, cudaStream_t st);
int launch_prepare_input(const HpdImage* images, int batch, float* out, int oh, int ow, const float* mean,
                         const float* stdv, cudaStream_t st);
int multi_scale_size(int h, int w, int input_size, double current_scale, double min_scale, int32_t* size_wh,
                     int32_t* center_xy, double* scale_wh);
int affine_transform_matrix(const double* center, const double* scale, const int32_t* out_wh, int inverse, double* m);

}  // namespace hpd

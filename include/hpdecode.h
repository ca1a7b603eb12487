/* hpdecode -- B200 (sm_100a) bottom-up keypoint decode: C ABI of libhpdecode.so
 *
 * Drop-in boundary for the HigherHRNet decode path of thawro/pytorch-human-pose.  The
 * reference has no FFI layer (it is pure Python); each entry point below is the device twin
 * of one reference method, cited as /root/reference-relative file:line.  INTEGRATION.md shows
 * the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *  - Every pointer is a DEVICE pointer unless the name ends in _host.  The caller owns every
 *    buffer; the library allocates nothing, never synchronises, and enqueues all work on the
 *    given stream (a cudaStream_t passed as void*; NULL = the legacy default stream).
 *  - All maps are float32.  Network outputs are NCHW views with explicit batch / channel
 *    strides in ELEMENTS and contiguous rows (HigherHRNet's hm_lo / tag are channel slices of
 *    one 34-channel tensor: src/keypoints/architectures/higher_hrnet.py:78-79).
 *  - Return value: 0 on success, an HPD_E* code otherwise; hpd_last_error_string() gives the
 *    message of the calling thread's last failure.  No C++ exception crosses the ABI.
 *  - Re-entrant: one workspace per in-flight call; safe from one host thread per GPU.
 *  - Inputs must be finite (NaN ordering of torch.topk / MaxPool2d is not reproduced).
 */
#ifndef HPDECODE_H_
#define HPDECODE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define HPD_EXPORT __attribute__((visibility("default")))
#else
#define HPD_EXPORT
#endif

#define HPD_ABI_VERSION 2
#define HPD_MAX_KPTS 32     /* joints per person (COCO: 17) */
#define HPD_MAX_PEOPLE 32   /* max_num_people == top-k; the warp-wide Hungarian handles n <= 32 */
#define HPD_MAX_EMB 2       /* tag embedding dims: 1 (no flip test) or 2 (flip test) */
#define HPD_MAX_SCALES 4    /* test scales averaged into one heatmap */

enum {
  HPD_OK = 0,
  HPD_EINVAL = 1,       /* bad argument / unsupported shape */
  HPD_EWORKSPACE = 2,   /* workspace too small */
  HPD_ECUDA = 3         /* a CUDA call failed (launch configuration, etc.) */
};

enum {
  HPD_F32 = 0,          /* float32 maps (inference path, model.py:78-111) */
  HPD_F16 = 1           /* IEEE half maps: the validation-time caller runs the net under autocast
                         * (module.py:78,100-110); values are widened on load, all arithmetic stays float32 */
};

/* One NCHW tensor view. */
typedef struct HpdMap {
  const void* ptr;    /* NULL = absent */
  int64_t stride_b;   /* elements between images */
  int64_t stride_c;   /* elements between channels; rows are contiguous (stride_y == w) */
  int32_t h, w;
  int32_t dtype;      /* HPD_F32 or HPD_F16; all maps of one call share it */
  int32_t reserved_;
} HpdMap;

/* Raw network outputs for one test scale: the un-flipped forward and, when the flip test is on,
 * the forward of the horizontally flipped image (model.py:83-86). */
typedef struct HpdScaleInputs {
  HpdMap hm_lo, hm_hi, tag;         /* stage-1 heatmaps (S/4), stage-2 heatmaps (S/2), tags (S/4) */
  HpdMap hm_lo_f, hm_hi_f, tag_f;   /* flipped run; ptr == NULL when flip test is off */
} HpdScaleInputs;

typedef struct HpdParams {
  int32_t batch;        /* B images per call */
  int32_t num_kpts;     /* K <= HPD_MAX_KPTS */
  int32_t out_h, out_w; /* H, W of the aggregated maps == network input size (results.py:219) */
  int32_t emb;          /* E: 1 or 2 (2 iff flip inputs are given to the aggregate stage) */
  int32_t max_people;   /* M = max_num_people = top-k (grouping.py:70,153); H*W >= 64*M required */
  int32_t num_scales;   /* heatmaps are averaged over this many scales (1 in the reference) */
  int32_t tag_scale;    /* index of the scale whose tag maps are used */
  int32_t do_adjust;    /* grouping.py:274 */
  int32_t do_refine;    /* grouping.py:278 */
  int32_t tags_preflipped; /* 1: tag_f is already un-flipped and joint-permuted, i.e. it is the second entry of
                            * the reference's tags_heatmaps list (model.py:91-94) as from_preds receives it */
  int32_t force_generic; /* testing only, 0 in production.  bit 0: generic aggregation kernel instead of the specialised
                          * ones, and in top-k the literal libstdc++ heap code (one lane, shared memory) over the whole
                          * row instead of the floor mode / warp-wide heap; bit 1: one warp per row in top-k even for
                          * small batches; bit 2: rows with ties are streamed by that one warp instead of being handed
                          * to the second (8 warps per row) launch.  All variants are bit-identical (tests). */
  int32_t batches_in_flight; /* hint, 0 or 1 = unknown / latency matters: small batches (<= 512 rows) give every top-k row eight
                          * warps; >= 8: the caller keeps that many batches in flight (DecodePipeline), so the row gets one
                          * warp and the batches overlap instead (+5 % throughput at 8-16 images per batch).  Same results. */
  int32_t reserved_;
  double det_thr;       /* grouping.py:71,100  (compared in float64) */
  double tag_thr;       /* grouping.py:72,135  (compared in float64) */
  int32_t flip_index[HPD_MAX_KPTS];   /* COCO_FLIP_INDEX (transforms.py:11) */
  int32_t joints_order[HPD_MAX_KPTS]; /* MPPEHeatmapParser.joints_order (grouping.py:63-65) */
} HpdParams;

/* Device buffers written by the stages.  WPR = (W + 31) / 32 words per image row.
 * Every field is required by hpd_decode; single-stage calls document what they touch. */
typedef struct HpdBuffers {
  float* agg_hm;        /* [B,K,H,W]      aggregated heatmaps (results.py:227)                  */
  float* agg_tags;      /* [B,K,H,W,E]    resized tags, E innermost (results.py:229-230)        */
  uint32_t* nms_mask;   /* [B,K,H,WPR]    bit x%32 of word (y, x/32): pixel survives NMS         */
  float* nms_wmax;      /* [B,K,H,WPR]    max of the NMS'd values of the word's pixels           */
  float* hm_wmax;       /* [B,K,H,WPR]    max of the raw aggregated values of the word's pixels  */
  float* scores_k;      /* [B,K,M]        top-k values, torch CPU order (grouping.py:153)        */
  int32_t* idx_k;       /* [B,K,M]        flat indices y*W+x                                     */
  int32_t* coords_k;    /* [B,K,M,2]      (x, y) (grouping.py:163-165)                           */
  float* tags_k;        /* [B,K,M,E]      (grouping.py:158-161)                                  */
  float* poses;         /* [B,M,K,3+E]    grouped joints [x,y,score,tag..]; zero-filled rows     */
  float* person_scores; /* [B,M]          grouping.py:276                                        */
  int32_t* n_person;    /* [B]            persons returned (<= M)                                */
  int32_t* flags;       /* [B]            bit0: empty-scene fallback fired (grouping.py:262-269)  */
  float* tag_bmin;      /* [B,K,HB,WPR]   HB = (H+3)/4: lower / upper bound of the first tag component over     */
  float* tag_bmax;      /* [B,K,HB,WPR]   4 image rows x one 32-pixel word (refine prefilter; -inf/+inf = none) */
  uint8_t* records;     /* [B][row_bytes] optional (NULL = not written): one result record per image, laid out as
                         * hpd_record_layout() says; written by the epilogue of hpd_adjust_refine / hpd_decode     */
  const double* inv_affine; /* [B][6] optional: per-image 2x3 matrix (row major) that maps network-input pixels back
                         * to the raw image, i.e. get_affine_transform(center, scale, 0, (W, H), inverse=True)
                         * (results.py:158-171, base/transforms/utils.py:25-57).  NULL = identity              */
} HpdBuffers;

/* Layout of one result record (all offsets in bytes from the start of the image's row; row_bytes % 8 == 0).
 * The record is everything the reference's callers read after a decode, ready for ONE device->host copy:
 *   coco          f64 [M][3K+1]  per person (x, y, 1) * K then the person score: the "keypoints" list and "score"
 *                                of evaluate_dataset's COCO record (bin/eval.py:31-47) with x, y back-projected to
 *                                the raw image (results.py:158-171,189-201,244).  Persons >= n_person: zeros.
 *                                Like the reference, back-projected coordinates are rounded to float32 (they are
 *                                written into a float32 array, results.py:165-170) -- except in the empty-scene
 *                                fallback, whose pseudo-person is float64 throughout (grouping.py:262-269).
 *   poses         f32 [M][K][3+E] grouped joints in network-input pixels (x, y, score, tags), grouping.py:283
 *   person_scores f32 [M]        grouping.py:276
 *   n_person      i32            persons returned
 *   flags         i32            bit 0: empty-scene fallback fired                                              */
typedef struct HpdRecordLayout {
  int64_t row_bytes;
  int64_t off_coco, off_poses, off_person_scores, off_n_person, off_flags;
  int32_t coco_stride;  /* doubles per person: 3K + 1 */
  int32_t reserved_;
} HpdRecordLayout;

/* One raw image for hpd_prepare_input: uint8, HWC, 3 channels, resident on the device. */
typedef struct HpdImage {
  const uint8_t* ptr;
  int64_t stride_row;   /* bytes between rows (>= 3*w) */
  int32_t h, w;
  double m[6];          /* the forward 2x3 matrix cv2.warpAffine receives (base/transforms/utils.py:95-96) */
} HpdImage;

HPD_EXPORT int hpd_abi_version(void);
HPD_EXPORT const char* hpd_last_error_string(void);

/* Bytes of scratch hpd_decode / hpd_adjust_refine need for these params. */
HPD_EXPORT int hpd_workspace_bytes(const HpdParams* p, size_t* out_bytes);

/* (a)+(b) fused aggregation + NMS: model.py:85-96, results.py:46-67,225-230, grouping.py:80-83.
 * Reads scales[0..num_scales); writes agg_hm, agg_tags, nms_mask, nms_wmax, hm_wmax, tag_bmin, tag_bmax. */
HPD_EXPORT int hpd_aggregate_nms(const HpdParams* p, const HpdScaleInputs* scales, const HpdBuffers* buf, void* stream);

/* Standalone bilinear resize with torch's CPU arithmetic (BaseKeypointsResult.match_heatmaps_size,
 * resize_heatmaps_list, resize_heatmaps: results.py:46-67).  in: [batch,channels,h,w] view;
 * out: contiguous [batch,channels,out_h,out_w]. */
HPD_EXPORT int hpd_resize_bilinear(const HpdMap* in, int batch, int channels, float* out, int out_h, int out_w,
                                   void* stream);

/* (b) alone, for callers that hand in aggregated maps (MPPEHeatmapParser.parse/top_k/nms,
 * grouping.py:80-83,150): reads agg_hm; writes nms_mask, nms_wmax, hm_wmax, tag_bmin/bmax (= no bound) and, if nms_out is
 * not NULL, the float NMS'd map [B,K,H,W] exactly as the reference's nms() returns it. */
HPD_EXPORT int hpd_nms(const HpdParams* p, const HpdBuffers* buf, float* nms_out, void* stream);

/* (c) per-joint top-k with torch-CPU (std::partial_sort) tie order: grouping.py:147-170.
 * Reads agg_hm, agg_tags, nms_mask, nms_wmax; writes scores_k, idx_k, coords_k, tags_k. */
HPD_EXPORT int hpd_topk(const HpdParams* p, const HpdBuffers* buf, void* stream);

/* (d) associative-embedding grouping: grouping.py:85-145 + py_max_match :55-59 + fallback :262-269.
 * Reads scores_k, coords_k, tags_k; writes poses, n_person, flags. */
HPD_EXPORT int hpd_group(const HpdParams* p, const HpdBuffers* buf, void* stream);

/* (e) adjust + person score + refine: grouping.py:172-191, :276, :193-250.
 * Reads agg_hm, agg_tags, hm_wmax, tag_bmin, tag_bmax, idx_k, scores_k, n_person, flags; updates poses in place;
 * writes person_scores and, if buf->records is given, the per-image result records (back-projection through
 * buf->inv_affine + COCO layout, results.py:158-201,240-244, bin/eval.py:31-47) in the epilogue of its last kernel. */
HPD_EXPORT int hpd_adjust_refine(const HpdParams* p, const HpdBuffers* buf, void* workspace, size_t workspace_bytes,
                      void* stream);

/* Whole path: hpd_aggregate_nms -> hpd_topk -> hpd_group -> hpd_adjust_refine on one stream.
 * If scales == NULL the aggregation is skipped and buf->agg_hm / agg_tags are taken as inputs
 * (the MPPEHeatmapParser.parse entry, grouping.py:252). */
HPD_EXPORT int hpd_decode(const HpdParams* p, const HpdScaleInputs* scales, const HpdBuffers* buf, void* workspace,
               size_t workspace_bytes, void* stream);

/* Offsets and size of the per-image result record for these params (num_kpts, max_people, emb). */
HPD_EXPORT int hpd_record_layout(const HpdParams* p, HpdRecordLayout* out);

/* ---- input side: InferenceKeypointsModel.prepare_input, model.py:70-76 ---------------------
 * Host-side geometry, float64 like the reference (no device work, no stream):
 * get_multi_scale_size (base/transforms/utils.py:60-87): resized (w, h), center, scale of an img_h x img_w image. */
HPD_EXPORT int hpd_multi_scale_size(int img_h, int img_w, int input_size, double current_scale, double min_scale,
                                    int32_t size_resized_wh[2], int32_t center_xy[2], double scale_wh[2]);

/* get_affine_transform(center, scale, rot=0, output_size, inverse) (base/transforms/utils.py:25-57): float32 point
 * triples, then cv2.getAffineTransform's 6x6 LU solve replayed in float64 (bit-identical to OpenCV 4.x). */
HPD_EXPORT int hpd_get_affine_transform(const double center_xy[2], const double scale_wh[2], const int32_t output_size_wh[2],
                                        int inverse, double m_out[6]);

/* Both of the above for n images in one call (the batched evaluation loop): per image i the resized size, center, scale,
 * the forward matrix (for hpd_prepare_input) and the inverse one (for HpdBuffers.inv_affine). */
HPD_EXPORT int hpd_prepare_geometry(int n, const int32_t* img_h, const int32_t* img_w, int input_size, double current_scale,
                                    double min_scale, int32_t* size_resized_wh /*[n][2]*/, int32_t* center_xy /*[n][2]*/,
                                    double* scale_wh /*[n][2]*/, double* m_forward /*[n][6]*/, double* m_inverse /*[n][6]*/);

/* resize_align_multi_scale's cv2.warpAffine (INTER_LINEAR, BORDER_CONSTANT 0; base/transforms/utils.py:96) followed by
 * T.ToTensor + T.Normalize (model.py:45-50): images_host[b] (uint8 HWC on the device, descriptors on the host) ->
 * out [batch,3,out_h,out_w] float32.  The warp replays OpenCV's fixed-point arithmetic (10-bit coordinates, 5-bit
 * sub-pixel positions, 15-bit weights) and the normalisation torch's float32 sequence, so out is bit-identical to
 * the reference's tensor.  mean / std: 3 floats each (host). */
HPD_EXPORT int hpd_prepare_input(const HpdImage* images_host, int batch, float* out, int out_h, int out_w,
                                 const float mean[3], const float std_[3], void* stream);

/* Number of kernel launches the previous call on this thread enqueued (for bench accounting). */
HPD_EXPORT int hpd_last_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* HPDECODE_H_ */

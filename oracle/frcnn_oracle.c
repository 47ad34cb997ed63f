/*
 * frcnn_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the region-proposal / RoI hot path of
 * antoineBarbez/2D_object_detection *including* the TensorFlow kernels that
 * path dispatches to.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library, and only as the
 * checker / the timed CPU baseline -- never as the thing shipped.
 *
 * PARITY UNPINNED: the reference ships no tests, fixtures or golden vectors
 * and TensorFlow 2 (un-vendored, unpinned: train_faster_rcnn.py:16 only asserts
 * tf.__version__.startswith("2")) cannot be installed here, so this file is
 * pinned by (i) hand-computed known-answer tests, (ii) an independent NumPy
 * restatement (oracle/reference_np.py), (iii) torchvision.ops.nms and
 * torch grid_sample/autograd cross-checks (tests/test_oracle_*.py).
 *
 * All arithmetic is fp32, evaluated in the reference's operation order and
 * compiled with -ffp-contract=off (no FMA contraction), matching TF's CPU
 * kernels (official x86 wheels are built without FMA).
 *
 * Box layout everywhere: [x_min, y_min, x_max, y_max].
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---------------------------------------------------------------------------
 * models/detectors/rpn_detector.py:162-199  RPNDetector._generate_anchors
 *   meshgrid(scales, ratios) is xy-indexed => flat shape index k = r*ns + s
 *   (:178-180); h = s/sqrt(r)*base_h, w = s*sqrt(r)*base_w (:182-184);
 *   centres at j*stride, i*stride with no half-stride offset (:186-188);
 *   anchor index = (i*Wf + j)*A + k (:190-199).
 * ------------------------------------------------------------------------- */
ORC_API void orc_generate_anchors(int Hf, int Wf, const float* scales, int ns, const float* ratios, int nr,
                                  float base_h, float base_w, float stride_h, float stride_w, float* out) {
    const int A = ns * nr;
    float* hs = (float*)malloc(sizeof(float) * A);
    float* ws = (float*)malloc(sizeof(float) * A);
    for (int r = 0; r < nr; ++r)
        for (int s = 0; s < ns; ++s) {
            const float rs = sqrtf(ratios[r]);
            float h = scales[s] / rs;
            h = h * base_h;
            float w = scales[s] * rs;
            w = w * base_w;
            hs[r * ns + s] = h;
            ws[r * ns + s] = w;
        }
    for (int i = 0; i < Hf; ++i)
        for (int j = 0; j < Wf; ++j) {
            const float cx = (float)j * stride_w;
            const float cy = (float)i * stride_h;
            for (int k = 0; k < A; ++k) {
                float* o = out + ((size_t)(i * Wf + j) * A + k) * 4;
                const float hw = 0.5f * ws[k];
                const float hh = 0.5f * hs[k];
                o[0] = cx - hw;
                o[1] = cy - hh;
                o[2] = cx + hw;
                o[3] = cy + hh;
            }
        }
    free(hs);
    free(ws);
}

/* rpn_detector.py:215-224: anchors fully inside the image; ascending index order.
 * Returns the number of valid anchors; idx_out may be NULL to count only. */
ORC_API int64_t orc_valid_anchor_indices(const float* anchors, int64_t n, float img_w, float img_h, int64_t* idx_out) {
    int64_t m = 0;
    for (int64_t i = 0; i < n; ++i) {
        const float* a = anchors + i * 4;
        if (a[0] >= 0.0f && a[1] >= 0.0f && a[2] <= img_w && a[3] <= img_h) {
            if (idx_out) idx_out[m] = i;
            ++m;
        }
    }
    return m;
}

/* utils/boxes.py:4-17 clip_to_window; window unpacked as x_min,y_min,x_max,y_max (:13). */
ORC_API void orc_clip_to_window(const float* boxes, int64_t n, const float* window, float* out) {
    const float lo[4] = {window[0], window[1], window[0], window[1]};
    const float hi[4] = {window[2], window[3], window[2], window[3]};
    for (int64_t i = 0; i < n * 4; ++i) {
        const int c = (int)(i & 3);
        out[i] = fmaxf(fminf(boxes[i], hi[c]), lo[c]);
    }
}

/* utils/boxes.py:20-41 decode for one box. */
static inline void decode_one(const float* t, const float* ref, float* o) {
    const float cxr = (ref[2] + ref[0]) / 2.0f, cyr = (ref[3] + ref[1]) / 2.0f; /* :34 */
    const float wr = ref[2] - ref[0], hr = ref[3] - ref[1];                     /* :35 */
    float cx = t[0] * wr;
    cx = cx + cxr; /* :38 mul then add, separately rounded */
    float cy = t[1] * hr;
    cy = cy + cyr;
    const float w = expf(t[2]) * wr; /* :39 */
    const float h = expf(t[3]) * hr;
    const float hw = 0.5f * w, hh = 0.5f * h; /* :41 */
    o[0] = cx - hw;
    o[1] = cy - hh;
    o[2] = cx + hw;
    o[3] = cy + hh;
}

/* utils/boxes.py:44-73 encode for one box. */
static inline void encode_one(const float* b, const float* ref, float* o) {
    const float cxr = (ref[2] + ref[0]) / 2.0f, cyr = (ref[3] + ref[1]) / 2.0f;
    const float wr = ref[2] - ref[0], hr = ref[3] - ref[1];
    const float cx = (b[2] + b[0]) / 2.0f, cy = (b[3] + b[1]) / 2.0f;
    const float w = b[2] - b[0], h = b[3] - b[1];
    o[0] = (cx - cxr) / wr;
    o[1] = (cy - cyr) / hr;
    o[2] = logf(w / wr);
    o[3] = logf(h / hr);
}

/* decode / encode with a reference broadcast of period ref_n boxes (ref index = i % ref_n). */
ORC_API void orc_decode(const float* deltas, const float* refs, int64_t n, int64_t ref_n, float* out) {
    for (int64_t i = 0; i < n; ++i) decode_one(deltas + i * 4, refs + (i % ref_n) * 4, out + i * 4);
}
ORC_API void orc_encode(const float* boxes, const float* refs, int64_t n, int64_t ref_n, float* out) {
    for (int64_t i = 0; i < n; ++i) encode_one(boxes + i * 4, refs + (i % ref_n) * 4, out + i * 4);
}

/* utils/boxes.py:76-93: to_absolute multiplies, to_relative truly divides by [W,H,W,H]. */
ORC_API void orc_to_absolute(const float* boxes, int64_t n, float img_h, float img_w, float* out) {
    const float s[4] = {img_w, img_h, img_w, img_h};
    for (int64_t i = 0; i < n * 4; ++i) out[i] = boxes[i] * s[i & 3];
}
ORC_API void orc_to_relative(const float* boxes, int64_t n, float img_h, float img_w, float* out) {
    const float s[4] = {img_w, img_h, img_w, img_h};
    for (int64_t i = 0; i < n * 4; ++i) out[i] = boxes[i] / s[i & 3];
}

/* utils/metrics.py:136-208 area / intersection / iou. */
static inline float ref_iou(const float* a, const float* b) {
    const float dw = fminf(a[2], b[2]) - fmaxf(a[0], b[0]); /* :173 */
    const float dh = fminf(a[3], b[3]) - fmaxf(a[1], b[1]); /* :174 */
    const float inter = fmaxf(0.0f, dw) * fmaxf(0.0f, dh);  /* :175 */
    const float area_a = (a[2] - a[0]) * (a[3] - a[1]);     /* :147 */
    const float area_b = (b[2] - b[0]) * (b[3] - b[1]);
    float uni = area_a + area_b; /* :206 */
    uni = uni - inter;
    return inter == 0.0f ? 0.0f : inter / uni; /* :208 */
}
ORC_API void orc_iou_pairwise(const float* b1, int64_t n, const float* b2, int64_t m, float* out) {
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = 0; j < m; ++j) out[i * m + j] = ref_iou(b1 + i * 4, b2 + j * 4);
}
ORC_API void orc_iou_elementwise(const float* b1, const float* b2, int64_t n, float* out) {
    for (int64_t i = 0; i < n; ++i) out[i] = ref_iou(b1 + i * 4, b2 + i * 4);
}
ORC_API void orc_area(const float* b, int64_t n, float* out) {
    for (int64_t i = 0; i < n; ++i) out[i] = (b[i * 4 + 2] - b[i * 4]) * (b[i * 4 + 3] - b[i * 4 + 1]);
}
ORC_API void orc_intersection_pairwise(const float* b1, int64_t n, const float* b2, int64_t m, float* out) {
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = 0; j < m; ++j) {
            const float* a = b1 + i * 4;
            const float* b = b2 + j * 4;
            const float dw = fminf(a[2], b[2]) - fmaxf(a[0], b[0]);
            const float dh = fminf(a[3], b[3]) - fmaxf(a[1], b[1]);
            out[i * m + j] = fmaxf(0.0f, dw) * fmaxf(0.0f, dh);
        }
}

/* ---------------------------------------------------------------------------
 * TF CombinedNonMaxSuppression IOU (tensorflow/core/kernels/image/
 * non_max_suppression_op.cc, restated from the published algorithm):
 * min/max coordinate normalisation, zero for non-positive areas, IEEE division.
 * ------------------------------------------------------------------------- */
static inline float tf_nms_iou(const float* a, const float* b) {
    const float a0 = fminf(a[0], a[2]), a1 = fminf(a[1], a[3]), a2 = fmaxf(a[0], a[2]), a3 = fmaxf(a[1], a[3]);
    const float b0 = fminf(b[0], b[2]), b1 = fminf(b[1], b[3]), b2 = fmaxf(b[0], b[2]), b3 = fmaxf(b[1], b[3]);
    const float area_a = (a2 - a0) * (a3 - a1);
    const float area_b = (b2 - b0) * (b3 - b1);
    if (area_a <= 0.0f || area_b <= 0.0f) return 0.0f;
    const float i0 = fmaxf(a0, b0), i1 = fmaxf(a1, b1), i2 = fminf(a2, b2), i3 = fminf(a3, b3);
    const float inter = fmaxf(i2 - i0, 0.0f) * fmaxf(i3 - i1, 0.0f);
    float uni = area_a + area_b;
    uni = uni - inter;
    return inter / uni;
}
ORC_API float orc_tf_nms_iou(const float* a, const float* b) { return tf_nms_iou(a, b); }

typedef struct {
    float score;
    int32_t idx;
} cand_t;

/* score descending, index ascending on ties (TF leaves ties unordered). */
static int cand_cmp(const void* pa, const void* pb) {
    const cand_t* a = (const cand_t*)pa;
    const cand_t* b = (const cand_t*)pb;
    if (a->score > b->score) return -1;
    if (a->score < b->score) return 1;
    return (a->idx > b->idx) - (a->idx < b->idx);
}

/*
 * Single-class greedy NMS in TF's order: candidates with score > score_thr
 * visited by descending score, kept iff no already kept box has IoU > iou_thr
 * (strict), stop at max_out kept.  boxes has stride box_stride floats, scores
 * stride score_stride.  pre_nms_top_n <= 0 means "all candidates" (reference).
 * Returns number kept; kept_idx[] receives the box indices in keep order.
 */
static int nms_single(const float* boxes, int64_t box_stride, const float* scores, int64_t score_stride, int n,
                      float score_thr, float iou_thr, int max_out, int pre_nms_top_n, int32_t* kept_idx) {
    cand_t* c = (cand_t*)malloc(sizeof(cand_t) * (size_t)(n > 0 ? n : 1));
    int m = 0;
    for (int i = 0; i < n; ++i) {
        const float s = scores[(int64_t)i * score_stride];
        if (s > score_thr) {
            c[m].score = s;
            c[m].idx = i;
            ++m;
        }
    }
    qsort(c, (size_t)m, sizeof(cand_t), cand_cmp);
    if (pre_nms_top_n > 0 && m > pre_nms_top_n) m = pre_nms_top_n;
    const int cap = max_out < n ? max_out : n; /* min(max_output_size_per_class, num_boxes) T17 */
    int kept = 0;
    for (int t = 0; t < m && kept < cap; ++t) {
        const float* bi = boxes + (int64_t)c[t].idx * box_stride;
        int keep = 1;
        for (int j = kept - 1; j >= 0; --j) { /* newest -> oldest, break on first hit */
            if (tf_nms_iou(bi, boxes + (int64_t)kept_idx[j] * box_stride) > iou_thr) {
                keep = 0;
                break;
            }
        }
        if (keep) kept_idx[kept++] = c[t].idx;
    }
    free(c);
    return kept;
}

/* Stand-alone single-class NMS on explicit boxes (for stage-level parity tests). */
ORC_API int orc_nms(const float* boxes, const float* scores, int n, float score_thr, float iou_thr, int max_out,
                    int pre_nms_top_n, int32_t* kept_idx) {
    return nms_single(boxes, 4, scores, 1, n, score_thr, iou_thr, max_out, pre_nms_top_n, kept_idx);
}

typedef struct {
    float score;
    int32_t cls;
    int32_t idx;
    int32_t ord; /* position in the class-major concatenation: tie-break */
} merged_t;

static int merged_cmp(const void* pa, const void* pb) {
    const merged_t* a = (const merged_t*)pa;
    const merged_t* b = (const merged_t*)pb;
    if (a->score > b->score) return -1;
    if (a->score < b->score) return 1;
    return (a->ord > b->ord) - (a->ord < b->ord);
}

/*
 * tf.image.combined_non_max_suppression(boxes[B,N,q,4], scores[B,N,nc], ...,
 * pad_per_class=False, clip_boxes=True) as called by
 * utils/post_processing.py:53-55.  q is 1 or nc.
 * Outputs are zero padded to max_total (R): boxes[B,R,4] (clipped to [0,1]),
 * scores[B,R], classes[B,R] (int32), valid[B]; kept_idx[B,R] (region index, -1
 * padded) is an extra output used by the index-set parity tests.
 */
ORC_API void orc_combined_nms(const float* boxes, const float* scores, int B, int N, int q, int nc,
                              int max_per_class, int max_total, float iou_thr, float score_thr,
                              int pre_nms_top_n, float* out_boxes, float* out_scores, int32_t* out_classes,
                              int32_t* out_valid, int32_t* out_idx) {
#pragma omp parallel for schedule(dynamic)
    for (int b = 0; b < B; ++b) {
        const float* bb = boxes + (size_t)b * N * q * 4;
        const float* sb = scores + (size_t)b * N * nc;
        const int per_cap = max_per_class < N ? max_per_class : N;
        merged_t* mg = (merged_t*)malloc(sizeof(merged_t) * (size_t)(nc * (per_cap > 0 ? per_cap : 1)));
        int32_t* kept = (int32_t*)malloc(sizeof(int32_t) * (size_t)(per_cap > 0 ? per_cap : 1));
        int total = 0;
        for (int c = 0; c < nc; ++c) {
            const int qi = q > 1 ? c : 0;
            const int k = nms_single(bb + qi * 4, (int64_t)q * 4, sb + c, nc, N, score_thr, iou_thr, max_per_class,
                                     pre_nms_top_n, kept);
            for (int t = 0; t < k; ++t) {
                mg[total].score = sb[(size_t)kept[t] * nc + c];
                mg[total].cls = c;
                mg[total].idx = kept[t];
                mg[total].ord = total;
                ++total;
            }
        }
        qsort(mg, (size_t)total, sizeof(merged_t), merged_cmp);
        const int nv = total < max_total ? total : max_total;
        for (int t = 0; t < max_total; ++t) {
            float* ob = out_boxes + ((size_t)b * max_total + t) * 4;
            if (t < nv) {
                const int qi = q > 1 ? mg[t].cls : 0;
                const float* src = bb + ((size_t)mg[t].idx * q + qi) * 4;
                for (int k = 0; k < 4; ++k) ob[k] = fmaxf(fminf(src[k], 1.0f), 0.0f); /* clip_boxes=True */
                out_scores[(size_t)b * max_total + t] = mg[t].score;
                out_classes[(size_t)b * max_total + t] = mg[t].cls;
                if (out_idx) out_idx[(size_t)b * max_total + t] = mg[t].idx;
            } else {
                ob[0] = ob[1] = ob[2] = ob[3] = 0.0f;
                out_scores[(size_t)b * max_total + t] = 0.0f;
                out_classes[(size_t)b * max_total + t] = 0;
                if (out_idx) out_idx[(size_t)b * max_total + t] = -1;
            }
        }
        out_valid[b] = nv;
        free(mg);
        free(kept);
    }
}

/*
 * utils/post_processing.py:6-63 postprocess_output.
 * regions: [N,4] (regions_batched=0) or [B,N,4] (=1), absolute pixels;
 * pred_scores [B,N,nc+1]; pred_boxes [B,N,nc,4] encoded deltas.
 * decoded_rel (optional, [B,N,nc,4]) receives the decoded relative boxes that
 * were fed to NMS (unclipped -- trap T2).
 */
ORC_API void orc_postprocess_output(float img_h, float img_w, const float* regions, int regions_batched,
                                    const float* pred_scores, const float* pred_boxes, int B, int N, int nc,
                                    float score_thr, float iou_thr, int max_per_class, int max_total,
                                    int pre_nms_top_n, float* out_boxes, float* out_scores, int32_t* out_classes,
                                    int32_t* out_valid, int32_t* out_idx, float* decoded_rel) {
    const size_t nbox = (size_t)B * N * nc;
    float* rel = decoded_rel ? decoded_rel : (float*)malloc(sizeof(float) * nbox * 4);
    float* fg = (float*)malloc(sizeof(float) * nbox);
    const float s[4] = {img_w, img_h, img_w, img_h};
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)B * N; ++i) {
        const int64_t n = i % N;
        const float* ref = regions + (regions_batched ? i : n) * 4; /* :40-43 tile over classes */
        for (int c = 0; c < nc; ++c) {
            float d[4];
            decode_one(pred_boxes + ((size_t)i * nc + c) * 4, ref, d);                  /* :46 */
            for (int k = 0; k < 4; ++k) rel[((size_t)i * nc + c) * 4 + k] = d[k] / s[k]; /* :49 */
            fg[(size_t)i * nc + c] = pred_scores[(size_t)i * (nc + 1) + 1 + c];          /* :52 */
        }
    }
    orc_combined_nms(rel, fg, B, N, nc, nc, max_per_class, max_total, iou_thr, score_thr, pre_nms_top_n, out_boxes,
                     out_scores, out_classes, out_valid, out_idx); /* :53-55 */
    if (!decoded_rel) free(rel);
    free(fg);
}

/* ---------------------------------------------------------------------------
 * models/detectors/fast_rcnn_detector.py:142-177 ROIPooling.call:
 * tf.image.crop_and_resize (bilinear, extrapolation 0, S = P*k) followed by
 * MaxPool2D(k) stride k VALID.  rois [B,R,4] relative [x1,y1,x2,y2]
 * (reordered to TF's [y1,x1,y2,x2] at :157-158).  out [B,R,P,P,C];
 * argmax (optional) [B,R,P,P,C] uint8 = dy*k+dx of the first maximum in
 * row-major window order (TF-CPU MaxPoolGrad rule).
 * ------------------------------------------------------------------------- */
typedef struct {
    int valid;
    int lo, hi;
    float lerp;
} samp_t;

static inline samp_t crop_coord(float c1, float c2, int extent, int S, int i) {
    samp_t r;
    float in;
    if (S > 1) {
        float scale = (c2 - c1) * (float)(extent - 1);
        scale = scale / (float)(S - 1);
        const float a = c1 * (float)(extent - 1);
        const float bb = (float)i * scale;
        in = a + bb;
    } else {
        in = 0.5f * (c1 + c2);
        in = in * (float)(extent - 1);
    }
    /* TF tests `in < 0 || in > extent - 1`; a NaN coordinate passes that test and then indexes out of bounds
     * (undefined behaviour in the reference).  Defined choice here and in the CUDA path: NaN -> not sampled. */
    if (!(in >= 0.0f && in <= (float)(extent - 1))) {
        r.valid = 0;
        r.lo = r.hi = 0;
        r.lerp = 0.0f;
        return r;
    }
    r.valid = 1;
    r.lo = (int)floorf(in);
    r.hi = (int)ceilf(in);
    r.lerp = in - (float)r.lo;
    return r;
}

ORC_API void orc_roi_pool_forward(const float* feat, int B, int Hf, int Wf, int C, const float* rois, int R, int P,
                                  int k, float* out, uint8_t* argmax) {
    const int S = P * k;
#pragma omp parallel for schedule(dynamic, 4)
    for (int br = 0; br < B * R; ++br) {
        const int b = br / R;
        const float* roi = rois + (size_t)br * 4;
        const float x1 = roi[0], y1 = roi[1], x2 = roi[2], y2 = roi[3];
        samp_t sy[64], sx[64];
        for (int i = 0; i < S; ++i) {
            sy[i] = crop_coord(y1, y2, Hf, S, i);
            sx[i] = crop_coord(x1, x2, Wf, S, i);
        }
        const float* fb = feat + (size_t)b * Hf * Wf * C;
        for (int py = 0; py < P; ++py)
            for (int px = 0; px < P; ++px) {
                float* o = out + (((size_t)br * P + py) * P + px) * C;
                uint8_t* am = argmax ? argmax + (((size_t)br * P + py) * P + px) * C : NULL;
                for (int c = 0; c < C; ++c) {
                    float best = -INFINITY;
                    int besti = 0;
                    for (int dy = 0; dy < k; ++dy)
                        for (int dx = 0; dx < k; ++dx) {
                            const samp_t* ys = &sy[py * k + dy];
                            const samp_t* xs = &sx[px * k + dx];
                            float v = 0.0f; /* extrapolation_value */
                            if (ys->valid && xs->valid) {
                                const float tl = fb[((size_t)ys->lo * Wf + xs->lo) * C + c];
                                const float tr = fb[((size_t)ys->lo * Wf + xs->hi) * C + c];
                                const float bl = fb[((size_t)ys->hi * Wf + xs->lo) * C + c];
                                const float brr = fb[((size_t)ys->hi * Wf + xs->hi) * C + c];
                                float t = (tr - tl) * xs->lerp;
                                t = tl + t;
                                float bo = (brr - bl) * xs->lerp;
                                bo = bl + bo;
                                float d = (bo - t) * ys->lerp;
                                v = t + d;
                            }
                            if (v > best || (dy == 0 && dx == 0)) { /* strict >: first max wins */
                                best = v;
                                besti = dy * k + dx;
                            }
                        }
                    o[c] = best;
                    if (am) am[c] = (uint8_t)besti;
                }
            }
    }
}

/*
 * Backward of the above w.r.t. feature maps (models/faster_rcnn.py:73,103):
 * MaxPoolGrad routes dY to the arg-max sample, CropAndResizeGradImage scatters
 * the four bilinear weights; out-of-range samples contribute nothing; no box
 * gradient (rois are stop_gradient, post_processing.py:58).
 * Accumulation order: RoIs in index order, bins row-major (sequential, as the
 * TF CPU kernel).  dfeat [B,Hf,Wf,C] is overwritten.
 */
ORC_API void orc_roi_pool_backward(const float* feat, int B, int Hf, int Wf, int C, const float* rois, int R, int P,
                                   int k, const float* dy_in, float* dfeat) {
    const int S = P * k;
    memset(dfeat, 0, sizeof(float) * (size_t)B * Hf * Wf * C);
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        const float* fb = feat + (size_t)b * Hf * Wf * C;
        float* gb = dfeat + (size_t)b * Hf * Wf * C;
        for (int r = 0; r < R; ++r) {
            const int br = b * R + r;
            const float* roi = rois + (size_t)br * 4;
            samp_t sy[64], sx[64];
            for (int i = 0; i < S; ++i) {
                sy[i] = crop_coord(roi[1], roi[3], Hf, S, i);
                sx[i] = crop_coord(roi[0], roi[2], Wf, S, i);
            }
            for (int py = 0; py < P; ++py)
                for (int px = 0; px < P; ++px) {
                    const float* g = dy_in + (((size_t)br * P + py) * P + px) * C;
                    for (int c = 0; c < C; ++c) {
                        float best = -INFINITY;
                        int bdy = 0, bdx = 0;
                        for (int dy = 0; dy < k; ++dy)
                            for (int dx = 0; dx < k; ++dx) {
                                const samp_t* ys = &sy[py * k + dy];
                                const samp_t* xs = &sx[px * k + dx];
                                float v = 0.0f;
                                if (ys->valid && xs->valid) {
                                    const float tl = fb[((size_t)ys->lo * Wf + xs->lo) * C + c];
                                    const float tr = fb[((size_t)ys->lo * Wf + xs->hi) * C + c];
                                    const float bl = fb[((size_t)ys->hi * Wf + xs->lo) * C + c];
                                    const float brr = fb[((size_t)ys->hi * Wf + xs->hi) * C + c];
                                    float t = (tr - tl) * xs->lerp;
                                    t = tl + t;
                                    float bo = (brr - bl) * xs->lerp;
                                    bo = bl + bo;
                                    float d = (bo - t) * ys->lerp;
                                    v = t + d;
                                }
                                if (v > best || (dy == 0 && dx == 0)) {
                                    best = v;
                                    bdy = dy;
                                    bdx = dx;
                                }
                            }
                        const samp_t* ys = &sy[py * k + bdy];
                        const samp_t* xs = &sx[px * k + bdx];
                        if (!(ys->valid && xs->valid)) continue;
                        const float gg = g[c];
                        const float dtop = (1.0f - ys->lerp) * gg;
                        const float dbot = ys->lerp * gg;
                        gb[((size_t)ys->lo * Wf + xs->lo) * C + c] += (1.0f - xs->lerp) * dtop;
                        gb[((size_t)ys->lo * Wf + xs->hi) * C + c] += xs->lerp * dtop;
                        gb[((size_t)ys->hi * Wf + xs->lo) * C + c] += (1.0f - xs->lerp) * dbot;
                        gb[((size_t)ys->hi * Wf + xs->hi) * C + c] += xs->lerp * dbot;
                    }
                }
        }
    }
}

/* ---------------------------------------------------------------------------
 * utils/training.py:7-77 generate_targets + :123-143 _get_labels_masks, one image.
 *   gt_labels [G, nc+1] one-hot (zero rows = padding), gt_boxes [G,4] relative,
 *   regions [N,4] absolute.  Outputs target_labels [N,nc+1], target_boxes
 *   [N,nc,4]; optional max_iou[N], argmax[N] (index into the *unpadded* gt
 *   list, as the reference's tf.gather sees it).
 * G' = 0 (no gt after padding removal) is undefined in the reference (T13):
 *   here every region gets max_iou = 0, no forced foreground, zero targets.
 * ------------------------------------------------------------------------- */
ORC_API void orc_generate_targets(const float* gt_labels, const float* gt_boxes, int G, int nc, const float* regions,
                                  int N, float img_h, float img_w, float fg_lo, float fg_hi, float bg_lo, float bg_hi,
                                  float* target_labels, float* target_boxes, float* max_iou_out,
                                  int32_t* argmax_out) {
    const int L = nc + 1;
    float* gabs = (float*)malloc(sizeof(float) * 4 * (size_t)(G > 0 ? G : 1));
    int* gsrc = (int*)malloc(sizeof(int) * (size_t)(G > 0 ? G : 1));
    int Gp = 0;
    for (int g = 0; g < G; ++g) { /* :43-45 remove padding rows */
        float s = 0.0f;
        for (int c = 0; c < L; ++c) s += gt_labels[(size_t)g * L + c];
        if (s != 0.0f) {
            gabs[Gp * 4 + 0] = gt_boxes[g * 4 + 0] * img_w; /* :48 to_absolute */
            gabs[Gp * 4 + 1] = gt_boxes[g * 4 + 1] * img_h;
            gabs[Gp * 4 + 2] = gt_boxes[g * 4 + 2] * img_w;
            gabs[Gp * 4 + 3] = gt_boxes[g * 4 + 3] * img_h;
            gsrc[Gp] = g;
            ++Gp;
        }
    }
    float* miou = (float*)malloc(sizeof(float) * (size_t)(N > 0 ? N : 1));
    int32_t* amax = (int32_t*)malloc(sizeof(int32_t) * (size_t)(N > 0 ? N : 1));
    float gmax = -INFINITY;
    int gmax_idx = -1;
    for (int i = 0; i < N; ++i) { /* :51-54 */
        float best = Gp > 0 ? -INFINITY : 0.0f;
        int bi = 0;
        for (int g = 0; g < Gp; ++g) {
            const float v = ref_iou(regions + (size_t)i * 4, gabs + g * 4);
            if (v > best) { /* argmax: first index on ties */
                best = v;
                bi = g;
            }
        }
        miou[i] = best;
        amax[i] = bi;
        if (Gp > 0 && best > gmax) { /* first region attaining the global max (:137) */
            gmax = best;
            gmax_idx = i;
        }
    }
    memset(target_labels, 0, sizeof(float) * (size_t)N * L);
    memset(target_boxes, 0, sizeof(float) * (size_t)N * nc * 4);
    for (int i = 0; i < N; ++i) {
        const float v = miou[i];
        const int bg = (v >= bg_lo) && (v < bg_hi);                     /* :134 */
        const int fg = ((v >= fg_lo) && (v < fg_hi)) || (i == gmax_idx); /* :139-140 */
        float* tl = target_labels + (size_t)i * L;
        if (bg) tl[0] = 1.0f;                                           /* :60,64 */
        if (fg && Gp > 0) {                                             /* :61,65 fg overrides bg */
            const float* gl = gt_labels + (size_t)gsrc[amax[i]] * L;
            for (int c = 0; c < L; ++c) tl[c] = gl[c];
        }
        /* :68-75 encode(gt[argmax], region) tiled over classes, masked by labels[:,1:] */
        if (Gp > 0) {
            float enc[4];
            int any = 0;
            for (int c = 0; c < nc; ++c) any |= (tl[1 + c] != 0.0f);
            if (any) {
                encode_one(gabs + amax[i] * 4, regions + (size_t)i * 4, enc);
                for (int c = 0; c < nc; ++c)
                    if (tl[1 + c] != 0.0f)
                        for (int k = 0; k < 4; ++k) target_boxes[((size_t)i * nc + c) * 4 + k] = enc[k];
            }
        }
        if (max_iou_out) max_iou_out[i] = v;
        if (argmax_out) argmax_out[i] = amax[i];
    }
    free(gabs);
    free(gsrc);
    free(miou);
    free(amax);
}

/* Batched driver used by the CPU baseline (tf.map_fn over images, rpn_detector.py:143-149 /
 * fast_rcnn_detector.py:113-119).  regions_batched selects [N,4] vs [B,N,4]. */
ORC_API void orc_generate_targets_batched(const float* gt_labels, const float* gt_boxes, int B, int G, int nc,
                                          const float* regions, int regions_batched, int N, float img_h, float img_w,
                                          float fg_lo, float fg_hi, float bg_lo, float bg_hi, float* target_labels,
                                          float* target_boxes) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b)
        orc_generate_targets(gt_labels + (size_t)b * G * (nc + 1), gt_boxes + (size_t)b * G * 4, G, nc,
                             regions + (regions_batched ? (size_t)b * N * 4 : 0), N, img_h, img_w, fg_lo, fg_hi, bg_lo,
                             bg_hi, target_labels + (size_t)b * N * (nc + 1), target_boxes + (size_t)b * N * nc * 4,
                             NULL, NULL);
}

// Target matching: utils/training.py:7-77 generate_targets and :123-143 _get_labels_masks, batched
// over images (the reference loops images with tf.map_fn, rpn_detector.py:143-149 /
// fast_rcnn_detector.py:113-119).  The [N,G] IoU matrix of training.py:51 is never materialised:
// each thread scans the image's ground-truth boxes from shared memory for one region.
//   match_kernel   row max / first arg-max over gt, plus the per-image global max and the FIRST region
//                  attaining it, packed as (iou_bits << 32 | ~region) and reduced with atomicMax.
//   labels_kernel  background / foreground intervals, forced foreground, labels, encode()d box targets.
#include "common.cuh"

namespace frcnn {

constexpr int kMaxGt = 1024;
constexpr int kTgtThreads = 256;

// Label row `g` of an image as generate_targets sees it.  kObj = false: the row itself (L = nc + 1 columns).
// kObj = true: the RPN's objectness labels, rpn_detector.py:141 folded in --
//   one_hot(cast(reduce_sum(gt_labels, -1), int32), depth=2): a padding row (sum 0) becomes [1, 0] and is therefore NOT
//   removed by training.py:43-45 (trap T11); a one-hot row (sum 1) becomes [0, 1]; any other integer gives [0, 0].
struct LabelRow {
    const float* row;  // the input row, Lin columns
    int Lin;
    int obj_idx;       // kObj: the one_hot index
    template <bool kObj>
    __device__ __forceinline__ static LabelRow load(const float* gt_labels, size_t row_index, int Lin) {
        LabelRow r{gt_labels + row_index * Lin, Lin, 0};
        if (kObj) {
            float s = 0.0f;
            for (int c = 0; c < Lin; ++c) s = fadd(s, r.row[c]);  // tf.reduce_sum, left to right
            r.obj_idx = (int)s;                                    // tf.cast float -> int32 truncates
        }
        return r;
    }
    template <bool kObj>
    __device__ __forceinline__ float at(int c) const {
        return kObj ? ((c == obj_idx) ? 1.0f : 0.0f) : row[c];
    }
    template <bool kObj>
    __device__ __forceinline__ float sum(int L) const {  // L = columns generate_targets sees (2 for objectness)
        float s = 0.0f;
        for (int c = 0; c < L; ++c) s = fadd(s, at<kObj>(c));
        return s;
    }
};

template <bool kObj>
__global__ void __launch_bounds__(kTgtThreads)
    match_kernel(const float* __restrict__ gt_labels, int Lin, const float4* __restrict__ gt_boxes, int G, int L,
                 const float4* __restrict__ regions, int regions_batched, int N, float img_h, float img_w,
                 float* __restrict__ max_iou, int32_t* __restrict__ argmax, unsigned long long* __restrict__ gmax,
                 int32_t* __restrict__ gt_src) {
    __shared__ float4 sgt[kMaxGt];
    __shared__ int ssrc[kMaxGt];
    __shared__ unsigned char sflag[kMaxGt];
    __shared__ int s_gp;
    __shared__ unsigned long long s_best;
    const int b = blockIdx.y;
    // training.py:43-45: drop rows whose label sum is 0, order preserved (ssrc doubles as the flag array first)
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        const float s = LabelRow::load<kObj>(gt_labels, (size_t)b * G + g, Lin).template sum<kObj>(L);
        sflag[g] = (s != 0.0f) ? 1 : 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int gp = 0;
        for (int g = 0; g < G; ++g)
            if (sflag[g]) ssrc[gp++] = g;
        s_gp = gp;
        s_best = 0ull;
    }
    __syncthreads();
    const int gp = s_gp;
    for (int g = threadIdx.x; g < gp; g += blockDim.x) {
        const float4 r = gt_boxes[(size_t)b * G + ssrc[g]];
        sgt[g] = make_float4(fmul(r.x, img_w), fmul(r.y, img_h), fmul(r.z, img_w), fmul(r.w, img_h));  // :48
        if (blockIdx.x == 0) gt_src[(size_t)b * (G + 1) + g] = ssrc[g];
    }
    __syncthreads();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long mine = 0ull;
    if (n < N) {
        const float4 reg = regions[regions_batched ? (size_t)b * N + n : (size_t)n];
        float best = gp > 0 ? -INFINITY : 0.0f;
        int bi = 0;
        for (int g = 0; g < gp; ++g) {
            const float v = ref_iou(reg, sgt[g]);  // metrics.py:183-208
            if (v > best) {                        // tf.argmax: first index on ties
                best = v;
                bi = g;
            }
        }
        max_iou[(size_t)b * N + n] = best;
        argmax[(size_t)b * N + n] = bi;
        if (gp > 0 && best >= 0.0f)
            mine = ((unsigned long long)__float_as_uint(best) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)n);
    }
    // block reduce, then one atomic per block
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long v = __shfl_xor_sync(0xffffffffu, mine, o);
        mine = v > mine ? v : mine;
    }
    if ((threadIdx.x & 31) == 0 && mine) atomicMax(&s_best, mine);
    __syncthreads();
    if (threadIdx.x == 0 && s_best) atomicMax(gmax + b, s_best);
    if (threadIdx.x == 0 && blockIdx.x == 0) gt_src[(size_t)b * (G + 1) + G] = gp;  // count in the extra slot
}

template <bool kObj>
__global__ void __launch_bounds__(kTgtThreads)
    labels_kernel(const float* __restrict__ gt_labels, int Lin, const float4* __restrict__ gt_boxes, int G, int nc,
                  const float4* __restrict__ regions, int regions_batched, int N, float img_h, float img_w, float fg_lo,
                  float fg_hi, float bg_lo, float bg_hi, const float* __restrict__ max_iou,
                  const int32_t* __restrict__ argmax, const unsigned long long* __restrict__ gmax,
                  const int32_t* __restrict__ gt_src, float* __restrict__ target_labels,
                  float4* __restrict__ target_boxes) {
    const int b = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int L = nc + 1;
    const int gp = gt_src[(size_t)b * (G + 1) + G];
    const float v = max_iou[(size_t)b * N + n];
    const unsigned long long gm = gmax[b];
    const bool forced = gp > 0 && gm != 0ull && (0xffffffffu - (uint32_t)(gm & 0xffffffffull)) == (uint32_t)n;
    const bool bg = (v >= bg_lo) && (v < bg_hi);                    // training.py:134
    const bool fg = gp > 0 && (((v >= fg_lo) && (v < fg_hi)) || forced);  // :139-140
    float* tl = target_labels + ((size_t)b * N + n) * L;
    float4* tb = target_boxes + ((size_t)b * N + n) * nc;
    if (fg) {
        const int g = gt_src[(size_t)b * (G + 1) + argmax[(size_t)b * N + n]];
        const LabelRow gl = LabelRow::load<kObj>(gt_labels, (size_t)b * G + g, Lin);
        const float4 r = gt_boxes[(size_t)b * G + g];
        const float4 gabs = make_float4(fmul(r.x, img_w), fmul(r.y, img_h), fmul(r.z, img_w), fmul(r.w, img_h));
        const float4 reg = regions[regions_batched ? (size_t)b * N + n : (size_t)n];
        const float4 enc = encode_box(gabs, reg);  // :68-69
        tl[0] = gl.at<kObj>(0);
        for (int c = 0; c < nc; ++c) {
            const float l = gl.at<kObj>(1 + c);
            tl[1 + c] = l;  // :61,65 foreground overrides background
            tb[c] = (l != 0.0f) ? enc : make_float4(0.f, 0.f, 0.f, 0.f);  // :70-75 masked by labels[:,1:]
        }
    } else {
        tl[0] = bg ? 1.0f : 0.0f;  // :60,64; neither -> all-zero (ignored)
        for (int c = 0; c < nc; ++c) {
            tl[1 + c] = 0.0f;
            tb[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

// rpn_detector.py:141 on its own (the Python helper rpn_objectness_labels): [rows, Lin] -> [rows, 2]
__global__ void objectness_kernel(const float* __restrict__ gt_labels, long long rows, int Lin, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const LabelRow r = LabelRow::load<true>(gt_labels, (size_t)i, Lin);
    out[2 * i] = r.at<true>(0);
    out[2 * i + 1] = r.at<true>(1);
}

struct TgtWs {
    float* max_iou;
    int32_t* argmax;
    unsigned long long* gmax;
    int32_t* gt_src;  // [B, G+1]: source row of each unpadded gt, count in slot G
    size_t bytes;
};
static TgtWs carve_tgt_ws(void* base, int B, int N, int G) {
    TgtWs w;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void* p = base ? static_cast<char*>(base) + off : nullptr;
        off += align_up(bytes, 256);
        return p;
    };
    w.max_iou = static_cast<float*>(take(sizeof(float) * (size_t)B * N));
    w.argmax = static_cast<int32_t*>(take(sizeof(int32_t) * (size_t)B * N));
    w.gmax = static_cast<unsigned long long*>(take(sizeof(unsigned long long) * (size_t)B));
    w.gt_src = static_cast<int32_t*>(take(sizeof(int32_t) * (size_t)B * (G + 1)));
    w.bytes = off;
    return w;
}

}  // namespace frcnn

using namespace frcnn;

extern "C" {

size_t frcnn_generate_targets_workspace_bytes(int B, int N) {
    if (B <= 0 || N < 0) return 0;
    return carve_tgt_ws(nullptr, B, N, kMaxGt).bytes;
}

static int generate_targets_impl(bool objectness, const float* gt_labels, int Lin, const float* gt_boxes, int B, int G, int nc,
                                 const float* regions, int regions_batched, int N, float img_h, float img_w, float fg_lo,
                                 float fg_hi, float bg_lo, float bg_hi, float* target_labels, float* target_boxes,
                                 float* max_iou, int32_t* argmax, void* workspace, size_t workspace_bytes,
                                 frcnn_stream_t stream) {
    FRCNN_REQUIRE(B > 0 && G >= 0 && nc > 0 && N >= 0 && Lin >= 1, FRCNN_EINVAL, "generate_targets: bad shape");
    FRCNN_REQUIRE(G <= kMaxGt, FRCNN_ELIMIT, "generate_targets: G = %d exceeds the %d limit", G, kMaxGt);
    FRCNN_REQUIRE(B <= 65535, FRCNN_ELIMIT, "generate_targets: B > 65535");
    FRCNN_REQUIRE(aligned16(gt_boxes) && aligned16(regions) && aligned16(target_boxes) && target_labels, FRCNN_EINVAL,
                  "generate_targets: pointers must be 16-byte aligned");
    FRCNN_REQUIRE(workspace && workspace_bytes >= frcnn_generate_targets_workspace_bytes(B, N), FRCNN_EWORKSPACE,
                  "generate_targets: workspace too small");
    if (N == 0) return 0;
    cudaStream_t st = as_stream(stream);
    TgtWs w = carve_tgt_ws(workspace, B, N, G);
    float* mi = max_iou ? max_iou : w.max_iou;
    int32_t* am = argmax ? argmax : w.argmax;
    FRCNN_CUDA(cudaMemsetAsync(w.gmax, 0, sizeof(unsigned long long) * (size_t)B, st));
    const dim3 grid((N + kTgtThreads - 1) / kTgtThreads, B);
    if (objectness) {
        match_kernel<true><<<grid, kTgtThreads, 0, st>>>(gt_labels, Lin, (const float4*)gt_boxes, G, nc + 1, (const float4*)regions,
                                                         regions_batched, N, img_h, img_w, mi, am, w.gmax, w.gt_src);
        FRCNN_LAUNCH_CHECK("match_kernel<objectness>");
        labels_kernel<true><<<grid, kTgtThreads, 0, st>>>(gt_labels, Lin, (const float4*)gt_boxes, G, nc, (const float4*)regions,
                                                          regions_batched, N, img_h, img_w, fg_lo, fg_hi, bg_lo, bg_hi, mi, am,
                                                          w.gmax, w.gt_src, target_labels, (float4*)target_boxes);
        FRCNN_LAUNCH_CHECK("labels_kernel<objectness>");
        return 0;
    }
    match_kernel<false><<<grid, kTgtThreads, 0, st>>>(gt_labels, Lin, (const float4*)gt_boxes, G, nc + 1, (const float4*)regions,
                                                      regions_batched, N, img_h, img_w, mi, am, w.gmax, w.gt_src);
    FRCNN_LAUNCH_CHECK("match_kernel");
    labels_kernel<false><<<grid, kTgtThreads, 0, st>>>(gt_labels, Lin, (const float4*)gt_boxes, G, nc, (const float4*)regions,
                                                       regions_batched, N, img_h, img_w, fg_lo, fg_hi, bg_lo, bg_hi, mi, am,
                                                       w.gmax, w.gt_src, target_labels, (float4*)target_boxes);
    FRCNN_LAUNCH_CHECK("labels_kernel");
    return 0;
}

int frcnn_generate_targets(const float* gt_labels, const float* gt_boxes, int B, int G, int nc, const float* regions,
                           int regions_batched, int N, float img_h, float img_w, float fg_lo, float fg_hi, float bg_lo,
                           float bg_hi, float* target_labels, float* target_boxes, float* max_iou, int32_t* argmax,
                           void* workspace, size_t workspace_bytes, frcnn_stream_t stream) {
    return generate_targets_impl(false, gt_labels, nc + 1, gt_boxes, B, G, nc, regions, regions_batched, N, img_h, img_w,
                                 fg_lo, fg_hi, bg_lo, bg_hi, target_labels, target_boxes, max_iou, argmax, workspace,
                                 workspace_bytes, stream);
}

int frcnn_generate_targets_objectness(const float* gt_labels, const float* gt_boxes, int B, int G, int num_classes_plus_1,
                                      const float* regions, int regions_batched, int N, float img_h, float img_w,
                                      float fg_lo, float fg_hi, float bg_lo, float bg_hi, float* target_labels,
                                      float* target_boxes, float* max_iou, int32_t* argmax, void* workspace,
                                      size_t workspace_bytes, frcnn_stream_t stream) {
    return generate_targets_impl(true, gt_labels, num_classes_plus_1, gt_boxes, B, G, 1, regions, regions_batched, N, img_h,
                                 img_w, fg_lo, fg_hi, bg_lo, bg_hi, target_labels, target_boxes, max_iou, argmax, workspace,
                                 workspace_bytes, stream);
}

int frcnn_objectness_labels(const float* gt_labels, int64_t rows, int num_classes_plus_1, float* out, frcnn_stream_t stream) {
    FRCNN_REQUIRE(rows >= 0 && num_classes_plus_1 >= 1 && (rows == 0 || (gt_labels && out)), FRCNN_EINVAL,
                  "objectness_labels: bad arguments");
    if (rows == 0) return 0;
    objectness_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, as_stream(stream)>>>(gt_labels, rows, num_classes_plus_1, out);
    FRCNN_LAUNCH_CHECK("objectness_kernel");
    return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// Average-precision matching: utils/metrics.py:57-84 AveragePrecision.update_state_single_image, batched over images
// (SURVEY 8(f) rank 4).  Ground-truth rows whose coordinates sum to 0 are padding (:67-68); every remaining gt box
// marks the prediction of highest IoU with it (tf.math.argmax: the first maximum) as a true positive when that IoU
// exceeds the threshold (:79-83).  IoU is the reference's own (utils/metrics.py:183-208), never materialised as a
// matrix.  One CTA per image, one warp per gt box.
// ------------------------------------------------------------------------------------------------
namespace frcnn {

__global__ void __launch_bounds__(256)
    ap_match_kernel(const float4* __restrict__ gt_boxes, int G, const float4* __restrict__ pred_boxes, int M, float iou_thr,
                    int32_t* __restrict__ true_pos, int32_t* __restrict__ num_gt) {
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    __shared__ int count;
    if (threadIdx.x == 0) count = 0;
    int32_t* tp = true_pos + (size_t)b * M;
    for (int i = threadIdx.x; i < M; i += blockDim.x) tp[i] = 0;
    __syncthreads();
    const float4* gt = gt_boxes + (size_t)b * G;
    const float4* pr = pred_boxes + (size_t)b * M;
    for (int g = warp; g < G; g += nwarps) {
        const float4 gb = gt[g];
        if (fadd(fadd(fadd(gb.x, gb.y), gb.z), gb.w) == 0.0f) continue;  // padding row
        if (lane == 0) atomicAdd(&count, 1);
        float best = -1.0f;
        int arg = 0x7fffffff;
        for (int i = lane; i < M; i += 32) {
            const float v = ref_iou(gb, pr[i]);
            if (v > best) {  // ascending i per lane: keeps the first maximum of the lane
                best = v;
                arg = i;
            }
        }
        for (int d = 16; d > 0; d >>= 1) {  // warp arg-max, lower index on ties (tf.math.argmax)
            const float ob = __shfl_xor_sync(0xffffffffu, best, d);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, d);
            if (ob > best || (ob == best && oa < arg)) {
                best = ob;
                arg = oa;
            }
        }
        if (lane == 0 && M > 0 && best > iou_thr) tp[arg] = 1;
    }
    __syncthreads();
    if (threadIdx.x == 0 && num_gt) num_gt[b] = count;
}

}  // namespace frcnn

extern "C" int frcnn_ap_match(const float* gt_boxes, int B, int G, const float* pred_boxes, int M, float iou_threshold,
                              int32_t* true_positives, int32_t* num_gt, frcnn_stream_t stream) {
    FRCNN_REQUIRE(B > 0 && G >= 0 && M >= 0 && B <= 65535, FRCNN_EINVAL, "ap_match: bad shape");
    FRCNN_REQUIRE(aligned16(gt_boxes) && aligned16(pred_boxes) && (true_positives || M == 0), FRCNN_EINVAL,
                  "ap_match: box tensors must be 16-byte aligned");
    frcnn::ap_match_kernel<<<B, 256, 0, frcnn::as_stream(stream)>>>((const float4*)gt_boxes, G, (const float4*)pred_boxes, M,
                                                                     iou_threshold, true_positives, num_gt);
    FRCNN_LAUNCH_CHECK("ap_match_kernel");
    return 0;
}

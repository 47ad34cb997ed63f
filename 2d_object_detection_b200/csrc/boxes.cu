// Box utilities: utils/boxes.py, utils/metrics.py:136-208 and the anchor helpers of
// models/detectors/rpn_detector.py:162-230.  All HBM-bound elementwise work: one float4 per box.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace frcnn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return static_cast<int>(e);
}

constexpr int kThreads = 256;
static inline unsigned grid_for(int64_t n, int threads = kThreads) { return static_cast<unsigned>((n + threads - 1) / threads); }

__global__ void clip_kernel(const float4* __restrict__ in, int64_t n, float4 lo, float4 hi, float4* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 b = in[i];
    out[i] = make_float4(fmaxf(fminf(b.x, hi.x), lo.x), fmaxf(fminf(b.y, hi.y), lo.y), fmaxf(fminf(b.z, hi.z), lo.z),
                         fmaxf(fminf(b.w, hi.w), lo.w));
}

template <bool kEncode>
__global__ void codec_kernel(const float4* __restrict__ in, const float4* __restrict__ refs, int64_t n, int64_t inner,
                             int64_t period, float4* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 r = refs[(i / inner) % period];
    out[i] = kEncode ? encode_box(in[i], r) : decode_box(in[i], r);
}

template <bool kDivide>
__global__ void scale_kernel(const float4* __restrict__ in, int64_t n, float w, float h, float4* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 b = in[i];
    out[i] = kDivide ? make_float4(fdiv(b.x, w), fdiv(b.y, h), fdiv(b.z, w), fdiv(b.w, h))
                     : make_float4(fmul(b.x, w), fmul(b.y, h), fmul(b.z, w), fmul(b.w, h));
}

struct AnchorShapes {
    float w[64];
    float h[64];
};

// anchor index (i*Wf + j)*A + k; centres at j*stride_w, i*stride_h (rpn_detector.py:186-199).
__global__ void anchors_kernel(int Hf, int Wf, int A, AnchorShapes sh, float stride_h, float stride_w,
                               float4* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t n = (int64_t)Hf * Wf * A;
    if (i >= n) return;
    const int k = (int)(i % A);
    const int64_t loc = i / A;
    const float cx = fmul((float)(loc % Wf), stride_w);
    const float cy = fmul((float)(loc / Wf), stride_h);
    const float hw = fmul(0.5f, sh.w[k]), hh = fmul(0.5f, sh.h[k]);
    out[i] = make_float4(fsub(cx, hw), fsub(cy, hh), fadd(cx, hw), fadd(cy, hh));
}

// Ordered stream compaction by one CTA (one-off at model construction; N <= a few 100k).
__global__ void __launch_bounds__(1024) valid_idx_kernel(const float4* __restrict__ anchors, int64_t n, float W, float H,
                                                         int64_t* __restrict__ idx_out, int32_t* __restrict__ count) {
    __shared__ int warp_tot[32];
    __shared__ int base_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) base_s = 0;
    __syncthreads();
    for (int64_t start = 0; start < n; start += blockDim.x) {
        const int64_t i = start + threadIdx.x;
        bool ok = false;
        if (i < n) {
            const float4 a = anchors[i];
            ok = a.x >= 0.0f && a.y >= 0.0f && a.z <= W && a.w <= H;
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) warp_tot[warp] = __popc(m);
        __syncthreads();
        int off = base_s;
        for (int w = 0; w < warp; ++w) off += warp_tot[w];
        if (ok) idx_out[off + __popc(m & ((1u << lane) - 1u))] = i;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += warp_tot[w];
            base_s += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = base_s;
}

__global__ void gather_rows_kernel(const float* __restrict__ in, int64_t batch, int64_t n, int64_t row, const int64_t* __restrict__ idx,
                                   int64_t m, float* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t total = batch * m * row;
    if (i >= total) return;
    const int64_t f = i % row;
    const int64_t r = (i / row) % m;
    const int64_t b = i / (row * m);
    out[i] = in[(b * n + idx[r]) * row + f];
}

__global__ void area_kernel(const float4* __restrict__ b, int64_t n, float* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = b[i];
    out[i] = fmul(fsub(a.z, a.x), fsub(a.w, a.y));
}

// mode 0: iou, 1: intersection.  One thread per (i, j); j fastest so writes are coalesced.
template <int kMode>
__global__ void pairwise_kernel(const float4* __restrict__ b1, int64_t n, const float4* __restrict__ b2, int64_t m,
                                float* __restrict__ out) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n * m) return;
    const float4 a = b1[t / m];
    const float4 b = b2[t % m];
    if (kMode == 0) {
        out[t] = ref_iou(a, b);
    } else {
        const float dw = fsub(fminf(a.z, b.z), fmaxf(a.x, b.x));
        const float dh = fsub(fminf(a.w, b.w), fmaxf(a.y, b.y));
        out[t] = fmul(fmaxf(0.0f, dw), fmaxf(0.0f, dh));
    }
}

template <int kMode>
__global__ void elementwise_kernel(const float4* __restrict__ b1, const float4* __restrict__ b2, int64_t n,
                                   float* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 a = b1[i], b = b2[i];
    if (kMode == 0) {
        out[i] = ref_iou(a, b);
    } else {
        const float dw = fsub(fminf(a.z, b.z), fmaxf(a.x, b.x));
        const float dh = fsub(fminf(a.w, b.w), fmaxf(a.y, b.y));
        out[i] = fmul(fmaxf(0.0f, dw), fmaxf(0.0f, dh));
    }
}

}  // namespace frcnn

using namespace frcnn;

extern "C" {

int frcnn_version(void) { return FRCNN_VERSION; }
const char* frcnn_last_error(void) { return g_err; }

int frcnn_clip_to_window(const float* boxes, int64_t n, float x0, float y0, float x1, float y1, float* out,
                         frcnn_stream_t stream) {
    FRCNN_REQUIRE(n >= 0 && aligned16(boxes) && aligned16(out), FRCNN_EINVAL, "clip_to_window: bad size/alignment");
    if (n == 0) return 0;
    clip_kernel<<<grid_for(n), kThreads, 0, as_stream(stream)>>>((const float4*)boxes, n, make_float4(x0, y0, x0, y0),
                                                                  make_float4(x1, y1, x1, y1), (float4*)out);
    FRCNN_LAUNCH_CHECK("clip_to_window");
    return 0;
}

int frcnn_decode(const float* deltas, const float* refs, int64_t n, int64_t inner, int64_t period, float* out,
                 frcnn_stream_t stream) {
    FRCNN_REQUIRE(n >= 0 && inner > 0 && period > 0 && aligned16(deltas) && aligned16(refs) && aligned16(out),
                  FRCNN_EINVAL, "decode: bad size/alignment");
    if (n == 0) return 0;
    codec_kernel<false><<<grid_for(n), kThreads, 0, as_stream(stream)>>>((const float4*)deltas, (const float4*)refs, n,
                                                                         inner, period, (float4*)out);
    FRCNN_LAUNCH_CHECK("decode");
    return 0;
}

int frcnn_encode(const float* boxes, const float* refs, int64_t n, int64_t inner, int64_t period, float* out,
                 frcnn_stream_t stream) {
    FRCNN_REQUIRE(n >= 0 && inner > 0 && period > 0 && aligned16(boxes) && aligned16(refs) && aligned16(out),
                  FRCNN_EINVAL, "encode: bad size/alignment");
    if (n == 0) return 0;
    codec_kernel<true><<<grid_for(n), kThreads, 0, as_stream(stream)>>>((const float4*)boxes, (const float4*)refs, n,
                                                                        inner, period, (float4*)out);
    FRCNN_LAUNCH_CHECK("encode");
    return 0;
}

int frcnn_scale_boxes(const float* boxes, int64_t n, float image_height, float image_width, int divide, float* out,
                      frcnn_stream_t stream) {
    FRCNN_REQUIRE(n >= 0 && aligned16(boxes) && aligned16(out), FRCNN_EINVAL, "scale_boxes: bad size/alignment");
    if (n == 0) return 0;
    if (divide)
        scale_kernel<true><<<grid_for(n), kThreads, 0, as_stream(stream)>>>((const float4*)boxes, n, image_width,
                                                                            image_height, (float4*)out);
    else
        scale_kernel<false><<<grid_for(n), kThreads, 0, as_stream(stream)>>>((const float4*)boxes, n, image_width,
                                                                             image_height, (float4*)out);
    FRCNN_LAUNCH_CHECK("scale_boxes");
    return 0;
}

int frcnn_generate_anchors(int Hf, int Wf, const float* scales, int ns, const float* ratios, int nr, float base_h,
                           float base_w, float stride_h, float stride_w, float* out, frcnn_stream_t stream) {
    FRCNN_REQUIRE(Hf > 0 && Wf > 0 && ns > 0 && nr > 0 && ns * nr <= 64 && scales && ratios && aligned16(out),
                  FRCNN_EINVAL, "generate_anchors: need 0 < ns*nr <= 64 and aligned output");
    AnchorShapes sh;
    // rpn_detector.py:178-184: meshgrid (xy) => k = r*ns + s; h = s/sqrt(r)*base_h, w = s*sqrt(r)*base_w.
    // Evaluated on the host in fp32 (IEEE sqrt/div/mul are correctly rounded everywhere).
    for (int r = 0; r < nr; ++r)
        for (int s = 0; s < ns; ++s) {
            volatile float rs = sqrtf(ratios[r]);
            volatile float h = scales[s] / rs;
            h = h * base_h;
            volatile float w = scales[s] * rs;
            w = w * base_w;
            sh.h[r * ns + s] = h;
            sh.w[r * ns + s] = w;
        }
    const int64_t n = (int64_t)Hf * Wf * ns * nr;
    anchors_kernel<<<grid_for(n), kThreads, 0, as_stream(stream)>>>(Hf, Wf, ns * nr, sh, stride_h, stride_w, (float4*)out);
    FRCNN_LAUNCH_CHECK("generate_anchors");
    return 0;
}

int frcnn_valid_anchor_indices(const float* anchors, int64_t n, float image_height, float image_width, int64_t* idx_out,
                               int32_t* count_out, frcnn_stream_t stream) {
    FRCNN_REQUIRE(n >= 0 && aligned16(anchors) && idx_out && count_out, FRCNN_EINVAL, "valid_anchor_indices: bad args");
    valid_idx_kernel<<<1, 1024, 0, as_stream(stream)>>>((const float4*)anchors, n, image_width, image_height, idx_out,
                                                         count_out);
    FRCNN_LAUNCH_CHECK("valid_anchor_indices");
    return 0;
}

int frcnn_gather_rows(const float* in, int64_t batch, int64_t n, int64_t row, const int64_t* idx, int64_t m, float* out,
                      frcnn_stream_t stream) {
    FRCNN_REQUIRE(batch >= 0 && n >= 0 && row > 0 && m >= 0, FRCNN_EINVAL, "gather_rows: bad sizes");
    const int64_t total = batch * m * row;
    if (total == 0) return 0;
    gather_rows_kernel<<<grid_for(total), kThreads, 0, as_stream(stream)>>>(in, batch, n, row, idx, m, out);
    FRCNN_LAUNCH_CHECK("gather_rows");
    return 0;
}

int frcnn_area(const float* boxes, int64_t n, float* out, frcnn_stream_t stream) {
    FRCNN_REQUIRE(n >= 0 && aligned16(boxes), FRCNN_EINVAL, "area: bad size/alignment");
    if (n == 0) return 0;
    area_kernel<<<grid_for(n), kThreads, 0, as_stream(stream)>>>((const float4*)boxes, n, out);
    FRCNN_LAUNCH_CHECK("area");
    return 0;
}

int frcnn_intersection_pairwise(const float* b1, int64_t n, const float* b2, int64_t m, float* out,
                                frcnn_stream_t stream) {
    FRCNN_REQUIRE(n >= 0 && m >= 0 && aligned16(b1) && aligned16(b2), FRCNN_EINVAL, "intersection: bad size/alignment");
    if (n * m == 0) return 0;
    pairwise_kernel<1><<<grid_for(n * m), kThreads, 0, as_stream(stream)>>>((const float4*)b1, n, (const float4*)b2, m, out);
    FRCNN_LAUNCH_CHECK("intersection_pairwise");
    return 0;
}

int frcnn_iou_pairwise(const float* b1, int64_t n, const float* b2, int64_t m, float* out, frcnn_stream_t stream) {
    FRCNN_REQUIRE(n >= 0 && m >= 0 && aligned16(b1) && aligned16(b2), FRCNN_EINVAL, "iou: bad size/alignment");
    if (n * m == 0) return 0;
    pairwise_kernel<0><<<grid_for(n * m), kThreads, 0, as_stream(stream)>>>((const float4*)b1, n, (const float4*)b2, m, out);
    FRCNN_LAUNCH_CHECK("iou_pairwise");
    return 0;
}

int frcnn_iou_elementwise(const float* b1, const float* b2, int64_t n, float* out, frcnn_stream_t stream) {
    FRCNN_REQUIRE(n >= 0 && aligned16(b1) && aligned16(b2), FRCNN_EINVAL, "iou: bad size/alignment");
    if (n == 0) return 0;
    elementwise_kernel<0><<<grid_for(n), kThreads, 0, as_stream(stream)>>>((const float4*)b1, (const float4*)b2, n, out);
    FRCNN_LAUNCH_CHECK("iou_elementwise");
    return 0;
}

int frcnn_intersection_elementwise(const float* b1, const float* b2, int64_t n, float* out, frcnn_stream_t stream) {
    FRCNN_REQUIRE(n >= 0 && aligned16(b1) && aligned16(b2), FRCNN_EINVAL, "intersection: bad size/alignment");
    if (n == 0) return 0;
    elementwise_kernel<1><<<grid_for(n), kThreads, 0, as_stream(stream)>>>((const float4*)b1, (const float4*)b2, n, out);
    FRCNN_LAUNCH_CHECK("intersection_elementwise");
    return 0;
}

}  // extern "C"

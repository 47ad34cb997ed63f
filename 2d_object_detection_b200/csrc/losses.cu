// Losses on the sampled rows: utils/losses.py:4-43, evaluated at models/faster_rcnn.py:87-101 on the outputs of
// get_training_samples, forward and backward in one launch each.
//
//   ClassificationLoss = tf.keras.losses.CategoricalCrossentropy() on probabilities:
//       q = p / sum(p);  c = clip(q, 1e-7, 1 - 1e-7);  row loss = -sum_j t_j log c_j;  mean over all B*S rows.
//   RegressionLoss = Huber(delta = 1, reduction NONE) on the rows whose target box is non-zero
//       (tf.where(reduce_sum(target, -1) != 0)): row value = mean over the 4 coordinates of
//       (|e| <= 1 ? 0.5 e^2 : |e| - 0.5), e = pred - target; loss = SUM over the kept rows (the reference's
//       reduce_sum(axis=-1) over the row axis followed by a reduce_mean of a scalar).
//
// One CTA; every thread accumulates a strided subset of rows in a fixed order and the partial sums are combined by a
// shared-memory tree, so the result is deterministic run to run.  The reduction ORDER differs from Eigen's (which is
// unspecified), hence parity within 1e-5 relative, as stated for floating point.
#include "common.cuh"

namespace frcnn {

namespace {

constexpr int kLossThreads = 1024;
constexpr float kEps = 1e-7f;  // tf.keras.backend.epsilon()

__device__ float block_sum(float v, float* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int d = 16; d > 0; d >>= 1) v = fadd(v, __shfl_down_sync(0xffffffffu, v, d));
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < (blockDim.x >> 5) ? scratch[lane] : 0.0f;
        for (int d = 16; d > 0; d >>= 1) v = fadd(v, __shfl_down_sync(0xffffffffu, v, d));
        if (lane == 0) scratch[0] = v;
    }
    __syncthreads();
    return scratch[0];
}

// losses[0] = classification, losses[1] = regression, losses[2] = number of regression rows kept
__global__ void __launch_bounds__(kLossThreads)
    sampled_losses_kernel(const float* __restrict__ t_labels, const float* __restrict__ p_scores, int M, int L,
                          const float4* __restrict__ t_boxes, const float4* __restrict__ p_boxes, int K, float g_cls,
                          float g_reg, float* __restrict__ losses, float* __restrict__ d_scores, float4* __restrict__ d_boxes) {
    __shared__ float scratch[32];
    float cls = 0.0f;
    const float inv_m = fdiv(1.0f, (float)max(M, 1));
    for (int r = threadIdx.x; r < M; r += kLossThreads) {
        const float* t = t_labels + (size_t)r * L;
        const float* p = p_scores + (size_t)r * L;
        float s = 0.0f;
        for (int j = 0; j < L; ++j) s = fadd(s, p[j]);
        float row = 0.0f, gp = 0.0f;  // gp = sum_j g_j p_j with g_j = dLoss/dc_j through the clip
        for (int j = 0; j < L; ++j) {
            const float q = fdiv(p[j], s);
            const float c = fminf(fmaxf(q, kEps), 1.0f - kEps);
            row = fsub(row, fmul(t[j], logf(c)));
            if (d_scores && q >= kEps && q <= 1.0f - kEps) gp = fadd(gp, fmul(fdiv(-t[j], c), p[j]));
        }
        cls = fadd(cls, row);
        if (d_scores) {  // d q_j / d p_k = (delta_jk s - p_j) / s^2
            const float scale = fmul(g_cls, inv_m);
            for (int k = 0; k < L; ++k) {
                const float q = fdiv(p[k], s);
                const float c = fminf(fmaxf(q, kEps), 1.0f - kEps);
                const float gk = (q >= kEps && q <= 1.0f - kEps) ? fdiv(-t[k], c) : 0.0f;
                d_scores[(size_t)r * L + k] = fmul(scale, fdiv(fsub(fmul(gk, s), gp), fmul(s, s)));
            }
        }
    }
    float reg = 0.0f, kept = 0.0f;
    for (int r = threadIdx.x; r < K; r += kLossThreads) {
        const float4 t = t_boxes[r], p = p_boxes[r];
        const float tsum = fadd(fadd(fadd(t.x, t.y), t.z), t.w);
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tsum != 0.0f) {
            const float e[4] = {fsub(p.x, t.x), fsub(p.y, t.y), fsub(p.z, t.z), fsub(p.w, t.w)};
            float h = 0.0f, ge[4];
            for (int j = 0; j < 4; ++j) {
                const float a = fabsf(e[j]);
                const bool quad = a <= 1.0f;
                h = fadd(h, quad ? fmul(0.5f, fmul(e[j], e[j])) : fsub(a, 0.5f));
                ge[j] = quad ? e[j] : (e[j] > 0.0f ? 1.0f : -1.0f);
            }
            reg = fadd(reg, fmul(h, 0.25f));
            kept += 1.0f;
            const float scale = fmul(g_reg, 0.25f);
            g = make_float4(fmul(scale, ge[0]), fmul(scale, ge[1]), fmul(scale, ge[2]), fmul(scale, ge[3]));
        }
        if (d_boxes) d_boxes[r] = g;
    }
    const float cls_sum = block_sum(cls, scratch);
    const float reg_sum = block_sum(reg, scratch);
    const float kept_sum = block_sum(kept, scratch);
    if (threadIdx.x == 0) {
        losses[0] = fmul(cls_sum, inv_m);
        losses[1] = reg_sum;
        losses[2] = kept_sum;
    }
}

}  // namespace

}  // namespace frcnn

using namespace frcnn;

extern "C" int frcnn_sampled_losses(const float* target_labels, const float* pred_scores, int64_t rows, int num_classes_plus_1,
                                    const float* target_boxes, const float* pred_boxes, int64_t box_rows, float grad_cls,
                                    float grad_reg, float* losses, float* d_pred_scores, float* d_pred_boxes,
                                    frcnn_stream_t stream) {
    FRCNN_REQUIRE(rows >= 0 && box_rows >= 0 && rows < (1ll << 31) && box_rows < (1ll << 31) && num_classes_plus_1 >= 1,
                  FRCNN_EINVAL, "sampled_losses: bad shape");
    FRCNN_REQUIRE(losses && (rows == 0 || (target_labels && pred_scores)) && (box_rows == 0 || (target_boxes && pred_boxes)),
                  FRCNN_EINVAL, "sampled_losses: missing tensor");
    FRCNN_REQUIRE(aligned16(target_boxes) && aligned16(pred_boxes) && aligned16(d_pred_boxes), FRCNN_EINVAL,
                  "sampled_losses: box tensors must be 16-byte aligned");
    sampled_losses_kernel<<<1, kLossThreads, 0, as_stream(stream)>>>(
        target_labels, pred_scores, (int)rows, num_classes_plus_1, (const float4*)target_boxes, (const float4*)pred_boxes,
        (int)box_rows, grad_cls, grad_reg, losses, d_pred_scores, (float4*)d_pred_boxes);
    FRCNN_LAUNCH_CHECK("sampled_losses_kernel");
    return 0;
}

// Training-sample selection, batched: utils/training.py:80-120 get_sample_indices (called per image through tf.map_fn
// at models/detectors/rpn_detector.py:151-153 and fast_rcnn_detector.py:121-123) fused with the four
// tf.gather(batch_dims=1) that follow it (rpn_detector.py:155-160, fast_rcnn_detector.py:125-130).
//
//   foreground = regions whose label row is non-zero and whose background column is 0   (ascending index, tf.where)
//   background = regions whose label row is non-zero and whose background column is 1
//   n_fg = min(#foreground, round(num_samples * foreground_proportion));  n_bg = num_samples - n_fg
//   take n_fg foregrounds WITHOUT replacement (tf.random.shuffle, then a prefix) and n_bg backgrounds WITH replacement
//   (tf.random.uniform indices into the background list); output = [foregrounds..., backgrounds...].
//
// The reference draws from TensorFlow's unseeded global generator, so parity is distributional, not bit-wise
// (SURVEY 8(f) rank 2).  Here every draw is a pure function of (seed, image, position) through Philox4x32-10:
//   * each foreground gets a 32-bit key; the n_fg smallest (key, list position) pairs, in key order, are a uniformly
//     random n_fg-prefix of a uniformly random permutation -- exactly "shuffle, then slice";
//   * background draw t is floor(u32 * #background / 2^32).
// One CTA per image: a block-wide stream compaction writes the two index lists to the workspace (ascending, like
// tf.where), the foreground keys are selected with a bitonic sort in shared memory (in chunks if there are more
// than 4096 of them), then the same CTA gathers the sampled rows of the four tensors.
#include "common.cuh"

namespace frcnn {

namespace {

constexpr int kSampThreads = 1024;
constexpr int kSortCap = 4096;  // composite keys sorted per pass (32 KB of shared memory)

struct Philox {
    uint32_t k0, k1;
    __device__ __forceinline__ uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            const uint32_t n0 = hi1 ^ c1 ^ a, n2 = hi0 ^ c3 ^ b;
            c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
            a += 0x9E3779B9u;
            b += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};

// bitonic sort of n (power of two) 64-bit keys in shared memory, ascending
__device__ void bitonic_sort_u64(unsigned long long* keys, int n) {
    for (int size = 2; size <= n; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < n / 2; t += blockDim.x) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long a = keys[lo], b = keys[hi];
                if ((a > b) == up) {
                    keys[lo] = b;
                    keys[hi] = a;
                }
            }
        }
    __syncthreads();
}

__global__ void __launch_bounds__(kSampThreads)
    sample_gather_kernel(const float* __restrict__ target_labels, int N, int L, int num_samples, int fg_quota, uint32_t seed_lo,
                         uint32_t seed_hi, int* __restrict__ ws, long long* __restrict__ indices, int* __restrict__ counts,
                         const float* __restrict__ pred_scores, const float* __restrict__ target_boxes,
                         const float* __restrict__ pred_boxes, float* __restrict__ out_labels, float* __restrict__ out_scores,
                         float* __restrict__ out_tboxes, float* __restrict__ out_pboxes) {
    __shared__ unsigned long long keys[kSortCap];
    __shared__ int warp_fg[32], warp_bg[32];
    __shared__ int base_fg, base_bg;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* tl = target_labels + (size_t)b * N * L;
    int* fg_list = ws + (size_t)b * 2 * N;
    int* bg_list = fg_list + N;
    if (tid == 0) base_fg = base_bg = 0;
    __syncthreads();
    // ---- stream compaction into the two ascending lists
    for (int start = 0; start < N; start += kSampThreads) {
        const int i = start + tid;
        bool is_fg = false, is_bg = false;
        if (i < N) {
            float sum = 0.0f;
            for (int c = 0; c < L; ++c) sum = fadd(sum, tl[(size_t)i * L + c]);  // tf.reduce_sum order, left to right
            const float c0 = tl[(size_t)i * L];
            is_fg = sum != 0.0f && c0 == 0.0f;
            is_bg = sum != 0.0f && c0 == 1.0f;
        }
        const unsigned mf = __ballot_sync(0xffffffffu, is_fg), mb = __ballot_sync(0xffffffffu, is_bg);
        if (lane == 0) {
            warp_fg[warp] = __popc(mf);
            warp_bg[warp] = __popc(mb);
        }
        __syncthreads();
        if (warp == 0) {  // exclusive scan of the 32 warp counts
            int vf = warp_fg[lane], vb = warp_bg[lane];
            const int tf_ = vf, tb_ = vb;
            for (int d = 1; d < 32; d <<= 1) {
                const int uf = __shfl_up_sync(0xffffffffu, vf, d), ub = __shfl_up_sync(0xffffffffu, vb, d);
                if (lane >= d) {
                    vf += uf;
                    vb += ub;
                }
            }
            warp_fg[lane] = base_fg + vf - tf_;
            warp_bg[lane] = base_bg + vb - tb_;
            __syncwarp();
            if (lane == 31) {
                base_fg += vf;
                base_bg += vb;
            }
        }
        __syncthreads();
        const unsigned lt = (1u << lane) - 1u;
        if (is_fg) fg_list[warp_fg[warp] + __popc(mf & lt)] = i;
        if (is_bg) bg_list[warp_bg[warp] + __popc(mb & lt)] = i;
        __syncthreads();
    }
    const int n_fg_all = base_fg, n_bg_all = base_bg;
    const int n_fg = min(n_fg_all, fg_quota);
    const int n_bg = num_samples - n_fg;
    if (tid == 0 && counts) {
        counts[2 * b] = n_fg_all;
        counts[2 * b + 1] = n_bg_all;
    }
    const Philox rng{seed_lo, seed_hi};
    long long* out_idx = indices + (size_t)b * num_samples;
    // ---- foregrounds: the n_fg smallest (key, position) pairs in key order.  Chunked: the best n_fg so far stay at the
    // front of the sort buffer and compete with the next kSortCap - n_fg candidates.
    if (n_fg > 0) {
        int kept = 0, next = 0;  // kept: valid keys at the front of the buffer; next: first unread list position
        while (next < n_fg_all) {
            const int take = min(kSortCap - kept, n_fg_all - next);
            for (int t = tid; t < take; t += kSampThreads) {
                const int j = next + t;
                const uint4 r = rng((uint32_t)j, (uint32_t)b, 0u, 0u);
                keys[kept + t] = ((unsigned long long)r.x << 32) | (uint32_t)j;
            }
            int n = kept + take, pow2 = 1;
            while (pow2 < n) pow2 <<= 1;
            for (int t = n + tid; t < pow2; t += kSampThreads) keys[t] = ~0ull;
            bitonic_sort_u64(keys, pow2);
            kept = min(n, n_fg);
            next += take;
        }
        for (int t = tid; t < n_fg; t += kSampThreads) out_idx[t] = fg_list[(uint32_t)keys[t]];
    }
    // ---- backgrounds: uniform draws with replacement
    for (int t = tid; t < n_bg; t += kSampThreads) {
        long long v = -1;  // no background to draw from: the reference's tf.random.uniform(maxval=0) fails; flagged to the host
        if (n_bg_all > 0) {
            const uint4 r = rng((uint32_t)t, (uint32_t)b, 1u, 0u);
            v = bg_list[__umulhi(r.x, (uint32_t)n_bg_all)];
        }
        out_idx[n_fg + t] = v;
    }
    __syncthreads();
    // ---- the four gathers (tf.gather(batch_dims=1)); one row = L or 4 * (L - 1) floats
    const int nc4 = 4 * (L - 1);
    for (int e = tid; e < num_samples * L; e += kSampThreads) {
        const int s = e / L, c = e - s * L;
        const long long src = out_idx[s];
        if (src < 0) continue;
        const size_t o = ((size_t)b * num_samples + s) * L + c, in = ((size_t)b * N + src) * L + c;
        if (out_labels) out_labels[o] = tl[(size_t)src * L + c];
        if (out_scores) out_scores[o] = pred_scores[in];
    }
    for (int e = tid; e < num_samples * nc4; e += kSampThreads) {
        const int s = e / nc4, c = e - s * nc4;
        const long long src = out_idx[s];
        if (src < 0) continue;
        const size_t o = ((size_t)b * num_samples + s) * nc4 + c, in = ((size_t)b * N + src) * nc4 + c;
        if (out_tboxes) out_tboxes[o] = target_boxes[in];
        if (out_pboxes) out_pboxes[o] = pred_boxes[in];
    }
}

}  // namespace

}  // namespace frcnn

using namespace frcnn;

extern "C" {

size_t frcnn_sample_indices_workspace_bytes(int B, int N) { return sizeof(int) * 2 * (size_t)(B > 0 ? B : 0) * (size_t)(N > 0 ? N : 0) + 16; }

int frcnn_sample_and_gather(const float* target_labels, int B, int N, int num_classes_plus_1, int num_samples,
                            double foreground_proportion, uint64_t seed, int64_t* indices, int32_t* counts,
                            const float* pred_scores, const float* target_boxes, const float* pred_boxes, float* out_labels,
                            float* out_scores, float* out_target_boxes, float* out_pred_boxes, void* workspace,
                            size_t workspace_bytes, frcnn_stream_t stream) {
    FRCNN_REQUIRE(B > 0 && N > 0 && num_classes_plus_1 >= 2 && num_samples > 0, FRCNN_EINVAL, "sample_and_gather: bad shape");
    FRCNN_REQUIRE(B <= 65535, FRCNN_ELIMIT, "sample_and_gather: B > 65535");
    FRCNN_REQUIRE(target_labels && indices, FRCNN_EINVAL, "sample_and_gather: target_labels and indices are required");
    FRCNN_REQUIRE((!out_scores || pred_scores) && (!out_target_boxes || target_boxes) && (!out_pred_boxes || pred_boxes),
                  FRCNN_EINVAL, "sample_and_gather: an output was requested without its source tensor");
    FRCNN_REQUIRE(workspace && workspace_bytes >= frcnn_sample_indices_workspace_bytes(B, N), FRCNN_EWORKSPACE,
                  "sample_and_gather: workspace too small");
    FRCNN_REQUIRE(foreground_proportion >= 0.0 && foreground_proportion <= 1.0, FRCNN_EINVAL,
                  "sample_and_gather: foreground_proportion %g outside [0, 1]", foreground_proportion);
    // tf.math.round(num_samples * foreground_proportion): the product is a Python float (double), then round-half-even
    int quota = (int)nearbyint((double)num_samples * foreground_proportion);
    quota = quota < 0 ? 0 : quota > num_samples ? num_samples : quota;
    // the chunked selection keeps the best `quota` keys at the front of a kSortCap-entry buffer and needs room for new ones
    FRCNN_REQUIRE(quota < kSortCap, FRCNN_ELIMIT, "sample_and_gather: foreground quota %d >= %d (num_samples * foreground_proportion too large)",
                  quota, kSortCap);
    sample_gather_kernel<<<B, kSampThreads, 0, as_stream(stream)>>>(
        target_labels, N, num_classes_plus_1, num_samples, quota, (uint32_t)seed, (uint32_t)(seed >> 32),
        (int*)workspace, (long long*)indices, counts, pred_scores, target_boxes, pred_boxes, out_labels, out_scores,
        out_target_boxes, out_pred_boxes);
    FRCNN_LAUNCH_CHECK("sample_gather_kernel");
    return 0;
}

}  // extern "C"

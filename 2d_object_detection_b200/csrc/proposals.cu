// Proposal selection: utils/post_processing.py:6-63 postprocess_output, i.e. decode -> to_relative ->
// drop background column -> tf.image.combined_non_max_suppression -> padded outputs.
//
// Three kernels, one launch each for the whole batch:
//   decode_filter_kernel  fused decode + normalise + score threshold, one float4 per box, 44 B/box.
//   nms_select_kernel     one persistent CTA per (image, class) segment.  It walks the candidates in
//                         (score desc, index asc) order in chunks of 512: an 11-bit radix *select* finds
//                         the next chunk (no full sort of N), a bitonic network orders it in shared
//                         memory, a 512x512 suppression bitmask is built from shared-memory box tiles and
//                         reduced 64 candidates at a time with register bit operations.  It stops as soon
//                         as max_output_size_per_class boxes are kept, exactly like TF's greedy loop, so
//                         the reference's "no pre-NMS top-k" behaviour (trap T1) costs nothing extra.
//   finalize_kernel       per image: merge classes by score, clip to [0,1], zero pad.
#include "common.cuh"

namespace frcnn {

// ------------------------------------------------------------------------------------------------
// Stage 1: decode + to_relative + background drop + score threshold.
// boxes_rel is class-major [B,nc,N,4] so each NMS segment is contiguous; keys are the order-preserving
// score keys (0 = not a candidate).  Boxes stay UNCLIPPED here (trap T2).
//
// kLogits: `scores` holds the classification head's raw outputs after the reshape (rpn_detector.py:81-83,
// fast_rcnn_detector.py:60-61) and the Softmax activation (rpn_detector.py:84) is computed here, so the
// probabilities never round-trip through HBM before the threshold (second half):
//   p_j = exp(l_j - max_k l_k) / sum_k exp(l_k - max_k l_k), sum left to right, IEEE division
// (the form of TF's GPU kernel; Eigen's CPU kernel multiplies by the reciprocal of the sum instead: <= 1 ulp apart).
// softmax_out (optional, [B,N,nc+1]) receives the probabilities the training losses need.
// ------------------------------------------------------------------------------------------------
template <bool kSingleClass, bool kLogits>
__global__ void __launch_bounds__(256)
    decode_filter_kernel(float img_h, float img_w, const float4* __restrict__ regions, int regions_batched,
                         const float* __restrict__ scores, const float4* __restrict__ deltas, int B, int N, int nc,
                         float thr, float4* __restrict__ boxes_rel, uint32_t* __restrict__ keys,
                         float* __restrict__ softmax_out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)B * N * nc;
    if (i >= total) return;
    int c;
    int64_t bn;
    if (kSingleClass) {
        c = 0;
        bn = i;
    } else {
        c = (int)(i % nc);
        bn = i / nc;
    }
    const int64_t n = bn % N;
    const int64_t b = bn / N;
    const float4 r = __ldg(regions + (regions_batched ? bn : n));
    const float4 t = ld_nc_f4(deltas + i);
    const float4 d = decode_box(t, r);                                                    // boxes.py:20-41
    const float4 rel = make_float4(fdiv(d.x, img_w), fdiv(d.y, img_h), fdiv(d.z, img_w), fdiv(d.w, img_h));  // :86-93
    float s;
    if (kSingleClass) {
        const float2 p = __ldg(reinterpret_cast<const float2*>(scores) + bn);  // [bg, fg] pair, one 8-byte load
        if (kLogits) {
            const float m = fmaxf(p.x, p.y);
            const float e0 = expf(fsub(p.x, m)), e1 = expf(fsub(p.y, m));
            const float sum = fadd(e0, e1);
            s = fdiv(e1, sum);
            if (softmax_out) reinterpret_cast<float2*>(softmax_out)[bn] = make_float2(fdiv(e0, sum), s);
        } else {
            s = p.y;
        }
    } else if (kLogits) {
        const float* row = scores + bn * (nc + 1);
        float m = __ldg(row);
        for (int j = 1; j <= nc; ++j) m = fmaxf(m, __ldg(row + j));
        float sum = 0.0f;
        for (int j = 0; j <= nc; ++j) sum = fadd(sum, expf(fsub(__ldg(row + j), m)));
        s = fdiv(expf(fsub(__ldg(row + 1 + c), m)), sum);
        if (softmax_out) {
            softmax_out[bn * (nc + 1) + 1 + c] = s;
            if (c == 0) softmax_out[bn * (nc + 1)] = fdiv(expf(fsub(__ldg(row), m)), sum);
        }
    } else {
        s = __ldg(scores + bn * (nc + 1) + 1 + c);  // post_processing.py:52 drop background column
    }
    const int64_t o = (b * nc + c) * N + n;
    boxes_rel[o] = rel;
    keys[o] = (s > thr) ? score_key(s) : 0u;  // strict > (trap T3); NaN compares false
}

__global__ void score_keys_kernel(const float* __restrict__ scores, int64_t n, float thr, uint32_t* __restrict__ keys) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float s = scores[i];
    keys[i] = (s > thr) ? score_key(s) : 0u;
}

// ------------------------------------------------------------------------------------------------
// Stage 2: segmented greedy NMS.
// ------------------------------------------------------------------------------------------------
constexpr int kNmsThreads = 1024;
constexpr int kChunk = 1024;         // candidates whose boxes are staged + resolved at a time (sub-chunk of the sorted list)
constexpr int kList = 4096;          // candidates selected + sorted per refill (a refill = one radix select over all N keys)
constexpr int kBins = 2048;          // 11-bit radix digits

// TF CombinedNMS IOU on corner-normalised boxes with precomputed areas; returns iou > thr.
//
// The exact rule divides: iou = inter / union (IEEE), compared with the threshold.  A warp tests 32 candidates against one
// box and some lane nearly always overlaps it, so the division (and its divergent branch) ran for almost every test.  The
// division is only needed when inter / union is within a few ulps of the threshold: with thr_lo = thr (1 - 2^-18) and
// thr_hi = thr (1 + 2^-18), inter > thr_hi * union implies rn(inter / union) > thr and inter < thr_lo * union implies
// rn(inter / union) < thr (the products and the quotient are each off by at most 2^-24 relative), so those two outcomes
// are decided with one multiplication each and only the sliver in between takes the exact path.  NmsThreshold::make
// disables the shortcut (thr_lo = -inf, thr_hi = +inf) for thresholds below 2^-10, negative or non-finite.
struct NmsThreshold {
    float thr, lo, hi;
    __device__ static NmsThreshold make(float thr) {
        NmsThreshold t{thr, -INFINITY, INFINITY};
        if (thr >= 0.0009765625f && thr <= 1.0e6f) {
            t.lo = fmul(thr, 1.0f - 3.814697265625e-6f);
            t.hi = fmul(thr, 1.0f + 3.814697265625e-6f);
        }
        return t;
    }
};
__device__ __forceinline__ bool tf_suppresses(const float4& a, float area_a, const float4& b, float area_b, const NmsThreshold& t) {
    const float d0 = fmaxf(fsub(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.0f);
    const float d1 = fmaxf(fsub(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.0f);
    const float inter = fmul(d0, d1);
    const float uni = fsub(fadd(area_a, area_b), inter);
    const bool valid = area_a > 0.0f && area_b > 0.0f && uni >= 1.0e-30f;  // (normal products: the error bounds hold)
    if (valid && inter > fmul(t.hi, uni)) return true;
    if (valid && inter < fmul(t.lo, uni)) return false;  // (includes inter == 0)
    float iou = 0.0f;  // exact rule: degenerate boxes, NaNs, and quotients within 2^-18 of the threshold
    if (!(area_a <= 0.0f || area_b <= 0.0f)) {
        if (inter == 0.0f && t.thr >= 0.0f) return false;  // 0 / positive = 0, never > thr
        iou = fdiv(inter, uni);
    }
    return iou > t.thr;
}

// The same rule on SANITISED areas (sane_area: NaN stands for "area <= 0 or NaN"), the form the NMS kernel's inner loops
// use: a degenerate box makes the union NaN, both shortcut comparisons fail, and the exact rule (on the areas recomputed
// from the corners) decides -- three compares and two products on the common path instead of three more area tests.
__device__ __forceinline__ float box_area(const float4& b) { return fmul(fsub(b.z, b.x), fsub(b.w, b.y)); }
__device__ __forceinline__ float sane_area(float area) { return area > 0.0f ? area : __int_as_float(0x7fc00000); }
__device__ __forceinline__ bool tf_suppresses_sane(const float4& a, float sa, const float4& b, float sb, const NmsThreshold& t) {
    const float d0 = fmaxf(fsub(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.0f);
    const float d1 = fmaxf(fsub(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.0f);
    const float inter = fmul(d0, d1);
    const float uni = fsub(fadd(sa, sb), inter);
    const bool normal = uni >= 1.0e-30f;  // (false for NaN: a degenerate box)
    // the common outcome first (most pairs do not overlap at all); `&`: one combined predicate, one branch
    if (normal & (inter < fmul(t.lo, uni))) return false;
    if (normal & (inter > fmul(t.hi, uni))) return true;
    return tf_suppresses(a, box_area(a), b, box_area(b), t);
}

__device__ __forceinline__ uint64_t composite(uint32_t key, uint32_t idx) {
    return (static_cast<uint64_t>(key) << 32) | static_cast<uint64_t>(0xffffffffu - idx);
}

// Outputs of postprocess_output for a single-class call (the RPN form): with one class per image there is nothing to merge,
// so the NMS CTA of an image writes the padded outputs itself and the finalize launch disappears (out_boxes == nullptr:
// not fused -- the multi-class form, or the stand-alone NMS entry point).
struct NmsFinalize {
    float4* out_boxes;
    float* out_scores;
    int32_t* out_classes;
    int32_t* out_valid;
    int32_t* out_idx;
    int32_t* out_exhausted;
    int R;
};

struct NmsShared {
    unsigned long long hi;        // exclusive upper bound of unvisited composites
    unsigned long long pfx, pmask;
    unsigned long long thresh;    // chunk = eligible composites >= thresh
    int need, want_max, want_min, expect, kept, budget, remaining, stop, select_done;
    unsigned int cnt;
    unsigned int wsum[32], wtotal; // block-wide scan of the histogram: candidates in the warps before a warp, all candidates
    unsigned int wmax[32], wmin[32]; // per-warp largest / smallest candidate key (keys in shared memory)
    // per 64-block, double-buffered by block parity (the buffer of the next block is zeroed while this block is resolved):
    unsigned int dead[2][2];      // candidates suppressed by already kept boxes
    unsigned int diag[2][128];    // 64x64 suppression bits among the block (row r: words 2r, 2r+1)
};

constexpr size_t kNmsSharedBytes = (sizeof(NmsShared) + 15) / 16 * 16;

// One row of the padded outputs (post_processing.py:57-63): the kept box clipped to [0,1] (clip_boxes=True), its score and
// class; zeros (and index -1) past the valid detections.
__device__ __forceinline__ void write_detection(float4* __restrict__ out_boxes, float* __restrict__ out_scores,
                                                int32_t* __restrict__ out_classes, int32_t* __restrict__ out_idx, size_t row,
                                                bool valid, const float4* __restrict__ seg_boxes,
                                                const uint32_t* __restrict__ seg_keys, int idx, int cls) {
    float4 ob = make_float4(0.f, 0.f, 0.f, 0.f);
    float os = 0.f;
    int oc = 0, oi = -1;
    if (valid) {
        const float4 v = seg_boxes[idx];
        ob = make_float4(fmaxf(fminf(v.x, 1.f), 0.f), fmaxf(fminf(v.y, 1.f), 0.f), fmaxf(fminf(v.z, 1.f), 0.f),
                         fmaxf(fminf(v.w, 1.f), 0.f));
        os = key_score(seg_keys[idx]);
        oc = cls;
        oi = idx;
    }
    out_boxes[row] = ob;
    out_scores[row] = os;
    out_classes[row] = oc;
    if (out_idx) out_idx[row] = oi;
}

__device__ __forceinline__ uint4 lds_u4(const uint32_t* p) {
    uint4 q;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w)
                 : "r"((unsigned)__cvta_generic_to_shared(p)));
    return q;
}
__device__ __forceinline__ uint32_t lds_u1(const uint32_t* p) {
    uint32_t q;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(q) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return q;
}

// One sweep over the N score keys of a segment: f(key, index) for every key.  kShared: the keys sit in shared memory (copied
// there once per launch when they fit); otherwise they are read from global memory.  128-bit loads, four of them in flight
// per thread before the first key is consumed (the sweeps are latency chains otherwise: ~4 dependent round trips per thread
// at N = 16848).
template <bool kShared, class F>
__device__ __forceinline__ void for_each_key(const uint32_t* __restrict__ base, bool vec4, int N, int tid, F&& f) {
    if (vec4) {
        const int n4 = N >> 2;
        auto load4 = [&](int v) { return kShared ? lds_u4(base + 4 * v) : __ldg(reinterpret_cast<const uint4*>(base) + v); };
        auto use4 = [&](const uint4& q, int v) {
            f(q.x, 4 * v);
            f(q.y, 4 * v + 1);
            f(q.z, 4 * v + 2);
            f(q.w, 4 * v + 3);
        };
        int v = tid;
        if (!kShared)  // (shared memory: the plain loop below -- short latency, and the unrolled body is a lot of cold code)
        for (; v + 3 * kNmsThreads < n4; v += 4 * kNmsThreads) {  // whole groups: four unconditional loads, then the work
            const uint4 q0 = load4(v), q1 = load4(v + kNmsThreads), q2 = load4(v + 2 * kNmsThreads), q3 = load4(v + 3 * kNmsThreads);
            use4(q0, v);
            use4(q1, v + kNmsThreads);
            use4(q2, v + 2 * kNmsThreads);
            use4(q3, v + 3 * kNmsThreads);
        }
        for (; v < n4; v += kNmsThreads) use4(load4(v), v);
        for (int n = (n4 << 2) + tid; n < N; n += kNmsThreads) f(kShared ? lds_u1(base + n) : __ldg(base + n), n);
    } else {
        for (int n = tid; n < N; n += kNmsThreads) f(kShared ? lds_u1(base + n) : __ldg(base + n), n);
    }
}

// One persistent CTA per (image, class) segment.
//   loop over REFILLS of the candidate list in (score desc, index asc) order:
//     SELECT   11-bit MSD radix select over the 64-bit composite (score key, ~index) finds a threshold that cuts
//              off the next unvisited candidates: <= 512 when at most 512 boxes are kept (the greedy loop then usually
//              ends within two lists), <= 4096 for larger caps; whole histogram bins are accepted as soon as they give
//              at least half a list, so unique scores need two histogram sweeps over the keys.
//     COLLECT  + bitonic SORT of the list in shared memory.
//     per 1024-candidate sub-chunk of the sorted list: boxes staged (corner-normalised) with their areas, then
//     GREEDY   64 candidates at a time: (a) all threads test the block against the kept list (shared memory,
//              broadcast reads), (b) all threads build the 64x64 bitmask of the block, (c) one warp resolves the
//              block with register bit operations and appends the survivors to the kept list.
//   The loop ends as soon as `cap` boxes are kept -- TF's greedy loop exactly, no pre-NMS cut needed (trap T1).
// keys_in_smem: the segment's N score keys are copied to shared memory once and every later sweep (two or three histogram
// sweeps + one collect sweep per refill) reads them there (N <= ~40k: the KITTI grids); larger segments sweep global memory.
__global__ void __launch_bounds__(kNmsThreads, 1)
    nms_select_kernel(const float4* __restrict__ boxes, const uint32_t* __restrict__ keys, int N, int cap, int top_n,
                      float iou_thr, int32_t* __restrict__ kept_idx, int32_t* __restrict__ kept_count,
                      int32_t* __restrict__ exhausted, int keys_in_smem, const NmsFinalize fin) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // 8-byte items first, then 16-byte boxes, then 4-byte arrays: every sub-array stays naturally aligned for any cap
    NmsShared* S = reinterpret_cast<NmsShared*>(smem_raw);
    unsigned long long* comp = reinterpret_cast<unsigned long long*>(smem_raw + kNmsSharedBytes);  // [kList]
    float4* cbox = reinterpret_cast<float4*>(comp + kList);                                 // [kChunk]
    float4* kbox = cbox + kChunk;                                                          // [cap]
    float* carea = reinterpret_cast<float*>(kbox + cap);                                    // [kChunk]
    float* karea = carea + kChunk;                                                         // [cap]
    unsigned int* hist = reinterpret_cast<unsigned int*>(karea + cap);                      // [kBins]
    uint32_t* skey_s = reinterpret_cast<uint32_t*>(                                         // [N] (keys_in_smem), 16-byte aligned
        (reinterpret_cast<uintptr_t>(hist + kBins) + 15) & ~static_cast<uintptr_t>(15));

    const int seg = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t* skeys = keys + (size_t)seg * N;
    const float4* sboxes = boxes + (size_t)seg * N;
    int32_t* sidx = kept_idx + (size_t)seg * cap;

    if (tid == 0) {
        S->hi = ~0ull;
        S->kept = 0;
        S->budget = top_n > 0 ? top_n : 0x7fffffff;
        S->remaining = 0;
        S->stop = 0;
    }
    int flag_exhausted = 0;
    const bool vec4 = (reinterpret_cast<uintptr_t>(skeys) & 15u) == 0;  // this segment's keys allow 128-bit loads
    // Keys are selected RELATIVE TO THE LARGEST candidate key: adding koff moves it to 0xfffffffe (a translation: the
    // order, and so every selected set, is unchanged; not 0xffffffff, so that every composite stays below the initial
    // bound hi = ~0), and the leading `lead` bits that the largest and the smallest candidate then share are known.  The
    // radix passes start below them.  Scores are probabilities: half of them lie
    // in [0.5, 1), which is FOUR bins of a histogram over the top 11 bits of the raw float key -- thousands of same-address
    // shared-memory atomics per sweep; below the common leading bits the same keys spread over 64 bins or more.
    uint32_t koff = 0u;
    uint32_t cand_lo = 1u;                          // no candidate's translated key lies below this (a key 0 is no candidate)
    unsigned long long pfx0 = 0ull, pmask0 = 0ull;  // the decided leading bits of every candidate's composite
    if (keys_in_smem) {
        uint32_t kmax = 0u, kmin = 0xffffffffu;
        auto seen = [&](uint32_t k) {
            if (k != 0u) {
                kmax = max(kmax, k);
                kmin = min(kmin, k);
            }
        };
        auto seen4 = [&](const uint4& q) {
            seen(q.x);
            seen(q.y);
            seen(q.z);
            seen(q.w);
        };
        if (vec4) {
            const int n4 = N >> 2;
            const uint4* k4 = reinterpret_cast<const uint4*>(skeys);
            uint4* s4 = reinterpret_cast<uint4*>(skey_s);
            int v = tid;
            for (; v + 3 * kNmsThreads < n4; v += 4 * kNmsThreads) {
                const uint4 q0 = __ldg(k4 + v), q1 = __ldg(k4 + v + kNmsThreads), q2 = __ldg(k4 + v + 2 * kNmsThreads),
                            q3 = __ldg(k4 + v + 3 * kNmsThreads);
                s4[v] = q0;
                s4[v + kNmsThreads] = q1;
                s4[v + 2 * kNmsThreads] = q2;
                s4[v + 3 * kNmsThreads] = q3;
                seen4(q0);
                seen4(q1);
                seen4(q2);
                seen4(q3);
            }
            for (; v < n4; v += kNmsThreads) {
                const uint4 q = __ldg(k4 + v);
                s4[v] = q;
                seen4(q);
            }
            for (int n = (n4 << 2) + tid; n < N; n += kNmsThreads) {
                const uint32_t k = __ldg(skeys + n);
                skey_s[n] = k;
                seen(k);
            }
        } else {
            for (int n = tid; n < N; n += kNmsThreads) {
                const uint32_t k = __ldg(skeys + n);
                skey_s[n] = k;
                seen(k);
            }
        }
        kmax = __reduce_max_sync(0xffffffffu, kmax);
        kmin = __reduce_min_sync(0xffffffffu, kmin);
        if (lane == 0) {
            S->wmax[warp] = kmax;
            S->wmin[warp] = kmin;
        }
        __syncthreads();
        kmax = __reduce_max_sync(0xffffffffu, S->wmax[lane]);
        kmin = __reduce_min_sync(0xffffffffu, S->wmin[lane]);
        if (kmax != 0u) {  // (no candidate at all: the first histogram is empty and the loop ends)
            koff = 0xfffffffeu - kmax;
            cand_lo = kmin + koff;
            const uint32_t x = 0xfffffffeu ^ (kmin + koff);  // translated extremes: everything between shares their prefix
            const int lead = x ? __clz((int)x) : 32;
            // undecided composite bits: [0, 64 - lead); all keys equal leaves the index half (bits 22..31 of ~idx are ones)
            pmask0 = lead == 0 ? 0ull : lead == 32 ? 0xffffffffffc00000ull : (~0ull << (64 - lead));
            pfx0 = 0xfffffffeffffffffull & pmask0;
        }
    } else {
        __syncthreads();
    }
#ifdef FRCNN_NMS_TIMING  // debug build: time spent per phase (block 0, thread 0; printed at the end)
    long long t_mark = clock64(), t_phase[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    int n_chunks = 0, n_blocks = 0;
#define NMS_TICK(i)                                   \
    do {                                              \
        const long long t_now__ = clock64();          \
        t_phase[i] += t_now__ - t_mark;               \
        t_mark = t_now__;                             \
    } while (0)
#else
#define NMS_TICK(i) do { } while (0)
#endif

    const NmsThreshold nthr = NmsThreshold::make(iou_thr);
    const bool key32_vec = (reinterpret_cast<uintptr_t>(hist) & 15u) == 0;  // (cap % 4 == 0: the usual 300 / 1000)

    while (true) {
        // ===================== SELECT ===================================================================
        if (tid == 0) {
            S->pfx = pfx0;
            S->pmask = pmask0;
            const int budget = S->budget;
            // list target per refill.  The greedy loop usually reaches `cap` long before pre_nms_top_n candidates have been
            // visited (KITTI: 300 kept after ~700 of 2000), so short lists that are selected lazily beat one exact select of
            // everything (measured: 63 us vs 86 us per launch): 512 candidates when at most 512 boxes are kept, otherwise
            // kList (dense-anchor configurations: fewer sweeps over all N keys).
            const int target = cap > kChunk / 2 ? kList : kChunk / 2;
            S->want_max = min(target, budget);
            // exact count only when pre_nms_top_n ends inside this chunk; otherwise any size in [half, full] will do
            S->want_min = budget <= target ? S->want_max : target / 2;
            S->need = S->want_max;
            S->select_done = 0;
            S->expect = 0;
            S->cnt = 0;
        }
        __syncthreads();
        const unsigned long long hi = S->hi;
        for (int pass = 0;; ++pass) {
            const unsigned long long pfx = S->pfx, pmask = S->pmask;  // (written before the barrier that ended the last pass)
            // digit layout: 11 bits at a time from the highest undecided bit of the (translated) key down to its bit 0 (the
            // last key digit takes what is left), then ~idx bits 21..11 and 10..0 (idx < 2^22).  pmask is a run of leading
            // ones, so the highest undecided bit is 63 - clz(~pmask).
            const int top = 64 - __clzll((long long)~pmask);
            const int width = top > 32 ? min(11, top - 32) : 11;
            const int shift = top - width;
            const int nb = 1 << width;
            for (int i = tid; i < nb; i += kNmsThreads) hist[i] = 0;
            __syncthreads();
            NMS_TICK(14);
            // one histogram sweep over the keys: 128-bit loads (four keys per thread and iteration, independent of each
            // other, so the loads of a thread overlap) and one fire-and-forget shared-memory atomic per participating key --
            // the match_any aggregation it replaces cost more than the bank conflicts it avoided (4.7 us per sweep; eight
            // private copies of the histogram were tried too: zeroing and summing them costs more than the conflicts)
            auto count_key = [&](uint32_t k, int n) {
                const unsigned long long c = composite(k + koff, (uint32_t)n);
                if ((k != 0u) && (c < hi) && ((c & pmask) == pfx))
                    atomicAdd(&hist[(unsigned int)(c >> shift) & (unsigned int)(nb - 1)], 1u);
            };
            // While the digit lies in the key half (nearly always the only passes), everything is decided on 32 bits: the
            // keys that carry the current prefix and lie below the bound form ONE interval [lo_r, lo_r + span], so a key
            // takes a subtraction and an unsigned compare (a key that is not a candidate, 0, lands below every candidate's
            // translated key); only a key equal to the bound's key half looks at its index.
            const unsigned int pm_k = (unsigned int)(pmask >> 32), pf_k = (unsigned int)(pfx >> 32);
            const unsigned int hi_k = (unsigned int)(hi >> 32), hi_i = (unsigned int)hi;
            const unsigned int lo_r = max(pf_k, cand_lo), hi_r = min(pf_k | ~pm_k, hi_k);
            const unsigned int span = hi_r - lo_r;
            const int sh_k = shift - 32;
            auto count_key32 = [&](uint32_t k, int n) {
                const unsigned int kk = k + koff;
                if (kk - lo_r <= span && (kk != hi_k || (0xffffffffu - (unsigned int)n) < hi_i))
                    atomicAdd(&hist[(kk >> sh_k) & (unsigned int)(nb - 1)], 1u);
            };
            if (shift >= 32) {
                if (hi_r >= lo_r) {
                    if (keys_in_smem) for_each_key<true>(skey_s, true, N, tid, count_key32);
                    else for_each_key<false>(skeys, vec4, N, tid, count_key32);
                }
            } else {
                if (keys_in_smem) for_each_key<true>(skey_s, true, N, tid, count_key);
                else for_each_key<false>(skeys, vec4, N, tid, count_key);
            }
            NMS_TICK(15);
            __syncthreads();
            {   // Block-wide scan of the histogram in descending digit order (position p <-> digit nb - 1 - p): every thread
                // sums its K digits, warp scans + one scan of the 32 warp totals give each thread the number of candidates
                // above its digits; the one thread whose digits contain rank `need` decides.  (A single warp walking 64
                // digits per lane did this before: as long as a sweep over the keys.)
                int need = S->need;            // still wanted from inside the current prefix (read before anybody writes)
                const int got = S->expect;     // already secured by bins above the current prefix
                const int K = nb > kNmsThreads ? nb / kNmsThreads : 1;  // 2 (11-bit digits) or 1 (narrower last key digit)
                const int top_bin = nb - 1 - tid * K;                   // (negative: a thread beyond a narrow digit's bins)
                unsigned int hk[2] = {0u, 0u};  // this thread's digits
#pragma unroll
                for (int q = 0; q < 2; ++q)
                    if (q < K && top_bin - q >= 0) hk[q] = hist[top_bin - q];
                const unsigned int sum = hk[0] + hk[1];
                unsigned int incl = sum;
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                if (lane == 31) S->wsum[warp] = incl;
                __syncthreads();
                if (warp == 0) {
                    const unsigned int w = S->wsum[lane];
                    unsigned int wi = w;
                    for (int o = 1; o < 32; o <<= 1) {
                        const unsigned int v = __shfl_up_sync(0xffffffffu, wi, o);
                        if (lane >= o) wi += v;
                    }
                    S->wsum[lane] = wi - w;  // candidates in the warps before this one
                    if (lane == 31) S->wtotal = wi;
                }
                __syncthreads();
                const unsigned int total = S->wtotal;
                const unsigned int excl = S->wsum[warp] + incl - sum;
                if (pass == 0) {               // total = all unvisited candidates of this segment
                    if ((int)total < need) need = (int)total;
                    if (tid == 0) S->remaining = (int)total;
                }
                if (need > 0 && excl < (unsigned)need && (unsigned)need <= excl + sum) {  // exactly one thread
                    unsigned int cum = excl;
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const unsigned int h = hk[q];
                        if (q < K && cum + h >= (unsigned)need) {
                            const unsigned long long bin = (unsigned long long)(top_bin - q);
                            unsigned long long npfx = pfx | (bin << shift);
                            unsigned long long nmask = pmask | ((unsigned long long)(nb - 1) << shift);
                            const int in_bin = need - (int)cum;
                            if ((int)h == in_bin) {  // whole bin taken: exact threshold
                                S->thresh = npfx;
                                S->expect = got + (int)cum + (int)h;
                                S->select_done = 1;
                            } else if (got + (int)cum >= S->want_min) {  // bins above this one are enough
                                S->thresh = npfx + (1ull << shift);      // first composite above the bin
                                S->expect = got + (int)cum;
                                S->select_done = 1;
                            } else {
                                if (shift == 32) {  // the key is decided; bits 22..31 of ~idx are constant ones
                                    npfx |= 0xffc00000ull;
                                    nmask |= 0xffc00000ull;
                                }
                                S->expect = got + (int)cum;
                                S->need = in_bin;
                            }
                            S->pfx = npfx;
                            S->pmask = nmask;
                            break;
                        }
                        cum += h;
                    }
                }
                if (need == 0 && tid == 0) S->select_done = 1;
            }
            __syncthreads();
            if (S->select_done) break;
        }
        const int expect = S->expect;
        NMS_TICK(0);
#ifdef FRCNN_NMS_TIMING
        ++n_chunks;
#endif
        if (expect == 0) break;  // no candidates left (uniform)
        const unsigned long long thresh = S->thresh;

        // ===================== COLLECT the chunk (unordered) ===========================================
        // Only ~ one key in 30 is taken, so every taker simply reserves its slot with a shared-memory atomic (the order
        // inside the list is irrelevant: it is sorted next); 128-bit key loads as in the histogram sweeps.
        auto collect_key = [&](uint32_t k, int n) {
            const unsigned long long c = composite(k + koff, (uint32_t)n);
            if ((k != 0u) && (c < hi) && (c >= thresh)) {
                const unsigned int pos = atomicAdd(&S->cnt, 1u);
                if (pos < (unsigned)kList) comp[pos] = c;
            }
        };
        if (keys_in_smem) {
            // Keys in shared memory: no atomics at all.  Sweep 1: every thread marks its takers in a 64-bit mask (bit 4 r + e
            // = key e of the thread's r-th 128-bit load; bit 63 = its key of the unaligned tail).  A block-wide exclusive
            // scan of the counts gives each thread its first slot (every warp scans the 32 warp totals itself: one barrier).
            // Sweep 2: the few threads that own a taker reload that load and write.  (One shared-memory atomic per taker on
            // the list counter, all 32 warps on one address, cost 5.6 k cycles per refill; this costs about 2 k.)
            const int n4 = N >> 2;
            const unsigned int hi_k = (unsigned int)(hi >> 32), th_k = (unsigned int)(thresh >> 32);
            // key half first: one subtraction + unsigned compare against [max(th_k, cand_lo), hi_k] (a key 0, not a
            // candidate, lands below cand_lo); only keys that tie with a bound compare whole composites
            const unsigned int lo_c = max(th_k, cand_lo), span_c = hi_k - lo_c;  // (hi_k >= th_k >= ... : thresh < hi)
            const bool th_tie = (unsigned int)thresh != 0u;
            auto takes = [&](uint32_t k, int n) {
                const unsigned int kk = k + koff;
                bool take = kk - lo_c <= span_c;
                if (take && (kk == hi_k || (th_tie && kk == th_k))) {
                    const unsigned long long c = composite(kk, (uint32_t)n);
                    take = c < hi && c >= thresh;
                }
                return take;
            };
            unsigned long long mask = 0ull;
            {
                int r = 0;
                for (int v = tid; v < n4; v += kNmsThreads, ++r) {
                    const uint4 q = lds_u4(skey_s + 4 * v);
                    const unsigned int nib = (takes(q.x, 4 * v) ? 1u : 0u) | (takes(q.y, 4 * v + 1) ? 2u : 0u) |
                                             (takes(q.z, 4 * v + 2) ? 4u : 0u) | (takes(q.w, 4 * v + 3) ? 8u : 0u);
                    mask |= (unsigned long long)nib << (4 * r);
                }
                const int n = (n4 << 2) + tid;
                if (n < N && takes(lds_u1(skey_s + n), n)) mask |= 1ull << 63;
            }
            NMS_TICK(9);
            const int mine = __popcll(mask);
            int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) S->wsum[warp] = (unsigned int)incl;
            __syncthreads();
            int wincl = (int)S->wsum[lane];
            NMS_TICK(10);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, wincl, o);
                if (lane >= o) wincl += t;
            }
            const int wbase = warp == 0 ? 0 : __shfl_sync(0xffffffffu, wincl, (warp + 31) & 31);
            if (tid == kNmsThreads - 1) S->cnt = (unsigned int)wincl;  // (lane 31: the block total)
            NMS_TICK(11);
            if (mask) {
                unsigned int pos = (unsigned int)(wbase + incl - mine);
                int r = 0;
                for (int v = tid; v < n4; v += kNmsThreads, ++r) {
                    const unsigned int nib = (unsigned int)(mask >> (4 * r)) & 15u;
                    if (nib == 0u) continue;
                    const uint4 q = lds_u4(skey_s + 4 * v);
                    const unsigned int kq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if ((nib >> e) & 1u) {
                            if (pos < (unsigned)kList) comp[pos] = composite(kq[e] + koff, (uint32_t)(4 * v + e));
                            ++pos;
                        }
                    }
                }
                if (mask >> 63) {
                    const int n = (n4 << 2) + tid;
                    if (pos < (unsigned)kList) comp[pos] = composite(lds_u1(skey_s + n) + koff, (uint32_t)n);
                }
            }
            NMS_TICK(12);
        } else {
            for_each_key<false>(skeys, vec4, N, tid, collect_key);
        }
        __syncthreads();
        const int m = min((int)S->cnt, kList);  // == expect by construction
        NMS_TICK(13);
        int p2 = 32;
        while (p2 < m) p2 <<= 1;
        // ===================== SORT descending (bitonic network) ========================================
        // Exchanges at distance >= 32 go through shared memory (one barrier each); the distances 16 .. 1 that end every
        // phase stay inside a warp and run on registers with shuffles -- 19 barriers instead of 45 for 512 candidates.
        //
        // Short lists (the usual 512-candidate refill) are not sent through the network at all: every key of the list lies
        // in [key half of thresh, key half of hi], normally a few hundred thousand key values, so (key - base + 1) << 9 | slot
        // fits a 32-bit word and orders the list whenever no two candidates share a key.  Those words are RANKED BY
        // COUNTING: two threads per candidate count the words above theirs in one half of the list each (broadcast 128-bit
        // loads, a compare and an add per word: 512 x 512 compares on 1024 threads), the rank is where the candidate's
        // composite goes.  No barrier, no dependent shuffle chain; 19 barriers and 45 exchange steps otherwise.  Equal keys
        // (their order is the index order the composite carries) show up as neighbours with the same key half and send the
        // list through the 64-bit network instead.
        const unsigned long long* sorted = comp;
        bool sorted_done = false;
        const unsigned int base_k = (unsigned int)(thresh >> 32);
        if (m <= 512 && (unsigned int)(hi >> 32) - base_k < (1u << 23) - 2u) {  // (uniform)
            unsigned int* key32 = hist;  // (the histogram is dead until the next refill; 16-byte aligned? see below)
            unsigned long long* comp2 = reinterpret_cast<unsigned long long*>(
                (reinterpret_cast<uintptr_t>(hist + 512) + 7) & ~static_cast<uintptr_t>(7));
            if (tid < 512)
                key32[tid] = tid < m ? ((((unsigned int)(comp[tid] >> 32) - base_k + 1u) << 9) | (unsigned int)tid) : 0u;
            __syncthreads();
            NMS_TICK(1);
            {
                const int e = tid >> 1, h = tid & 1;
                const unsigned int mine = key32[e];  // (0 for e >= m: counts everything, never used)
                const int mh = ((m + 7) >> 3) << 2;   // words per half, a multiple of 4 (the list is zero-padded to 512)
                const unsigned int* half = key32 + mh * h;
                unsigned int c0 = 0u, c1 = 0u, c2 = 0u, c3 = 0u;  // four independent counters: no dependent chain
                auto count = [&](unsigned int& c, unsigned int q) {  // one compare + one guarded add
                    asm("{\n\t.reg .pred p;\n\tsetp.gt.u32 p, %1, %2;\n\t@p add.u32 %0, %0, 1;\n\t}" : "+r"(c) : "r"(q), "r"(mine));
                };
#pragma unroll 4
                for (int u = 0; u < mh; u += 4) {
                    // (four words per load; hist is only 4-byte aligned in general, so plain 32-bit loads when it is not)
                    unsigned int q0, q1, q2, q3;
                    if (key32_vec) {
                        const uint4 q = *reinterpret_cast<const uint4*>(half + u);
                        q0 = q.x; q1 = q.y; q2 = q.z; q3 = q.w;
                    } else {
                        q0 = half[u]; q1 = half[u + 1]; q2 = half[u + 2]; q3 = half[u + 3];
                    }
                    count(c0, q0);
                    count(c1, q1);
                    count(c2, q2);
                    count(c3, q3);
                }
                int above = (int)(c0 + c1 + c2 + c3);
                above += __shfl_xor_sync(0xffffffffu, above, 1);
                if (h == 0 && e < m) comp2[above] = comp[e];
            }
            __syncthreads();
            const bool tie = tid > 0 && tid < m && (unsigned int)(comp2[tid] >> 32) == (unsigned int)(comp2[tid - 1] >> 32);
            if (!__syncthreads_or(tie ? 1 : 0)) {
                sorted = comp2;
                sorted_done = true;
            }
        }
        if (!sorted_done) {
        for (int i = m + tid; i < p2; i += kNmsThreads) comp[i] = 0ull;
        __syncthreads();

        NMS_TICK(1);
        for (int k = 2; k <= p2; k <<= 1) {
            int j = k >> 1;
            for (; j >= 32; j >>= 1) {
                for (int t = tid; t < p2; t += kNmsThreads) {
                    const int ixj = t ^ j;
                    if (ixj > t) {
                        const unsigned long long a = comp[t], b = comp[ixj];
                        const bool desc = (t & k) == 0;
                        if (desc ? (a < b) : (a > b)) {
                            comp[t] = b;
                            comp[ixj] = a;
                        }
                    }
                }
                __syncthreads();
            }
            for (int t = tid; t < p2; t += kNmsThreads) {  // (p2 is a multiple of 32: whole warps take part)
                unsigned long long v = comp[t];
                const bool desc = (t & k) == 0;
                for (int jj = j; jj > 0; jj >>= 1) {
                    const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, jj);
                    const bool keep_max = (((t & jj) == 0) == desc);  // the lower index of a descending pair keeps the larger
                    v = keep_max ? (v > o ? v : o) : (v < o ? v : o);
                }
                comp[t] = v;
            }
            __syncthreads();
        }
        }

        NMS_TICK(2);
      for (int sub = 0; sub < m && S->kept < cap; sub += kChunk) {  // (uniform: S->kept is written before a barrier)
        const unsigned long long* scomp = sorted + sub;
        const int ms = min(kChunk, m - sub);
        // ===================== LOAD boxes (corner-normalised) + areas of the sub-chunk ==================
        if (tid < ms) {
            const uint32_t idx = 0xffffffffu - (uint32_t)(scomp[tid] & 0xffffffffull);
            const float4 b = __ldg(sboxes + idx);
            const float4 nb4 = make_float4(fminf(b.x, b.z), fminf(b.y, b.w), fmaxf(b.x, b.z), fmaxf(b.y, b.w));
            cbox[tid] = nb4;
            carea[tid] = sane_area(box_area(nb4));  // (carea / karea hold sanitised areas: tf_suppresses_sane)
        }
        if (tid < 128) S->diag[0][tid] = 0u;  // bitmask buffers of the first 64-block (block b uses buffer b & 1)
        if (tid < 2) S->dead[0][tid] = 0u;
        __syncthreads();

        NMS_TICK(3);
        // ===================== GREEDY, 64 candidates at a time ==========================================
        const int nblk = (ms + 63) / 64;
        for (int blk = 0; blk < nblk; ++blk) {
            const int kept = S->kept;  // uniform: written before the barrier that ended the previous block
            if (kept >= cap) break;
            const int base = blk * 64;
            const int nvalid = min(64, ms - base);
            unsigned int* const diag = S->diag[blk & 1];
            unsigned int* const dead = S->dead[blk & 1];
            NMS_TICK(5);
#ifdef FRCNN_NMS_TIMING
            ++n_blocks;
#endif
            {   // (a) block vs kept list: thread = (candidate, one of 16 interleaved slices of the kept list)
                const int cand = tid & 63, part = tid >> 6;
                if (cand < nvalid && part < kept) {
                    const float4 bt = cbox[base + cand];
                    const float at = carea[base + cand];
                    bool sup = false;
                    // this thread's slice of the kept list: boxes part, part + 16, ... ; four per round at immediate offsets
                    constexpr int kStep = kNmsThreads / 64;
                    const float4* kb = kbox + part;
                    const float* ka = karea + part;
                    int left = (kept - part + kStep - 1) / kStep;
                    for (; left >= 4 && !sup; left -= 4, kb += 4 * kStep, ka += 4 * kStep) {
                        sup = tf_suppresses_sane(bt, at, kb[0], ka[0], nthr) || tf_suppresses_sane(bt, at, kb[kStep], ka[kStep], nthr) ||
                              tf_suppresses_sane(bt, at, kb[2 * kStep], ka[2 * kStep], nthr) ||
                              tf_suppresses_sane(bt, at, kb[3 * kStep], ka[3 * kStep], nthr);
                    }
                    for (; left > 0 && !sup; --left, kb += kStep, ka += kStep) sup = tf_suppresses_sane(bt, at, kb[0], ka[0], nthr);
                    if (sup) atomicOr(&dead[cand >> 5], 1u << (cand & 31));
                }
            }
            {   // (b) 64x64 upper triangle of the block: thread = (row, group of 4 columns)
                const int row = tid >> 4, g = tid & 15;
                if (row < nvalid && 4 * g + 3 > row) {
                    const float4 bt = cbox[base + row];
                    const float at = carea[base + row];
                    unsigned int bits = 0u;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int u = 4 * g + q;
                        if (u > row && u < nvalid && tf_suppresses_sane(bt, at, cbox[base + u], carea[base + u], nthr))
                            bits |= 1u << q;
                    }
                    if (bits) atomicOr(&diag[2 * row + (g >> 3)], bits << ((4 * g) & 31));
                }
            }
            NMS_TICK(6);
            __syncthreads();
            NMS_TICK(7);
            if (warp == 0) {  // (c) resolve: lane l holds the mask rows l and l+32, as 32-bit words (row r only has bits > r)
                const unsigned int vlo = nvalid >= 32 ? ~0u : ((1u << nvalid) - 1u);
                const unsigned int vhi = nvalid == 64 ? ~0u : nvalid > 32 ? ((1u << (nvalid - 32)) - 1u) : 0u;
                // cur: suppressed (or beyond the block).  Row i only sets bits above i, so bit i is final when step i reads it
                // and the survivors are simply the bits still clear at the end: one test and two guarded ORs per step.
                unsigned int cur_lo = dead[0] | ~vlo, cur_hi = dead[1] | ~vhi;
                const unsigned int r0_lo = diag[2 * lane], r0_hi = diag[2 * lane + 1], r1_hi = diag[2 * (lane + 32) + 1];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const unsigned int lo = __shfl_sync(0xffffffffu, r0_lo, i), hi = __shfl_sync(0xffffffffu, r0_hi, i);
                    if (!(cur_lo & (1u << i))) {
                        cur_lo |= lo;
                        cur_hi |= hi;
                    }
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const unsigned int hi = __shfl_sync(0xffffffffu, r1_hi, i);
                    if (!(cur_hi & (1u << i))) cur_hi |= hi;
                }
                unsigned long long keepbits = ~((unsigned long long)cur_lo | ((unsigned long long)cur_hi << 32));
                // the cap ends the greedy loop: only the first cap - kept survivors count (earlier decisions do not depend on
                // later candidates, so cutting the tail afterwards is the same loop)
                const int room = cap - kept;
                while (__popcll(keepbits) > room) keepbits &= ~(1ull << (63 - __clzll(keepbits)));
#pragma unroll
                for (int h = 0; h < 2; ++h) {  // append survivors (lane handles bits lane and lane+32)
                    const int i = lane + 32 * h;
                    if ((keepbits >> i) & 1ull) {
                        const int pos = kept + __popcll(keepbits & ((1ull << i) - 1ull));
                        kbox[pos] = cbox[base + i];
                        karea[pos] = carea[base + i];
                        sidx[pos] = (int32_t)(0xffffffffu - (uint32_t)(scomp[base + i] & 0xffffffffull));
                    }
                }
                if (lane == 0) S->kept = kept + __popcll(keepbits);
            } else if (warp <= 4) {  // meanwhile: zero the bitmask buffers of the next block
                S->diag[(blk + 1) & 1][tid - 32] = 0u;
                if (tid < 34) S->dead[(blk + 1) & 1][tid - 32] = 0u;
            }
            __syncthreads();
            NMS_TICK(8);
        }
        __syncthreads();  // the next sub-chunk overwrites cbox / carea
        NMS_TICK(4);
      }
        if (tid == 0) {
            S->hi = sorted[m - 1];
            S->budget -= m;
            S->remaining -= m;
            int stop = 0;
            if (S->kept >= cap) stop = 1;
            else if (S->remaining <= 0) stop = 1;  // every candidate visited
            else if (S->budget <= 0) stop = 2;     // pre_nms_top_n cut the list short
            S->stop = stop;
        }
        __syncthreads();
        const int stop = S->stop;
        if (stop) {
            flag_exhausted = (stop == 2);
            break;
        }
    }
    __syncthreads();
    const int kept = S->kept;
#ifdef FRCNN_NMS_TIMING
    if (tid == 0)
        printf("nms timing block %d refills %d blocks %d (cycles): select %lld collect %lld sort %lld load %lld greedy-rest %lld | per block: zero+bar %lld  a+b(thread 0) %lld  "
               "bar %lld  resolve+bar %lld   kept %d | collect: sweep1 %lld bar %lld scan %lld sweep2 %lld bar %lld pad %lld | select: zero+bar %lld sweep %lld scan+bar %lld\n", (int)blockIdx.x, n_chunks, n_blocks, t_phase[0], t_phase[1], t_phase[2], t_phase[3], t_phase[4], t_phase[5], t_phase[6],
               t_phase[7], t_phase[8], kept, t_phase[9], t_phase[10], t_phase[11], t_phase[12], t_phase[13], t_phase[1], t_phase[14], t_phase[15], t_phase[0]);
#endif
    for (int i = kept + tid; i < cap; i += kNmsThreads) sidx[i] = -1;
    if (tid == 0) {
        kept_count[seg] = kept;
        if (exhausted) exhausted[seg] = flag_exhausted;
    }
    if (fin.out_boxes) {  // single class: this segment is image `seg`; the kept list is already in score order
        const int nv = min(kept, fin.R);
        for (int t = tid; t < fin.R; t += kNmsThreads)
            write_detection(fin.out_boxes, fin.out_scores, fin.out_classes, fin.out_idx, (size_t)seg * fin.R + t, t < nv, sboxes, skeys,
                            t < nv ? sidx[t] : -1, 0);  // (sidx: written by this CTA before the barrier above)
        if (tid == 0) {
            fin.out_valid[seg] = nv;
            if (fin.out_exhausted) fin.out_exhausted[seg] = flag_exhausted;
        }
    }
}

static size_t nms_smem_bytes(int cap) {
    return sizeof(unsigned long long) * kList + sizeof(float4) * (kChunk + (size_t)cap) +
           sizeof(float) * (kChunk + (size_t)cap) + sizeof(unsigned int) * kBins + kNmsSharedBytes + 16;
}

// ------------------------------------------------------------------------------------------------
// Stage 3: per image merge across classes (score desc, then class, then rank), clip, pad.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
    finalize_kernel(const float4* __restrict__ boxes_rel, const uint32_t* __restrict__ keys, int N, int nc, int cap,
                    int R, const int32_t* __restrict__ kept_idx, const int32_t* __restrict__ kept_count,
                    const int32_t* __restrict__ seg_exhausted, float4* __restrict__ out_boxes,
                    float* __restrict__ out_scores, int32_t* __restrict__ out_classes, int32_t* __restrict__ out_valid,
                    int32_t* __restrict__ out_idx, int32_t* __restrict__ out_exhausted, int p2) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* comp = reinterpret_cast<unsigned long long*>(smem_raw);  // [p2] (nc > 1 only)
    const int b = blockIdx.x, tid = threadIdx.x;
    int total = 0;
    for (int c = 0; c < nc; ++c) total += kept_count[b * nc + c];
    const int nv = min(total, R);
    if (nc > 1) {
        for (int i = tid; i < p2; i += blockDim.x) {
            unsigned long long v = 0ull;
            const int c = i / cap, t = i % cap;
            if (c < nc && t < kept_count[b * nc + c]) {
                const int idx = kept_idx[((size_t)b * nc + c) * cap + t];
                v = composite(keys[((size_t)b * nc + c) * N + idx], (uint32_t)i);  // ord = c*cap + t
            }
            comp[i] = v;
        }
        __syncthreads();
        for (int k = 2; k <= p2; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = tid; t < p2; t += blockDim.x) {
                    const int ixj = t ^ j;
                    if (ixj > t) {
                        const unsigned long long a = comp[t], bb = comp[ixj];
                        const bool desc = (t & k) == 0;
                        if (desc ? (a < bb) : (a > bb)) {
                            comp[t] = bb;
                            comp[ixj] = a;
                        }
                    }
                }
                __syncthreads();
            }
    }
    for (int t = tid; t < R; t += blockDim.x) {
        int c = 0, rank = t;
        if (t < nv && nc > 1) {
            const uint32_t ord = 0xffffffffu - (uint32_t)(comp[t] & 0xffffffffull);
            c = (int)(ord / (uint32_t)cap);
            rank = (int)(ord % (uint32_t)cap);
        }
        const size_t seg = (size_t)b * nc + c;
        write_detection(out_boxes, out_scores, out_classes, out_idx, (size_t)b * R + t, t < nv, boxes_rel + seg * N, keys + seg * N,
                        t < nv ? kept_idx[seg * cap + rank] : -1, c);
    }
    if (tid == 0) {
        out_valid[b] = nv;
        if (out_exhausted) {
            int e = 0;
            for (int c = 0; c < nc; ++c) e |= seg_exhausted[b * nc + c];
            out_exhausted[b] = e;
        }
    }
}

// workspace layout for postprocess: boxes_rel | keys | kept_idx | kept_count | exhausted
struct PostWs {
    float4* boxes_rel;
    uint32_t* keys;
    int32_t* kept_idx;
    int32_t* kept_count;
    int32_t* exhausted;
    size_t bytes;
};
static PostWs carve_post_ws(void* base, int B, int N, int nc, int cap) {
    PostWs w;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void* p = base ? static_cast<char*>(base) + off : nullptr;
        off += align_up(bytes, 256);
        return p;
    };
    const size_t segs = (size_t)B * nc;
    w.boxes_rel = static_cast<float4*>(take(sizeof(float4) * segs * N));
    w.keys = static_cast<uint32_t*>(take(sizeof(uint32_t) * segs * N));
    w.kept_idx = static_cast<int32_t*>(take(sizeof(int32_t) * segs * (size_t)(cap > 0 ? cap : 1)));
    w.kept_count = static_cast<int32_t*>(take(sizeof(int32_t) * segs));
    w.exhausted = static_cast<int32_t*>(take(sizeof(int32_t) * segs));
    w.bytes = off;
    return w;
}

static int launch_nms(const float4* boxes, const uint32_t* keys, int segments, int N, int cap, int top_n, float iou_thr,
                      int32_t* kept_idx, int32_t* kept_count, int32_t* exhausted, cudaStream_t st,
                      const NmsFinalize& fin = NmsFinalize{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0}) {
    FRCNN_REQUIRE(cap <= 4096, FRCNN_ELIMIT, "nms: min(max_output, N) = %d exceeds the 4096 limit", cap);
    FRCNN_REQUIRE(N < (1 << 22), FRCNN_ELIMIT, "nms: N = %d exceeds the 2^22 limit", N);
    size_t smem = nms_smem_bytes(cap);
    // the segment's keys stay in shared memory across the sweeps when they fit beside the lists (227 KB per CTA on sm_100)
    const size_t with_keys = smem + sizeof(uint32_t) * (size_t)N + 16;
    const int keys_in_smem = with_keys <= (size_t)227 * 1024 && N <= 15 * 4 * kNmsThreads ? 1 : 0;  // (64-bit taker masks)
    if (keys_in_smem) smem = with_keys;
    FRCNN_CUDA(ensure_dynamic_smem<0>((const void*)nms_select_kernel, smem));
    nms_select_kernel<<<segments, kNmsThreads, smem, st>>>(boxes, keys, N, cap, top_n, iou_thr, kept_idx, kept_count,
                                                           exhausted, keys_in_smem, fin);
    FRCNN_LAUNCH_CHECK("nms_select_kernel");
    return 0;
}

}  // namespace frcnn

using namespace frcnn;

extern "C" {

size_t frcnn_postprocess_workspace_bytes(int B, int N, int nc, int max_per_class, int /*max_total*/) {
    if (B <= 0 || N < 0 || nc <= 0) return 0;
    const int cap = max_per_class < N ? max_per_class : N;
    return carve_post_ws(nullptr, B, N, nc, cap).bytes;
}

static int launch_decode_filter(float img_h, float img_w, const float* regions, int regions_batched,
                                const float* pred_scores, int logits, float* softmax_out, const float* pred_boxes, int B,
                                int N, int nc, float score_threshold, float* boxes_rel, uint32_t* keys,
                                frcnn_stream_t stream) {
    FRCNN_REQUIRE(B > 0 && N >= 0 && nc > 0, FRCNN_EINVAL, "decode_filter: bad shape B=%d N=%d nc=%d", B, N, nc);
    FRCNN_REQUIRE(aligned16(regions) && aligned16(pred_boxes) && aligned16(boxes_rel) &&
                      (reinterpret_cast<uintptr_t>(pred_scores) & 7u) == 0 &&
                      (reinterpret_cast<uintptr_t>(softmax_out) & 7u) == 0,
                  FRCNN_EINVAL, "decode_filter: pointers must be 16-byte aligned");
    FRCNN_REQUIRE(logits || !softmax_out, FRCNN_EINVAL, "decode_filter: softmax_out needs logits input");
    const int64_t total = (int64_t)B * N * nc;
    if (total == 0) return 0;
    const unsigned grid = (unsigned)((total + 255) / 256);
#define FRCNN_DF(S, L)                                                                                                  \
    decode_filter_kernel<S, L><<<grid, 256, 0, as_stream(stream)>>>(img_h, img_w, (const float4*)regions, regions_batched, \
                                                                    pred_scores, (const float4*)pred_boxes, B, N, nc,  \
                                                                    score_threshold, (float4*)boxes_rel, keys, softmax_out)
    if (nc == 1) {
        if (logits) FRCNN_DF(true, true); else FRCNN_DF(true, false);
    } else {
        if (logits) FRCNN_DF(false, true); else FRCNN_DF(false, false);
    }
#undef FRCNN_DF
    FRCNN_LAUNCH_CHECK("decode_filter_kernel");
    return 0;
}

int frcnn_decode_filter(float img_h, float img_w, const float* regions, int regions_batched, const float* pred_scores,
                        const float* pred_boxes, int B, int N, int nc, float score_threshold, float* boxes_rel,
                        uint32_t* keys, frcnn_stream_t stream) {
    return launch_decode_filter(img_h, img_w, regions, regions_batched, pred_scores, 0, nullptr, pred_boxes, B, N, nc,
                                score_threshold, boxes_rel, keys, stream);
}

size_t frcnn_nms_workspace_bytes(int segments, int N) {
    if (segments <= 0 || N < 0) return 0;
    return align_up(sizeof(uint32_t) * (size_t)segments * N, 256);
}

int frcnn_nms(const float* boxes, const float* scores, int segments, int N, float score_threshold, float iou_threshold,
              int max_output, int pre_nms_top_n, int32_t* kept_idx, int32_t* kept_count, int32_t* exhausted,
              void* workspace, size_t workspace_bytes, frcnn_stream_t stream) {
    FRCNN_REQUIRE(segments > 0 && N >= 0 && max_output >= 0, FRCNN_EINVAL, "nms: bad shape");
    FRCNN_REQUIRE(aligned16(boxes) && kept_idx && kept_count, FRCNN_EINVAL, "nms: bad pointers");
    FRCNN_REQUIRE(workspace_bytes >= frcnn_nms_workspace_bytes(segments, N) && (workspace || N == 0), FRCNN_EWORKSPACE,
                  "nms: workspace too small");
    const int cap = max_output < N ? max_output : N;
    uint32_t* keys = static_cast<uint32_t*>(workspace);
    const int64_t total = (int64_t)segments * N;
    if (total > 0) {
        score_keys_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(scores, total,
                                                                                         score_threshold, keys);
        FRCNN_LAUNCH_CHECK("score_keys_kernel");
    }
    return launch_nms((const float4*)boxes, keys, segments, N, cap, pre_nms_top_n, iou_threshold, kept_idx, kept_count,
                      exhausted, as_stream(stream));
}

static int postprocess_impl(float img_h, float img_w, const float* regions, int regions_batched, const float* pred_scores,
                            int logits, float* softmax_out, const float* pred_boxes, int B, int N, int nc,
                            float score_threshold, float iou_threshold, int max_per_class, int max_total, int pre_nms_top_n,
                            float* out_boxes, float* out_scores, int32_t* out_classes, int32_t* out_valid, int32_t* out_idx,
                            int32_t* out_exhausted, void* workspace, size_t workspace_bytes, frcnn_stream_t stream) {
    FRCNN_REQUIRE(B > 0 && N >= 0 && nc > 0 && max_per_class >= 0 && max_total > 0, FRCNN_EINVAL,
                  "postprocess: bad shape B=%d N=%d nc=%d per_class=%d total=%d", B, N, nc, max_per_class, max_total);
    FRCNN_REQUIRE(aligned16(out_boxes) && out_scores && out_classes && out_valid, FRCNN_EINVAL,
                  "postprocess: bad output pointers");
    const int cap = max_per_class < N ? max_per_class : N;  // min(max_output_size_per_class, num_boxes) (T17)
    FRCNN_REQUIRE((int64_t)nc * cap <= 16384, FRCNN_ELIMIT, "postprocess: nc*cap = %lld exceeds the 16384 limit",
                  (long long)nc * cap);
    FRCNN_REQUIRE(workspace_bytes >= frcnn_postprocess_workspace_bytes(B, N, nc, max_per_class, max_total) && workspace,
                  FRCNN_EWORKSPACE, "postprocess: workspace too small");
    cudaStream_t st = as_stream(stream);
    PostWs w = carve_post_ws(workspace, B, N, nc, cap);
    int rc = launch_decode_filter(img_h, img_w, regions, regions_batched, pred_scores, logits, softmax_out, pred_boxes, B,
                                  N, nc, score_threshold, (float*)w.boxes_rel, w.keys, stream);
    if (rc) return rc;
    if (nc == 1)  // one class per image: the NMS CTA writes the padded outputs itself (no finalize launch)
        return launch_nms(w.boxes_rel, w.keys, B, N, cap, pre_nms_top_n, iou_threshold, w.kept_idx, w.kept_count, w.exhausted, st,
                          NmsFinalize{(float4*)out_boxes, out_scores, out_classes, out_valid, out_idx, out_exhausted, max_total});
    rc = launch_nms(w.boxes_rel, w.keys, B * nc, N, cap, pre_nms_top_n, iou_threshold, w.kept_idx, w.kept_count,
                    w.exhausted, st);
    if (rc) return rc;
    int p2 = 0;
    size_t smem = 0;
    if (nc > 1) {
        p2 = 32;
        while (p2 < nc * cap) p2 <<= 1;
        smem = sizeof(unsigned long long) * (size_t)p2;
        if (smem > 48 * 1024) FRCNN_CUDA(ensure_dynamic_smem<1>((const void*)finalize_kernel, smem));
    }
    finalize_kernel<<<B, nc > 1 ? 1024 : 256, smem, st>>>(w.boxes_rel, w.keys, N, nc, cap > 0 ? cap : 1, max_total,
                                                          w.kept_idx, w.kept_count, w.exhausted, (float4*)out_boxes,
                                                          out_scores, out_classes, out_valid, out_idx, out_exhausted,
                                                          p2);
    FRCNN_LAUNCH_CHECK("finalize_kernel");
    return 0;
}

int frcnn_postprocess(float img_h, float img_w, const float* regions, int regions_batched, const float* pred_scores,
                      const float* pred_boxes, int B, int N, int nc, float score_threshold, float iou_threshold,
                      int max_per_class, int max_total, int pre_nms_top_n, float* out_boxes, float* out_scores,
                      int32_t* out_classes, int32_t* out_valid, int32_t* out_idx, int32_t* out_exhausted,
                      void* workspace, size_t workspace_bytes, frcnn_stream_t stream) {
    return postprocess_impl(img_h, img_w, regions, regions_batched, pred_scores, 0, nullptr, pred_boxes, B, N, nc,
                            score_threshold, iou_threshold, max_per_class, max_total, pre_nms_top_n, out_boxes, out_scores,
                            out_classes, out_valid, out_idx, out_exhausted, workspace, workspace_bytes, stream);
}

int frcnn_postprocess_logits(float img_h, float img_w, const float* regions, int regions_batched, const float* pred_logits,
                             float* softmax_out, const float* pred_boxes, int B, int N, int nc, float score_threshold,
                             float iou_threshold, int max_per_class, int max_total, int pre_nms_top_n, float* out_boxes,
                             float* out_scores, int32_t* out_classes, int32_t* out_valid, int32_t* out_idx,
                             int32_t* out_exhausted, void* workspace, size_t workspace_bytes, frcnn_stream_t stream) {
    return postprocess_impl(img_h, img_w, regions, regions_batched, pred_logits, 1, softmax_out, pred_boxes, B, N, nc,
                            score_threshold, iou_threshold, max_per_class, max_total, pre_nms_top_n, out_boxes, out_scores,
                            out_classes, out_valid, out_idx, out_exhausted, workspace, workspace_bytes, stream);
}

}  // extern "C"

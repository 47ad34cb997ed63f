// Row-streaming RoI pooling for the reference configuration (pooled_size 7, kernel_size 2: a 14 x 14 crop_and_resize
// max-pooled 2 x 2; models/detectors/fast_rcnn_detector.py:134-177), forward and backward.
//
// This file holds what the two persistent kernels share -- the sampling arithmetic, the packed fp32 lerp, the crop-row
// fold (forward) and the gradient routing (backward), the TMA / mbarrier wrappers, the host-side tensor maps -- and the
// launchers; the kernels themselves live in roi_pool_persist.cuh (records + forward) and roi_pool_persist_bwd.cuh.
//
// The walk (both directions): a warp owns 64 channels (one float2 per lane) of one RoI.
//   * The feature columns [cmin, cmax] under the 14 x-samples select one of eight TMA tensor maps (box = 64 channels x W
//     pixels, W >= the span); the 14 y-samples become the list of feature rows the downward walk needs, in order, and a
//     64-bit bit string that says after which staged row each crop row is complete (roi_records_kernel, once per RoI).
//   * Forward: ONE cp.async.bulk.tensor per feature row (W consecutive pixels x 64 channels of the [B*Hf*Wf, C] view of the
//     NHWC maps) into a slot of the warp's private shared-memory ring; a staged row is x-interpolated ONCE at all 14
//     sample positions into one of two register row sets.  Walk position k lands in set k & 1, so the newest row is
//     always the bottom neighbour; each finished crop row is 14 y-interpolations folded straight into the 7 running
//     bin maxima, stored after every second crop row.  Every register index is static, all control flow warp-uniform.
//   * Backward: the mirror image -- the two register sets ACCUMULATE the routed gradient per x-sample, and a feature row
//     the walk has left is scattered along x and added to dFeature with one cp.reduce.async.bulk.tensor.
// Per sample the fp32 operations and their order are those of TF's CPU kernel (t = TL + (TR - TL) * xl, b likewise,
// out = t + (b - t) * yl, each op rounded separately; packed FADD2/FMUL2/FFMA2 with a run-time 1.0, see lerp2): forward
// results are bit-identical to the oracle.  Compared with the bin-row walk (roi_pool.cu) every feature pixel is read
// from L2 once per RoI instead of once per bin-row, and a row is x-interpolated once per RoI.
// RoIs whose x-samples span more than 51 feature columns do not fit the ring; they take a plain gather path.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stddef.h>

#include <atomic>

#include "roi_axis.cuh"

namespace frcnn {

namespace {

constexpr int kS = 14;                         // crop size = pooled_size * kernel_size
constexpr int kP = 7;
constexpr int kWarpCh = 64;                    // channels per warp: one float2 per lane
constexpr int kPixBytes = kWarpCh * 4;         // one pixel in a warp's ring
constexpr int kRingPix = 52;                   // pixels per warp ring (forward): 13 KB per warp, 16 warps per SM
constexpr int kWarpRingBytes = kRingPix * kPixBytes;
constexpr int kNumBoxes = 8;
// box widths: the smallest W with floor(52 / W) = 8, 7, 6, 5, 4, 3, 2, 1 slots
__host__ __device__ constexpr int box_width(int i) {
    return i == 0 ? 6 : i == 1 ? 7 : i == 2 ? 8 : i == 3 ? 10 : i == 4 ? 13 : i == 5 ? 17 : i == 6 ? 26 : 52;
}


struct TensorMaps {
    CUtensorMap box[kNumBoxes];
};


__device__ __constant__ float kOne = 1.0f;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
#ifdef FRCNN_DEBUG_SPIN
    // debug build: a wait that does not complete within 50 ms reports itself and traps instead of hanging the GPU
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        unsigned ok;
        asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity)
                     : "memory");
        if (ok) return;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 50000000ull) {
            if ((threadIdx.x & 31) == 0)
                printf("mbar_wait stuck: block %d warp %d bar %u parity %u\n", blockIdx.x, threadIdx.x >> 5, bar, parity);
            __trap();
        }
    }
#else
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "RS_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni RS_DONE;\n"
        "bra.uni RS_WAIT;\n"
        "RS_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
#endif
}

// box (c0 = first channel, c1 = first pixel) of the [pixels, channels] view -> shared memory, completion on `bar`
__device__ __forceinline__ void tma_box_g2s(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}

typedef unsigned long long u64;  // two packed fp32 channels

__device__ __forceinline__ u64 pack2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 unpack2(u64 v) {
    float2 r;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
// a + (b - a) * w on two packed channels: sub.rn, mul.rn and add.rn rounded separately.  The add is written as
// fma(p, ONE, a) with ONE = 1.0 read at run time so that ptxas can neither fold it nor contract mul+add into one
// rounding (it does contract mul.rn.f32x2 + add.rn.f32x2 even under -fmad=false); p * 1 + a rounds exactly like p + a.
__device__ __forceinline__ u64 lerp2(u64 a, u64 b, u64 ww, u64 one) {
    u64 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(b), "l"(a));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(d), "l"(ww));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(d), "l"(one), "l"(a));
    return d;
}

// The same with a sub that the compiler may not hoist or share: in the walk two or three consecutive crop rows often sit
// between the same pair of feature rows, and hoisting the 14 packed differences (b - a) out of the crop-row loop keeps
// 28 more registers alive across it -- exactly the registers the walk state needs.  Recomputing them costs a few percent
// more packed adds and removes the spills.
__device__ __forceinline__ u64 lerp2_nohoist(u64 a, u64 b, u64 ww, u64 one) {
    u64 d;
    asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(b), "l"(a));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(d), "l"(ww));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(d), "l"(one), "l"(a));
    return d;
}

// One crop_and_resize sampling position along an axis, in registers; op order of crop_and_resize_op.cc (= axis_sample).
struct Sample {
    int lo, hi;
    float lerp;
    bool ok;
};
__device__ __forceinline__ Sample sample_axis(float c1, float c2, int extent, int i) {
    const float em1 = (float)(extent - 1);
    const float scale = fdiv(fmul(fsub(c2, c1), em1), (float)(kS - 1));
    const float in = fadd(fmul(c1, em1), fmul((float)i, scale));
    Sample s;
    s.ok = (in >= 0.0f) && (in <= em1);  // NaN -> not sampled
    const float fl = s.ok ? floorf(in) : 0.0f;
    s.lo = (int)fl;
    s.hi = s.ok ? (int)ceilf(in) : 0;
    s.lerp = s.ok ? fsub(in, fl) : 0.0f;
    return s;
}

// Feature columns the 14 x-samples of an RoI need (full warp; lanes 0..13 hold the samples).  The right neighbour of a
// sample is read one pixel after the left one even at an integer position (weight exactly 0), so a sample sitting
// exactly on the last needed column asks for one more pixel (real data of the map, or TMA zero fill past its end).
struct Span {
    int cmin, need;  // first column, pixels to stage (0: no sample inside the map)
    bool any_outside;
};
__device__ __forceinline__ Span column_span(const Sample& xs, int lane) {
    const unsigned all = 0xffffffffu;
    const bool ok = lane < kS && xs.ok;
    const int cmin = __reduce_min_sync(all, ok ? xs.lo : 0x7fffffff);
    const int cmax = __reduce_max_sync(all, ok ? xs.hi : -1);
    const bool spare = __any_sync(all, ok && xs.hi == xs.lo && xs.lo == cmax) != 0;
    Span s;
    s.cmin = cmin;
    s.need = cmax < 0 ? 0 : cmax - cmin + 1 + (spare ? 1 : 0);
    s.any_outside = __any_sync(all, lane < kS && !xs.ok) != 0;
    return s;
}

// Strict-greater update of a running maximum and its 2-bit window position (first maximum in row-major window order
// wins: trap T8).  `pos` = dy * 2 + dx of the candidate; the code of channel lane `c` of bin j sits at bit 4 j + 2 c.
__device__ __forceinline__ void take_if_greater(float& m, unsigned& cw, float v, unsigned pos, int shift) {
    if (v > m) {
        m = v;
        cw = (cw & ~(3u << shift)) | (pos << shift);
    }
}

// Arg-max tracking of the streaming walk: instead of rewriting a 2-bit position per candidate, every candidate after the
// first only records "I beat the running maximum" in a flag bit (one predicated OR) and the maximum itself is a plain
// FMNMX.  The last candidate that won is the first maximum in row-major window order (strict >: trap T8), so the 2-bit
// code is rebuilt from the three flags with word-wide logic once per bin-row (codes_from_flags).
//   word a: bit 4j + 2c     window position 1 won,   bit 4j + 2c + 1   position 2 won      (c = channel lane 0 / 1)
//   word b: bit 4j + 2c     window position 3 won
__device__ __forceinline__ void flag_if_greater(float& m, unsigned& word, float v, unsigned bit) {
#ifdef FRCNN_FLAG_IF
    if (v > m) word |= bit;
#else
    asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(word) : "f"(v), "f"(m), "r"(bit));
#endif
    m = fmaxf(m, v);
}
__device__ __forceinline__ unsigned codes_from_flags(unsigned a, unsigned b) {
    const unsigned won2 = (a >> 1) & 0x05555555u;                 // position 2 flags moved to the even bits
    const unsigned lo = b | (a & 0x05555555u & ~won2);            // low code bit: position 3, or position 1 not beaten by 2
    return lo | ((won2 | b) << 1);                                // high code bit: position 2 or 3
}

// Fold one finished crop row (top / bottom feature rows T / Bm, weight yl) into the 7 running bin maxima.  kFirst: first
// crop row of the bin-row.  kCodes: also track which of the four window samples holds the maximum (training forward).
template <bool kFirst, bool kCodes, bool kNoHoist = false>
__device__ __forceinline__ void crop_row(const u64 (&T)[kS], const u64 (&Bm)[kS], float yl, u64 one, float2 (&M)[kP],
                                         unsigned& ca, unsigned& cb) {
    const u64 ww = pack2(yl, yl);
#pragma unroll
    for (int j = 0; j < kP; ++j) {
        const float2 v0 = unpack2(kNoHoist ? lerp2_nohoist(T[2 * j], Bm[2 * j], ww, one) : lerp2(T[2 * j], Bm[2 * j], ww, one));
        const float2 v1 = unpack2(kNoHoist ? lerp2_nohoist(T[2 * j + 1], Bm[2 * j + 1], ww, one)
                                           : lerp2(T[2 * j + 1], Bm[2 * j + 1], ww, one));
        if (kCodes) {
            if (kFirst) {
                M[j] = v0;  // window position 0; the flags were cleared for this bin-row
                flag_if_greater(M[j].x, ca, v1.x, 1u << (4 * j));
                flag_if_greater(M[j].y, ca, v1.y, 4u << (4 * j));
            } else {
                flag_if_greater(M[j].x, ca, v0.x, 2u << (4 * j));
                flag_if_greater(M[j].y, ca, v0.y, 8u << (4 * j));
                flag_if_greater(M[j].x, cb, v1.x, 1u << (4 * j));
                flag_if_greater(M[j].y, cb, v1.y, 4u << (4 * j));
            }
        } else if (kFirst) {
            M[j].x = fmaxf(v0.x, v1.x);
            M[j].y = fmaxf(v0.y, v1.y);
        } else {
            M[j].x = fmaxf(fmaxf(M[j].x, v0.x), v1.x);
            M[j].y = fmaxf(fmaxf(M[j].y, v0.y), v1.y);
        }
    }
}

#include "roi_pool_persist.cuh"

// ================================================================================================
// Backward from saved arg-max codes: the mirror image of the walk above (models/faster_rcnn.py:73,103 differentiates
// ROIPooling; MaxPoolGrad routes dY to the arg-max sample, CropAndResizeGradImage scatters the four bilinear shares).
//   * dY and the codes of one bin-row (7 bins x 64 channels = 1792 B, 7 x 16 code bytes) arrive by TMA in a 3-slot ring.
//   * The two register row sets now ACCUMULATE: a crop row adds (1 - yl) g to the set of its top feature row and yl g to
//     the set of its bottom row, at its 14 x-sample positions (g = dY of the bin if this sample is the bin's arg-max
//     for the channel, else 0).
//   * When the walk leaves a feature row its set is scattered along x -- sample i gives (1 - xl) G[i] to pixel lo and
//     xl G[i] to pixel lo + 1 of a lane-private shared-memory row buffer (plain read-modify-write, no atomics) -- and
//     ONE cp.reduce.async.bulk.tensor (TMA reduce-add, fp32) adds the buffer to dFeature: W pixels x 64 channels per
//     instruction instead of W x 16 vector atomics.
// The gradient has no reference summation order (TF: atomics on GPU, a sequential loop on CPU); the bar is 1e-5
// relative, so contributions are fused multiply-adds and the y-sum is formed before the x-weights are applied.
// RoIs that do not fit (x-samples more than 29 columns apart) take a plain atomic path.
// ================================================================================================
constexpr int kBwdMaxBufs = 4;
constexpr int kBwdDySlots = 3;
constexpr int kBwdDyBytes = kP * kPixBytes;            // 7 bins x 64 channels x 4 B
constexpr int kBwdCodeBytes = 128;                     // 7 bins x 16 code bytes, padded
constexpr int kBwdNumBoxes = 8;                        // row-buffer widths 6 .. 26 as in the forward; the last map is a 2-pixel box:
                                                       // RoIs of the widest class (and reversed ones) reduce sample by sample
__host__ __device__ constexpr int bwd_box_width(int i) { return i == kBwdNumBoxes - 1 ? 2 : box_width(i); }

struct BwdMaps {
    CUtensorMap dfeat[kBwdNumBoxes];  // [B*Hf*Wf, C] fp32, box 64 x W: reduce-add target
    CUtensorMap dy;                   // [B*R*49, C] fp32, box 64 x 7
    CUtensorMap codes;                // [B*R*49, C/4] u8, box 16 x 7
};


__device__ __forceinline__ void tma_reduce_add_s2g(const CUtensorMap* map, int c0, int c1, unsigned src) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map),
                 "r"(c0), "r"(c1), "r"(src)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_reduce_add_s2g_nocommit(const CUtensorMap* map, int c0, int c1, unsigned src) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map),
                 "r"(c0), "r"(c1), "r"(src)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read(int pending) {  // wait until at most `pending` groups still read smem
    switch (pending) {
        case 0: asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory"); break;
        default: asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory"); break;
    }
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ void sts64(unsigned addr, u64 v) { asm volatile("st.shared.b64 [%0], %1;" ::"r"(addr), "l"(v) : "memory"); }


// One crop row inside the map: route dY of its 7 bins to the samples that won the max, weight by (1 - yl) / yl.
// One copy of this code serves every crop row (the hot loop should stay small: the instruction caches are 6 KB / 32 KB):
//   g0_top  which register set holds the crop row's top feature row (the walk alternates) -- only swaps the two weights;
//   second  second crop row of its bin-row -- flips the row bit of all codes once instead of comparing per sample.
// acc_top += g * w_top, acc_bot += g * w_bot iff the sample's code field is clear (it holds the window maximum)
__device__ __forceinline__ void routed_fma(unsigned field, float g, float w_top, float w_bot, float& acc_top, float& acc_bot) {
    asm("{\n\t"
        ".reg .pred p;\n\t"
        "setp.eq.u32 p, %2, 0;\n\t"
        "@p fma.rn.f32 %0, %3, %4, %0;\n\t"
        "@p fma.rn.f32 %1, %3, %5, %1;\n\t"
        "}"
        : "+f"(acc_top), "+f"(acc_bot)
        : "r"(field), "f"(g), "f"(w_top), "f"(w_bot));
}

__device__ __forceinline__ void bwd_crop_row(u64 (&G0)[kS], u64 (&G1)[kS], bool g0_top, bool second, float yl, unsigned dy_addr,
                                             unsigned cw) {
    const float wt = fsub(1.0f, yl);
    const float w0f = g0_top ? wt : yl, w1f = g0_top ? yl : wt;
    const u64 w0 = pack2(w0f, w0f), w1 = pack2(w1f, w1f);
    // window position = 2 * (crop row in the bin) + (crop column in the bin).  c0 / c1 hold the codes XOR the position of
    // this crop row's first / second column, so "this sample won" is "its two code bits are clear": one predicate-producing
    // LOP3 and one FSEL per sample and channel (no shift / compare).
    const unsigned c0 = second ? (cw ^ 0x0AAAAAAAu) : cw;
    const unsigned c1 = c0 ^ 0x05555555u;
#pragma unroll
    for (int j = 0; j < kP; ++j) {
        u64 dyv;
        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(dyv) : "r"(dy_addr + j * kPixBytes));
        const float2 d = unpack2(dyv);
        const unsigned mx = 3u << (4 * j), my = 12u << (4 * j);
#ifdef FRCNN_AB_BWD_FSEL  // A/B: select the gradient (LOP3 + FSEL per sample and channel), packed FFMA2 on both channels
        const u64 g0 = pack2((c0 & mx) == 0u ? d.x : 0.0f, (c0 & my) == 0u ? d.y : 0.0f);
        const u64 g1 = pack2((c1 & mx) == 0u ? d.x : 0.0f, (c1 & my) == 0u ? d.y : 0.0f);
        G0[2 * j] = fma2(g0, w0, G0[2 * j]);
        G1[2 * j] = fma2(g0, w1, G1[2 * j]);
        G0[2 * j + 1] = fma2(g1, w0, G0[2 * j + 1]);
        G1[2 * j + 1] = fma2(g1, w1, G1[2 * j + 1]);
#else
        // One predicate per sample and channel guards the two scalar FMAs of that sample: 3 instructions where the select
        // form needs LOP3 + FSEL + its share of two FFMA2 plus the copy ptxas adds to half of the packed results.
        float2 a0 = unpack2(G0[2 * j]), b0 = unpack2(G1[2 * j]), a1 = unpack2(G0[2 * j + 1]), b1 = unpack2(G1[2 * j + 1]);
        routed_fma(c0 & mx, d.x, w0f, w1f, a0.x, b0.x);
        routed_fma(c0 & my, d.y, w0f, w1f, a0.y, b0.y);
        routed_fma(c1 & mx, d.x, w0f, w1f, a1.x, b1.x);
        routed_fma(c1 & my, d.y, w0f, w1f, a1.y, b1.y);
        G0[2 * j] = pack2(a0.x, a0.y);
        G1[2 * j] = pack2(b0.x, b0.y);
        G0[2 * j + 1] = pack2(a1.x, a1.y);
        G1[2 * j + 1] = pack2(b1.x, b1.y);
#endif
    }
}

__device__ __forceinline__ u64 lds64(unsigned addr) {
    u64 v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
    return v;
}

#include "roi_pool_persist_bwd.cuh"

// Host side: the eight tensor maps of a feature tensor, cached by (pointer, shape).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encoder(EncodeTiledFn* out) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        FRCNN_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        FRCNN_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, FRCNN_EINVAL,
                      "roi_pool: cuTensorMapEncodeTiled not available in this driver");
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    *out = encode;
    return 0;
}

// row-major [rows, cols] matrix of `esize`-byte elements; box = box_cols x box_rows
int encode_2d(CUtensorMap* map, const void* ptr, CUtensorMapDataType dt, int esize, long long rows, long long cols, int box_cols,
              int box_rows) {
    EncodeTiledFn encode = nullptr;
    const int rc = get_encoder(&encode);
    if (rc) return rc;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * esize};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult cr = encode(map, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FRCNN_REQUIRE(cr == CUDA_SUCCESS, FRCNN_EINVAL, "roi_pool: cuTensorMapEncodeTiled failed (CUresult %d)", (int)cr);
    return 0;
}

struct MapCacheEntry {
    const void* ptr = nullptr;
    long long pixels = 0;
    int C = 0, dev = -1;
    TensorMaps maps;
};

int get_tensor_maps(const float* feat, long long pixels, int C, int dev, const TensorMaps** out) {
    static thread_local MapCacheEntry cache[16];
    static thread_local int next = 0;
    for (auto& e : cache)
        if (e.ptr == feat && e.pixels == pixels && e.C == C && e.dev == dev) {
            *out = &e.maps;
            return 0;
        }
    MapCacheEntry& e = cache[next];
    next = (next + 1) % 16;
    e.ptr = nullptr;
    for (int i = 0; i < kNumBoxes; ++i) {
        const int rc = encode_2d(&e.maps.box[i], feat, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, pixels, C, kWarpCh, box_width(i));
        if (rc) return rc;
    }
    e.ptr = feat;
    e.pixels = pixels;
    e.C = C;
    e.dev = dev;
    *out = &e.maps;
    return 0;
}

struct BwdCacheEntry {
    const void *dfeat = nullptr, *dy = nullptr, *codes = nullptr;
    long long pixels = 0, bins = 0;
    int C = 0, dev = -1;
    BwdMaps maps;
};

int get_bwd_maps(float* dfeat, long long pixels, const float* dy, const unsigned char* codes, long long bins, int C, int dev,
                 const BwdMaps** out) {
    static thread_local BwdCacheEntry cache[16];
    static thread_local int next = 0;
    for (auto& e : cache)
        if (e.dfeat == dfeat && e.dy == dy && e.codes == codes && e.pixels == pixels && e.bins == bins && e.C == C && e.dev == dev) {
            *out = &e.maps;
            return 0;
        }
    BwdCacheEntry& e = cache[next];
    next = (next + 1) % 16;
    e.dfeat = nullptr;
    for (int i = 0; i < kBwdNumBoxes; ++i) {
        const int rc = encode_2d(&e.maps.dfeat[i], dfeat, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, pixels, C, kWarpCh, bwd_box_width(i));
        if (rc) return rc;
    }
    int rc = encode_2d(&e.maps.dy, dy, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, bins, C, kWarpCh, kP);
    if (rc) return rc;
    rc = encode_2d(&e.maps.codes, codes, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, bins, C / 4, kWarpCh / 4, kP);
    if (rc) return rc;
    e.dfeat = dfeat;
    e.dy = dy;
    e.codes = codes;
    e.pixels = pixels;
    e.bins = bins;
    e.C = C;
    e.dev = dev;
    *out = &e.maps;
    return 0;
}

}  // namespace

// (512: the shapes this path is tuned and cross-checked for; narrower maps keep the bin-row walk)
bool roi_pool_stream_supported(int C, int P, int k) { return C % 512 == 0 && C % kWarpCh == 0 && P == kP && k == 2; }

static std::atomic<int> g_roi_pool_max_sms{0};
void roi_pool_set_max_sms(int max_sms) { g_roi_pool_max_sms.store(max_sms > 0 ? max_sms : 0, std::memory_order_relaxed); }

size_t roi_pool_stream_workspace_bytes(int B, int R) { return 256 + (size_t)kRecStride * (size_t)B * (size_t)R; }

// The records pre-kernel alone (frcnn_roi_pool_prepare): a caller that pipelines steps runs it with the proposal path, off
// the pooling stream, and launches the pooling kernels with mode kPoolPrepared.
int launch_roi_records(int B, int Hf, int Wf, const float* rois, int R, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    FRCNN_REQUIRE(workspace && workspace_bytes >= roi_pool_stream_workspace_bytes(B, R) && aligned16(workspace), FRCNN_EWORKSPACE,
                  "roi_pool_prepare: workspace too small (frcnn_roi_pool_workspace_bytes)");
    FRCNN_REQUIRE(Hf <= 65535 && Wf <= 65535, FRCNN_ELIMIT, "roi_pool_prepare: feature map side beyond 65535");
    unsigned* counter = static_cast<unsigned*>(workspace);
    RoiRecord* recs = reinterpret_cast<RoiRecord*>(static_cast<unsigned char*>(workspace) + 256);
    const int n_rois = B * R;
    roi_records_kernel<<<(n_rois + 3) / 4, 128, 0, st>>>((const float4*)rois, n_rois, R, Hf, Wf, recs, counter);
    FRCNN_LAUNCH_CHECK("roi_records_kernel");
    return 0;
}

// Persistent forward: records pre-kernel (unless prepared: the workspace already holds this batch's records and fresh
// counters) + warp-persistent pooling kernel.  workspace: [forward counter, backward counter | records].
int launch_roi_pool_fwd_persist(const float* feat, int B, int Hf, int Wf, int C, const float* rois, int R, float* out,
                                unsigned char* codes, void* workspace, size_t workspace_bytes, cudaStream_t st, bool prepared) {
    FRCNN_REQUIRE(workspace && workspace_bytes >= roi_pool_stream_workspace_bytes(B, R) && aligned16(workspace), FRCNN_EWORKSPACE,
                  "roi_pool_forward: workspace too small (frcnn_roi_pool_workspace_bytes)");
    FRCNN_REQUIRE(Hf <= 65535 && Wf <= 65535, FRCNN_ELIMIT, "roi_pool_forward: feature map side beyond 65535");
    FRCNN_REQUIRE((int64_t)B * Hf * Wf < (1ll << 31) && (int64_t)B * R * (C / kWarpCh) < (1ll << 30), FRCNN_ELIMIT,
                  "roi_pool_forward: tensor too large");
    const size_t smem = (size_t)kPWarps * kPWarpBytes;
    int dev = 0;
    FRCNN_CUDA(cudaGetDevice(&dev));
    static bool configured[64] = {};
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        FRCNN_CUDA(cudaFuncSetAttribute(roi_pool_fwd_persist_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        FRCNN_CUDA(cudaFuncSetAttribute(roi_pool_fwd_persist_kernel<1024, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        FRCNN_CUDA(cudaFuncSetAttribute(roi_pool_fwd_persist_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        FRCNN_CUDA(cudaFuncSetAttribute(roi_pool_fwd_persist_kernel<1024, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    static int sm_count[64] = {};
    int sms = (dev >= 0 && dev < 64) ? sm_count[dev] : 0;
    if (!sms) {
        FRCNN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if (dev >= 0 && dev < 64) sm_count[dev] = sms;
    }
    const TensorMaps* maps = nullptr;
    const int rc = get_tensor_maps(feat, (long long)B * Hf * Wf, C, dev, &maps);
    if (rc) return rc;
    unsigned* counter = static_cast<unsigned*>(workspace);
    RoiRecord* recs = reinterpret_cast<RoiRecord*>(static_cast<unsigned char*>(workspace) + 256);
    const int n_rois = B * R, n_items = n_rois * (C / kWarpCh);
    const int budget = g_roi_pool_max_sms.load(std::memory_order_relaxed);
    const int ctas = min((n_items + kPWarps - 1) / kPWarps, budget > 0 ? min(budget, sms) : sms);  // one CTA per SM
    if (!prepared) {
        roi_records_kernel<<<(n_rois + 3) / 4, 128, 0, st>>>((const float4*)rois, n_rois, R, Hf, Wf, recs, counter);
        FRCNN_LAUNCH_CHECK("roi_records_kernel");
    }
    if (codes) {
        if (C == 1024) roi_pool_fwd_persist_kernel<1024, true><<<ctas, kPThreads, smem, st>>>(*maps, feat, Hf, Wf, C, recs, n_items, out, codes, counter);
        else roi_pool_fwd_persist_kernel<0, true><<<ctas, kPThreads, smem, st>>>(*maps, feat, Hf, Wf, C, recs, n_items, out, codes, counter);
    } else {
        if (C == 1024) roi_pool_fwd_persist_kernel<1024, false><<<ctas, kPThreads, smem, st>>>(*maps, feat, Hf, Wf, C, recs, n_items, out, nullptr, counter);
        else roi_pool_fwd_persist_kernel<0, false><<<ctas, kPThreads, smem, st>>>(*maps, feat, Hf, Wf, C, recs, n_items, out, nullptr, counter);
    }
    FRCNN_LAUNCH_CHECK("roi_pool_fwd_persist_kernel");
    return 0;
}

// Persistent backward: records pre-kernel + warp-persistent kernel.  dfeat must have been zeroed on the same stream.
int launch_roi_pool_bwd_persist(int B, int Hf, int Wf, int C, const float* rois, int R, const float* dy,
                                const unsigned char* codes, float* dfeat, void* workspace, size_t workspace_bytes, cudaStream_t st,
                                bool prepared) {
    FRCNN_REQUIRE(workspace && workspace_bytes >= roi_pool_stream_workspace_bytes(B, R) && aligned16(workspace), FRCNN_EWORKSPACE,
                  "roi_pool_backward: workspace too small (frcnn_roi_pool_workspace_bytes)");
    FRCNN_REQUIRE(Hf <= 65535 && Wf <= 65535, FRCNN_ELIMIT, "roi_pool_backward: feature map side beyond 65535");
    FRCNN_REQUIRE((int64_t)B * Hf * Wf < (1ll << 31) && (int64_t)B * R * kP * kP < (1ll << 31) &&
                      (int64_t)B * R * (C / kWarpCh) < (1ll << 30),
                  FRCNN_ELIMIT, "roi_pool_backward: tensor too large");
    const size_t smem = (size_t)kPWarps * kPBwdWarpBytes;
    int dev = 0;
    FRCNN_CUDA(cudaGetDevice(&dev));
    static bool configured[64] = {};
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        FRCNN_CUDA(cudaFuncSetAttribute(roi_pool_bwd_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    static int sm_count[64] = {};
    int sms = (dev >= 0 && dev < 64) ? sm_count[dev] : 0;
    if (!sms) {
        FRCNN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if (dev >= 0 && dev < 64) sm_count[dev] = sms;
    }
    const BwdMaps* maps = nullptr;
    const int rc = get_bwd_maps(dfeat, (long long)B * Hf * Wf, dy, codes, (long long)B * R * kP * kP, C, dev, &maps);
    if (rc) return rc;
    unsigned* counter = static_cast<unsigned*>(workspace);
    RoiRecord* recs = reinterpret_cast<RoiRecord*>(static_cast<unsigned char*>(workspace) + 256);
    const int n_rois = B * R, n_items = n_rois * (C / kWarpCh);
    const int budget = g_roi_pool_max_sms.load(std::memory_order_relaxed);
    const int ctas = min((n_items + kPWarps - 1) / kPWarps, budget > 0 ? min(budget, sms) : sms);
    if (!prepared) {
        roi_records_kernel<<<(n_rois + 3) / 4, 128, 0, st>>>((const float4*)rois, n_rois, R, Hf, Wf, recs, counter);
        FRCNN_LAUNCH_CHECK("roi_records_kernel");
    }
    roi_pool_bwd_persist_kernel<<<ctas, kPThreads, smem, st>>>(*maps, Hf, Wf, C, recs, n_items, dy, codes, dfeat, counter + 1);
    FRCNN_LAUNCH_CHECK("roi_pool_bwd_persist_kernel");
    return 0;
}


}  // namespace frcnn

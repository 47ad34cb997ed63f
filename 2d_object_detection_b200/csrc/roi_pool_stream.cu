// Row-streaming RoI pooling forward for the reference configuration (pooled_size 7, kernel_size 2: a 14 x 14
// crop_and_resize max-pooled 2 x 2; models/detectors/fast_rcnn_detector.py:134-177).
//
// CTA = one RoI x 256 channels, 4 warps; a warp owns 64 channels (one float2 per lane; 128 registers, 4 CTAs per SM)
// and runs its OWN copy pipeline -- after the prologue the warps never synchronise with each other.
//   * Prologue (registers + shuffles; warp 0 the columns, warp 1 the rows, in parallel): the feature columns
//     [cmin, cmax] under the 14 x-samples select one of eight TMA tensor maps (box = 64 channels x W pixels, W >= the
//     span); the 14 y-samples become the list of feature rows the downward walk needs, in order, and a 64-bit bit
//     string that says after which staged row each crop row is complete.
//   * Staging: ONE cp.async.bulk.tensor per feature row and warp (box of W consecutive pixels x 64 channels of the
//     [B*Hf*Wf, C] view of the NHWC maps) into slot `k % depth` of the warp's private shared-memory ring, completion
//     on the slot's mbarrier.  The ring depth adapts to the RoI: narrow RoIs (small boxes) get a deep ring.  When a
//     warp has consumed row k it immediately re-arms the slot with row k + depth -- no producer warp, no hand-shake.
//   * A staged row is x-interpolated ONCE at all 14 sample positions into one of two register row sets (operands from
//     shared memory, addresses from a 14-entry table).  Walk position k lands in set k & 1, so the newest row is
//     always the bottom neighbour; each finished crop row is 14 y-interpolations folded straight into the 7 running
//     bin maxima, stored after every second crop row.  All control flow is CTA-uniform, every register index static.
// Per sample the fp32 operations and their order are those of TF's CPU kernel (t = TL + (TR - TL) * xl, b likewise,
// out = t + (b - t) * yl, each op rounded separately; packed FADD2/FMUL2/FFMA2 with a run-time 1.0, see lerp2):
// results are bit-identical to the oracle.  Compared with the bin-row walk (roi_pool.cu) every feature pixel is read
// from L2 once per RoI instead of once per bin-row, a row is x-interpolated once per RoI, and nobody waits on L2.
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
constexpr int kChunk = 256;                    // channels per CTA
constexpr int kWarpCh = 64;                    // channels per warp: one float2 per lane
constexpr int kWarps = kChunk / kWarpCh;
static_assert(kWarps >= 2, "the prologue uses two warps (columns / rows)");
constexpr int kThreads = kWarps * 32;
constexpr int kPixBytes = kWarpCh * 4;         // one pixel in a warp's ring
constexpr int kRingPix = 52;                   // pixels per warp ring: 4 warps x 52 x 256 B = 52 KB per CTA, 208 KB per SM
constexpr int kWarpRingBytes = kRingPix * kPixBytes;
constexpr int kRingBytes = kWarps * kWarpRingBytes;
constexpr int kMaxDepth = 8;
constexpr int kNumBoxes = 8;
// box widths: the smallest W with floor(52 / W) = 8, 7, 6, 5, 4, 3, 2, 1 slots
__host__ __device__ constexpr int box_width(int i) {
    return i == 0 ? 6 : i == 1 ? 7 : i == 2 ? 8 : i == 3 ? 10 : i == 4 ? 13 : i == 5 ? 17 : i == 6 ? 26 : 52;
}


struct TensorMaps {
    CUtensorMap box[kNumBoxes];
};

struct StreamShared {
    int4 xt[kS];          // per crop column: {left pixel byte offset in a slot, 0, x_lerp bits, x_lerp bits}
    float ylerp[kS];
    int rows[2 * kS];     // feature rows in the order the walk consumes them
    int n_rows, box, cmin, depth, slot_bytes, first_px, wide;
    unsigned prog_lo, prog_hi;  // the walk as a bit string, consumed LSB first: 0 = stage-and-interpolate the next feature
                                // row, 1 = the next crop row is complete (both its rows are in registers): emit it
    int n_before, n_after;      // crop rows outside the map before / after the rows inside it
    int xlo[kS], xhi[kS], xok[kS], ylo[kS], yhi[kS], yok[kS];  // plain sample tables, used by the wide-RoI path only
    float xlerp[kS];
    unsigned long long full[kWarps][kMaxDepth];
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
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_test(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    return ok != 0;
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

struct Walk {
    StreamShared* sh;
    const unsigned char* ring;  // this lane's channels inside pixel 0 of slot 0 of the warp's ring
    const CUtensorMap* map;
    unsigned full0;             // shared address of this warp's full[0]
    int k, n_rows, depth, slot_bytes, slot, slot_off;
    unsigned phase;
    u64 one;
};

// Stage feature row rows[idx] of the walk into the slot at byte offset slot_off of the warp's ring: one TMA box, issued
// by lane 0 (which re-derives the few parameters it needs from shared memory instead of keeping them in registers).
__device__ __forceinline__ void stage_row(const Walk& w, int idx, int slot, int slot_off, int warp_ch, int b_pix0, int Wf) {
    if ((threadIdx.x & 31) == 0) {
        const StreamShared& sh = *w.sh;
        const unsigned bar = w.full0 + slot * 8;
        const int first_px = sh.first_px;
        mbar_expect_tx(bar, (unsigned)(w.slot_bytes - first_px * kPixBytes));
        tma_box_g2s(smem_u32(w.ring) + slot_off + first_px * kPixBytes, w.map, warp_ch, b_pix0 + sh.cmin + sh.rows[idx] * Wf, bar);
    }
}

// x-interpolate the next staged feature row (row k of the walk) at the 14 sample positions, then re-arm its slot.
// aoff[i] = byte offset of sample i's left pixel in a slot, wx[i] = its x_lerp: both live in registers for the whole RoI.
__device__ __forceinline__ void load_row(Walk& w, u64 (&X)[kS], const int (&aoff)[kS], const float (&wx)[kS], int warp_ch,
                                         int b_pix0, int Wf) {
    mbar_wait(w.full0 + w.slot * 8, w.phase);
#pragma unroll
    for (int i = 0; i < kS; ++i) {
        // aoff[i] is a complete shared-memory address for slot 0; the slot offset is warp-uniform, so the address is one
        // register + one uniform register + an immediate (the right neighbour is the next pixel)
        u64 l, r;
        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(l) : "r"(aoff[i] + w.slot_off));
        asm volatile("ld.shared.b64 %0, [%1+256];" : "=l"(r) : "r"(aoff[i] + w.slot_off));
        X[i] = lerp2(l, r, pack2(wx[i], wx[i]), w.one);
    }
    __syncwarp();  // every lane has consumed its operands: the slot may be overwritten
    if (w.k + w.depth < w.n_rows) stage_row(w, w.k + w.depth, w.slot, w.slot_off, warp_ch, b_pix0, Wf);
    ++w.k;
    w.slot_off += w.slot_bytes;
    if (++w.slot == w.depth) {
        w.slot = 0;
        w.slot_off = 0;
        w.phase ^= 1u;
    }
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

// kC = number of channels when known at compile time (1024: the reference's backbone; store offsets become immediates), 0 = any
// kCodes: training forward, also writes the arg-max codes ([B,R,P,P,C/4] bytes: four 2-bit window positions per byte)
template <int kC, bool kCodes>
__global__ void __launch_bounds__(kThreads, 512 / kThreads)
    roi_pool_fwd_stream_kernel(const __grid_constant__ TensorMaps maps, const float* __restrict__ feat, int Hf, int Wf, int C_,
                               const float4* __restrict__ rois, int R, int nchunk, float* __restrict__ out,
                               unsigned char* __restrict__ codes) {
    const int C = kC ? kC : C_;
    extern __shared__ __align__(128) unsigned char smem[];
    StreamShared& sh = *reinterpret_cast<StreamShared*>(smem);
    unsigned char* ring = smem + (sizeof(StreamShared) + 127) / 128 * 128;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = blockIdx.x / nchunk, ch0 = (blockIdx.x % nchunk) * kChunk, b = blockIdx.y;
    const unsigned full_base = smem_u32(&sh.full[0][0]);

    // every warp initialises the mbarriers of its own ring
    if (lane < kMaxDepth) {
        mbar_init(full_base + (warp * kMaxDepth + lane) * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // Prologue: warp 0 prepares the columns, warp 1 the rows (in parallel).
    if (tid < 32) {
        const float4 roi = rois[(size_t)b * R + r];  // x1,y1,x2,y2 (fast_rcnn_detector.py:157-158 reorders for TF)
        const Sample xs = sample_axis(roi.x, roi.z, Wf, lane);  // lanes 0..13 are the crop columns
        // ---- columns: a slot holds the pixels cmin .. cmin + W - 1 of one feature row; a sample outside the map reads
        // a zeroed pixel in front of them: 0 + (p - 0) * 0.
        const Span sp = column_span(xs, lane);
        const int first_px = sp.any_outside ? 1 : 0;
        int box = 0;
#pragma unroll
        for (int i = 0; i < kNumBoxes; ++i) box += (sp.need > box_width(i)) ? 1 : 0;
        int W = 0;
#pragma unroll
        for (int i = 0; i < kNumBoxes; ++i) W = (box == i) ? box_width(i) : W;
        // does not fit the ring (or no column inside the map at all): plain gather path below
        const bool wide = box >= kNumBoxes || W + first_px > kRingPix || sp.need == 0;
        if (lane < kS) {
            const int lerp = xs.ok ? __float_as_int(xs.lerp) : 0;
            sh.xt[lane] = make_int4(xs.ok ? (xs.lo - sp.cmin + first_px) * kPixBytes : 0, 0, lerp, lerp);
            if (wide) {
                sh.xlo[lane] = xs.lo; sh.xhi[lane] = xs.hi; sh.xok[lane] = xs.ok; sh.xlerp[lane] = xs.lerp;
            }
        }
        if (lane == 0) {
            const int px = W + first_px;
            sh.box = box;
            sh.cmin = sp.cmin;
            sh.first_px = first_px;
            sh.wide = wide ? 1 : 0;
            sh.slot_bytes = px * kPixBytes;
            sh.depth = max(1, min(kMaxDepth, kRingPix / max(px, 1)));
        }
    } else if (tid < 64) {
        const float4 roi = rois[(size_t)b * R + r];
        const unsigned all = 0xffffffffu;
        const Sample ys = sample_axis(roi.y, roi.w, Hf, lane);  // lanes 0..13 are the crop rows
        // ---- rows: the downward walk.  The walk stages feature rows in the order rows[0], rows[1], ...; position k
        // goes to register row set k & 1, so the most recently staged row is always the bottom neighbour and the set
        // it replaces held the old top row.
        const bool v = lane < kS && ys.ok;
        const int c = ys.lo;
        const int pv_i = __shfl_up_sync(all, (int)v, 1);  // (collectives first: no short-circuit around a shuffle)
        const bool pv = pv_i != 0 && lane > 0;
        const int pc = __shfl_up_sync(all, c, 1);
        const bool reuse = v && pv && c == pc;          // same two feature rows as the previous crop row
        const bool shift = v && pv && c == pc + 1;      // one feature row further down: stage one new bottom row
        const bool load2 = v && !reuse && !shift;       // anything else: stage both rows
        const unsigned lt = (1u << lane) - 1u;
        const unsigned m1 = __ballot_sync(all, shift || load2), m2 = __ballot_sync(all, load2);
        const int off = __popc(m1 & lt) + __popc(m2 & lt);
        const unsigned mv = __ballot_sync(all, v);
        const int last = Hf - 1;
        if (load2) {
            sh.rows[off] = c;
            sh.rows[off + 1] = min(c + 1, last);
        } else if (shift) {
            sh.rows[off] = min(c + 1, last);
        }
        // bit position of crop row sy = loads up to and including its bottom row + crop rows inside the map before it
        const int bitpos = off + (load2 ? 2 : shift ? 1 : 0) + __popc(mv & lt);
        const unsigned prog_lo = __reduce_or_sync(all, (v && bitpos < 32) ? (1u << bitpos) : 0u);
        const unsigned prog_hi = __reduce_or_sync(all, (v && bitpos >= 32) ? (1u << (bitpos - 32)) : 0u);
        const int n_before = __popc(__ballot_sync(all, lane < kS && !v && !(mv & lt)));
        if (lane < kS) {
            sh.ylerp[lane] = ys.lerp;
            sh.ylo[lane] = ys.lo; sh.yhi[lane] = ys.hi; sh.yok[lane] = ys.ok;
        }
        if (lane == 0) {
            sh.n_rows = __popc(m1) + __popc(m2);
            sh.prog_lo = prog_lo;
            sh.prog_hi = prog_hi;
            sh.n_before = n_before;
            sh.n_after = kS - n_before - __popc(mv);
        }
    }
    __syncthreads();
    if (sh.wide) {  // rare (x-samples more than 51 columns apart): one bilinear gather per sample, same arithmetic
        const float onef = kOne;
        const u64 one = pack2(onef, onef);
        const float* fb = feat + (size_t)b * Hf * Wf * C + ch0 + tid * 2;
        float2* ob = reinterpret_cast<float2*>(out + ((size_t)b * R + r) * (kP * kP) * C + ch0) + tid;
        unsigned char* cb = kCodes ? codes + ((size_t)b * R + r) * (kP * kP) * (C / 4) + (ch0 + tid * 2) / 4 : nullptr;
#pragma unroll 1
        for (int bin = 0; bin < kP * kP; ++bin) {
            const int py = bin / kP, px = bin - py * kP;
            float2 best = make_float2(0.0f, 0.0f);
            unsigned cw = 0u;
#pragma unroll 1
            for (int s4 = 0; s4 < 4; ++s4) {
                const int sy = 2 * py + (s4 >> 1), sx = 2 * px + (s4 & 1);
                float2 val = make_float2(0.0f, 0.0f);  // extrapolation value
                if (sh.yok[sy] && sh.xok[sx]) {
                    const float* top = fb + (size_t)sh.ylo[sy] * Wf * C;
                    const float* bot = fb + (size_t)sh.yhi[sy] * Wf * C;
                    const size_t xl = (size_t)sh.xlo[sx] * C, xh = (size_t)sh.xhi[sx] * C;
                    const u64 tl = *reinterpret_cast<const u64*>(top + xl), tr = *reinterpret_cast<const u64*>(top + xh);
                    const u64 bl = *reinterpret_cast<const u64*>(bot + xl), br = *reinterpret_cast<const u64*>(bot + xh);
                    const u64 wx = pack2(sh.xlerp[sx], sh.xlerp[sx]), wy = pack2(sh.ylerp[sy], sh.ylerp[sy]);
                    val = unpack2(lerp2(lerp2(tl, tr, wx, one), lerp2(bl, br, wx, one), wy, one));
                }
                if (s4 == 0) {
                    best = val;
                } else {
                    take_if_greater(best.x, cw, val.x, (unsigned)s4, 0);
                    take_if_greater(best.y, cw, val.y, (unsigned)s4, 2);
                }
            }
            __stcs(ob + (size_t)bin * (C / 2), best);
            if (kCodes) {
                const unsigned pw = __shfl_xor_sync(0xffffffffu, cw, 1);
                if (!(tid & 1)) cb[(size_t)bin * (C / 4)] = (unsigned char)((cw & 15u) | ((pw & 15u) << 4));
            }
        }
        return;
    }

    const unsigned all = 0xffffffffu;
    Walk w;
    w.sh = &sh;
    w.full0 = full_base + warp * (kMaxDepth * 8);
    // (reduce_or of a warp-uniform value = the value, in a uniform register: loop control and slot offsets stay uniform)
    w.n_rows = __reduce_or_sync(all, sh.n_rows);
    w.depth = __reduce_or_sync(all, sh.depth);
    w.slot_bytes = __reduce_or_sync(all, sh.slot_bytes);
    w.map = &maps.box[sh.box];
    w.ring = ring + warp * kWarpRingBytes + lane * 8;
    w.k = 0;
    w.slot = 0;
    w.slot_off = 0;
    w.phase = 0;
    {
        const float onef = kOne;
        w.one = pack2(onef, onef);
    }
    const int warp_ch = ch0 + warp * kWarpCh, b_pix0 = b * Hf * Wf;
    if (w.n_rows > 0) {
        for (int s = 0; s < w.depth && s < w.n_rows; ++s) stage_row(w, s, s, s * w.slot_bytes, warp_ch, b_pix0, Wf);
        if (sh.first_px)  // the zero pixel of every slot (each lane later reads only what it wrote itself)
            for (int s = 0; s < w.depth; ++s) *reinterpret_cast<u64*>(const_cast<unsigned char*>(w.ring) + s * w.slot_bytes) = 0ull;
    }
    int aoff[kS];
    float wx[kS];
#pragma unroll
    for (int i = 0; i < kS; ++i) {
        const int4 e = sh.xt[i];
        aoff[i] = e.x + (int)smem_u32(w.ring);
        wx[i] = __int_as_float(e.z);
    }
    u64 prog = ((u64)__reduce_or_sync(all, sh.prog_hi) << 32) | __reduce_or_sync(all, sh.prog_lo);
    const int n_before = sh.n_before, n_after = sh.n_after;

    u64 X0[kS], X1[kS];
#pragma unroll
    for (int i = 0; i < kS; ++i) X0[i] = X1[i] = 0ull;
    float2 M[kP];
    float2* optr = reinterpret_cast<float2*>(out + ((size_t)b * R + r) * (kP * kP) * C + ch0) + tid;
    const int bin_stride = C / 2;  // float2 per bin
    int sy = 0;
    float yl_next = sh.ylerp[0];  // the y weight of the next crop row, loaded one crop row ahead
    unsigned ca = 0u, cb = 0u;  // kCodes: "candidate beat the running maximum" flags of the bin-row (see flag_if_greater)
    unsigned char* cptr = kCodes ? codes + ((size_t)b * R + r) * (kP * kP) * (C / 4) + (ch0 + tid * 2) / 4 : nullptr;
    // One finished crop row: fold it into the bin maxima; after the second row of a bin-row store the 7 bins.
    auto emit = [&](const u64 (&T)[kS], const u64 (&Bm)[kS], bool outside) {
        const float yl = yl_next;
        yl_next = sh.ylerp[min(sy + 1, kS - 1)];
        if (!(sy & 1)) {
            ca = cb = 0u;
            if (outside) {  // crop row outside the map: extrapolation value 0 takes part in the max
#pragma unroll
                for (int j = 0; j < kP; ++j) M[j] = make_float2(0.0f, 0.0f);
            } else {
                crop_row<true, kCodes>(T, Bm, yl, w.one, M, ca, cb);
            }
        } else {
            if (outside) {
#pragma unroll
                for (int j = 0; j < kP; ++j) {
                    if (kCodes) {  // window positions 2 and 3 are both 0: position 2 wins if 0 beats the first row
                        flag_if_greater(M[j].x, ca, 0.0f, 2u << (4 * j));
                        flag_if_greater(M[j].y, ca, 0.0f, 8u << (4 * j));
                    } else {
                        M[j] = make_float2(fmaxf(M[j].x, 0.0f), fmaxf(M[j].y, 0.0f));
                    }
                }
            } else {
                crop_row<false, kCodes>(T, Bm, yl, w.one, M, ca, cb);
            }
#pragma unroll
            for (int j = 0; j < kP; ++j) __stcs(optr + j * bin_stride, M[j]);
            optr += kP * bin_stride;
            if (kCodes) {  // a byte holds four channels = this lane's nibble and its odd neighbour's
                const unsigned cw = codes_from_flags(ca, cb);
                const unsigned pw = __shfl_xor_sync(0xffffffffu, cw, 1);
                if (!(tid & 1)) {
#pragma unroll
                    for (int j = 0; j < kP; ++j)
                        cptr[j * (C / 4)] = (unsigned char)(((cw >> (4 * j)) & 15u) | (((pw >> (4 * j)) & 15u) << 4));
                }
                cptr += kP * (C / 4);
            }
        }
        ++sy;
    };
    for (int i = 0; i < n_before; ++i) emit(X0, X1, true);
#pragma unroll 1
    for (int k = 0; k < w.n_rows; k += 2) {
        load_row(w, X0, aoff, wx, warp_ch, b_pix0, Wf);  // even position: the new bottom row lands in set 0, set 1 holds the top row
        prog >>= 1;
        while (prog & 1ull) {
            emit(X1, X0, false);
            prog >>= 1;
        }
        if (k + 1 < w.n_rows) {
            load_row(w, X1, aoff, wx, warp_ch, b_pix0, Wf);
            prog >>= 1;
            while (prog & 1ull) {
                emit(X0, X1, false);
                prog >>= 1;
            }
        }
    }
    for (int i = 0; i < n_after; ++i) emit(X0, X1, true);
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
constexpr int kBwdRowPix = 30;                         // per-warp row-buffer ring, pixels
constexpr int kBwdMaxBufs = 4;
constexpr int kBwdDySlots = 3;
constexpr int kBwdDyBytes = kP * kPixBytes;            // 7 bins x 64 channels x 4 B
constexpr int kBwdCodeBytes = 128;                     // 7 bins x 16 code bytes, padded
constexpr int kBwdWarpBytes = kBwdRowPix * kPixBytes + 2 * kPixBytes + kBwdDySlots * (kBwdDyBytes + kBwdCodeBytes);
constexpr int kBwdNumBoxes = 8;                        // row-buffer widths 6 .. 26 as in the forward, then the whole ring
__host__ __device__ constexpr int bwd_box_width(int i) { return i < 7 ? box_width(i) : kBwdRowPix; }

struct BwdMaps {
    CUtensorMap dfeat[kBwdNumBoxes];  // [B*Hf*Wf, C] fp32, box 64 x W: reduce-add target
    CUtensorMap dy;                   // [B*R*49, C] fp32, box 64 x 7
    CUtensorMap codes;                // [B*R*49, C/4] u8, box 16 x 7
};

struct BwdShared {
    int4 xt[kS];          // per crop column: {left pixel byte offset in a row buffer, 0, x_lerp bits, x_lerp bits}
    float ylerp[kS];
    int rows[2 * kS];
    int n_rows, box, cmin, W, nbufs, slow;
    unsigned prog_lo, prog_hi, xpair;  // xpair: 3 bits per bin = how its two crop columns relate (see bwd_flush)
    int n_before, n_after;
    int xlo[kS], xhi[kS], xok[kS], ylo[kS], yhi[kS], yok[kS];  // plain sample tables for the slow path
    float xlerp[kS];
    unsigned long long dfull[kWarps][kBwdDySlots];
};

__device__ __forceinline__ void tma_reduce_add_s2g(const CUtensorMap* map, int c0, int c1, unsigned src) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map),
                 "r"(c0), "r"(c1), "r"(src)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
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
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ void sts64(unsigned addr, u64 v) { asm volatile("st.shared.b64 [%0], %1;" ::"r"(addr), "l"(v) : "memory"); }

struct BwdWalk {
    BwdShared* sh;
    const BwdMaps* maps;
    unsigned warp_smem;   // shared address of this warp's region: row buffers | 2 scratch pixels | dY slots | code slots
    unsigned lane_smem;   // warp_smem + this lane's channel offset
    unsigned dfull0;      // shared address of this warp's dfull[0]
    int warp_ch, W, nbufs, buf, buf_off, issued;
    unsigned xpair;       // 3 bits per bin: 0 both columns in the same cell, 1 second column one cell to the right,
                          // 2 both inside the map but further apart, 3 only the first inside, 4 only the second, 5 none
};

// One crop row inside the map: route dY of its 7 bins to the samples that won the max, weight by (1 - yl) / yl.
// One copy of this code serves every crop row (the hot loop should stay small: the instruction caches are 6 KB / 32 KB):
//   g0_top  which register set holds the crop row's top feature row (the walk alternates) -- only swaps the two weights;
//   second  second crop row of its bin-row -- flips the row bit of all codes once instead of comparing per sample.
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
        const u64 g0 = pack2((c0 & mx) == 0u ? d.x : 0.0f, (c0 & my) == 0u ? d.y : 0.0f);
        const u64 g1 = pack2((c1 & mx) == 0u ? d.x : 0.0f, (c1 & my) == 0u ? d.y : 0.0f);
        G0[2 * j] = fma2(g0, w0, G0[2 * j]);
        G1[2 * j] = fma2(g0, w1, G1[2 * j]);
        G0[2 * j + 1] = fma2(g1, w0, G0[2 * j + 1]);
        G1[2 * j + 1] = fma2(g1, w1, G1[2 * j + 1]);
    }
}

__device__ __forceinline__ u64 lds64(unsigned addr) {
    u64 v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
    return v;
}

// The walk has left feature row rows[kk]: scatter its accumulated set along x into a zeroed row buffer -- sample i adds
// (1 - xl) G[i] to pixel lo and xl G[i] to pixel lo + 1; every lane owns its own two channels of the buffer, so these are
// plain shared-memory read-modify-writes in program order -- and hand the buffer to the TMA engine, which adds it to
// dFeature (one cp.reduce.async.bulk.tensor per feature row and warp).
__device__ __forceinline__ void bwd_flush(BwdWalk& w, u64 (&G)[kS], const float (&wx)[kS], const int (&aoff)[kS], int kk,
                                          int b_pix0, int Wf) {
    const int lane = threadIdx.x & 31;
    if (w.issued >= w.nbufs) {  // the buffer about to be reused must have been read by its reduce
        if (lane == 0) bulk_wait_read(w.nbufs - 1);
        __syncwarp();
    }
    unsigned buf = w.lane_smem + w.buf_off;
    asm volatile("" : "+r"(buf));  // keep the address in a register (ptxas otherwise rebuilds it from %tid for every sample)
#pragma unroll 2
    for (int p = 0; p < w.W; ++p) sts64(buf + p * kPixBytes, 0ull);
    // One bin (two crop columns a, b) at a time.  When both columns sit in the same cell or in neighbouring cells -- the
    // common cases -- their pixels are distinct addresses, so the loads of the pair go out together and the chain of
    // dependent shared-memory round trips is 7 long instead of 14; the per-pixel order of additions is unchanged.
    auto one = [&](int i) {
        const float xl = wx[i];
        const float xr = fsub(1.0f, xl);
        const unsigned a = buf + aoff[i];
        const u64 lo = lds64(a), hi = lds64(a + kPixBytes);
        sts64(a, fma2(G[i], pack2(xr, xr), lo));
        sts64(a + kPixBytes, fma2(G[i], pack2(xl, xl), hi));
    };
#pragma unroll
    for (int j = 0; j < kP; ++j) {
        const unsigned pc = (w.xpair >> (3 * j)) & 7u;  // warp-uniform
        const int ia = 2 * j, ib = 2 * j + 1;
        if (pc <= 1u) {
            const float xla = wx[ia], xlb = wx[ib];
            const u64 wa1 = pack2(xla, xla), wb1 = pack2(xlb, xlb);
            const float xra = fsub(1.0f, xla), xrb = fsub(1.0f, xlb);
            const u64 wa0 = pack2(xra, xra), wb0 = pack2(xrb, xrb);
            const unsigned a = buf + aoff[ia];
            if (pc == 0u) {  // same cell: two pixels
                u64 p0 = lds64(a), p1 = lds64(a + kPixBytes);
                p0 = fma2(G[ib], wb0, fma2(G[ia], wa0, p0));
                p1 = fma2(G[ib], wb1, fma2(G[ia], wa1, p1));
                sts64(a, p0);
                sts64(a + kPixBytes, p1);
            } else {  // neighbouring cells: three pixels
                u64 p0 = lds64(a), p1 = lds64(a + kPixBytes), p2 = lds64(a + 2 * kPixBytes);
                p0 = fma2(G[ia], wa0, p0);
                p1 = fma2(G[ib], wb0, fma2(G[ia], wa1, p1));
                p2 = fma2(G[ib], wb1, p2);
                sts64(a, p0);
                sts64(a + kPixBytes, p1);
                sts64(a + 2 * kPixBytes, p2);
            }
        } else {  // further apart, or a column outside the map (which receives no gradient)
            if (pc == 2u || pc == 3u) one(ia);
            if (pc == 2u || pc == 4u) one(ib);
        }
        G[ia] = 0ull;
        G[ib] = 0ull;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the TMA engine
    __syncwarp();
    if (lane == 0)
        tma_reduce_add_s2g(&w.maps->dfeat[w.sh->box], w.warp_ch, b_pix0 + w.sh->cmin + w.sh->rows[kk] * Wf, w.warp_smem + w.buf_off);
    ++w.issued;
    w.buf_off += w.W * kPixBytes;
    if (++w.buf == w.nbufs) {
        w.buf = 0;
        w.buf_off = 0;
    }
}

__global__ void __launch_bounds__(kThreads, 512 / kThreads)
    roi_pool_bwd_stream_kernel(const __grid_constant__ BwdMaps maps, int Hf, int Wf, int C, const float4* __restrict__ rois, int R,
                               int nchunk, const float* __restrict__ dy, const unsigned char* __restrict__ codes,
                               float* __restrict__ dfeat) {
    extern __shared__ __align__(128) unsigned char smem[];
    BwdShared& sh = *reinterpret_cast<BwdShared*>(smem);
    unsigned char* region = smem + (sizeof(BwdShared) + 127) / 128 * 128;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = blockIdx.x / nchunk, ch0 = (blockIdx.x % nchunk) * kChunk, b = blockIdx.y;
    const unsigned dfull_base = smem_u32(&sh.dfull[0][0]);
    const unsigned all = 0xffffffffu;

    if (lane < kBwdDySlots) {  // every warp initialises the mbarriers of its own dY ring
        mbar_init(dfull_base + (warp * kBwdDySlots + lane) * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {  // warp 0: the columns
        const float4 roi = rois[(size_t)b * R + r];
        const Sample xs = sample_axis(roi.x, roi.z, Wf, lane);
        const Span sp = column_span(xs, lane);
        int box = 0;
#pragma unroll
        for (int i = 0; i < kBwdNumBoxes; ++i) box += (sp.need > bwd_box_width(i)) ? 1 : 0;
        int W = 0;
#pragma unroll
        for (int i = 0; i < kBwdNumBoxes; ++i) W = (box == i) ? bwd_box_width(i) : W;
        // how the two crop columns of a bin relate (lane 2j holds the first column of bin j, lane 2j + 1 the second)
        const int lo_n = __shfl_down_sync(all, xs.lo, 1), ok_n = __shfl_down_sync(all, (int)xs.ok, 1);
        const unsigned pcode = (!xs.ok && !ok_n) ? 5u : !ok_n ? 3u : !xs.ok ? 4u : lo_n == xs.lo ? 0u : lo_n == xs.lo + 1 ? 1u : 2u;
        const unsigned xpair = __reduce_or_sync(all, (lane < kS && !(lane & 1)) ? (pcode << (3 * (lane >> 1))) : 0u);
        const bool slow = box >= kBwdNumBoxes || sp.need == 0;  // too wide for a row buffer, or no column inside the map
        if (lane < kS) {
            const int lerp = xs.ok ? __float_as_int(xs.lerp) : 0;
            sh.xt[lane] = make_int4(xs.ok ? (xs.lo - sp.cmin) * kPixBytes : 0, 0, lerp, lerp);
            sh.xlo[lane] = xs.lo; sh.xhi[lane] = xs.hi; sh.xok[lane] = xs.ok; sh.xlerp[lane] = xs.lerp;
        }
        if (lane == 0) {
            sh.xpair = xpair;
            sh.box = box;
            sh.cmin = sp.cmin;
            sh.W = W;
            sh.nbufs = max(1, min(kBwdMaxBufs, kBwdRowPix / max(W, 1)));
            sh.slow = slow ? 1 : 0;
        }
    } else if (tid < 64) {  // warp 1: the downward walk, exactly as in the forward kernel
        const float4 roi = rois[(size_t)b * R + r];
        const Sample ys = sample_axis(roi.y, roi.w, Hf, lane);
        if (lane < kS) {
            sh.ylo[lane] = ys.lo; sh.yhi[lane] = ys.hi; sh.yok[lane] = ys.ok; sh.ylerp[lane] = ys.lerp;
        }
        const bool v = lane < kS && ys.ok;
        const int c = ys.lo;
        const int pv_i = __shfl_up_sync(all, (int)v, 1);
        const bool pv = pv_i != 0 && lane > 0;
        const int pc = __shfl_up_sync(all, c, 1);
        const bool reuse = v && pv && c == pc;
        const bool shift = v && pv && c == pc + 1;
        const bool load2 = v && !reuse && !shift;
        const unsigned lt = (1u << lane) - 1u;
        const unsigned m1 = __ballot_sync(all, shift || load2), m2 = __ballot_sync(all, load2);
        const int off = __popc(m1 & lt) + __popc(m2 & lt);
        const unsigned mv = __ballot_sync(all, v);
        const int last = Hf - 1;
        if (load2) {
            sh.rows[off] = c;
            sh.rows[off + 1] = min(c + 1, last);
        } else if (shift) {
            sh.rows[off] = min(c + 1, last);
        }
        const int bitpos = off + (load2 ? 2 : shift ? 1 : 0) + __popc(mv & lt);
        const unsigned prog_lo = __reduce_or_sync(all, (v && bitpos < 32) ? (1u << bitpos) : 0u);
        const unsigned prog_hi = __reduce_or_sync(all, (v && bitpos >= 32) ? (1u << (bitpos - 32)) : 0u);
        const int n_before = __popc(__ballot_sync(all, lane < kS && !v && !(mv & lt)));
        if (lane == 0) {
            sh.n_rows = __popc(m1) + __popc(m2);
            sh.prog_lo = prog_lo;
            sh.prog_hi = prog_hi;
            sh.n_before = n_before;
            sh.n_after = kS - n_before - __popc(mv);
        }
    }
    __syncthreads();

    const size_t bin0 = ((size_t)b * R + r) * (kP * kP);
    if (sh.slow) {  // rare: one atomic per bilinear share, TF's operation order
        const float* gy = dy + bin0 * C + ch0 + tid * 2;
        const unsigned char* cb = codes + bin0 * (C / 4) + (ch0 + tid * 2) / 4;
        float* gb = dfeat + (size_t)b * Hf * Wf * C + ch0 + tid * 2;
#pragma unroll 1
        for (int bin = 0; bin < kP * kP; ++bin) {
            const int py = bin / kP, px = bin - py * kP;
            const float2 g2 = *reinterpret_cast<const float2*>(gy + (size_t)bin * C);
            const unsigned nib = (cb[(size_t)bin * (C / 4)] >> ((tid & 1) * 4)) & 15u;
#pragma unroll 1
            for (int cidx = 0; cidx < 2; ++cidx) {
                const unsigned pos = (nib >> (2 * cidx)) & 3u;
                const int sy = 2 * py + (int)(pos >> 1), sx = 2 * px + (int)(pos & 1u);
                if (!(sh.yok[sy] && sh.xok[sx])) continue;  // the extrapolation value won: no gradient
                const float g = cidx ? g2.y : g2.x;
                const float yl = sh.ylerp[sy], xl = sh.xlerp[sx];
                const float dtop = fmul(fsub(1.0f, yl), g), dbot = fmul(yl, g);
                float* top = gb + (size_t)sh.ylo[sy] * Wf * C + cidx;
                float* bot = gb + (size_t)sh.yhi[sy] * Wf * C + cidx;
                atomicAdd(top + (size_t)sh.xlo[sx] * C, fmul(fsub(1.0f, xl), dtop));
                atomicAdd(top + (size_t)sh.xhi[sx] * C, fmul(xl, dtop));
                atomicAdd(bot + (size_t)sh.xlo[sx] * C, fmul(fsub(1.0f, xl), dbot));
                atomicAdd(bot + (size_t)sh.xhi[sx] * C, fmul(xl, dbot));
            }
        }
        return;
    }

    BwdWalk w;
    w.sh = &sh;
    w.maps = &maps;
    w.warp_smem = smem_u32(region) + warp * kBwdWarpBytes;
    asm volatile("" : "+r"(w.warp_smem));
    w.lane_smem = w.warp_smem + lane * 8;
    asm volatile("" : "+r"(w.lane_smem));
    w.dfull0 = dfull_base + warp * (kBwdDySlots * 8);
    w.warp_ch = ch0 + warp * kWarpCh;
    w.W = __reduce_or_sync(all, sh.W);
    w.nbufs = __reduce_or_sync(all, sh.nbufs);
    w.buf = 0;
    w.buf_off = 0;
    w.issued = 0;
    w.xpair = __reduce_or_sync(all, sh.xpair);
    const int n_rows = __reduce_or_sync(all, sh.n_rows);
    const int b_pix0 = b * Hf * Wf;
    const unsigned dy_slots = w.warp_smem + (kBwdRowPix + 2) * kPixBytes;
    const unsigned code_slots = dy_slots + kBwdDySlots * kBwdDyBytes;
    // stage dY + codes of bin-row `py` into slot py % 3 (lane 0)
    auto stage_bin_row = [&](int py, int slot) {
        if (lane == 0) {
            const unsigned bar = w.dfull0 + slot * 8;
            mbar_expect_tx(bar, kBwdDyBytes + kP * 16);
            tma_box_g2s(dy_slots + slot * kBwdDyBytes, &maps.dy, w.warp_ch, (int)bin0 + py * kP, bar);
            tma_box_g2s(code_slots + slot * kBwdCodeBytes, &maps.codes, w.warp_ch / 4, (int)bin0 + py * kP, bar);
        }
    };
    for (int py = 0; py < kBwdDySlots; ++py) stage_bin_row(py, py);

    float wx[kS];
    int aoff[kS];
#pragma unroll
    for (int i = 0; i < kS; ++i) {
        const int4 e = sh.xt[i];
        aoff[i] = e.x;
        wx[i] = __int_as_float(e.z);
    }
    u64 prog = ((u64)__reduce_or_sync(all, sh.prog_hi) << 32) | __reduce_or_sync(all, sh.prog_lo);
    const int n_before = sh.n_before, n_after = sh.n_after;
    u64 G0[kS], G1[kS];
#pragma unroll
    for (int i = 0; i < kS; ++i) G0[i] = G1[i] = 0ull;
    int sy = 0, dslot = 0, cw_py = -1;
    unsigned dphase = 0, cw = 0u;  // cw: this lane's 2-bit arg-max codes (2 channels x 7 bins) of bin-row cw_py
    float yl_next = sh.ylerp[0];   // the y weight of the next crop row, loaded one crop row ahead
    // One crop row: wait for its bin-row's dY, accumulate (unless it lies outside the map), release the slot after the
    // second row of the bin-row and re-arm it with the bin-row three further down.
    auto crop = [&](bool g0_top, bool outside) {
        const float yl = yl_next;
        yl_next = sh.ylerp[min(sy + 1, kS - 1)];
        mbar_wait(w.dfull0 + dslot * 8, dphase);
        if (!outside) {
            if (cw_py != (sy >> 1)) {  // first crop row of this bin-row that lies inside the map
                cw = 0u;
#pragma unroll
                for (int j = 0; j < kP; ++j) {
                    unsigned byte;
                    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(byte) : "r"(code_slots + dslot * kBwdCodeBytes + j * 16 + (lane >> 1)));
                    cw |= ((byte >> ((lane & 1) * 4)) & 15u) << (4 * j);
                }
                cw_py = sy >> 1;
            }
            unsigned dy_addr = w.lane_smem + (kBwdRowPix + 2) * kPixBytes + dslot * kBwdDyBytes;
            asm volatile("" : "+r"(dy_addr));
            bwd_crop_row(G0, G1, g0_top, (sy & 1) != 0, yl, dy_addr, cw);
        }
        if (sy & 1) {
            __syncwarp();
            const int py = sy >> 1;
            if (py + kBwdDySlots < kP) stage_bin_row(py + kBwdDySlots, dslot);
            if (++dslot == kBwdDySlots) {
                dslot = 0;
                dphase ^= 1u;
            }
        }
        ++sy;
    };
    for (int i = 0; i < n_before; ++i) crop(false, true);
    // Walk position k enters register set k & 1: the row that set held (position k - 2) is complete and is flushed; then
    // the crop rows between positions k - 1 (top, the other set) and k (bottom, this set) are accumulated.
#pragma unroll 1
    for (int k = 0; k < n_rows + 2; ++k) {
        const bool odd = (k & 1) != 0;
        if (k >= 2) {
            if (odd) bwd_flush(w, G1, wx, aoff, k - 2, b_pix0, Wf);
            else bwd_flush(w, G0, wx, aoff, k - 2, b_pix0, Wf);
        }
        if (k < n_rows) {
            prog >>= 1;
            while (prog & 1ull) {
                crop(odd, false);  // odd position: the top row sits in set 0
                prog >>= 1;
            }
        }
    }
    for (int i = 0; i < n_after; ++i) crop(false, true);
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all reduce-adds performed before exit
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
bool roi_pool_stream_supported(int C, int P, int k) { return C % 512 == 0 && C % kChunk == 0 && P == kP && k == 2; }

int launch_roi_pool_fwd_stream(const float* feat, int B, int Hf, int Wf, int C, const float* rois, int R, float* out,
                               unsigned char* codes, cudaStream_t st) {
    const size_t smem = (sizeof(StreamShared) + 127) / 128 * 128 + kRingBytes;
    static bool configured[64] = {};  // the attribute is per function and per device; setting it twice is harmless
    int dev = 0;
    FRCNN_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        FRCNN_CUDA(cudaFuncSetAttribute(roi_pool_fwd_stream_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        FRCNN_CUDA(cudaFuncSetAttribute(roi_pool_fwd_stream_kernel<1024, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        FRCNN_CUDA(cudaFuncSetAttribute(roi_pool_fwd_stream_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        FRCNN_CUDA(cudaFuncSetAttribute(roi_pool_fwd_stream_kernel<1024, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    const int nchunk = C / kChunk;
    FRCNN_REQUIRE((int64_t)R * nchunk < (1ll << 31), FRCNN_ELIMIT, "roi_pool_forward: R * C / 512 too large");
    FRCNN_REQUIRE((int64_t)B * Hf * Wf < (1ll << 31), FRCNN_ELIMIT, "roi_pool_forward: B * Hf * Wf too large");
    const TensorMaps* maps = nullptr;
    const int rc = get_tensor_maps(feat, (long long)B * Hf * Wf, C, dev, &maps);
    if (rc) return rc;
    const float4* r4 = (const float4*)rois;
    const dim3 grid(R * nchunk, B);
    if (codes) {
        if (C == 1024) roi_pool_fwd_stream_kernel<1024, true><<<grid, kThreads, smem, st>>>(*maps, feat, Hf, Wf, C, r4, R, nchunk, out, codes);
        else roi_pool_fwd_stream_kernel<0, true><<<grid, kThreads, smem, st>>>(*maps, feat, Hf, Wf, C, r4, R, nchunk, out, codes);
    } else {
        if (C == 1024) roi_pool_fwd_stream_kernel<1024, false><<<grid, kThreads, smem, st>>>(*maps, feat, Hf, Wf, C, r4, R, nchunk, out, nullptr);
        else roi_pool_fwd_stream_kernel<0, false><<<grid, kThreads, smem, st>>>(*maps, feat, Hf, Wf, C, r4, R, nchunk, out, nullptr);
    }
    FRCNN_LAUNCH_CHECK("roi_pool_fwd_stream_kernel");
    return 0;
}

static std::atomic<int> g_roi_pool_max_sms{0};
void roi_pool_set_max_sms(int max_sms) { g_roi_pool_max_sms.store(max_sms > 0 ? max_sms : 0, std::memory_order_relaxed); }

size_t roi_pool_stream_workspace_bytes(int B, int R) { return 256 + (size_t)kRecStride * (size_t)B * (size_t)R; }

// Persistent forward: records pre-kernel + warp-persistent pooling kernel.  workspace: [counter | records].
int launch_roi_pool_fwd_persist(const float* feat, int B, int Hf, int Wf, int C, const float* rois, int R, float* out,
                                unsigned char* codes, void* workspace, size_t workspace_bytes, cudaStream_t st) {
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
    roi_records_kernel<<<(n_rois + 3) / 4, 128, 0, st>>>((const float4*)rois, n_rois, R, Hf, Wf, recs, counter,
                                                        2u * (unsigned)(ctas * kPWarps));
    FRCNN_LAUNCH_CHECK("roi_records_kernel");
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
                                const unsigned char* codes, float* dfeat, void* workspace, size_t workspace_bytes, cudaStream_t st) {
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
    roi_records_kernel<<<(n_rois + 3) / 4, 128, 0, st>>>((const float4*)rois, n_rois, R, Hf, Wf, recs, counter,
                                                        2u * (unsigned)(ctas * kPWarps));
    FRCNN_LAUNCH_CHECK("roi_records_kernel");
    roi_pool_bwd_persist_kernel<<<ctas, kPThreads, smem, st>>>(*maps, Hf, Wf, C, recs, n_items, dy, codes, dfeat, counter);
    FRCNN_LAUNCH_CHECK("roi_pool_bwd_persist_kernel");
    return 0;
}

// dfeat must have been zeroed on the same stream.
int launch_roi_pool_bwd_stream(int B, int Hf, int Wf, int C, const float* rois, int R, const float* dy,
                               const unsigned char* codes, float* dfeat, cudaStream_t st) {
    const size_t smem = (sizeof(BwdShared) + 127) / 128 * 128 + (size_t)kWarps * kBwdWarpBytes;
    static bool configured[64] = {};
    int dev = 0;
    FRCNN_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        FRCNN_CUDA(cudaFuncSetAttribute(roi_pool_bwd_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    const int nchunk = C / kChunk;
    FRCNN_REQUIRE((int64_t)R * nchunk < (1ll << 31), FRCNN_ELIMIT, "roi_pool_backward: R * C / 512 too large");
    FRCNN_REQUIRE((int64_t)B * Hf * Wf < (1ll << 31) && (int64_t)B * R * kP * kP < (1ll << 31), FRCNN_ELIMIT,
                  "roi_pool_backward: tensor too large");
    const BwdMaps* maps = nullptr;
    const int rc = get_bwd_maps(dfeat, (long long)B * Hf * Wf, dy, codes, (long long)B * R * kP * kP, C, dev, &maps);
    if (rc) return rc;
    const dim3 grid(R * nchunk, B);
    roi_pool_bwd_stream_kernel<<<grid, kThreads, smem, st>>>(*maps, Hf, Wf, C, (const float4*)rois, R, nchunk, dy, codes, dfeat);
    FRCNN_LAUNCH_CHECK("roi_pool_bwd_stream_kernel");
    return 0;
}

}  // namespace frcnn

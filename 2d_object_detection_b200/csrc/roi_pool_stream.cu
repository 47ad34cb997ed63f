// Row-streaming RoI pooling forward for the reference configuration (pooled_size 7, kernel_size 2: a 14 x 14
// crop_and_resize max-pooled 2 x 2; models/detectors/fast_rcnn_detector.py:134-177).
//
// CTA = one RoI x 512 channels, 8 warps; a warp owns 64 channels (one float2 per lane; 128 registers, 2 CTAs per SM)
// and runs its OWN copy pipeline -- after the prologue the warps never synchronise with each other.
//   * Prologue (warp 0, registers + shuffles): the feature columns [cmin, cmax] under the 14 x-samples select one of
//     eight TMA tensor maps (box = 64 channels x W pixels, W >= the span); the 14 y-samples become the list of feature
//     rows the downward walk needs, in order, and per crop row the walk position that completes it.
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

#include "roi_axis.cuh"

namespace frcnn {

namespace {

constexpr int kS = 14;                         // crop size = pooled_size * kernel_size
constexpr int kP = 7;
constexpr int kChunk = 512;                    // channels per CTA
constexpr int kWarpCh = 64;                    // channels per warp: one float2 per lane
constexpr int kWarps = kChunk / kWarpCh;
constexpr int kThreads = kWarps * 32;
constexpr int kPixBytes = kWarpCh * 4;         // one pixel in a warp's ring
constexpr int kRingPix = 52;                   // pixels per warp ring: 8 warps x 52 x 256 B = 104 KB per CTA
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

// Fold one finished crop row (top / bottom feature rows T / Bm, weight yl) into the 7 running bin maxima.  kFirst: first
// crop row of the bin-row.  kCodes: also track which of the four window samples holds the maximum (training forward).
template <bool kFirst, bool kCodes>
__device__ __forceinline__ void crop_row(const u64 (&T)[kS], const u64 (&Bm)[kS], float yl, u64 one, float2 (&M)[kP],
                                         unsigned& cw) {
    const u64 ww = pack2(yl, yl);
#pragma unroll
    for (int j = 0; j < kP; ++j) {
        const float2 v0 = unpack2(lerp2(T[2 * j], Bm[2 * j], ww, one));
        const float2 v1 = unpack2(lerp2(T[2 * j + 1], Bm[2 * j + 1], ww, one));
        if (kCodes) {
            if (kFirst) {
                M[j] = v0;  // window position 0; cw was cleared for this bin-row
                take_if_greater(M[j].x, cw, v1.x, 1u, 4 * j);
                take_if_greater(M[j].y, cw, v1.y, 1u, 4 * j + 2);
            } else {
                take_if_greater(M[j].x, cw, v0.x, 2u, 4 * j);
                take_if_greater(M[j].y, cw, v0.y, 2u, 4 * j + 2);
                take_if_greater(M[j].x, cw, v1.x, 3u, 4 * j);
                take_if_greater(M[j].y, cw, v1.y, 3u, 4 * j + 2);
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
__global__ void __launch_bounds__(kThreads, 2)
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

    if (tid < 32) {
        const float4 roi = rois[(size_t)b * R + r];  // x1,y1,x2,y2 (fast_rcnn_detector.py:157-158 reorders for TF)
        const unsigned all = 0xffffffffu;
        const Sample xs = sample_axis(roi.x, roi.z, Wf, lane);  // lanes 0..13 are the crop columns / rows
        const Sample ys = sample_axis(roi.y, roi.w, Hf, lane);
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
        const bool wide = box >= kNumBoxes || W + first_px > kRingPix;  // does not fit the ring: plain gather path below
        if (wide && lane < kS) {
            sh.xlo[lane] = xs.lo; sh.xhi[lane] = xs.hi; sh.xok[lane] = xs.ok; sh.xlerp[lane] = xs.lerp;
            sh.ylo[lane] = ys.lo; sh.yhi[lane] = ys.hi; sh.yok[lane] = ys.ok;
        }
        if (lane < kS) {
            const int lerp = xs.ok ? __float_as_int(xs.lerp) : 0;
            sh.xt[lane] = make_int4(xs.ok ? (xs.lo - sp.cmin + first_px) * kPixBytes : 0, 0, lerp, lerp);
        }
        // ---- rows: the downward walk.  The walk stages feature rows in the order rows[0], rows[1], ...; position k
        // goes to register row set k & 1, so the most recently staged row is always the bottom neighbour and the set
        // it replaces held the old top row.  A crop row is complete when the row at pos[sy] has arrived.
        const bool anyx = sp.need > 0;
        const bool v = lane < kS && ys.ok && anyx;
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
        if (lane < kS) sh.ylerp[lane] = ys.lerp;
        if (lane == 0) {
            const int px = W + first_px;
            sh.n_rows = __popc(m1) + __popc(m2);
            sh.prog_lo = prog_lo;
            sh.prog_hi = prog_hi;
            sh.n_before = n_before;
            sh.n_after = kS - n_before - __popc(mv);
            sh.box = box;
            sh.cmin = sp.cmin;
            sh.first_px = first_px;
            sh.wide = wide ? 1 : 0;
            sh.slot_bytes = px * kPixBytes;
            sh.depth = max(1, min(kMaxDepth, kRingPix / max(px, 1)));
        }
    } else if (tid < 32 + kWarps * kMaxDepth) {
        mbar_init(full_base + (tid - 32) * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (sh.wide) {  // rare (x-samples more than 51 columns apart): one bilinear gather per sample, same arithmetic
        const float onef = kOne;
        const u64 one = pack2(onef, onef);
        const float* fb = feat + (size_t)b * Hf * Wf * C + ch0 + tid * 2;
        float2* ob = reinterpret_cast<float2*>(out + ((size_t)b * R + r) * (kP * kP) * C + ch0) + tid;
#pragma unroll 1
        unsigned char* cb = kCodes ? codes + ((size_t)b * R + r) * (kP * kP) * (C / 4) + (ch0 + tid * 2) / 4 : nullptr;
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
    unsigned cw = 0u;  // kCodes: the 2-bit window positions of this lane's two channels for the 7 bins of the bin-row
    unsigned char* cptr = kCodes ? codes + ((size_t)b * R + r) * (kP * kP) * (C / 4) + (ch0 + tid * 2) / 4 : nullptr;
    // One finished crop row: fold it into the bin maxima; after the second row of a bin-row store the 7 bins.
    auto emit = [&](const u64 (&T)[kS], const u64 (&Bm)[kS], bool outside) {
        if (!(sy & 1)) {
            cw = 0u;
            if (outside) {  // crop row outside the map: extrapolation value 0 takes part in the max
#pragma unroll
                for (int j = 0; j < kP; ++j) M[j] = make_float2(0.0f, 0.0f);
            } else {
                crop_row<true, kCodes>(T, Bm, sh.ylerp[sy], w.one, M, cw);
            }
        } else {
            if (outside) {
#pragma unroll
                for (int j = 0; j < kP; ++j) {
                    if (kCodes) {  // window positions 2 and 3 are both 0: position 2 wins if 0 beats the first row
                        take_if_greater(M[j].x, cw, 0.0f, 2u, 4 * j);
                        take_if_greater(M[j].y, cw, 0.0f, 2u, 4 * j + 2);
                    } else {
                        M[j] = make_float2(fmaxf(M[j].x, 0.0f), fmaxf(M[j].y, 0.0f));
                    }
                }
            } else {
                crop_row<false, kCodes>(T, Bm, sh.ylerp[sy], w.one, M, cw);
            }
#pragma unroll
            for (int j = 0; j < kP; ++j) __stcs(optr + j * bin_stride, M[j]);
            optr += kP * bin_stride;
            if (kCodes) {  // a byte holds four channels = this lane's nibble and its odd neighbour's
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

// Host side: the eight tensor maps of a feature tensor, cached by (pointer, shape).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct MapCacheEntry {
    const void* ptr = nullptr;
    long long pixels = 0;
    int C = 0, dev = -1;
    TensorMaps maps;
};

int get_tensor_maps(const float* feat, long long pixels, int C, int dev, const TensorMaps** out) {
    static thread_local MapCacheEntry cache[16];
    static thread_local int next = 0;
    static EncodeTiledFn encode = nullptr;
    for (auto& e : cache)
        if (e.ptr == feat && e.pixels == pixels && e.C == C && e.dev == dev) {
            *out = &e.maps;
            return 0;
        }
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        FRCNN_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        FRCNN_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, FRCNN_EINVAL,
                      "roi_pool_forward: cuTensorMapEncodeTiled not available in this driver");
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    MapCacheEntry& e = cache[next];
    next = (next + 1) % 16;
    e.ptr = nullptr;
    for (int i = 0; i < kNumBoxes; ++i) {
        const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)pixels};
        const cuuint64_t strides[1] = {(cuuint64_t)C * sizeof(float)};
        const cuuint32_t box[2] = {(cuuint32_t)kWarpCh, (cuuint32_t)box_width(i)};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult cr = encode(&e.maps.box[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(feat), dims, strides,
                                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        FRCNN_REQUIRE(cr == CUDA_SUCCESS, FRCNN_EINVAL, "roi_pool_forward: cuTensorMapEncodeTiled failed (CUresult %d)", (int)cr);
    }
    e.ptr = feat;
    e.pixels = pixels;
    e.C = C;
    e.dev = dev;
    *out = &e.maps;
    return 0;
}

}  // namespace

bool roi_pool_stream_supported(int C, int P, int k) { return C % kChunk == 0 && P == kP && k == 2; }

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
    const dim3 grid(R * nchunk, B);
    const float4* r4 = (const float4*)rois;
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

}  // namespace frcnn

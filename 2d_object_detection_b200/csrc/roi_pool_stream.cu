// Row-streaming RoI pooling forward for the reference configuration (pooled_size 7, kernel_size 2: a 14 x 14
// crop_and_resize max-pooled 2 x 2; models/detectors/fast_rcnn_detector.py:134-177).
//
// CTA = one RoI x 512 channels, 256 threads, one float2 of channels per thread (128 registers, 2 CTAs per SM).
//   * Prologue (warp 0, registers + shuffles): the 14 x-samples of the RoI touch at most 28 distinct feature columns;
//     they are compacted into a sorted list (match_any + rank).  The 14 y-samples become a short program: per crop
//     row which of the two register row sets must be (re)loaded and which one holds the top row, plus the list of
//     feature rows the downward walk needs, in order.
//   * Staging: a feature row is brought in with ONE cp.async.bulk per distinct column (TMA engine, 2 KB = 512 channels
//     of one pixel, issued lane-parallel by one warp) into slot `k % depth` of a shared-memory ring; a "full" mbarrier
//     per slot carries the byte count.  The ring depth adapts to the RoI: narrow RoIs (few columns, small slots) get a
//     deep ring.  There is no producer warp: the LAST warp to finish reading a slot (shared-memory arrival counter)
//     refills it at once with the row `depth` further down, so nobody ever waits for a slot to drain.
//   * A staged row is x-interpolated ONCE at all 14 sample positions into one of two register row sets (operands from
//     shared memory, addresses from a 14-entry table); the sets alternate as top / bottom neighbour while the walk
//     moves down; each crop row is then 14 y-interpolations folded straight into the 7 running bin maxima, stored
//     after every second crop row.  All control flow is CTA-uniform, every register index is static.
// Per sample the fp32 operations and their order are those of TF's CPU kernel (t = TL + (TR - TL) * xl, b likewise,
// out = t + (b - t) * yl, each op rounded separately; packed FADD2/FMUL2/FFMA2 with a run-time 1.0, see lerp2):
// results are bit-identical to the oracle.  Compared with the bin-row walk (roi_pool.cu) every feature pixel is read
// from L2 once per RoI instead of once per bin-row, a row is x-interpolated once per RoI, and nobody waits on L2.
#include "roi_axis.cuh"

namespace frcnn {

namespace {

constexpr int kS = 14;                         // crop size = pooled_size * kernel_size
constexpr int kP = 7;
constexpr int kChunk = 512;                    // channels per CTA
constexpr int kThreads = kChunk / 2;           // one float2 of channels per thread; 8 warps
constexpr int kWarps = kThreads / 32;
constexpr int kPixBytes = kChunk * 4;          // one pixel's channel chunk
constexpr int kRingPix = 52;                   // 104 KB ring: depth(28 columns) = 1, (25) = 2, (16) = 3, (12) = 4
constexpr int kRingBytes = kRingPix * kPixBytes;
constexpr int kMaxDepth = 8;

constexpr int kPosBefore = -1;  // crop row outside the map, emitted before the first staged row ...
constexpr int kPosAfter = 99;   // ... or after the last one

struct StreamShared {
    int4 xt[kS];          // per crop column: {left pixel byte offset in a slot, 0, x_lerp bits, x_lerp bits}
    int pos[kS];          // per crop row: walk position of its bottom feature row (the row whose arrival completes it)
    float ylerp[kS];
    int rows[2 * kS];     // feature rows in the order the walk consumes them
    int cols[32];         // sorted distinct feature columns (compact index -> column)
    int n_rows, ncols, depth, slot_bytes, first_px, spare_px;
    int done[kMaxDepth];  // warps that have finished reading the slot's current row
    unsigned long long full[kMaxDepth];
};

__device__ __constant__ float kOne = 1.0f;
#ifdef FRCNN_DEBUG_SPIN
__device__ int g_dbg[8 + 16 * 8];
#define g_dbg_k 0
__device__ volatile int* g_trace = nullptr;
#define TRACE(v) do { if (g_trace && blockIdx.x == 0 && blockIdx.y == 0 && (threadIdx.x & 31) == 0) { g_trace[threadIdx.x >> 5] = (v); __threadfence_system(); } } while (0)
#else
#define TRACE(v) do {} while (0)
#endif

__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
#ifdef FRCNN_DEBUG_SPIN
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {  // debug build: a wait that does not complete within 20 ms is recorded and the thread exits
        unsigned ok;
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return;
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 20000000ull) {
            const int n = atomicAdd(&g_dbg[0], 1);
            if (n < 15) {
                int* d = g_dbg + 8 + n * 8;
                d[0] = blockIdx.x; d[1] = blockIdx.y; d[2] = threadIdx.x; d[3] = (int)bar; d[4] = (int)parity;
                unsigned long long st;
                asm volatile("ld.shared.b64 %0, [%1];" : "=l"(st) : "r"(bar));
                d[5] = (int)(st >> 32); d[6] = (int)st; d[7] = g_dbg_k;
            }
            asm volatile("exit;");
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
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
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

struct Walk {
    StreamShared* sh;
    const unsigned char* ring;  // this thread's channels inside pixel 0 of slot 0
    const float* src;           // staging duty: this lane's feature column (row 0) of the image, channel chunk ch0
    size_t row_stride;          // floats per feature row
    unsigned dst;               // staging duty: shared address of this lane's pixel in slot 0
    unsigned full0;             // shared address of full[0]
    int k, n_rows, ncols, depth, slot_bytes, slot;
    unsigned phase;
    u64 one;
};

// Stage feature row rows[idx] into `slot` (one 2 KB bulk copy per lane = per distinct column); warp-collective.
__device__ __forceinline__ void stage_row(const Walk& w, int idx, int slot) {
    const int lane = threadIdx.x & 31;
    const unsigned bar = w.full0 + slot * 8;
    if (lane == 0) mbar_expect_tx(bar, (unsigned)(w.ncols * kPixBytes));
    __syncwarp();
    if (lane < w.ncols) bulk_g2s(w.dst + slot * w.slot_bytes, w.src + w.sh->rows[idx] * w.row_stride, kPixBytes, bar);
}

// x-interpolate the next staged feature row (row k of the walk) at the 14 sample positions.
__device__ __forceinline__ void load_row(Walk& w, u64 (&X)[kS]) {
    TRACE(1000 + w.k);
    mbar_wait(w.full0 + w.slot * 8, w.phase);
    TRACE(2000 + w.k);
    const unsigned char* base = w.ring + w.slot * w.slot_bytes;
#pragma unroll
    for (int i = 0; i < kS; ++i) {
        const int4 e = w.sh->xt[i];
        const u64 l = *reinterpret_cast<const u64*>(base + e.x);
        const u64 r = *reinterpret_cast<const u64*>(base + e.x + kPixBytes);  // the right neighbour is the next slot pixel
        X[i] = lerp2(l, r, pack2(__int_as_float(e.z), __int_as_float(e.w)), w.one);
    }
    // the last warp to finish with this slot refills it with the row `depth` further down the walk
    int old = 0;
    if ((threadIdx.x & 31) == 0) old = atomicAdd(&w.sh->done[w.slot], 1);
    old = __shfl_sync(0xffffffffu, old, 0);
    TRACE(3000 + w.k + 100 * old);
    if (old == kWarps - 1) {
        if ((threadIdx.x & 31) == 0) w.sh->done[w.slot] = 0;
        if (w.k + w.depth < w.n_rows) stage_row(w, w.k + w.depth, w.slot);
    }
    TRACE(4000 + w.k);
    ++w.k;
    if (++w.slot == w.depth) {
        w.slot = 0;
        w.phase ^= 1u;
    }
}

template <bool kFirst>
__device__ __forceinline__ void crop_row(const u64 (&T)[kS], const u64 (&Bm)[kS], float yl, u64 one, float2 (&M)[kP]) {
    const u64 ww = pack2(yl, yl);
#pragma unroll
    for (int j = 0; j < kP; ++j) {
        const float2 v0 = unpack2(lerp2(T[2 * j], Bm[2 * j], ww, one));
        const float2 v1 = unpack2(lerp2(T[2 * j + 1], Bm[2 * j + 1], ww, one));
        if (kFirst) {
            M[j].x = fmaxf(v0.x, v1.x);
            M[j].y = fmaxf(v0.y, v1.y);
        } else {
            M[j].x = fmaxf(fmaxf(M[j].x, v0.x), v1.x);
            M[j].y = fmaxf(fmaxf(M[j].y, v0.y), v1.y);
        }
    }
}

__global__ void __launch_bounds__(kThreads, 2)
    roi_pool_fwd_stream_kernel(const float* __restrict__ feat, int Hf, int Wf, int C, const float4* __restrict__ rois, int R,
                               int nchunk, float* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char smem[];
    StreamShared& sh = *reinterpret_cast<StreamShared*>(smem);
    unsigned char* ring = smem + (sizeof(StreamShared) + 127) / 128 * 128;
    const int tid = threadIdx.x, lane = tid & 31;
    const int r = blockIdx.x / nchunk, ch0 = (blockIdx.x % nchunk) * kChunk, b = blockIdx.y;
    const unsigned full0 = smem_u32(&sh.full[0]);
    TRACE(7);

    if (tid < 32) {
        const float4 roi = rois[(size_t)b * R + r];  // x1,y1,x2,y2 (fast_rcnn_detector.py:157-158 reorders for TF)
        const unsigned all = 0xffffffffu;
        const Sample xs = sample_axis(roi.x, roi.z, Wf, lane);  // lanes 0..13 are the crop columns / rows
        const Sample ys = sample_axis(roi.y, roi.w, Hf, lane);
        // ---- columns: the distinct feature columns under the 14 x-samples -> sorted compact list; a slot holds them
        // in that order.  The right neighbour of a sample (ceil) is the next list entry, or -- at an integer position,
        // where its weight is exactly 0 -- whatever follows: that must be finite, so a zeroed spare pixel closes the
        // slot when the last column can be hit that way.  A sample outside the map reads a zeroed pixel in front of
        // the slot: 0 + (p - 0) * 0.
        const int i = lane >> 1;
        const int ok_i = __shfl_sync(all, (int)xs.ok, i);  // (collectives first: no short-circuit around a shuffle)
        const bool ok = lane < 2 * kS && ok_i != 0;
        const int lo_i = __shfl_sync(all, xs.lo, i), hi_i = __shfl_sync(all, xs.hi, i);
        const int lerp_i = __shfl_sync(all, __float_as_int(xs.lerp), i);
        const int val = ok ? ((lane & 1) ? hi_i : lo_i) : 0x7fffffff;
        const unsigned same = __match_any_sync(all, val);
        const bool leader = ok && (__ffs(same) - 1 == lane);
        int rank = 0;
#pragma unroll
        for (int j = 0; j < 2 * kS; ++j) {
            const int vj = __shfl_sync(all, val, j);
            const int lj = __shfl_sync(all, (int)leader, j);
            rank += (lj && vj < val) ? 1 : 0;
        }
        TRACE(8);
        const int ncols = __popc(__ballot_sync(all, leader));
        const bool even = lane < 2 * kS && !(lane & 1);
        const int first_px = __any_sync(all, even && !ok) ? 1 : 0;
        const int spare_px = __any_sync(all, even && ok && hi_i == lo_i && rank == ncols - 1) ? 1 : 0;
        if (leader) sh.cols[rank] = val;
        if (even) sh.xt[i] = make_int4(ok ? (rank + first_px) * kPixBytes : 0, 0, ok ? lerp_i : 0, ok ? lerp_i : 0);
        // ---- rows: the downward walk.  The walk stages feature rows in the order rows[0], rows[1], ...; position k
        // goes to register row set k & 1, so the most recently staged row is always the bottom neighbour and the set
        // it replaces held the old top row.  A crop row is complete when the row at pos[sy] has arrived.
        const bool anyx = ncols > 0;
        const bool v = lane < kS && ys.ok && anyx;
        const int c = ys.lo;
        const int pv_i = __shfl_up_sync(all, (int)v, 1);
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
        if (lane < kS) {
            sh.pos[lane] = v ? off + (load2 ? 2 : shift ? 1 : 0) - 1 : ((mv & lt) ? kPosAfter : kPosBefore);
            sh.ylerp[lane] = ys.lerp;
        }
        if (lane == 0) {
            const int px = ncols + first_px + spare_px;
            sh.n_rows = __popc(m1) + __popc(m2);
            sh.ncols = ncols;
            sh.first_px = first_px;
            sh.spare_px = spare_px;
            sh.slot_bytes = px * kPixBytes;
            sh.depth = px ? min(kMaxDepth, kRingPix / px) : 1;
        }
        TRACE(9);
    } else if (tid < 32 + kMaxDepth) {
        mbar_init(full0 + (tid - 32) * 8, 1);
        sh.done[tid - 32] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

#ifdef FRCNN_DEBUG_SPIN
    if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) {
        g_dbg[1] = sh.n_rows; g_dbg[2] = sh.ncols; g_dbg[3] = sh.depth; g_dbg[4] = sh.slot_bytes; g_dbg[5] = sh.first_px;
        g_dbg[6] = sh.spare_px; g_dbg[7] = (int)full0;
    }
#endif
    TRACE(1);
    Walk w;
    w.sh = &sh;
    w.full0 = full0;
    w.ncols = sh.ncols;
    w.n_rows = sh.n_rows;
    w.depth = sh.depth;
    w.slot_bytes = sh.slot_bytes;
    w.row_stride = (size_t)Wf * C;
    w.src = feat + ((size_t)b * Hf * Wf + (lane < w.ncols ? sh.cols[lane] : 0)) * C + ch0;
    w.dst = smem_u32(ring) + (lane + sh.first_px) * kPixBytes;
    w.ring = ring + tid * 8;
    w.k = 0;
    w.slot = 0;
    w.phase = 0;
    {
        const float onef = kOne;
        w.one = pack2(onef, onef);
    }
    if (w.n_rows > 0) {
        if (tid < 32)
            for (int s = 0; s < w.depth && s < w.n_rows; ++s) stage_row(w, s, s);
        TRACE(2);
        // zero pixels of every slot (each thread later reads only what it wrote itself)
        for (int s = 0; s < w.depth; ++s) {
            unsigned char* slot0 = ring + s * w.slot_bytes + tid * 8;
            if (sh.first_px) *reinterpret_cast<u64*>(slot0) = 0ull;
            if (sh.spare_px) *reinterpret_cast<u64*>(slot0 + (sh.first_px + w.ncols) * kPixBytes) = 0ull;
        }
    }

    TRACE(3);
    u64 X0[kS], X1[kS];
#pragma unroll
    for (int i = 0; i < kS; ++i) X0[i] = X1[i] = 0ull;
    float2 M[kP];
    float2* optr = reinterpret_cast<float2*>(out + ((size_t)b * R + r) * (kP * kP) * C + ch0) + tid;
    const int bin_stride = C / 2;  // float2 per bin
    int sy = 0;
    int next_pos = sh.pos[0];
    // One finished crop row: fold it into the bin maxima; after the second row of a bin-row store the 7 bins.
    auto emit = [&](const u64 (&T)[kS], const u64 (&Bm)[kS], bool outside) {
        const float yl = sh.ylerp[sy];
        if (!(sy & 1)) {
            if (outside) {  // crop row outside the map: extrapolation value 0 takes part in the max
#pragma unroll
                for (int j = 0; j < kP; ++j) M[j] = make_float2(0.0f, 0.0f);
            } else {
                crop_row<true>(T, Bm, yl, w.one, M);
            }
        } else {
            if (outside) {
#pragma unroll
                for (int j = 0; j < kP; ++j) M[j] = make_float2(fmaxf(M[j].x, 0.0f), fmaxf(M[j].y, 0.0f));
            } else {
                crop_row<false>(T, Bm, yl, w.one, M);
            }
#pragma unroll
            for (int j = 0; j < kP; ++j) __stcs(optr + j * bin_stride, M[j]);
            optr += kP * bin_stride;
        }
        ++sy;
        next_pos = sy < kS ? sh.pos[sy] : 1000;
    };
    while (next_pos == kPosBefore) emit(X0, X1, true);
#pragma unroll 1
    for (int k = 0; k < w.n_rows; k += 2) {
        load_row(w, X0);  // even position: the new bottom row lands in set 0, set 1 holds the top row
        while (next_pos == k) emit(X1, X0, false);
        if (k + 1 < w.n_rows) {
            load_row(w, X1);
            while (next_pos == k + 1) emit(X0, X1, false);
        }
    }
    while (sy < kS) emit(X0, X1, true);
    TRACE(9999);
}

}  // namespace

#ifdef FRCNN_DEBUG_SPIN
extern "C" __attribute__((visibility("default"))) int frcnn_debug_set_trace(int* mapped) {
    return (int)cudaMemcpyToSymbol(g_trace, &mapped, sizeof(mapped));
}
extern "C" __attribute__((visibility("default"))) int frcnn_debug_read(int* host_dst) {
    cudaDeviceSynchronize();
    int z[8 + 16 * 8] = {};
    cudaError_t e = cudaMemcpyFromSymbol(host_dst, g_dbg, sizeof(z));
    cudaMemcpyToSymbol(g_dbg, z, sizeof(z));
    return (int)e;
}
#endif

bool roi_pool_stream_supported(int C, int P, int k) { return C % kChunk == 0 && P == kP && k == 2; }

int launch_roi_pool_fwd_stream(const float* feat, int B, int Hf, int Wf, int C, const float* rois, int R, float* out,
                               cudaStream_t st) {
    const size_t smem = (sizeof(StreamShared) + 127) / 128 * 128 + kRingBytes;
    static bool configured[64] = {};  // the attribute is per function and per device; setting it twice is harmless
    int dev = 0;
    FRCNN_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        FRCNN_CUDA(cudaFuncSetAttribute(roi_pool_fwd_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    const int nchunk = C / kChunk;
    FRCNN_REQUIRE((int64_t)R * nchunk < (1ll << 31), FRCNN_ELIMIT, "roi_pool_forward: R * C / 512 too large");
    const dim3 grid(R * nchunk, B);
    roi_pool_fwd_stream_kernel<<<grid, kThreads, smem, st>>>(feat, Hf, Wf, C, (const float4*)rois, R, nchunk, out);
    FRCNN_LAUNCH_CHECK("roi_pool_fwd_stream_kernel");
    return 0;
}

}  // namespace frcnn

// Persistent row-streaming RoI pooling forward (included by roi_pool_stream.cu, inside its anonymous namespace: it reuses
// the sampling helpers, the packed lerp and the crop-row fold of the kernel above -- the arithmetic, and therefore every
// output bit, is unchanged).
//
// What changes is everything around the arithmetic.  The per-CTA kernel pays, for every (RoI, 256 channels) CTA, a
// two-warp prologue all four warps wait for, and one exposed L2 -> shared-memory latency before the first feature row can
// be interpolated; its per-row staging code re-derives the copy parameters from shared memory (~45 instructions).  Here:
//   * roi_records_kernel (one warp per RoI, a few microseconds per launch) turns every RoI into a 512-byte RECORD once:
//     sample offsets / weights, the walk's row list and bit string, box class, ring geometry (fast_rcnn_detector.py:157-168
//     reads the same boxes once per crop_and_resize call).
//   * roi_pool_fwd_persist_kernel runs 16 independent WARPS per CTA, one CTA per SM, and every warp is its own worker: it
//     pulls (RoI, 64-channel) items from a global counter -- the first two are static, then one atomicAdd per item, issued
//     two items ahead -- copies the item's record into shared memory with one cp.async.bulk (triple-buffered), and never
//     synchronises with another warp.
//   * The warp's 52-pixel ring is two halves.  An item whose slot is at most 13 pixels ("small": 3/4 of the KITTI
//     proposals) lives in one half; while the warp works on it, the first rows of the NEXT item are already on their way
//     into the other half, so the L2 latency of an item's first rows hides behind the previous item.  Wider items take
//     the whole ring as before and start cold.
//   * Staging a row = one elected lane: a 16-bit row number from the record, one multiply-add for the TMA coordinate,
//     expect_tx + cp.async.bulk.tensor.
#pragma once

constexpr int kRecBytes = 256;            // the part of a record the walk needs (bulk-copied to shared memory)
constexpr int kRecStride = 512;           // records are 512 B apart in global memory (second half: gather tables)
constexpr int kHalfPix = kRingPix / 2;    // 26
constexpr int kSmallPix = kHalfPix / 2;   // 13: a half-ring holds at least two slots
constexpr int kHalfDepth = 4;             // slots (and mbarriers) per half
constexpr int kRowBars = 2 * kHalfDepth;  // 8 row barriers per warp
constexpr int kRecBufs = 3;
constexpr int kPBarBytes = 128;           // 8 row + 3 record barriers, padded
constexpr int kPWarpBytes = kWarpRingBytes + kRecBufs * kRecBytes + kPBarBytes;  // 14208 B per warp, 56832 B per CTA
static_assert(kPWarpBytes % 128 == 0, "per-warp shared-memory region must keep the rings 128-byte aligned");
// ONE 16-warp CTA per SM (its warps never synchronise, so the CTA size is free): a grid of n CTAs occupies exactly n SMs and
// leaves the others completely empty -- the caller can keep a few SMs for kernels of other streams (the proposal path of
// the next batch needs whole SMs: 1024-thread NMS CTAs), see frcnn_roi_pool_set_max_sms.
constexpr int kPWarps = 16;
constexpr int kPThreads = kPWarps * 32;
static_assert(kPWarps * kPWarpBytes <= 227 * 1024, "one CTA per SM must fit");

enum : unsigned char { kClsSmall = 0, kClsLarge = 1, kClsGather = 2 };

struct __align__(16) RoiRecord {
    float wx[kS];                           //   0 x_lerp of sample i (0 outside the map); 16-byte aligned: read with 128-bit loads
    int aoff[kS];                           //  56 byte offset of sample i's left pixel inside a slot (0: the zero pixel)
    float ylerp[kS];                        // 112 y_lerp of crop row i
    int pix0;                               // 168 b * Hf * Wf + cmin: TMA pixel coordinate of column cmin in feature row 0
    unsigned prog_lo, prog_hi;              // 172 the walk as a bit string (see StreamShared)
    unsigned short slot_bytes, tx_bytes;    // 180 (W + first_px) * 256, W * 256
    unsigned char n_rows, box, first_px, cls, depth, n_before, n_after, pad0;  // 184
    unsigned short rows[2 * kS];            // 192 feature rows in the order the walk consumes them
    unsigned adv_mask;                      // 248 2 bits per crop column: how far the next column inside the map lies (0, 1, >= 2
                                            //     pixels; the last one inside: 2; a column outside: 0) -- backward x-scatter
    unsigned char mono, pad1[3];            // 252 1: the columns inside the map walk left to right (not a reversed box)
    // ---- second half: read from global memory by the gather path only (cls == kClsGather)
    unsigned short xlo[kS], xhi[kS], ylo[kS], yhi[kS];  // 256
    unsigned xok, yok;                      // 368 bit i: sample i lies inside the map
    int b, pad2;                            // 376 image index
    unsigned char pad3[128];                // 384
};
static_assert(sizeof(RoiRecord) == kRecStride, "record layout");
static_assert(offsetof(RoiRecord, wx) == 0 && offsetof(RoiRecord, aoff) == 56 && offsetof(RoiRecord, ylerp) == 112 &&
                  offsetof(RoiRecord, pix0) == 168 && offsetof(RoiRecord, prog_lo) == 172 && offsetof(RoiRecord, slot_bytes) == 180 &&
                  offsetof(RoiRecord, n_rows) == 184 && offsetof(RoiRecord, depth) == 188 && offsetof(RoiRecord, rows) == 192,
              "the kernel addresses record fields by offset");

// One warp per RoI: the prologue of roi_pool_fwd_stream_kernel, once per RoI instead of once per (RoI, 256 channels) CTA.
__global__ void __launch_bounds__(128) roi_records_kernel(const float4* __restrict__ rois, int n_rois, int R, int Hf, int Wf,
                                                          RoiRecord* __restrict__ recs, unsigned* __restrict__ counter) {
    const int lane = threadIdx.x & 31, gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const unsigned all = 0xffffffffu;
    // the work counters of the pooling kernels that follow on the stream (word 0: forward, word 1: backward); tickets count
    // from 0, a kernel adds the number of items it hands out statically (first_ticket: depends on its own grid)
    if (gw == 0 && lane < 2) counter[lane] = 0u;
    if (gw >= n_rois) return;
    const int b = gw / R;
    const float4 roi = rois[gw];  // x1,y1,x2,y2 (fast_rcnn_detector.py:157-158 reorders for TF)
    RoiRecord& rec = recs[gw];
    // ---- columns
    const Sample xs = sample_axis(roi.x, roi.z, Wf, lane);  // lanes 0..13 are the crop columns
    const Span sp = column_span(xs, lane);
    const int first_px = sp.any_outside ? 1 : 0;
    int box = 0;
#pragma unroll
    for (int i = 0; i < kNumBoxes; ++i) box += (sp.need > box_width(i)) ? 1 : 0;
    int W = 0;
#pragma unroll
    for (int i = 0; i < kNumBoxes; ++i) W = (box == i) ? box_width(i) : W;
    const int px = W + first_px;
    const bool gather = box >= kNumBoxes || px > kRingPix || sp.need == 0;
    // ---- rows: the downward walk (see roi_pool_fwd_stream_kernel)
    const Sample ys = sample_axis(roi.y, roi.w, Hf, lane);
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
        rec.rows[off] = (unsigned short)c;
        rec.rows[off + 1] = (unsigned short)min(c + 1, last);
    } else if (shift) {
        rec.rows[off] = (unsigned short)min(c + 1, last);
    }
    const int bitpos = off + (load2 ? 2 : shift ? 1 : 0) + __popc(mv & lt);
    const unsigned prog_lo = __reduce_or_sync(all, (v && bitpos < 32) ? (1u << bitpos) : 0u);
    const unsigned prog_hi = __reduce_or_sync(all, (v && bitpos >= 32) ? (1u << (bitpos - 32)) : 0u);
    const int n_before = __popc(__ballot_sync(all, lane < kS && !v && !(mv & lt)));
    const unsigned xok = __ballot_sync(all, lane < kS && xs.ok), yok = mv;
    // columns as the backward's x-scatter walks them: distance to the next column inside the map
    const int lo_n = __shfl_down_sync(all, xs.lo, 1);
    const bool ok_c = lane < kS && xs.ok, ok_n = lane + 1 < kS && ((xok >> (lane + 1)) & 1u);
    const int adv = !ok_c ? 0 : !ok_n ? 2 : min(max(lo_n - xs.lo, 0), 2);
    const unsigned adv_mask = __reduce_or_sync(all, lane < kS ? ((unsigned)adv << (2 * lane)) : 0u);
    const bool mono = __all_sync(all, !(ok_c && ok_n) || lo_n >= xs.lo);
    if (lane < kS) {
        rec.aoff[lane] = xs.ok ? (xs.lo - sp.cmin + first_px) * kPixBytes : 0;
        rec.wx[lane] = xs.ok ? xs.lerp : 0.0f;
        rec.ylerp[lane] = ys.lerp;
        rec.xlo[lane] = (unsigned short)xs.lo;
        rec.xhi[lane] = (unsigned short)xs.hi;
        rec.ylo[lane] = (unsigned short)ys.lo;
        rec.yhi[lane] = (unsigned short)ys.hi;
    }
    if (lane == 0) {
        const bool small = px <= kSmallPix;
        rec.pix0 = b * Hf * Wf + (gather ? 0 : sp.cmin);
        rec.prog_lo = prog_lo;
        rec.prog_hi = prog_hi;
        rec.slot_bytes = (unsigned short)(px * kPixBytes);
        rec.tx_bytes = (unsigned short)(W * kPixBytes);
        rec.n_rows = (unsigned char)(__popc(m1) + __popc(m2));
        rec.box = (unsigned char)(gather ? 0 : box);
        rec.first_px = (unsigned char)first_px;
        rec.cls = gather ? kClsGather : small ? kClsSmall : kClsLarge;
        rec.depth = (unsigned char)(gather ? 1 : small ? min(kHalfDepth, kHalfPix / px) : max(1, min(kHalfDepth, kRingPix / px)));
        rec.n_before = (unsigned char)n_before;
        rec.n_after = (unsigned char)(kS - n_before - __popc(mv));
        rec.pad0 = 0;
        rec.adv_mask = adv_mask;
        rec.mono = mono ? 1 : 0;
        rec.xok = xok;
        rec.yok = yok;
        rec.b = b;
        rec.pad2 = 0;
    }
}

__device__ __forceinline__ unsigned lds_u32(unsigned addr) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds_f32(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned lds_u16(unsigned addr) {
    unsigned v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned uniform(unsigned v) { return __reduce_or_sync(0xffffffffu, v); }  // warp-uniform -> uniform register
// Next work item: one atom per item, issued by lane 0 two items ahead of its use.  ptxas turns every integer atom.add /
// atom.inc (atomicAdd() and inline PTX alike) into its warp-aggregated form (VOTE / POPC / ATOMG / SHFL), which waits
// for the round trip to L2 at the issue point: 3-4 % of the stall samples.  A float counter (atom.add.f32 is left
// alone) removes that wait, but the time moves into the row-barrier waits: 0.1687 vs 0.1695 ms forward, 0.3434 vs
// 0.3421 ms backward in an alternating A/B on one box -- not worth a 2^24-item limit, so the counter stays an integer.
// The first item of a warp is static (its global warp index).  The second one is static too (index + number of warps) when
// the launch is long -- no atom round trip before the first item -- and comes from the counter when every warp only gets a
// few items (a batch of one image: 4800 items for 2240 warps): there the static pair fixes the makespan at the worst pair
// of items, while handing out everything behind the first item lets the warps that drew short items take more of them.
__host__ __device__ __forceinline__ bool second_item_is_static(int n_items, int total_warps) { return n_items >= 8 * total_warps; }
__host__ __device__ __forceinline__ unsigned first_ticket(int n_items, int total_warps) {
    return (second_item_is_static(n_items, total_warps) ? 2u : 1u) * (unsigned)total_warps;
}
__device__ __forceinline__ unsigned take_ticket(unsigned* counter) {
    unsigned t;
    asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(t) : "l"(counter) : "memory");
    return t;
}

// 256-byte record, global -> shared, completion on `bar` (issued by one lane)
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}

// The scalars of an item's record the walk keeps in (uniform) registers.
struct PItem {
    unsigned rec;       // shared address of the record
    int n_rows, depth, slot_bytes, tx_bytes, first_off, pix0, box, cls;
    unsigned area;      // byte offset of the item's ring area inside the warp's ring (0 or half a ring)
    unsigned bar;       // index of the first row barrier of the item's set (0 or kHalfDepth)
};
__device__ __forceinline__ PItem load_item(unsigned rec, int half) {
    PItem it;
    it.rec = rec;
    const unsigned s = uniform(lds_u32(rec + 180)), f = uniform(lds_u32(rec + 184)), g = uniform(lds_u32(rec + 188));
    it.slot_bytes = (int)(s & 0xffffu);
    it.tx_bytes = (int)(s >> 16);
    it.n_rows = (int)(f & 0xffu);
    it.box = (int)((f >> 8) & 0xffu);
    it.first_off = (int)((f >> 16) & 0xffu) * kPixBytes;
    it.cls = (int)(f >> 24);
    it.depth = (int)(g & 0xffu);
    it.pix0 = (int)uniform(lds_u32(rec + 168));
    const unsigned h = uniform((unsigned)half);  // (the caller's loop-carried flag, made provably warp-uniform)
    it.area = (it.cls == kClsSmall && h) ? (unsigned)(kHalfPix * kPixBytes) : 0u;
    it.bar = (it.cls == kClsSmall && h) ? (unsigned)kHalfDepth : 0u;
    return it;
}

// kC = number of channels when known at compile time (1024), 0 = any; kCodes: also write the arg-max codes.
template <int kC, bool kCodes>
__global__ void __launch_bounds__(kPThreads, 1)
    roi_pool_fwd_persist_kernel(const __grid_constant__ TensorMaps maps, const float* __restrict__ feat, int Hf, int Wf, int C_,
                                const RoiRecord* __restrict__ recs, int n_items, float* __restrict__ out,
                                unsigned char* __restrict__ codes, unsigned* __restrict__ counter) {
    const int C = kC ? kC : C_;
    const int nch = C / kWarpCh;  // 64-channel items per RoI
    extern __shared__ __align__(128) unsigned char smem[];
    const unsigned all = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = (int)uniform(threadIdx.x >> 5);  // (provably warp-uniform: control flow below)
    const unsigned ring_s = smem_u32(smem) + warp * kPWarpBytes;  // this warp's ring | 3 records | barriers
    const unsigned rec_s = ring_s + kWarpRingBytes;
    const unsigned bar_s = rec_s + kRecBufs * kRecBytes;          // row barriers 0..7, record barriers 8..10
    if (lane < kRowBars + kRecBufs) {
        mbar_init(bar_s + lane * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int total_warps = gridDim.x * kPWarps;
    int idx_cur = blockIdx.x * kPWarps + warp;  // the first item (or two: second_item_is_static) is static
    int idx_nxt = idx_cur + total_warps;
    if (idx_cur >= n_items) return;
    const float onef = kOne;
    const u64 one = pack2(onef, onef);

    auto fetch_record = [&](int item, int buf) {  // (uniform arguments)
        if (lane == 0) {
            const unsigned bar = bar_s + (kRowBars + buf) * 8;
            mbar_expect_tx(bar, kRecBytes);
            bulk_g2s(rec_s + buf * kRecBytes, reinterpret_cast<const unsigned char*>(recs) + (size_t)(item / nch) * kRecStride, kRecBytes,
                     bar);
        }
    };
    // Stage walk row k of an item into slot `slot` of its ring area: one TMA box, issued by one lane.
    auto stage_row = [&](const PItem& it, int ch0, int row, int slot) {  // feature row `row` of the item's image
        if (lane == 0) {
            const unsigned bar = bar_s + (it.bar + slot) * 8;
            mbar_expect_tx(bar, (unsigned)it.tx_bytes);
            tma_box_g2s(ring_s + it.area + slot * it.slot_bytes + it.first_off, &maps.box[it.box], ch0, it.pix0 + row * Wf, bar);
        }
    };
    auto stage = [&](const PItem& it, int ch0, int k, int slot) { stage_row(it, ch0, (int)lds_u16(it.rec + 192 + 2 * k), slot); };
    auto stage_first_rows = [&](const PItem& it, int ch0) {
        if (it.cls == kClsGather) return;
        const int n = min(it.depth, it.n_rows);
#pragma unroll 1
        for (int s = 0; s < n; ++s) stage(it, ch0, s, s);
    };

    unsigned row_phase = 0u, rec_phase = 0u;  // parity of the next completion of every barrier
    int buf = 0;                               // record buffer of the current item; the next item's is (buf + 1) % 3
    int half = 0;
    bool prefetched = false;
    fetch_record(idx_cur, 0);
    unsigned ticket = 0u;                      // lane 0: index of the item after next (an atomicAdd in flight)
    const unsigned ticket0 = first_ticket(n_items, total_warps);  // tickets count from 0: item = ticket + ticket0
    if (!second_item_is_static(n_items, total_warps)) {  // (the round trip overlaps the fetch of the first record)
        if (lane == 0) ticket = take_ticket(counter) + ticket0;
        idx_nxt = (int)uniform(lane == 0 ? ticket : 0u);
    }
    if (idx_nxt < n_items) fetch_record(idx_nxt, 1);
    if (lane == 0) ticket = take_ticket(counter) + ticket0;
    mbar_wait(bar_s + kRowBars * 8, 0u);       // the first record
    rec_phase = 1u;

#pragma unroll 1
    for (;;) {
        const PItem it = load_item(rec_s + buf * kRecBytes, half);
        const int roi = idx_cur / nch, ch0 = (idx_cur - roi * nch) * kWarpCh;
        if (!prefetched) stage_first_rows(it, ch0);  // cold start: nothing of this item is in flight yet
        // ---- look ahead: the item after next gets its record fetched, the next item (record already here) its first rows
        const int idx_n2 = (int)uniform(lane == 0 ? ticket : 0u);  // lane 0's ticket, broadcast into a uniform register
        const int nbuf = buf == kRecBufs - 1 ? 0 : buf + 1, n2buf = nbuf == kRecBufs - 1 ? 0 : nbuf + 1;
        bool next_prefetched = false;
        if (idx_nxt < n_items) {
            if (idx_n2 < n_items) {
                fetch_record(idx_n2, n2buf);
                if (lane == 0) ticket = take_ticket(counter) + ticket0;
            }
            mbar_wait(bar_s + (kRowBars + nbuf) * 8, (rec_phase >> nbuf) & 1u);  // fetched one item ago: normally complete
            rec_phase ^= 1u << nbuf;
            if (it.cls == kClsSmall) {
                const PItem nx = load_item(rec_s + nbuf * kRecBytes, half ^ 1);
                if (nx.cls == kClsSmall) {
                    const int nroi = idx_nxt / nch;
                    stage_first_rows(nx, (idx_nxt - nroi * nch) * kWarpCh);
                    next_prefetched = true;
                }
            }
        }

        float2* optr = reinterpret_cast<float2*>(out + (size_t)roi * (kP * kP) * C + ch0) + lane;
        const int bin_stride = C / 2;  // float2 per bin
        unsigned char* cptr = kCodes ? codes + (size_t)roi * (kP * kP) * (C / 4) + (ch0 + lane * 2) / 4 : nullptr;

        if (it.cls == kClsGather) {
            // rare (x-samples more than 51 columns apart, or no column inside the map): one bilinear gather per sample
            const RoiRecord& g = recs[roi];
            const float* fb = feat + (size_t)g.b * Hf * Wf * C + ch0 + lane * 2;
#pragma unroll 1
            for (int bin = 0; bin < kP * kP; ++bin) {
                const int py = bin / kP, px = bin - py * kP;
                float2 best = make_float2(0.0f, 0.0f);
                unsigned cw = 0u;
#pragma unroll 1
                for (int s4 = 0; s4 < 4; ++s4) {
                    const int sy = 2 * py + (s4 >> 1), sx = 2 * px + (s4 & 1);
                    float2 val = make_float2(0.0f, 0.0f);  // extrapolation value
                    if (((g.yok >> sy) & 1u) && ((g.xok >> sx) & 1u)) {
                        const float* top = fb + (size_t)g.ylo[sy] * Wf * C;
                        const float* bot = fb + (size_t)g.yhi[sy] * Wf * C;
                        const size_t xl = (size_t)g.xlo[sx] * C, xh = (size_t)g.xhi[sx] * C;
                        const u64 tl = *reinterpret_cast<const u64*>(top + xl), tr = *reinterpret_cast<const u64*>(top + xh);
                        const u64 bl = *reinterpret_cast<const u64*>(bot + xl), br = *reinterpret_cast<const u64*>(bot + xh);
                        const u64 wx = pack2(g.wx[sx], g.wx[sx]), wy = pack2(g.ylerp[sy], g.ylerp[sy]);
                        val = unpack2(lerp2(lerp2(tl, tr, wx, one), lerp2(bl, br, wx, one), wy, one));
                    }
                    if (s4 == 0) {
                        best = val;
                    } else {
                        take_if_greater(best.x, cw, val.x, (unsigned)s4, 0);
                        take_if_greater(best.y, cw, val.y, (unsigned)s4, 2);
                    }
                }
                __stcs(optr + (size_t)bin * bin_stride, best);
                if (kCodes) {
                    const unsigned pw = __shfl_xor_sync(all, cw, 1);
                    if (!(lane & 1)) cptr[(size_t)bin * (C / 4)] = (unsigned char)((cw & 15u) | ((pw & 15u) << 4));
                }
            }
        } else {
            // ---- the walk of roi_pool_fwd_stream_kernel, on this warp's own ring area
            const unsigned lane_ring = ring_s + lane * 8;
            if (it.first_off) {  // the zero pixel of every slot (each lane later reads only what it wrote itself)
                for (int s = 0; s < it.depth; ++s)
                    asm volatile("st.shared.b64 [%0], %1;" ::"r"(lane_ring + it.area + s * it.slot_bytes), "l"(0ull) : "memory");
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // later TMA writes may land on these bytes
            }
            // aoff[i]: this lane's shared-memory address of sample i's left pixel in slot 0 of the warp's ring (registers for
            // the whole item; the slot offset is warp-uniform, so a row's operand is `register + uniform register +
            // immediate`).  The x weights stay in the record: a row re-reads them with four broadcast 128-bit loads, which is
            // cheaper than 14 more live registers next to the two row sets.
            int aoff[kS];
#pragma unroll
            for (int i = 0; i < kS; ++i) aoff[i] = (int)(lds_u32(it.rec + 56 + 4 * i) + lane_ring);
            u64 prog = ((u64)uniform(lds_u32(it.rec + 176)) << 32) | uniform(lds_u32(it.rec + 172));
            const unsigned nba = uniform(lds_u32(it.rec + 188));
            const int n_before = (int)((nba >> 8) & 0xffu), n_after = (int)((nba >> 16) & 0xffu);
            u64 X0[kS], X1[kS];  // (always written by load_row before a crop row inside the map reads them)
            float2 M[kP];
            int sy = 0;
            float yl_next = lds_f32(it.rec + 112);  // the y weight of the next crop row, loaded one crop row ahead
            unsigned ca = 0u, cb = 0u;
            int k = 0, slot = 0;
            unsigned slot_off = it.area;
            // x-interpolate the next staged feature row into X, then re-arm its slot with row k + depth
            auto load_row = [&](u64(&X)[kS]) {
                // (the feature row that will re-arm this slot: read from the record before the wait, so that the elected
                // lane does not sit on the shared-memory latency when it builds the TMA coordinate)
                const int next_row = (int)lds_u16(it.rec + 192 + 2 * (k + it.depth));
                const unsigned bit = it.bar + slot;
                mbar_wait(bar_s + bit * 8, (row_phase >> bit) & 1u);
                row_phase ^= 1u << bit;
#pragma unroll
                for (int h = 0; h < 2; ++h) {  // samples 0..7, then 8..13: the x weights come eight at a time (two broadcast
                    float wx[8];               // 128-bit loads of record bytes 0..63), which keeps fewer of them alive
#pragma unroll
                    for (int q = 0; q < 2; ++q)
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                                     : "=f"(wx[4 * q]), "=f"(wx[4 * q + 1]), "=f"(wx[4 * q + 2]), "=f"(wx[4 * q + 3])
                                     : "r"(it.rec + 32 * h + 16 * q));
#pragma unroll
                    for (int i = 8 * h; i < (h ? kS : 8); ++i) {
                        u64 l, r;
                        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(l) : "r"(aoff[i] + slot_off));
                        asm volatile("ld.shared.b64 %0, [%1+256];" : "=l"(r) : "r"(aoff[i] + slot_off));
                        X[i] = lerp2(l, r, pack2(wx[i - 8 * h], wx[i - 8 * h]), one);
                    }
                }
                __syncwarp();  // every lane has consumed its operands: the slot may be overwritten
                if (k + it.depth < it.n_rows) stage_row(it, ch0, next_row, slot);
                ++k;
                slot_off += it.slot_bytes;
                if (++slot == it.depth) {
                    slot = 0;
                    slot_off = it.area;
                }
            };
            // One finished crop row: fold it into the bin maxima; after the second row of a bin-row store the 7 bins.
            auto emit = [&](const u64(&T)[kS], const u64(&Bm)[kS], bool outside) {
                const float yl = yl_next;
                yl_next = lds_f32(it.rec + 112 + 4 * (sy + 1));  // (after the last crop row this reads the word behind ylerp: unused)
                if (!(sy & 1)) {
                    ca = cb = 0u;
                    if (outside) {  // crop row outside the map: extrapolation value 0 takes part in the max
#pragma unroll
                        for (int j = 0; j < kP; ++j) M[j] = make_float2(0.0f, 0.0f);
                    } else {
                        crop_row<true, kCodes, true>(T, Bm, yl, one, M, ca, cb);
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
                        crop_row<false, kCodes, true>(T, Bm, yl, one, M, ca, cb);
                    }
#pragma unroll
                    for (int j = 0; j < kP; ++j) __stcs(optr + j * bin_stride, M[j]);
                    optr += kP * bin_stride;
                    if (kCodes) {  // a byte holds four channels = this lane's nibble and its odd neighbour's
                        const unsigned cw = codes_from_flags(ca, cb);
                        const unsigned pw = __shfl_xor_sync(all, cw, 1);
                        if (!(lane & 1)) {
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
            if (it.n_rows > 0) {
#pragma unroll 1
                for (;;) {  // (the exit in the middle keeps the second half unconditional: no merge copies of a row set)
                    load_row(X0);  // even position: the new bottom row lands in set 0, set 1 holds the top row
                    prog >>= 1;
                    while (prog & 1ull) {
                        emit(X1, X0, false);
                        prog >>= 1;
                    }
                    if (k >= it.n_rows) break;
                    load_row(X1);
                    prog >>= 1;
                    while (prog & 1ull) {
                        emit(X0, X1, false);
                        prog >>= 1;
                    }
                    if (k >= it.n_rows) break;
                }
            }
            for (int i = 0; i < n_after; ++i) emit(X0, X1, true);
        }
        // ---- next item
        idx_cur = (int)uniform((unsigned)idx_nxt);  // uniform loop exit: everything carried around the loop stays uniform
        idx_nxt = idx_n2;
        if (idx_cur >= n_items) break;
        buf = nbuf;
        prefetched = uniform(next_prefetched ? 1u : 0u) != 0u;
        half = (int)uniform(next_prefetched ? (unsigned)(half ^ 1) : 0u);
    }
}

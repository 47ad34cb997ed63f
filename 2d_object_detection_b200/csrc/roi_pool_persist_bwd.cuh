// Persistent RoI pooling backward from saved arg-max codes (included by roi_pool_stream.cu after the per-CTA backward: it
// reuses BwdMaps, bwd_crop_row, the TMA reduce-add helpers -- the arithmetic is unchanged).
//
// Same worker model as the persistent forward (roi_pool_persist.cuh): one 16-warp CTA per SM, every warp pulls
// (RoI, 64-channel) items from a global counter, reads the RoI's record (built once by roi_records_kernel) from shared
// memory and never synchronises with another warp.  What the per-CTA kernel paid per (RoI, 256 channels) CTA -- the
// two-warp prologue behind a __syncthreads, the exposed latency of the first dY rows, a three-way branch per bin in the
// x-scatter -- is gone:
//   * dY + codes arrive as a CONTINUOUS stream of bin-rows (7 bins x 64 channels + 7 x 16 code bytes per TMA pair) through
//     a 3-slot ring that runs across item boundaries: while the last three bin-rows of an item are consumed, the first
//     three of the next item are already in flight.
//   * the x-scatter of a finished feature row keeps the two pixels under the current sample in REGISTERS: the crop columns
//     walk left to right, so a pixel is complete as soon as the walk moves past it and is stored once (the record says how
//     far each column advances) -- no read-modify-write chain through shared memory.  One cp.reduce.async.bulk.tensor
//     then adds the row buffer to dFeature.  Reversed boxes (and RoIs wider than the ring) take a plain atomic path.
#pragma once

constexpr int kPBwdRowPix = 28;                              // row-buffer ring of a warp, pixels
constexpr int kPBwdDyOff = kPBwdRowPix * kPixBytes;
constexpr int kPBwdSlotBytes = kBwdDyBytes + kBwdCodeBytes;  // 1792 B of dY + 128 B of codes
constexpr int kPBwdRecOff = kPBwdDyOff + kBwdDySlots * kPBwdSlotBytes;
constexpr int kPBwdRecBufs = 2;
constexpr int kPBwdBarOff = kPBwdRecOff + kPBwdRecBufs * kRecBytes;  // 3 dY barriers + 2 record barriers, padded
constexpr int kPBwdWarpBytes = kPBwdBarOff + 128;
static_assert(kPBwdWarpBytes % 128 == 0, "per-warp region must stay 128-byte aligned");
static_assert(kPWarps * kPBwdWarpBytes <= 227 * 1024, "one CTA per SM must fit");

__global__ void __launch_bounds__(kPThreads, 1)
    roi_pool_bwd_persist_kernel(const __grid_constant__ BwdMaps maps, int Hf, int Wf, int C, const RoiRecord* __restrict__ recs,
                                int n_items, const float* __restrict__ dy, const unsigned char* __restrict__ codes,
                                float* __restrict__ dfeat, unsigned* __restrict__ counter) {
    const int nch = C / kWarpCh;
    extern __shared__ __align__(128) unsigned char smem[];
    const unsigned all = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = (int)uniform(threadIdx.x >> 5);
    const unsigned warp_s = smem_u32(smem) + warp * kPBwdWarpBytes;  // row buffers | scratch | dY slots | records | barriers
    const unsigned lane_s = warp_s + lane * 8;
    const unsigned dy_s = warp_s + kPBwdDyOff, rec_s = warp_s + kPBwdRecOff, bar_s = warp_s + kPBwdBarOff;
    if (lane < kBwdDySlots + kPBwdRecBufs) {
        mbar_init(bar_s + lane * 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int total_warps = gridDim.x * kPWarps;
    int idx_cur = blockIdx.x * kPWarps + warp;  // the first item (or two: second_item_is_static) is static
    int idx_nxt = idx_cur + total_warps;
    if (idx_cur >= n_items) return;

    auto fetch_record = [&](int item, int buf) {
        if (lane == 0) {
            const unsigned bar = bar_s + (kBwdDySlots + buf) * 8;
            mbar_expect_tx(bar, kRecBytes);
            bulk_g2s(rec_s + buf * kRecBytes, reinterpret_cast<const unsigned char*>(recs) + (size_t)(item / nch) * kRecStride, kRecBytes,
                     bar);
        }
    };
    // dY + codes of bin-row `py` of `item` into ring slot `slot` (one lane, two TMA boxes on one barrier)
    // (`bin` = first bin of the bin-row in the [B*R*49, C] view of dY, `ch0` = first channel of the item)
    auto stage_bin_row = [&](int ch0, int bin, int slot) {
        if (lane == 0) {
            const unsigned bar = bar_s + slot * 8;
            mbar_expect_tx(bar, kBwdDyBytes + kP * 16);
            tma_box_g2s(dy_s + slot * kPBwdSlotBytes, &maps.dy, ch0, bin, bar);
            tma_box_g2s(dy_s + slot * kPBwdSlotBytes + kBwdDyBytes, &maps.codes, ch0 / 4, bin, bar);
        }
    };

    unsigned dphase = 0u, rec_phase = 0u;  // parity of the next completion of the dY / record barriers
    int dslot = 0, buf = 0;
    fetch_record(idx_cur, 0);
    unsigned ticket = 0u;
    const unsigned ticket0 = first_ticket(n_items, total_warps);  // tickets count from 0: item = ticket + ticket0
    if (!second_item_is_static(n_items, total_warps)) {  // few items per warp: the second one comes from the counter too
        if (lane == 0) ticket = take_ticket(counter) + ticket0;
        idx_nxt = (int)uniform(lane == 0 ? ticket : 0u);
    }
    if (lane == 0) ticket = take_ticket(counter) + ticket0;
    {  // cold start of the stream of bin-rows
        const int roi = idx_cur / nch, ch0 = (idx_cur - roi * nch) * kWarpCh;
#pragma unroll 1
        for (int py = 0; py < kBwdDySlots; ++py) stage_bin_row(ch0, roi * (kP * kP) + py * kP, py);
    }

#pragma unroll 1
    for (;;) {
        // ---- this item's record (fetched during the previous item), the next item's record and index
        mbar_wait(bar_s + (kBwdDySlots + buf) * 8, (rec_phase >> buf) & 1u);
        rec_phase ^= 1u << buf;
        const int idx_n2 = (int)uniform(lane == 0 ? ticket : 0u);
        const bool has_next = idx_nxt < n_items;
        if (has_next) {
            fetch_record(idx_nxt, buf ^ 1);
            if (idx_n2 < n_items && lane == 0) ticket = take_ticket(counter) + ticket0;
        }
        const unsigned rec = rec_s + buf * kRecBytes;
        const unsigned f = uniform(lds_u32(rec + 184)), g8 = uniform(lds_u32(rec + 188));
        const int n_rows = (int)(f & 0xffu), box = (int)((f >> 8) & 0xffu), cls = (int)(f >> 24);
        const int first_off = (int)((f >> 16) & 0xffu) * kPixBytes;
        const int n_before = (int)((g8 >> 8) & 0xffu), n_after = (int)((g8 >> 16) & 0xffu);
        const int pix0 = (int)uniform(lds_u32(rec + 168));
        const int roi = idx_cur / nch, ch0 = (idx_cur - roi * nch) * kWarpCh;
        const int roi_n = has_next ? idx_nxt / nch : 0, ch0_n = has_next ? (idx_nxt - roi_n * nch) * kWarpCh : 0;
        const int bin0 = roi * (kP * kP), bin0_n = roi_n * (kP * kP);
        const bool mono = (uniform(lds_u32(rec + 252)) & 0xffu) != 0u;
        // no column inside the map or wider than the forward's ring: plain atomics.  The widest box class (27 .. 52 columns)
        // and reversed boxes (columns right to left) keep the register sets but scatter sample by sample ("wide").
        const bool slow = cls == kClsGather;
        const bool wide = !slow && (box >= kBwdNumBoxes - 1 || !mono);
        const int row_bytes = (int)(uniform(lds_u32(rec + 180)) >> 16);  // tx_bytes of the record = W * 256: the forward's box
        const int W = row_bytes / kPixBytes;
        const int nbufs = slow ? 1 : wide ? 2 : max(1, min(kBwdMaxBufs, kPBwdRowPix / max(W, 1)));

        int sy = 0;
        float yl_next = lds_f32(rec + 112);
        unsigned cw = 0u;
        int cw_py = -1;
        // One bin-row consumed: release its slot to the bin-row three further down the stream (this item's, or the next's)
        auto release_bin_row = [&](int py) {
            __syncwarp();
            // bin-row py + 3 of this item, or bin-row py + 3 - 7 of the next one
            const bool mine = py + kBwdDySlots < kP;
            if (mine || has_next)
                stage_bin_row(mine ? ch0 : ch0_n, (mine ? bin0 + kBwdDySlots * kP : bin0_n + (kBwdDySlots - kP) * kP) + py * kP, dslot);
            if (++dslot == kBwdDySlots) dslot = 0;
        };
        auto wait_bin_row = [&]() {  // first crop row of a bin-row: its dY + codes must have landed
            mbar_wait(bar_s + dslot * 8, (dphase >> dslot) & 1u);
            dphase ^= 1u << dslot;
        };
        auto load_codes = [&]() {  // this lane's 2-bit arg-max codes (2 channels x 7 bins) of the bin-row in the current slot
            cw = 0u;
#pragma unroll
            for (int j = 0; j < kP; ++j) {
                unsigned byte;
                asm volatile("ld.shared.u8 %0, [%1];" : "=r"(byte) : "r"(dy_s + dslot * kPBwdSlotBytes + kBwdDyBytes + j * 16 + (lane >> 1)));
                cw |= ((byte >> ((lane & 1) * 4)) & 15u) << (4 * j);
            }
        };

        if (slow) {
            // rare: one atomic per bilinear share, TF's operation order; dY / codes still come through the ring
            const RoiRecord& g = recs[roi];
            float* gb = dfeat + (size_t)g.b * Hf * Wf * C + ch0 + lane * 2;
#pragma unroll 1
            for (int py = 0; py < kP; ++py) {
                wait_bin_row();
                load_codes();
#pragma unroll 1
                for (int px = 0; px < kP; ++px) {
                    const float2 g2 = unpack2(lds64(lane_s + kPBwdDyOff + dslot * kPBwdSlotBytes + px * kPixBytes));
#pragma unroll 1
                    for (int cidx = 0; cidx < 2; ++cidx) {
                        const unsigned pos = (cw >> (4 * px + 2 * cidx)) & 3u;
                        const int ssy = 2 * py + (int)(pos >> 1), ssx = 2 * px + (int)(pos & 1u);
                        if (!(((g.yok >> ssy) & 1u) && ((g.xok >> ssx) & 1u))) continue;  // the extrapolation value won
                        const float gv = cidx ? g2.y : g2.x;
                        const float yl = g.ylerp[ssy], xl = g.wx[ssx];
                        const float dtop = fmul(fsub(1.0f, yl), gv), dbot = fmul(yl, gv);
                        float* top = gb + (size_t)g.ylo[ssy] * Wf * C + cidx;
                        float* bot = gb + (size_t)g.yhi[ssy] * Wf * C + cidx;
                        atomicAdd(top + (size_t)g.xlo[ssx] * C, fmul(fsub(1.0f, xl), dtop));
                        atomicAdd(top + (size_t)g.xhi[ssx] * C, fmul(xl, dtop));
                        atomicAdd(bot + (size_t)g.xlo[ssx] * C, fmul(fsub(1.0f, xl), dbot));
                        atomicAdd(bot + (size_t)g.xhi[ssx] * C, fmul(xl, dbot));
                    }
                }
                release_bin_row(py);
            }
        } else {
            // aoff[i]: this lane's shared address of sample i's left pixel in row buffer 0.  A sample outside the map (record
            // offset 0 while the record reserves a zero pixel) may hold a gradient -- the extrapolation value can win a
            // window -- that goes nowhere: it scatters to pixel 0 with weights okf - xl = 0 - 0 and xl = 0, while a sample
            // inside the map gets okf - xl = 1 - xl, the same rounding as before.
            int aoff[kS];
            float okf[kS];
#pragma unroll
            for (int i = 0; i < kS; ++i) {
                const int a = (int)lds_u32(rec + 56 + 4 * i);
                const bool outside = first_off != 0 && a == 0;
                aoff[i] = (int)lane_s + (outside ? 0 : a - first_off);
                okf[i] = outside ? 0.0f : 1.0f;
            }
            u64 prog = ((u64)uniform(lds_u32(rec + 176)) << 32) | uniform(lds_u32(rec + 172));
            const unsigned adv_mask = uniform(lds_u32(rec + 248));
            u64 G0[kS], G1[kS];
#pragma unroll
            for (int i = 0; i < kS; ++i) G0[i] = G1[i] = 0ull;
            int rbuf = 0, issued = 0;
            unsigned rbuf_off = 0u;
            // One crop row: wait for its bin-row's dY, accumulate (unless it lies outside the map), release the slot after
            // the second row of the bin-row.
            auto crop = [&](bool g0_top, bool outside) {
                const float yl = yl_next;
                yl_next = lds_f32(rec + 112 + 4 * (sy + 1));
                if (!(sy & 1)) wait_bin_row();
                if (!outside) {
                    if (cw_py != (sy >> 1)) {
                        load_codes();
                        cw_py = sy >> 1;
                    }
                    unsigned dy_addr = lane_s + kPBwdDyOff + dslot * kPBwdSlotBytes;
                    asm volatile("" : "+r"(dy_addr));
                    bwd_crop_row(G0, G1, g0_top, (sy & 1) != 0, yl, dy_addr, cw);
                }
                if (sy & 1) release_bin_row(sy >> 1);
                ++sy;
            };
            // The walk has left feature row `kk` of the record: scatter its set along x into a zeroed row buffer and hand the
            // buffer to the TMA engine, which adds it to dFeature.
            auto flush = [&](u64(&G)[kS], int kk) {
                if (wide) {
                    // No carry: sample i writes its two weighted pixels into its own pair of buffer pixels and gets its own
                    // 2-pixel reduce-add (overlapping pixels of neighbouring samples meet in L2).  The 28-pixel ring is two
                    // halves of 7 samples, one bulk group each, so a half is rewritten while the other is still being read.
                    float wx[16];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                                     : "=f"(wx[4 * q]), "=f"(wx[4 * q + 1]), "=f"(wx[4 * q + 2]), "=f"(wx[4 * q + 3])
                                     : "r"(rec + 16 * q));
                    const int row = (int)lds_u16(rec + 192 + 2 * kk);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (issued >= 2) {
                            if (lane == 0) bulk_wait_read(1);
                            __syncwarp();
                        }
                        const unsigned half_off = h * (kS * kPixBytes);
#pragma unroll
                        for (int q = 0; q < kS / 2; ++q) {
                            const int i = (kS / 2) * h + q;
                            const float xl = wx[i];
                            const float xr = fsub(okf[i], xl);
                            sts64(lane_s + half_off + (2 * q) * kPixBytes, mul2(G[i], pack2(xr, xr)));
                            sts64(lane_s + half_off + (2 * q + 1) * kPixBytes, mul2(G[i], pack2(xl, xl)));
                            asm volatile("sub.rn.f32x2 %0, %0, %0;" : "+l"(G[i]));
                        }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) {
#pragma unroll 1
                            for (int q = 0; q < kS / 2; ++q) {
                                const int a = (int)lds_u32(rec + 56 + 4 * ((kS / 2) * h + q));  // 0: outside, both weights 0
                                const int xoff = a == 0 ? 0 : (a - first_off) / kPixBytes;
                                tma_reduce_add_s2g_nocommit(&maps.dfeat[kBwdNumBoxes - 1], ch0, pix0 + row * Wf + xoff,
                                                            warp_s + half_off + (2 * q) * kPixBytes);
                            }
                            bulk_commit();
                        }
                        ++issued;
                    }
                    return;
                }
                if (issued >= nbufs) {  // the buffer about to be reused must have been read by its reduce
                    if (lane == 0) bulk_wait_read(nbufs - 1);
                    __syncwarp();
                }
                const unsigned bufa = lane_s + rbuf_off;
                {  // zero the W pixels of the buffer: four per iteration, then the remainder (W is 6 .. 26)
                    int p = 0;
#pragma unroll 1
                    for (; p + 4 <= W; p += 4) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) sts64(bufa + (p + q) * kPixBytes, 0ull);
                    }
                    if (p + 2 <= W) {
                        sts64(bufa + p * kPixBytes, 0ull);
                        sts64(bufa + (p + 1) * kPixBytes, 0ull);
                        p += 2;
                    }
                    if (p < W) sts64(bufa + p * kPixBytes, 0ull);
                }
                {
                    // the two pixels under the current sample live in registers; a pixel is stored once, when the walk has
                    // moved past it (adv: 1 = one pixel on, 2 = both pixels done); same order of additions as the
                    // read-modify-write form, so the same bits
                    u64 acc_lo, acc_hi;
                    asm volatile("mov.b64 %0, 0;" : "=l"(acc_lo));
                    asm volatile("mov.b64 %0, 0;" : "=l"(acc_hi));
                    float wx[16];
#pragma unroll
                    for (int q = 0; q < 4; ++q)  // record bytes 0 .. 63: the 14 x weights
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                                     : "=f"(wx[4 * q]), "=f"(wx[4 * q + 1]), "=f"(wx[4 * q + 2]), "=f"(wx[4 * q + 3])
                                     : "r"(rec + 16 * q));
#pragma unroll
                    for (int i = 0; i < kS; ++i) {
                        const float xl = wx[i];
                        const float xr = fsub(okf[i], xl);
                        acc_lo = fma2(G[i], pack2(xr, xr), acc_lo);
                        acc_hi = fma2(G[i], pack2(xl, xl), acc_hi);
                        // G - G = 0 in ONE packed instruction (a literal zero here becomes a uniform-register move plus two copies;
                        // ptxas cannot fold x - x for floats)
                        asm volatile("sub.rn.f32x2 %0, %0, %0;" : "+l"(G[i]));
                        const unsigned m = (adv_mask >> (2 * i)) & 3u;  // warp-uniform
                        if (m) {
                            const unsigned a = (unsigned)aoff[i] + rbuf_off;
                            sts64(a, acc_lo);
                            if (m & 2u) {
                                sts64(a + kPixBytes, acc_hi);
                                acc_lo = 0ull;
                            } else {
                                acc_lo = acc_hi;
                            }
                            acc_hi = 0ull;
                        }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the TMA engine
                __syncwarp();
                if (lane == 0) {
                    const int row = (int)lds_u16(rec + 192 + 2 * kk);
                    tma_reduce_add_s2g(&maps.dfeat[box], ch0, pix0 + row * Wf, warp_s + rbuf_off);
                }
                ++issued;
                rbuf_off += row_bytes;
                if (++rbuf == nbufs) {
                    rbuf = 0;
                    rbuf_off = 0u;
                }
            };
            for (int i = 0; i < n_before; ++i) crop(false, true);
            // Walk position k enters register set k & 1: the row that set held (position k - 2) is complete and is flushed;
            // then the crop rows between positions k - 1 (top, the other set) and k (bottom, this set) are accumulated.
#pragma unroll 1
            for (int k = 0; k < n_rows + 2; ++k) {
                const bool odd = (k & 1) != 0;
                if (k >= 2) {
                    if (odd) flush(G1, k - 2);
                    else flush(G0, k - 2);
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
            // the next item has its own row-buffer geometry: every reduce of this item must have read its buffer
            if (lane == 0) bulk_wait_read(0);
            __syncwarp();
        }
        // ---- next item
        idx_cur = (int)uniform((unsigned)idx_nxt);
        idx_nxt = idx_n2;
        if (idx_cur >= n_items) break;
        buf ^= 1;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all reduce-adds performed before exit
}

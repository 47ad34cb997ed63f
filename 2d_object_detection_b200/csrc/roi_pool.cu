// RoI pooling: models/detectors/fast_rcnn_detector.py:133-177 (ROIPooling) forward and the backward that
// tf.GradientTape derives for it (models/faster_rcnn.py:73,103): tf.image.crop_and_resize to S x S
// (S = pooled_size * kernel_size, one bilinear sample per crop pixel, (H-1) scaling, extrapolation 0)
// fused with MaxPool2D(kernel_size), never materialising the S x S crop (traps T6-T8).
//
// Layout: feature maps NHWC, so the 4 neighbours of a sample are 4 contiguous C-float rows; a thread
// owns one float4 of channels and walks the bins of one RoI, so every global access is a coalesced
// 128-bit load/store.  Arithmetic follows the TF CPU kernel operation by operation with explicit _rn
// intrinsics: the forward result is bit-identical to the CPU oracle.
#include "roi_axis.cuh"

namespace frcnn {

// Compile-time A/B switch (FRCNN_DEFINES=FRCNN_AB_BINROW_WALK): route the reference shape to the older bin-row walk
// kernels instead of the row-streaming ones.  Forward codes and backward always switch together.  Never set in the
// shipped build: one path per shape, no run-time dispatch on the environment.
#ifdef FRCNN_AB_BINROW_WALK
constexpr bool kForceBinRowWalk = true;
#else
constexpr bool kForceBinRowWalk = false;
#endif

// ---- tiny vector abstraction: float4 (C % 4 == 0) or float -------------------------------------
__device__ __forceinline__ float4 vload(const float4* p) { return __ldg(p); }
__device__ __forceinline__ float vload(const float* p) { return __ldg(p); }
__device__ __forceinline__ float4 vzero(float4) { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float vzero(float) { return 0.f; }
__device__ __forceinline__ float lerp1(float a, float b, float w) { return fadd(a, fmul(fsub(b, a), w)); }
__device__ __forceinline__ float lerpv(float a, float b, float w) { return lerp1(a, b, w); }
__device__ __forceinline__ float4 lerpv(const float4& a, const float4& b, float w) {
    return make_float4(lerp1(a.x, b.x, w), lerp1(a.y, b.y, w), lerp1(a.z, b.z, w), lerp1(a.w, b.w, w));
}
__device__ __forceinline__ void vstore_stream(float4* p, const float4& v) { st_cs_f4(p, v); }
__device__ __forceinline__ void vstore_stream(float* p, float v) { __stcs(p, v); }

// One bilinear sample (crop pixel) for this thread's channels: top = TL + (TR-TL)*xl, etc.
template <typename V>
__device__ __forceinline__ V bilinear(const V* __restrict__ fb, int Wf, int CV, int ylo, int yhi, float yl, int xlo,
                                      int xhi, float xl, int cv) {
    const V tl = vload(fb + ((size_t)ylo * Wf + xlo) * CV + cv);
    const V tr = vload(fb + ((size_t)ylo * Wf + xhi) * CV + cv);
    const V bl = vload(fb + ((size_t)yhi * Wf + xlo) * CV + cv);
    const V br = vload(fb + ((size_t)yhi * Wf + xhi) * CV + cv);
    const V top = lerpv(tl, tr, xl);
    const V bot = lerpv(bl, br, xl);
    return lerpv(top, bot, yl);
}

__device__ __forceinline__ void max_first(float& best, float v, int& arg, int s) {
    if (v > best) {  // strict: the first maximum in row-major window order wins (T8)
        best = v;
        arg = s;
    }
}

template <typename V>
__global__ void __launch_bounds__(kRoiThreads)
    roi_pool_fwd_kernel(const V* __restrict__ feat, int Hf, int Wf, int CV, const float4* __restrict__ rois, int R, int P,
                        int k, V* __restrict__ out) {
    __shared__ AxisTab ty, tx;
    const int r = blockIdx.x, b = blockIdx.y;
    const int S = P * k;
    const float4 roi = rois[(size_t)b * R + r];  // x1,y1,x2,y2 (fast_rcnn_detector.py:157-158 reorders for TF)
    if (threadIdx.x < S) axis_sample(roi.y, roi.w, Hf, S, threadIdx.x, ty);
    else if (threadIdx.x >= 32 && threadIdx.x < 32 + S) axis_sample(roi.x, roi.z, Wf, S, threadIdx.x - 32, tx);
    __syncthreads();
    const V* fb = feat + (size_t)b * Hf * Wf * CV;
    V* ob = out + ((size_t)b * R + r) * P * P * CV;
    for (int cv = threadIdx.x; cv < CV; cv += blockDim.x) {
        for (int py = 0; py < P; ++py)
            for (int px = 0; px < P; ++px) {
                V best = vzero(V());
                bool first = true;
                for (int dy = 0; dy < k; ++dy) {
                    const int sy = py * k + dy;
                    for (int dx = 0; dx < k; ++dx) {
                        const int sx = px * k + dx;
                        V v = vzero(V());  // extrapolation_value = 0
                        if (ty.valid[sy] && tx.valid[sx])
                            v = bilinear(fb, Wf, CV, ty.lo[sy], ty.hi[sy], ty.lerp[sy], tx.lo[sx], tx.hi[sx],
                                         tx.lerp[sx], cv);
                        if (first) {
                            best = v;
                            first = false;
                        } else {
                            if constexpr (sizeof(V) == 16) {
                                best.x = v.x > best.x ? v.x : best.x;
                                best.y = v.y > best.y ? v.y : best.y;
                                best.z = v.z > best.z ? v.z : best.z;
                                best.w = v.w > best.w ? v.w : best.w;
                            } else {
                                best = v > best ? v : best;
                            }
                        }
                    }
                }
                vstore_stream(ob + ((size_t)py * P + px) * CV + cv, best);
            }
    }
}

// ------------------------------------------------------------------------------------------------
// Fast forward path for kernel_size == 2 (the reference configuration), float4 channels.
// A thread owns one float4 of channels and walks one bin-row (two crop rows) from left to right.
// The four feature rows it touches (top/bottom of both crop rows) and the current left/right
// feature columns live in registers: when the next sample falls in the same cell nothing is loaded,
// when it moves one cell right the right column is shifted into the left one and only one new column
// is fetched.  Rows shared between the two crop rows (the common case: crop step < 1 feature cell)
// are fetched and x-interpolated once.  All branch conditions depend on the RoI only, so they are
// uniform over the CTA.  The arithmetic per sample is unchanged (bit-identical to the generic kernel).
// ------------------------------------------------------------------------------------------------
struct AxisY {             // per crop row: element offsets (in float4) of the two feature rows
    int lo[kMaxCrop];
    int hi[kMaxCrop];
    float lerp[kMaxCrop];
    int valid[kMaxCrop];
};
struct AxisX {             // per crop column: feature cell (left column index, -1 = outside the map) and x_lerp
    int cell[kMaxCrop];
    float lerp[kMaxCrop];
};

// Packed fp32x2 lerp a + (b - a) * w for a float4 (Blackwell FADD2 / FMUL2 / FFMA2: two fp32 lanes per issue slot).
// Every element still sees sub.rn, mul.rn, add.rn SEPARATELY rounded, i.e. the same bits as lerpv() and as the
// reference's unfused CPU arithmetic.  The catch: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 (single
// rounding) even under -fmad=false, and also folds fma(p, 1.0f, a).  So the final add is written as
// fma.rn.f32x2(p, ONE, a) with ONE read from constant memory at run time: the compiler cannot fold it, and
// p * 1 + a rounds exactly like p + a.
__device__ __constant__ float kRuntimeOne = 1.0f;

__device__ __forceinline__ float4 lerp_packed(const float4& a, const float4& b, float w) {
    unsigned long long a0, a1, b0, b1, d0, d1, ww, one;
    const float onef = kRuntimeOne;
    asm("mov.b64 %0, {%1,%2};" : "=l"(a0) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1,%2};" : "=l"(a1) : "f"(a.z), "f"(a.w));
    asm("mov.b64 %0, {%1,%2};" : "=l"(b0) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1,%2};" : "=l"(b1) : "f"(b.z), "f"(b.w));
    asm("mov.b64 %0, {%1,%1};" : "=l"(ww) : "f"(w));
    asm("mov.b64 %0, {%1,%1};" : "=l"(one) : "f"(onef));
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d0) : "l"(b0), "l"(a0));
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d1) : "l"(b1), "l"(a1));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d0) : "l"(d0), "l"(ww));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d1) : "l"(d1), "l"(ww));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d0) : "l"(d0), "l"(one), "l"(a0));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d1) : "l"(d1), "l"(one), "l"(a1));
    float4 r;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(d0));
    asm("mov.b64 {%0,%1}, %2;" : "=f"(r.z), "=f"(r.w) : "l"(d1));
    return r;
}

__device__ __forceinline__ float4 vmax4(const float4& a, const float4& b) {
    return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}

// Walk one bin-row (two crop rows) left to right, cell by cell.
// NR = number of distinct feature rows under the two crop rows:
//   2: both crop rows in the same cell        (rows t, b;          s0 = y(X0,X1), s1 = y(X0,X1))
//   3: second crop row one cell further down  (rows t0, b0=t1, b1;   s0 = y(X0,X1), s1 = y(X1,X2))
//   4: anything else, including crop rows outside the map (kCheckY) -- no sharing assumed.
// VPT = float4 channel vectors per thread (thread t owns vectors t, t+T, ...: each warp access stays a
// contiguous 512 B); VPT = 2 halves the per-sample index/branch overhead relative to the fp32 work.
// Two register column sets A and B alternate as left/right neighbour: moving one cell to the right loads
// ONE new column into the set that held the old left column and swaps roles (two copies of the sample
// code, no register shuffling); every feature value is loaded exactly once per bin-row.  The right
// neighbour of an integer position is taken as column c+1 instead of TF's ceil() == c: its weight is
// x_lerp == 0, so the interpolated value is identical.  All branches are CTA-uniform.
template <int NR, bool kCheckY, int VPT, int T>
__device__ __forceinline__ void walk_bin_row(const float4* __restrict__ f, const int (&roff)[4], float yl0, float yl1,
                                             bool yv0, bool yv1, const AxisX& ax, int S, int last_col, int CV,
                                             float4* __restrict__ orow) {
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* fr[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) fr[k] = f + roff[k];
    float4 A[NR][VPT], B[NR][VPT];
    float4 best[VPT];
    float4* optr = orow;
    auto load_col = [&](float4 (&dst)[NR][VPT], int col) {
        const int off = min(col, last_col) * CV;
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            const float4* p = fr[k] + off;
#pragma unroll
            for (int v = 0; v < VPT; ++v) dst[k][v] = __ldg(p + v * T);
        }
    };
    auto emit_zero = [&](int sx) {
        if (sx & 1) {
#pragma unroll
            for (int v = 0; v < VPT; ++v) st_cs_f4(optr + v * T, vmax4(best[v], zero));
            optr += CV;
        } else {
#pragma unroll
            for (int v = 0; v < VPT; ++v) best[v] = zero;
        }
    };
    auto sample_v = [&](const float4 (&L)[NR][VPT], const float4 (&Rt)[NR][VPT], float xl, int v) -> float4 {
        float4 X[NR];
#pragma unroll
        for (int k = 0; k < NR; ++k) X[k] = lerp_packed(L[k][v], Rt[k][v], xl);
        float4 s0 = lerp_packed(X[0], X[1], yl0);
        float4 s1 = NR == 2 ? lerp_packed(X[0], X[1], yl1)
                            : NR == 3 ? lerp_packed(X[1], X[2], yl1) : lerp_packed(X[NR - 2], X[NR - 1], yl1);
        if (kCheckY) {
            if (!yv0) s0 = zero;  // extrapolation_value = 0 for a crop row outside the map
            if (!yv1) s1 = zero;
        }
        return vmax4(s0, s1);
    };
    // even crop column: first half of a bin, result goes straight into `best`; odd: finish the bin and store it
    auto emit = [&](int sx, const float4 (&L)[NR][VPT], const float4 (&Rt)[NR][VPT]) {
        const float xl = ax.lerp[sx];
        if (sx & 1) {
#pragma unroll
            for (int v = 0; v < VPT; ++v) st_cs_f4(optr + v * T, vmax4(best[v], sample_v(L, Rt, xl, v)));
            optr += CV;
        } else {
#pragma unroll
            for (int v = 0; v < VPT; ++v) best[v] = sample_v(L, Rt, xl, v);
        }
    };
    int sx = 0;
    while (sx < S) {
        int c = ax.cell[sx];
        if (c < 0) {  // crop column outside the map
            emit_zero(sx);
            ++sx;
            continue;
        }
        load_col(A, c);
        load_col(B, c + 1);
        for (;;) {
            do {  // left = A, right = B
                emit(sx, A, B);
                ++sx;
            } while (sx < S && ax.cell[sx] == c);
            if (sx >= S || ax.cell[sx] != c + 1) break;
            ++c;
            load_col(A, c + 1);
            do {  // left = B, right = A
                emit(sx, B, A);
                ++sx;
            } while (sx < S && ax.cell[sx] == c);
            if (sx >= S || ax.cell[sx] != c + 1) break;
            ++c;
            load_col(B, c + 1);
        }
    }
}

// grid (R, B); T threads; each thread owns VPT float4 channel vectors per pass over the channels.
template <int VPT, int T>
__global__ void __launch_bounds__(T, 512 / T)
    roi_pool_fwd_k2_kernel(const float4* __restrict__ feat, int Hf, int Wf, int CV, const float4* __restrict__ rois, int R,
                           int P, float4* __restrict__ out) {
    __shared__ AxisTab ty, tx;
    __shared__ AxisY ay;
    __shared__ AxisX ax;
    const int r = blockIdx.x, b = blockIdx.y;
    const int S = 2 * P;
    const float4 roi = rois[(size_t)b * R + r];
    if (threadIdx.x < S) {
        const int i = threadIdx.x;
        axis_sample(roi.y, roi.w, Hf, S, i, ty);
        ay.lo[i] = ty.lo[i] * Wf * CV;
        ay.hi[i] = ty.hi[i] * Wf * CV;
        ay.lerp[i] = ty.lerp[i];
        ay.valid[i] = ty.valid[i];
    } else if (threadIdx.x >= 32 && threadIdx.x < 32 + S) {
        const int i = threadIdx.x - 32;
        axis_sample(roi.x, roi.z, Wf, S, i, tx);
        ax.cell[i] = tx.valid[i] ? tx.lo[i] : -1;
        ax.lerp[i] = tx.lerp[i];
    }
    __syncthreads();
    const float4* fb = feat + (size_t)b * Hf * Wf * CV;
    float4* ob = out + ((size_t)b * R + r) * P * P * CV;
    for (int cv = threadIdx.x; cv < CV; cv += T * VPT) {  // host guarantees CV % (T*VPT) == 0 when VPT > 1
        const float4* f = fb + cv;
        for (int py = 0; py < P; ++py) {
            const int sy0 = 2 * py, sy1 = sy0 + 1;
            const int t0 = ay.lo[sy0], b0 = ay.hi[sy0], t1 = ay.lo[sy1], b1 = ay.hi[sy1];
            const float yl0 = ay.lerp[sy0], yl1 = ay.lerp[sy1];
            const bool yv0 = ay.valid[sy0] != 0, yv1 = ay.valid[sy1] != 0;
            float4* orow = ob + (size_t)py * P * CV + cv;
            if (yv0 && yv1 && t1 == t0 && b1 == b0) {
                const int roff[4] = {t0, b0, 0, 0};
                walk_bin_row<2, false, VPT, T>(f, roff, yl0, yl1, true, true, ax, S, Wf - 1, CV, orow);
            } else if (yv0 && yv1 && t1 == b0) {
                const int roff[4] = {t0, b0, b1, 0};
                walk_bin_row<3, false, VPT, T>(f, roff, yl0, yl1, true, true, ax, S, Wf - 1, CV, orow);
            } else {
                // rare / large-RoI shape: one channel vector at a time keeps the register budget of the kernel at
                // what the common 2- and 3-row shapes need (occupancy is set by the worst path)
                const int roff[4] = {t0, b0, t1, b1};
#pragma unroll 1
                for (int v = 0; v < VPT; ++v)
                    walk_bin_row<4, true, 1, T>(f + v * T, roff, yl0, yl1, yv0, yv1, ax, S, Wf - 1, CV, orow + v * T);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Backward.  Each thread re-derives the arg-max sample of its bins from the feature map (no saved
// state), then scatters the four bilinear weights with 16-byte vector reductions (red.global.add.v4.f32).
// Channels whose arg-max is a different sample of the window contribute an exact 0 to that reduction.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add(float4* p, const float4& v) { red_add_f4(p, v); }
__device__ __forceinline__ void red_add(float* p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ float4 vscale(const float4& a, float w) {
    return make_float4(fmul(w, a.x), fmul(w, a.y), fmul(w, a.z), fmul(w, a.w));
}
__device__ __forceinline__ float vscale(float a, float w) { return fmul(w, a); }

template <typename V>
__global__ void __launch_bounds__(kRoiThreads)
    roi_pool_bwd_kernel(const V* __restrict__ feat, int Hf, int Wf, int CV, const float4* __restrict__ rois, int R, int P,
                        int k, const V* __restrict__ dy_in, V* __restrict__ dfeat) {
    __shared__ AxisTab ty, tx;
    const int r = blockIdx.x, b = blockIdx.y;
    const int S = P * k;
    const float4 roi = rois[(size_t)b * R + r];
    if (threadIdx.x < S) axis_sample(roi.y, roi.w, Hf, S, threadIdx.x, ty);
    else if (threadIdx.x >= 32 && threadIdx.x < 32 + S) axis_sample(roi.x, roi.z, Wf, S, threadIdx.x - 32, tx);
    __syncthreads();
    const V* fb = feat + (size_t)b * Hf * Wf * CV;
    V* gb = dfeat + (size_t)b * Hf * Wf * CV;
    const V* gy = dy_in + ((size_t)b * R + r) * P * P * CV;
    constexpr int kLanes = sizeof(V) / 4;
    for (int cv = threadIdx.x; cv < CV; cv += blockDim.x) {
        for (int py = 0; py < P; ++py)
            for (int px = 0; px < P; ++px) {
                float best[kLanes];
                int arg[kLanes];
                for (int dy = 0; dy < k; ++dy)
                    for (int dx = 0; dx < k; ++dx) {
                        const int sy = py * k + dy, sx = px * k + dx, s = dy * k + dx;
                        V v = vzero(V());
                        if (ty.valid[sy] && tx.valid[sx])
                            v = bilinear(fb, Wf, CV, ty.lo[sy], ty.hi[sy], ty.lerp[sy], tx.lo[sx], tx.hi[sx],
                                         tx.lerp[sx], cv);
                        const float* vf = reinterpret_cast<const float*>(&v);
#pragma unroll
                        for (int l = 0; l < kLanes; ++l) {
                            if (s == 0) {
                                best[l] = vf[l];
                                arg[l] = 0;
                            } else {
                                max_first(best[l], vf[l], arg[l], s);
                            }
                        }
                    }
                V g;
                if constexpr (kLanes == 4) g = ld_nc_f4(gy + ((size_t)py * P + px) * CV + cv);
                else g = __ldg(gy + ((size_t)py * P + px) * CV + cv);
                const float* gf = reinterpret_cast<const float*>(&g);
                for (int dy = 0; dy < k; ++dy)
                    for (int dx = 0; dx < k; ++dx) {
                        const int sy = py * k + dy, sx = px * k + dx, s = dy * k + dx;
                        if (!(ty.valid[sy] && tx.valid[sx])) continue;  // out-of-range samples contribute nothing
                        V gs;
                        float* gsf = reinterpret_cast<float*>(&gs);
                        bool any = false;
#pragma unroll
                        for (int l = 0; l < kLanes; ++l) {
                            const bool sel = arg[l] == s;
                            gsf[l] = sel ? gf[l] : 0.0f;
                            any |= sel;
                        }
                        if (!any) continue;
                        const float yl = ty.lerp[sy], xl = tx.lerp[sx];
                        const V dtop = vscale(gs, fsub(1.0f, yl));
                        V* top = gb + ((size_t)ty.lo[sy] * Wf) * CV + cv;
                        red_add(top + (size_t)tx.lo[sx] * CV, vscale(dtop, fsub(1.0f, xl)));
                        if (xl != 0.0f) red_add(top + (size_t)tx.hi[sx] * CV, vscale(dtop, xl));
                        if (yl != 0.0f) {
                            const V dbot = vscale(gs, yl);
                            V* bot = gb + ((size_t)ty.hi[sy] * Wf) * CV + cv;
                            red_add(bot + (size_t)tx.lo[sx] * CV, vscale(dbot, fsub(1.0f, xl)));
                            if (xl != 0.0f) red_add(bot + (size_t)tx.hi[sx] * CV, vscale(dbot, xl));
                        }
                    }
            }
    }
}

// ------------------------------------------------------------------------------------------------
// Fast backward path for kernel_size == 2, float4 channels, pooled_size <= 8.  Per bin-row, two walks:
//   pass 1  the forward walk again, but instead of storing the maximum it records WHICH of the four crop
//           samples won, per channel (first maximum in row-major window order, strict >: trap T8), as 2-bit
//           codes in one 64-bit register per channel vector (7 bins x 4 channels x 2 bits);
//   pass 2  the same cells once more with the incoming gradient: every sample adds its four bilinear shares
//           (weights (1-xl)(1-yl) etc., one packed FMA each) into REGISTER accumulators for the
//           left / right feature column of the current cell; a column is flushed with one 16-byte vector
//           reduction per row (red.global.add.v4.f32) only when the walk leaves it.  That aggregates all
//           samples of a bin-row that share a feature pixel into a single atomic, instead of up to 16 per bin.
// ------------------------------------------------------------------------------------------------
template <int NR, bool kCheckY, int VPT, int T>
__device__ __forceinline__ void argmax_bin_row(const float4* __restrict__ f, const int (&roff)[4], float yl0, float yl1,
                                               bool yv0, bool yv1, const AxisX& ax, int S, int last_col, int CV,
                                               unsigned long long (&codes)[VPT], float4* __restrict__ orow = nullptr) {
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* fr[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) fr[k] = f + roff[k];
    float4 A[NR][VPT], B[NR][VPT];
    float4 e0[VPT], e1[VPT];  // the two even-column samples of the current bin (window positions (0,0) and (1,0))
#pragma unroll
    for (int v = 0; v < VPT; ++v) codes[v] = 0ull;
    auto load_col = [&](float4 (&dst)[NR][VPT], int col) {
        const int off = min(col, last_col) * CV;
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            const float4* p = fr[k] + off;
#pragma unroll
            for (int v = 0; v < VPT; ++v) dst[k][v] = __ldg(p + v * T);
        }
    };
    auto pick = [&](float a00, float a01, float a10, float a11) -> unsigned {
        float best = a00;
        unsigned arg = 0u;
        if (a01 > best) { best = a01; arg = 1u; }
        if (a10 > best) { best = a10; arg = 2u; }
        if (a11 > best) { arg = 3u; }
        return arg;
    };
    auto finish = [&](int sx, int v, const float4& o0, const float4& o1) {
        const unsigned cb = pick(e0[v].x, o0.x, e1[v].x, o1.x) | (pick(e0[v].y, o0.y, e1[v].y, o1.y) << 2) |
                            (pick(e0[v].z, o0.z, e1[v].z, o1.z) << 4) | (pick(e0[v].w, o0.w, e1[v].w, o1.w) << 6);
        codes[v] |= (unsigned long long)cb << ((sx >> 1) * 8);
        if (orow)  // forward-with-argmax: also the pooled value (same max as roi_pool_fwd_k2_kernel)
            st_cs_f4(orow + (size_t)(sx >> 1) * CV + v * T, vmax4(vmax4(e0[v], o0), vmax4(e1[v], o1)));
    };
    auto emit_zero = [&](int sx) {
#pragma unroll
        for (int v = 0; v < VPT; ++v) {
            if (sx & 1) finish(sx, v, zero, zero);
            else e0[v] = e1[v] = zero;
        }
    };
    auto emit = [&](int sx, const float4 (&L)[NR][VPT], const float4 (&Rt)[NR][VPT]) {
        const float xl = ax.lerp[sx];
#pragma unroll
        for (int v = 0; v < VPT; ++v) {
            float4 X[NR];
#pragma unroll
            for (int k = 0; k < NR; ++k) X[k] = lerp_packed(L[k][v], Rt[k][v], xl);
            float4 s0 = lerp_packed(X[0], X[1], yl0);
            float4 s1 = NR == 2 ? lerp_packed(X[0], X[1], yl1)
                                : NR == 3 ? lerp_packed(X[1], X[2], yl1) : lerp_packed(X[NR - 2], X[NR - 1], yl1);
            if (kCheckY) {
                if (!yv0) s0 = zero;
                if (!yv1) s1 = zero;
            }
            if (sx & 1) finish(sx, v, s0, s1);
            else { e0[v] = s0; e1[v] = s1; }
        }
    };
    int sx = 0;
    while (sx < S) {
        int c = ax.cell[sx];
        if (c < 0) {
            emit_zero(sx);
            ++sx;
            continue;
        }
        load_col(A, c);
        load_col(B, c + 1);
        for (;;) {
            do {
                emit(sx, A, B);
                ++sx;
            } while (sx < S && ax.cell[sx] == c);
            if (sx >= S || ax.cell[sx] != c + 1) break;
            ++c;
            load_col(A, c + 1);
            do {
                emit(sx, B, A);
                ++sx;
            } while (sx < S && ax.cell[sx] == c);
            if (sx >= S || ax.cell[sx] != c + 1) break;
            ++c;
            load_col(B, c + 1);
        }
    }
}

// acc + w * x with packed fused multiply-adds (FFMA2).  The gradient has no reference summation order (TF's GPU
// kernel uses atomics, its CPU kernel a sequential loop) and its bar is 1e-5 relative, so the backward may fuse:
// one rounding per contribution instead of TF's three ((1-yl)*g, (1-xl)*dtop, +=) -- closer to the exact sum.
__device__ __forceinline__ void axpy_packed(float4& acc, float w, const float4& x) {
    unsigned long long a0, a1, x0, x1, ww;
    asm("mov.b64 %0, {%1,%2};" : "=l"(a0) : "f"(acc.x), "f"(acc.y));
    asm("mov.b64 %0, {%1,%2};" : "=l"(a1) : "f"(acc.z), "f"(acc.w));
    asm("mov.b64 %0, {%1,%2};" : "=l"(x0) : "f"(x.x), "f"(x.y));
    asm("mov.b64 %0, {%1,%2};" : "=l"(x1) : "f"(x.z), "f"(x.w));
    asm("mov.b64 %0, {%1,%1};" : "=l"(ww) : "f"(w));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a0) : "l"(x0), "l"(ww));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a1) : "l"(x1), "l"(ww));
    asm("mov.b64 {%0,%1}, %2;" : "=f"(acc.x), "=f"(acc.y) : "l"(a0));
    asm("mov.b64 {%0,%1}, %2;" : "=f"(acc.z), "=f"(acc.w) : "l"(a1));
}

template <int NR, bool kCheckY, int VPT, int T>
__device__ __forceinline__ void scatter_bin_row(float4* __restrict__ g_base, const int (&roff)[4], float yl0, float yl1,
                                                bool yv0, bool yv1, const AxisX& ax, int S, int last_col, int CV,
                                                const float4* __restrict__ gy_row,
                                                const unsigned long long (&codes)[VPT]) {
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    float4* gr[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) gr[k] = g_base + roff[k];
    float4 A[NR][VPT], B[NR][VPT];  // gradient accumulated for feature columns c (left) and c+1 (right)
    float4 g[VPT];
    int g_bin = -1;
    auto clear = [&](float4 (&acc)[NR][VPT]) {
#pragma unroll
        for (int k = 0; k < NR; ++k)
#pragma unroll
            for (int v = 0; v < VPT; ++v) acc[k][v] = zero;
    };
    auto flush = [&](const float4 (&acc)[NR][VPT], int col) {
        const int off = min(col, last_col) * CV;
#pragma unroll
        for (int k = 0; k < NR; ++k)
#pragma unroll
            for (int v = 0; v < VPT; ++v) red_add_f4(gr[k] + off + v * T, acc[k][v]);
    };
    // rows under crop row 1: NR==2 -> (0,1), NR==3 -> (1,2), NR==4 -> (2,3); crop row 0 always (0,1)
    constexpr int kT1 = NR == 2 ? 0 : NR == 3 ? 1 : 2;
    constexpr int kB1 = kT1 + 1;
    auto accumulate = [&](int sx, float4 (&L)[NR][VPT], float4 (&Rt)[NR][VPT]) {
        const int px = sx >> 1;
        const unsigned dx = (unsigned)(sx & 1);
        if (g_bin != px) {
#pragma unroll
            for (int v = 0; v < VPT; ++v) g[v] = ld_nc_f4(gy_row + (size_t)px * CV + v * T);
            g_bin = px;
        }
        const float xl = ax.lerp[sx];
        const float wl = fsub(1.0f, xl);
#pragma unroll
        for (int v = 0; v < VPT; ++v) {
            const unsigned cb = (unsigned)(codes[v] >> (px * 8)) & 0xffu;
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
                if (kCheckY && !(dy == 0 ? yv0 : yv1)) continue;  // crop row outside the map contributes nothing
                const unsigned id = (unsigned)dy * 2u + dx;
                const float4 gs = make_float4(((cb >> 0) & 3u) == id ? g[v].x : 0.f, ((cb >> 2) & 3u) == id ? g[v].y : 0.f,
                                              ((cb >> 4) & 3u) == id ? g[v].z : 0.f, ((cb >> 6) & 3u) == id ? g[v].w : 0.f);
                const float yl = dy == 0 ? yl0 : yl1;
                const float wt = fsub(1.0f, yl);
                const int kt = dy == 0 ? 0 : kT1, kb = dy == 0 ? 1 : kB1;
                axpy_packed(L[kt][v], fmul(wl, wt), gs);   // (1-xl)(1-yl) g -> top-left
                axpy_packed(Rt[kt][v], fmul(xl, wt), gs);  //    xl (1-yl) g -> top-right
                axpy_packed(L[kb][v], fmul(wl, yl), gs);   // (1-xl)   yl  g -> bottom-left
                axpy_packed(Rt[kb][v], fmul(xl, yl), gs);  //    xl    yl  g -> bottom-right
            }
        }
    };
    int sx = 0;
    while (sx < S) {
        int c = ax.cell[sx];
        if (c < 0) {  // crop column outside the map: no gradient
            ++sx;
            continue;
        }
        clear(A);
        clear(B);
        for (;;) {
            do {  // left = A (column c), right = B (column c+1)
                accumulate(sx, A, B);
                ++sx;
            } while (sx < S && ax.cell[sx] == c);
            if (sx >= S || ax.cell[sx] != c + 1) {
                flush(A, c);
                flush(B, c + 1);
                break;
            }
            flush(A, c);
            clear(A);
            ++c;
            do {  // left = B (column c), right = A (column c+1)
                accumulate(sx, B, A);
                ++sx;
            } while (sx < S && ax.cell[sx] == c);
            if (sx >= S || ax.cell[sx] != c + 1) {
                flush(B, c);
                flush(A, c + 1);
                break;
            }
            flush(B, c);
            clear(B);
            ++c;
        }
    }
}

template <int VPT, int T>
__global__ void __launch_bounds__(T, VPT == 2 ? 3 : 512 / T)
    roi_pool_bwd_k2_kernel(const float4* __restrict__ feat, int Hf, int Wf, int CV, const float4* __restrict__ rois, int R,
                           int P, const float4* __restrict__ dy_in, float4* __restrict__ dfeat) {
    __shared__ AxisTab ty, tx;
    __shared__ AxisY ay;
    __shared__ AxisX ax;
    const int r = blockIdx.x, b = blockIdx.y;
    const int S = 2 * P;
    const float4 roi = rois[(size_t)b * R + r];
    if (threadIdx.x < S) {
        const int i = threadIdx.x;
        axis_sample(roi.y, roi.w, Hf, S, i, ty);
        ay.lo[i] = ty.lo[i] * Wf * CV;
        ay.hi[i] = ty.hi[i] * Wf * CV;
        ay.lerp[i] = ty.lerp[i];
        ay.valid[i] = ty.valid[i];
    } else if (threadIdx.x >= 32 && threadIdx.x < 32 + S) {
        const int i = threadIdx.x - 32;
        axis_sample(roi.x, roi.z, Wf, S, i, tx);
        ax.cell[i] = tx.valid[i] ? tx.lo[i] : -1;
        ax.lerp[i] = tx.lerp[i];
    }
    __syncthreads();
    const float4* fb = feat + (size_t)b * Hf * Wf * CV;
    float4* gb = dfeat + (size_t)b * Hf * Wf * CV;
    const float4* gy = dy_in + ((size_t)b * R + r) * P * P * CV;
    for (int cv = threadIdx.x; cv < CV; cv += T * VPT) {
        for (int py = 0; py < P; ++py) {
            const int sy0 = 2 * py, sy1 = sy0 + 1;
            const int t0 = ay.lo[sy0], b0 = ay.hi[sy0], t1 = ay.lo[sy1], b1 = ay.hi[sy1];
            const float yl0 = ay.lerp[sy0], yl1 = ay.lerp[sy1];
            const bool yv0 = ay.valid[sy0] != 0, yv1 = ay.valid[sy1] != 0;
            const float4* gy_row = gy + (size_t)py * P * CV + cv;
            if (yv0 && yv1 && t1 == t0 && b1 == b0) {
                const int roff[4] = {t0, b0, 0, 0};
                unsigned long long codes[VPT];
                argmax_bin_row<2, false, VPT, T>(fb + cv, roff, yl0, yl1, true, true, ax, S, Wf - 1, CV, codes);
                scatter_bin_row<2, false, VPT, T>(gb + cv, roff, yl0, yl1, true, true, ax, S, Wf - 1, CV, gy_row, codes);
            } else if (yv0 && yv1 && t1 == b0) {
                const int roff[4] = {t0, b0, b1, 0};
                unsigned long long codes[VPT];
                argmax_bin_row<3, false, VPT, T>(fb + cv, roff, yl0, yl1, true, true, ax, S, Wf - 1, CV, codes);
                scatter_bin_row<3, false, VPT, T>(gb + cv, roff, yl0, yl1, true, true, ax, S, Wf - 1, CV, gy_row, codes);
            } else {
                const int roff[4] = {t0, b0, t1, b1};
#pragma unroll 1
                for (int v = 0; v < VPT; ++v) {
                    unsigned long long codes[1];
                    argmax_bin_row<4, true, 1, T>(fb + cv + v * T, roff, yl0, yl1, yv0, yv1, ax, S, Wf - 1, CV, codes);
                    scatter_bin_row<4, true, 1, T>(gb + cv + v * T, roff, yl0, yl1, yv0, yv1, ax, S, Wf - 1, CV,
                                                   gy_row + v * T, codes);
                }
            }
        }
    }
}

// Training-mode forward: pooled output plus the arg-max codes ([B,R,P,P,CV] bytes: four 2-bit codes per float4),
// so that the backward pass needs neither the feature maps nor a second interpolation pass.
// MODE 0: forward + codes.  MODE 1: backward from saved codes (pass 2 of roi_pool_bwd_k2_kernel only).
template <int VPT, int T, int MODE>
__global__ void __launch_bounds__(T, VPT == 2 ? 3 : 512 / T)
    roi_pool_codes_kernel(const float4* __restrict__ feat, int Hf, int Wf, int CV, const float4* __restrict__ rois, int R,
                          int P, float4* __restrict__ out, unsigned char* __restrict__ codes_io,
                          const float4* __restrict__ dy_in, float4* __restrict__ dfeat) {
    __shared__ AxisTab ty, tx;
    __shared__ AxisY ay;
    __shared__ AxisX ax;
    const int r = blockIdx.x, b = blockIdx.y;
    const int S = 2 * P;
    const float4 roi = rois[(size_t)b * R + r];
    if (threadIdx.x < S) {
        const int i = threadIdx.x;
        axis_sample(roi.y, roi.w, Hf, S, i, ty);
        ay.lo[i] = ty.lo[i] * Wf * CV;
        ay.hi[i] = ty.hi[i] * Wf * CV;
        ay.lerp[i] = ty.lerp[i];
        ay.valid[i] = ty.valid[i];
    } else if (threadIdx.x >= 32 && threadIdx.x < 32 + S) {
        const int i = threadIdx.x - 32;
        axis_sample(roi.x, roi.z, Wf, S, i, tx);
        ax.cell[i] = tx.valid[i] ? tx.lo[i] : -1;
        ax.lerp[i] = tx.lerp[i];
    }
    __syncthreads();
    const size_t bin0 = ((size_t)b * R + r) * P * P;
    for (int cv = threadIdx.x; cv < CV; cv += T * VPT) {
        for (int py = 0; py < P; ++py) {
            const int sy0 = 2 * py, sy1 = sy0 + 1;
            const int t0 = ay.lo[sy0], b0 = ay.hi[sy0], t1 = ay.lo[sy1], b1 = ay.hi[sy1];
            const float yl0 = ay.lerp[sy0], yl1 = ay.lerp[sy1];
            const bool yv0 = ay.valid[sy0] != 0, yv1 = ay.valid[sy1] != 0;
            unsigned char* crow = codes_io + (bin0 + (size_t)py * P) * CV + cv;  // + px*CV + v*T
            const int shape = (yv0 && yv1 && t1 == t0 && b1 == b0) ? 2 : (yv0 && yv1 && t1 == b0) ? 3 : 4;
            if (MODE == 0) {
                const float4* f = feat + (size_t)b * Hf * Wf * CV + cv;
                float4* orow = out + (bin0 + (size_t)py * P) * CV + cv;
                unsigned long long codes[VPT];
                if (shape == 2) {
                    const int roff[4] = {t0, b0, 0, 0};
                    argmax_bin_row<2, false, VPT, T>(f, roff, yl0, yl1, true, true, ax, S, Wf - 1, CV, codes, orow);
                } else if (shape == 3) {
                    const int roff[4] = {t0, b0, b1, 0};
                    argmax_bin_row<3, false, VPT, T>(f, roff, yl0, yl1, true, true, ax, S, Wf - 1, CV, codes, orow);
                } else {
                    const int roff[4] = {t0, b0, t1, b1};
#pragma unroll 1
                    for (int v = 0; v < VPT; ++v) {
                        unsigned long long c1[1];
                        argmax_bin_row<4, true, 1, T>(f + v * T, roff, yl0, yl1, yv0, yv1, ax, S, Wf - 1, CV, c1, orow + v * T);
                        codes[v] = c1[0];
                    }
                }
#pragma unroll
                for (int v = 0; v < VPT; ++v)
                    for (int px = 0; px < P; ++px) crow[(size_t)px * CV + v * T] = (unsigned char)(codes[v] >> (px * 8));
            } else {
                float4* g = dfeat + (size_t)b * Hf * Wf * CV + cv;
                const float4* gy_row = dy_in + (bin0 + (size_t)py * P) * CV + cv;
                unsigned long long codes[VPT];
#pragma unroll
                for (int v = 0; v < VPT; ++v) {
                    codes[v] = 0ull;
                    for (int px = 0; px < P; ++px)
                        codes[v] |= (unsigned long long)__ldg(crow + (size_t)px * CV + v * T) << (px * 8);
                }
                if (shape == 2) {
                    const int roff[4] = {t0, b0, 0, 0};
                    scatter_bin_row<2, false, VPT, T>(g, roff, yl0, yl1, true, true, ax, S, Wf - 1, CV, gy_row, codes);
                } else if (shape == 3) {
                    const int roff[4] = {t0, b0, b1, 0};
                    scatter_bin_row<3, false, VPT, T>(g, roff, yl0, yl1, true, true, ax, S, Wf - 1, CV, gy_row, codes);
                } else {
                    const int roff[4] = {t0, b0, t1, b1};
#pragma unroll 1
                    for (int v = 0; v < VPT; ++v) {
                        const unsigned long long c1[1] = {codes[v]};
                        scatter_bin_row<4, true, 1, T>(g + v * T, roff, yl0, yl1, yv0, yv1, ax, S, Wf - 1, CV, gy_row + v * T, c1);
                    }
                }
            }
        }
    }
}

template <int MODE>
static int launch_codes(const float* feat, int B, int Hf, int Wf, int C, const float* rois, int R, int P, float* out,
                        unsigned char* codes, const float* dy, float* dfeat, cudaStream_t st) {
    const dim3 grid(R, B);
    const int cv = C / 4;
    if (cv % 256 == 0)
        roi_pool_codes_kernel<2, 128, MODE><<<grid, 128, 0, st>>>((const float4*)feat, Hf, Wf, cv, (const float4*)rois, R, P,
                                                                  (float4*)out, codes, (const float4*)dy, (float4*)dfeat);
    else if (cv > 64)
        roi_pool_codes_kernel<1, 128, MODE><<<grid, 128, 0, st>>>((const float4*)feat, Hf, Wf, cv, (const float4*)rois, R, P,
                                                                  (float4*)out, codes, (const float4*)dy, (float4*)dfeat);
    else
        roi_pool_codes_kernel<1, 64, MODE><<<grid, 64, 0, st>>>((const float4*)feat, Hf, Wf, cv, (const float4*)rois, R, P,
                                                                (float4*)out, codes, (const float4*)dy, (float4*)dfeat);
    FRCNN_LAUNCH_CHECK(MODE == 0 ? "roi_pool_codes_kernel<fwd>" : "roi_pool_codes_kernel<bwd>");
    return 0;
}

}  // namespace frcnn

using namespace frcnn;

extern "C" {

int frcnn_roi_pool_set_max_sms(int max_sms) {
    roi_pool_set_max_sms(max_sms);
    return 0;
}

size_t frcnn_roi_pool_workspace_bytes(int B, int R, int C, int P, int k) {
    if (B <= 0 || R <= 0 || !roi_pool_stream_supported(C, P, k) || kForceBinRowWalk) return 0;
    return roi_pool_stream_workspace_bytes(B, R);
}

int frcnn_roi_pool_forward(const float* feat, int B, int Hf, int Wf, int C, const float* rois, int R, int P, int k,
                           float* out, void* workspace, size_t workspace_bytes, frcnn_stream_t stream) {
    FRCNN_REQUIRE(B > 0 && Hf > 0 && Wf > 0 && C > 0 && R >= 0 && P > 0 && k > 0, FRCNN_EINVAL,
                  "roi_pool_forward: bad shape");
    FRCNN_REQUIRE(P * k <= kMaxCrop, FRCNN_ELIMIT, "roi_pool_forward: pooled_size*kernel_size = %d > %d", P * k, kMaxCrop);
    FRCNN_REQUIRE(aligned16(feat) && aligned16(rois) && aligned16(out), FRCNN_EINVAL,
                  "roi_pool_forward: pointers must be 16-byte aligned");
    FRCNN_REQUIRE(B <= 65535, FRCNN_ELIMIT, "roi_pool_forward: B > 65535");
    if (R == 0) return 0;
    const dim3 grid(R, B);
    if (roi_pool_stream_supported(C, P, k) && !kForceBinRowWalk)
        return launch_roi_pool_fwd_persist(feat, B, Hf, Wf, C, rois, R, out, nullptr, workspace, workspace_bytes, as_stream(stream));
    if (C % 4 == 0 && k == 2 && (int64_t)Hf * Wf * (C / 4) < (1ll << 31)) {
        const int cv = C / 4;
        if (cv % 256 == 0)
            roi_pool_fwd_k2_kernel<2, 128><<<grid, 128, 0, as_stream(stream)>>>((const float4*)feat, Hf, Wf, cv,
                                                                                (const float4*)rois, R, P, (float4*)out);
        else if (cv > 64)
            roi_pool_fwd_k2_kernel<1, 128><<<grid, 128, 0, as_stream(stream)>>>((const float4*)feat, Hf, Wf, cv,
                                                                                (const float4*)rois, R, P, (float4*)out);
        else
            roi_pool_fwd_k2_kernel<1, 64><<<grid, 64, 0, as_stream(stream)>>>((const float4*)feat, Hf, Wf, cv,
                                                                              (const float4*)rois, R, P, (float4*)out);
    } else if (C % 4 == 0)
        roi_pool_fwd_kernel<float4><<<grid, kRoiThreads, 0, as_stream(stream)>>>((const float4*)feat, Hf, Wf, C / 4,
                                                                                  (const float4*)rois, R, P, k,
                                                                                  (float4*)out);
    else
        roi_pool_fwd_kernel<float><<<grid, kRoiThreads, 0, as_stream(stream)>>>(feat, Hf, Wf, C, (const float4*)rois, R,
                                                                                 P, k, out);
    FRCNN_LAUNCH_CHECK("roi_pool_fwd_kernel");
    return 0;
}

int frcnn_roi_pool_backward(const float* feat, int B, int Hf, int Wf, int C, const float* rois, int R, int P, int k,
                            const float* dy, float* dfeat, frcnn_stream_t stream) {
    FRCNN_REQUIRE(B > 0 && Hf > 0 && Wf > 0 && C > 0 && R >= 0 && P > 0 && k > 0, FRCNN_EINVAL,
                  "roi_pool_backward: bad shape");
    FRCNN_REQUIRE(P * k <= kMaxCrop, FRCNN_ELIMIT, "roi_pool_backward: pooled_size*kernel_size = %d > %d", P * k, kMaxCrop);
    FRCNN_REQUIRE(aligned16(feat) && aligned16(rois) && aligned16(dy) && aligned16(dfeat), FRCNN_EINVAL,
                  "roi_pool_backward: pointers must be 16-byte aligned");
    FRCNN_REQUIRE(B <= 65535, FRCNN_ELIMIT, "roi_pool_backward: B > 65535");
    FRCNN_CUDA(cudaMemsetAsync(dfeat, 0, sizeof(float) * (size_t)B * Hf * Wf * C, as_stream(stream)));
    if (R == 0) return 0;
    const dim3 grid(R, B);
    if (C % 4 == 0 && k == 2 && P <= 8 && (int64_t)Hf * Wf * (C / 4) < (1ll << 31)) {
        const int cv = C / 4;
        if (cv % 256 == 0)
            roi_pool_bwd_k2_kernel<2, 128><<<grid, 128, 0, as_stream(stream)>>>(
                (const float4*)feat, Hf, Wf, cv, (const float4*)rois, R, P, (const float4*)dy, (float4*)dfeat);
        else if (cv > 64)
            roi_pool_bwd_k2_kernel<1, 128><<<grid, 128, 0, as_stream(stream)>>>(
                (const float4*)feat, Hf, Wf, cv, (const float4*)rois, R, P, (const float4*)dy, (float4*)dfeat);
        else
            roi_pool_bwd_k2_kernel<1, 64><<<grid, 64, 0, as_stream(stream)>>>(
                (const float4*)feat, Hf, Wf, cv, (const float4*)rois, R, P, (const float4*)dy, (float4*)dfeat);
    } else if (C % 4 == 0)
        roi_pool_bwd_kernel<float4><<<grid, kRoiThreads, 0, as_stream(stream)>>>(
            (const float4*)feat, Hf, Wf, C / 4, (const float4*)rois, R, P, k, (const float4*)dy, (float4*)dfeat);
    else
        roi_pool_bwd_kernel<float><<<grid, kRoiThreads, 0, as_stream(stream)>>>(feat, Hf, Wf, C, (const float4*)rois, R,
                                                                                 P, k, dy, dfeat);
    FRCNN_LAUNCH_CHECK("roi_pool_bwd_kernel");
    return 0;
}

int frcnn_roi_pool_prepare(const float* rois, int B, int R, int Hf, int Wf, int C, int P, int k, void* workspace,
                           size_t workspace_bytes, frcnn_stream_t stream) {
    FRCNN_REQUIRE(B > 0 && Hf > 0 && Wf > 0 && C > 0 && R >= 0 && B <= 65535, FRCNN_EINVAL, "roi_pool_prepare: bad shape");
    FRCNN_REQUIRE(roi_pool_stream_supported(C, P, k) && !kForceBinRowWalk, FRCNN_EINVAL,
                  "roi_pool_prepare: only the row-streaming configuration (pooled_size 7, kernel_size 2, C %% 512 == 0) has records");
    FRCNN_REQUIRE(aligned16(rois), FRCNN_EINVAL, "roi_pool_prepare: rois must be 16-byte aligned");
    if (R == 0) return 0;
    return launch_roi_records(B, Hf, Wf, rois, R, workspace, workspace_bytes, as_stream(stream));
}

int frcnn_roi_pool_forward_prepared(const float* feat, int B, int Hf, int Wf, int C, int R, int P, int k, float* out,
                                    unsigned char* argmax_codes, void* workspace, size_t workspace_bytes,
                                    frcnn_stream_t stream) {
    FRCNN_REQUIRE(B > 0 && Hf > 0 && Wf > 0 && C > 0 && R >= 0 && B <= 65535, FRCNN_EINVAL, "roi_pool_forward_prepared: bad shape");
    FRCNN_REQUIRE(roi_pool_stream_supported(C, P, k) && !kForceBinRowWalk, FRCNN_EINVAL,
                  "roi_pool_forward_prepared: only the row-streaming configuration has records");
    FRCNN_REQUIRE(aligned16(feat) && aligned16(out), FRCNN_EINVAL, "roi_pool_forward_prepared: pointers must be 16-byte aligned");
    if (R == 0) return 0;
    return launch_roi_pool_fwd_persist(feat, B, Hf, Wf, C, nullptr, R, out, argmax_codes, workspace, workspace_bytes,
                                       as_stream(stream), true);
}

int frcnn_roi_pool_backward_prepared(int B, int Hf, int Wf, int C, int R, int P, int k, const float* dy,
                                     const unsigned char* argmax_codes, float* dfeat, void* workspace, size_t workspace_bytes,
                                     frcnn_stream_t stream) {
    FRCNN_REQUIRE(B > 0 && Hf > 0 && Wf > 0 && C > 0 && R >= 0 && B <= 65535, FRCNN_EINVAL, "roi_pool_backward_prepared: bad shape");
    FRCNN_REQUIRE(roi_pool_stream_supported(C, P, k) && !kForceBinRowWalk, FRCNN_EINVAL,
                  "roi_pool_backward_prepared: only the row-streaming configuration has records");
    FRCNN_REQUIRE(aligned16(dy) && aligned16(dfeat) && argmax_codes, FRCNN_EINVAL,
                  "roi_pool_backward_prepared: pointers must be 16-byte aligned");
    FRCNN_CUDA(cudaMemsetAsync(dfeat, 0, sizeof(float) * (size_t)B * Hf * Wf * C, as_stream(stream)));
    if (R == 0) return 0;
    FRCNN_REQUIRE(workspace && workspace_bytes >= 256, FRCNN_EWORKSPACE, "roi_pool_backward_prepared: workspace too small");
    // the backward's work counter (second word of the workspace) is re-armed here, next to the zero-init of dfeature, so
    // the same records serve any number of backward calls
    FRCNN_CUDA(cudaMemsetAsync(static_cast<unsigned*>(workspace) + 1, 0, sizeof(unsigned), as_stream(stream)));
    return launch_roi_pool_bwd_persist(B, Hf, Wf, C, nullptr, R, dy, argmax_codes, dfeat, workspace, workspace_bytes,
                                       as_stream(stream), true);
}

int frcnn_roi_pool_argmax_supported(int C, int pooled_size, int kernel_size) {
    return (C % 4 == 0 && kernel_size == 2 && pooled_size >= 1 && pooled_size <= 8) ? 1 : 0;
}

int frcnn_roi_pool_forward_argmax(const float* feat, int B, int Hf, int Wf, int C, const float* rois, int R, int P, int k,
                                  float* out, unsigned char* argmax_codes, void* workspace, size_t workspace_bytes,
                                  frcnn_stream_t stream) {
    FRCNN_REQUIRE(B > 0 && Hf > 0 && Wf > 0 && C > 0 && R >= 0 && B <= 65535, FRCNN_EINVAL, "roi_pool_forward_argmax: bad shape");
    FRCNN_REQUIRE(frcnn_roi_pool_argmax_supported(C, P, k) && (int64_t)Hf * Wf * (C / 4) < (1ll << 31), FRCNN_ELIMIT,
                  "roi_pool_forward_argmax: needs C %% 4 == 0, kernel_size == 2, pooled_size <= 8");
    FRCNN_REQUIRE(aligned16(feat) && aligned16(rois) && aligned16(out) && argmax_codes, FRCNN_EINVAL,
                  "roi_pool_forward_argmax: pointers must be 16-byte aligned");
    if (R == 0) return 0;
    if (roi_pool_stream_supported(C, P, k) && !kForceBinRowWalk)
        return launch_roi_pool_fwd_persist(feat, B, Hf, Wf, C, rois, R, out, argmax_codes, workspace, workspace_bytes,
                                           as_stream(stream));
    return launch_codes<0>(feat, B, Hf, Wf, C, rois, R, P, out, argmax_codes, nullptr, nullptr, as_stream(stream));
}

int frcnn_roi_pool_backward_argmax(int B, int Hf, int Wf, int C, const float* rois, int R, int P, int k, const float* dy,
                                   const unsigned char* argmax_codes, float* dfeat, void* workspace, size_t workspace_bytes,
                                   frcnn_stream_t stream) {
    FRCNN_REQUIRE(B > 0 && Hf > 0 && Wf > 0 && C > 0 && R >= 0 && B <= 65535, FRCNN_EINVAL, "roi_pool_backward_argmax: bad shape");
    FRCNN_REQUIRE(frcnn_roi_pool_argmax_supported(C, P, k) && (int64_t)Hf * Wf * (C / 4) < (1ll << 31), FRCNN_ELIMIT,
                  "roi_pool_backward_argmax: needs C %% 4 == 0, kernel_size == 2, pooled_size <= 8");
    FRCNN_REQUIRE(aligned16(rois) && aligned16(dy) && aligned16(dfeat) && argmax_codes, FRCNN_EINVAL,
                  "roi_pool_backward_argmax: pointers must be 16-byte aligned");
    FRCNN_CUDA(cudaMemsetAsync(dfeat, 0, sizeof(float) * (size_t)B * Hf * Wf * C, as_stream(stream)));
    if (R == 0) return 0;
    if (roi_pool_stream_supported(C, P, k) && !kForceBinRowWalk)
        return launch_roi_pool_bwd_persist(B, Hf, Wf, C, rois, R, dy, argmax_codes, dfeat, workspace, workspace_bytes,
                                           as_stream(stream));
    return launch_codes<1>(nullptr, B, Hf, Wf, C, rois, R, P, nullptr, const_cast<unsigned char*>(argmax_codes), dy, dfeat,
                           as_stream(stream));
}

}  // extern "C"

// Shared helpers for libfrcnn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/frcnn_b200.h"

namespace frcnn {

// thread-local error text behind frcnn_last_error()
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define FRCNN_REQUIRE(cond, code, ...)        \
    do {                                      \
        if (!(cond)) {                        \
            ::frcnn::set_error(__VA_ARGS__);  \
            return (code);                    \
        }                                     \
    } while (0)

#define FRCNN_LAUNCH_CHECK(what)                                   \
    do {                                                           \
        cudaError_t e__ = cudaGetLastError();                      \
        if (e__ != cudaSuccess) return ::frcnn::cuda_fail(e__, what); \
    } while (0)

#define FRCNN_CUDA(call)                                                \
    do {                                                                \
        cudaError_t e__ = (call);                                       \
        if (e__ != cudaSuccess) return ::frcnn::cuda_fail(e__, #call);  \
    } while (0)

static inline cudaStream_t as_stream(frcnn_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Raise a kernel's dynamic shared-memory limit.  The attribute is per function AND per device, so the high-water mark is
// kept per device (a host thread may drive several GPUs); kSite gives every call site its own table.
template <int kSite>
static inline cudaError_t ensure_dynamic_smem(const void* func, size_t bytes) {
    static thread_local size_t configured[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const bool cached = dev >= 0 && dev < 64;
    if (cached && bytes <= configured[dev]) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess && cached) configured[dev] = bytes;
    return e;
}

// ---------------------------------------------------------------------------------------------
// Exactly-rounded fp32 helpers.  The library is compiled with -fmad=false, and the parity-critical
// expressions additionally use the _rn intrinsics so that no later flag change can contract them.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// Order-preserving map float -> uint32 (larger float <=> larger key).  Smallest image of a real
// number is key(-inf) = 0x007fffff, so 0 is free to mean "not a candidate".
__device__ __forceinline__ uint32_t score_key(float s) {
    uint32_t b = __float_as_uint(s);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_score(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// 128-bit streaming accesses: read-only path without L1 allocation / evict-first stores.
__device__ __forceinline__ float4 ld_nc_f4(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void st_cs_f4(float4* p, const float4& v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
// sm_90+ vector reduction: one 16-byte fp32x4 atomic add, no return value.
__device__ __forceinline__ void red_add_f4(float4* p, const float4& v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// ---- reference IoU flavours --------------------------------------------------------------------
// utils/metrics.py:136-208 (target matching): no corner normalisation, 0 when inter == 0.
__device__ __forceinline__ float ref_iou(const float4& a, const float4& b) {
    const float dw = fsub(fminf(a.z, b.z), fmaxf(a.x, b.x));
    const float dh = fsub(fminf(a.w, b.w), fmaxf(a.y, b.y));
    const float inter = fmul(fmaxf(0.0f, dw), fmaxf(0.0f, dh));
    const float area_a = fmul(fsub(a.z, a.x), fsub(a.w, a.y));
    const float area_b = fmul(fsub(b.z, b.x), fsub(b.w, b.y));
    const float uni = fsub(fadd(area_a, area_b), inter);
    return inter == 0.0f ? 0.0f : fdiv(inter, uni);
}

// utils/boxes.py:20-41 decode, :44-73 encode (one box).
__device__ __forceinline__ float4 decode_box(const float4& t, const float4& r) {
    const float cxr = fdiv(fadd(r.z, r.x), 2.0f), cyr = fdiv(fadd(r.w, r.y), 2.0f);
    const float wr = fsub(r.z, r.x), hr = fsub(r.w, r.y);
    const float cx = fadd(fmul(t.x, wr), cxr);
    const float cy = fadd(fmul(t.y, hr), cyr);
    const float hw = fmul(0.5f, fmul(expf(t.z), wr));
    const float hh = fmul(0.5f, fmul(expf(t.w), hr));
    return make_float4(fsub(cx, hw), fsub(cy, hh), fadd(cx, hw), fadd(cy, hh));
}
__device__ __forceinline__ float4 encode_box(const float4& b, const float4& r) {
    const float cxr = fdiv(fadd(r.z, r.x), 2.0f), cyr = fdiv(fadd(r.w, r.y), 2.0f);
    const float wr = fsub(r.z, r.x), hr = fsub(r.w, r.y);
    const float cx = fdiv(fadd(b.z, b.x), 2.0f), cy = fdiv(fadd(b.w, b.y), 2.0f);
    const float w = fsub(b.z, b.x), h = fsub(b.w, b.y);
    return make_float4(fdiv(fsub(cx, cxr), wr), fdiv(fsub(cy, cyr), hr), logf(fdiv(w, wr)), logf(fdiv(h, hr)));
}

}  // namespace frcnn

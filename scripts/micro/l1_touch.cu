// Micro-benchmark: does a byte "touch" (LDG.U8) make a later 128-bit load of the same sectors an L1 hit?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l1_touch l1_touch.cu && ./l1_touch
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float4 ldg_nc(const float4* p) { return __ldg(p); }
__global__ void k(const float4* __restrict__ buf, size_t stride_vec, int mode, int delay, long long* out, float* sink) {
    // one warp per CTA; each iteration uses a fresh 512-byte row (never touched before)
    const int lane = threadIdx.x;
    __shared__ float2 dummy[64];
    long long tot = 0;
    float acc = 0.f;
    for (int it = 0; it < 64; ++it) {
        const float4* p = buf + ((size_t)blockIdx.x * 64 + it) * stride_vec + lane;
        if (mode == 1) {  // byte touch
            unsigned int s;
            asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(s) : "l"(p));
            acc += (float)s * 0.f;
        } else if (mode == 2) {  // full 128-bit touch
            float4 v = ldg_nc(p);
            acc += v.x * 0.f;
        } else if (mode == 3) {  // prefetch.global.L1
            asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
        } else if (mode == 4) {  // 32-bit touch per lane
            unsigned int s;
            asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(s) : "l"(p));
            acc += (float)s * 0.f;
        } else if (mode == 5) {  // 64-bit touch per lane
            float2 v = __ldg(reinterpret_cast<const float2*>(p));
            acc += v.x * 0.f;
        } else if (mode == 6) {  // 128-bit touch by even lanes only (one per 32-byte sector)
            if ((lane & 1) == 0) { float4 v = ldg_nc(p); acc += v.x * 0.f; }
        } else if (mode == 7) {  // 128-bit touch by every 8th lane (one per 128-byte line)
            if ((lane & 7) == 0) { float4 v = ldg_nc(p); acc += v.x * 0.f; }
        } else if (mode == 9) {  // cp.async.ca 8 bytes per lane into a dummy smem slot (no registers, fire and forget)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(&dummy[lane * 2])), "l"(p));
        } else if (mode == 10) {  // cp.async.ca 16 bytes by even lanes
            if ((lane & 1) == 0)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(&dummy[lane * 2])), "l"(p));
        } else if (mode == 11) {  // cp.async.ca 4 bytes per lane
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(&dummy[lane * 2])), "l"(p));
        } else if (mode == 8) {  // byte touch through the non-.nc path (ld.global.ca)
            unsigned int s;
            asm volatile("ld.global.ca.u8 %0, [%1];" : "=r"(s) : "l"(p));
            acc += (float)s * 0.f;
        }
        long long t0 = clock64();
        while (clock64() - t0 < delay) {}
        __syncwarp();
        long long a = clock64();
        float4 v = ldg_nc(p);
        acc += v.x + v.y + v.z + v.w;   // forces completion
        long long b = clock64();
        if (acc == 123456.f) tot += 1;
        tot += b - a;
    }
    if (lane == 0) out[blockIdx.x] = tot / 64;
    if (acc == 987654.f) *sink = acc;
}
int main() {
    const size_t stride_vec = 4096 / 16 * 37;  // rows far apart
    const int ctas = 148;
    size_t n = (size_t)ctas * 64 * stride_vec + 64;
    float4* buf; cudaMalloc(&buf, n * sizeof(float4)); cudaMemset(buf, 0, n * sizeof(float4));
    long long* out; cudaMalloc(&out, ctas * sizeof(long long)); float* sink; cudaMalloc(&sink, 4);
    long long h[148];
    const char* names[] = {"cold (no touch)", "byte touch (LDG.U8)", "128-bit touch", "prefetch.global.L1", "32-bit touch", "64-bit touch", "128-bit even lanes", "128-bit every 8th lane", "byte touch ld.ca", "cp.async.ca 8B/lane", "cp.async.ca 16B even lanes", "cp.async.ca 4B/lane"};
    for (int warm = 0; warm < 2; ++warm) if (warm == 1 || true)
    for (int mode = 0; mode < 12; ++mode) {
        for (int delay : {2000}) {
            // flush L2-resident state? keep: data stays in L2 after first pass (second pass = L2-hot, like the kernel)
            k<<<ctas, 32>>>(buf, stride_vec, mode, delay, out, sink);
            cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
            double s = 0; for (int i = 0; i < ctas; ++i) s += h[i];
            printf("pass %d  %-22s delay %5d cycles -> load latency %.0f cycles\n", warm, names[mode], delay, s / ctas);
        }
    }
    return 0;
}

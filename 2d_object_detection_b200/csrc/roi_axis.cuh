// Shared by the RoI pooling translation units: crop_and_resize sampling positions along one axis.
#pragma once
#include "common.cuh"

namespace frcnn {

constexpr int kMaxCrop = 32;  // S = P*k <= 32
constexpr int kRoiThreads = 256;

struct AxisTab {
    int lo[kMaxCrop];
    int hi[kMaxCrop];
    float lerp[kMaxCrop];
    int valid[kMaxCrop];
};

// Sampling position i of crop_and_resize along one axis; op order of crop_and_resize_op.cc.
__device__ __forceinline__ void axis_sample(float c1, float c2, int extent, int S, int i, AxisTab& t) {
    const float em1 = (float)(extent - 1);
    float in;
    if (S > 1) {
        const float scale = fdiv(fmul(fsub(c2, c1), em1), (float)(S - 1));
        in = fadd(fmul(c1, em1), fmul((float)i, scale));
    } else {
        in = fmul(fmul(0.5f, fadd(c1, c2)), em1);
    }
    const bool ok = (in >= 0.0f) && (in <= em1);  // NaN -> not sampled
    t.valid[i] = ok ? 1 : 0;
    const float fl = ok ? floorf(in) : 0.0f;
    t.lo[i] = (int)fl;
    t.hi[i] = ok ? (int)ceilf(in) : 0;
    t.lerp[i] = ok ? fsub(in, fl) : 0.0f;
}

// Row-streaming forward for pooled_size 7 / kernel_size 2 / C % 512 == 0 (roi_pool_stream.cu).
bool roi_pool_stream_supported(int C, int P, int k);
// Per-RoI records + warp-persistent pooling kernel; workspace = counter + records.
// codes != nullptr: training forward, also writes the arg-max codes ([B,R,P,P,C/4] bytes).
size_t roi_pool_stream_workspace_bytes(int B, int R);
void roi_pool_set_max_sms(int max_sms);
// prepared: the workspace already holds the records of `rois` and fresh counters (launch_roi_records ran before, possibly
// on another stream the caller has ordered this one after) -- only the pooling kernel is launched and `rois` is not read.
int launch_roi_records(int B, int Hf, int Wf, const float* rois, int R, void* workspace, size_t workspace_bytes, cudaStream_t st);
int launch_roi_pool_fwd_persist(const float* feat, int B, int Hf, int Wf, int C, const float* rois, int R, float* out,
                                unsigned char* codes, void* workspace, size_t workspace_bytes, cudaStream_t st,
                                bool prepared = false);

// Backward from the saved arg-max codes on the same walk (TMA reduce-add rows); dfeat must already be zeroed.

int launch_roi_pool_bwd_persist(int B, int Hf, int Wf, int C, const float* rois, int R, const float* dy,
                                const unsigned char* codes, float* dfeat, void* workspace, size_t workspace_bytes, cudaStream_t st,
                                bool prepared = false);

}  // namespace frcnn

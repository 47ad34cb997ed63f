// Host-buffer entry point for the inference slice of FasterRCNN.call (models/faster_rcnn.py:52-55):
// postprocess_output(rpn_output) -> ROIPooling(feature_maps, proposals), taking and returning HOST
// buffers.  This is the call a non-CUDA host (the reference runs CombinedNMS on the CPU) would bind.
//
// Copies and kernels are software-pipelined per image over two streams: the small RPN tensors go up
// first and NMS for the whole batch starts while the feature maps are still in flight; each image's
// RoI pooling starts when its feature map has landed and its pooled block is copied back while the
// next image is pooled, so PCIe traffic in both directions overlaps the kernels.
//
// submit / wait: the device buffers exist twice, so a caller may submit batch i + 1 before waiting for batch i -- the
// host-to-device copies of the next batch then run on the upstream copy engine while the pooled features of the current
// one still drain downstream (PCIe is full duplex).  frcnn_pipeline_run = submit + wait.
#include <new>

#include "common.cuh"

namespace {
constexpr int kSlots = 2;
struct Slot {
    float *d_regions, *d_scores, *d_deltas, *d_feat, *d_rois, *d_roi_scores, *d_pooled;
    int32_t *d_classes, *d_valid;
    void* d_ws;
    unsigned char* d_pool_ws;  // B pooling workspaces (records + counter of one image each)
    cudaEvent_t ev_nms, ev_done[3];
    bool in_flight;
};
}  // namespace

struct frcnn_pipeline {
    int B, Hf, Wf, C, N, R, P, k;
    Slot slot[kSlots];
    size_t ws_bytes, pool_ws_bytes;
    cudaStream_t s_main, s_img[2];
    unsigned next_ticket;
};

using namespace frcnn;

static int wait_slot(frcnn_pipeline* p, int s) {
    Slot& sl = p->slot[s];
    if (!sl.in_flight) return 0;
    for (int i = 0; i < 3; ++i) FRCNN_CUDA(cudaEventSynchronize(sl.ev_done[i]));
    sl.in_flight = false;
    return 0;
}

extern "C" {

void frcnn_pipeline_destroy(frcnn_pipeline_t* p) {
    if (!p) return;
    for (int s = 0; s < kSlots; ++s) {
        Slot& sl = p->slot[s];
        cudaFree(sl.d_regions);
        cudaFree(sl.d_scores);
        cudaFree(sl.d_deltas);
        cudaFree(sl.d_feat);
        cudaFree(sl.d_rois);
        cudaFree(sl.d_roi_scores);
        cudaFree(sl.d_pooled);
        cudaFree(sl.d_classes);
        cudaFree(sl.d_valid);
        cudaFree(sl.d_ws);
        cudaFree(sl.d_pool_ws);
        if (sl.ev_nms) cudaEventDestroy(sl.ev_nms);
        for (int i = 0; i < 3; ++i)
            if (sl.ev_done[i]) cudaEventDestroy(sl.ev_done[i]);
    }
    if (p->s_main) cudaStreamDestroy(p->s_main);
    for (int i = 0; i < 2; ++i)
        if (p->s_img[i]) cudaStreamDestroy(p->s_img[i]);
    delete p;
}

int frcnn_pipeline_create(frcnn_pipeline_t** out, int B, int Hf, int Wf, int C, int N, int max_total, int P, int k) {
    FRCNN_REQUIRE(out && B > 0 && Hf > 0 && Wf > 0 && C > 0 && N > 0 && max_total > 0 && P > 0 && k > 0, FRCNN_EINVAL,
                  "pipeline_create: bad shape");
    frcnn_pipeline* p = new (std::nothrow) frcnn_pipeline();
    FRCNN_REQUIRE(p, FRCNN_EINVAL, "pipeline_create: out of host memory");
    *p = frcnn_pipeline{};
    p->B = B; p->Hf = Hf; p->Wf = Wf; p->C = C; p->N = N; p->R = max_total; p->P = P; p->k = k;
    p->ws_bytes = frcnn_postprocess_workspace_bytes(B, N, 1, max_total, max_total);
    p->pool_ws_bytes = align_up(frcnn_roi_pool_workspace_bytes(1, max_total, C, P, k), 256);
    const size_t R = max_total;
    cudaError_t e = cudaSuccess;
    auto alloc = [&](void** ptr, size_t bytes) {
        if (e == cudaSuccess) e = cudaMalloc(ptr, bytes ? bytes : 16);
    };
    for (int s = 0; s < kSlots; ++s) {
        Slot& sl = p->slot[s];
        alloc((void**)&sl.d_regions, sizeof(float) * 4 * (size_t)N);
        alloc((void**)&sl.d_scores, sizeof(float) * 2 * (size_t)B * N);
        alloc((void**)&sl.d_deltas, sizeof(float) * 4 * (size_t)B * N);
        alloc((void**)&sl.d_feat, sizeof(float) * (size_t)B * Hf * Wf * C);
        alloc((void**)&sl.d_rois, sizeof(float) * 4 * B * R);
        alloc((void**)&sl.d_roi_scores, sizeof(float) * B * R);
        alloc((void**)&sl.d_pooled, sizeof(float) * B * R * P * P * C);
        alloc((void**)&sl.d_classes, sizeof(int32_t) * B * R);
        alloc((void**)&sl.d_valid, sizeof(int32_t) * B);
        alloc(&sl.d_ws, p->ws_bytes);
        alloc((void**)&sl.d_pool_ws, p->pool_ws_bytes * B);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&sl.ev_nms, cudaEventDisableTiming);
        for (int i = 0; i < 3 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&sl.ev_done[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->s_main, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaStreamCreateWithFlags(&p->s_img[i], cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        frcnn_pipeline_destroy(p);
        return cuda_fail(e, "pipeline_create");
    }
    *out = p;
    return 0;
}

int frcnn_pipeline_submit(frcnn_pipeline_t* p, float img_h, float img_w, const float* regions, const float* pred_scores,
                          const float* pred_boxes, const float* feature_maps, float score_threshold, float iou_threshold,
                          int pre_nms_top_n, float* rois, float* roi_scores, int32_t* num_valid, float* pooled,
                          int32_t* ticket) {
    FRCNN_REQUIRE(p && regions && pred_scores && pred_boxes && feature_maps && rois && roi_scores && num_valid && ticket,
                  FRCNN_EINVAL, "pipeline_submit: null argument");
    const unsigned t = p->next_ticket;
    const int s = (int)(t % kSlots);
    int rc = wait_slot(p, s);  // the batch submitted two calls ago must have left this slot's buffers
    if (rc) return rc;
    Slot& sl = p->slot[s];
    const size_t B = p->B, N = p->N, R = p->R;
    const size_t feat_img = (size_t)p->Hf * p->Wf * p->C, pooled_img = R * p->P * p->P * p->C;
    // 1. small RPN tensors up, proposals for the whole batch
    FRCNN_CUDA(cudaMemcpyAsync(sl.d_regions, regions, sizeof(float) * 4 * N, cudaMemcpyHostToDevice, p->s_main));
    FRCNN_CUDA(cudaMemcpyAsync(sl.d_scores, pred_scores, sizeof(float) * 2 * B * N, cudaMemcpyHostToDevice, p->s_main));
    FRCNN_CUDA(cudaMemcpyAsync(sl.d_deltas, pred_boxes, sizeof(float) * 4 * B * N, cudaMemcpyHostToDevice, p->s_main));
    rc = frcnn_postprocess(img_h, img_w, sl.d_regions, 0, sl.d_scores, sl.d_deltas, p->B, p->N, 1, score_threshold,
                           iou_threshold, p->R, p->R, pre_nms_top_n, sl.d_rois, sl.d_roi_scores, sl.d_classes, sl.d_valid,
                           nullptr, nullptr, sl.d_ws, p->ws_bytes, p->s_main);
    if (rc) return rc;
    FRCNN_CUDA(cudaEventRecord(sl.ev_nms, p->s_main));
    FRCNN_CUDA(cudaMemcpyAsync(rois, sl.d_rois, sizeof(float) * 4 * B * R, cudaMemcpyDeviceToHost, p->s_main));
    FRCNN_CUDA(cudaMemcpyAsync(roi_scores, sl.d_roi_scores, sizeof(float) * B * R, cudaMemcpyDeviceToHost, p->s_main));
    FRCNN_CUDA(cudaMemcpyAsync(num_valid, sl.d_valid, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, p->s_main));
    // 2. per image: feature map up, pool, pooled block down (alternating streams)
    for (size_t b = 0; b < B; ++b) {
        cudaStream_t st = p->s_img[b & 1];
        FRCNN_CUDA(cudaMemcpyAsync(sl.d_feat + b * feat_img, feature_maps + b * feat_img, sizeof(float) * feat_img,
                                   cudaMemcpyHostToDevice, st));
        FRCNN_CUDA(cudaStreamWaitEvent(st, sl.ev_nms, 0));
        rc = frcnn_roi_pool_forward(sl.d_feat + b * feat_img, 1, p->Hf, p->Wf, p->C, sl.d_rois + b * R * 4, p->R, p->P, p->k,
                                    sl.d_pooled + b * pooled_img, sl.d_pool_ws + b * p->pool_ws_bytes, p->pool_ws_bytes, st);
        if (rc) return rc;
        if (pooled)
            FRCNN_CUDA(cudaMemcpyAsync(pooled + b * pooled_img, sl.d_pooled + b * pooled_img, sizeof(float) * pooled_img,
                                       cudaMemcpyDeviceToHost, st));
    }
    FRCNN_CUDA(cudaEventRecord(sl.ev_done[0], p->s_main));
    FRCNN_CUDA(cudaEventRecord(sl.ev_done[1], p->s_img[0]));
    FRCNN_CUDA(cudaEventRecord(sl.ev_done[2], p->s_img[1]));
    sl.in_flight = true;
    *ticket = (int32_t)(t & 0x7fffffffu);
    p->next_ticket = t + 1;
    return 0;
}

int frcnn_pipeline_wait(frcnn_pipeline_t* p, int32_t ticket) {
    FRCNN_REQUIRE(p && ticket >= 0, FRCNN_EINVAL, "pipeline_wait: bad argument");
    return wait_slot(p, (int)((unsigned)ticket % kSlots));
}

int frcnn_pipeline_run(frcnn_pipeline_t* p, float img_h, float img_w, const float* regions, const float* pred_scores,
                       const float* pred_boxes, const float* feature_maps, float score_threshold, float iou_threshold,
                       int pre_nms_top_n, float* rois, float* roi_scores, int32_t* num_valid, float* pooled) {
    int32_t ticket = -1;
    const int rc = frcnn_pipeline_submit(p, img_h, img_w, regions, pred_scores, pred_boxes, feature_maps, score_threshold,
                                         iou_threshold, pre_nms_top_n, rois, roi_scores, num_valid, pooled, &ticket);
    if (rc) return rc;
    return frcnn_pipeline_wait(p, ticket);
}

}  // extern "C"

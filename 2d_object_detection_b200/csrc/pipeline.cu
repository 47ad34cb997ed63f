// Host-buffer entry point for the inference slice of FasterRCNN.call (models/faster_rcnn.py:52-55):
// postprocess_output(rpn_output) -> ROIPooling(feature_maps, proposals), taking and returning HOST
// buffers.  This is the call a non-CUDA host (the reference runs CombinedNMS on the CPU) would bind.
//
// Copies and kernels are software-pipelined per image over two streams: the small RPN tensors go up
// first and NMS for the whole batch starts while the feature maps are still in flight; each image's
// RoI pooling starts when its feature map has landed and its pooled block is copied back while the
// next image is pooled, so PCIe traffic in both directions overlaps the kernels.
#include <new>

#include "common.cuh"

struct frcnn_pipeline {
    int B, Hf, Wf, C, N, R, P, k;
    float *d_regions, *d_scores, *d_deltas, *d_feat, *d_rois, *d_roi_scores, *d_pooled;
    int32_t *d_classes, *d_valid;
    void* d_ws;
    size_t ws_bytes;
    cudaStream_t s_main, s_img[2];
    cudaEvent_t ev_nms;
};

using namespace frcnn;

extern "C" {

void frcnn_pipeline_destroy(frcnn_pipeline_t* p) {
    if (!p) return;
    cudaFree(p->d_regions);
    cudaFree(p->d_scores);
    cudaFree(p->d_deltas);
    cudaFree(p->d_feat);
    cudaFree(p->d_rois);
    cudaFree(p->d_roi_scores);
    cudaFree(p->d_pooled);
    cudaFree(p->d_classes);
    cudaFree(p->d_valid);
    cudaFree(p->d_ws);
    if (p->s_main) cudaStreamDestroy(p->s_main);
    for (int i = 0; i < 2; ++i)
        if (p->s_img[i]) cudaStreamDestroy(p->s_img[i]);
    if (p->ev_nms) cudaEventDestroy(p->ev_nms);
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
    const size_t R = max_total;
    cudaError_t e = cudaSuccess;
    auto alloc = [&](void** ptr, size_t bytes) {
        if (e == cudaSuccess) e = cudaMalloc(ptr, bytes ? bytes : 16);
    };
    alloc((void**)&p->d_regions, sizeof(float) * 4 * (size_t)N);
    alloc((void**)&p->d_scores, sizeof(float) * 2 * (size_t)B * N);
    alloc((void**)&p->d_deltas, sizeof(float) * 4 * (size_t)B * N);
    alloc((void**)&p->d_feat, sizeof(float) * (size_t)B * Hf * Wf * C);
    alloc((void**)&p->d_rois, sizeof(float) * 4 * B * R);
    alloc((void**)&p->d_roi_scores, sizeof(float) * B * R);
    alloc((void**)&p->d_pooled, sizeof(float) * B * R * P * P * C);
    alloc((void**)&p->d_classes, sizeof(int32_t) * B * R);
    alloc((void**)&p->d_valid, sizeof(int32_t) * B);
    alloc(&p->d_ws, p->ws_bytes);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->s_main, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaStreamCreateWithFlags(&p->s_img[i], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_nms, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        frcnn_pipeline_destroy(p);
        return cuda_fail(e, "pipeline_create");
    }
    *out = p;
    return 0;
}

int frcnn_pipeline_run(frcnn_pipeline_t* p, float img_h, float img_w, const float* regions, const float* pred_scores,
                       const float* pred_boxes, const float* feature_maps, float score_threshold, float iou_threshold,
                       int pre_nms_top_n, float* rois, float* roi_scores, int32_t* num_valid, float* pooled) {
    FRCNN_REQUIRE(p && regions && pred_scores && pred_boxes && feature_maps && rois && roi_scores && num_valid,
                  FRCNN_EINVAL, "pipeline_run: null argument");
    const size_t B = p->B, N = p->N, R = p->R;
    const size_t feat_img = (size_t)p->Hf * p->Wf * p->C, pooled_img = R * p->P * p->P * p->C;
    // 1. small RPN tensors up, proposals for the whole batch
    FRCNN_CUDA(cudaMemcpyAsync(p->d_regions, regions, sizeof(float) * 4 * N, cudaMemcpyHostToDevice, p->s_main));
    FRCNN_CUDA(cudaMemcpyAsync(p->d_scores, pred_scores, sizeof(float) * 2 * B * N, cudaMemcpyHostToDevice, p->s_main));
    FRCNN_CUDA(cudaMemcpyAsync(p->d_deltas, pred_boxes, sizeof(float) * 4 * B * N, cudaMemcpyHostToDevice, p->s_main));
    int rc = frcnn_postprocess(img_h, img_w, p->d_regions, 0, p->d_scores, p->d_deltas, p->B, p->N, 1, score_threshold,
                               iou_threshold, p->R, p->R, pre_nms_top_n, p->d_rois, p->d_roi_scores, p->d_classes,
                               p->d_valid, nullptr, nullptr, p->d_ws, p->ws_bytes, p->s_main);
    if (rc) return rc;
    FRCNN_CUDA(cudaEventRecord(p->ev_nms, p->s_main));
    FRCNN_CUDA(cudaMemcpyAsync(rois, p->d_rois, sizeof(float) * 4 * B * R, cudaMemcpyDeviceToHost, p->s_main));
    FRCNN_CUDA(cudaMemcpyAsync(roi_scores, p->d_roi_scores, sizeof(float) * B * R, cudaMemcpyDeviceToHost, p->s_main));
    FRCNN_CUDA(cudaMemcpyAsync(num_valid, p->d_valid, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, p->s_main));
    // 2. per image: feature map up, pool, pooled block down (alternating streams)
    for (size_t b = 0; b < B; ++b) {
        cudaStream_t s = p->s_img[b & 1];
        FRCNN_CUDA(cudaMemcpyAsync(p->d_feat + b * feat_img, feature_maps + b * feat_img, sizeof(float) * feat_img,
                                   cudaMemcpyHostToDevice, s));
        FRCNN_CUDA(cudaStreamWaitEvent(s, p->ev_nms, 0));
        rc = frcnn_roi_pool_forward(p->d_feat + b * feat_img, 1, p->Hf, p->Wf, p->C, p->d_rois + b * R * 4, p->R, p->P,
                                    p->k, p->d_pooled + b * pooled_img, s);
        if (rc) return rc;
        if (pooled)
            FRCNN_CUDA(cudaMemcpyAsync(pooled + b * pooled_img, p->d_pooled + b * pooled_img, sizeof(float) * pooled_img,
                                       cudaMemcpyDeviceToHost, s));
    }
    FRCNN_CUDA(cudaStreamSynchronize(p->s_main));
    FRCNN_CUDA(cudaStreamSynchronize(p->s_img[0]));
    FRCNN_CUDA(cudaStreamSynchronize(p->s_img[1]));
    return 0;
}

}  // extern "C"

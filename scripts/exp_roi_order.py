import sys, importlib, numpy as np, torch
sys.path.insert(0,'/root/repo')
det = importlib.import_module('2d_object_detection_b200'); syn = det.synthetic
w = syn.WORKLOADS['kitti_b8']; rng = np.random.default_rng(5); dev='cuda'
anchors = det.clip_to_window(det.generate_anchors(w.grid, w.scales, w.aspect_ratios, w.base_anchor_shape), [0,0,w.image_shape[1],w.image_shape[0]])
scores, deltas = syn.rpn_head_outputs(rng, w.batch, w.num_anchors)
feat = torch.as_tensor(syn.feature_maps(rng, w.batch, w.grid, w.channels)).to(dev)
out = det.postprocess_output(w.image_shape, anchors, torch.as_tensor(scores).to(dev), torch.as_tensor(deltas).to(dev), **w.nms_kwargs(), pre_nms_top_n=w.pre_nms_top_n)
rois = out['pred_boxes']
flush = torch.empty(256<<20, dtype=torch.uint8, device=dev)
def timeit(r):
    ts=[]
    for i in range(25):
        flush.zero_()
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); det.roi_pool_forward(feat, r, 7, 2); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts[5:]))
print('score order      ', timeit(rois))
h = (rois[...,3]-rois[...,1]); wd=(rois[...,2]-rois[...,0])
for name,key in (('height desc',-h),('height asc',h),('area desc',-(h*wd)),('rows+ desc', -(h*23 + 3))):
    idx = torch.argsort(key, dim=1)
    r2 = torch.gather(rois, 1, idx[...,None].expand(-1,-1,4)).contiguous()
    print(f'{name:16s}', timeit(r2))
# global LPT across the batch is not expressible through [B,R] layout; interleave images instead: handled by grid order

"""Quick bit-exactness check of the RoI forward against the oracle on awkward RoIs (default: row-streaming kernel;
FRCNN_FWD_PERSIST=1: persistent variant; FRCNN_FWD_WALK=1: bin-row walk).  Test infrastructure, like tests/."""
import sys, importlib, os, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tests/ -> repo root; sys.path.insert(0, ROOT)
det = importlib.import_module('2d_object_detection_b200'); syn = det.synthetic
import oracle
rng = np.random.default_rng(0)
for grid, C, R in [((24, 78), 1024, 40), ((9, 11), 512, 8), ((50, 84), 512, 30), ((3, 2), 512, 6)]:
    f = syn.feature_maps(rng, 2, grid, C, relu=True)
    rois = syn.random_rois(rng, 2, R, num_valid=[R - 2, R]); rois[1, 0] = [0.25, 0.1, 1.0, 1.0]
    rois[1, 1] = [0.0, 0.0, 1.0, 1.0]; rois[1, 2] = [0.9, 0.8, 0.1, 0.2]; rois[1, 3] = [-0.2, 0.3, 0.4, 1.3]; rois[0, 0] = [0.02, 0.0, 0.97, 0.5]
    out = det.roi_pool_forward(torch.as_tensor(f).cuda(), torch.as_tensor(rois).cuda(), 7, 2); torch.cuda.synchronize()
    ref = oracle.roi_pool_forward(f, rois, 7, 2)
    got = out.cpu().numpy()
    bad = [(b, r) for b in range(2) for r in range(R) if not np.array_equal(got[b, r], ref[b, r])]
    print(grid, C, 'ok' if not bad else ('MISMATCH', bad[:10]), flush=True)

import sys, importlib, ctypes, time, os, numpy as np, torch
sys.path.insert(0,'/root/repo')
det = importlib.import_module('2d_object_detection_b200'); syn = det.synthetic
rng = np.random.default_rng(0)
L = ctypes.CDLL('/root/repo/2d_object_detection_b200/libfrcnn_b200.so')
trace = torch.zeros(64, dtype=torch.int32).pin_memory()
print('set trace', L.frcnn_debug_set_trace(ctypes.c_void_p(trace.data_ptr())), flush=True)
grid, C, R = (24, 78), 1024, 40
f = syn.feature_maps(rng, 2, grid, C, relu=True)
rois = syn.random_rois(rng, 2, R, num_valid=[R-2,R]); rois[1,0]=[0.25,0.1,1.0,1.0]
ft = torch.as_tensor(f).cuda(); rt = torch.as_tensor(rois[:, 0:1].copy()).cuda(); torch.cuda.synchronize()
out = det.roi_pool_forward(ft, rt, 7, 2)
time.sleep(2.0)
print('trace', trace[:8].tolist(), flush=True)
buf = (ctypes.c_int * 136)()
os._exit(0)

"""cfg4 exactly as bench_configs.cfg4 runs it, stand-alone (no 1080p workload before it), with per-call timing"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pydensecrf_b200.densecrf as dcrf
from pydensecrf_b200.utils import pinned_empty
from tests import workloads as wl
W = H = 512
L, n, chunk, distinct = 21, 5, 64, 16
imgs = np.stack([wl.smooth_image(H, W, seed=100 + i) for i in range(distinct)])
Us = np.stack([wl.softmax_unary(H, W, L, seed=100 + i) for i in range(distinct)])
ub = pinned_empty((chunk, L, H * W), np.float32)
ib = pinned_empty((chunk, H, W, 3), np.uint8)
for j in range(chunk):
    ub[j], ib[j] = Us[j % distinct], imgs[j % distinct]
dcrf.map_batch_2d(ub[:16], ib[:16], n_iter=n, device=0)
for rep in range(6):
    t0 = time.perf_counter()
    lab = dcrf.map_batch_2d(ub, ib, n_iter=n, device=0)
    print("call %d: %.1f ms" % (rep, (time.perf_counter() - t0) * 1e3), flush=True)

"""map_batch_2d on 512x512 images: wall time per call and free device memory after it, per (chunk, workers) setting."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pydensecrf_b200.densecrf as dcrf
from pydensecrf_b200.utils import pinned_empty
from tests import workloads as wl
import torch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
W = H = 512; L = 21
ub = pinned_empty((B, L, H * W), np.float32); ib = pinned_empty((B, H, W, 3), np.uint8)
for j in range(B):
    ub[j] = wl.softmax_unary(H, W, L, seed=100 + j % 4); ib[j] = wl.smooth_image(H, W, seed=100 + j % 4)
for chunk, workers in ((0, 0),):
    os.environ["DCRF_B200_BATCH_CHUNK"] = str(chunk)
    for rep in range(4):
        t0 = time.perf_counter()
        lab = dcrf.map_batch_2d(ub, ib, n_iter=5, workers=workers)
        dt = time.perf_counter() - t0
        free, total = torch.cuda.mem_get_info()
        print("chunk %2d workers %d call %d: %7.1f ms for %d images = %6.0f img/s, device memory in use %.1f GB" % (
            chunk, workers, rep, dt * 1e3, B, B / dt, (total - free) / 1e9), flush=True)

"""setUnaryEnergy from ordinary (pageable) numpy arrays at several sizes, arrays taken in turn so that they are not cache-hot;
DCRF_B200_UPLOAD_THREADS=0 shows the driver's own staging for comparison."""
import os, sys, time, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pydensecrf_b200.densecrf as dcrf
for (W, H, L) in ((320, 240, 21), (512, 512, 8), (512, 512, 21), (640, 480, 21), (1024, 1024, 21), (1920, 1080, 21)):
    d = dcrf.DenseCRF2D(W, H, L)
    Us = [np.random.default_rng(i).random((L, W * H), dtype=np.float32) for i in range(4)]
    for _ in range(3): d.setUnaryEnergy(Us[0])
    t0 = time.perf_counter()
    for i in range(20): d.setUnaryEnergy(Us[i % 4])
    dt = (time.perf_counter() - t0) / 20
    print("%4dx%4dx%2d %6.1f MB pageable: %.3f ms (%.1f GB/s)" % (W, H, L, Us[0].nbytes / 1e6, dt * 1e3, Us[0].nbytes / dt / 1e9))

"""Create / destroy many models and result matrices and watch device memory (pools must plateau)."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pydensecrf_b200.densecrf as dcrf
from tests import workloads as wl
def used():
    return int(subprocess.run(["nvidia-smi", "--query-gpu=memory.used", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.split()[0])
W = H = 384; L = 21
img = wl.smooth_image(H, W, seed=0); U = wl.softmax_unary(H, W, L, seed=0)
marks = []
for i in range(300):
    d = dcrf.DenseCRF2D(W, H, L)
    d.setUnaryEnergy(U); d.addPairwiseGaussian(3, 3); d.addPairwiseBilateral(80, 13, img, 10)
    q = np.asarray(d.inference(3)); q2 = d.inference(3); q3 = d.inference(3); kl = d.klDivergence(q3)
    if i % 50 == 49: marks.append(used())
print("device MiB after every 50 models:", marks)
assert marks[-1] - marks[1] <= 64, marks
print("leak-ok")

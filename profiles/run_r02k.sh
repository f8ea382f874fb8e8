(timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6)
for sb in 1 0; do DCRF_B200_STAGED_BLUR=$sb python profiles/quick_bench.py staged_blur=$sb 3; done
DCRF_B200_STAGED_BLUR=1 python profiles/batch_probe.py 128 2>&1 | grep "chunk  0"
DCRF_B200_STAGED_BLUR=0 python profiles/batch_probe.py 128 2>&1 | grep "chunk  0"

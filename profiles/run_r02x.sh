(timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_fuzz.py -m gpu -x -q -k "batch or super or concurrent or staged" 2>&1 | tail -3) | tee gpurun_out/r02x_pytest.log
for sl in 1 0 1; do DCRF_B200_BATCH_SPLIT_LOCK=$sl timeout 200 python profiles/batch_probe.py 128 2>&1 | tail -2 | sed "s/^/split_lock=$sl /"; done | tee gpurun_out/r02x_batch.log

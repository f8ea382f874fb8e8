(timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q -k "tensor or batch or super" -s 2>&1 | grep -E "KL |passed|failed|Error" | tail -12) | tee gpurun_out/r02w_pytest.log
for sl in 1 0 1 0; do DCRF_B200_BATCH_SPLIT_LOCK=$sl timeout 200 python profiles/batch_probe.py 128 2>&1 | tail -2 | sed "s/^/split_lock=$sl /"; done | tee gpurun_out/r02w_batch.log

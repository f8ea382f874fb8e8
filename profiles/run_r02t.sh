(LD_LIBRARY_PATH=build/variants/tma_kb16 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor" -s 2>&1 | grep -E "tensor-core|passed|failed|Error" | tail -14) | tee gpurun_out/r02t_kb16_pytest.log
timeout 200 python profiles/cfg3_probe.py kb32 3 2>&1 | tail -1 | tee gpurun_out/r02t_probe.log
LD_LIBRARY_PATH=build/variants/tma_kb16 timeout 200 python profiles/cfg3_probe.py kb16 3 2>&1 | tail -1 | tee -a gpurun_out/r02t_probe.log
LD_LIBRARY_PATH=build/variants/tma_kb16_timing timeout 200 python profiles/cfg3_probe.py kb16_timing 1 2>&1 | tail -8 | tee gpurun_out/r02t_timing.log

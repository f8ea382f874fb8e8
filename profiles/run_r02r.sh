(timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor or matrix" 2>&1 | tail -4) | tee gpurun_out/r02r_pytest.log
timeout 200 python profiles/cfg3_probe.py pingpong 3 2>&1 | tail -3 | tee gpurun_out/r02r_probe.log
LD_LIBRARY_PATH=build/variants/tma_timing timeout 200 python profiles/cfg3_probe.py timing 1 2>&1 | tail -8 | tee gpurun_out/r02r_timing.log
(LD_LIBRARY_PATH=build/variants/tma_trunc timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor" -s 2>&1 | grep -E "tensor-core|passed|failed|Error" | tail -14) | tee gpurun_out/r02r_trunc_pytest.log
LD_LIBRARY_PATH=build/variants/tma_trunc timeout 200 python profiles/cfg3_probe.py trunc 3 2>&1 | tail -3 | tee -a gpurun_out/r02r_probe.log
LD_LIBRARY_PATH=build/variants/tma_trunc_timing timeout 200 python profiles/cfg3_probe.py trunc_timing 1 2>&1 | tail -8 | tee -a gpurun_out/r02r_timing.log

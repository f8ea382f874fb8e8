(timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor or matrix" 2>&1 | tail -4) | tee gpurun_out/r02q_pytest.log
timeout 200 python profiles/cfg3_probe.py nb3 3 2>&1 | tail -3 | tee gpurun_out/r02q_probe.log
LD_LIBRARY_PATH=build/variants/tma_timing timeout 200 python profiles/cfg3_probe.py timing 1 2>&1 | tail -8 | tee gpurun_out/r02q_timing.log

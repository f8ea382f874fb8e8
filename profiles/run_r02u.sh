(timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor or matrix" 2>&1 | tail -3) | tee gpurun_out/r02u_pytest.log
timeout 200 python profiles/cfg3_probe.py relaxed 3 2>&1 | tail -1 | tee gpurun_out/r02u_probe.log
LD_LIBRARY_PATH=build/variants/tma_timing timeout 200 python profiles/cfg3_probe.py timing 1 2>&1 | tail -8 | tee gpurun_out/r02u_timing.log

cd "$(dirname "$0")/.."
(timeout -s KILL 150 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_core_update" -s 2>&1 | grep -E "tensor-core|passed|failed|Error|error" | tail -20) | tee gpurun_out/r03i_pytest.log
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
timeout -s KILL 200 python profiles/cfg3_probe.py pair 3 2>&1 | tail -2 | tee gpurun_out/r03i_probe.log
DCRF_B200_TCU_PAIR=0 timeout -s KILL 200 python profiles/cfg3_probe.py single 3 2>&1 | tail -1 | tee -a gpurun_out/r03i_probe.log
LD_LIBRARY_PATH=build/variants/tma_timing timeout -s KILL 200 python profiles/cfg3_probe.py timing 1 2>&1 | tail -12 | tee gpurun_out/r03i_timing.log

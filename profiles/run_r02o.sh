(timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor or matrix" -s 2>&1 | tail -40) > gpurun_out/r02o_pytest.log; tail -30 gpurun_out/r02o_pytest.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --configs cfg3 --no-pipelined 2>gpurun_out/r02o_cfg3.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps(d['configs']['cfg3'])[:1500])" | tee gpurun_out/r02o_cfg3.txt
DCRF_B200_TCU_TMA=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --configs cfg3 --no-pipelined 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('old kernel', json.dumps(d['configs']['cfg3']['per_launch_ms']), d['configs']['cfg3']['ms_per_iteration'])" | tee -a gpurun_out/r02o_cfg3.txt
tail -5 gpurun_out/r02o_cfg3.err

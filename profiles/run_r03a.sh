(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/r03a_pytest.log; tail -4 gpurun_out/r03a_pytest.log
timeout 200 python profiles/cfg3_probe.py init_tma 3 2>&1 | tail -1 | tee gpurun_out/r03a_probe.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --configs cfg3 --no-pipelined 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps(d['configs']['cfg3'])[:900])" | tee -a gpurun_out/r03a_probe.log

(timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/r02e_pytest.log; tail -8 gpurun_out/r02e_pytest.log
for c in 0 8 16 32 64; do DCRF_B200_BATCH_CHUNK=$c python bench.py --steps 3 --warmup 3 --no-cpu-baseline --configs cfg4 --no-pipelined 2>gpurun_out/r02e_err_$c.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('chunk $c', d['configs']['cfg4'].get('value'), d['configs']['cfg4'].get('error'))"; done
DCRF_B200_BATCH_MODE=0 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --configs cfg4 --no-pipelined 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('per-image mode', d['configs']['cfg4'].get('value'))"

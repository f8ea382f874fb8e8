cd "$(dirname "$0")/.."
for p in 1 0 1 0; do DCRF_B200_AUX_PRIO=$p timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --configs cfg3 --no-pipelined 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); c=d['configs']['cfg3']; print('prio=$p', c['ms_per_inference'], c['per_launch_ms'])"; done

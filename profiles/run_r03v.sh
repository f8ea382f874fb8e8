cd "$(dirname "$0")/.."
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2)
timeout 300 python profiles/cfg5_probe.py 2>&1 | tail -1
DCRF_B200_COOP_SPLAT=0 timeout 300 python profiles/cfg5_probe.py 2>&1 | tail -1

cd "$(dirname "$0")/.."
for p in 1 0 1 0; do DCRF_B200_AUX_PRIO=$p timeout 120 python profiles/quick_bench.py prio=$p 3 2>&1 | grep "mask="; done
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3)

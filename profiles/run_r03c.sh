cd "$(dirname "$0")/.."
for p in 0 1 -1 0 1; do DCRF_B200_AUX_PRIO=$p timeout 120 python profiles/quick_bench.py prio=$p 3 2>&1 | grep "mask="; done

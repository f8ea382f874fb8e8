cd "$(dirname "$0")/.."
for env in "DCRF_B200_TCU_PAIR=0" "DCRF_B200_TCU_TMA=0" "DCRF_B200_AUX_PRIO=0" "DCRF_B200_BATCH_SPLIT_LOCK=0" "DCRF_B200_NO_GRAPHS=1" "DCRF_B200_OVERLAP=0"; do
  echo "$env: $(env $env timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -1)"
done

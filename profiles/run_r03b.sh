cd "$(dirname "$0")/.."
timeout 120 python profiles/quick_bench.py base 3 2>&1 | grep -v "^$"
for dir in build/variants/splat_*/; do
  name=$(basename "$dir")
  LD_LIBRARY_PATH=$dir timeout 120 python profiles/quick_bench.py "$name" 3 2>&1 | grep -v "^$"
done
timeout 120 python profiles/quick_bench.py base 3 2>&1 | grep -v "^$"

cd "$(dirname "$0")/.."
timeout 120 python profiles/quick_bench.py base 3 2>&1 | grep "mask="
for v in 16 32; do LD_LIBRARY_PATH=build/variants/splat_al$v timeout 120 python profiles/quick_bench.py timing_only_aligned$v 3 2>&1 | grep "mask="; done
timeout 120 python profiles/quick_bench.py base 3 2>&1 | grep "mask="

#!/bin/bash
# GPU-box side of the round-2 kernel A/B: every library variant under build/variants/ (profiles/build_variant.py).
cd "$(dirname "$0")/.."
for dir in build/variants/*/; do
  name=$(basename "$dir")
  LD_LIBRARY_PATH=$dir timeout 120 python profiles/quick_bench.py "$name" $MASKS 2>&1 | grep -v "^$"
done

cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2)
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 200 gpurun_out/final_bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/final_bench_ref.json 2>/dev/null; head -c 300 gpurun_out/final_bench_ref.json

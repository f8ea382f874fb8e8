cd "$(dirname "$0")/.."
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/r03f_pytest.log; tail -2 gpurun_out/r03f_pytest.log
timeout 900 python bench.py > gpurun_out/r03f_bench.json 2> gpurun_out/r03f_bench.err; tail -c 300 gpurun_out/r03f_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r03f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-pipelined > gpurun_out/r03f_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_fused_update|k_splat|k_blur' -s 130 -c 13 -o gpurun_out/prof_r03f python profiles/prof_workload.py > gpurun_out/r03f_ncu_full.log 2>&1; tail -2 gpurun_out/r03f_ncu_full.log

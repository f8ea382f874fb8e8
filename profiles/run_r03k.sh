cd "$(dirname "$0")/.."
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_update_matrix_tc_tma -s 3 -c 1 -o gpurun_out/prof_r03k_tma_pair python profiles/cfg3_probe.py ncu 1 > gpurun_out/r03k_ncu.log 2>&1; tail -2 gpurun_out/r03k_ncu.log
timeout 900 python bench.py > gpurun_out/r03k_bench.json 2> gpurun_out/r03k_bench.err; tail -c 300 gpurun_out/r03k_bench.err

cd "$(dirname "$0")/.."
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/r03j_pytest.log; tail -2 gpurun_out/r03j_pytest.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --configs cfg3 --no-pipelined 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); c=d['configs']['cfg3']; print('pair', c['ms_per_inference'], c['per_launch_ms'])" | tee gpurun_out/r03j_cfg3.log
DCRF_B200_TCU_PAIR=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --configs cfg3 --no-pipelined 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); c=d['configs']['cfg3']; print('single', c['ms_per_inference'], c['per_launch_ms'])" | tee -a gpurun_out/r03j_cfg3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_update_matrix_tc_tma -s 2 -c 1 -o gpurun_out/prof_r03j_tma_pair python profiles/cfg3_probe.py ncu 1 > gpurun_out/r03j_ncu.log 2>&1; tail -2 gpurun_out/r03j_ncu.log
python profiles/sanitize_workload.py > /dev/null 2>&1; for tool in memcheck racecheck; do timeout 900 compute-sanitizer --tool $tool python profiles/sanitize_workload.py > gpurun_out/r03j_$tool.log 2>&1; echo "$tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r03j_$tool.log | tail -1) $(grep -c sanitize-ok gpurun_out/r03j_$tool.log)"; done | tee gpurun_out/r03j_sanitizer.txt

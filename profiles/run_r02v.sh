(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r02v_pytest.log; tail -5 gpurun_out/r02v_pytest.log
timeout 900 python bench.py > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err; tail -c 400 gpurun_out/r02v_bench.err

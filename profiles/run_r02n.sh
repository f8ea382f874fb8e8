(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r02n_pytest.log; tail -5 gpurun_out/r02n_pytest.log
timeout 600 python bench.py > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err; tail -c 600 gpurun_out/r02n_bench.err

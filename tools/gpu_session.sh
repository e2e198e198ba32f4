mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/s35_pytest.log 2>&1; tail -4 gpurun_out/s35_pytest.log
(time timeout 1200 python bench.py) > gpurun_out/s35_bench.json 2> gpurun_out/s35_bench.err; tail -c 3000 gpurun_out/s35_bench.json; tail -5 gpurun_out/s35_bench.err

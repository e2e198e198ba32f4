mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_beagle_shim.py -x -q) > gpurun_out/s33_pytest.log 2>&1; tail -15 gpurun_out/s33_pytest.log
timeout 300 python tools/beagle_shim_bench.py > gpurun_out/r02_beagle_shim_bench.json 2> gpurun_out/s33_bench.err; cat gpurun_out/r02_beagle_shim_bench.json; tail -3 gpurun_out/s33_bench.err

mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/s18_pytest.log 2>&1; tail -8 gpurun_out/s18_pytest.log

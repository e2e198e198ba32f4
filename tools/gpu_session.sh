mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_integration_gpu.py -x -q) > gpurun_out/s14_pytest.log 2>&1; tail -12 gpurun_out/s14_pytest.log

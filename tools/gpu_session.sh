mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gp_gpu.py tests/test_integration_gpu.py -x -q) > gpurun_out/s34_pytest.log 2>&1; tail -15 gpurun_out/s34_pytest.log
timeout 120 python tools/gp_kernel_time.py 2>&1 | tail -1

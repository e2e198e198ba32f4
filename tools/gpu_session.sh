mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gp_gpu.py tests/test_integration_gpu.py -x -q) > gpurun_out/s37_pytest.log 2>&1; tail -5 gpurun_out/s37_pytest.log
timeout 120 python tools/gp_kernel_time.py 2>&1 | tail -1
SBNB_GP_FUSE=0 timeout 120 python tools/gp_kernel_time.py 2>&1 | tail -1
timeout 600 python tools/gp_bench.py 2>&1 | tail -3

mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_sharding_gpu.py tests/test_group_gpu.py tests/test_engine_gpu.py -x -q) > gpurun_out/s21_pytest.log 2>&1; tail -12 gpurun_out/s21_pytest.log

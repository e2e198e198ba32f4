mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_full_size_gpu.py -m gpu -x -q -k beagle_compatible) > gpurun_out/s59_pytest.log 2>&1; tail -25 gpurun_out/s59_pytest.log

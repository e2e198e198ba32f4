mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gp_gpu.py -m gpu -x -q -k "cluster_exchange") > gpurun_out/s70_pytest.log 2>&1; tail -15 gpurun_out/s70_pytest.log

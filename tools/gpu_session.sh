mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_site_pattern.py tests/test_integration_gpu.py -m gpu -x -q) > gpurun_out/s57_pytest.log 2>&1; tail -3 gpurun_out/s57_pytest.log
timeout 900 python tools/compress_bench.py 2>&1 | tee gpurun_out/r02_compress_bench.jsonl | cut -c1-420

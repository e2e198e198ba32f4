mkdir -p gpurun_out
python tools/gp_bench.py 2>&1 | tail -4 | cut -c1-700 | tee gpurun_out/r02_gp_bench_ds1.jsonl
(timeout 900 python -m pytest tests/test_gp_gpu.py tests/test_integration_gpu.py -x -q) 2>&1 | tail -3

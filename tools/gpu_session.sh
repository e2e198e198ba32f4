mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/s32_pytest.log 2>&1; tail -5 gpurun_out/s32_pytest.log
timeout 120 python tools/gp_kernel_time.py 2>&1 | tail -1 > gpurun_out/r02_gp_kernel_time_ds1.json; cat gpurun_out/r02_gp_kernel_time_ds1.json
timeout 600 python tools/gp_bench.py > gpurun_out/r02_gp_bench_ds1_v2.jsonl 2>&1; cat gpurun_out/r02_gp_bench_ds1_v2.jsonl

mkdir -p gpurun_out
nvidia-smi -L
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3) > gpurun_out/s11_bench_2gpu.json 2> gpurun_out/s11_bench_2gpu.err; tail -c 800 gpurun_out/s11_bench_2gpu.err; wc -c gpurun_out/s11_bench_2gpu.json
(timeout 900 python -m pytest tests/test_group_gpu.py tests/test_sharding_gpu.py -x -q) > gpurun_out/s11_pytest.log 2>&1; tail -5 gpurun_out/s11_pytest.log

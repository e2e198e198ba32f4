mkdir -p gpurun_out
(time python bench.py --impl reference --steps 3 --warmup 1) > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_ref.err; tail -c 200 gpurun_out/r02_ref.err
(time python bench.py) > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -c 300 gpurun_out/r02_bench.err; wc -c gpurun_out/r02_bench.json

# Development aid: the command list of one `gpurun -- 'bash tools/gpu_session.sh'` call (edited per session).
mkdir -p gpurun_out
(time python bench.py > gpurun_out/s58_bench.json) 2> gpurun_out/s58_bench.err; tail -4 gpurun_out/s58_bench.err; cut -c1-300 gpurun_out/s58_bench.json

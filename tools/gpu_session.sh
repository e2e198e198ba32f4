# Development aid: the command list of one `gpurun -- 'bash tools/gpu_session.sh'` call (edited per session).
mkdir -p gpurun_out
timeout 900 python tools/compress_bench.py 2>&1 | tee gpurun_out/r02_compress_bench.jsonl | cut -c1-200

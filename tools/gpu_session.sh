# Development aid: the command list of one `gpurun -- 'bash tools/gpu_session.sh'` call (edited per session).
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"HashKernel|EmitKernel|TileKernel" -s 4 -c 4 -o gpurun_out/r02_site_pattern python tools/compress_bench.py --only-repeats > gpurun_out/s61.log 2>&1; tail -2 gpurun_out/s61.log | cut -c1-200

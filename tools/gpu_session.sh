# Development aid: the command list of one `gpurun -- 'bash tools/gpu_session.sh'` call (edited per session).
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck python tools/sanitizer_cases_shim_patterns.py > gpurun_out/r02_sanitizer_memcheck_shim_patterns.log 2>&1; tail -4 gpurun_out/r02_sanitizer_memcheck_shim_patterns.log
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitizer_cases_shim_patterns.py > gpurun_out/r02_sanitizer_racecheck_shim_patterns.log 2>&1; tail -4 gpurun_out/r02_sanitizer_racecheck_shim_patterns.log

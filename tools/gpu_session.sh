# Development aid: the command list of one `gpurun -- 'bash tools/gpu_session.sh'` call (edited per session).
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitizer_cases_gp_beagle.py > gpurun_out/r02_sanitizer_racecheck_gp_beagle.log 2>&1; tail -3 gpurun_out/r02_sanitizer_racecheck_gp_beagle.log
timeout 1200 compute-sanitizer --tool memcheck python tools/sanitizer_cases_gp_beagle.py > gpurun_out/r02_sanitizer_memcheck_gp_beagle.log 2>&1; tail -2 gpurun_out/r02_sanitizer_memcheck_gp_beagle.log
(timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/session_pytest.log 2>&1; tail -3 gpurun_out/session_pytest.log

# Development aid: the command list of one `gpurun -- 'bash tools/gpu_session.sh'` call (edited per session).
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gp_gpu.py tests/test_integration_gpu.py -m gpu -x -q) > gpurun_out/s64_pytest.log 2>&1; tail -3 gpurun_out/s64_pytest.log
SBNB_GP_NO_CLUSTER=1 timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitizer_cases_gp_beagle.py > gpurun_out/r02_sanitizer_racecheck_gp_beagle.log 2>&1; tail -2 gpurun_out/r02_sanitizer_racecheck_gp_beagle.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitizer_cases_gp_beagle.py > gpurun_out/r02_sanitizer_racecheck_gp_cluster.log 2>&1; grep -c "Race reported" gpurun_out/r02_sanitizer_racecheck_gp_cluster.log; grep "Race reported" gpurun_out/r02_sanitizer_racecheck_gp_cluster.log | grep -o "gp_engine.cu:[0-9]*" | sort | uniq -c; grep "and .* access" gpurun_out/r02_sanitizer_racecheck_gp_cluster.log | grep -o "gp_engine.cu:[0-9]*" | sort | uniq -c
timeout 900 compute-sanitizer --tool memcheck python tools/sanitizer_cases_gp_beagle.py > gpurun_out/r02_sanitizer_memcheck_gp_beagle.log 2>&1; tail -1 gpurun_out/r02_sanitizer_memcheck_gp_beagle.log
timeout 120 python tools/gp_kernel_time.py --repeats 30 2>/dev/null | tail -1 | tee gpurun_out/r02_gp_kernel_time_ds1.json | cut -c1-300

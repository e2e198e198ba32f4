# Development aid: the command list of one `gpurun -- 'bash tools/gpu_session.sh'` call (edited per session).
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests/test_gp_gpu.py -x -q) > gpurun_out/session_pytest.log 2>&1; tail -15 gpurun_out/session_pytest.log

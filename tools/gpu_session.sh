# Development aid: the command list of one `gpurun -- 'bash tools/gpu_session.sh'` call (edited per session).
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/session_pytest.log 2>&1; tail -3 gpurun_out/session_pytest.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')") 2>&1 | tail -2
(time python bench.py > gpurun_out/session_bench.json) 2> gpurun_out/session_bench.err; tail -4 gpurun_out/session_bench.err; cut -c1-200 gpurun_out/session_bench.json

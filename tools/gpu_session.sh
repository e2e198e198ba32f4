# Development aid: the command list of one `gpurun -- 'bash tools/gpu_session.sh'` call (edited per session).
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/s60_pytest.log 2>&1; tail -3 gpurun_out/s60_pytest.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')") 2>&1 | tail -2
timeout 300 python tools/beagle_shim_bench.py | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['branch_gradient_call_sequence_ms'], d['device_ms_per_logl_plus_gradient'])"

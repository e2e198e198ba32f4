# Development aid: the command list of one `gpurun -- 'bash tools/gpu_session.sh'` call (edited per session).
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_beagle_shim.py -m gpu -x -q) > gpurun_out/s43_pytest.log 2>&1; tail -3 gpurun_out/s43_pytest.log
for k in 1 2; do echo "K=$k"; SBNB_BEAGLE_PATTERNS_PER_THREAD=$k timeout 300 python tools/beagle_shim_bench.py | tee gpurun_out/s43_shim_K$k.json | cut -c100-900; done

mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/s7_pytest.log 2>&1; tail -5 gpurun_out/s7_pytest.log
python tools/debug2.py 2>&1 | tail -1
TREES=296 bash tools/variants.sh "SBNB_EXTRA_SMEM=0" 2>&1 | tail -2

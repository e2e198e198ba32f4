# Development aid: the command list of one `gpurun -- 'bash tools/gpu_session.sh'` call (edited per session).
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_beagle_shim.py -m gpu -x -q) > gpurun_out/s55_pytest.log 2>&1; tail -3 gpurun_out/s55_pytest.log
timeout 300 python tools/beagle_shim_bench.py | tee gpurun_out/s55_shim.json | cut -c1-1500
for k in "UpdatePartialsPipelinedKernel<\(bool\)0" "UpdatePartialsPipelinedKernel<\(bool\)1" "EdgeDerivativesKernel"; do
  tag=$(echo "$k" | tr -cd 'A-Za-z01' | cut -c1-40)
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$k" -s 2 -c 1 -o gpurun_out/r02_shim_libsbn_$tag \
    python tools/beagle_shim_bench.py --repeats 2 > gpurun_out/r02_shim_ncu_$tag.log 2>&1
done

# Development aid: the command list of one `gpurun -- 'bash tools/gpu_session.sh'` call (edited per session).
mkdir -p gpurun_out
for k in "UpdatePartialsPipelinedKernel<\(bool\)0" "UpdatePartialsPipelinedKernel<\(bool\)1" "EdgeDerivativesKernel"; do
  tag=$(echo "$k" | tr -cd 'A-Za-z01' | cut -c1-40)
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$k" -s 2 -c 1 -o gpurun_out/r02_shim_$tag \
    python tools/beagle_shim_bench.py --repeats 2 > gpurun_out/r02_shim_ncu_$tag.log 2>&1
done

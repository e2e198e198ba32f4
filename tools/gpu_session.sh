mkdir -p gpurun_out
rm -f gpurun_out/ab.log
TREES=592 ROUNDS=2 bash tools/ab.sh new logl4

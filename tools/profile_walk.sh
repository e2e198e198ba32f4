#!/bin/bash
# ncu evidence for the tree-walk kernel (run under gpurun; writes gpurun_out/).
#   1. launch list of the bench command (per-launch durations, cold cache, serialised)
#   2. one --set full capture of the gradient walk on 148 trees
#   3. one --set full capture of the logL-only walk on 148 trees
# Summarise here with: python profiles/ncu_summary.py gpurun_out/<TAG>_treewalk.ncu-rep OUT.json 148 "note"
TAG=${1:-r02_v6}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-traffic-probe --quick > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"TreeWalkOeKernel<\(int\)4, \(int\)4, \(bool\)1" -c 1 -o gpurun_out/${TAG}_treewalk \
  python bench.py --_traffic-probe --trees 148 > gpurun_out/${TAG}_ncu.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"TreeWalkOeKernel<\(int\)4, \(int\)4, \(bool\)0" -c 1 -o gpurun_out/${TAG}_treewalk_logl \
  python bench.py --_traffic-probe --trees 148 > gpurun_out/${TAG}_ncu_logl.log 2>&1
ls -la gpurun_out

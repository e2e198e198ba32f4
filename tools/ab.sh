#!/bin/bash
# A/B timing of library builds on one GPU (development aid):  TREES=592 ROUNDS=2 tools/ab.sh base new ...
# runs bench.py --quick once per build per round, interleaved so that power-cap drift hits all alike.
mkdir -p gpurun_out
for round in $(seq 1 ${ROUNDS:-2}); do
  for name in "$@"; do
    lib=$PWD/libsbn_b200/lib/libsbn_$name.so
    [ "$name" = "new" ] && lib=$PWD/libsbn_b200/lib/libsbn_b200.so
    SBNB_LIBRARY=$lib python bench.py --trees ${TREES:-592} --steps 3 --warmup 3 --no-cpu-baseline --no-traffic-probe --quick 2>>gpurun_out/ab.err |
      python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name', round(d['value'],1), 'logL-only', round(d['log_likelihood_only']['value'],1), 'clock', d['clocks']['sm_mhz'])" | tee -a gpurun_out/ab.log
  done
done

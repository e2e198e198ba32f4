#!/bin/bash
# A/B timing of engine settings on one GPU (development aid): one short bench run per
# argument, each argument a list of environment assignments, e.g.
#   TREES=1024 tools/variants.sh "SBNB_TREES_IN_FLIGHT=8" "SBNB_TREES_IN_FLIGHT=64"
mkdir -p gpurun_out
for v in "$@"; do
  echo "== $v" | tee -a gpurun_out/variants.log
  env $v python bench.py --trees ${TREES:-296} --steps 2 --warmup 1 --no-cpu-baseline --no-traffic-probe --quick 2>>gpurun_out/variants.err |
    python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps({k:d[k] for k in ('value','ms_per_step','mean_log_likelihood')}), d['roofline']['kernel'], d['roofline']['kernel_ms'], 'logL-only', d['log_likelihood_only']['value'])" | tee -a gpurun_out/variants.log
done

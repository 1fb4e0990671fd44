#!/bin/bash
# Same-box A/B of an environment switch of the library (e.g. QNN_TC_NOSPLIT=1): alternating bench lines with and without it.
#   bash tools/ab_env.sh <tag> <VAR=value> [workloads...]      -> gpurun_out/r2/<tag>_abenv.txt
TAG=${1:-abenv}; SW=$2; shift 2
WLS=${@:-cfg2 dense}
OUT=gpurun_out/r2
mkdir -p $OUT
cd $GRAFT_REPO_ROOT
run() {  # $1 = workload, $2 = "VAR=value" or ""
  env $2 timeout 200 python bench.py --workload $1 --steps 50 --warmup 5 --no-secondary 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().splitlines()[-1]); print('%.5f %.5f' % (d['ms_per_step'], d['sustained']['ms_per_step_median']))"
}
for i in $(seq 1 ${ROUNDS:-3}); do
  for wl in $WLS; do
    echo "$wl run $i: default $(run $wl X_=1) | $SW $(run $wl $SW)" >> $OUT/${TAG}_abenv.txt
  done
done

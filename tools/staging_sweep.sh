#!/bin/bash
# Sweep of the output-staging / x-stage split of k_hamilton_tc's shared memory (QNN_TC_STAGING x QNN_TC_XSTAGES) on cfg 2,
# the dense north-star shape and the cfg 3 stack.  One JSON line per point -> gpurun_out/r2/<tag>_staging_sweep.json
TAG=${1:-s}
OUT=gpurun_out/r2
mkdir -p $OUT
cd $GRAFT_REPO_ROOT
: > $OUT/${TAG}_staging_sweep.json
for wl in cfg2 dense stack; do
  for st in 2 4 8; do
    for xs in 2 4; do
      QNN_TC_STAGING=$st QNN_TC_XSTAGES=$xs timeout 200 python bench.py --workload $wl --steps 50 --warmup 5 --no-secondary 2> $OUT/${TAG}_sweep.err | \
        python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps({'workload':'$wl','staging':$st,'x_stages':$xs,'ms_per_step':d['ms_per_step'],'sustained_median':d['sustained']['ms_per_step_median'],'parity_max_rel':d['parity']['max_rel']}))" >> $OUT/${TAG}_staging_sweep.json 2>> $OUT/${TAG}_sweep.err || echo "{\"workload\":\"$wl\",\"staging\":$st,\"x_stages\":$xs,\"error\":true}" >> $OUT/${TAG}_staging_sweep.json
    done
  done
done
echo done > $OUT/${TAG}_sweep_done

#!/bin/bash
# A/B of two library builds on ONE box: alternating runs of the cfg 2 / dense bench lines.  bash tools/ab_bench.sh <tag> <other .so>
TAG=${1:-ab}
OTHER=$2
OUT=gpurun_out/r2
mkdir -p $OUT
cd $GRAFT_REPO_ROOT
: > $OUT/${TAG}_ab.txt
for i in 1 2 3; do
  for wl in cfg2 dense; do
    a=$(timeout 200 python bench.py --workload $wl --steps 50 --warmup 5 --no-secondary 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().splitlines()[-1]); print(d['ms_per_step'], d['sustained']['ms_per_step_median'])")
    b=$(QNN_LIB_PATH=$OTHER timeout 200 python bench.py --workload $wl --steps 50 --warmup 5 --no-secondary 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().splitlines()[-1]); print(d['ms_per_step'], d['sustained']['ms_per_step_median'])")
    echo "$wl run $i: current $a | other $b" >> $OUT/${TAG}_ab.txt
  done
done

#!/bin/bash
# Same-box A/B of library builds: alternating runs of the cfg 2 / dense bench lines for the default build and every other
# build given.  bash tools/ab_bench.sh <tag> <name=path.so> [<name=path.so> ...]   (paths relative to the repo root)
TAG=${1:-ab}
shift
OUT=gpurun_out/r2
mkdir -p $OUT
cd $GRAFT_REPO_ROOT
: > $OUT/${TAG}_ab.txt
run() {  # $1 = workload, $2 = lib path or ""
  QNN_LIB_PATH=$2 timeout 200 python bench.py --workload $1 --steps 50 --warmup 5 --no-secondary 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().splitlines()[-1]); print('%.5f %.5f' % (d['ms_per_step'], d['sustained']['ms_per_step_median']))"
}
for i in $(seq 1 ${ROUNDS:-3}); do
  for wl in cfg2 dense; do
    line="$wl run $i: default $(run $wl '')"
    for v in "$@"; do line="$line | ${v%%=*} $(run $wl $GRAFT_REPO_ROOT/${v#*=})"; done
    echo "$line" >> $OUT/${TAG}_ab.txt
  done
done

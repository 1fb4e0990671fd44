#!/bin/bash
# Same-box experiments on variant builds under gpurun_variants/*/libqnn_*.so (tag -> gpurun_out/r2/<tag>_*):
# per-CTA traces of cfg 2 / dense, alternating bench lines, and (PARITY=name) the parity tests on one variant.
#   bash tools/diag_variants.sh <tag>
TAG=${1:-diag}
OUT=gpurun_out/r2
mkdir -p $OUT
cd $GRAFT_REPO_ROOT
for wl in cfg2 dense; do
  echo "######## default $wl" >> $OUT/${TAG}_diag.txt
  timeout 120 python tools/tc_trace.py $wl 2>&1 | head -34 >> $OUT/${TAG}_diag.txt
  for so in gpurun_variants/*/libqnn_*.so; do
    echo "######## $so $wl" >> $OUT/${TAG}_diag.txt
    QNN_LIB_PATH=$GRAFT_REPO_ROOT/$so timeout 120 python tools/tc_trace.py $wl 2>&1 | head -34 >> $OUT/${TAG}_diag.txt
  done
done
if [ -n "$PARITY" ]; then
  so=$(ls gpurun_variants/$PARITY/libqnn_*.so)
  QNN_LIB_PATH=$GRAFT_REPO_ROOT/$so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $OUT/${TAG}_pytest_$PARITY.log 2>&1
  echo "rc=$?" >> $OUT/${TAG}_pytest_$PARITY.log
fi
if [ -n "$BENCH" ]; then
  args=""
  for n in $BENCH; do args="$args $n=$(ls gpurun_variants/$n/libqnn_*.so)"; done
  ROUNDS=${ROUNDS:-2} bash tools/ab_bench.sh $TAG $args
fi

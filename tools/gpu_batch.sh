#!/bin/bash
# One gpurun call: build check, GPU test-suite, isolated risky tests, bench lines.  Usage: bash tools/gpu_batch.sh <tag>
TAG=${1:-g}
OUT=gpurun_out/r2
mkdir -p $OUT
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
if grep -q "rc=0" $OUT/${TAG}_pytest.log; then
  timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "rc=$?" >> $OUT/${TAG}_smoke.log
  timeout 400 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench20.json 2> $OUT/${TAG}_bench20.err
  timeout 400 python bench.py --steps 200 --warmup 10 --no-secondary > $OUT/${TAG}_bench200.json 2> $OUT/${TAG}_bench200.err
  timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/${TAG}_ref.json 2> $OUT/${TAG}_ref.err
else
  # first failure in isolation with a full trace
  timeout 300 python -m pytest tests -m gpu -x -q --tb=long 2>&1 | tail -n 150 > $OUT/${TAG}_fail.log
fi
timeout 200 python tools/bench_smallk.py > $OUT/${TAG}_smallk.json 2> $OUT/${TAG}_smallk.err
timeout 120 python tools/tc_trace.py cfg2 > $OUT/${TAG}_tc_trace_cfg2.log 2>&1
timeout 120 python tools/tc_trace.py dense > $OUT/${TAG}_tc_trace_dense.log 2>&1
if [ -f baseline/_ref/working_example.py ]; then
  timeout 900 python tools/run_working_example.py --model QDNN > $OUT/${TAG}_working_example_qdnn.json 2> $OUT/${TAG}_working_example_qdnn.log
  timeout 900 python tools/run_working_example.py --model QCNN > $OUT/${TAG}_working_example_qcnn.json 2> $OUT/${TAG}_working_example_qcnn.log
fi
echo done > $OUT/${TAG}_done

mkdir -p gpurun_out/r2
T=${TAG:-g27}
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2/${T}_pytest.log
ROUNDS=3 bash tools/ab_env.sh $T QNN_TC_W_EARLY=0 cfg2 dense
timeout 100 python tools/tc_trace.py cfg2 > gpurun_out/r2/${T}_tc_trace_cfg2.log 2>&1
for i in 1 2; do
QNN_TC_W_EARLY=0 timeout 400 python bench.py --steps 20 --warmup 3 --workload stack --no-secondary > gpurun_out/r2/${T}_stack_noearly_$i.json 2>/dev/null
timeout 400 python bench.py --steps 20 --warmup 3 --workload stack --no-secondary > gpurun_out/r2/${T}_stack_early_$i.json 2>/dev/null
done

mkdir -p gpurun_out/r2
T=${TAG:-g25}
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2/${T}_pytest.log
ROUNDS=2 bash tools/ab_bench.sh $T base=gpurun_variants/base/libqnn_base.so
timeout 100 python tools/tc_trace.py cfg2 > gpurun_out/r2/${T}_tc_trace_cfg2.log 2>&1
timeout 100 python tools/tc_trace.py dense > gpurun_out/r2/${T}_tc_trace_dense.log 2>&1

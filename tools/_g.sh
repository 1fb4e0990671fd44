mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2/g23_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2/g23_pytest.log
ROUNDS=2 bash tools/ab_env.sh g23 QNN_TC_NOSPLIT=1 cfg2 dense
timeout 100 python tools/tc_trace.py cfg2 > gpurun_out/r2/g23_tc_trace_cfg2.log 2>&1
timeout 100 python tools/tc_trace.py dense > gpurun_out/r2/g23_tc_trace_dense.log 2>&1

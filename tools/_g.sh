mkdir -p gpurun_out/r2
T=${TAG:-g31}
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2/${T}_pytest.log
ROUNDS=2 bash tools/ab_env.sh $T QNN_RAGGED_STREAM=0 stack train

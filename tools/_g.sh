mkdir -p gpurun_out/r2
T=${TAG:-g35}
QNN_LIB_PATH=$GRAFT_REPO_ROOT/gpurun_variants/dbg/libqnn_dbg.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "channels_last_conv2d or stack or timit or golden" > gpurun_out/r2/${T}_pytest_dbg.log 2>&1; rc=$?; echo "rc=$rc" >> gpurun_out/r2/${T}_pytest_dbg.log
if [ $rc -eq 0 ]; then
  timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2/${T}_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2/${T}_pytest.log
  ROUNDS=2 bash tools/ab_env.sh $T QNN_RAGGED_STREAM=0 stack train
fi

#!/bin/bash
TAG=${1:-s}
OUT=gpurun_out/r2
mkdir -p $OUT
cd $GRAFT_REPO_ROOT
SEL="test_kat_on_gpu or (test_conv_forward_vs_reference_golden and _tc_) or (test_dense_forward_vs_reference_golden and d_tc_) or (test_tensor_core_dgrad_conv1d_vs_oracle and causal_k2_relu) or (test_small_k_dense_vs_oracle)"
timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > $OUT/${TAG}_sanitizer_synccheck.log 2>&1
echo "rc=$?" >> $OUT/${TAG}_sanitizer_synccheck.log

#!/bin/bash
# One gpurun call (1 GPU): ncu launch lists, one full capture per hot kernel, compute-sanitizer logs.  bash tools/gpu_profile.sh <tag>
TAG=${1:-p}
OUT=gpurun_out/r2
mkdir -p $OUT
cd $GRAFT_REPO_ROOT
NCU="ncu --clock-control none"
# launch lists (cold-cache, serialised: compare SHARES)
timeout 600 $NCU --metrics gpu__time_duration.sum -c 300 --csv --log-file $OUT/${TAG}_launches_cfg2.csv python bench.py --steps 20 --warmup 3 --no-secondary > $OUT/${TAG}_l1.log 2>&1
timeout 600 $NCU --metrics gpu__time_duration.sum -c 300 --csv --log-file $OUT/${TAG}_launches_stack.csv python bench.py --steps 5 --warmup 3 --workload stack > $OUT/${TAG}_l2.log 2>&1
timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $OUT/${TAG}_launches_train.csv python bench.py --steps 3 --warmup 3 --workload train > $OUT/${TAG}_l3.log 2>&1
# full captures of the hot kernels (one launch each)
timeout 600 $NCU --set full --import-source on -k regex:k_hamilton_tc -s 6 -c 1 -o $OUT/${TAG}_tc_cfg2_full python bench.py --steps 5 --warmup 3 --no-secondary > $OUT/${TAG}_f1.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:k_hamilton_tc -s 6 -c 1 -o $OUT/${TAG}_tc_dense_full python bench.py --steps 5 --warmup 3 --workload dense > $OUT/${TAG}_f2.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:k_hamilton_wgrad -s 2 -c 1 -o $OUT/${TAG}_wgrad_full python bench.py --steps 3 --warmup 3 --workload train > $OUT/${TAG}_f3.log 2>&1
# compute-sanitizer on small shapes of every tensor-core kernel
SEL="test_kat_on_gpu or (test_conv_forward_vs_reference_golden and _tc_) or (test_dense_forward_vs_reference_golden and d_tc_) or (test_tensor_core_dgrad_conv1d_vs_oracle and causal_k2_relu) or (test_small_k_dense_vs_oracle) or test_split_last_round_is_bit_identical or (test_tensor_core_channels_last_conv2d_vs_oracle and (q9_ or q41_ or q12_ or q43_))"
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > $OUT/${TAG}_sanitizer_$tool.log 2>&1
  echo "rc=$?" >> $OUT/${TAG}_sanitizer_$tool.log
done
echo done > $OUT/${TAG}_done

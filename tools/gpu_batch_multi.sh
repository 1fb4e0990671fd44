#!/bin/bash
# One gpurun --gpus N call: multi-GPU tests and bench lines.  Usage: bash tools/gpu_batch_multi.sh <tag> <ngpus>
TAG=${1:-m}
N=${2:-2}
OUT=gpurun_out/r2
mkdir -p $OUT
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $OUT/${TAG}_multi_pytest.log 2>&1; echo "rc=$?" >> $OUT/${TAG}_multi_pytest.log
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 3 --workload train > $OUT/${TAG}_train_n$N.json 2> $OUT/${TAG}_train_n$N.err; echo "rc=$?" >> $OUT/${TAG}_train_n$N.err
QNN_BENCH_NCCL_IN_GRAPH=1 timeout 120 $TR bench.py --gpus $N --steps 20 --warmup 3 --workload train > $OUT/${TAG}_train_graph_n$N.json 2> $OUT/${TAG}_train_graph_n$N.err; echo "rc=$?" >> $OUT/${TAG}_train_graph_n$N.err
timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 3 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err; echo "rc=$?" >> $OUT/${TAG}_bench_n$N.err
timeout 300 $TR bench.py --gpus $N --steps 5 --warmup 3 --workload cfg5 > $OUT/${TAG}_cfg5_n$N.json 2> $OUT/${TAG}_cfg5_n$N.err; echo "rc=$?" >> $OUT/${TAG}_cfg5_n$N.err
nvidia-smi topo -m > $OUT/${TAG}_topo.log 2>&1
echo done > $OUT/${TAG}_done

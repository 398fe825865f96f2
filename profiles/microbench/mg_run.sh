#!/bin/bash
# Multi-GPU evidence run (one box, N GPUs visible): correctness checks at N, then the bench lines.
# usage: mg_run.sh <N> [tag]    outputs under gpurun_out/
N=${1:-8}
TAG=${2:-r2}
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
mkdir -p gpurun_out
$T --nproc-per-node $N --master-port 29531 tests/multi_gpu/shared_occupancy_check.py > gpurun_out/${TAG}_mg${N}_shared.log 2>&1; echo "shared rc=$?"
$T --nproc-per-node $N --master-port 29532 tests/multi_gpu/exchange_check.py > gpurun_out/${TAG}_mg${N}_xchg.log 2>&1; echo "xchg rc=$?"
for n in 1 2 4 8; do
  [ $n -gt $N ] && continue
  if [ $n -eq 1 ]; then
    python bench.py --gpus 1 --e2e-steps 100 > gpurun_out/${TAG}_weak_n1.json 2> gpurun_out/${TAG}_weak_n1.err
  else
    $T --nproc-per-node $n --master-port $((29540+n)) bench.py --gpus $n --e2e-steps 100 > gpurun_out/${TAG}_weak_n$n.json 2> gpurun_out/${TAG}_weak_n$n.err
  fi
  echo "weak n=$n rc=$?"
done
$T --nproc-per-node $N --master-port 29551 bench.py --gpus $N --scaling strong --e2e-steps 100 > gpurun_out/${TAG}_strong_n$N.json 2> gpurun_out/${TAG}_strong_n$N.err; echo "strong rc=$?"
$T --nproc-per-node $N --master-port 29552 bench.py --gpus $N --workload C5 --e2e-steps 40 > gpurun_out/${TAG}_c5_n$N.json 2> gpurun_out/${TAG}_c5_n$N.err; echo "c5 rc=$?"
M=$(( N < 4 ? N : 4 ))
$T --nproc-per-node $M --master-port 29553 bench.py --gpus $M --workload C4 > gpurun_out/${TAG}_c4_n$M.json 2> gpurun_out/${TAG}_c4_n$M.err; echo "c4 rc=$?"
tail -2 gpurun_out/${TAG}_mg${N}_shared.log gpurun_out/${TAG}_mg${N}_xchg.log

#!/bin/bash
# N-GPU sanity of the driver's bench command (config 3 split over the ranks):  tools/gpu_n4.sh TAG N
TAG=${1:-n4}; N=${2:-4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
source tools/summ.sh
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( time timeout 600 $TR --master-port 29518 bench.py --gpus $N --steps 5 --warmup 3 ) > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
summ $OUT/bench_n$N.json; tail -4 $OUT/bench_n$N.err

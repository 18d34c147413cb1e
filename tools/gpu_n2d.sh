#!/bin/bash
# the driver's N-GPU bench command on the current code:  tools/gpu_n2d.sh TAG N
TAG=${1:-n2d}; N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
source tools/summ.sh
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( time timeout 400 $TR --master-port 29518 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --stream-rounds 0 ) > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
summ $OUT/bench_n$N.json; tail -3 $OUT/bench_n$N.err

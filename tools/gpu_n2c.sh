#!/bin/bash
# N-GPU probe: host->device bandwidth per rank alone / together, then the config-3 bench with the
# wave pipeline on and off:  tools/gpu_n2c.sh TAG N
TAG=${1:-n2c}; N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( timeout 300 $TR --master-port 29511 tools/h2d_probe.py ) > $OUT/h2d.log 2>&1
grep "^rank" $OUT/h2d.log
nvidia-smi topo -m > $OUT/topo.txt 2>&1; head -12 $OUT/topo.txt
lscpu | grep -E "NUMA|Model name|^CPU\(s\)" > $OUT/cpu.txt; cat $OUT/cpu.txt
source tools/summ.sh
for e in SMB_PIPELINE=off SMB_PIPELINE=on; do
  echo "== bench --gpus $N $e"
  ( env $e timeout 600 $TR --master-port 29518 bench.py --gpus $N --steps 3 --warmup 2 --no-cpu-baseline --stream-rounds 0 ) > $OUT/bench_$e.json 2> $OUT/bench_$e.err
  summ $OUT/bench_$e.json; tail -2 $OUT/bench_$e.err
done

#!/bin/bash
# Multi-GPU check (gpurun --gpus N): NCCL correctness of both sharding modes and of the index
# broadcast, then the driver's own commands: bench.py at N (config 3, reads split over the ranks)
# and the reference arm under torchrun.
#   tools/gpu_n2.sh TAG N
TAG=${1:-n2}; N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/smi.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( time timeout 900 $TR --master-port 29517 tools/shard_nccl_check.py ) > $OUT/nccl_check.log 2>&1
grep -E "rank|PASS|Error|error" $OUT/nccl_check.log | head -20
source tools/summ.sh
echo "== bench --gpus $N (config 3, strong scaling)"
( time timeout 900 $TR --master-port 29518 bench.py --gpus $N --steps 3 --warmup 1 ) > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
summ $OUT/bench_n$N.json; tail -3 $OUT/bench_n$N.err
echo "== bench --gpus $N --shard contigs (8 contigs x 2.5 Mbp, 4000 reads)"
( time timeout 900 $TR --master-port 29519 bench.py --gpus $N --steps 1 --warmup 1 --shard contigs --ref-bp 20000000 --contigs 8 \
    --reads 4000 --mode default --no-cpu-baseline ) > $OUT/bench_n${N}_contigs.json 2> $OUT/bench_n${N}_contigs.err
summ $OUT/bench_n${N}_contigs.json; tail -3 $OUT/bench_n${N}_contigs.err
echo "== reference arm under torchrun"
( time timeout 900 $TR --master-port 29520 bench.py --impl reference --gpus $N --steps 2 --warmup 1 ) > $OUT/ref_n$N.json 2> $OUT/ref_n$N.err
tail -c 900 $OUT/ref_n$N.json; tail -3 $OUT/ref_n$N.err
ls -la $OUT

#!/bin/bash
# Two-GPU check: the NCCL contig-shard test, then bench.py at N=2 in both sharding modes.
TAG=${1:-two}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/smi.txt 2>&1
( time timeout 600 python -m pytest tests/test_gpu_contig_shard.py -m gpu -x -q ) > $OUT/pytest_shard.log 2>&1
tail -4 $OUT/pytest_shard.log
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 2 ) > $OUT/bench_n2.json 2> $OUT/bench_n2.err
tail -c 1200 $OUT/bench_n2.json; tail -2 $OUT/bench_n2.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 29518 bench.py --gpus 2 --steps 1 --warmup 1 --shard contigs --contigs 8 --reads 4000 ) \
    > $OUT/bench_n2_contigs.json 2> $OUT/bench_n2_contigs.err
tail -c 1500 $OUT/bench_n2_contigs.json; tail -2 $OUT/bench_n2_contigs.err

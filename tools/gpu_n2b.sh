#!/bin/bash
# Multi-GPU evidence (gpurun --gpus N): NCCL parity of both sharding modes, the driver's own
# commands at N (config 3 split over the ranks; the reference arm under torchrun), the contig-
# sharded bench at config scale, and one genome-scale contig-sharded run (own cloud parts).
#   tools/gpu_n2b.sh TAG N BIG_BP BIG_READS
TAG=${1:-n2}; N=${2:-2}; BIG_BP=${3:-400000000}; BIG_READS=${4:-1000}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,memory.total --format=csv > $OUT/smi.txt 2>&1
free -g > $OUT/host_mem.txt; nproc >> $OUT/host_mem.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( time timeout 600 $TR --master-port 29517 tools/shard_nccl_check.py ) > $OUT/nccl_check.log 2>&1
grep -E "^rank|PASS|Error|error" $OUT/nccl_check.log | head -20
source tools/summ.sh
echo "== bench --gpus $N (config 3, strong scaling)"
( time timeout 600 $TR --master-port 29518 bench.py --gpus $N --steps 3 --warmup 3 ) > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
summ $OUT/bench_n$N.json; tail -3 $OUT/bench_n$N.err
echo "== bench --gpus $N --shard contigs (config 3's reference, 4000 reads)"
( time timeout 600 $TR --master-port 29519 bench.py --gpus $N --steps 2 --warmup 1 --shard contigs --workload c3 \
    --reads 4000 --no-cpu-baseline --stream-rounds 0 ) > $OUT/bench_n${N}_contigs.json 2> $OUT/bench_n${N}_contigs.err
summ $OUT/bench_n${N}_contigs.json; tail -3 $OUT/bench_n${N}_contigs.err
echo "== reference arm under torchrun"
( time timeout 600 $TR --master-port 29520 bench.py --impl reference --gpus $N --steps 1 --warmup 1 ) > $OUT/ref_n$N.json 2> $OUT/ref_n$N.err
tail -c 700 $OUT/ref_n$N.json; tail -3 $OUT/ref_n$N.err
echo "== genome-scale contig-sharded run: $BIG_BP bp x 16 contigs, $BIG_READS reads"
( time timeout 900 $TR --master-port 29521 bench.py --gpus $N --steps 1 --warmup 1 --shard contigs --workload c3 \
    --ref-bp $BIG_BP --contigs 16 --reads $BIG_READS --no-cpu-baseline --stream-rounds 0 ) > $OUT/bench_big.json 2> $OUT/bench_big.err
summ $OUT/bench_big.json; tail -5 $OUT/bench_big.err
python - $OUT/bench_big.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print({k:d.get(k) for k in ('value','ms_per_step','setup_s','mapped_reads','truth_concordant_reads')}); print(d['config']); print(d['pipeline']['counters_per_step'])
PY
ls -la $OUT

#!/bin/bash
# suite + default bench + the same bench with the Morton-ordered index:  tools/gpu_r2i.sh TAG
TAG=${1:-r2i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q ) > $OUT/pytest_gpu.log 2>&1
echo "== pytest: $(grep -E 'passed|failed|error' $OUT/pytest_gpu.log | tail -1)"
grep -E "^(FAILED|ERROR)|^E  " $OUT/pytest_gpu.log | head -30
source tools/summ.sh
echo "== bench default (KD-ordered index)"
( timeout 900 python bench.py --steps 4 --warmup 3 ) > $OUT/bench.json 2> $OUT/bench.err
summ $OUT/bench.json; tail -3 $OUT/bench.err
echo "== bench SMB_INDEX=morton"
( SMB_INDEX=morton timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline ) > $OUT/bench_morton.json 2> $OUT/bench_morton.err
summ $OUT/bench_morton.json; tail -3 $OUT/bench_morton.err

#!/bin/bash
# suite + default bench + A/B legs through SMB_* options:  tools/gpu_r2j.sh TAG "ENV1" "ENV2" ...
TAG=${1:-r2j}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q ) > $OUT/pytest_gpu.log 2>&1
echo "== pytest: $(grep -E 'passed|failed|error' $OUT/pytest_gpu.log | tail -1)"
grep -E "^(FAILED|ERROR)|^E  " $OUT/pytest_gpu.log | cut -c1-300 | head -30
source tools/summ.sh
echo "== bench default"
( timeout 900 python bench.py --steps 4 --warmup 3 ) > $OUT/bench.json 2> $OUT/bench.err
summ $OUT/bench.json; tail -3 $OUT/bench.err
k=0
for e in "$@"; do
  k=$((k+1))
  echo "== bench $e"
  ( env $e timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline ) > $OUT/bench_ab$k.json 2> $OUT/bench_ab$k.err
  summ $OUT/bench_ab$k.json; tail -3 $OUT/bench_ab$k.err
done

#!/bin/bash
# suite + default bench + launch list of one config-2 pass + A/B legs:  tools/gpu_r2k.sh TAG "ENV1" ...
TAG=${1:-r2k}; shift
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
python - $OUT/bench.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('    slowest rounds', d['latency']['slowest_rounds'][:2])
PY
k=0
for e in "$@"; do
  k=$((k+1))
  echo "== bench $e"
  ( env $e timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline ) > $OUT/bench_ab$k.json 2> $OUT/bench_ab$k.err
  summ $OUT/bench_ab$k.json; tail -3 $OUT/bench_ab$k.err
done
echo "== launch list (config 2, one pass + 4 read-until rounds)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-config3 --stream-rounds 4 > $OUT/launches_bench.log 2>&1
python profiles/summarize_launches.py $OUT/launches.csv > $OUT/launches_summary.txt 2>&1
head -24 $OUT/launches_summary.txt

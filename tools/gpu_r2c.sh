#!/bin/bash
# r2c: the suite, the default bench, and the launch list (gpu__time_duration of every kernel of
# one config-2 step + the read-until rounds) the share-of-step numbers come from.
TAG=${1:-r2c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 1500 python -m pytest tests -m gpu -q ) > $OUT/pytest_gpu.log 2>&1
echo "== pytest: $(grep -E 'passed|failed|error' $OUT/pytest_gpu.log | tail -1)"
grep -E "^(FAILED|ERROR)|^E  " $OUT/pytest_gpu.log | head -30
source tools/summ.sh
echo "== bench (default: c2 + c3 leg)"
( timeout 900 python bench.py --steps 4 --warmup 2 ) > $OUT/bench.json 2> $OUT/bench.err
summ $OUT/bench.json; tail -3 $OUT/bench.err
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-config3 --stream-rounds 4 > $OUT/launches_bench.log 2>&1
python profiles/summarize_launches.py $OUT/launches.csv > $OUT/launches_summary.txt 2>&1
head -40 $OUT/launches_summary.txt
ls -la $OUT

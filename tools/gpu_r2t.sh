#!/bin/bash
# suite + default bench + the config 3 leg with the small staging:  tools/gpu_r2t.sh TAG
TAG=${1:-r2t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 1500 python -m pytest tests -m gpu -q ) > $OUT/pytest_gpu.log 2>&1
echo "== pytest: $(grep -E 'passed|failed|error' $OUT/pytest_gpu.log | tail -1)"
grep -E "^(FAILED|ERROR)|^E  " $OUT/pytest_gpu.log | cut -c1-300 | head -30
source tools/summ.sh
echo "== bench default"
( timeout 900 python bench.py --steps 4 --warmup 3 ) > $OUT/bench.json 2> $OUT/bench.err
summ $OUT/bench.json; tail -3 $OUT/bench.err
echo "== config 3 with SMB_STAGE=small"
( SMB_STAGE=small timeout 600 python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline --stream-rounds 0 ) > $OUT/bench_c3small.json 2> $OUT/bench_c3small.err
summ $OUT/bench_c3small.json; tail -2 $OUT/bench_c3small.err

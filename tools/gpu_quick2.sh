#!/bin/bash
# suite + default bench only:  tools/gpu_quick2.sh TAG
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 1500 python -m pytest tests -m gpu -q ) > $OUT/pytest_gpu.log 2>&1
echo "== pytest: $(grep -E 'passed|failed|error' $OUT/pytest_gpu.log | tail -1)"
grep -E "^(FAILED|ERROR)|^E  " $OUT/pytest_gpu.log | head -30
source tools/summ.sh
( timeout 900 python bench.py --steps ${STEPS:-4} --warmup 2 ) > $OUT/bench.json 2> $OUT/bench.err
summ $OUT/bench.json; tail -3 $OUT/bench.err

#!/bin/bash
# suite + default bench + read-until leg on config 3's reference:  tools/gpu_r2s.sh TAG
TAG=${1:-r2s}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 1500 python -m pytest tests -m gpu -q ) > $OUT/pytest_gpu.log 2>&1
echo "== pytest: $(grep -E 'passed|failed|error' $OUT/pytest_gpu.log | tail -1)"
grep -E "^(FAILED|ERROR)|^E  " $OUT/pytest_gpu.log | cut -c1-300 | head -30
source tools/summ.sh
echo "== bench default"
( timeout 900 python bench.py --steps 4 --warmup 3 ) > $OUT/bench.json 2> $OUT/bench.err
summ $OUT/bench.json; tail -3 $OUT/bench.err
echo "== read-until on config 3's reference"
( timeout 600 python bench.py --workload c3 --reads 6000 --steps 1 --warmup 1 --no-cpu-baseline ) > $OUT/bench_lat3.json 2> $OUT/bench_lat3.err
python - $OUT/bench_lat3.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value']/1e9,4), 'latency p50', round(d['latency']['p50'],3), d['latency']['slowest_rounds'][-1:])
PY

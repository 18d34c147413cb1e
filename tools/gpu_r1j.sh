#!/bin/bash
# r1j: GPU parity tests, default bench, A/B of the overlapped event blocks and of the
# warp-per-chunk event kernel on the read-until leg.
TAG=${1:-r1j}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log
( time timeout 600 python bench.py ) > $OUT/bench.json 2> $OUT/bench.err
tail -c 3500 $OUT/bench.json
tail -3 $OUT/bench.err
summ() {
python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); lat=d.get('latency') or {}
        print(round(d['value']/1e9,4), round(d['ms_per_step'],1), d['pipeline']['kernel_ms_per_step'], 'p50', lat.get('p50'), 'p90', lat.get('p90'), (lat.get('slowest_rounds') or [None])[0])
PY
}
echo "== overlap off"
( SMB_EVENTS_OVERLAP=0 timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --stream-rounds 0 ) > $OUT/bench_nooverlap.json 2> $OUT/bench_nooverlap.err
summ $OUT/bench_nooverlap.json
echo "== events thread-per-chunk only (latency leg)"
( SMB_EVENTS=thread timeout 300 python bench.py --reads 3000 --steps 1 --warmup 1 --no-cpu-baseline ) > $OUT/bench_thread.json 2> $OUT/bench_thread.err
summ $OUT/bench_thread.json
echo "== default, small read set (latency leg)"
( timeout 300 python bench.py --reads 3000 --steps 1 --warmup 1 --no-cpu-baseline ) > $OUT/bench_small.json 2> $OUT/bench_small.err
summ $OUT/bench_small.json

#!/bin/bash
# read-until leg on config 3's reference (12 Mbp x 16 contigs), per-tile DP pass on / off
TAG=${1:-lat3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for e in SMB_X=1 SMB_DP_TILES=0; do
  echo "== $e"
  ( env $e timeout 600 python bench.py --workload c3 --reads 6000 --steps 1 --warmup 1 --no-cpu-baseline ) > $OUT/bench_$e.json 2> $OUT/bench_$e.err
  python - $OUT/bench_$e.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value']/1e9,4), 'latency p50', round(d['latency']['p50'],3), d['latency']['slowest_rounds'][:1])
PY
  tail -2 $OUT/bench_$e.err
done

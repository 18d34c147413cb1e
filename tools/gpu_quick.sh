#!/bin/bash
# One short gpurun call: GPU parity tests + the default bench + bench variants selected by
# environment toggles.  tools/gpu_quick.sh TAG "VAR=val VAR2=val" "VAR=val" ...
TAG=${1:-q}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log
( time timeout 600 python bench.py ) > $OUT/bench.json 2> $OUT/bench.err
tail -c 2500 $OUT/bench.json
i=0
for V in "$@"; do
  i=$((i+1))
  echo "== variant $i: $V"
  ( env $V timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --stream-rounds 0 ) > $OUT/bench_v$i.json 2> $OUT/bench_v$i.err
  echo "# $V" >> $OUT/bench_v$i.json
  python - $OUT/bench_v$i.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); c=d['pipeline']['counters_per_step']; print(round(d['value']/1e9,4), round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['pipeline']['kernel_ms_per_step'].items()}, 'steps', c.get('steps'), 'seg', c.get('seg_sort_steps'), 'part', c.get('part_sort_steps'))
PY
done

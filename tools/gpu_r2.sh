#!/bin/bash
# One round-2 GPU call: the whole -m gpu suite, the default bench (config 2 + the config 3 leg +
# CPU reference + concordance), and a full ncu capture of the search kernels of one config-2 step.
#   tools/gpu_r2.sh TAG [skip-ncu]
TAG=${1:-r2b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
echo "== pytest: $(grep -E 'passed|failed|error' $OUT/pytest_gpu.log | tail -1)"
grep -E "^(FAILED|ERROR)|Error|assert " $OUT/pytest_gpu.log | head -20
source tools/summ.sh
echo "== bench (default: c2 + c3 leg)"
( timeout 900 python bench.py --steps 4 --warmup 2 ) > $OUT/bench.json 2> $OUT/bench.err
summ $OUT/bench.json; tail -3 $OUT/bench.err
if [ -z "$2" ]; then
  echo "== ncu: search kernels of one c2 step"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_search_lean|k_radius_search" -s 40 -c 40 \
      -o $OUT/search_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-config3 --stream-rounds 0 \
      > $OUT/search_full_bench.log 2>&1
  ncu -i $OUT/search_full.ncu-rep --page raw --csv > $OUT/search_full_raw.csv 2>/dev/null
  python profiles/ncu_summary.py < $OUT/search_full_raw.csv > $OUT/search_full_summary.md 2>&1
  head -12 $OUT/search_full_summary.md
  SZ=$(stat -c %s $OUT/search_full.ncu-rep 2>/dev/null || echo 0)
  if [ "$SZ" -gt 40000000 ]; then rm -f $OUT/search_full.ncu-rep; fi
fi
ls -la $OUT

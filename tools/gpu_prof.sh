#!/bin/bash
# ncu --set full capture of selected kernels on a reduced read set, summarised ON THE BOX so the
# (large) .ncu-rep need not travel:  tools/gpu_prof.sh TAG 'REGEX' SKIP COUNT [hot-line kernel ...]
# Outputs under gpurun_out/TAG/: prof_raw.csv (ncu raw page), summary.md, hot_<kernel>.txt.
TAG=${1:-prof}; REGEX=${2:-k_radius_search}; SKIP=${3:-20}; COUNT=${4:-8}
shift 4
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -s $SKIP -c $COUNT \
    -o $OUT/prof python bench.py --reads 4000 --steps 1 --warmup 1 --no-cpu-baseline --stream-rounds 0 \
    > $OUT/prof_bench.log 2>&1
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/prof_raw.csv 2>/dev/null
python profiles/ncu_summary.py < $OUT/prof_raw.csv > $OUT/summary.md 2>&1
for K in "$@"; do
  python profiles/hot_lines.py $OUT/prof.ncu-rep "$K" 0 40 > $OUT/hot_$K.txt 2>&1
done
SZ=$(stat -c %s $OUT/prof.ncu-rep 2>/dev/null || echo 0)
if [ "$SZ" -gt 30000000 ]; then rm -f $OUT/prof.ncu-rep; fi
ls -la $OUT
cat $OUT/summary.md

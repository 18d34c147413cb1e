#!/bin/bash
# ncu --set full capture of selected kernels on a reduced read set: tools/gpu_ncu.sh TAG REGEX [SKIP] [COUNT]
TAG=${1:-ncu}; REGEX=${2:-k_radius_search}; SKIP=${3:-6}; COUNT=${4:-3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -s $SKIP -c $COUNT \
    -o $OUT/prof python bench.py --reads 4000 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/prof_bench.log 2>&1
ls -la $OUT

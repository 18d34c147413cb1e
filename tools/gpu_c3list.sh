#!/bin/bash
# ncu launch list of config 3 (30 000 reads of it) on one GPU:  tools/gpu_c3list.sh TAG
OUT=gpurun_out/${1:-c3list}
mkdir -p $OUT
timeout 800 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv \
    python bench.py --workload c3 --reads 30000 --steps 1 --warmup 1 --no-cpu-baseline --stream-rounds 0 > $OUT/launches_bench.log 2>&1
python profiles/summarize_launches.py $OUT/launches.csv > $OUT/launches_summary.txt 2>&1
head -22 $OUT/launches_summary.txt
tail -c 600 $OUT/launches_bench.log

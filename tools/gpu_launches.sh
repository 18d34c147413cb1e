#!/bin/bash
# ncu launch list (per-launch gpu__time_duration of every kernel of one reduced bench run) and one
# ncu --set full capture of the hot kernels: tools/gpu_launches.sh TAG
TAG=${1:-launches}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file $OUT/launches.csv python bench.py --reads 4000 --steps 1 --warmup 1 --no-cpu-baseline --stream-rounds 0 \
    > $OUT/launches_bench.log 2>&1
python profiles/summarize_launches.py $OUT/launches.csv > $OUT/launches_summary.txt 2>&1
head -32 $OUT/launches_summary.txt
bash tools/gpu_prof.sh $TAG 'k_radius_search|k_chain_dp|k_chain_prep|k_sel_trace|k_sel_final|k_part_sort|k_ev_features' 70 14 \
    k_radius_search k_chain_dp k_part_sort k_chain_prep > $OUT/prof.log 2>&1
tail -18 $OUT/summary.md

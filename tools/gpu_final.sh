#!/bin/bash
# Full single-GPU evidence run: the suite, the default bench (driver's steps), the launch list of
# one config-2 pass, and `ncu --set full` captures of the hot kernels (summarised on the box).
#   tools/gpu_final.sh TAG
TAG=${1:-r2z}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q ) > $OUT/pytest_gpu.log 2>&1
echo "== pytest: $(grep -E 'passed|failed|error' $OUT/pytest_gpu.log | tail -1)"
grep -E "^(FAILED|ERROR)|^E  " $OUT/pytest_gpu.log | head -30
source tools/summ.sh
echo "== bench (default)"
( timeout 1200 python bench.py --steps ${STEPS:-8} --warmup 3 ) > $OUT/bench.json 2> $OUT/bench.err
summ $OUT/bench.json; tail -3 $OUT/bench.err
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-config3 --stream-rounds 4 > $OUT/launches_bench.log 2>&1
python profiles/summarize_launches.py $OUT/launches.csv > $OUT/launches_summary.txt 2>&1
head -14 $OUT/launches_summary.txt
echo "== ncu --set full: search kernels of one pass"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_search_lean|k_radius_search" -s 40 -c 40 \
    -o $OUT/search_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-config3 --stream-rounds 0 \
    > $OUT/search_full_bench.log 2>&1
ncu -i $OUT/search_full.ncu-rep --page raw --csv > $OUT/search_full_raw.csv 2>/dev/null
python profiles/ncu_summary.py < $OUT/search_full_raw.csv > $OUT/search_full_summary.md 2>&1
head -6 $OUT/search_full_summary.md
echo "== ncu --set full: sort / chain kernels (first four steps of the second pass)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_part_sort|k_chain_prep|k_chain_dp|k_sel_trace|k_sel_commit" -s 100 -c 20 \
    -o $OUT/hot_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-config3 --stream-rounds 0 \
    > $OUT/hot_full_bench.log 2>&1
ncu -i $OUT/hot_full.ncu-rep --page raw --csv > $OUT/hot_full_raw.csv 2>/dev/null
python profiles/ncu_summary.py < $OUT/hot_full_raw.csv > $OUT/hot_full_summary.md 2>&1
head -12 $OUT/hot_full_summary.md
for f in $OUT/search_full.ncu-rep $OUT/hot_full.ncu-rep; do
  SZ=$(stat -c %s $f 2>/dev/null || echo 0); if [ "$SZ" -gt 25000000 ]; then rm -f $f; fi
done
ls -la $OUT

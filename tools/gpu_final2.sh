#!/bin/bash
# Full single-GPU evidence run: the suite, the driver's own bench command (both arms), the launch
# list of one config-2 pass, and `ncu --set full` captures of the hot kernels (summarised on the box).
#   tools/gpu_final2.sh TAG
TAG=${1:-r2z}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q ) > $OUT/pytest_gpu.log 2>&1
echo "== pytest: $(grep -E 'passed|failed|error' $OUT/pytest_gpu.log | tail -1)"
grep -E "^(FAILED|ERROR)|^E  " $OUT/pytest_gpu.log | cut -c1-300 | head -30
source tools/summ.sh
echo "== smoke"
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > $OUT/smoke.log 2>&1; tail -4 $OUT/smoke.log
echo "== bench (the driver's command)"
( time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $OUT/bench.json 2> $OUT/bench.err
summ $OUT/bench.json; tail -4 $OUT/bench.err
echo "== reference arm (the driver's command)"
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > $OUT/ref.json 2> $OUT/ref.err
tail -c 600 $OUT/ref.json; tail -4 $OUT/ref.err
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-config3 --stream-rounds 4 > $OUT/launches_bench.log 2>&1
python profiles/summarize_launches.py $OUT/launches.csv > $OUT/launches_summary.txt 2>&1
head -16 $OUT/launches_summary.txt
echo "== ncu --set full: search kernels of one pass"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_search_lean|k_radius_search" -s 40 -c 40 \
    -o $OUT/search_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-config3 --stream-rounds 0 \
    > $OUT/search_full_bench.log 2>&1
ncu -i $OUT/search_full.ncu-rep --page raw --csv > $OUT/search_full_raw.csv 2>/dev/null
python profiles/ncu_summary.py < $OUT/search_full_raw.csv > $OUT/search_full_summary.md 2>&1
head -5 $OUT/search_full_summary.md
python tools/make_search_traffic.py $OUT/search_full_raw.csv $OUT/search_traffic.json > /dev/null 2>&1; cat $OUT/search_traffic.json | head -12
python profiles/hot_lines.py $OUT/search_full.ncu-rep k_search_lean 4 40 > $OUT/hot_k_search_lean.txt 2>&1
echo "== ncu --set full: sort / chain kernels (four steps of the second pass)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_part_sort|k_chain_prep|k_chain_dp|k_sel_trace|k_sel_commit" -s 100 -c 20 \
    -o $OUT/hot_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-config3 --stream-rounds 0 \
    > $OUT/hot_full_bench.log 2>&1
ncu -i $OUT/hot_full.ncu-rep --page raw --csv > $OUT/hot_full_raw.csv 2>/dev/null
python profiles/ncu_summary.py < $OUT/hot_full_raw.csv > $OUT/hot_full_summary.md 2>&1
head -8 $OUT/hot_full_summary.md
for K in k_part_sort k_chain_prep k_chain_dp; do
  python profiles/hot_lines.py $OUT/hot_full.ncu-rep $K 0 30 > $OUT/hot_$K.txt 2>&1
done
for f in $OUT/search_full.ncu-rep $OUT/hot_full.ncu-rep; do
  SZ=$(stat -c %s $f 2>/dev/null || echo 0); if [ "$SZ" -gt 25000000 ]; then rm -f $f; fi
done
ls -la $OUT

#!/bin/bash
# One gpurun call: GPU parity tests, the default bench (both arms), an ncu launch list and one
# ncu --set full capture of the hot kernels.  Outputs under gpurun_out/<tag>/.
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
( time timeout 600 python bench.py ) > $OUT/bench.json 2> $OUT/bench.err
tail -c 3000 $OUT/bench.json
if [ "${SKIP_REF:-0}" != "1" ]; then
( time timeout 400 python bench.py --impl reference --steps 2 --warmup 1 ) > $OUT/bench_ref.json 2> $OUT/bench_ref.err
tail -c 1500 $OUT/bench_ref.json
fi
if [ "${SKIP_NCU:-0}" != "1" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file $OUT/launches.csv python bench.py --reads 4000 --steps 1 --warmup 1 --no-cpu-baseline --stream-rounds 0 \
    > $OUT/launches_bench.log 2>&1
python profiles/summarize_launches.py $OUT/launches.csv > $OUT/launches_summary.txt 2>&1
head -30 $OUT/launches_summary.txt
bash tools/gpu_prof.sh $TAG 'k_radius_search|k_chain_dp|k_chain_prep|k_sel_trace|k_sel_final|k_part_sort|k_ev_features' 70 14 \
    k_radius_search k_chain_dp k_part_sort k_chain_prep k_ev_features > $OUT/prof.log 2>&1
tail -20 $OUT/summary.md
fi

#!/bin/bash
# First GPU call of the next round (about 5 GPU-minutes): what round 1 left unvalidated.
#   1. the parity suite as is, and again with the experimental binary16 box test (SMB_BOX=half):
#      the radius hit-set tests decide whether its per-query slack is conservative in practice;
#   2. the fallback sorts (SMB_SORT=entry / global) through the whole suite -- they are not
#      covered by a test of their own;
#   3. bench A/B: default vs SMB_BOX=half;
#   4. config 3 of BASELINE.json (12 Mbp x 16 contigs, default stop rules): the first 7-level
#      index (needs the > 48 KB shared-memory opt-in added at the end of round 1).
TAG=${1:-r2a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run_suite() {  # name, env...
  local name=$1; shift
  ( time env "$@" timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_$name.log 2>&1
  echo "== pytest [$name] $(tail -4 $OUT/pytest_$name.log | grep -E 'passed|failed|error' | tail -1)"
}
run_suite default SMB_NOTHING=1
run_suite box_half SMB_BOX=half
run_suite sort_entry SMB_SORT=entry
run_suite sort_global SMB_SORT=global
summ() {
python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value']/1e9,4), round(d['ms_per_step'],1), {k:round(v,1) for k,v in d['pipeline']['kernel_ms_per_step'].items()}, d['mapped_reads'], d['truth_concordant_reads'], d['pipeline']['counters_per_step']['linked'])
PY
}
for V in "SMB_NOTHING=1" "SMB_BOX=half"; do
  echo "== bench [$V]"
  ( env $V timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --stream-rounds 0 ) > $OUT/bench_$V.json 2> $OUT/bench_$V.err
  summ $OUT/bench_$V.json
done
echo "== bench config 3 (12 Mbp x 16 contigs, 100 000 reads, default stop rules)"
( timeout 600 python bench.py --ref-bp 12000000 --contigs 16 --reads 100000 --mode default --steps 1 --warmup 1 \
    --no-cpu-baseline --stream-rounds 0 ) > $OUT/bench_c3.json 2> $OUT/bench_c3.err
summ $OUT/bench_c3.json; tail -2 $OUT/bench_c3.err
echo "== CUDA vs oracle on noisy, multi-contig, spiked reads"
( timeout 600 python tools/check_noisy_gpu.py ) > $OUT/noisy.log 2>&1; tail -3 $OUT/noisy.log

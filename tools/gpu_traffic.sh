#!/bin/bash
# DRAM traffic of the roofline kernel on the DEFAULT bench workload: one `ncu --set full` capture of
# the 20 k_radius_search launches of one whole step (= one map pass over the 20 000 reads), summarised
# on the box.  Writes gpurun_out/TAG/search_full_raw.csv, search_full_summary.md and
# search_traffic.json (copy the latter to profiles/ -- bench.py reports it as roofline.traffic).
TAG=${1:-traffic}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_radius_search -s 20 -c 20 \
    -o $OUT/search_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --stream-rounds 0 \
    > $OUT/search_full_bench.log 2>&1
ncu -i $OUT/search_full.ncu-rep --page raw --csv > $OUT/search_full_raw.csv 2>/dev/null
python profiles/ncu_summary.py < $OUT/search_full_raw.csv > $OUT/search_full_summary.md 2>&1
python - $OUT/search_full_raw.csv $OUT/search_traffic.json <<'PY'
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def col(name):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
    out = []
    for r in rows[2:]:
        out.append(float(r[ix[name]].replace(",", "")) * scale.get(units[ix[name]], 1))
    return out
rd, wr, t = col("dram__bytes_read.sum"), col("dram__bytes_write.sum"), col("gpu__time_duration.sum")
n = len(rd)
json.dump({"kernel": "k_radius_search", "launches": n,
           "dram_bytes_per_launch": (sum(rd) + sum(wr)) / n,
           "dram_read_bytes_per_launch": sum(rd) / n, "dram_write_bytes_per_launch": sum(wr) / n,
           "ncu_ms_per_launch": 1e3 * sum(t) / n,
           "how": "ncu --set full --clock-control none, the 20 launches of one whole step of the default "
                  "bench.py workload (launches 21..40 = second map pass); dram__bytes_read.sum + "
                  "dram__bytes_write.sum averaged per launch"}, open(sys.argv[2], "w"), indent=1)
print(open(sys.argv[2]).read())
PY
python profiles/hot_lines.py $OUT/search_full.ncu-rep k_radius_search 0 40 > $OUT/hot_k_radius_search.txt 2>&1
SZ=$(stat -c %s $OUT/search_full.ncu-rep 2>/dev/null || echo 0)
if [ "$SZ" -gt 30000000 ]; then rm -f $OUT/search_full.ncu-rep; fi
ls -la $OUT

#!/bin/bash
# launch list of read-until rounds on config 3's reference
OUT=gpurun_out/${1:-lat3l}
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv \
    python bench.py --workload c3 --reads 3000 --steps 1 --warmup 1 --no-cpu-baseline --stream-rounds 6 > $OUT/launches_bench.log 2>&1
python - $OUT/launches.csv <<'PY'
import csv,sys,re,collections
lines=[l for l in open(sys.argv[1]) if not l.startswith("==")]
rows=[r for r in csv.DictReader(lines) if r.get("Metric Name")=="gpu__time_duration.sum"]
# the last 6*~30 launches are the stream rounds: print per-kernel time of the last 200 launches
tail=rows[-200:]
tot=collections.defaultdict(float); cnt=collections.Counter()
for r in tail:
    name=re.sub(r"<.*","",re.sub(r"\(.*","",r["Kernel Name"]))[:50]
    v=float(r["Metric Value"].replace(",",""))*{"ns":1e-3,"us":1,"ms":1e3}[r["Metric Unit"]]
    tot[name]+=v; cnt[name]+=1
for k,v in sorted(tot.items(), key=lambda x:-x[1])[:16]: print(f"{v:10.1f} us {cnt[k]:4d} x {v/cnt[k]:8.1f}  {k}")
PY

#!/usr/bin/env python
"""profiles/search_traffic.json from the raw page of one `ncu --set full` capture of the search
kernels of ONE whole map pass of the default bench workload (lean + general kernel per step):
  python tools/make_search_traffic.py search_full_raw.csv out.json"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
         "inst": 1, "%": 1}


def col(name, r):
    return float(r[ix[name]].replace(",", "")) * scale.get(units[ix[name]], 1)


body = [r for r in rows[2:] if len(r) == len(hdr)]
lean = [r for r in body if "k_search_lean" in r[ix["Kernel Name"]]]
gen = [r for r in body if "k_radius_search" in r[ix["Kernel Name"]]]
steps = max(len(lean), 1)
rd = sum(col("dram__bytes_read.sum", r) for r in body)
wr = sum(col("dram__bytes_write.sum", r) for r in body)
t = sum(col("gpu__time_duration.sum", r) for r in body)
tg = sum(col("gpu__time_duration.sum", r) for r in gen)
inst = sum(col("smsp__inst_executed.sum", r) for r in body) if "smsp__inst_executed.sum" in ix else None
issue = [col("sm__inst_issued.avg.pct_of_peak_sustained_active", r) for r in lean] if "sm__inst_issued.avg.pct_of_peak_sustained_active" in ix else []
l2 = [col("lts__t_sector_hit_rate.pct", r) for r in lean] if "lts__t_sector_hit_rate.pct" in ix else []
out = {"kernel": "k_search_lean + k_radius_search (overflow queries)", "launches": steps,
       "dram_bytes_per_launch": (rd + wr) / steps, "dram_read_bytes_per_launch": rd / steps,
       "dram_write_bytes_per_launch": wr / steps, "ncu_ms_per_launch": 1e3 * t / steps,
       "how": "ncu --set full --clock-control none, the search launches (lean + general kernel each) of one whole "
              "pass of the default bench.py workload; dram__bytes_read.sum + dram__bytes_write.sum summed per step",
       "ncu": {"issue_slots_pct": sum(issue) / len(issue) if issue else None,
               "l2_hit_pct": sum(l2) / len(l2) if l2 else None,
               "warp_instructions_per_launch": inst / steps if inst else None,
               "general_kernel_share_of_time": tg / t if t else None},
       "limiter": "instruction issue: index records come from L1/L2 (queries in Morton order), DRAM traffic is "
                  "below the algorithmic bytes"}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))

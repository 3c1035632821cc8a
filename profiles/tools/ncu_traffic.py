#!/usr/bin/env python
"""ncu_traffic.py <out traffic.json> <workload>=<raw.csv>=<summary path> ...: collects, per workload, what one `ncu --set full`
capture of its force kernel measured (DRAM bytes of the launch, FP32-pipe / issue utilisation) into the file bench.py reads
for roofline.traffic / roofline.ncu."""
import csv
import json
import sys


def num(s):
    return float(s.replace(",", ""))


def unit_scale(u):
    u = u.lower()
    return {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "us": 1.0, "usecond": 1.0, "ns": 1e-3, "nsecond": 1e-3, "ms": 1e3,
            "msecond": 1e3}.get(u, 1.0)


out = {"note": "dram__bytes_read.sum + dram__bytes_write.sum of ONE k_force launch, ncu --set full --clock-control none "
               "(profiles/tools/run_r2_n.sh); bench.py copies the figure of its workload into roofline.traffic"}
for arg in sys.argv[2:]:
    wl, raw, summary = arg.split("=")
    rows = list(csv.reader(open(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}

    def get(k):
        return num(d[k][0]) * unit_scale(d[k][1])
    out[wl] = {"bytes": int(get("dram__bytes_read.sum") + get("dram__bytes_write.sum")),
               "kernel": d["Kernel Name"][0].split("(")[0].replace("void <unnamed>::", "").replace("(int)", "").replace("(bool)", ""),
               "source": summary,
               "fma_pipe_cycles_active_pct": round(num(d["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"][0]), 1),
               "issue_active_per_cycle": round(num(d["smsp__issue_active.avg.per_cycle_active"][0]), 2),
               "sm_cycles_active_over_elapsed": round(num(d["sm__cycles_active.avg"][0]) / num(d["sm__cycles_elapsed.avg"][0]), 3),
               "warp_instructions": int(num(d["smsp__inst_executed.sum"][0])),
               "kernel_us_under_ncu": round(get("gpu__time_duration.sum"), 2)}
json.dump(out, open(sys.argv[1], "w"), indent=1)
print(json.dumps(out, indent=1))

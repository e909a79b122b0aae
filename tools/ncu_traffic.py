"""Summarise an ncu CSV (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum per launch of one network
call, tools/one_call.py) into profiles/traffic.json + a per-kernel markdown table.
Usage: python tools/ncu_traffic.py gpurun_out/traffic.csv profiles/traffic.json profiles/rNN_launches.md"""
import collections
import csv
import json
import re
import sys

src, out_json, out_md = sys.argv[1:4]
lines = [l for l in open(src) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("ccedit::", "").strip()
    name = re.sub(r"<.*", "", name)
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    a = agg.setdefault(name, {"ids": set(), "read": 0.0, "write": 0.0, "ns": 0.0})
    a["ids"].add(row["ID"])
    m = row["Metric Name"]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1,
             "usecond": 1e3, "msecond": 1e6}.get(unit, 1)
    if m == "dram__bytes_read.sum":
        a["read"] += v * scale
    elif m == "dram__bytes_write.sum":
        a["write"] += v * scale
    elif m == "gpu__time_duration.sum":
        a["ns"] += v * scale
tot_ns = sum(a["ns"] for a in agg.values())
kern = {}
with open(out_md, "w") as f:
    f.write("| kernel | launches | ms (ncu, serialised) | share | DRAM read MB | DRAM write MB | DRAM bytes / launch |\n|---|---|---|---|---|---|---|\n")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        n = len(a["ids"])
        kern[name] = {"launches": n, "bytes_per_launch": (a["read"] + a["write"]) / n, "ms": a["ns"] / 1e6}
        f.write(f"| {name} | {n} | {a['ns'] / 1e6:.3f} | {100 * a['ns'] / max(tot_ns, 1):.1f} % | {a['read'] / 1e6:.1f} | "
                f"{a['write'] / 1e6:.1f} | {(a['read'] + a['write']) / n / 1e6:.2f} MB |\n")
    f.write(f"\ntotal {tot_ns / 1e6:.2f} ms over {sum(len(a['ids']) for a in agg.values())} launches\n")
json.dump({"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none, "
                     "one network call at CFG batch 2 x 17 x 64 x 96 (tools/one_call.py)", "kernels": kern},
          open(out_json, "w"), indent=1)
print(open(out_md).read())

"""DRAM bytes per launch per kernel from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv`
over one bench.py run (identical steps): python tools/traffic_summary.py gpurun_out/traffic.csv profiles/r02_final_traffic.json"""
import collections, csv, json, sys

src, dst = sys.argv[1], sys.argv[2]
rows = list(csv.DictReader(l for l in open(src) if not l.startswith("==")))
tot = collections.OrderedDict()
for r in rows:
    name = r["Kernel Name"].split("(")[0].replace("void ", "").split("<")[0]
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"].lower()
    v *= {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)
    e = tot.setdefault(name, {"bytes": 0.0, "ids": set()})
    e["bytes"] += v
    e["ids"].add(r["ID"])
out = {}
for k, e in tot.items():
    n = len(e["ids"])
    out[k] = {"dram_bytes_per_launch": e["bytes"] / n, "launches": n, "dram_bytes_total": e["bytes"],
              "captured": f"average over all {n} launches of `bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-secondary --no-job` (4 identical steps of 640 trees + the 40-tree CRC fit): {e['bytes'] / 1e9:.1f} GB in total",
              "source": f"{src}: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none"}
json.dump(out, open(dst, "w"), indent=1)
for k, e in sorted(out.items(), key=lambda kv: -kv[1]["dram_bytes_total"])[:12]:
    print(f"{k:28s} {e['launches']:5d} launches {e['dram_bytes_total'] / 1e9:9.2f} GB total {e['dram_bytes_per_launch'] / 1e6:10.1f} MB / launch")

"""Turn gpurun_out ncu artefacts into the tracked summaries under profiles/.

    python tools/profile_summary.py launches gpurun_out/launches.csv profiles/r01_launches.md
    python tools/profile_summary.py kernel   gpurun_out/prof_big.ncu-rep profiles/r01_k_grow_big.md [traffic.json key]
"""
import collections
import csv
import json
import subprocess
import sys


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return 0.0


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    seq = []
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0].replace("void ", "")
        v, unit = num(row["Metric Value"]), row["Metric Unit"]
        ms = v / 1e6 if unit.startswith("n") else (v / 1e3 if unit.startswith("u") else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += ms
        seq.append((name, ms))
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` — per-launch times are "
                "cold-cache and serialised: compare SHARES.\n\n| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {a[0]} | {a[1]:.3f} | {a[1] / tot:.3f} |\n")
        f.write(f"\ntotal {tot:.3f} ms over {len(seq)} launches\n")
    print(open(dst).read())


KEYS = ["Duration", "DRAM Throughput", "Memory Throughput", "L1/TEX Hit Rate", "L2 Hit Rate", "Executed Ipc Active", "Issue Slots Busy",
        "Registers Per Thread", "Theoretical Occupancy", "Achieved Occupancy", "Grid Size", "Block Size", "Dynamic Shared Memory Per Block"]
RAW = ["dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed_op_shared_atom.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum",
       "lts__t_sectors_op_read.sum", "gpu__time_duration.sum", "smsp__inst_executed.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
       # the LSU data pipe (wavefronts as a share of its peak rate) and its shared-memory parts; issue slots
       "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum.pct_of_peak_sustained_elapsed", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
       "sm__inst_issued.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def kernel(src, dst, traffic_json=None):
    det = subprocess.run(["ncu", "-i", src, "--page", "details"], capture_output=True, text=True).stdout
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    out = [f"# ncu --set full summary ({src})\n"]
    launches_ = []
    cur = None
    for l in det.splitlines():
        s = l.strip()
        if "Context" in s and "Stream" in s and "(" in s:
            cur = {"name": s.split(" Context")[0]}
            launches_.append(cur)
        elif cur is not None:
            for k in KEYS:
                if s.startswith(k + " ") and k not in cur:
                    cur[k] = " ".join(s.split()[len(k.split()):])
    out.append("| launch | " + " | ".join(KEYS) + " |\n|---|" + "---|" * len(KEYS))
    for i, c in enumerate(launches_):
        out.append(f"| {i}: `{c['name'][:60]}` | " + " | ".join(c.get(k, "") for k in KEYS) + " |")
    out.append("\n## raw counters per launch\n")
    idx = {h: i for i, h in enumerate(hdr)}
    sel = [h for h in RAW if h in idx]
    out.append("| launch | " + " | ".join(sel) + " |\n|---|" + "---|" * len(sel))
    traffic = []
    for i, r in enumerate(rows[2:]):
        if len(r) != len(hdr):
            continue
        out.append(f"| {i} | " + " | ".join(r[idx[h]] for h in sel) + " |")
        if "dram__bytes_read.sum" in idx:
            traffic.append(num(r[idx["dram__bytes_read.sum"]]) + num(r[idx["dram__bytes_write.sum"]]))
    units = {h: rows[1][idx[h]] for h in sel}
    out.append("\nunits: " + ", ".join(f"{h}: {u}" for h, u in units.items()))
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out)[:3000])
    if traffic_json and traffic:
        u = units.get("dram__bytes_read.sum", "byte").lower()
        scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        try:
            allk = json.load(open(traffic_json))
        except Exception:
            allk = {}
        per = {}
        for c, tr in zip(launches_, traffic):
            kname = c["name"].replace("void ", "").split("<")[0].split("(")[0].strip()
            per.setdefault(kname, []).append(tr)
        for kname, v in per.items():
            allk[kname] = {"dram_bytes_per_launch": sum(v) * scale / len(v), "launches": len(v), "source": dst}
        json.dump(allk, open(traffic_json, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        kernel(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)

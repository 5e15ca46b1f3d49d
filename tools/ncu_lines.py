"""Summarise an ncu report per CUDA source line: python tools/ncu_lines.py report.ncu-rep [top] [kernel-substr]"""
import csv, subprocess, sys

rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
def num(x):
    try: return float(x.replace(',', ''))
    except Exception: return 0.0
rows, fname, hdr, kern, seen = [], None, None, None, set()
for l in out.splitlines():
    if l.startswith('"Kernel Name"'):
        kern = l
        if len(seen) >= 1 and kern not in seen: pass
        seen.add(kern); continue
    if l.startswith('"File Path"'):
        fname = l.split(',')[1].strip('"').split('/')[-1]; continue
    if l.startswith('"Line No"'):
        hdr = next(csv.reader([l])); continue
    if hdr is None or not l.startswith('"'): continue
    r = next(csv.reader([l]))
    if len(r) != len(hdr) or r[0] == '' : continue
    d = {}
    for k, v in zip(hdr, r):
        if k not in d: d[k] = v
    rows.append((fname, d))
print(sorted(seen)[0][:100] if seen else '')
tot_s = sum(num(r['# Samples']) for _, r in rows) or 1
tot_i = sum(num(r['Instructions Executed']) for _, r in rows) or 1
print(f"total samples {tot_s:.0f}, warp instructions {tot_i:.3e} (all captured launches of all kernels in the report)")
# one row per (file, line): the report repeats every line once per captured launch
agg = {}
for f, r in rows:
    k = (f, r['Line No'])
    if k not in agg:
        agg[k] = dict(r); agg[k]['_n'] = 1
        for kk, v in r.items():
            if kk.startswith('stall_') or kk in ('# Samples', 'Instructions Executed', 'Avg. Threads Executed'): agg[k][kk] = num(v)
    else:
        a = agg[k]; a['_n'] += 1
        for kk, v in r.items():
            if kk.startswith('stall_') or kk in ('# Samples', 'Instructions Executed', 'Avg. Threads Executed'): a[kk] += num(v)
rows = []
for (f, ln), a in agg.items():
    a['Avg. Threads Executed'] = f"{a['Avg. Threads Executed'] / a['_n']:.0f}"
    for kk in list(a):
        if kk.startswith('stall_') or kk in ('# Samples', 'Instructions Executed'): a[kk] = str(a[kk])
    rows.append((f, a))
rows.sort(key=lambda fr: -num(fr[1]['# Samples']))
for f, r in rows[:top]:
    ns = max(1, num(r['# Samples']))
    stalls = {k[6:]: num(v) for k, v in r.items() if k.startswith('stall_') and 'Not Issued' not in k and num(v) > 0}
    ts = sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
    st = ' '.join(f'{k}={v / ns * 100:.0f}%' for k, v in ts)
    print(f"{f}:{r['Line No']:>4} samp {num(r['# Samples']) / tot_s * 100:5.1f}% inst {num(r['Instructions Executed']) / tot_i * 100:5.1f}% "
          f"thr {r['Avg. Threads Executed']:>4} {st:42s} | {r['Source'].strip()[:80]}")

#!/bin/bash
# ncu --set full of the HEAVIEST launches of one kernel in the timed step of the headline bench.
#   tools/ncu_heaviest.sh <kernel regex> <output name> [launches to capture]
# KBASE=mangled matches the regex against the mangled name, which tells template instantiations apart: k_grow_mid4ILi4ELi2ELb0 (ncu splits -k values at commas, so demangled template lists cannot be matched)
# Pass 1 lists every launch of the kernel with its duration, pass 2 captures the 3 (default) launches around the longest one.  Run under gpurun; the bench command must have exited 0 without ncu before (done here first).
K=$1; OUT=$2; N=${3:-3}
KB=${KBASE:+--kernel-name-base $KBASE}
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-secondary --no-job"
$CMD > gpurun_out/${OUT}_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none $KB -k "regex:$K" --csv --log-file gpurun_out/${OUT}_list.csv $CMD > /dev/null 2>&1
IDX=$(python - <<PY
import csv
rows=[r for r in csv.DictReader(l for l in open("gpurun_out/${OUT}_list.csv") if not l.startswith("=="))]
v=[float(r["Metric Value"].replace(",","")) for r in rows]
i=max(range(len(v)), key=lambda k:v[k])   # every step does the same work: the longest launch of any of them
print(max(0, i-1))
PY
)
echo "capturing $N launches of $K from launch $IDX"
ncu --set full --clock-control none --import-source on $KB -k "regex:$K" -s $IDX -c $N -o gpurun_out/$OUT -f $CMD > gpurun_out/${OUT}_ncu.log 2>&1
ls -la gpurun_out/$OUT.ncu-rep

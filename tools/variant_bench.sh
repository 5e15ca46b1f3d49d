#!/bin/bash
# A/B of build variants on one box: tools/variant_bench.sh name[:ENV=VAL,...] ...   (libs in build/variants/libtst_<name>.so; "base" = the in-tree library)
for spec in "$@"; do
  name=${spec%%:*}; envs=""
  if [[ "$spec" == *:* ]]; then envs=$(echo "${spec#*:}" | tr ',' ' '); fi
  lib=""; [[ "$name" != base ]] && lib="TST_LIB=$PWD/build/variants/libtst_$name.so"
  out=$(env $lib $envs timeout 200 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary 2>&1 | tail -1)
  python - "$spec" <<PY "$out"
import json,sys
try:
    d=json.loads(sys.argv[2]); t=d["roofline"]["tiers"]; w=d["roofline"]["whole_tree"]
    print(sys.argv[1], "value %.1f ms %.1f |"%(d["value"], d["ms_per_step"]), {k[7:]:round(v["ms"]/d["steps"],1) for k,v in t.items()}, "tiny %.1f prep %.1f infer %.1f rest %.1f"%(w.get("ms_tiny",0)/d["steps"], w["ms_prep"]/d["steps"], w["ms_infer"]/d["steps"], d["ms_per_step"]-(sum(v["ms"] for v in t.values())+w.get("ms_tiny",0)+w["ms_prep"]+w["ms_infer"])/d["steps"]), flush=True)
except Exception as e:
    print(sys.argv[1], "FAILED", sys.argv[2][-300:])
PY
done

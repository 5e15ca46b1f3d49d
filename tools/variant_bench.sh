#!/bin/bash
# A/B of build variants (TST_LIB) and knobs on the headline step; prints value + per-tier ms/step.  Usage: tools/variant_bench.sh "VAR=.. VAR=.." ...
for v in "$@"; do
  env $v python bench.py --steps 3 --warmup 3 --no-cpu --no-secondary --no-job --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
t = d['roofline']['tiers']; w = d['roofline']['whole_tree']; s = d['steps']
print('$v', '| trees/s %.0f ms/step %.1f |' % (d['value'], d['ms_per_step']), ' '.join('%s %.1f' % (k.replace('k_grow_', ''), x['ms'] / s) for k, x in t.items()), 'tiny %.1f prep %.1f infer %.1f' % (w['ms_tiny'] / s, w['ms_prep'] / s, w['ms_infer'] / s))
"
done

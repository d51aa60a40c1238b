#!/bin/bash
# A/B of environment knobs on the three config-2 passes: bash scripts/ab_env.sh "VAR=a" "VAR=b OTHER=c" ...
for v in "$@"; do echo "== $v"; env $v timeout 200 python bench.py --steps 20 --no-cpu-baseline --no-networks --no-double 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['value'], d['e2e']['value'], [(k['pass'][:12], k['ms']) for k in d['kernels']], d['clocks'].get('sm_mhz'))
    else: print(l.rstrip()[-300:])
"; done

for v in "CATTL3_TC_PAIRS=0" "CATTL3_TC_PAIRS=3"; do echo "== $v"; env $v timeout 200 python bench.py --steps 20 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['value'], d['e2e']['value'], [(k['pass'][:12], k['ms']) for k in d['kernels']], d['clocks'])
    else: print(l.rstrip()[-300:])
"; done

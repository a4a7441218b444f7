#!/bin/bash
# A/B of an environment switch on one box: tools/bench_ab.sh VAR v1 v2 [v1 v2 ...]
var=$1; shift
for v in "$@"; do
  env $var=$v timeout 400 python bench.py --no-cpu --no-exact 2>/dev/null > /tmp/ab.json
  python - "$var" "$v" <<'PY'
import json, sys
d = json.loads(open('/tmp/ab.json').read().strip().splitlines()[-1])
print(sys.argv[1], sys.argv[2], round(d['value'], 2), 'steps/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value'], 2))
PY
done

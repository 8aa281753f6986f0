#!/bin/bash
# usage: tools/sweep.sh ENVVAR v1 v2 ...  -- bench the lookahead kernel for each value of an environment knob
var=$1; shift
for v in "$@"; do
  env $var=$v timeout 100 python bench.py --steps 15 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$var=$v', round(d['value']), round(d['kernel_ms']['lookahead'],3))"
done

#!/bin/bash
# usage (on the GPU box): tools/ab_env.sh "bench args" VAR val1 val2 ... - runs bench.py with VAR set to each value, twice, interleaved
cd "$(dirname "$0")/.."
args=$1; var=$2; shift 2
for round in 1 2; do
for v in "$@"; do
  env $var=$v timeout 300 python bench.py $args 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$var=$v', '%.4g' % d['value'], '%.4f' % d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"
done; done

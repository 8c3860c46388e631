#!/bin/bash
# usage (on the GPU box): tools/ab_variants.sh "bench args" NAME...  - runs bench.py with each variant library, twice, interleaved
cd "$(dirname "$0")/.."
args=$1; shift
cp sid_b200/libsidgpu.so /tmp/libsidgpu_orig.so
for round in 1 2; do
for v in "$@"; do
  if [ "$v" = base ]; then cp /tmp/libsidgpu_orig.so sid_b200/libsidgpu.so; else cp sid_b200/variants/libsidgpu_$v.so sid_b200/libsidgpu.so; fi
  timeout 300 python bench.py $args 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', '%.4g' % d['value'], '%.4f' % d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"
done; done
cp /tmp/libsidgpu_orig.so sid_b200/libsidgpu.so

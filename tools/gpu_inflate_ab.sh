#!/bin/bash
# device inflate: parity tests and kernel throughput of every variant (SID_INFLATE = 0 warp per member; 4/8/16 lockstep)
mkdir -p gpurun_out/r2b
for v in 4 8 16 0; do
  echo "== SID_INFLATE=$v"
  SID_INFLATE=$v timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bgzf" 2>&1 | tail -2
  SID_INFLATE=$v timeout 300 python tools/inflate_bench.py ${1:-20000000} --kernel-only 2>&1 | tail -1 | python -c "
import json,sys
r=json.loads(sys.stdin.read()); print('whole', r['kernel_whole_file'], 'chunk', r['kernel_4000_members'])"
done

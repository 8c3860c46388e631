#!/bin/bash
# device inflate: parity tests and kernel throughput (+ an ncu capture when a name is given)
mkdir -p gpurun_out/r2b
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bgzf" 2>&1 | tail -2
timeout 300 python tools/inflate_bench.py ${1:-20000000} --kernel-only 2>&1 | tail -1 | python -c "
import json,sys
r=json.loads(sys.stdin.read()); print('whole', r['kernel_whole_file'], 'chunk', r['kernel_4000_members'])"
if [ -n "$2" ]; then
ncu --set full --clock-control none --import-source on -k regex:k_inflate -s 1 -c 1 -f -o gpurun_out/r2b/$2 python tools/inflate_bench.py 4000000 --kernel-only > gpurun_out/r2b/ncu_inflate.log 2>&1; tail -1 gpurun_out/r2b/ncu_inflate.log
fi

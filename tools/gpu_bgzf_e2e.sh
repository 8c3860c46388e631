#!/bin/bash
mkdir -p gpurun_out/r2b
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -m gpu -x -q -k "bgzf or gzip or call_io" 2>&1 | tail -2
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other 2>gpurun_out/r2b/e2e_bgzf.err | tail -1 > gpurun_out/r2b/e2e_bgzf.json
python -c "
import json
d=json.loads(open('gpurun_out/r2b/e2e_bgzf.json').read())
print('e2e', d['e2e']['ms_per_step'], d['e2e']['value']); print('het', d['e2e_het_only']['ms_per_step']); print('bgzf', d['e2e_bgzf']); print('cli', {k:v for k,v in d['cli_e2e'].items() if k!='bgzf'}); print('cli bgzf', d['cli_e2e'].get('bgzf'))"

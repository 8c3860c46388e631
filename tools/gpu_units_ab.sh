#!/bin/bash
# stage 2 by units (default build) against the 64-bit window form (variant "win"): parity tests, then interleaved bench runs
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tokenizer or csv_matches or fuzz or deep or full_size" 2>&1 | tail -3
tools/ab_variants.sh "--steps 5 --warmup 3 --sites 20000000 --no-e2e --no-cpu-baseline --no-other" base win
tools/ab_variants.sh "--steps 5 --warmup 3 --sites 2000000 --depth depth500 --no-e2e --no-cpu-baseline --no-other" base win

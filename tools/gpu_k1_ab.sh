#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tokenizer or csv_matches or fuzz or deep or full_size or small_chunks" 2>&1 | tail -2
tools/ab_variants.sh "--steps 5 --warmup 3 --sites 20000000 --no-e2e --no-cpu-baseline --no-other" base prev
tools/ab_variants.sh "--steps 5 --warmup 3 --sites 10000000 --depth depth60 --no-e2e --no-cpu-baseline --no-other" base prev

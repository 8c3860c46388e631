#!/bin/bash
# K1 A/B on one box: the committed library against the variant built from the commit before (tools/build_variant.sh prev after git stash); parity tests first
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tokenizer or csv_matches or fuzz or deep or full_size or small_chunks or strand" 2>&1 | tail -2
tools/ab_variants.sh "--steps 5 --warmup 3 --sites 20000000 --no-e2e --no-cpu-baseline --no-other" base prev

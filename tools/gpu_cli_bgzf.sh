#!/bin/bash
# where the streaming host spends its time on a BGZF file (SIDGPU_IO_TIMING) against the same text as a plain file
python - <<'PY'
import os, sys
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
from inflate_bench import bgzip
from sid_b200 import synth
text = synth.generate(int(os.environ.get("SITES", "25000000")), seed=1, **synth.CONFIGS["depth30"]).tobytes()
open("/dev/shm/t.plp", "wb").write(text)
open("/dev/shm/t.plp.gz", "wb").write(bgzip(text))
PY
for i in 1 2; do
for f in "/dev/shm/t.plp" "/dev/shm/t.plp.gz" "--host-inflate /dev/shm/t.plp.gz"; do
  echo "== sid -m local $f $EXTRA"
  SID_TIMING=1 SIDGPU_IO_TIMING=1 host/sid -m local $EXTRA $f 2>&1 >/dev/null | grep "^#"
done
done
rm -f /dev/shm/t.plp /dev/shm/t.plp.gz

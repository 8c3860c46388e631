# k_tok3 (streaming K1): parity subset, A/B against the site-store path, ncu capture
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests -m gpu -x -q -k "feed_rows or csv_matches or small_chunks or het_only or malformed or empty or deep or fuzz or cli" > gpurun_out/r2/t3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/t3_pytest.log
tail -15 gpurun_out/r2/t3_pytest.log
B="--steps 5 --warmup 3 --sites 20000000 --no-e2e --no-cpu-baseline"
rm -f gpurun_out/r2/t3_ab.txt
run() { name=$1; shift; timeout 300 env $ENVV python bench.py $B "$@" > gpurun_out/r2/t3_$name.json 2> gpurun_out/r2/t3_$name.err
  tail -1 gpurun_out/r2/t3_$name.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', '%.4g' % d['value'], '%.4f' % d['ms_per_step'], '%.3f' % d['roofline']['frac'], {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items() if v})" >> gpurun_out/r2/t3_ab.txt 2>&1; }
ENVV="X=1" run tok3
ENVV="SIDGPU_CHUNK_KB=32" run tok3_c32
ENVV="SIDGPU_CHUNK_KB=128" run tok3_c128
ENVV="SIDGPU_CHUNK_KB=256" run tok3_c256
ENVV="X=1" run unfused --unfused
ENVV="X=1" run d500 --depth depth500 --sites 1000000
ENVV="X=1" run d60 --depth depth60 --sites 10000000
cat gpurun_out/r2/t3_ab.txt
P="--steps 1 --warmup 1 --sites 5000000 --no-e2e --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tok3 -s 1 -c 1 -o gpurun_out/r2/prof_tok3 python bench.py $P > gpurun_out/r2/ncu_tok3.log 2>&1
ls -la gpurun_out/r2/*.ncu-rep

mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q -k "quality or deep or call_io or csv_matches" > gpurun_out/r2/pytest_h.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_h.log
tail -4 gpurun_out/r2/pytest_h.log
B="--steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-other"
rm -f gpurun_out/r2/q3_ab.txt
run() { name=$1; shift; timeout 300 env $ENVV python bench.py $B "$@" > gpurun_out/r2/q3_$name.json 2> gpurun_out/r2/q3_$name.err
  tail -1 gpurun_out/r2/q3_$name.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', '%.4g' % d['value'], '%.4f' % d['ms_per_step'], '%.3f' % d['roofline']['frac'], {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items() if v})" >> gpurun_out/r2/q3_ab.txt 2>&1; }
ENVV="X=1" run quality_c2 --method quality --sites 20000000
cp sid_b200/libsidgpu.so /tmp/orig.so
cp sid_b200/variants/libsidgpu_q3.so sid_b200/libsidgpu.so
ENVV="X=1" run quality_c3 --method quality --sites 20000000
cp /tmp/orig.so sid_b200/libsidgpu.so
cat gpurun_out/r2/q3_ab.txt
P="--steps 1 --warmup 1 --method quality --sites 5000000 --no-e2e --no-cpu-baseline --no-other"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_quality|k_tok2" -c 2 -o gpurun_out/r2/prof_quality2 python bench.py $P > gpurun_out/r2/ncu_quality2.log 2>&1
ls -la gpurun_out/r2/prof_quality2.ncu-rep

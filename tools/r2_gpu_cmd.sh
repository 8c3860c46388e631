mkdir -p gpurun_out/r2
rm -f gpurun_out/r2/ab.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest3.log
tail -3 gpurun_out/r2/pytest3.log
B="--steps 5 --warmup 3 --sites 20000000 --no-e2e --no-cpu-baseline"
run() { name=$1; shift; timeout 200 env "$@" python bench.py $B $EXTRA > gpurun_out/r2/ab_$name.json 2> gpurun_out/r2/ab_$name.err
  tail -1 gpurun_out/r2/ab_$name.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', '%.4g' % d['value'], '%.4f' % d['ms_per_step'], '%.3f' % d['roofline']['frac'], {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items() if v})" >> gpurun_out/r2/ab.txt 2>&1; }
EXTRA=""
run fused X=1
run fused_s2 SIDGPU_TOK_STAGES=2
EXTRA="--unfused"
run unfused X=1
run unfused_s2 SIDGPU_TOK_STAGES=2
EXTRA=""
cp sid_b200/libsidgpu.so /tmp/orig.so
for v in noasm nosfx nojoin nostage2; do cp sid_b200/variants/libsidgpu_$v.so sid_b200/libsidgpu.so; run whatif_$v X=1; done
cp /tmp/orig.so sid_b200/libsidgpu.so
cat gpurun_out/r2/ab.txt
P="--steps 1 --warmup 1 --sites 5000000 --no-e2e --no-cpu-baseline"
python bench.py $P > gpurun_out/r2/plain_rows.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_tok2 -s 1 -c 1 -o gpurun_out/r2/prof_rows python bench.py $P > gpurun_out/r2/ncu_rows.log 2>&1
python bench.py $P --unfused > gpurun_out/r2/plain_sites.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_tok2 -s 1 -c 1 -o gpurun_out/r2/prof_sites python bench.py $P --unfused > gpurun_out/r2/ncu_sites.log 2>&1
ls -la gpurun_out/r2/*.ncu-rep

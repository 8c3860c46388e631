# baseline of the committed build: GPU suite, A/B of fused vs unfused local, ncu of both k_tok2 forms
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest.log
tail -3 gpurun_out/r2/pytest.log
B="--steps 5 --warmup 3 --sites 20000000 --no-e2e --no-cpu-baseline"
rm -f gpurun_out/r2/ab.txt
run() { name=$1; shift; timeout 300 python bench.py $B "$@" > gpurun_out/r2/ab_$name.json 2> gpurun_out/r2/ab_$name.err
  tail -1 gpurun_out/r2/ab_$name.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', '%.4g' % d['value'], '%.4f' % d['ms_per_step'], '%.3f' % d['roofline']['frac'], {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items() if v})" >> gpurun_out/r2/ab.txt 2>&1; }
run fused
run unfused --unfused
run d500 --depth depth500 --sites 1000000
run d500u --depth depth500 --sites 1000000 --unfused
run d60bayes --depth depth60 --method bayes
run quality --method quality --sites 10000000
cat gpurun_out/r2/ab.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2/bench_full.json 2> gpurun_out/r2/bench_full.err; tail -c 3000 gpurun_out/r2/bench_full.json
P="--steps 1 --warmup 1 --sites 5000000 --no-e2e --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tok2 -s 1 -c 1 -o gpurun_out/r2/prof_rows python bench.py $P > gpurun_out/r2/ncu_rows.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tok2 -s 1 -c 1 -o gpurun_out/r2/prof_sites python bench.py $P --unfused > gpurun_out/r2/ncu_sites.log 2>&1
ls -la gpurun_out/r2/*.ncu-rep

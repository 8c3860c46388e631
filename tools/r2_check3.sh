# after the DEEP instantiation: suite, A/B, k_quality capture, cli timing
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_e.log
tail -6 gpurun_out/r2/pytest_e.log
B="--steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-other"
rm -f gpurun_out/r2/deep3_ab.txt
run() { name=$1; shift; timeout 300 env $ENVV python bench.py $B "$@" > gpurun_out/r2/deep3_$name.json 2> gpurun_out/r2/deep3_$name.err
  tail -1 gpurun_out/r2/deep3_$name.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', '%.4g' % d['value'], '%.4f' % d['ms_per_step'], '%.3f' % d['roofline']['frac'], {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items() if v})" >> gpurun_out/r2/deep3_ab.txt 2>&1; }
ENVV="X=1" run d30 --sites 20000000
ENVV="X=1" run d500_on --depth depth500 --sites 2000000
ENVV="SIDGPU_DEEP_LINES=0" run d500_off --depth depth500 --sites 2000000
ENVV="SIDGPU_DEEP_LINES=1" run d60_on --depth depth60 --sites 10000000
ENVV="X=1" run d60 --depth depth60 --sites 10000000
ENVV="X=1" run d60bayes --depth depth60 --sites 20000000 --method bayes
cat gpurun_out/r2/deep3_ab.txt
timeout 300 python bench.py --steps 3 --warmup 3 --no-other --no-cpu-baseline > gpurun_out/r2/bench_cli.json 2> gpurun_out/r2/bench_cli.err
python -c "import json; d=json.loads(open('gpurun_out/r2/bench_cli.json').read().strip().splitlines()[-1]); print(d['cli_e2e'])"
P="--steps 1 --warmup 1 --method quality --sites 5000000 --no-e2e --no-cpu-baseline --no-other"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_quality -c 1 -o gpurun_out/r2/prof_quality python bench.py $P > gpurun_out/r2/ncu_quality.log 2>&1
ls -la gpurun_out/r2/prof_quality.ncu-rep

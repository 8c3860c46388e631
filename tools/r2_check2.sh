# full GPU suite; A/B of the deep path; the whole default bench line
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_d.log
tail -6 gpurun_out/r2/pytest_d.log
B="--steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-other"
rm -f gpurun_out/r2/deep2_ab.txt
run() { name=$1; shift; timeout 300 env $ENVV python bench.py $B "$@" > gpurun_out/r2/deep2_$name.json 2> gpurun_out/r2/deep2_$name.err
  tail -1 gpurun_out/r2/deep2_$name.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', '%.4g' % d['value'], '%.4f' % d['ms_per_step'], '%.3f' % d['roofline']['frac'], {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items() if v})" >> gpurun_out/r2/deep2_ab.txt 2>&1; }
ENVV="X=1" run d30 --sites 20000000
ENVV="X=1" run d500_on --depth depth500 --sites 2000000
ENVV="SIDGPU_DEEP_LINES=0" run d500_off --depth depth500 --sites 2000000
ENVV="X=1" run d60bayes --depth depth60 --sites 20000000 --method bayes
cat gpurun_out/r2/deep2_ab.txt
timeout 900 python bench.py > gpurun_out/r2/bench_full2.json 2> gpurun_out/r2/bench_full2.err
echo "bench rc=$?"; tail -c 7000 gpurun_out/r2/bench_full2.json; tail -3 gpurun_out/r2/bench_full2.err

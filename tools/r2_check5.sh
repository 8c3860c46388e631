mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_g.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_g.log
tail -6 gpurun_out/r2/pytest_g.log
B="--steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-other"
rm -f gpurun_out/r2/q2_ab.txt
run() { name=$1; shift; timeout 300 env $ENVV python bench.py $B "$@" > gpurun_out/r2/q2_$name.json 2> gpurun_out/r2/q2_$name.err
  tail -1 gpurun_out/r2/q2_$name.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', '%.4g' % d['value'], '%.4f' % d['ms_per_step'], '%.3f' % d['roofline']['frac'], {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items() if v})" >> gpurun_out/r2/q2_ab.txt 2>&1; }
ENVV="X=1" run quality --method quality --sites 20000000
ENVV="SIDGPU_QUALITY_K1=0" run quality_old --method quality --sites 20000000
ENVV="X=1" run d30 --sites 20000000
ENVV="X=1" run d60 --depth depth60 --sites 10000000
cat gpurun_out/r2/q2_ab.txt

# full GPU suite, then A/B of the window-per-lane stage 2 on deep pileups (and that depth 30 did not move)
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_c.log
tail -6 gpurun_out/r2/pytest_c.log
B="--steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-other"
rm -f gpurun_out/r2/deep_ab.txt
run() { name=$1; shift; timeout 300 env $ENVV python bench.py $B "$@" > gpurun_out/r2/deep_$name.json 2> gpurun_out/r2/deep_$name.err
  tail -1 gpurun_out/r2/deep_$name.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', '%.4g' % d['value'], '%.4f' % d['ms_per_step'], '%.3f' % d['roofline']['frac'], {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items() if v})" >> gpurun_out/r2/deep_ab.txt 2>&1; }
ENVV="X=1" run d30 --sites 20000000
ENVV="X=1" run d500_on --depth depth500 --sites 2000000
ENVV="SIDGPU_DEEP_LINES=0" run d500_off --depth depth500 --sites 2000000
ENVV="X=1" run d60 --depth depth60 --sites 10000000
ENVV="SIDGPU_DEEP_LINES=1" run d60_on --depth depth60 --sites 10000000
cat gpurun_out/r2/deep_ab.txt
P="--steps 1 --warmup 2 --depth depth500 --sites 1000000 --no-e2e --no-cpu-baseline --no-other"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tok2 -s 2 -c 1 -o gpurun_out/r2/prof_deep python bench.py $P > gpurun_out/r2/ncu_deep.log 2>&1
ls -la gpurun_out/r2/prof_deep.ncu-rep

# end-of-round evidence on one GPU: suite, the default bench line, the reference arm, the launch list under ncu
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_final.log
python __graft_entry__.py --smoke 2>&1 | tail -1
tail -4 gpurun_out/r2/pytest_final.log
timeout 900 python bench.py > gpurun_out/r2/bench_final.json 2> gpurun_out/r2/bench_final.err
echo "bench rc=$?"; tail -c 2500 gpurun_out/r2/bench_final.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2/bench_reference.json 2> gpurun_out/r2/bench_reference.err
echo "ref rc=$?"; tail -c 600 gpurun_out/r2/bench_reference.json
P="--steps 2 --warmup 1 --sites 20000000 --no-e2e --no-cpu-baseline --no-other"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2/launches.csv python bench.py $P > gpurun_out/r2/ncu_launches.log 2>&1
tail -3 gpurun_out/r2/launches.csv
B="--steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-other"
timeout 300 python bench.py $B --method quality --sites 20000000 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('quality', '%.4g' % d['value'], '%.4f' % d['ms_per_step'], {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items() if v})"

# full GPU suite + a reduced-size run of the whole bench line (flow check of other_configs / cli_e2e)
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_b.log
tail -12 gpurun_out/r2/pytest_b.log
timeout 600 python bench.py --sites 20000000 --lynch-sites 20000000 --deep-sites 500000 --quality-sites 5000000 --steps 3 --warmup 3 > gpurun_out/r2/bench_small.json 2> gpurun_out/r2/bench_small.err
echo "bench rc=$?"; tail -c 6000 gpurun_out/r2/bench_small.json; tail -5 gpurun_out/r2/bench_small.err

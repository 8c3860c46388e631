# single GPU: the GPU suite, then bench.py kernel-only lines for the in-tree library and for every variant under sid_b200/variants/
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests -m gpu -x -q ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/r2/pytest_ab.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_ab.log
tail -4 gpurun_out/r2/pytest_ab.log
B="--steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-other"
rm -f gpurun_out/r2/ab.txt
run() { name=$1; shift; timeout 300 python bench.py $B "$@" > gpurun_out/r2/ab_$name.json 2> gpurun_out/r2/ab_$name.err
  tail -1 gpurun_out/r2/ab_$name.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', '%.4g' % d['value'], '%.4f' % d['ms_per_step'], '%.3f' % d['roofline']['frac'], {k: round(v,3) for k,v in d['roofline']['kernel_ms_per_step'].items() if v})" >> gpurun_out/r2/ab.txt 2>&1; }
run quality --method quality --sites 20000000
run d30 --sites 20000000
cp sid_b200/libsidgpu.so /tmp/orig.so
for f in sid_b200/variants/libsidgpu_*.so; do [ -e "$f" ] || continue; v=$(basename $f .so | sed 's/libsidgpu_//'); cp $f sid_b200/libsidgpu.so; run quality_$v --method quality --sites 20000000; done
cp /tmp/orig.so sid_b200/libsidgpu.so
cat gpurun_out/r2/ab.txt

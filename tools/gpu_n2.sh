# two GPUs: the multi-rank GPU tests, the whole bench line at N=2 (NCCL paths of other_configs), the copy bound
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests -m gpu -x -q -k "multi or cpp_api or call_io or pipe" > gpurun_out/r2/pytest_n2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_n2.log
tail -5 gpurun_out/r2/pytest_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2/bench_n2.json 2> gpurun_out/r2/bench_n2.err
echo "bench rc=$?"; tail -c 5000 gpurun_out/r2/bench_n2.json; tail -5 gpurun_out/r2/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/copy_bench.py > gpurun_out/r2/copy_n2.json 2> gpurun_out/r2/copy_n2.err
echo "copy rc=$?"; tail -c 1500 gpurun_out/r2/copy_n2.json

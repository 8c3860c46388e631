# eight GPUs: the concurrent copy bound of the box, then the whole bench line (NCCL paths of other_configs included)
mkdir -p gpurun_out/r2
nvidia-smi topo -m > gpurun_out/r2/topo_n8.txt 2>&1
lscpu | head -30 >> gpurun_out/r2/topo_n8.txt 2>&1
numactl -H >> gpurun_out/r2/topo_n8.txt 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/copy_bench.py --reps 2 > gpurun_out/r2/copy_n8.json 2> gpurun_out/r2/copy_n8.err
echo "copy rc=$?"; tail -c 600 gpurun_out/r2/copy_n8.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2/bench_n8.json 2> gpurun_out/r2/bench_n8.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/r2/bench_n8.json; tail -5 gpurun_out/r2/bench_n8.err

# weak-scaling bench on N GPUs of one box.  usage (via gpurun --gpus N): bash scripts/gpu_multi.sh N <tag>
N=${1:-8}; TAG=${2:-x}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 3000 --warmup 300 --no-cpu-baseline --e2e-steps 50 > gpurun_out/bench_${TAG}_${N}gpu.json 2> gpurun_out/bench_${TAG}_${N}gpu.err
tail -c 1500 gpurun_out/bench_${TAG}_${N}gpu.json; tail -3 gpurun_out/bench_${TAG}_${N}gpu.err

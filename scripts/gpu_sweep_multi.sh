# Config-5 sweep (BASELINE.json configs[4]) at N GPUs: total env batch 256 -> 65536 at 50 UE x 10 BS, split evenly over
# the GPUs (the 8192 row is the strong-scaling line of the north-star batch).  usage: bash scripts/gpu_sweep_multi.sh N tag
N=${1:-1}; TAG=${2:-r02}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
OUT=gpurun_out/sweep_${TAG}_n${N}.jsonl
: > $OUT
for T in 256 1024 4096 8192 16384 65536; do
  if [ "$N" = 1 ]; then
    timeout 300 python bench.py --gpus 1 --total-envs $T --steps 1000 --warmup 100 --reps 3 --no-cpu-baseline --e2e-steps 50 >> $OUT 2>> gpurun_out/sweep_${TAG}_n${N}.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --total-envs $T --steps 1000 --warmup 100 --reps 3 --no-cpu-baseline --e2e-steps 50 >> $OUT 2>> gpurun_out/sweep_${TAG}_n${N}.err
  fi
  tail -1 $OUT | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('total', $T, 'N', d['n_gpus'], 'env-steps/s %.4e'%d['value'], 'us/step %.2f'%(1e3*d['ms_per_step']), 'frac %.3f'%d['roofline']['frac'], 'e2e %.3e'%d['e2e']['value'])"
done
# the driver's own invocation at this N (weak scaling, 1024 envs per GPU)
if [ "$N" = 1 ]; then
  timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_driver_n${N}.json 2> gpurun_out/bench_${TAG}_driver_n${N}.err
else
  NCCL_DEBUG=INFO timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_driver_n${N}.json 2> gpurun_out/bench_${TAG}_driver_n${N}.err
fi
tail -1 gpurun_out/bench_${TAG}_driver_n${N}.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('driver flags N', d['n_gpus'], 'env-steps/s %.4e'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'e2e %.3e'%d['e2e']['value'], 'ceil %.3e'%d['e2e']['pcie_ceiling']['value'], d['rep_ms'])"
grep -c "NCCL INFO" gpurun_out/bench_${TAG}_driver_n${N}.err; grep -o "nranks [0-9]*" gpurun_out/bench_${TAG}_driver_n${N}.err | sort | uniq -c | head -3

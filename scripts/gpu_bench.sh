# usage (on the GPU box, via gpurun): bash scripts/gpu_bench.sh <tag>
TAG=${1:-r01}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/pytest_$TAG.log 2>&1; tail -3 gpurun_out/pytest_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 3000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
# launch list (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 300 --warmup 100 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_launch_$TAG.log 2>&1
# full capture of the dominant kernel (the 100-step fragment launches; skip the buffer-allocation and warm-up ones)
ncu --set full --clock-control none --import-source on -k regex:dcb_step_kernel -s 7 -c 1 -o gpurun_out/prof_$TAG -f \
    python bench.py --steps 300 --warmup 100 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out/

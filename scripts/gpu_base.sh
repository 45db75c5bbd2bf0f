# baseline check: GPU tests, quick bench, one full ncu capture of a fragment launch.  usage: bash scripts/gpu_base.sh <tag>
TAG=${1:-x}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_$TAG.log 2>&1; tail -3 gpurun_out/pytest_$TAG.log
timeout 200 python bench.py --steps 3000 --warmup 300 --no-cpu-baseline --e2e-steps 20 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 1500 gpurun_out/bench_$TAG.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dcb_step_kernel -s 7 -c 1 -o gpurun_out/prof_$TAG -f \
    python bench.py --steps 300 --warmup 100 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out/

# usage: bash scripts/gpu_prof.sh <tag> [bench args]  -- full ncu capture of one 100-step fragment launch of the fused kernel
TAG=${1:-x}; shift
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dcb_step_kernel -s 5 -c 1 -o gpurun_out/prof_$TAG -f \
    python bench.py --steps 300 --warmup 100 --reps 1 --no-cpu-baseline --e2e-steps 3 "$@" > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out/prof_$TAG.ncu-rep

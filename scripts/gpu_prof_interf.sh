# full ncu capture of one wide-kernel launch at BASELINE config 4 with the interference extension.  usage: bash scripts/gpu_prof_interf.sh <tag>
TAG=${1:-x}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dcb_wide_kernel -s 4 -c 1 -o gpurun_out/prof_interf_$TAG -f \
    python bench.py --n-ue 1000 --n-bs 50 --envs 1024 --fragment 4 --steps 12 --warmup 4 --reps 1 --no-cpu-baseline --e2e-steps 1 --interference > gpurun_out/ncu_interf_$TAG.log 2>&1
ls -la gpurun_out/prof_interf_$TAG.ncu-rep; tail -3 gpurun_out/ncu_interf_$TAG.log

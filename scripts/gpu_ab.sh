cd $GRAFT_REPO_ROOT
run() { timeout 120 python bench.py --steps 3000 --warmup 300 --no-cpu-baseline --e2e-steps 5 "$@" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('env-steps/s %.4e'%d['value'], 'us/step %.2f'%(1e3*d['ms_per_step']), d['config']['launch_geometry'])"; }
for rep in 1 2; do
echo "A (alt lib)"; DCB_LIB_PATH=$GRAFT_REPO_ROOT/gpurun_ab_v6.so run
echo "B (tree lib)"; run
done
echo "large K:"; 
echo A; DCB_LIB_PATH=$GRAFT_REPO_ROOT/gpurun_ab_v6.so run --envs 16384 --fragment 25 --steps 500 --warmup 100
echo B; run --envs 16384 --fragment 25 --steps 500 --warmup 100

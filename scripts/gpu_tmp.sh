cd $GRAFT_REPO_ROOT
run() { timeout 120 python bench.py --steps 3000 --warmup 300 --no-cpu-baseline --e2e-steps 5 "$@" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('env-steps/s %.4e'%d['value'], 'us/step %.2f'%(1e3*d['ms_per_step']), d['config']['launch_geometry'])"; }
for s in 2 4 8; do echo "S=$s"; DCB_REDUCE_LANES=$s run; done
echo central; run --kind central
echo cfg3; run --n-ue 200 --n-bs 20 --envs 512 --fragment 50 --steps 1000 --warmup 100
echo k16384; run --envs 16384 --fragment 25 --steps 200 --warmup 50

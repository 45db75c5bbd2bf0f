cd $GRAFT_REPO_ROOT
for KF in "256 100" "1024 100" "4096 100" "8192 50" "16384 25" "65536 10"; do
  set -- $KF
  python bench.py --envs $1 --fragment $2 --steps 500 --warmup 100 --no-cpu-baseline --e2e-steps 3 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('K', d['config']['envs_per_gpu'], 'env-steps/s %.3e'%d['value'], 'us/step %.2f'%(1e3*d['ms_per_step']), 'frac %.3f'%d['roofline']['frac'], d['run']['launch_geometry'])"
done

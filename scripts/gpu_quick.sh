# quick GPU check with hang protection: bash scripts/gpu_quick.sh
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests -q -m gpu -x 2>&1 | tail -15
timeout 120 python bench.py --steps 2000 --warmup 200 --no-cpu-baseline --e2e-steps 20 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('env-steps/s %.3e'%d['value'], 'us/step %.2f'%(1e3*d['ms_per_step']), 'frac %.3f'%d['roofline']['frac'], d['run']['launch_geometry'], 'e2e %.3e'%d['e2e']['value'])"

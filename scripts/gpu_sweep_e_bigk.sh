cd $GRAFT_REPO_ROOT
for E in 1 2 3 4 5 7; do
  DCB_ENVS_PER_CTA=$E timeout 120 python bench.py --envs 16384 --fragment 25 --steps 300 --warmup 50 --no-cpu-baseline --e2e-steps 2 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('E', d['run']['launch_geometry']['envs_per_cta'], 'env-steps/s %.3e'%d['value'], d['run']['launch_geometry'])"
done
echo central:
timeout 120 python bench.py --kind central --steps 3000 --warmup 300 --no-cpu-baseline --e2e-steps 20 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('central env-steps/s %.3e'%d['value'], 'frac %.3f'%d['roofline']['frac'], d['run']['launch_geometry'], 'e2e %.3e'%d['e2e']['value'])"
echo 200x20x512:
timeout 120 python bench.py --n-ue 200 --n-bs 20 --envs 512 --steps 1000 --warmup 100 --no-cpu-baseline --e2e-steps 5 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('200x20 env-steps/s %.3e'%d['value'], 'frac %.3f'%d['roofline']['frac'], d['run']['launch_geometry'])"

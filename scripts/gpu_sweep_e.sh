cd $GRAFT_REPO_ROOT
for E in 1 2 3 4 5 7; do
  echo "E=$E"; DCB_ENVS_PER_CTA=$E python bench.py --steps 2000 --warmup 200 --no-cpu-baseline --e2e-steps 5 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['run']['launch_geometry'])"
done

# A/B of experiment builds: bash scripts/gpu_exp.sh <lib1.so> <lib2.so> ...   ("tree" = the in-tree library)
cd $GRAFT_REPO_ROOT
run() { timeout 120 python bench.py --steps 3000 --warmup 300 --no-cpu-baseline --e2e-steps 5 "$@" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('env-steps/s %.4e'%d['value'], 'us/step %.2f'%(1e3*d['ms_per_step']), d['run']['launch_geometry'])"; }
for rep in 1 2; do
for lib in "$@"; do
  if [ "$lib" = tree ]; then echo "== tree"; run; else echo "== $lib"; DCB_LIB_PATH=$GRAFT_REPO_ROOT/$lib run; fi
done
done

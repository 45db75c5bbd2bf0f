# A/B of library builds on the headline bench only: bash scripts/gpu_ab_only.sh <lib1.so> ... ("tree" = in-tree)
cd $GRAFT_REPO_ROOT
run() { timeout 120 python bench.py --steps 3000 --warmup 300 --reps 3 --no-cpu-baseline --e2e-steps 5 "$@" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('env-steps/s %.4e'%d['value'], 'us/step %.3f'%(1e3*d['ms_per_step']), 'frac %.4f'%d['roofline']['frac'])"; }
for rep in 1 2; do
for lib in "$@"; do
  if [ "$lib" = tree ]; then echo "== tree"; run; else echo "== $lib"; DCB_LIB_PATH=$GRAFT_REPO_ROOT/$lib run; fi
done
done

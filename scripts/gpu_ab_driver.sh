# A/B of library builds on the DRIVER's invocation (one 20-step launch per repetition) and on the long run, then the fused
# kernel's parity tests on the LAST library named.  usage: bash scripts/gpu_ab_driver.sh <lib1.so> ... ("tree" = in-tree)
cd $GRAFT_REPO_ROOT
fmt='import json,sys; d=json.loads(sys.stdin.read()); print("  us/step %.3f  frac %.4f  launch us %.2f  reps %s" % (1e3*d["ms_per_step"], d["roofline"]["frac"], 1e3*d["roofline"]["avg_launch_ms"], [round(1e3*x,1) for x in d["rep_ms"]]))'
short() { timeout 120 python bench.py --steps 20 --warmup 5 --reps 9 --no-cpu-baseline --e2e-steps 5 | python -c "$fmt"; }
long() { timeout 120 python bench.py --steps 3000 --warmup 300 --reps 3 --no-cpu-baseline --e2e-steps 5 | python -c "$fmt"; }
for rep in 1 2; do
for lib in "$@"; do
  [ "$lib" = tree ] && unset DCB_LIB_PATH || export DCB_LIB_PATH=$GRAFT_REPO_ROOT/$lib
  echo "== $lib (driver flags)"; short
done
done
for lib in "$@"; do
  [ "$lib" = tree ] && unset DCB_LIB_PATH || export DCB_LIB_PATH=$GRAFT_REPO_ROOT/$lib
  echo "== $lib (long run)"; long
done
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -q -m gpu -x -k "golden or step_many or auto_reset or partial_reset or step_host or policies or continuous" 2>&1 | tail -3

# A/B of wide-kernel builds at BASELINE config 4: every gpurun_exp_W*.so against the in-tree library
cd $GRAFT_REPO_ROOT
fmt='import json,sys; d=json.loads(sys.stdin.read()); print("  %-8s us/step %.2f  frac %.4f" % (d["config"].get("kind", "?") if isinstance(d["config"], dict) else "?", 1e3*d["ms_per_step"], d["roofline"]["frac"]))'
run() { timeout 300 python bench.py --no-cpu-baseline --n-ue 1000 --n-bs 50 --envs 1024 --fragment 10 --steps 100 --warmup 20 --reps 3 --e2e-steps 3 "$@" 2>/dev/null | python -c "$fmt"; }
for lib in deepcomp_b200/libdeepcomp_b200.so gpurun_exp_W*.so; do
  echo "== $lib"
  DCB_LIB_PATH=$GRAFT_REPO_ROOT/$lib run
  DCB_LIB_PATH=$GRAFT_REPO_ROOT/$lib run --kind central
done

cd $GRAFT_REPO_ROOT
fmt='import json,sys; d=json.loads(sys.stdin.read()); print("%s: env-steps/s %.4e  us/step %.2f  frac %.3f  %s" % (d["config"]["workload"][:44], d["value"], 1e3*d["ms_per_step"], d["roofline"]["frac"], d["config"]["launch_geometry"]))'
run() { timeout 600 python bench.py --no-cpu-baseline --e2e-steps 3 "$@" 2>/dev/null | python -c "$fmt"; }
run --n-ue 1000 --n-bs 50 --envs 1024 --fragment 10 --steps 100 --warmup 20
run --n-ue 1000 --n-bs 50 --envs 1024 --fragment 10 --steps 100 --warmup 20 --kind central
timeout 600 python -m pytest tests -x -q -m gpu -k "wide" 2>&1 | tail -3

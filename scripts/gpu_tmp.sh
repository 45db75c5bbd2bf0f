cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
fmt='import json,sys; d=json.loads(sys.stdin.read()); print("%s: env-steps/s %.4e  us/step %.2f  frac %.3f  %s" % (d["config"]["workload"][:44], d["value"], 1e3*d["ms_per_step"], d["roofline"]["frac"], d["config"]["launch_geometry"]))'
run() { timeout 600 python bench.py --no-cpu-baseline --e2e-steps 3 "$@" 2>/dev/null | python -c "$fmt"; }
run --steps 3000 --warmup 300
run --n-ue 200 --n-bs 20 --envs 512 --fragment 50 --steps 1000 --warmup 100
run --n-ue 200 --n-bs 20 --envs 4096 --fragment 10 --steps 200 --warmup 50
run --kind central --steps 3000 --warmup 300

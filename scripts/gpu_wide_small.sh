cd $GRAFT_REPO_ROOT
fmt='import json,sys; d=json.loads(sys.stdin.read()); print("%s %s: env-steps/s %.4e  us/step %.2f  frac %.4f" % (d["config"]["workload"][:40], d["roofline"]["kernel"], d["value"], 1e3*d["ms_per_step"], d["roofline"]["frac"]))'
run() { timeout 600 python bench.py --no-cpu-baseline "$@" 2>/dev/null | python -c "$fmt"; }
DCB_FORCE_WIDE=1 run --n-ue 200 --n-bs 20 --envs 512 --fragment 50 --steps 500 --warmup 100 --reps 3 --e2e-steps 3
DCB_FORCE_WIDE=1 run --n-ue 200 --n-bs 20 --envs 512 --fragment 50 --steps 500 --warmup 100 --reps 3 --e2e-steps 3 --kind central
DCB_FORCE_WIDE=1 run --steps 1000 --warmup 100 --reps 3 --e2e-steps 3

cd $GRAFT_REPO_ROOT
fmt='import json,sys; d=json.loads(sys.stdin.read()); print("%s: env-steps/s %.4e  us/step %.2f  frac %.3f  %s %s" % (d["config"]["workload"][:40], d["value"], 1e3*d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel"], d["run"]["launch_geometry"]))'
echo "cfg3 forced wide"; DCB_FORCE_WIDE=1 timeout 300 python bench.py --no-cpu-baseline --n-ue 200 --n-bs 20 --envs 512 --fragment 50 --steps 500 --warmup 100 --reps 3 --e2e-steps 5 | python -c "$fmt"
echo "cfg3 fused"; timeout 300 python bench.py --no-cpu-baseline --n-ue 200 --n-bs 20 --envs 512 --fragment 50 --steps 500 --warmup 100 --reps 3 --e2e-steps 5 | python -c "$fmt"
echo "K=256 forced wide"; DCB_FORCE_WIDE=1 timeout 300 python bench.py --no-cpu-baseline --envs 256 --steps 1000 --warmup 100 --reps 3 --e2e-steps 5 | python -c "$fmt"
for E in 1 2 3; do echo "K=256 fused E=$E"; DCB_ENVS_PER_CTA=$E timeout 300 python bench.py --no-cpu-baseline --envs 256 --steps 1000 --warmup 100 --reps 3 --e2e-steps 5 | python -c "$fmt"; done
echo "headline forced wide"; DCB_FORCE_WIDE=1 timeout 300 python bench.py --no-cpu-baseline --steps 1000 --warmup 100 --reps 3 --e2e-steps 5 | python -c "$fmt"

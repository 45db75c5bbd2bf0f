# Wide-kernel check: the whole GPU suite, the parity / API suites again with the wide kernel forced, then config 4.
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6
echo "== forced wide (failures named *fused* / kernel_name are the tests that insist on the fused kernel)"
DCB_FORCE_WIDE=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_gpu_soak.py -q -m gpu 2>&1 | tail -15
fmt='import json,sys; d=json.loads(sys.stdin.read()); print("%s: env-steps/s %.4e  us/step %.2f  frac %.4f  e2e %.3e" % (d["config"]["workload"][:50], d["value"], 1e3*d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"]))'
run() { timeout 600 python bench.py --no-cpu-baseline "$@" 2>/dev/null | python -c "$fmt"; }
run --n-ue 1000 --n-bs 50 --envs 1024 --fragment 10 --steps 100 --warmup 20 --e2e-steps 3
run --n-ue 1000 --n-bs 50 --envs 1024 --fragment 10 --steps 100 --warmup 20 --e2e-steps 3 --kind central
run --n-ue 1000 --n-bs 50 --envs 1024 --fragment 10 --steps 100 --warmup 20 --e2e-steps 3 --interference

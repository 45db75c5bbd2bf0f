# BASELINE.json configs beyond the headline one, and the env-batch sweep.  usage: bash scripts/gpu_configs.sh <tag>
TAG=${1:-x}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
fmt='import json,sys; d=json.loads(sys.stdin.read()); print("%s: env-steps/s %.4e  us/step %.2f  frac %.3f  e2e %.3e  %s" % (d["config"]["workload"][:44], d["value"], 1e3*d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["run"]["launch_geometry"]))'
run() { name=$1; shift; timeout 600 python bench.py --no-cpu-baseline "$@" > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err; python -c "$fmt" < gpurun_out/bench_${TAG}_$name.json || tail -5 gpurun_out/bench_${TAG}_$name.err; }
run cfg3 --n-ue 200 --n-bs 20 --envs 512 --fragment 50 --steps 1000 --warmup 100 --e2e-steps 10
run cfg4 --n-ue 1000 --n-bs 50 --envs 1024 --fragment 10 --steps 100 --warmup 20 --e2e-steps 3
run cfg4c --n-ue 1000 --n-bs 50 --envs 1024 --fragment 10 --steps 100 --warmup 20 --e2e-steps 3 --kind central
# config 4 with the interference extension (SINR; not in the reference, not parity-graded)
run cfg4i --n-ue 1000 --n-bs 50 --envs 1024 --fragment 10 --steps 100 --warmup 20 --e2e-steps 3 --interference
run cfg3c --n-ue 200 --n-bs 20 --envs 512 --fragment 50 --steps 1000 --warmup 100 --e2e-steps 10 --kind central
run central --kind central --steps 3000 --warmup 300 --e2e-steps 50

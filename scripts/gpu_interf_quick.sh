# Quick check of the interference extension: its parity tests, then BASELINE config 4 with it.  usage: bash scripts/gpu_interf_quick.sh <tag>
TAG=${1:-x}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "interference" > gpurun_out/pytest_$TAG.log 2>&1; tail -5 gpurun_out/pytest_$TAG.log
fmt='import json,sys; d=json.loads(sys.stdin.read()); print("%s: env-steps/s %.4e  us/step %.2f  frac %.3f  e2e %.3e  %s" % (d["config"]["workload"][:44], d["value"], 1e3*d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["run"]["launch_geometry"]))'
timeout 300 python bench.py --no-cpu-baseline --n-ue 1000 --n-bs 50 --envs 1024 --fragment 10 --steps 100 --warmup 20 --reps 3 --e2e-steps 3 --interference > gpurun_out/bench_${TAG}_cfg4i.json 2> gpurun_out/bench_${TAG}_cfg4i.err
python -c "$fmt" < gpurun_out/bench_${TAG}_cfg4i.json || tail -5 gpurun_out/bench_${TAG}_cfg4i.err

# Last call of a round, most important first (the GPU budget may cut the tail): full GPU suite, default bench, the driver's
# invocation, smoke, launch list, full ncu captures of a 100-step and of the 20-step launches.  usage: bash scripts/gpu_last.sh <tag>
TAG=${1:-x}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests -q -m gpu > $O/pytest_$TAG.log 2>&1; tail -2 $O/pytest_$TAG.log
timeout 200 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; tail -c 300 $O/bench_$TAG.json; echo
timeout 100 python bench.py --steps 20 --warmup 5 > $O/bench_${TAG}_driver.json 2> $O/bench_${TAG}_driver.err; tail -c 200 $O/bench_${TAG}_driver.json; echo
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke_$TAG.log 2>&1; tail -1 $O/smoke_$TAG.log
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 25 > $O/ncu_launch_$TAG.log 2>&1
timeout 100 ncu --set full --clock-control none --import-source on -k regex:dcb_step_kernel -s 5 -c 1 -o $O/prof_$TAG -f \
    python bench.py --steps 300 --warmup 100 --reps 1 --no-cpu-baseline --e2e-steps 3 > $O/ncu_full_$TAG.log 2>&1
timeout 100 ncu --set full --clock-control none -k regex:dcb_step_kernel -s 6 -c 8 -o $O/prof_f20_$TAG -f \
    python bench.py --steps 20 --warmup 5 --reps 2 --no-cpu-baseline --e2e-steps 3 > $O/ncu_full_f20_$TAG.log 2>&1
ls $O | grep $TAG

# Round artefacts in one GPU call.  usage: bash scripts/gpu_final.sh <tag>
TAG=${1:-x}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > $O/pytest_$TAG.log 2>&1; tail -2 $O/pytest_$TAG.log
# the default bench, the driver's invocation, the reference arm
timeout 600 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; tail -c 600 $O/bench_$TAG.json; tail -3 $O/bench_$TAG.err
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_${TAG}_driver.json 2> $O/bench_${TAG}_driver.err; tail -c 300 $O/bench_${TAG}_driver.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref_$TAG.json 2>> $O/bench_$TAG.err; tail -c 400 $O/bench_ref_$TAG.json
# launch list of the driver's command (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 25 > $O/ncu_launch_$TAG.log 2>&1
# full captures: one 100-step fragment launch (default bench), one 20-step launch (the driver's fragment length), and the
# wide kernel at config 4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dcb_step_kernel -s 5 -c 1 -o $O/prof_$TAG -f \
    python bench.py --steps 300 --warmup 100 --reps 1 --no-cpu-baseline --e2e-steps 3 > $O/ncu_full_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:dcb_step_kernel -s 6 -c 8 -o $O/prof_f20_$TAG -f \
    python bench.py --steps 20 --warmup 5 --reps 2 --no-cpu-baseline --e2e-steps 3 > $O/ncu_full_f20_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dcb_wide_kernel -s 4 -c 1 -o $O/prof_wide_$TAG -f \
    python bench.py --n-ue 1000 --n-bs 50 --envs 1024 --fragment 4 --steps 12 --warmup 4 --reps 1 --no-cpu-baseline --e2e-steps 1 > $O/ncu_wide_$TAG.log 2>&1
# phase timeline of CTA 0 (instrumented build: scripts/build_trace_lib.sh)
[ -f gpurun_exp_TRACE.so ] && DCB_LIB_PATH=$GRAFT_REPO_ROOT/gpurun_exp_TRACE.so timeout 300 python scripts/trace_timeline.py > $O/timeline_$TAG.txt 2>&1
# racecheck: both kernels, small shapes, multi-step fragments
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_api.py -q -x -k "step_many or auto_reset or rollout_equals" > $O/racecheck_$TAG.log 2>&1; tail -3 $O/racecheck_$TAG.log
DCB_FORCE_WIDE=1 timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_api.py tests/test_gpu_parity.py -q -x -k "step_many or auto_reset or interference" > $O/racecheck_wide_$TAG.log 2>&1; tail -3 $O/racecheck_wide_$TAG.log
bash scripts/gpu_configs.sh $TAG
ls $O | head -80
# build + smoke entry point of the driver
timeout 600 python __graft_entry__.py > $O/graft_entry_$TAG.log 2>&1; tail -2 $O/graft_entry_$TAG.log

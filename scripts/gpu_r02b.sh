# round 2: new bench structure -- driver flags, default flags, e2e; suite on the host-table build
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02b_driver.json 2> gpurun_out/bench_r02b_driver.err; tail -c 3000 gpurun_out/bench_r02b_driver.json; tail -5 gpurun_out/bench_r02b_driver.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err; tail -c 2500 gpurun_out/bench_r02b.json; tail -5 gpurun_out/bench_r02b.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 | tail -c 600
timeout 900 python -m pytest tests/test_gpu_soak.py -q -m gpu -x -s 2>&1 | tail -12

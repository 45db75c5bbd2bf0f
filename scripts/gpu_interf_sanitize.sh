# compute-sanitizer (racecheck, memcheck) over the interference extension's tests.  usage: bash scripts/gpu_interf_sanitize.sh <tag>
TAG=${1:-x}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 170 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -x -k "interference_fragment or (interference_extension and 12-5)" > gpurun_out/racecheck_interf_$TAG.log 2>&1; tail -4 gpurun_out/racecheck_interf_$TAG.log
timeout 120 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k "interference_fragment and central" > gpurun_out/memcheck_interf_$TAG.log 2>&1; tail -4 gpurun_out/memcheck_interf_$TAG.log

# compute-sanitizer memcheck over the newer code paths (wide kernel, variable population, brute force, padding).  usage: bash scripts/gpu_memcheck.sh <tag>
TAG=${1:-x}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -q -x \
    -k "pop_updown or pop_3up2down or (wide_kernel and 5-3) or (wide_kernel and 33-64) or fragments or utilstep_multi_min" > gpurun_out/memcheck_$TAG.log 2>&1
echo "rc=$?"; tail -4 gpurun_out/memcheck_$TAG.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_api.py -q -x \
    -k "brute_force_matches or padding_slots or arriving_and_departing" > gpurun_out/memcheck_api_$TAG.log 2>&1
echo "rc=$?"; tail -4 gpurun_out/memcheck_api_$TAG.log

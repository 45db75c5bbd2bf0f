# compute-sanitizer memcheck over the round-2 code paths (data-rate observations on both kernels, interference, UniformMovement,
# sequential stepping, continuous stepping, host-buffer fragments, variable population through the adapters).
# usage: bash scripts/gpu_memcheck.sh <tag>
TAG=${1:-x}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -q -x \
    -k "datarate or normdr or (interference and 12-5) or (interference and 20-10) or uniform_multi or seq_multi_avg or pop_3up2down or fragments" > gpurun_out/memcheck_$TAG.log 2>&1
echo "rc=$?"; tail -4 gpurun_out/memcheck_$TAG.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_api.py -q -x \
    -k "continuous_stepping or step_many_host or rllib_adapters_with or datarate_facades or step_host_with" > gpurun_out/memcheck_api_$TAG.log 2>&1
echo "rc=$?"; tail -4 gpurun_out/memcheck_api_$TAG.log

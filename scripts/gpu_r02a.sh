# round 2, first call: suite on the split build, FX_AGG experiment (parity subset + A/B bench)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
DCB_LIB_PATH=$GRAFT_REPO_ROOT/gpurun_exp_FXAGG.so timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -3
bash scripts/gpu_exp.sh tree gpurun_exp_FXAGG.so 2>&1 | tee gpurun_out/exp_r02a.txt
